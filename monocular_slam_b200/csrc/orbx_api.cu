// orbx_api.cu -- C ABI of liborbx.so: handle management, level geometry, and the extraction entry points
// declared in include/orbx.h (the drop-in for reference src/FeatureExtractor.cpp:17,19).
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace orbx {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// defined in the kernel files
void pyr_down_smem_extent(const int* ofs_x, int dw, const int* ofs_y, int dh, int sw, int sh, int* s_w, int* s_h);
cudaError_t launch_pyr_down_level(uint8_t* slots, size_t slot_stride, const LevelGeom& src, const LevelGeom& dst, int s_w,
                                  int s_h, int nframes, cudaStream_t s);
void fast_tile_dims(int* tw, int* th);
cudaError_t launch_ingest(const uint8_t* d_frames, size_t frame_pitch, size_t stride, int w, int h, uint8_t* slots,
                          size_t slot_stride, const LevelGeom& L0, int nframes, cudaStream_t s);
cudaError_t launch_ingest_bgr(const uint8_t* d_frames, size_t frame_pitch, size_t stride, int w, int h, uint8_t* slots,
                              size_t slot_stride, const LevelGeom& L0, int nframes, cudaStream_t s);
cudaError_t harris_select_prepare(int max_surv_cap);
cudaError_t launch_harris_select(const FrameGeom& g, const uint8_t* slots, size_t slot_stride, const Cand* surv,
                                 size_t surv_stride, Sel* sel, size_t sel_stride, FrameCounters* ctr, int nframes,
                                 int max_surv_cap, float s4, cudaStream_t s);

}  // namespace orbx

using namespace orbx;

// One of the batches that may be in flight through orbx_submit_batch / orbx_wait_batch.  Lane l owns slots
// [l * max_batch, (l + 1) * max_batch) of every per-frame array of the handle.
struct orbx_lane {
    bool busy;
    int nframes, cap;
    cudaEvent_t uploaded, computed, done;
    int32_t* counts;      // caller's arrays, filled by orbx_wait_batch
    int64_t* ngood;
    int32_t* ninliers;
    int back;             // 0: consecutive pairs (one per frame); >= 1: `back` pairs per frame (orbx_submit_batch_back)
};
#ifndef ORBX_LANES_N
#define ORBX_LANES_N 3
#endif
constexpr int ORBX_LANES = ORBX_LANES_N;
constexpr int ORBX_MAX_BACK = 8;
constexpr int ORBX_SPLIT_MIN = 8;
constexpr int ORBX_MAX_SPLIT = 8;       // a batch is split in two when each half has at least this many frames

// The host reads the per-level counts, the total and the overflow flags of each frame; the score histograms behind them
// (8 KB per frame, K3's scratch) stay on the device: a strided copy of the leading bytes of every FrameCounters.
static cudaError_t copy_counters_d2h(FrameCounters* dst, const FrameCounters* src, int nframes, cudaStream_t s) {
    return cudaMemcpy2DAsync(dst, sizeof(FrameCounters), src, sizeof(FrameCounters), offsetof(FrameCounters, hist), (size_t)nframes,
                             cudaMemcpyDeviceToHost, s);
}

struct orbx_context {
    int device;
    orbx_params p;
    int max_w, max_h, max_batch;
    cudaStream_t own_stream, stream;
    cudaStream_t copy_stream;                 // host-buffer path: uploads of chunk i+1 overlap the kernels of chunk i
    std::vector<cudaEvent_t>* copy_events;
    cudaEvent_t order_event;                  // makes the copy stream wait for work already queued on the compute stream
    // geometry: `gmax` sizes the allocations, `g` is the geometry of the frame size last used
    FrameGeom gmax, g;
    int geom_w, geom_h;
    int pyr_sw[ORBX_MAX_LEVELS], pyr_sh[ORBX_MAX_LEVELS];
    size_t slot_stride, cand_stride, surv_stride, sel_stride;
    int max_surv_cap, dev_cap;
    float harris_s4;
    // device memory
    uint8_t* d_slots;
    Cand* d_cand;
    Cand* d_surv;
    Sel* d_sel;
    FrameCounters* d_ctr;
    orbx_keypoint* d_kps;
    uint8_t* d_desc;
    int32_t* d_counts;
    uint8_t* d_tab;
    size_t tab_bytes;
    int32_t* d_fast_hint;                     // per level: corner-dense in the previous batch (K3 -> K2 speed hint)
    // pinned host memory
    FrameCounters* h_ctr;
    int32_t* h_counts;
    std::vector<uint8_t>* h_tab;
    int dev_pending;    // > 0: _dev submissions of up to this many frames have not been checked for overflow yet
    bool dev_unordered; // such a submission may still be running and the copy stream has not been ordered behind it yet
    // sequence mode (orbx_match_consecutive)
    int last_nframes, last_cap;
    uint8_t* d_prev_desc;
    int32_t* d_prev_count;
    bool have_prev;
    int prev_cap;                             // row capacity the kept previous frame was extracted with
    orbx_dmatch* d_good;
    int64_t* d_ngood;
    int64_t* h_ngood;
    // orbx_filter_consecutive: keypoints of the last frame of the previous batch (double-buffered, allocated on first match)
    orbx_keypoint* d_prev_kps[2];
    int prev_kps_cur;
    const orbx_keypoint* filter_prev_kps;     // what frame 0 of the batch last matched pairs with (NULL: nothing)
    int filter_nframes, filter_cap;           // batch orbx_match_consecutive last ran on (0: none)
    uint8_t* d_fstatus; double* d_fF; int32_t* d_finfo; int32_t* h_finfo;
    // orbx_match_back: history of the last ORBX_MAX_BACK frames (double-buffered), allocated on first use
    uint8_t* d_hist[2];
    int32_t* d_hist_counts[2];
    int hist_cur, nhist, hist_cap;
    orbx_dmatch* d_good_back;
    int64_t* d_ngood_back;
    int64_t* h_ngood_back;
    // orbx_filter_back: keypoint history parallel to d_hist, and the geometry of the batch orbx_match_back last ran on
    orbx_keypoint* d_hist_kps[2];
    int back_n, back_back, back_cap, back_nhist;
    uint8_t* d_fstatus_back; double* d_fF_back; int32_t* d_finfo_back; int32_t* h_finfo_back;
    // single-frame path: the launch sequence of run_extract_on(frame 0) captured as CUDA graphs, one per (mode, capacity);
    // dropped whenever the geometry changes (the resize tables' addresses are baked into the kernel arguments)
    struct { cudaGraphExec_t exec; int mode, cap; } graphs[4];
    int ngraphs;
    bool use_graphs;
    // two-way split of large batches (run_extract)
    int split;                                // number of parts (1 = off)
    cudaStream_t sub_stream[ORBX_MAX_SPLIT];
    cudaEvent_t fork_event, join_event[ORBX_MAX_SPLIT];
    // 3-channel input (orbx_set_input_channels): packed BGR staging for the host paths, allocated on first use
    int channels;
    uint8_t* d_bgr;
    // pipelined host path
    orbx_lane lanes[ORBX_LANES];
    int lane_head, lane_next;                 // oldest batch in flight, lane the next submission uses
    cudaStream_t d2h_stream;
    // stage profiling
    bool profiling;
    std::vector<cudaEvent_t>* events;
    size_t events_used;
};

static inline int rne_f(float v) { return (int)lrintf(v); }
static int require_idle(orbx_handle h, const char* fn);

// The sequence entry points run the caller's matcher / filter on the extractor's stream (ordered after the extraction,
// no extra synchronisation) for the duration of one call; the stream the caller installed on that handle is put back
// when the scope ends, so later hamx_*_dev / fmx_*_dev calls stay ordered on the caller's stream.
struct ScopedHamxStream {
    hamx_handle m; void* saved; int rc;
    ScopedHamxStream(hamx_handle m_, cudaStream_t s) : m(m_), saved(nullptr), rc(hamx_get_stream(m_, &saved)) { if (!rc) rc = hamx_set_stream(m, (void*)s); }
    ~ScopedHamxStream() { if (m) hamx_set_stream(m, saved); }
};
struct ScopedFmxStream {
    fmx_handle m; void* saved; int rc;
    ScopedFmxStream(fmx_handle m_, cudaStream_t s) : m(m_), saved(nullptr), rc(fmx_get_stream(m_, &saved)) { if (!rc) rc = fmx_set_stream(m, (void*)s); }
    ~ScopedFmxStream() { if (m) fmx_set_stream(m, saved); }
};

// Level sizes, quotas, list capacities and buffer offsets for a w x h frame.
static int build_geometry(const orbx_params& p, int w, int h, FrameGeom* g)
{
    memset(g, 0, sizeof(*g));
    g->nlevels = p.nlevels;
    g->w = w;
    g->h = h;
    g->score_type = p.score_type;
    g->fast_threshold = p.fast_threshold;
    int tw, th;
    fast_tile_dims(&tw, &th);

    int quota[ORBX_MAX_LEVELS];
    {
        float factor = (float)(1.0 / (double)p.scale_factor);
        float nd = (float)p.nfeatures * (1.f - factor) / (1.f - (float)pow((double)factor, (double)p.nlevels));
        int sum = 0;
        for (int l = 0; l < p.nlevels - 1; l++) {
            quota[l] = rne_f(nd);
            sum += quota[l];
            nd *= factor;
        }
        quota[p.nlevels - 1] = std::max(p.nfeatures - sum, 0);
    }

    size_t img_off = 0, cand_off = 0, surv_off = 0, sel_off = 0;
    int tile_start = 0;
    for (int l = 0; l < p.nlevels; l++) {
        LevelGeom& L = g->lv[l];
        L.scale = (float)pow((double)p.scale_factor, (double)(l - p.first_level));
        L.inv_scale = 1.f / L.scale;
        L.w = rne_f((float)w / L.scale);
        L.h = rne_f((float)h / L.scale);
        if (L.w < 1 || L.h < 1) { set_error("pyramid level %d of a %dx%d frame is empty", l, w, h); return ORBX_E_INVALID; }
        L.pitch = (int)align_up((size_t)L.w, 128);
        L.quota = quota[l];
        const int iw = std::max(L.w - 62, 0), ih = std::max(L.h - 62, 0);   // interior [31, w-31) x [31, h-31)
        L.tiles_x = div_up(iw, tw);
        L.tiles_y = div_up(ih, th);
        if (iw == 0 || ih == 0) L.tiles_x = L.tiles_y = 0;
        L.tile_start = tile_start;
        tile_start += L.tiles_x * L.tiles_y;
        L.cand_cap = (iw == 0 || ih == 0) ? 0 : (iw / 2 + 1) * (ih / 2 + 1);   // NMS: at most one maximum per 2x2
        const int target = p.score_type == ORBX_HARRIS_SCORE ? 2 * L.quota : L.quota;
        L.surv_cap = std::min(L.cand_cap, target + std::max(target, 4096));     // room for ties at the cut
        L.img_off = img_off;
        img_off += align_up((size_t)L.pitch * L.h, 256);
        L.cand_off = cand_off; cand_off += (size_t)L.cand_cap;
        L.surv_off = surv_off; surv_off += (size_t)L.surv_cap;
        L.sel_off = sel_off;   sel_off += (size_t)L.surv_cap;
    }
    g->total_tiles = tile_start;
    return ORBX_OK;
}

static void geometry_totals(const FrameGeom& g, size_t* img, size_t* cand, size_t* surv, int* max_surv)
{
    const LevelGeom& L = g.lv[g.nlevels - 1];
    *img = L.img_off + align_up((size_t)L.pitch * L.h, 256);
    *cand = L.cand_off + L.cand_cap;
    *surv = L.surv_off + L.surv_cap;
    int m = 0;
    for (int l = 0; l < g.nlevels; l++) m = std::max(m, g.lv[l].surv_cap);
    *max_surv = m;
}

// INTER_LINEAR_EXACT coefficient tables, computed in double exactly like OpenCV.
static void linear_table(int n, int m, int* ofs, uint16_t* c1)
{
    const double s = 1.0 / ((double)m / (double)n);
    for (int d = 0; d < m; d++) {
        double f = s * (d + 0.5) - 0.5;
        int i = (int)floor(f);
        if (i < 0) { ofs[d] = 0; c1[d] = 0; }
        else if (i >= n - 1) { ofs[d] = n - 1; c1[d] = 0; }
        else { ofs[d] = i; c1[d] = (uint16_t)lrint((f - i) * 256.0); }
    }
}

static size_t table_bytes(const FrameGeom& g)
{
    size_t b = 0;
    for (int l = 1; l < g.nlevels; l++) b += align_up((size_t)(g.lv[l].w + g.lv[l].h) * 4, 16) + align_up((size_t)(g.lv[l].w + g.lv[l].h) * 2, 16);
    return b + 256;
}

static void drop_graphs(orbx_handle h)
{
    for (int i = 0; i < h->ngraphs; i++) cudaGraphExecDestroy(h->graphs[i].exec);
    h->ngraphs = 0;
}

static int set_geometry(orbx_handle h, int w, int hh)
{
    if (h->geom_w == w && h->geom_h == hh) return ORBX_OK;
    drop_graphs(h);
    ORBX_REQUIRE(w >= 1 && hh >= 1 && w <= h->max_w && hh <= h->max_h, "frame %dx%d outside the handle's limits %dx%d", w, hh,
                 h->max_w, h->max_h);
    FrameGeom g;
    int rc = build_geometry(h->p, w, hh, &g);
    if (rc) return rc;
    std::vector<uint8_t>& tab = *h->h_tab;
    size_t off = 0;
    for (int l = 1; l < g.nlevels; l++) {
        LevelGeom& D = g.lv[l];
        const LevelGeom& S = g.lv[l - 1];
        int* ox = (int*)(tab.data() + off);
        int* oy = ox + D.w;
        size_t ioff = off;
        off += align_up((size_t)(D.w + D.h) * 4, 16);
        uint16_t* cx = (uint16_t*)(tab.data() + off);
        uint16_t* cy = cx + D.w;
        size_t coff = off;
        off += align_up((size_t)(D.w + D.h) * 2, 16);
        linear_table(S.w, D.w, ox, cx);
        linear_table(S.h, D.h, oy, cy);
        pyr_down_smem_extent(ox, D.w, oy, D.h, S.w, S.h, &h->pyr_sw[l], &h->pyr_sh[l]);
        ORBX_REQUIRE((size_t)h->pyr_sw[l] * 32 * 2 + 16 <= 200 * 1024,
                     "scale factor %.3f needs more shared memory than one CTA has", (double)h->p.scale_factor);
        D.ofs_x = (const int*)(h->d_tab + ioff);
        D.ofs_y = D.ofs_x + D.w;
        D.c1x = (const uint16_t*)(h->d_tab + coff);
        D.c1y = D.c1x + D.w;
    }
    ORBX_REQUIRE(off <= h->tab_bytes, "internal: resize tables (%zu bytes) exceed their buffer (%zu)", off, h->tab_bytes);
    // the previous tables may still be in use by work queued on the stream; the copy below is stream-ordered, but the
    // pageable source is overwritten on the host, so drain first (geometry changes are rare: once per frame size)
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    if (off) ORBX_CUDA(cudaMemcpyAsync(h->d_tab, tab.data(), off, cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    h->g = g;
    h->geom_w = w;
    h->geom_h = hh;
    return ORBX_OK;
}

extern "C" const char* orbx_last_error(void) { return g_error; }
extern "C" const char* orbx_version(void) { return "orbx 0.1 (sm_100a)"; }

extern "C" int orbx_host_alloc(size_t bytes, void** out)
{
    ORBX_REQUIRE(out != nullptr && bytes > 0, "orbx_host_alloc: bad arguments");
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) { set_error("orbx_host_alloc: cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); *out = nullptr; return ORBX_E_ALLOC; }
    return ORBX_OK;
}

extern "C" int orbx_host_alloc_wc(size_t bytes, void** out)
{
    ORBX_REQUIRE(out != nullptr && bytes > 0, "orbx_host_alloc_wc: bad arguments");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocWriteCombined);
    if (e != cudaSuccess) { set_error("orbx_host_alloc_wc: cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); *out = nullptr; return ORBX_E_ALLOC; }
    return ORBX_OK;
}

extern "C" int orbx_host_free(void* p)
{
    if (p) ORBX_CUDA(cudaFreeHost(p));
    return ORBX_OK;
}

extern "C" int orbx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" void orbx_default_params(orbx_params* p)
{
    if (!p) return;
    p->nfeatures = 500; p->scale_factor = 1.2f; p->nlevels = 8; p->edge_threshold = 31; p->first_level = 0; p->wta_k = 2;
    p->score_type = ORBX_HARRIS_SCORE; p->patch_size = 31; p->fast_threshold = 20;
}

extern "C" int orbx_create(orbx_handle* out, const orbx_params* params, int device, int max_w, int max_h, int max_batch)
{
    ORBX_REQUIRE(out != nullptr, "orbx_create: out is NULL");
    *out = nullptr;
    orbx_params p;
    orbx_default_params(&p);
    if (params) p = *params;
    ORBX_REQUIRE(p.nlevels >= 1 && p.nlevels <= ORBX_MAX_LEVELS, "nlevels %d outside [1,%d]", p.nlevels, ORBX_MAX_LEVELS);
    ORBX_REQUIRE(p.scale_factor > 1.0f && p.scale_factor <= 2.0f, "scale_factor %.3f outside (1,2]", (double)p.scale_factor);
    ORBX_REQUIRE(p.nfeatures >= 0 && p.nfeatures <= 20000, "nfeatures %d outside [0,20000]", p.nfeatures);
    ORBX_REQUIRE(p.edge_threshold == 31 && p.first_level == 0 && p.wta_k == 2 && p.patch_size == 31,
                 "only edge_threshold 31, first_level 0, wta_k 2, patch_size 31 (the reference's defaults) are supported");
    ORBX_REQUIRE(p.score_type == ORBX_HARRIS_SCORE || p.score_type == ORBX_FAST_SCORE, "bad score_type %d", p.score_type);
    ORBX_REQUIRE(p.fast_threshold >= 0 && p.fast_threshold <= 254, "fast_threshold %d outside [0,254]", p.fast_threshold);
    ORBX_REQUIRE(max_w >= 1 && max_h >= 1 && max_w <= 16384 && max_h <= 16384, "max frame size %dx%d outside [1,16384]", max_w, max_h);
    ORBX_REQUIRE(max_batch >= 1 && max_batch <= 1024, "max_batch %d outside [1,1024]", max_batch);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("orbx_create: no CUDA device (%s); liborbx has no CPU fallback", cudaGetErrorString(e));
        return ORBX_E_CUDA;
    }
    ORBX_REQUIRE(device >= 0 && device < ndev, "orbx_create: device %d out of range [0,%d)", device, ndev);
    ORBX_CUDA(cudaSetDevice(device));

    orbx_context* h = new orbx_context();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->p = p;
    h->max_w = max_w; h->max_h = max_h; h->max_batch = max_batch;
    h->geom_w = h->geom_h = -1;
    h->channels = 1;
    {
        float scale = 1.f / ((1 << 2) * 7 * 255.f);    // OpenCV HarrisResponses, blockSize 7
        h->harris_s4 = scale * scale * scale * scale;
    }
    int rc = build_geometry(p, max_w, max_h, &h->gmax);
    if (rc) { delete h; return rc; }
    size_t img, cand, surv;
    geometry_totals(h->gmax, &img, &cand, &surv, &h->max_surv_cap);
    if (harris_select_smem(h->max_surv_cap) > 220 * 1024) {
        set_error("nfeatures %d needs %zu bytes of shared memory in the Harris selection (limit 220 KB)", p.nfeatures,
                  harris_select_smem(h->max_surv_cap));
        delete h;
        return ORBX_E_INVALID;
    }
    h->slot_stride = align_up(img + 256, 256);
    h->cand_stride = cand + 1;
    h->surv_stride = surv + 1;
    h->sel_stride = surv + 1;
    h->dev_cap = (int)std::min<size_t>(surv, 1u << 20);
    h->tab_bytes = table_bytes(h->gmax) + 4096;
    h->h_tab = new std::vector<uint8_t>(h->tab_bytes);

    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), orbx_destroy(h));
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking), orbx_destroy(h));
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking), orbx_destroy(h));
    {
        const char* e = getenv("ORBX_SPLIT");
        h->split = e ? atoi(e) : 1;
        if (h->split < 1) h->split = 1;
        if (h->split > ORBX_MAX_SPLIT) h->split = ORBX_MAX_SPLIT;
        e = getenv("ORBX_GRAPHS");
        h->use_graphs = !(e && e[0] == '0');
        for (int i = 0; i < ORBX_MAX_SPLIT; i++) {
            ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->sub_stream[i], cudaStreamNonBlocking), orbx_destroy(h));
            ORBX_CUDA_OR(cudaEventCreateWithFlags(&h->join_event[i], cudaEventDisableTiming), orbx_destroy(h));
        }
        ORBX_CUDA_OR(cudaEventCreateWithFlags(&h->fork_event, cudaEventDisableTiming), orbx_destroy(h));
    }
    for (int l = 0; l < ORBX_LANES; l++) {
        ORBX_CUDA_OR(cudaEventCreateWithFlags(&h->lanes[l].uploaded, cudaEventDisableTiming), orbx_destroy(h));
        ORBX_CUDA_OR(cudaEventCreateWithFlags(&h->lanes[l].computed, cudaEventDisableTiming), orbx_destroy(h));
        ORBX_CUDA_OR(cudaEventCreateWithFlags(&h->lanes[l].done, cudaEventDisableTiming), orbx_destroy(h));
    }
    h->copy_events = new std::vector<cudaEvent_t>();
    ORBX_CUDA_OR(cudaEventCreateWithFlags(&h->order_event, cudaEventDisableTiming), orbx_destroy(h));
    h->stream = h->own_stream;
    const size_t B = (size_t)max_batch * ORBX_LANES;   // blocking calls use lane 0 only
#define ORBX_ALLOC(ptr, bytes)                                                                            \
    do {                                                                                                  \
        cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                                             \
        if (e_ != cudaSuccess) {                                                                          \
            set_error("orbx_create: cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e_)); \
            orbx_destroy(h);                                                                              \
            return ORBX_E_ALLOC;                                                                          \
        }                                                                                                 \
    } while (0)
    ORBX_ALLOC(h->d_slots, B * h->slot_stride);
    ORBX_ALLOC(h->d_cand, B * h->cand_stride * sizeof(Cand));
    ORBX_ALLOC(h->d_surv, B * h->surv_stride * sizeof(Cand));
    ORBX_ALLOC(h->d_sel, B * h->sel_stride * sizeof(Sel));
    ORBX_ALLOC(h->d_ctr, B * sizeof(FrameCounters));
    ORBX_ALLOC(h->d_kps, B * (size_t)h->dev_cap * sizeof(orbx_keypoint) + 256);
    ORBX_ALLOC(h->d_desc, B * (size_t)h->dev_cap * 32 + 256);
    ORBX_ALLOC(h->d_counts, B * sizeof(int32_t) + 256);
    ORBX_ALLOC(h->d_tab, h->tab_bytes);
    ORBX_ALLOC(h->d_fast_hint, ORBX_MAX_LEVELS * sizeof(int32_t));
    ORBX_ALLOC(h->d_prev_desc, (size_t)h->dev_cap * 32 + 256);
    ORBX_ALLOC(h->d_prev_count, 256);
    ORBX_ALLOC(h->d_good, B * (size_t)h->dev_cap * sizeof(orbx_dmatch) + 256);
    ORBX_ALLOC(h->d_ngood, B * sizeof(int64_t) + 256);
#undef ORBX_ALLOC
    ORBX_CUDA_OR(cudaMallocHost((void**)&h->h_ngood, B * sizeof(int64_t)), orbx_destroy(h));
    h->events = new std::vector<cudaEvent_t>();
    ORBX_CUDA_OR(cudaMemset(h->d_slots, 0, B * h->slot_stride), orbx_destroy(h));
    ORBX_CUDA_OR(cudaMemset(h->d_fast_hint, 0, ORBX_MAX_LEVELS * sizeof(int32_t)), orbx_destroy(h));   // padding bytes are read (never used) by vector loads
    ORBX_CUDA_OR(cudaMallocHost((void**)&h->h_ctr, B * sizeof(FrameCounters)), orbx_destroy(h));
    ORBX_CUDA_OR(cudaMallocHost((void**)&h->h_counts, B * sizeof(int32_t)), orbx_destroy(h));
    ORBX_CUDA_OR(harris_select_prepare(h->max_surv_cap), orbx_destroy(h));
    *out = h;
    return ORBX_OK;
}

extern "C" int orbx_destroy(orbx_handle h)
{
    if (!h) return ORBX_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graphs(h);
    cudaFree(h->d_slots); cudaFree(h->d_cand); cudaFree(h->d_surv); cudaFree(h->d_sel); cudaFree(h->d_ctr);
    cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_counts); cudaFree(h->d_tab); cudaFree(h->d_fast_hint);
    cudaFree(h->d_prev_desc); cudaFree(h->d_prev_count); cudaFree(h->d_good); cudaFree(h->d_ngood); cudaFree(h->d_bgr);
    cudaFree(h->d_prev_kps[0]); cudaFree(h->d_prev_kps[1]); cudaFree(h->d_fstatus); cudaFree(h->d_fF); cudaFree(h->d_finfo);
    if (h->h_finfo) cudaFreeHost(h->h_finfo);
    cudaFree(h->d_hist[0]); cudaFree(h->d_hist[1]); cudaFree(h->d_hist_counts[0]); cudaFree(h->d_hist_counts[1]);
    cudaFree(h->d_good_back); cudaFree(h->d_ngood_back);
    cudaFree(h->d_hist_kps[0]); cudaFree(h->d_hist_kps[1]); cudaFree(h->d_fstatus_back); cudaFree(h->d_fF_back); cudaFree(h->d_finfo_back);
    if (h->h_finfo_back) cudaFreeHost(h->h_finfo_back);
    if (h->h_ngood_back) cudaFreeHost(h->h_ngood_back);
    if (h->h_ngood) cudaFreeHost(h->h_ngood);
    if (h->events) { for (cudaEvent_t e : *h->events) cudaEventDestroy(e); delete h->events; }
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) { cudaStreamSynchronize(h->d2h_stream); cudaStreamDestroy(h->d2h_stream); }
    for (int i = 0; i < ORBX_MAX_SPLIT; i++) {
        if (h->sub_stream[i]) { cudaStreamSynchronize(h->sub_stream[i]); cudaStreamDestroy(h->sub_stream[i]); }
        if (h->join_event[i]) cudaEventDestroy(h->join_event[i]);
    }
    if (h->fork_event) cudaEventDestroy(h->fork_event);
    for (int l = 0; l < ORBX_LANES; l++) {
        if (h->lanes[l].uploaded) cudaEventDestroy(h->lanes[l].uploaded);
        if (h->lanes[l].computed) cudaEventDestroy(h->lanes[l].computed);
        if (h->lanes[l].done) cudaEventDestroy(h->lanes[l].done);
    }
    if (h->order_event) cudaEventDestroy(h->order_event);
    if (h->copy_events) { for (cudaEvent_t e : *h->copy_events) cudaEventDestroy(e); delete h->copy_events; }
    delete h->h_tab;
    delete h;
    return ORBX_OK;
}

extern "C" int orbx_set_stream(orbx_handle h, void* cuda_stream)
{
    ORBX_REQUIRE(h != nullptr, "orbx_set_stream: NULL handle");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return ORBX_OK;
}

extern "C" int orbx_synchronize(orbx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "orbx_synchronize: NULL handle");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int orbx_set_input_channels(orbx_handle h, int channels)
{
    ORBX_REQUIRE(h != nullptr, "orbx_set_input_channels: NULL handle");
    ORBX_REQUIRE(channels == 1 || channels == 3, "orbx_set_input_channels: %d channels (only CV_8UC1 and CV_8UC3 BGR are supported)", channels);
    int rc = require_idle(h, "orbx_set_input_channels");
    if (rc) return rc;
    h->channels = channels;
    return ORBX_OK;
}

extern "C" int orbx_max_keypoints(orbx_handle h) { return h ? h->dev_cap : ORBX_E_INVALID; }

extern "C" int orbx_level_info(orbx_handle h, int w, int hh, int32_t* widths, int32_t* heights, float* scales, int32_t* quotas)
{
    ORBX_REQUIRE(h != nullptr, "orbx_level_info: NULL handle");
    FrameGeom g;
    int rc = build_geometry(h->p, w, hh, &g);
    if (rc) return rc;
    for (int l = 0; l < g.nlevels; l++) {
        if (widths) widths[l] = g.lv[l].w;
        if (heights) heights[l] = g.lv[l].h;
        if (scales) scales[l] = g.lv[l].scale;
        if (quotas) quotas[l] = g.lv[l].quota;
    }
    return ORBX_OK;
}

// ---------------------------------------------------------------------------------------------- pipeline
static int build_pyramids(orbx_handle h, int f0, int nframes)
{
    for (int l = 1; l < h->g.nlevels; l++)
        ORBX_CUDA(launch_pyr_down_level(h->d_slots + (size_t)f0 * h->slot_stride, h->slot_stride, h->g.lv[l - 1], h->g.lv[l],
                                        h->pyr_sw[l], h->pyr_sh[l], nframes, h->stream));
    return ORBX_OK;
}

// frames[f0 .. f0 + nframes) -> slots slot0 + f0 ..  (3-channel frames go through the packed BGR staging buffer and
// the conversion kernel, queued on the same stream as the copies)
static int upload_frames(orbx_handle h, const uint8_t* const* frames, int f0, int nframes, int w, int hh, size_t stride,
                         cudaMemcpyKind kind, cudaStream_t s, int slot0 = 0)
{
    if (h->channels == 1) {
        // frames that are equally spaced in host memory with tightly packed rows (a video buffer) go in ONE 2-D copy whose
        // "rows" are whole frames: 64 separate copies cost ~0.4 ms of per-copy setup on the PCIe path
        bool uniform = nframes > 1 && stride == (size_t)w && h->g.lv[0].pitch == w && frames[f0 + 1] > frames[f0];
        const size_t gap = uniform ? (size_t)(frames[f0 + 1] - frames[f0]) : 0;
        for (int f = f0 + 1; uniform && f < f0 + nframes; f++) uniform = frames[f] == frames[f - 1] + gap;
        if (uniform && gap >= (size_t)w * hh) {
            ORBX_CUDA(cudaMemcpy2DAsync(h->d_slots + (size_t)(slot0 + f0) * h->slot_stride + h->g.lv[0].img_off, h->slot_stride, frames[f0], gap,
                                        (size_t)w * hh, (size_t)nframes, kind, s));
            return ORBX_OK;
        }
        for (int f = f0; f < f0 + nframes; f++)
            ORBX_CUDA(cudaMemcpy2DAsync(h->d_slots + (size_t)(slot0 + f) * h->slot_stride + h->g.lv[0].img_off, h->g.lv[0].pitch, frames[f],
                                        stride, (size_t)w, (size_t)hh, kind, s));
        return ORBX_OK;
    }
    ORBX_REQUIRE(stride >= (size_t)w * 3, "3-channel frame: stride %zu is smaller than 3 * width %d", stride, w);
    const size_t frame_bytes = align_up((size_t)h->max_w * 3, 16) * (size_t)h->max_h;
    if (!h->d_bgr) {
        cudaError_t e = cudaMalloc((void**)&h->d_bgr, frame_bytes * (size_t)h->max_batch * ORBX_LANES);
        if (e != cudaSuccess) { h->d_bgr = nullptr; set_error("cudaMalloc of the BGR staging buffer failed: %s", cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    }
    const size_t row = align_up((size_t)w * 3, 16);
    for (int f = f0; f < f0 + nframes; f++)
        ORBX_CUDA(cudaMemcpy2DAsync(h->d_bgr + (size_t)(slot0 + f) * frame_bytes, row, frames[f], stride, (size_t)w * 3, (size_t)hh, kind, s));
    ORBX_CUDA(launch_ingest_bgr(h->d_bgr + (size_t)(slot0 + f0) * frame_bytes, frame_bytes, row, w, hh,
                                h->d_slots + (size_t)(slot0 + f0) * h->slot_stride, h->slot_stride, h->g.lv[0], nframes, s));
    return ORBX_OK;
}

static int require_idle(orbx_handle h, const char* fn)
{
    for (int l = 0; l < ORBX_LANES; l++)
        ORBX_REQUIRE(!h->lanes[l].busy, "%s: a batch submitted with orbx_submit_batch is still in flight; call orbx_wait_batch first", fn);
    return ORBX_OK;
}

// pyramids + FAST + selection + orientation (+ descriptors); results land in d_out / d_desc / d_counts
static int stage_mark(orbx_handle h)
{
    if (!h->profiling) return ORBX_OK;
    if (h->events_used == h->events->size()) {
        cudaEvent_t e;
        ORBX_CUDA(cudaEventCreate(&e));
        h->events->push_back(e);
    }
    ORBX_CUDA(cudaEventRecord((*h->events)[h->events_used++], h->stream));
    return ORBX_OK;
}

// frames [f0, f0 + nframes) of the handle's slots on stream s; their results go to rows [o0, o0 + nframes) of d_out
// ([.][cap]) / d_desc ([.][cap][32]) / d_counts
static int run_extract_on(orbx_handle h, int f0, int nframes, int mode, orbx_keypoint* d_out, uint8_t* d_desc, int cap, int32_t* d_counts,
                          cudaStream_t s, bool mark, int o0)
{
    uint8_t* slots = h->d_slots + (size_t)f0 * h->slot_stride;
    Cand* cand = h->d_cand + (size_t)f0 * h->cand_stride;
    Cand* surv = h->d_surv + (size_t)f0 * h->surv_stride;
    Sel* sel = h->d_sel + (size_t)f0 * h->sel_stride;
    FrameCounters* ctr = h->d_ctr + f0;
    int rc = ORBX_OK;
    ORBX_CUDA(cudaMemsetAsync(ctr, 0, (size_t)nframes * sizeof(FrameCounters), s));
    if (mark && (rc = stage_mark(h))) return rc;
    for (int l = 1; l < h->g.nlevels; l++)
        ORBX_CUDA(launch_pyr_down_level(slots, h->slot_stride, h->g.lv[l - 1], h->g.lv[l], h->pyr_sw[l], h->pyr_sh[l], nframes, s));
    if (mark && (rc = stage_mark(h))) return rc;
    ORBX_CUDA(launch_fast(h->g, slots, h->slot_stride, cand, h->cand_stride, ctr, nframes, h->d_fast_hint, s));
    if (mark && (rc = stage_mark(h))) return rc;
    ORBX_CUDA(launch_select(h->g, cand, h->cand_stride, surv, h->surv_stride, ctr, nframes, h->d_fast_hint, s));
    if (mark && (rc = stage_mark(h))) return rc;
    ORBX_CUDA(launch_harris_select(h->g, slots, h->slot_stride, surv, h->surv_stride, sel, h->sel_stride, ctr, nframes,
                                   h->max_surv_cap, h->harris_s4, s));
    if (mark && (rc = stage_mark(h))) return rc;
    ORBX_CUDA(launch_orient_describe(h->g, slots, h->slot_stride, sel, h->sel_stride, ctr, d_out + (size_t)o0 * cap,
                                     (mode & ORBX_DO_DESC) ? d_desc + (size_t)o0 * cap * 32 : nullptr, cap, d_counts + o0, nframes,
                                     mode, s));
    if (mark && (rc = stage_mark(h))) return rc;
    return ORBX_OK;
}

// The stages of one frame depend on each other, the frames of a batch do not.  A large batch can therefore be cut in
// `split` parts (ORBX_SPLIT=n) that run on as many internal streams (forked from / joined to the handle's stream with
// events), so that the short, latency-bound launches of one part overlap the bulk kernels of another.  That paid 3 % while
// the Harris selection and the score cut were slow; with the current kernels one stream is fastest (measured: 24.6 k
// frames/s unsplit vs 24.1-24.5 k with 2-8 parts), so the default is 1 = off.  Stage profiling always keeps one stream.
static int run_extract(orbx_handle h, int f0, int nframes, int mode, orbx_keypoint* d_out, uint8_t* d_desc, int cap, int32_t* d_counts, int o0)
{
    int parts = std::min(h->split, nframes / ORBX_SPLIT_MIN);
    if (h->profiling || parts < 2)
        return run_extract_on(h, f0, nframes, mode, d_out, d_desc, cap, d_counts, h->stream, h->profiling, o0);
    ORBX_CUDA(cudaEventRecord(h->fork_event, h->stream));
    int b = 0;
    for (int i = 0; i < parts; i++) {
        const int n = (nframes - b) / (parts - i);
        ORBX_CUDA(cudaStreamWaitEvent(h->sub_stream[i], h->fork_event, 0));
        int rc = run_extract_on(h, f0 + b, n, mode, d_out, d_desc, cap, d_counts, h->sub_stream[i], false, o0 + b);
        if (rc) return rc;
        ORBX_CUDA(cudaEventRecord(h->join_event[i], h->sub_stream[i]));
        ORBX_CUDA(cudaStreamWaitEvent(h->stream, h->join_event[i], 0));
        b += n;
    }
    return ORBX_OK;
}

extern "C" int orbx_set_profiling(orbx_handle h, int enabled)
{
    ORBX_REQUIRE(h != nullptr, "orbx_set_profiling: NULL handle");
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    h->profiling = enabled != 0;
    h->events_used = 0;
    return ORBX_OK;
}

extern "C" int orbx_read_profile(orbx_handle h, float* stage_ms, int* nbatches)
{
    ORBX_REQUIRE(h != nullptr && stage_ms != nullptr, "orbx_read_profile: NULL argument");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    const size_t per = ORBX_NSTAGES + 1;
    const size_t n = h->events_used / per;
    for (int s = 0; s < ORBX_NSTAGES; s++) stage_ms[s] = 0.f;
    for (size_t b = 0; b < n; b++)
        for (int s = 0; s < ORBX_NSTAGES; s++) {
            float ms = 0.f;
            ORBX_CUDA(cudaEventElapsedTime(&ms, (*h->events)[b * per + s], (*h->events)[b * per + s + 1]));
            stage_ms[s] += ms;
        }
    if (n) for (int s = 0; s < ORBX_NSTAGES; s++) stage_ms[s] /= (float)n;
    if (nbatches) *nbatches = (int)n;
    h->events_used = 0;
    return ORBX_OK;
}

static int check_counters(orbx_handle h, int nframes, int cap, int slot0 = 0)
{
    for (int f = 0; f < nframes; f++) {
        const int ov = h->h_ctr[slot0 + f].overflow;
        if (ov & 1) { set_error("frame %d: FAST candidate list overflow (internal capacity)", f); return ORBX_E_CAPACITY; }
        if (ov & 2) { set_error("frame %d: more than %d tied keypoints at a retention cut; raise nfeatures", f, h->max_surv_cap); return ORBX_E_CAPACITY; }
        if (ov & 4) { set_error("frame %d: %d keypoints exceed the output capacity %d", f, h->h_ctr[slot0 + f].total, cap); return ORBX_E_CAPACITY; }
    }
    return ORBX_OK;
}

static int common_checks(orbx_handle h, const void* img, int w, int hh, size_t stride, const char* fn)
{
    ORBX_REQUIRE(h != nullptr, "%s: NULL handle", fn);
    ORBX_REQUIRE(img != nullptr, "%s: NULL image", fn);
    ORBX_REQUIRE(w >= 1 && hh >= 1 && stride >= (size_t)w * h->channels, "%s: bad image geometry %dx%d (%d channel(s)) stride %zu", fn, w,
                 hh, h->channels, stride);
    int rc = require_idle(h, fn);
    if (rc) return rc;
    ORBX_CUDA(cudaSetDevice(h->device));
    return set_geometry(h, w, hh);
}

// frame slot 0 through all stages: replay of a captured graph when possible, plain launches otherwise
static int run_extract_single(orbx_handle h, int mode, int dcap)
{
    if (!h->use_graphs || h->profiling)
        return run_extract_on(h, 0, 1, mode, h->d_kps, h->d_desc, dcap, h->d_counts, h->stream, h->profiling, 0);
    for (int i = 0; i < h->ngraphs; i++)
        if (h->graphs[i].mode == mode && h->graphs[i].cap == dcap) {
            ORBX_CUDA(cudaGraphLaunch(h->graphs[i].exec, h->stream));
            return ORBX_OK;
        }
    // capture on the handle's own stream (a caller-provided stream may be the legacy default stream, which cannot capture)
    cudaGraph_t graph = nullptr;
    ORBX_CUDA(cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeThreadLocal));
    int rc = run_extract_on(h, 0, 1, mode, h->d_kps, h->d_desc, dcap, h->d_counts, h->own_stream, false, 0);
    cudaError_t e = cudaStreamEndCapture(h->own_stream, &graph);
    if (rc || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->use_graphs = false;      // capture is not possible here: fall back to plain launches for good
        return run_extract_on(h, 0, 1, mode, h->d_kps, h->d_desc, dcap, h->d_counts, h->stream, false, 0);
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        cudaGetLastError();
        h->use_graphs = false;
        return run_extract_on(h, 0, 1, mode, h->d_kps, h->d_desc, dcap, h->d_counts, h->stream, false, 0);
    }
    if (h->ngraphs == 4) { cudaGraphExecDestroy(h->graphs[0].exec); h->graphs[0] = h->graphs[3]; h->ngraphs = 3; }
    h->graphs[h->ngraphs].exec = exec;
    h->graphs[h->ngraphs].mode = mode;
    h->graphs[h->ngraphs].cap = dcap;
    h->ngraphs++;
    ORBX_CUDA(cudaGraphLaunch(exec, h->stream));
    return ORBX_OK;
}

static int extract_host(orbx_handle h, const uint8_t* const* frames, int nframes, int w, int hh, size_t stride, orbx_keypoint* out,
                        uint8_t* desc, int cap, int32_t* counts, int mode, const char* fn)
{
    ORBX_REQUIRE(h != nullptr, "%s: NULL handle", fn);
    ORBX_REQUIRE(frames && out && counts && (desc || !(mode & ORBX_DO_DESC)), "%s: NULL pointer", fn);
    ORBX_REQUIRE(nframes >= 1 && nframes <= h->max_batch, "%s: %d frames outside [1, max_batch=%d]", fn, nframes, h->max_batch);
    ORBX_REQUIRE(cap >= 1, "%s: capacity must be positive", fn);
    int rc = common_checks(h, frames[0], w, hh, stride, fn);
    if (rc) return rc;
    const int dcap = std::min(cap, h->dev_cap);
    if (nframes == 1) {
        // one frame per call is the reference's own pattern (src/FeatureExtractor.cpp:17,19): keep it lean -- everything on
        // one stream, the kernel sequence replayed as a CUDA graph, results copied speculatively at full capacity so that
        // a single synchronisation ends the call
        rc = upload_frames(h, frames, 0, 1, w, hh, stride, cudaMemcpyHostToDevice, h->stream);
        if (rc) return rc;
        rc = run_extract_single(h, mode, dcap);
        if (rc) return rc;
        h->last_nframes = (mode & ORBX_DO_DESC) ? 1 : 0;
        h->filter_nframes = 0;
        h->back_n = 0;
        h->last_cap = dcap;
        ORBX_CUDA(cudaMemcpyAsync(h->h_ctr, h->d_ctr, offsetof(FrameCounters, hist), cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(out, h->d_kps, (size_t)dcap * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, h->stream));
        if (mode & ORBX_DO_DESC) ORBX_CUDA(cudaMemcpyAsync(desc, h->d_desc, (size_t)dcap * 32, cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(cudaStreamSynchronize(h->stream));
        counts[0] = h->h_ctr[0].total;
        return check_counters(h, 1, dcap);
    }
    // chunks of 16 frames: the upload of chunk i+1 (copy stream) overlaps the kernels of chunk i (compute stream)
    const int chunk = 16, nchunks = div_up(nframes, chunk);
    while ((int)h->copy_events->size() < nchunks) {
        cudaEvent_t e;
        ORBX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->copy_events->push_back(e);
    }
    ORBX_CUDA(cudaEventRecord(h->order_event, h->stream));          // e.g. an earlier asynchronous _dev submission
    ORBX_CUDA(cudaStreamWaitEvent(h->copy_stream, h->order_event, 0));
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * chunk, n = std::min(chunk, nframes - f0);
        rc = upload_frames(h, frames, f0, n, w, hh, stride, cudaMemcpyHostToDevice, h->copy_stream);
        if (rc) return rc;
        ORBX_CUDA(cudaEventRecord((*h->copy_events)[c], h->copy_stream));
    }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * chunk, n = std::min(chunk, nframes - f0);
        ORBX_CUDA(cudaStreamWaitEvent(h->stream, (*h->copy_events)[c], 0));
        rc = run_extract(h, f0, n, mode, h->d_kps, h->d_desc, dcap, h->d_counts, f0);
        if (rc) return rc;
    }
    h->last_nframes = (mode & ORBX_DO_DESC) ? nframes : 0;
    h->filter_nframes = 0;
    h->back_n = 0;
    h->last_cap = dcap;
    ORBX_CUDA(copy_counters_d2h(h->h_ctr, h->d_ctr, nframes, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < nframes; f++) counts[f] = h->h_ctr[f].total;
    rc = check_counters(h, nframes, dcap);
    if (rc) return rc;
    // device rows are laid out [frame][dcap]; the caller's are [frame][cap]
    int64_t total = 0;
    for (int f = 0; f < nframes; f++) total += counts[f];
    if (dcap == cap && total * 2 >= (int64_t)nframes * cap) {
        ORBX_CUDA(cudaMemcpyAsync(out, h->d_kps, (size_t)nframes * cap * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, h->stream));
        if (mode & ORBX_DO_DESC)
            ORBX_CUDA(cudaMemcpyAsync(desc, h->d_desc, (size_t)nframes * cap * 32, cudaMemcpyDeviceToHost, h->stream));
    } else {
        for (int f = 0; f < nframes; f++) {
            if (!counts[f]) continue;
            ORBX_CUDA(cudaMemcpyAsync(out + (size_t)f * cap, h->d_kps + (size_t)f * dcap, (size_t)counts[f] * sizeof(orbx_keypoint),
                                      cudaMemcpyDeviceToHost, h->stream));
            if (mode & ORBX_DO_DESC)
                ORBX_CUDA(cudaMemcpyAsync(desc + (size_t)f * cap * 32, h->d_desc + (size_t)f * dcap * 32, (size_t)counts[f] * 32,
                                          cudaMemcpyDeviceToHost, h->stream));
        }
    }
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int orbx_detect(orbx_handle h, const uint8_t* gray, int w, int hh, size_t stride, orbx_keypoint* out, int cap, int* n)
{
    ORBX_REQUIRE(n != nullptr, "orbx_detect: n is NULL");
    int32_t cnt = 0;
    const uint8_t* frames[1] = { gray };
    int rc = extract_host(h, frames, 1, w, hh, stride, out, nullptr, cap, &cnt, ORBX_DO_ANGLE, "orbx_detect");
    *n = cnt;
    return rc;
}

extern "C" int orbx_detect_and_compute(orbx_handle h, const uint8_t* gray, int w, int hh, size_t stride, orbx_keypoint* out,
                                       uint8_t* desc, int cap, int* n)
{
    ORBX_REQUIRE(n != nullptr, "orbx_detect_and_compute: n is NULL");
    int32_t cnt = 0;
    const uint8_t* frames[1] = { gray };
    int rc = extract_host(h, frames, 1, w, hh, stride, out, desc, cap, &cnt, ORBX_DO_ANGLE | ORBX_DO_DESC, "orbx_detect_and_compute");
    *n = cnt;
    return rc;
}

extern "C" int orbx_extract_batch(orbx_handle h, const uint8_t* const* frames, int nframes, int w, int hh, size_t stride,
                                  orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts)
{
    return extract_host(h, frames, nframes, w, hh, stride, out, desc, cap, counts, ORBX_DO_ANGLE | ORBX_DO_DESC, "orbx_extract_batch");
}

extern "C" int orbx_extract_batch_dev(orbx_handle h, const uint8_t* d_frames, size_t frame_pitch_bytes, int nframes, int w, int hh,
                                      size_t stride, orbx_keypoint* d_out, uint8_t* d_desc, int cap, int32_t* d_counts)
{
    ORBX_REQUIRE(h != nullptr, "orbx_extract_batch_dev: NULL handle");
    ORBX_REQUIRE(d_frames && d_out && d_desc && d_counts, "orbx_extract_batch_dev: NULL pointer");
    ORBX_REQUIRE(nframes >= 1 && nframes <= h->max_batch, "orbx_extract_batch_dev: %d frames outside [1, max_batch=%d]", nframes, h->max_batch);
    ORBX_REQUIRE(cap >= 1 && frame_pitch_bytes >= stride * (size_t)hh, "orbx_extract_batch_dev: bad capacity or frame pitch");
    if (((uintptr_t)d_desc & 3) || ((uintptr_t)d_out & 3)) { set_error("orbx_extract_batch_dev: output pointers must be 4-byte aligned"); return ORBX_E_ALIGN; }
    int rc = common_checks(h, d_frames, w, hh, stride, "orbx_extract_batch_dev");
    if (rc) return rc;
    if (h->channels == 3)
        ORBX_CUDA(launch_ingest_bgr(d_frames, frame_pitch_bytes, stride, w, hh, h->d_slots, h->slot_stride, h->g.lv[0], nframes, h->stream));
    else
        ORBX_CUDA(launch_ingest(d_frames, frame_pitch_bytes, stride, w, hh, h->d_slots, h->slot_stride, h->g.lv[0], nframes, h->stream));
    rc = run_extract(h, 0, nframes, ORBX_DO_ANGLE | ORBX_DO_DESC, d_out, d_desc, cap, d_counts, 0);
    if (rc) return rc;
    h->dev_pending = std::max(h->dev_pending, nframes);
    h->dev_unordered = true;
    h->last_nframes = h->filter_nframes = h->back_n = 0;     // the handle's own arrays were not written
    return ORBX_OK;
}

extern "C" int orbx_check_dev(orbx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "orbx_check_dev: NULL handle");
    ORBX_CUDA(cudaSetDevice(h->device));
    if (!h->dev_pending) { ORBX_CUDA(cudaStreamSynchronize(h->stream)); return ORBX_OK; }
    // only the slots the unchecked submissions used: a flag left in a higher slot by an earlier, larger batch is stale
    const int n = h->dev_pending;
    ORBX_CUDA(copy_counters_d2h(h->h_ctr, h->d_ctr, n, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    h->dev_pending = 0;
    h->dev_unordered = false;
    for (int f = 0; f < n; f++)
        if (h->h_ctr[f].overflow) {
            set_error("orbx_check_dev: frame slot %d overflowed (flags %d, %d keypoints)", f, h->h_ctr[f].overflow, h->h_ctr[f].total);
            return ORBX_E_CAPACITY;
        }
    return ORBX_OK;
}

extern "C" int orbx_compute(orbx_handle h, const uint8_t* gray, int w, int hh, size_t stride, orbx_keypoint* kps, int* n, uint8_t* desc)
{
    ORBX_REQUIRE(n != nullptr && *n >= 0, "orbx_compute: bad keypoint count");
    ORBX_REQUIRE(*n == 0 || (kps && desc), "orbx_compute: NULL pointer");
    int rc = common_checks(h, gray, w, hh, stride, "orbx_compute");
    if (rc) return rc;
    // OpenCV: KeyPointsFilter::runByImageBorder on full-resolution coordinates (rounded), then a stable regroup by octave
    const int b = h->p.edge_threshold;
    std::vector<orbx_keypoint> kept;
    kept.reserve(*n);
    if (!(hh <= 2 * b || w <= 2 * b)) {
        for (int i = 0; i < *n; i++) {
            ORBX_REQUIRE(kps[i].octave >= 0 && kps[i].octave < h->p.nlevels, "orbx_compute: keypoint %d has octave %d outside [0,%d)", i,
                         kps[i].octave, h->p.nlevels);
            const int xi = rne_f(kps[i].x), yi = rne_f(kps[i].y);
            if (xi >= b && xi < w - b && yi >= b && yi < hh - b) kept.push_back(kps[i]);
        }
        std::stable_sort(kept.begin(), kept.end(), [](const orbx_keypoint& a, const orbx_keypoint& c) { return a.octave < c.octave; });
    }
    const int m = (int)kept.size();
    ORBX_REQUIRE(m <= h->dev_cap, "orbx_compute: %d keypoints exceed the handle's capacity %d", m, h->dev_cap);
    *n = m;
    if (m == 0) return ORBX_OK;
    memcpy(kps, kept.data(), (size_t)m * sizeof(orbx_keypoint));
    h->last_nframes = h->filter_nframes = h->back_n = 0;    // slot 0 and d_kps / d_desc are overwritten below
    const uint8_t* frames[1] = { gray };
    rc = upload_frames(h, frames, 0, 1, w, hh, stride, cudaMemcpyHostToDevice, h->stream);
    if (rc) return rc;
    rc = build_pyramids(h, 0, 1);   // the reference path rebuilds the pyramid in compute() as well
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(h->d_kps, kps, (size_t)m * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA(launch_describe_given(h->g, h->d_slots, h->d_kps, m, h->d_desc, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(desc, h->d_desc, (size_t)m * 32, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

// ---------------------------------------------------------------------------------------------- sequence mode
extern "C" int hamx_set_stream(hamx_handle h, void* cuda_stream);

extern "C" int orbx_reset_sequence(orbx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "orbx_reset_sequence: NULL handle");
    h->have_prev = false;
    h->nhist = 0;
    return ORBX_OK;
}

static int ensure_filter_buffers(orbx_handle h);

// history (descriptors, counts, keypoints; double-buffered) and the per-lane result buffers of the `back`-predecessor mode
static int ensure_back_buffers(orbx_handle h)
{
    if (h->d_hist[0]) return ORBX_OK;
    for (int i = 0; i < 2; i++) {
        ORBX_CUDA(cudaMalloc((void**)&h->d_hist[i], (size_t)ORBX_MAX_BACK * h->dev_cap * 32 + 256));
        ORBX_CUDA(cudaMalloc((void**)&h->d_hist_counts[i], 256));
        ORBX_CUDA(cudaMalloc((void**)&h->d_hist_kps[i], (size_t)ORBX_MAX_BACK * h->dev_cap * sizeof(orbx_keypoint) + 256));
    }
    const size_t slots = (size_t)ORBX_LANES * h->max_batch * ORBX_MAX_BACK;
    ORBX_CUDA(cudaMalloc((void**)&h->d_good_back, slots * h->dev_cap * sizeof(orbx_dmatch) + 256));
    ORBX_CUDA(cudaMalloc((void**)&h->d_ngood_back, slots * sizeof(int64_t) + 256));
    ORBX_CUDA(cudaMallocHost((void**)&h->h_ngood_back, slots * sizeof(int64_t)));
    ORBX_CUDA(cudaMalloc((void**)&h->d_fstatus_back, slots * h->dev_cap + 256));
    ORBX_CUDA(cudaMalloc((void**)&h->d_fF_back, slots * 9 * sizeof(double) + 256));
    ORBX_CUDA(cudaMalloc((void**)&h->d_finfo_back, slots * 4 * sizeof(int32_t) + 256));
    ORBX_CUDA(cudaMallocHost((void**)&h->h_finfo_back, slots * 4 * sizeof(int32_t)));
    h->nhist = 0;
    h->hist_cap = 0;
    return ORBX_OK;
}

// D2H of a [rows][dcap] device array into the caller's [rows][cap] array (cap >= dcap), `elem` bytes per entry
static cudaError_t copy_rows_d2h(void* dst, int cap, const void* src, int dcap, size_t rows, size_t elem, cudaStream_t s)
{
    if (cap == dcap) return cudaMemcpyAsync(dst, src, rows * (size_t)cap * elem, cudaMemcpyDeviceToHost, s);
    return cudaMemcpy2DAsync(dst, (size_t)cap * elem, src, (size_t)dcap * elem, (size_t)dcap * elem, rows, cudaMemcpyDeviceToHost, s);
}

// the caller states the geometry of its buffers; it must be the batch the handle holds
static int check_seq_buffers(int nframes, int cap, int have_n, int have_cap, const char* fn)
{
    ORBX_REQUIRE(nframes == have_n, "%s: the caller's buffers are for %d frames, the handle's last batch has %d", fn, nframes, have_n);
    ORBX_REQUIRE(cap >= have_cap, "%s: the caller's rows hold %d entries, the handle's last batch has rows of %d (the cap of that extract call, "
                 "limited to orbx_max_keypoints())", fn, cap, have_cap);
    return ORBX_OK;
}

extern "C" int orbx_match_consecutive(orbx_handle h, hamx_handle m, float ratio, int nframes, int cap, orbx_dmatch* good, int64_t* ngood)
{
    ORBX_REQUIRE(h != nullptr && m != nullptr, "orbx_match_consecutive: NULL handle");
    ORBX_REQUIRE(good && ngood, "orbx_match_consecutive: NULL pointer");
    ORBX_REQUIRE(h->last_nframes >= 1, "orbx_match_consecutive: no batch with descriptors has been extracted on this handle");
    { int rc_ = require_idle(h, "orbx_match_consecutive"); if (rc_) return rc_; }
    { int rc_ = check_seq_buffers(nframes, cap, h->last_nframes, h->last_cap, "orbx_match_consecutive"); if (rc_) return rc_; }
    ORBX_CUDA(cudaSetDevice(h->device));
    const int n = h->last_nframes, dcap = h->last_cap;
    if (h->have_prev && h->prev_cap != dcap) h->have_prev = false;     // the kept frame was cut at another capacity
    int rc;
    {
        ScopedHamxStream on(m, h->stream);   // same stream as the extraction: ordered after it, no extra sync
        if (on.rc) return on.rc;
        rc = hamx_match_consecutive_dev(m, h->d_desc, h->d_counts, n, dcap, h->have_prev ? h->d_prev_desc : nullptr,
                                        h->have_prev ? h->d_prev_count : nullptr, ratio, h->d_good, h->d_ngood);
    }
    if (rc) return rc;
    ORBX_CUDA(cudaMemcpyAsync(h->h_ngood, h->d_ngood, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(copy_rows_d2h(good, cap, h->d_good, dcap, (size_t)n, sizeof(orbx_dmatch), h->stream));
    // keep the last frame's descriptors (and keypoints, for orbx_filter_consecutive) for the next batch
    ORBX_CUDA(cudaMemcpyAsync(h->d_prev_desc, h->d_desc + (size_t)(n - 1) * dcap * 32, (size_t)dcap * 32, cudaMemcpyDeviceToDevice, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->d_prev_count, h->d_counts + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    { int rc_ = ensure_filter_buffers(h); if (rc_) return rc_; }
    h->filter_prev_kps = h->have_prev ? h->d_prev_kps[h->prev_kps_cur] : nullptr;
    h->prev_kps_cur ^= 1;
    ORBX_CUDA(cudaMemcpyAsync(h->d_prev_kps[h->prev_kps_cur], h->d_kps + (size_t)(n - 1) * dcap, (size_t)dcap * sizeof(orbx_keypoint),
                              cudaMemcpyDeviceToDevice, h->stream));
    h->filter_nframes = n;
    h->filter_cap = dcap;
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    h->have_prev = true;
    h->prev_cap = dcap;
    for (int f = 0; f < n; f++) ngood[f] = h->h_ngood[f];
    return ORBX_OK;
}

// outputs of the outlier filter for every lane's slots, and the keypoints of the frame before the batch
static int ensure_filter_buffers(orbx_handle h)
{
    if (!h->d_fstatus) {
        const size_t slots = (size_t)ORBX_LANES * h->max_batch;
        ORBX_CUDA(cudaMalloc((void**)&h->d_fstatus, slots * h->dev_cap + 256));
        ORBX_CUDA(cudaMalloc((void**)&h->d_fF, slots * 9 * sizeof(double) + 256));
        ORBX_CUDA(cudaMalloc((void**)&h->d_finfo, slots * 4 * sizeof(int32_t) + 256));
        ORBX_CUDA(cudaMallocHost((void**)&h->h_finfo, slots * 4 * sizeof(int32_t)));
    }
    if (!h->d_prev_kps[0])
        for (int i = 0; i < 2; i++) ORBX_CUDA(cudaMalloc((void**)&h->d_prev_kps[i], (size_t)h->dev_cap * sizeof(orbx_keypoint) + 256));
    return ORBX_OK;
}

extern "C" int orbx_filter_consecutive(orbx_handle h, fmx_handle fm, double max_distance, double confidence, int nframes, int cap,
                                       uint8_t* status, double* F, int32_t* ninliers)
{
    ORBX_REQUIRE(h != nullptr && fm != nullptr, "orbx_filter_consecutive: NULL handle");
    ORBX_REQUIRE(status && F && ninliers, "orbx_filter_consecutive: NULL pointer");
    ORBX_REQUIRE(h->filter_nframes >= 1, "orbx_filter_consecutive: orbx_match_consecutive has not run on this handle's last batch");
    { int rc_ = require_idle(h, "orbx_filter_consecutive"); if (rc_) return rc_; }
    { int rc_ = check_seq_buffers(nframes, cap, h->filter_nframes, h->filter_cap, "orbx_filter_consecutive"); if (rc_) return rc_; }
    ORBX_CUDA(cudaSetDevice(h->device));
    const int n = h->filter_nframes, dcap = h->filter_cap;
    int rc = ensure_filter_buffers(h);
    if (rc) return rc;
    {
        ScopedFmxStream on(fm, h->stream);
        if (on.rc) return on.rc;
        rc = fmx_filter_consecutive_dev(fm, h->d_kps, h->filter_prev_kps, n, dcap, h->d_good, h->d_ngood, max_distance, confidence,
                                        h->d_fstatus, h->d_fF, h->d_finfo);
    }
    if (rc) return rc;
    ORBX_CUDA(copy_rows_d2h(status, cap, h->d_fstatus, dcap, (size_t)n, 1, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(F, h->d_fF, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->h_finfo, h->d_finfo, (size_t)n * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < n; f++) ninliers[f] = h->h_finfo[4 * f];
    return ORBX_OK;
}

extern "C" int orbx_filter_back(orbx_handle h, fmx_handle fm, double max_distance, double confidence, int nframes, int back, int cap,
                                uint8_t* status, double* F, int32_t* ninliers)
{
    ORBX_REQUIRE(h != nullptr && fm != nullptr, "orbx_filter_back: NULL handle");
    ORBX_REQUIRE(status && F && ninliers, "orbx_filter_back: NULL pointer");
    ORBX_REQUIRE(h->back_n >= 1, "orbx_filter_back: orbx_match_back has not run on this handle's last batch");
    { int rc_ = require_idle(h, "orbx_filter_back"); if (rc_) return rc_; }
    { int rc_ = check_seq_buffers(nframes, cap, h->back_n, h->back_cap, "orbx_filter_back"); if (rc_) return rc_; }
    ORBX_REQUIRE(back == h->back_back, "orbx_filter_back: the caller's buffers are for %d predecessors, orbx_match_back ran with %d", back, h->back_back);
    ORBX_CUDA(cudaSetDevice(h->device));
    const int n = h->back_n, dcap = h->back_cap, npairs = n * back;
    { int rc_ = ensure_back_buffers(h); if (rc_) return rc_; }
    int rc;
    {
        ScopedFmxStream on(fm, h->stream);
        if (on.rc) return on.rc;
        // orbx_match_back matched against the history that is now the inactive buffer (hist_cur was flipped after the update)
        rc = fmx_filter_back_dev(fm, h->d_kps, n, dcap, back, h->d_hist_kps[h->hist_cur ^ 1], h->back_nhist, h->d_good_back, h->d_ngood_back,
                                 max_distance, confidence, h->d_fstatus_back, h->d_fF_back, h->d_finfo_back);
    }
    if (rc) return rc;
    ORBX_CUDA(copy_rows_d2h(status, cap, h->d_fstatus_back, dcap, (size_t)npairs, 1, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(F, h->d_fF_back, (size_t)npairs * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->h_finfo_back, h->d_finfo_back, (size_t)npairs * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < npairs; i++) ninliers[i] = h->h_finfo_back[4 * i];
    return ORBX_OK;
}

extern "C" int orbx_match_back(orbx_handle h, hamx_handle m, int back, float ratio, int nframes, int cap, orbx_dmatch* good, int64_t* ngood)
{
    ORBX_REQUIRE(h != nullptr && m != nullptr, "orbx_match_back: NULL handle");
    ORBX_REQUIRE(good && ngood, "orbx_match_back: NULL pointer");
    ORBX_REQUIRE(back >= 1 && back <= ORBX_MAX_BACK, "orbx_match_back: back %d outside [1, %d]", back, ORBX_MAX_BACK);
    ORBX_REQUIRE(h->last_nframes >= 1, "orbx_match_back: no batch with descriptors has been extracted on this handle");
    { int rc_ = require_idle(h, "orbx_match_back"); if (rc_) return rc_; }
    { int rc_ = check_seq_buffers(nframes, cap, h->last_nframes, h->last_cap, "orbx_match_back"); if (rc_) return rc_; }
    ORBX_CUDA(cudaSetDevice(h->device));
    const int n = h->last_nframes, dcap = h->last_cap;
    { int rc_ = ensure_back_buffers(h); if (rc_) return rc_; }
    if (h->nhist && h->hist_cap != dcap) h->nhist = 0;      // a history laid out for another capacity cannot be indexed
    const int cur = h->hist_cur;
    int rc;
    {
        ScopedHamxStream on(m, h->stream);
        if (on.rc) return on.rc;
        rc = hamx_match_back_dev(m, h->d_desc, h->d_counts, n, dcap, back, h->d_hist[cur], h->d_hist_counts[cur], std::min(h->nhist, back), ratio,
                                 h->d_good_back, h->d_ngood_back);
        if (!rc) rc = hamx_update_history_dev(m, h->d_desc, h->d_counts, n, dcap, ORBX_MAX_BACK, h->d_hist[cur], h->d_hist_counts[cur], h->nhist,
                                              h->d_hist[cur ^ 1], h->d_hist_counts[cur ^ 1]);
    }
    if (rc) return rc;
    // the keypoint history follows the descriptor history (most recent frame first), for orbx_filter_back
    for (int hi = 0; hi < ORBX_MAX_BACK; hi++) {
        const int src = n - 1 - hi;
        const orbx_keypoint* from = src >= 0 ? h->d_kps + (size_t)src * dcap : (-src - 1 < h->nhist ? h->d_hist_kps[cur] + (size_t)(-src - 1) * dcap : nullptr);
        if (from)
            ORBX_CUDA(cudaMemcpyAsync(h->d_hist_kps[cur ^ 1] + (size_t)hi * dcap, from, (size_t)dcap * sizeof(orbx_keypoint), cudaMemcpyDeviceToDevice, h->stream));
    }
    h->back_n = n; h->back_back = back; h->back_cap = dcap; h->back_nhist = std::min(h->nhist, back);
    h->hist_cur = cur ^ 1;
    h->nhist = std::min(ORBX_MAX_BACK, h->nhist + n);
    h->hist_cap = dcap;
    ORBX_CUDA(cudaMemcpyAsync(h->h_ngood_back, h->d_ngood_back, (size_t)n * back * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(copy_rows_d2h(good, cap, h->d_good_back, dcap, (size_t)n * back, sizeof(orbx_dmatch), h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n * back; i++) ngood[i] = h->h_ngood_back[i];
    return ORBX_OK;
}

// ---------------------------------------------------------------------------------------------- pipelined host path
// Up to ORBX_LANES batches in flight: while batch k is being extracted and matched on the compute stream, the frames of the
// next batches cross PCIe on the copy stream (with three lanes the copy engine never waits for the host to collect a
// result) and the results of batch k-1 return on the D2H stream.
extern "C" int orbx_submit_batch(orbx_handle h, hamx_handle m, const uint8_t* const* frames, int nframes, int w, int hh, size_t stride,
                                 float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts, orbx_dmatch* good,
                                 int64_t* ngood)
{
    return orbx_submit_batch_filtered(h, m, nullptr, frames, nframes, w, hh, stride, ratio, out, desc, cap, counts, good, ngood, 0., 0.,
                                      nullptr, nullptr, nullptr);
}

static int submit_impl(orbx_handle h, hamx_handle m, fmx_handle fm, int back, const uint8_t* const* frames, int nframes, int w, int hh,
                       size_t stride, float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts, orbx_dmatch* good,
                       int64_t* ngood, double max_distance, double confidence, uint8_t* status, double* F, int32_t* ninliers);

extern "C" int orbx_submit_batch_filtered(orbx_handle h, hamx_handle m, fmx_handle fm, const uint8_t* const* frames, int nframes, int w, int hh,
                                          size_t stride, float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts,
                                          orbx_dmatch* good, int64_t* ngood, double max_distance, double confidence, uint8_t* status,
                                          double* F, int32_t* ninliers)
{
    return submit_impl(h, m, fm, 0, frames, nframes, w, hh, stride, ratio, out, desc, cap, counts, good, ngood, max_distance, confidence,
                       status, F, ninliers);
}

extern "C" int orbx_submit_batch_back(orbx_handle h, hamx_handle m, fmx_handle fm, int back, const uint8_t* const* frames, int nframes, int w,
                                      int hh, size_t stride, float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts,
                                      orbx_dmatch* good, int64_t* ngood, double max_distance, double confidence, uint8_t* status, double* F,
                                      int32_t* ninliers)
{
    ORBX_REQUIRE(m != nullptr, "orbx_submit_batch_back: a matcher is required");
    ORBX_REQUIRE(back >= 1 && back <= ORBX_MAX_BACK, "orbx_submit_batch_back: back %d outside [1, %d]", back, ORBX_MAX_BACK);
    return submit_impl(h, m, fm, back, frames, nframes, w, hh, stride, ratio, out, desc, cap, counts, good, ngood, max_distance, confidence,
                       status, F, ninliers);
}

static int submit_impl(orbx_handle h, hamx_handle m, fmx_handle fm, int back, const uint8_t* const* frames, int nframes, int w, int hh,
                       size_t stride, float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts, orbx_dmatch* good,
                       int64_t* ngood, double max_distance, double confidence, uint8_t* status, double* F, int32_t* ninliers)
{
    ORBX_REQUIRE(h != nullptr, "orbx_submit_batch: NULL handle");
    ORBX_REQUIRE(frames && out && desc && counts && (m == nullptr || (good && ngood)), "orbx_submit_batch: NULL pointer");
    ORBX_REQUIRE(fm == nullptr || (m != nullptr && status && F && ninliers), "orbx_submit_batch_filtered: the filter needs a matcher and its output buffers");
    ORBX_REQUIRE(nframes >= 1 && nframes <= h->max_batch, "orbx_submit_batch: %d frames outside [1, max_batch=%d]", nframes, h->max_batch);
    ORBX_REQUIRE(cap >= 1 && cap <= h->dev_cap, "orbx_submit_batch: capacity %d outside [1, %d]", cap, h->dev_cap);
    ORBX_REQUIRE(frames[0] && w >= 1 && hh >= 1 && stride >= (size_t)w * h->channels, "orbx_submit_batch: bad image geometry %dx%d stride %zu", w, hh, stride);
    const int li = h->lane_next;
    orbx_lane& L = h->lanes[li];
    ORBX_REQUIRE(!L.busy, "orbx_submit_batch: %d batches are already in flight; call orbx_wait_batch first", ORBX_LANES);
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = set_geometry(h, w, hh);      // drains the compute stream if the frame size changed
    if (rc) return rc;
    const int s0 = li * h->max_batch;
    // the lane's slots are free: its previous batch was collected by orbx_wait_batch and the blocking entry points drain
    // the compute stream before they return.  Only an unchecked asynchronous _dev submission can still be using lane 0;
    // then (and only then: the wait would also serialise this upload behind the other lane's kernels) order behind it.
    if (h->dev_unordered) {     // once: later uploads follow this one on the copy stream
        ORBX_CUDA(cudaEventRecord(h->order_event, h->stream));
        ORBX_CUDA(cudaStreamWaitEvent(h->copy_stream, h->order_event, 0));
        h->dev_unordered = false;
    }
    rc = upload_frames(h, frames, 0, nframes, w, hh, stride, cudaMemcpyHostToDevice, h->copy_stream, s0);
    if (rc) return rc;
    ORBX_CUDA(cudaEventRecord(L.uploaded, h->copy_stream));
    ORBX_CUDA(cudaStreamWaitEvent(h->stream, L.uploaded, 0));
    // every lane owns max_batch rows of dev_cap entries of the per-frame arrays; the rows of this batch have stride `cap`
    // inside the lane's region, so batches in flight with different capacities cannot overlap
    const size_t l0 = (size_t)s0 * h->dev_cap;
    orbx_keypoint* d_kps = h->d_kps + l0;
    uint8_t* d_desc = h->d_desc + l0 * 32;
    orbx_dmatch* d_good = h->d_good + l0;
    uint8_t* d_fstatus = h->d_fstatus ? h->d_fstatus + l0 : nullptr;
    rc = run_extract(h, s0, nframes, ORBX_DO_ANGLE | ORBX_DO_DESC, d_kps, d_desc, cap, h->d_counts + s0, 0);
    if (rc) return rc;
    const size_t b0 = (size_t)s0 * ORBX_MAX_BACK;          // this lane's first pair in the `back` buffers
    const size_t bl0 = b0 * h->dev_cap;
    if (m && back >= 1) {
        // the steady-state loop of pnpPoseEstimation (src/CameraPoseEstimator.cpp:405-419): every frame against its `back`
        // predecessors, matchFeatures + computeFundamentalMatrix per pair; the history (descriptors, keypoints) of the
        // frames before the batch is double-buffered on the device and advanced in stream order
        rc = ensure_back_buffers(h);
        if (rc) return rc;
        if (h->nhist && h->hist_cap != cap) h->nhist = 0;
        const int cur = h->hist_cur, nh = std::min(h->nhist, back);
        {
            ScopedHamxStream on(m, h->stream);
            if (on.rc) return on.rc;
            rc = hamx_match_back_dev(m, d_desc, h->d_counts + s0, nframes, cap, back, h->d_hist[cur], h->d_hist_counts[cur], nh, ratio,
                                     h->d_good_back + bl0, h->d_ngood_back + b0);
            if (!rc) rc = hamx_update_history_dev(m, d_desc, h->d_counts + s0, nframes, cap, ORBX_MAX_BACK, h->d_hist[cur], h->d_hist_counts[cur],
                                                  h->nhist, h->d_hist[cur ^ 1], h->d_hist_counts[cur ^ 1]);
        }
        if (rc) return rc;
        if (fm) {
            ScopedFmxStream on(fm, h->stream);
            if (on.rc) return on.rc;
            rc = fmx_filter_back_dev(fm, d_kps, nframes, cap, back, h->d_hist_kps[cur], nh, h->d_good_back + bl0,
                                     h->d_ngood_back + b0, max_distance, confidence, h->d_fstatus_back + bl0, h->d_fF_back + b0 * 9,
                                     h->d_finfo_back + b0 * 4);
            if (rc) return rc;
        }
        for (int hi = 0; hi < ORBX_MAX_BACK; hi++) {
            const int src = nframes - 1 - hi;
            const orbx_keypoint* from = src >= 0 ? d_kps + (size_t)src * cap
                                                 : (-src - 1 < h->nhist ? h->d_hist_kps[cur] + (size_t)(-src - 1) * cap : nullptr);
            if (from)
                ORBX_CUDA(cudaMemcpyAsync(h->d_hist_kps[cur ^ 1] + (size_t)hi * cap, from, (size_t)cap * sizeof(orbx_keypoint),
                                          cudaMemcpyDeviceToDevice, h->stream));
        }
        h->hist_cur = cur ^ 1;
        h->nhist = std::min(ORBX_MAX_BACK, h->nhist + nframes);
        h->hist_cap = cap;
    } else if (m) {
        if (h->have_prev && h->prev_cap != cap) h->have_prev = false;     // the kept frame is laid out for another capacity
        {
            ScopedHamxStream on(m, h->stream);
            if (on.rc) return on.rc;
            rc = hamx_match_consecutive_dev(m, d_desc, h->d_counts + s0, nframes, cap, h->have_prev ? h->d_prev_desc : nullptr,
                                            h->have_prev ? h->d_prev_count : nullptr, ratio, d_good, h->d_ngood + s0);
        }
        if (rc) return rc;
        ORBX_CUDA(cudaMemcpyAsync(h->d_prev_desc, d_desc + (size_t)(nframes - 1) * cap * 32, (size_t)cap * 32, cudaMemcpyDeviceToDevice, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(h->d_prev_count, h->d_counts + s0 + (nframes - 1), sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
        // computeFundamentalMatrix of every (frame, predecessor) pair; the keypoints of the batch's last frame are kept for
        // the next batch either way (stream order: the filter reads the old copy before it is overwritten)
        rc = ensure_filter_buffers(h);
        if (rc) return rc;
        d_fstatus = h->d_fstatus + l0;
        orbx_keypoint* prev_kps = h->d_prev_kps[h->prev_kps_cur];
        if (fm) {
            ScopedFmxStream on(fm, h->stream);
            if (on.rc) return on.rc;
            rc = fmx_filter_consecutive_dev(fm, d_kps, h->have_prev ? prev_kps : nullptr, nframes, cap, d_good,
                                            h->d_ngood + s0, max_distance, confidence, d_fstatus,
                                            h->d_fF + (size_t)s0 * 9, h->d_finfo + (size_t)s0 * 4);
            if (rc) return rc;
        }
        ORBX_CUDA(cudaMemcpyAsync(prev_kps, d_kps + (size_t)(nframes - 1) * cap, (size_t)cap * sizeof(orbx_keypoint),
                                  cudaMemcpyDeviceToDevice, h->stream));
        h->have_prev = true;
        h->prev_cap = cap;
    }
    ORBX_CUDA(cudaEventRecord(L.computed, h->stream));
    ORBX_CUDA(cudaStreamWaitEvent(h->d2h_stream, L.computed, 0));
    ORBX_CUDA(copy_counters_d2h(h->h_ctr + s0, h->d_ctr + s0, nframes, h->d2h_stream));
    ORBX_CUDA(cudaMemcpyAsync(out, d_kps, (size_t)nframes * cap * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, h->d2h_stream));
    ORBX_CUDA(cudaMemcpyAsync(desc, d_desc, (size_t)nframes * cap * 32, cudaMemcpyDeviceToHost, h->d2h_stream));
    if (m && back >= 1) {
        const size_t np = (size_t)nframes * back;
        ORBX_CUDA(cudaMemcpyAsync(good, h->d_good_back + bl0, np * cap * sizeof(orbx_dmatch), cudaMemcpyDeviceToHost, h->d2h_stream));
        ORBX_CUDA(cudaMemcpyAsync(h->h_ngood_back + b0, h->d_ngood_back + b0, np * sizeof(int64_t), cudaMemcpyDeviceToHost, h->d2h_stream));
        if (fm) {
            ORBX_CUDA(cudaMemcpyAsync(status, h->d_fstatus_back + bl0, np * cap, cudaMemcpyDeviceToHost, h->d2h_stream));
            ORBX_CUDA(cudaMemcpyAsync(F, h->d_fF_back + b0 * 9, np * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
            ORBX_CUDA(cudaMemcpyAsync(h->h_finfo_back + b0 * 4, h->d_finfo_back + b0 * 4, np * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                      h->d2h_stream));
        }
    } else if (m) {
        ORBX_CUDA(cudaMemcpyAsync(good, d_good, (size_t)nframes * cap * sizeof(orbx_dmatch), cudaMemcpyDeviceToHost, h->d2h_stream));
        ORBX_CUDA(cudaMemcpyAsync(h->h_ngood + s0, h->d_ngood + s0, (size_t)nframes * sizeof(int64_t), cudaMemcpyDeviceToHost, h->d2h_stream));
    }
    if (fm && back == 0) {
        ORBX_CUDA(cudaMemcpyAsync(status, d_fstatus, (size_t)nframes * cap, cudaMemcpyDeviceToHost, h->d2h_stream));
        ORBX_CUDA(cudaMemcpyAsync(F, h->d_fF + (size_t)s0 * 9, (size_t)nframes * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
        ORBX_CUDA(cudaMemcpyAsync(h->h_finfo + (size_t)s0 * 4, h->d_finfo + (size_t)s0 * 4, (size_t)nframes * 4 * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, h->d2h_stream));
    }
    ORBX_CUDA(cudaEventRecord(L.done, h->d2h_stream));
    L.busy = true;
    L.nframes = nframes;
    L.cap = cap;
    L.counts = counts;
    L.ngood = m ? ngood : nullptr;
    L.ninliers = fm ? ninliers : nullptr;
    L.back = m ? back : 0;
    h->last_nframes = 0;            // orbx_match_consecutive pairs with orbx_extract_batch only
    h->filter_nframes = 0;
    h->back_n = 0;
    h->lane_next = (li + 1) % ORBX_LANES;
    return ORBX_OK;
}

extern "C" int orbx_wait_batch(orbx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "orbx_wait_batch: NULL handle");
    const int li = h->lane_head;
    orbx_lane& L = h->lanes[li];
    ORBX_REQUIRE(L.busy, "orbx_wait_batch: no batch in flight");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaEventSynchronize(L.done));
    L.busy = false;
    h->lane_head = (li + 1) % ORBX_LANES;
    const int s0 = li * h->max_batch;
    for (int f = 0; f < L.nframes; f++) {
        L.counts[f] = h->h_ctr[s0 + f].total;
        if (L.back) continue;
        if (L.ngood) L.ngood[f] = h->h_ngood[s0 + f];
        if (L.ninliers) L.ninliers[f] = h->h_finfo[(size_t)(s0 + f) * 4];
    }
    if (L.back) {
        const size_t b0 = (size_t)s0 * ORBX_MAX_BACK;
        for (int p = 0; p < L.nframes * L.back; p++) {
            L.ngood[p] = h->h_ngood_back[b0 + p];
            if (L.ninliers) L.ninliers[p] = h->h_finfo_back[(b0 + p) * 4];
        }
    }
    return check_counters(h, L.nframes, L.cap, s0);
}

extern "C" int orbx_pipeline_depth(orbx_handle h) { return h ? ORBX_LANES : ORBX_E_INVALID; }

extern "C" int orbx_batches_in_flight(orbx_handle h)
{
    if (!h) return ORBX_E_INVALID;
    int n = 0;
    for (int l = 0; l < ORBX_LANES; l++) n += h->lanes[l].busy ? 1 : 0;
    return n;
}

// ---------------------------------------------------------------------------------------------- debug taps
extern "C" int orbx_debug_pyramid_level(orbx_handle h, const uint8_t* gray, int w, int hh, size_t stride, int level, uint8_t* out)
{
    int rc = common_checks(h, gray, w, hh, stride, "orbx_debug_pyramid_level");
    if (rc) return rc;
    ORBX_REQUIRE(out && level >= 0 && level < h->g.nlevels, "orbx_debug_pyramid_level: bad level %d", level);
    h->last_nframes = h->filter_nframes = h->back_n = 0;
    const uint8_t* frames[1] = { gray };
    rc = upload_frames(h, frames, 0, 1, w, hh, stride, cudaMemcpyHostToDevice, h->stream);
    if (rc) return rc;
    rc = build_pyramids(h, 0, 1);
    if (rc) return rc;
    const LevelGeom& L = h->g.lv[level];
    ORBX_CUDA(cudaMemcpy2DAsync(out, L.w, h->d_slots + L.img_off, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int orbx_debug_fast_level(orbx_handle h, const uint8_t* gray, int w, int hh, size_t stride, int level, int32_t* xs,
                                     int32_t* ys, int32_t* scores, int cap, int* n)
{
    int rc = common_checks(h, gray, w, hh, stride, "orbx_debug_fast_level");
    if (rc) return rc;
    ORBX_REQUIRE(xs && ys && scores && n && level >= 0 && level < h->g.nlevels, "orbx_debug_fast_level: bad arguments");
    h->last_nframes = h->filter_nframes = h->back_n = 0;
    const uint8_t* frames[1] = { gray };
    rc = upload_frames(h, frames, 0, 1, w, hh, stride, cudaMemcpyHostToDevice, h->stream);
    if (rc) return rc;
    ORBX_CUDA(cudaMemsetAsync(h->d_ctr, 0, sizeof(FrameCounters), h->stream));
    rc = build_pyramids(h, 0, 1);
    if (rc) return rc;
    ORBX_CUDA(launch_fast(h->g, h->d_slots, h->slot_stride, h->d_cand, h->cand_stride, h->d_ctr, 1, nullptr, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    const LevelGeom& L = h->g.lv[level];
    const int cnt = h->h_ctr[0].ncand[level];
    *n = cnt;
    if (cnt > L.cand_cap) { set_error("orbx_debug_fast_level: candidate overflow (%d > %d)", cnt, L.cand_cap); return ORBX_E_CAPACITY; }
    std::vector<Cand> v((size_t)cnt);
    if (cnt) ORBX_CUDA(cudaMemcpy(v.data(), h->d_cand + L.cand_off, (size_t)cnt * sizeof(Cand), cudaMemcpyDeviceToHost));
    std::sort(v.begin(), v.end(), [](const Cand& a, const Cand& c) { return a.xy < c.xy; });   // raster order (y, x)
    for (int i = 0; i < cnt && i < cap; i++) {
        xs[i] = (int32_t)(v[i].xy & 0xFFFFu);
        ys[i] = (int32_t)(v[i].xy >> 16);
        scores[i] = (int32_t)v[i].score;
    }
    return ORBX_OK;
}
