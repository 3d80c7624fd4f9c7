// common.cuh -- shared declarations of liborbx.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <exception>
#include <new>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>

#include "../../include/orbx.h"

#define ORBX_MAX_LEVELS 16

namespace orbx {

void set_error(const char* fmt, ...);

#define ORBX_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            orbx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return ORBX_E_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

// The same for constructors: `cleanup` (e.g. orbx_destroy(h)) releases what has been allocated so far.
#define ORBX_CUDA_OR(call, cleanup)                                                              \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            orbx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            cleanup;                                                                             \
            return ORBX_E_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

// Entry points whose host side allocates (std::vector, std::thread): no C++ exception may cross the C ABI.
#define ORBX_NOTHROW(expr)                                                                                   \
    try { return (expr); }                                                                                   \
    catch (const std::bad_alloc&) { orbx::set_error("%s: out of host memory", __func__); return ORBX_E_ALLOC; } \
    catch (const std::exception& e_) { orbx::set_error("%s: %s", __func__, e_.what()); return ORBX_E_INVALID; }

#define ORBX_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            orbx::set_error(__VA_ARGS__);  \
            return ORBX_E_INVALID;         \
        }                                  \
    } while (0)

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Geometry of one pyramid level inside a frame slot (all offsets in bytes from the slot base).
struct LevelGeom {
    int w, h;          // level size in pixels
    int pitch;         // row pitch in bytes (multiple of 128)
    int quota;         // q[l]: keypoints wanted on this level
    int cand_cap;      // capacity of the FAST candidate list
    int surv_cap;      // capacity of the survivor list (after the FAST-score cut)
    int tiles_x, tiles_y, tile_start;  // FAST tile grid over the interior [31, w-31) x [31, h-31)
    float scale;       // layer scale s_l
    float inv_scale;   // 1.f / s_l
    size_t img_off;    // offset of the level image
    size_t cand_off;   // index (in uint2 entries) of this level's candidate list inside the slot's list
    size_t surv_off;   // index of this level's survivor list
    size_t sel_off;    // index of this level's selected list
    // resize tables (device pointers, valid for l >= 1): source offset and second-tap weight (0..256) per dst col/row
    const int* ofs_x; const int* ofs_y;
    const uint16_t* c1x; const uint16_t* c1y;
};

struct FrameGeom {
    int nlevels;
    int w, h;
    int score_type;
    int fast_threshold;
    int total_tiles;
    LevelGeom lv[ORBX_MAX_LEVELS];
};

// A detected corner on its level.
struct Cand { uint32_t xy; uint32_t score; };       // xy = y << 16 | x
struct Sel  { uint32_t xy; float response; };

// Per-frame counters in device memory (zeroed at the start of every submission).
struct FrameCounters {
    int32_t ncand[ORBX_MAX_LEVELS];
    int32_t nsurv[ORBX_MAX_LEVELS];
    int32_t nsel[ORBX_MAX_LEVELS];
    int32_t total;        // keypoints written for this frame
    int32_t overflow;     // bit 0: candidate list, bit 1: survivor list, bit 2: output capacity
    uint32_t hist[ORBX_MAX_LEVELS][256];   // histogram of FAST scores per level
};

// kernel launchers (defined next to their kernels)
// fast_hint: one int per level in device memory, written by K3 (launch_select) and read by the next K2 launch: non-zero =
// the level is corner-dense, skip K2's compass pre-test (a speed hint; may be NULL)
cudaError_t launch_fast(const FrameGeom& g, uint8_t* slots, size_t slot_stride, Cand* cand, size_t cand_stride,
                        FrameCounters* ctr, int nframes, const int32_t* fast_hint, cudaStream_t s);
cudaError_t launch_select(const FrameGeom& g, const Cand* cand, size_t cand_stride, Cand* surv, size_t surv_stride,
                          FrameCounters* ctr, int nframes, int32_t* fast_hint, cudaStream_t s);
// mode bits for launch_orient_describe
enum { ORBX_DO_ANGLE = 1, ORBX_DO_DESC = 2 };
cudaError_t launch_orient_describe(const FrameGeom& g, const uint8_t* slots, size_t slot_stride, const Sel* sel,
                                   size_t sel_stride, FrameCounters* ctr, orbx_keypoint* out, uint8_t* desc, int cap,
                                   int32_t* counts, int nframes, int mode, cudaStream_t s);
cudaError_t launch_describe_given(const FrameGeom& g, const uint8_t* slot, const orbx_keypoint* kps, int n, uint8_t* desc,
                                  cudaStream_t s);
size_t harris_select_smem(int max_surv_cap);

}  // namespace orbx
