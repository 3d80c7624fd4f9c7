// jpeg.cu -- frame ingest for baseline JPEG files, grey-scale and YCbCr colour (sm_100a): what cv::imread does for the reference's frame loader
// (src/FrameLoader.cpp:62, imread(path, CV_LOAD_IMAGE_UNCHANGED); for .jpg that is libjpeg behind OpenCV, JDCT_ISLOW), for a
// whole batch of files at once and straight into the device-resident frames the extractor reads.  Only the compressed bytes
// cross PCIe (a fifth of the raw frames at quality 90).
//
//   host      marker parsing (ITU-T T.81 Annex B): DQT, SOF0/1, DHT, DRI, SOS of an 8-bit Huffman file with one component or
//             three (YCbCr 4:2:0 / 4:2:2 / 4:4:4, one interleaved scan); the
//             entropy-coded segment is cut at its restart markers into intervals (T.81 E.2.4) -- the unit of parallelism
//   K17 k_jpeg_unstuff, k_jpeg_sync / k_jpeg_huff   a warp per restart interval strips the stuffed zero bytes (FF 00 -> FF) into
//             a scratch copy (whole files without restart markers: a CTA per 8 KB chunk); the Huffman decoder of T.81 F.2.2 (10-bit lookahead tables, the DC predictor restarting with the
//             interval, one symbol per loop iteration so that lanes on different data stay in step) then runs either one lane
//             per interval (short intervals) or, for block rows and whole files, one lane per subsequence of <= 1024 bits,
//             iterated until every subsequence was decoded from its predecessor's true exit state; the non-zero coefficients
//             of each 8x8 block are scattered into a zeroed array
//   K18 k_jpeg_idct   one thread per block: dequantisation and libjpeg's jpeg_idct_islow (jidctint.c: 13-bit constants, two
//             passes, DESCALE) in registers, the range-limit table of jdmaster.c as arithmetic, 8-byte row stores that
//             coalesce across the blocks of a block row; for colour files into whole-block Y / Cb / Cr planes
//   K19 k_jpeg_ycc    colour files: libjpeg's fancy chroma upsampling (jdsample.c) and fixed-point YCbCr -> RGB (jdcolor.c), BGR out
//
// Bit-exact with cv2.imdecode (libjpeg-turbo 3.1.2) on every file of tests/golden/jpeg_cases.npz, with and without restart
// markers, grey and colour.  Anything else (progressive, 12-bit, arithmetic, other samplings) is refused with
// ORBX_E_UNSUPPORTED: the caller keeps its CPU decoder for those.
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"

namespace orbx {
namespace {

constexpr int JP_LOOK = 10;                       // lookahead bits of the Huffman tables
constexpr int JP_HUFF_THREADS = 128;
constexpr int JP_IDCT_THREADS = 128;
constexpr uint32_t JP_SUB_BITS = 1024;            // bits of an interval one lane decodes in the self-synchronising path
constexpr uint32_t JP_MAX_INTERVAL_BITS = 0xfffff000u;  // bit positions are 32-bit

struct JpHuff {                                   // one Huffman table on the device
    uint16_t look[1 << JP_LOOK];                  // (length << 8) | symbol for codes of at most JP_LOOK bits, 0: longer
    int32_t maxcode[18];                          // T.81 F.2.2.3, per code length; [17] = sentinel
    int32_t valoff[17];                           // valptr[l] - mincode[l]
    uint8_t vals[256];
};
struct JpTables {                                 // what one file's scan needs, per component (a grey file uses entry 0)
    JpHuff dc[3], ac[3];
    uint16_t quant[3][64];                        // natural order
};
constexpr int JP_MAX_SLOTS = 6;                   // blocks of an MCU: 1 (grey), 3 (4:4:4), 4 (4:2:2), 6 (4:2:0)
struct JpGeom {                                   // the same for every file of a batch
    int ncomp, bpm;                               // components; blocks per MCU
    int mcus_x;                                   // MCUs per row
    int slot_comp[JP_MAX_SLOTS], slot_bx[JP_MAX_SLOTS], slot_by[JP_MAX_SLOTS];   // block s of an MCU: its component and place in it
    int comp_h[3], comp_v[3];                     // sampling factors (blocks of the component per MCU, across and down)
    int comp_bw[3];                               // blocks per row of the component's plane
    uint32_t comp_off[3];                         // first block of the plane in a file's coefficient array
    uint32_t blocks_per_file;                     // all planes
};
// where block number s of the scan (MCU by MCU, the blocks of an MCU in T.81 A.2.3 order) lives in the file's coefficient array
__device__ __forceinline__ uint32_t jp_dest(const JpGeom& g, uint32_t s)
{
    if (g.bpm == 1) return s;
    const uint32_t mcu = s / (uint32_t)g.bpm, slot = s - mcu * (uint32_t)g.bpm;
    const uint32_t my = mcu / (uint32_t)g.mcus_x, mx = mcu - my * (uint32_t)g.mcus_x;
    const int c = g.slot_comp[slot];
    return g.comp_off[c] + (my * g.comp_v[c] + g.slot_by[slot]) * g.comp_bw[c] + mx * g.comp_h[c] + g.slot_bx[slot];
}
struct JpInterval {
    uint32_t src, src_len;                        // bytes of the interval in the uploaded stream
    uint32_t scratch;                             // byte offset of its stripped copy (multiple of 4)
    uint32_t first_block, nblocks;                // blocks of its file, in scan order (whole MCUs)
    uint32_t file;
    uint32_t first_sub, nsub;                     // its subsequences of JP_SUB_BITS bits (the self-synchronising path)
};

__constant__ uint8_t c_natural_order[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// bit reader over the stripped bytes of an interval (a whole number of 32-bit words; zero bits past the end, as libjpeg pads)
struct JpBits {
    const uint32_t* words; uint32_t nwords, wi;
    unsigned long long buf; int nbits;            // `nbits` valid bits at the top of buf
    __device__ __forceinline__ void fill()
    {
        if (nbits <= 32) {
            const uint32_t w = wi < nwords ? words[wi] : 0u;
            wi++;
            buf |= (unsigned long long)__byte_perm(w, 0, 0x0123) << (32 - nbits);
            nbits += 32;
        }
    }
    __device__ __forceinline__ void seek(uint32_t bit) { wi = bit >> 5; buf = 0; nbits = 0; fill(); skip((int)(bit & 31)); }
    __device__ __forceinline__ uint32_t pos() const { return wi * 32u - (uint32_t)nbits; }      // bits consumed so far
    __device__ __forceinline__ uint32_t peek16() const { return (uint32_t)(buf >> 48); }
    __device__ __forceinline__ void skip(int n) { buf <<= n; nbits -= n; }
};

__device__ __forceinline__ int jp_decode(JpBits& b, const JpHuff* __restrict__ t)     // at least 16 bits buffered
{
    const uint32_t p = b.peek16();
    const uint32_t e = __ldg(&t->look[p >> (16 - JP_LOOK)]);
    if (e) { b.skip((int)(e >> 8)); return (int)(e & 255); }
    for (int l = JP_LOOK + 1; l <= 16; l++) {
        const int code = (int)(p >> (16 - l));
        if (code <= __ldg(&t->maxcode[l])) { b.skip(l); return __ldg(&t->vals[(code + __ldg(&t->valoff[l])) & 255]); }
    }
    b.skip(16);
    return 0;                                     // no such code: corrupt data; libjpeg substitutes a zero, too
}

// ---- K17a: one warp per interval copies it without the zero byte that follows every FF (T.81 B.1.1.5), pads the last word
__global__ void __launch_bounds__(JP_HUFF_THREADS)
k_jpeg_unstuff(const uint8_t* __restrict__ stream, const JpInterval* __restrict__ intervals, int nintervals, uint8_t* __restrict__ scratch,
               uint32_t* __restrict__ nwords_out)
{
    const int lane = threadIdx.x & 31;
    const int it = blockIdx.x * (JP_HUFF_THREADS / 32) + (threadIdx.x >> 5);
    if (it >= nintervals) return;
    const JpInterval iv = intervals[it];
    uint8_t* dst = scratch + iv.scratch;
    uint32_t out = 0;
    uint32_t prev_ff = 0;                         // was the last byte of the previous round FF?
    for (uint32_t i0 = 0; i0 < iv.src_len; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool in = i < iv.src_len;
        const uint32_t byte = in ? stream[iv.src + i] : 0u;
        uint32_t before = __shfl_up_sync(0xffffffffu, byte, 1);
        if (lane == 0) before = prev_ff ? 0xFFu : 0u;
        const bool keep = in && !(byte == 0 && before == 0xFF);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) dst[out + __popc(bal & ((1u << lane) - 1))] = (uint8_t)byte;
        out += __popc(bal);
        prev_ff = __shfl_sync(0xffffffffu, byte, 31) == 0xFF && (i0 + 31 < iv.src_len);
    }
    const uint32_t nwords = (out + 3) / 4;
    if (lane < 4 && out + lane < nwords * 4) dst[out + lane] = 0;
    if (lane == 0) nwords_out[it] = nwords;
}

// the same for long intervals (a whole file without restart markers): one CTA per interval, 4096 bytes per round, four
// consecutive bytes per thread
__global__ void __launch_bounds__(1024)
k_jpeg_unstuff_block(const uint8_t* __restrict__ stream, const JpInterval* __restrict__ intervals, uint8_t* __restrict__ scratch,
                     uint32_t* __restrict__ nwords_out)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const JpInterval iv = intervals[blockIdx.x];
    const uint8_t* src = stream + iv.src;
    uint8_t* dst = scratch + iv.scratch;
    uint32_t out = 0;
    for (uint32_t i0 = 0; i0 < iv.src_len; i0 += 4096) {
        const uint32_t i = i0 + threadIdx.x * 4;
        uint32_t byte[4], keep = 0, n = 0;
        uint32_t before = i && i <= iv.src_len ? src[i - 1] : 0u;       // the byte in front of this thread's four
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool in = i + k < iv.src_len;
            byte[k] = in ? src[i + k] : 0u;
            if (in && !(byte[k] == 0 && before == 0xFF)) { keep |= 1u << k; n++; }
            before = byte[k];
        }
        // exclusive scan of the per-thread counts over the CTA
        uint32_t incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = s_warp[lane];
            uint32_t wi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
            s_warp[lane] = wi - v;
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        uint32_t at = out + s_warp[warp] + incl - n;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (keep & (1u << k)) dst[at++] = (uint8_t)byte[k];
        out += s_total;
        __syncthreads();
    }
    const uint32_t nwords = (out + 3) / 4;
    if (threadIdx.x < 4 && out + threadIdx.x < nwords * 4) dst[out + threadIdx.x] = 0;
    if (threadIdx.x == 0) nwords_out[blockIdx.x] = nwords;
}

// the same with several CTAs per interval, when few long intervals (a batch of files without restart markers) would leave most SMs
// idle: whether a byte stays depends on its predecessor alone, so every JP_CHUNK source bytes are handled by a CTA of their own --
// a first kernel counts the bytes each chunk keeps, the second adds up the counts in front of its chunk and compacts it there.
constexpr uint32_t JP_CHUNK = 8192;
constexpr int JP_CHUNK_THREADS = 256;

// thread's four source bytes at offset i of the interval -> keep mask (bit k: byte k stays); `aligned`: 32-bit loads are possible
// (the uploaded stream is padded, so a word that starts inside the interval can be read whole)
__device__ __forceinline__ uint32_t jp_keep4(const uint8_t* __restrict__ src, uint32_t src_len, uint32_t i, bool aligned, uint32_t* byte)
{
    if (i >= src_len) { byte[0] = byte[1] = byte[2] = byte[3] = 0; return 0u; }
    if (aligned) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(src + i);
        byte[0] = w & 255u; byte[1] = (w >> 8) & 255u; byte[2] = (w >> 16) & 255u; byte[3] = w >> 24;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) byte[k] = i + k < src_len ? src[i + k] : 0u;
    }
    uint32_t before = i ? src[i - 1] : 0u, keep = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (i + k < src_len && !(byte[k] == 0 && before == 0xFF)) keep |= 1u << k;
        before = byte[k];
    }
    return keep;
}

__global__ void __launch_bounds__(JP_CHUNK_THREADS)
k_jpeg_unstuff_count(const uint8_t* __restrict__ stream, const JpInterval* __restrict__ intervals, uint32_t max_chunks, uint32_t* __restrict__ chunk_kept)
{
    __shared__ uint32_t s_warp[JP_CHUNK_THREADS / 32];
    const JpInterval iv = intervals[blockIdx.y];
    const uint32_t c0 = blockIdx.x * JP_CHUNK;
    if (c0 >= iv.src_len) return;                 // (the chunks an interval does not have are never read)
    const uint8_t* src = stream + iv.src;
    const bool aligned = (iv.src & 3u) == 0;
    uint32_t n = 0;
#pragma unroll
    for (uint32_t r = 0; r < JP_CHUNK / (JP_CHUNK_THREADS * 4); r++) {
        uint32_t byte[4];
        n += __popc(jp_keep4(src, iv.src_len, c0 + r * (JP_CHUNK_THREADS * 4) + threadIdx.x * 4, aligned, byte));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < JP_CHUNK_THREADS / 32; i++) t += s_warp[i];
        chunk_kept[(size_t)blockIdx.y * max_chunks + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(JP_CHUNK_THREADS)
k_jpeg_unstuff_chunk(const uint8_t* __restrict__ stream, const JpInterval* __restrict__ intervals, uint32_t max_chunks, const uint32_t* __restrict__ chunk_kept,
                     uint8_t* __restrict__ scratch, uint32_t* __restrict__ nwords_out)
{
    constexpr int NW = JP_CHUNK_THREADS / 32;
    __shared__ uint32_t s_warp[2][NW];
    __shared__ uint32_t s_base[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const JpInterval iv = intervals[blockIdx.y];
    const uint32_t c0 = blockIdx.x * JP_CHUNK;
    if (c0 >= iv.src_len && blockIdx.x) return;   // an empty interval still gets its word count (0) from chunk 0
    const uint8_t* src = stream + iv.src;
    uint8_t* dst = scratch + iv.scratch;
    const bool aligned = (iv.src & 3u) == 0;
    // bytes kept in front of this chunk
    uint32_t b = 0;
    for (uint32_t c = threadIdx.x; c < blockIdx.x; c += JP_CHUNK_THREADS) b += chunk_kept[(size_t)blockIdx.y * max_chunks + c];
#pragma unroll
    for (int o = 16; o; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0) s_base[warp] = b;
    __syncthreads();
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) out += s_base[i];
    for (uint32_t r = 0; r < JP_CHUNK / (JP_CHUNK_THREADS * 4); r++) {
        const uint32_t i = c0 + r * (JP_CHUNK_THREADS * 4) + threadIdx.x * 4;
        if (c0 + r * (JP_CHUNK_THREADS * 4) >= iv.src_len) break;      // (uniform over the CTA)
        uint32_t byte[4];
        const uint32_t keep = jp_keep4(src, iv.src_len, i, aligned, byte);
        const uint32_t n = __popc(keep);
        uint32_t incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) s_warp[r & 1][warp] = incl;
        __syncthreads();                          // (the buffers alternate: one barrier per round)
        uint32_t at = out + incl - n, total = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) { const uint32_t v = s_warp[r & 1][k]; total += v; at += k < warp ? v : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (keep & (1u << k)) dst[at++] = (uint8_t)byte[k];
        out += total;
    }
    if (c0 + JP_CHUNK >= iv.src_len) {            // the interval's last chunk: pad the last word, publish the word count
        const uint32_t nwords = (out + 3) / 4;
        if (threadIdx.x < 4 && out + threadIdx.x < nwords * 4) dst[out + threadIdx.x] = 0;
        if (threadIdx.x == 0) nwords_out[blockIdx.y] = nwords;
    }
}

// ---- the decoder proper.  State = (bit position, block of the MCU, coefficient index); decodes from `state` up to bit `end`.
// WRITE: stores coefficients, `blk` being the scan-order number of the block the state is in and dc0..2 the DC predictors.
__device__ __forceinline__ unsigned long long jp_state(uint32_t pos, int slot, int k) { return ((unsigned long long)pos << 16) | ((unsigned)slot << 8) | (unsigned)k; }

template <bool WRITE>
__device__ __forceinline__ unsigned long long jp_run(JpBits& b, const JpTables* __restrict__ tb, const JpGeom& g, const uint8_t* s_natural,
                                                     unsigned long long state, uint32_t end, int& nblk, int& dc0, int& dc1, int& dc2,
                                                     int16_t* coefs_file, uint32_t blk, uint32_t blocks_left)
{
    int k = (int)(state & 255), slot = (int)((state >> 8) & 255);
    b.seek((uint32_t)(state >> 16));
    int c = g.slot_comp[slot];
    int16_t* co = WRITE ? coefs_file + (size_t)jp_dest(g, blk) * 64 : nullptr;
    while (b.pos() < end && (!WRITE || blocks_left)) {
        // one refill per symbol: at least 33 bits are buffered, a code takes at most 16 and its value bits at most 15.  The body
        // is written with selects, not branches: the lanes of a warp are at different places of different blocks
        b.fill();
        const bool is_dc = k == 0;
        const int sym = jp_decode(b, is_dc ? &tb->dc[c] : &tb->ac[c]);
        const int r = is_dc ? 0 : sym >> 4, sz = sym & 15;
        const int raw = (int)(b.peek16() >> (16 - sz));                   // the sz bits after the code (0 for sz = 0)
        b.skip(sz);
        const int half = (1 << sz) >> 1;
        const int val = raw < half ? raw - (1 << sz) + 1 : raw;         // HUFF_EXTEND; 0 for sz = 0
        int d = c == 0 ? dc0 : (c == 1 ? dc1 : dc2);
        d += is_dc ? val : 0;
        dc0 = c == 0 ? d : dc0; dc1 = c == 1 ? d : dc1; dc2 = c == 2 ? d : dc2;
        const int kk = k + r;                                             // the coefficient this symbol sets (AC with sz != 0)
        if (WRITE) {
            const int store = is_dc ? d : val;
            const bool doit = is_dc ? d != 0 : (sz != 0 && kk <= 63);
            if (doit) co[is_dc ? 0 : s_natural[kk & 63]] = (int16_t)store;
        }
        // next coefficient index: after a DC 1; after a value kk + 1 (a run past 63 is corrupt data and ends the block, as the
        // reference's loop does); ZRL skips 16; end of block
        k = is_dc ? 1 : (sz ? (kk > 63 ? 64 : kk + 1) : (r == 15 ? k + 16 : 64));
        if (k >= 64) {
            k = 0; nblk++;
            slot = slot + 1 == g.bpm ? 0 : slot + 1;
            c = g.slot_comp[slot];
            if (WRITE) { blk++; blocks_left--; co = coefs_file + (size_t)jp_dest(g, blk) * 64; }
        }
    }
    return jp_state(b.pos(), slot, k);
}

// ---- K17b: one lane per interval runs the sequential decoder of T.81 F.2.2 over its stripped bytes.  The loop body decodes
// one symbol -- a DC difference when the lane is at the start of a block, else an AC run/size -- so the 32 intervals of a
// warp execute the same instructions whatever their data; only the rare codes longer than JP_LOOK bits diverge.
__global__ void __launch_bounds__(JP_HUFF_THREADS)
k_jpeg_huff(const JpInterval* __restrict__ intervals, int nintervals, const JpTables* __restrict__ tables, const int32_t* __restrict__ file_tables,
            const uint8_t* __restrict__ scratch, const uint32_t* __restrict__ nwords_in, int16_t* __restrict__ coefs, const JpGeom g, int lane_step)
{
    // every lane_step-th lane of a warp owns an interval: with few intervals spreading them over more warps costs issue slots but
    // shortens every warp's memory gathers; with many, all 32 lanes work
    __shared__ uint8_t s_natural[64];             // the lanes index it with different k: shared memory, not the constant cache
    if (threadIdx.x < 64) s_natural[threadIdx.x] = c_natural_order[threadIdx.x];
    __syncthreads();
    const long long t = (long long)blockIdx.x * JP_HUFF_THREADS + threadIdx.x;
    if (t % lane_step) return;
    const long long it = t / lane_step;
    if (it >= nintervals) return;
    const JpInterval iv = intervals[it];
    JpBits b = {reinterpret_cast<const uint32_t*>(scratch + iv.scratch), nwords_in[it], 0, 0ull, 0};
    int nblk = 0, dc0 = 0, dc1 = 0, dc2 = 0;
    jp_run<true>(b, tables + file_tables[iv.file], g, s_natural, jp_state(0, 0, 0), 0xffffffffu, nblk, dc0, dc1, dc2,
                 coefs + (size_t)iv.file * g.blocks_per_file * 64, iv.first_block, iv.nblocks);
}

// ---- K17c: the self-synchronising path for intervals longer than JP_SUB_BITS (whole files without restart markers, block rows):
// an interval is cut into subsequences of JP_SUB_BITS bits and every lane decodes ONE of them.  A lane does not know where the
// first code of its subsequence starts nor which coefficient of a block it belongs to, so it starts on a guess (its first bit, a
// DC code) -- Huffman streams re-synchronise quickly, so its exit state (bit position and coefficient index where the next
// subsequence begins) is usually right even then.  Every following round re-decodes a subsequence from its predecessor's exit
// state if that has changed; subsequence 0 starts from the truth, so the fixed point IS the sequential decode, and a round in
// which nothing changes proves it has been reached.  The per-subsequence block counts and DC sums then tell every lane which
// block and which DC value it starts with, and a last decode writes the coefficients.  (The idea of self-synchronising
// subsequences is from A. Weissenberger and B. Schmidt, "Massively parallel Huffman decoding on GPUs", 2018.)
// One CTA per interval, its lanes striding over the interval's subsequences.  Round 0: every subsequence from its own first bit.
// Rounds 1..: a subsequence whose predecessor's exit state is not the state it was last decoded from is decoded again.  A round
// with no such subsequence ends the loop (at the latest after as many rounds as there are subsequences: a perfectly periodic
// stream -- a blank image -- never re-synchronises by itself and is walked front to back).  Then thread 0 adds up the block counts
// and DC sums, and every lane decodes its subsequences once more, writing coefficients.
#ifdef JP_PROFILE
__device__ unsigned int g_jp_rounds[3];           // debug builds: intervals, sum and maximum of the rounds they took
#endif
// CL > 1: a thread-block cluster of CL CTAs shares one interval (whole files without restart markers in a small batch would
// otherwise occupy one SM each); the rounds are separated by cluster barriers and the "anything changed" flags are read through
// distributed shared memory.
template <int CL>
__global__ void __launch_bounds__(1024, 2)
k_jpeg_sync(const JpInterval* __restrict__ intervals, const JpTables* __restrict__ tables, const int32_t* __restrict__ file_tables,
            const uint8_t* __restrict__ scratch, const uint32_t* __restrict__ nwords_in, unsigned long long* exit_state,
            unsigned long long* used_state, int32_t* sub_counts /* [nsubs][4]: blocks, DC sums of the components */, uint32_t* work_list,
            int16_t* __restrict__ coefs, const JpGeom g, uint32_t sub_bits)
{
    namespace cg = cooperative_groups;
    __shared__ uint8_t s_natural[64];
    __shared__ int s_count;                       // subsequences this CTA decodes again in the current round
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_natural[i] = c_natural_order[i];      // (a CTA can be a single warp)
    const unsigned ivx = blockIdx.x / CL, rank = CL > 1 ? cg::this_cluster().block_rank() : 0u;
    auto barrier = [&]() { if (CL > 1) cg::this_cluster().sync(); else __syncthreads(); };
    const JpInterval iv = intervals[ivx];
    const JpTables* tb = tables + file_tables[iv.file];
    const uint32_t nwords = nwords_in[ivx], total = nwords * 32u;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(scratch + iv.scratch);
    volatile unsigned long long* ex = exit_state + iv.first_sub;
    unsigned long long* used = used_state + iv.first_sub;
    int4* cnt = reinterpret_cast<int4*>(sub_counts) + iv.first_sub;
    const uint32_t lane0 = rank * blockDim.x + threadIdx.x, stride = CL * blockDim.x;
    // this CTA's list of subsequences to decode again: it owns at most ceil(nsub / stride) * blockDim of them (the lists of an
    // interval take nsub + stride entries in all)
    uint32_t* list = work_list + iv.first_sub + (size_t)ivx * stride + (size_t)rank * ((iv.nsub + stride - 1) / stride) * blockDim.x;
    // round 0: every subsequence from its own first bit (subsequence 0: from the truth)
    for (uint32_t j = lane0; j < iv.nsub; j += stride) {
        const unsigned long long start = jp_state(j * sub_bits, 0, 0);
        JpBits b = {words, nwords, 0, 0ull, 0};
        int nblk = 0, d0 = 0, d1 = 0, d2 = 0;
        const unsigned long long e = jp_run<false>(b, tb, g, nullptr, start, min((j + 1) * sub_bits, total), nblk, d0, d1, d2, nullptr, 0, 0);
        used[j] = start;
        cnt[j] = make_int4(nblk, d0, d1, d2);
        ex[j] = e;
    }
    barrier();
    // rounds 1..: the subsequences whose predecessor's exit state is not the state they were last decoded from are collected into a
    // dense list first, so that a round costs what it has to decode, not a pass of every warp over mostly settled subsequences
    for (int round = 1;; round++) {
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        if (round > 1) {                          // (round 1 decodes nearly every subsequence again: no list)
            for (uint32_t j = lane0; j < iv.nsub; j += stride)
                if (j && ex[j - 1] != used[j]) list[atomicAdd(&s_count, 1)] = j;
        }
        barrier();
        const int n = round > 1 ? s_count : (int)((iv.nsub - min(iv.nsub, lane0 - threadIdx.x) + stride - 1) / stride) * (int)blockDim.x;
        int any = round > 1 ? n : 1;
        if (CL > 1 && round > 1)
            for (unsigned r = 0; r < (unsigned)CL; r++) any |= *cg::this_cluster().map_shared_rank(&s_count, r);
        if (!any) {                               // nothing left to decode again anywhere: the fixed point
#ifdef JP_PROFILE
            if (threadIdx.x == 0 && rank == 0) { atomicAdd(&g_jp_rounds[0], 1u); atomicAdd(&g_jp_rounds[1], (unsigned)round); atomicMax(&g_jp_rounds[2], (unsigned)round); }
#endif
            break;
        }
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            // round 1: the CTA's own subsequences in order, t-th of them = rank * blockDim + t + (t / blockDim) * (stride - blockDim)
            const uint32_t j = round > 1 ? list[t] : lane0 + (uint32_t)(t / (int)blockDim.x) * stride;
            if (j == 0 || j >= iv.nsub) continue;
            const unsigned long long start = ex[j - 1];
            if (start == used[j]) continue;
            JpBits b = {words, nwords, 0, 0ull, 0};
            int nblk = 0, d0 = 0, d1 = 0, d2 = 0;
            const unsigned long long e = jp_run<false>(b, tb, g, nullptr, start, min((j + 1) * sub_bits, total), nblk, d0, d1, d2, nullptr, 0, 0);
            used[j] = start;
            cnt[j] = make_int4(nblk, d0, d1, d2);
            if (e != ex[j]) ex[j] = e;
        }
        barrier();                                // (also: every CTA of the cluster has read the counts before they are reset)
    }
    // first block and DC predictors of every subsequence: an exclusive prefix sum, in place (L2 loads: other CTAs of the cluster
    // wrote some of the counts, and will read the sums)
    if (threadIdx.x == 0 && rank == 0) {
        int4 run = make_int4(0, 0, 0, 0);
        for (uint32_t j = 0; j < iv.nsub; j++) {
            const int4 v = __ldcg(cnt + j);
            cnt[j] = run;
            run.x += v.x; run.y += v.y; run.z += v.z; run.w += v.w;
        }
    }
    barrier();
    for (uint32_t j = lane0; j < iv.nsub; j += stride) {
        const int4 at = __ldcg(cnt + j);
        const uint32_t first = (uint32_t)at.x;
        if (first >= iv.nblocks) continue;
        JpBits b = {words, nwords, 0, 0ull, 0};
        int nblk = 0, d0 = at.y, d1 = at.z, d2 = at.w;
        // the lane that reaches the end of the data goes on over zero bits until the interval's blocks are complete, as libjpeg
        // (and the sequential kernel) do with a truncated file
        const uint32_t end = (j + 1) * sub_bits >= total ? 0xffffffffu : (j + 1) * sub_bits;
        jp_run<true>(b, tb, g, s_natural, j ? ex[j - 1] : jp_state(0, 0, 0), end, nblk, d0, d1, d2, coefs + (size_t)iv.file * g.blocks_per_file * 64,
                     iv.first_block + first, iv.nblocks - first);
    }
}

// ---- K18: jidctint.c jpeg_idct_islow
#define JP_FIX_0_298631336 2446
#define JP_FIX_0_390180644 3196
#define JP_FIX_0_541196100 4433
#define JP_FIX_0_765366865 6270
#define JP_FIX_0_899976223 7373
#define JP_FIX_1_175875602 9633
#define JP_FIX_1_501321110 12299
#define JP_FIX_1_847759065 15137
#define JP_FIX_1_961570560 16069
#define JP_FIX_2_053119869 16819
#define JP_FIX_2_562915447 20995
#define JP_FIX_3_072711026 25172

// one 8-point pass: in[0..7] -> out[0..7] before the final shift
__device__ __forceinline__ void jp_idct8(const int* in, int* o)
{
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * JP_FIX_0_541196100;
    int tmp2 = z1 + z3 * (-JP_FIX_1_847759065);
    int tmp3 = z1 + z2 * JP_FIX_0_765366865;
    int tmp0 = (int)((unsigned)(in[0] + in[4]) << 13);
    int tmp1 = (int)((unsigned)(in[0] - in[4]) << 13);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * JP_FIX_1_175875602;
    tmp0 *= JP_FIX_0_298631336; tmp1 *= JP_FIX_2_053119869; tmp2 *= JP_FIX_3_072711026; tmp3 *= JP_FIX_1_501321110;
    z1 *= -JP_FIX_0_899976223; z2 *= -JP_FIX_2_562915447; z3 *= -JP_FIX_1_961570560; z4 *= -JP_FIX_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    o[0] = tmp10 + tmp3; o[7] = tmp10 - tmp3;
    o[1] = tmp11 + tmp2; o[6] = tmp11 - tmp2;
    o[2] = tmp12 + tmp1; o[5] = tmp12 - tmp1;
    o[3] = tmp13 + tmp0; o[4] = tmp13 - tmp0;
}
// sample_range_limit + CENTERJSAMPLE indexed with & RANGE_MASK (jdmaster.c prepare_range_limit_table)
__device__ __forceinline__ uint32_t jp_range_limit(int x)
{
    x &= 1023;
    return x < 128 ? x + 128 : (x < 512 ? 255 : (x < 896 ? 0 : x - 896));
}

struct JpOut {                                    // where the planes go: the caller's frames (grey) or the internal planes (colour)
    uint8_t* base[3]; size_t pitch[3], stride[3];
    int cw[3], ch[3];                             // samples to write (the image size for grey frames, whole blocks for internal planes)
};

__global__ void __launch_bounds__(JP_IDCT_THREADS)
k_jpeg_idct(const int16_t* __restrict__ coefs, const JpTables* __restrict__ tables, const int32_t* __restrict__ file_tables, int nfiles, const JpGeom g,
            const JpOut o)
{
    const uint32_t blk = blockIdx.x * JP_IDCT_THREADS + threadIdx.x;
    const int file = blockIdx.y;
    if (blk >= g.blocks_per_file || file >= nfiles) return;
    const int c = g.ncomp == 1 || blk < g.comp_off[1] ? 0 : (blk < g.comp_off[2] ? 1 : 2);
    const uint16_t* q = tables[file_tables[file]].quant[c];
    const uint4* cp = reinterpret_cast<const uint4*>(coefs + ((size_t)file * g.blocks_per_file + blk) * 64);
    int ws[64];
#pragma unroll
    for (int r = 0; r < 8; r++) {                 // dequantise (row r of the block)
        const uint4 v = __ldg(cp + r);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
            const int coef = (int)(short)((cc & 1) ? (u[cc >> 1] >> 16) : (u[cc >> 1] & 0xffff));
            ws[r * 8 + cc] = coef * (int)__ldg(q + r * 8 + cc);
        }
    }
#pragma unroll
    for (int cc = 0; cc < 8; cc++) {              // pass 1: columns
        int in[8], ov[8];
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = ws[r * 8 + cc];
        jp_idct8(in, ov);
#pragma unroll
        for (int r = 0; r < 8; r++) ws[r * 8 + cc] = (ov[r] + (1 << 10)) >> 11;
    }
    const uint32_t i = blk - g.comp_off[c];
    const int bx = (int)(i % (uint32_t)g.comp_bw[c]), by = (int)(i / (uint32_t)g.comp_bw[c]);
    uint8_t* out = o.base[c] + (size_t)file * o.pitch[c] + (size_t)by * 8 * o.stride[c] + (size_t)bx * 8;
    const int cols = min(8, o.cw[c] - bx * 8), rows = min(8, o.ch[c] - by * 8);
    const bool aligned = cols == 8 && ((reinterpret_cast<uintptr_t>(out) | o.stride[c]) & 7) == 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {                 // pass 2: rows
        int ov[8];
        jp_idct8(ws + r * 8, ov);
        uint32_t px[8];
#pragma unroll
        for (int cc = 0; cc < 8; cc++) px[cc] = jp_range_limit((ov[cc] + (1 << 17)) >> 18);
        if (r < rows) {
            uint8_t* row = out + (size_t)r * o.stride[c];
            if (aligned) {
                *reinterpret_cast<uint2*>(row) = make_uint2(px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24), px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24));
            } else {
#pragma unroll
                for (int cc = 0; cc < 8; cc++)
                    if (cc < cols) row[cc] = (uint8_t)px[cc];
            }
        }
    }
}

// ---- K19: chroma upsampling (jdsample.c h2v2_fancy_upsample / h2v1_fancy_upsample) and YCbCr -> BGR (jdcolor.c
// ycc_rgb_convert, 16-bit fixed point); jp_chroma: one sample, the general form
__device__ __forceinline__ int jp_chroma(const uint8_t* __restrict__ pl, int stride, int cw, int ch, int x, int y, int hs, int vs)
{
    if (hs == 1) return pl[(size_t)y * stride + x];
    const int cx = x >> 1;
    // jdsample.c jinit_upsampler: the fancy filters only for components more than two samples wide, replication below that
    if (cw <= 2) return pl[(size_t)(vs == 2 ? y >> 1 : y) * stride + cx];
    if (vs == 1) {
        const uint8_t* r = pl + (size_t)y * stride;
        if (x & 1) return cx == cw - 1 ? r[cx] : (3 * r[cx] + r[cx + 1] + 2) >> 2;
        return cx == 0 ? r[0] : (3 * r[cx] + r[cx - 1] + 1) >> 2;
    }
    // the nearer row weighs 3, the row above (even output rows) or below (odd) 1 -- the first / last real row at the edges --
    // then the same across
    const int cy = y >> 1;
    const int oy = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
    const uint8_t* r0 = pl + (size_t)cy * stride;
    const uint8_t* r1 = pl + (size_t)oy * stride;
    const int cur = 3 * r0[cx] + r1[cx];
    if (x & 1) return cx == cw - 1 ? (cur * 4 + 7) >> 4 : (3 * cur + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
    return cx == 0 ? (cur * 4 + 8) >> 4 : (3 * cur + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

// one pixel of the colour conversion: jdcolor.c ycc_rgb_convert
__device__ __forceinline__ void jp_ycc_px(int Y, int cb, int cr, uint8_t* bgr)
{
    cb -= 128; cr -= 128;
    const int r = Y + ((91881 * cr + 32768) >> 16);           // FIX(1.40200), ONE_HALF
    const int b = Y + ((116130 * cb + 32768) >> 16);          // FIX(1.77200)
    const int gg = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);   // FIX(0.34414), FIX(0.71414)
    bgr[0] = (uint8_t)min(max(b, 0), 255); bgr[1] = (uint8_t)min(max(gg, 0), 255); bgr[2] = (uint8_t)min(max(r, 0), 255);
}

// sixteen upsampled chroma samples of row y from column x0 (a multiple of 16, x0 + 16 <= w) of a plane subsampled 2:1 across
// (HS == 2; down as well when vs == 2): the eight plane samples under them as one 8-byte load per row, their two neighbours as
// byte loads.  At the image's left and right edge jdsample.c's special cases equal the general formula with the missing
// neighbour replaced by the sample itself.
__device__ __forceinline__ void jp_chroma16(const uint8_t* __restrict__ pl, int stride, int cw, int ch, int x0, int y, int vs, int* out)
{
    const int cx0 = x0 >> 1;
    const int cy = vs == 2 ? y >> 1 : y;
    const uint8_t* r0 = pl + (size_t)cy * stride + cx0;
    int cur[10];                                  // columns cx0 - 1 .. cx0 + 8
    {
        const uint2 v = *reinterpret_cast<const uint2*>(r0);
#pragma unroll
        for (int i = 0; i < 8; i++) cur[1 + i] = (int)(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 255u);
        cur[0] = cx0 > 0 ? (int)r0[-1] : cur[1];
        cur[9] = cx0 + 8 <= cw - 1 ? (int)r0[8] : cur[8];
    }
    if (vs == 2) {                                // the nearer row weighs 3, the other one 1
        const int oy = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
        const uint8_t* r1 = pl + (size_t)oy * stride + cx0;
        const uint2 v = *reinterpret_cast<const uint2*>(r1);
        int far_[10];
#pragma unroll
        for (int i = 0; i < 8; i++) far_[1 + i] = (int)(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 255u);
        far_[0] = cx0 > 0 ? (int)r1[-1] : far_[1];
        far_[9] = cx0 + 8 <= cw - 1 ? (int)r1[8] : far_[8];
#pragma unroll
        for (int i = 0; i < 10; i++) cur[i] = 3 * cur[i] + far_[i];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            out[2 * i] = (3 * cur[1 + i] + cur[i] + 8) >> 4;
            out[2 * i + 1] = (3 * cur[1 + i] + cur[2 + i] + 7) >> 4;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            out[2 * i] = (3 * cur[1 + i] + cur[i] + 1) >> 2;
            out[2 * i + 1] = (3 * cur[1 + i] + cur[2 + i] + 2) >> 2;
        }
    }
}

// sixteen pixels of a row per thread: the luma and chroma samples arrive as 8-byte loads, the 48 bytes of B, G, R leave as three
// 16-byte stores; the threads at the right edge of the image (and every thread of a frame whose rows are not 16-byte aligned) take
// the samples one by one
template <int HS>
__global__ void __launch_bounds__(128)
k_jpeg_ycc(const uint8_t* __restrict__ planes, size_t planes_pitch, size_t off1, size_t off2, int stride0, int stride1, int w, int h, int vs,
           uint8_t* __restrict__ frames, size_t frame_pitch, size_t stride)
{
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 16, y = blockIdx.y, file = blockIdx.z;
    if (x0 >= w) return;
    const uint8_t* p0 = planes + (size_t)file * planes_pitch;
    const uint8_t* p1 = p0 + off1;
    const uint8_t* p2 = p0 + off2;
    const int cw = (w + HS - 1) / HS, ch = (h + vs - 1) / vs;
    uint8_t* out = frames + (size_t)file * frame_pitch + (size_t)y * stride + (size_t)x0 * 3;
    if (x0 + 16 > w || (reinterpret_cast<uintptr_t>(out) & 15) != 0 || (HS == 2 && cw <= 2)) {
        for (int i = 0; i < 16 && x0 + i < w; i++) {
            const int x = x0 + i;
            jp_ycc_px(p0[(size_t)y * stride0 + x], jp_chroma(p1, stride1, cw, ch, x, y, HS, vs), jp_chroma(p2, stride1, cw, ch, x, y, HS, vs), out + 3 * i);
        }
        return;
    }
    int cb[16], cr[16];
    if (HS == 2) {
        jp_chroma16(p1, stride1, cw, ch, x0, y, vs, cb);
        jp_chroma16(p2, stride1, cw, ch, x0, y, vs, cr);
    } else {
        const uint2* q1 = reinterpret_cast<const uint2*>(p1 + (size_t)y * stride1 + x0);
        const uint2* q2 = reinterpret_cast<const uint2*>(p2 + (size_t)y * stride1 + x0);
        const uint2 a0 = q1[0], a1 = q1[1], b0 = q2[0], b1 = q2[1];
        const uint32_t aw[4] = {a0.x, a0.y, a1.x, a1.y}, bw[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
        for (int i = 0; i < 16; i++) { cb[i] = (int)((aw[i >> 2] >> (8 * (i & 3))) & 255u); cr[i] = (int)((bw[i >> 2] >> (8 * (i & 3))) & 255u); }
    }
    const uint2* qy = reinterpret_cast<const uint2*>(p0 + (size_t)y * stride0 + x0);
    const uint2 y0 = qy[0], y1 = qy[1];
    const uint32_t yw[4] = {y0.x, y0.y, y1.x, y1.y};
    uint8_t bgr[48];
#pragma unroll
    for (int i = 0; i < 16; i++) jp_ycc_px((int)((yw[i >> 2] >> (8 * (i & 3))) & 255u), cb[i], cr[i], bgr + 3 * i);
    uint32_t ow[12];
#pragma unroll
    for (int i = 0; i < 12; i++) ow[i] = bgr[4 * i] | (bgr[4 * i + 1] << 8) | (bgr[4 * i + 2] << 16) | ((uint32_t)bgr[4 * i + 3] << 24);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    o4[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    o4[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
    o4[2] = make_uint4(ow[8], ow[9], ow[10], ow[11]);
}

// ---- host: marker parsing
struct HostHuff { bool present; uint8_t bits[17]; uint8_t vals[256]; };
struct HostFile {
    int width, height, restart;                   // restart interval in MCUs
    int ncomp, hs, vs;                            // components; luma sampling factors (chroma is 1x1)
    uint16_t quant[3][64];                        // per component
    HostHuff dc[3], ac[3];
    size_t scan, scan_len;                        // entropy-coded segment
};

static bool build_table(const HostHuff& h, JpHuff& t)
{
    memset(&t, 0, sizeof(t));
    memcpy(t.vals, h.vals, 256);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        t.valoff[l] = k - code;
        for (int i = 0; i < h.bits[l]; i++, code++, k++) {
            if (code >= (1 << l) || k >= 256) return false;        // over-subscribed table
            if (l <= JP_LOOK)
                for (int fill = 0; fill < (1 << (JP_LOOK - l)); fill++) t.look[(code << (JP_LOOK - l)) | fill] = (uint16_t)((l << 8) | h.vals[k]);
        }
        t.maxcode[l] = h.bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    t.maxcode[17] = 0x7fffffff;
    return true;
}

// returns ORBX_OK, ORBX_E_UNSUPPORTED (a valid JPEG of another kind) or ORBX_E_INVALID (not a JPEG / damaged headers)
static int parse_jpeg(const uint8_t* f, size_t n, HostFile& out, const char** why)
{
    static const uint8_t natural[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                        35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
#define JP_FAIL(code, msg) do { *why = msg; return code; } while (0)
    if (n < 4 || f[0] != 0xFF || f[1] != 0xD8) JP_FAIL(ORBX_E_INVALID, "not a JPEG file (no SOI)");
    uint16_t quant[4][64];
    bool have_quant[4] = {false, false, false, false}, have_sof = false;
    HostHuff dc[4], ac[4];
    for (int i = 0; i < 4; i++) dc[i].present = ac[i].present = false;
    int tq[3] = {0, 0, 0}, cid[3] = {0, 0, 0};
    out.restart = 0;
    size_t p = 2;
    for (;;) {
        if (p + 4 > n || f[p] != 0xFF) JP_FAIL(ORBX_E_INVALID, "damaged marker structure");
        while (p < n && f[p] == 0xFF) p++;
        if (p >= n) JP_FAIL(ORBX_E_INVALID, "damaged marker structure");
        const int m = f[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) JP_FAIL(ORBX_E_INVALID, "end of image before the scan");
        if (p + 2 > n) JP_FAIL(ORBX_E_INVALID, "truncated header");
        const size_t len = ((size_t)f[p] << 8) | f[p + 1];
        if (len < 2 || p + len > n) JP_FAIL(ORBX_E_INVALID, "truncated header");
        const uint8_t* s = f + p + 2;
        const size_t sl = len - 2;
        if (m == 0xC0 || m == 0xC1) {
            if (sl < 6) JP_FAIL(ORBX_E_INVALID, "short SOF");
            if (s[0] != 8) JP_FAIL(ORBX_E_UNSUPPORTED, "sample precision is not 8 bits");
            out.height = (s[1] << 8) | s[2];
            out.width = (s[3] << 8) | s[4];
            out.ncomp = s[5];
            if (out.ncomp != 1 && out.ncomp != 3) JP_FAIL(ORBX_E_UNSUPPORTED, "neither a grey-scale nor a three-component file");
            if (sl < (size_t)(6 + 3 * out.ncomp) || out.width == 0 || out.height == 0) JP_FAIL(ORBX_E_INVALID, "bad SOF");
            for (int c = 0; c < out.ncomp; c++) {
                cid[c] = s[6 + 3 * c];
                tq[c] = s[8 + 3 * c] & 3;
                const int hh_ = s[7 + 3 * c] >> 4, vv_ = s[7 + 3 * c] & 15;
                if (c == 0) { out.hs = hh_; out.vs = vv_; }
                else if (hh_ != 1 || vv_ != 1) JP_FAIL(ORBX_E_UNSUPPORTED, "chroma sampling other than 1x1");
            }
            if (out.ncomp == 1) out.hs = out.vs = 1;       // a single-component scan is not interleaved: one block per MCU
            else if (!((out.hs == 2 && out.vs == 2) || (out.hs == 2 && out.vs == 1) || (out.hs == 1 && out.vs == 1)))
                JP_FAIL(ORBX_E_UNSUPPORTED, "luma sampling other than 2x2, 2x1 or 1x1");
            have_sof = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            JP_FAIL(ORBX_E_UNSUPPORTED, "progressive, lossless, hierarchical or arithmetic-coded file");
        } else if (m == 0xC4) {
            size_t q = 0;
            while (q < sl) {
                if (q + 17 > sl) JP_FAIL(ORBX_E_INVALID, "short DHT");
                const int tc = s[q] >> 4, th = s[q] & 15;
                if (tc > 1 || th > 3) JP_FAIL(ORBX_E_INVALID, "bad DHT selector");
                HostHuff& t = tc ? ac[th] : dc[th];
                int count = 0;
                t.bits[0] = 0;
                for (int l = 1; l <= 16; l++) { t.bits[l] = s[q + l]; count += t.bits[l]; }
                q += 17;
                if (count > 256 || q + (size_t)count > sl) JP_FAIL(ORBX_E_INVALID, "bad DHT counts");
                memset(t.vals, 0, sizeof(t.vals));
                memcpy(t.vals, s + q, (size_t)count);
                q += (size_t)count;
                t.present = true;
            }
        } else if (m == 0xDB) {
            size_t q = 0;
            while (q < sl) {
                const int pq = s[q] >> 4, t = s[q] & 15;
                if (t > 3 || pq > 1) JP_FAIL(ORBX_E_INVALID, "bad DQT selector");
                q++;
                if (q + (size_t)(pq ? 128 : 64) > sl) JP_FAIL(ORBX_E_INVALID, "short DQT");
                for (int i = 0; i < 64; i++) {
                    quant[t][natural[i]] = (uint16_t)(pq ? ((s[q] << 8) | s[q + 1]) : s[q]);
                    q += pq ? 2 : 1;
                }
                have_quant[t] = true;
            }
        } else if (m == 0xDD) {
            if (sl < 2) JP_FAIL(ORBX_E_INVALID, "short DRI");
            out.restart = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {
            if (!have_sof) JP_FAIL(ORBX_E_INVALID, "scan before the frame header");
            if (sl < (size_t)(4 + 2 * out.ncomp) || s[0] != out.ncomp) JP_FAIL(ORBX_E_UNSUPPORTED, "a scan that does not hold every component");
            const uint8_t* e = s + 1 + 2 * out.ncomp;
            if (e[0] != 0 || e[1] != 63 || e[2] != 0) JP_FAIL(ORBX_E_UNSUPPORTED, "not a sequential scan");
            for (int c = 0; c < out.ncomp; c++) {
                if (s[1 + 2 * c] != cid[c]) JP_FAIL(ORBX_E_UNSUPPORTED, "scan components out of frame order");
                const int td = s[2 + 2 * c] >> 4, ta = s[2 + 2 * c] & 15;
                if (td > 3 || ta > 3 || !dc[td].present || !ac[ta].present || !have_quant[tq[c]]) JP_FAIL(ORBX_E_INVALID, "scan refers to a missing table");
                out.dc[c] = dc[td]; out.ac[c] = ac[ta];
                memcpy(out.quant[c], quant[tq[c]], sizeof(out.quant[c]));
            }
            for (int c = out.ncomp; c < 3; c++) { out.dc[c] = out.dc[0]; out.ac[c] = out.ac[0]; memcpy(out.quant[c], out.quant[0], sizeof(out.quant[c])); }
            out.scan = p + len;
            out.scan_len = n - out.scan;
            return ORBX_OK;
        }
        p += len;
    }
#undef JP_FAIL
}

}  // namespace
}  // namespace orbx

using namespace orbx;

struct JpSet {                                    // what one call uploads: pinned staging + its device copy
    uint8_t* h_stream; size_t h_stream_bytes;     // the entropy-coded segments of a batch, one after the other
    uint8_t* d_stream; size_t d_stream_bytes;
    JpInterval* h_intervals; size_t h_intervals_n;
    JpInterval* d_intervals; size_t d_intervals_bytes;
    JpTables* h_tables; JpTables* d_tables; int tables_cap;
    int32_t* h_file_tables; size_t h_file_tables_n;
    int32_t* d_file_tables; size_t d_file_tables_bytes;
    cudaEvent_t uploaded;                         // on the copy stream: the uploads have left the pinned buffers
    cudaEvent_t decoded;                          // on the caller's stream: the kernels that read the device copy are done
    bool uploaded_pending, decoded_pending;
};

struct jpgx_context {
    int device;
    cudaStream_t own_stream, stream;
    int sm_count;
    cudaStream_t copy_stream;                     // uploads of call k+1 run beside the kernels of call k
    JpSet set[2];                                 // the buffers of two calls in flight, used in turn
    int next_set;
    uint8_t* d_scratch; size_t d_scratch_bytes;   // the same without stuffed bytes
    uint32_t* d_nwords; size_t d_nwords_bytes;    // 32-bit words of every interval's stripped copy
    uint32_t* d_chunk_kept; size_t d_chunk_kept_bytes;   // chunked unstuffing: bytes every chunk keeps, [interval][chunk]
    unsigned long long* d_subw; size_t d_subw_bytes;   // per subsequence: exit state, state it was last decoded from, block count and DC sums
    uint8_t* d_planes; size_t d_planes_bytes;     // colour files: the component planes before upsampling
    int16_t* d_coefs; size_t d_coefs_bytes;
    uint8_t* d_frames; size_t d_frames_bytes;     // staging of the host-output form
};

template <typename T>
static int jp_grow(T** p, size_t* have, size_t want)
{
    if (*p && *have >= want) return ORBX_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *have = 0; }
    const size_t bytes = align_up(want + want / 4, 256);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e != cudaSuccess) { *p = nullptr; set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    *have = bytes;
    return ORBX_OK;
}
// write_combined: for buffers the CPU only ever writes front to back (the DMA engine then reads memory that is not sitting
// dirty in a cache: measured 19 MB in 0.35 ms instead of 1.1 ms)
template <typename T>
static int jp_grow_host(T** p, size_t* have_n, size_t want_n, bool write_combined = false)
{
    if (*p && *have_n >= want_n) return ORBX_OK;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *have_n = 0; }
    const size_t n = want_n + want_n / 4 + 64;
    cudaError_t e = cudaHostAlloc((void**)p, n * sizeof(T), write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) { *p = nullptr; set_error("cudaMallocHost(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return ORBX_E_ALLOC; }
    *have_n = n;
    return ORBX_OK;
}

extern "C" int jpgx_destroy(jpgx_handle h)
{
    if (!h) return ORBX_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    for (JpSet& S : h->set) {
        cudaFree(S.d_stream); cudaFree(S.d_intervals); cudaFree(S.d_tables); cudaFree(S.d_file_tables);
        if (S.h_stream) cudaFreeHost(S.h_stream);
        if (S.h_intervals) cudaFreeHost(S.h_intervals);
        if (S.h_tables) cudaFreeHost(S.h_tables);
        if (S.h_file_tables) cudaFreeHost(S.h_file_tables);
        if (S.uploaded) cudaEventDestroy(S.uploaded);
        if (S.decoded) cudaEventDestroy(S.decoded);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    cudaFree(h->d_scratch); cudaFree(h->d_nwords); cudaFree(h->d_chunk_kept); cudaFree(h->d_subw); cudaFree(h->d_planes);
    cudaFree(h->d_coefs); cudaFree(h->d_frames);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return ORBX_OK;
}

extern "C" int jpgx_create(jpgx_handle* out, int device)
{
    ORBX_REQUIRE(out != nullptr, "jpgx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_error("jpgx_create: no CUDA device (%s); liborbx has no CPU fallback", cudaGetErrorString(e)); return ORBX_E_CUDA; }
    ORBX_REQUIRE(device >= 0 && device < ndev, "jpgx_create: device %d out of range [0,%d)", device, ndev);
    ORBX_CUDA(cudaSetDevice(device));
    jpgx_context* h = new jpgx_context();
    memset(h, 0, sizeof(*h));
    h->device = device;
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking), jpgx_destroy(h));
    h->stream = h->own_stream;
    ORBX_CUDA_OR(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking), jpgx_destroy(h));
    ORBX_CUDA_OR(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device), jpgx_destroy(h));
    for (JpSet& S : h->set) {
        S.tables_cap = 16;
        ORBX_CUDA_OR(cudaEventCreateWithFlags(&S.uploaded, cudaEventDisableTiming), jpgx_destroy(h));
        ORBX_CUDA_OR(cudaEventCreateWithFlags(&S.decoded, cudaEventDisableTiming), jpgx_destroy(h));
        ORBX_CUDA_OR(cudaMallocHost((void**)&S.h_tables, sizeof(JpTables) * (size_t)S.tables_cap), jpgx_destroy(h));
        ORBX_CUDA_OR(cudaMalloc((void**)&S.d_tables, sizeof(JpTables) * (size_t)S.tables_cap), jpgx_destroy(h));
    }
    *out = h;
    return ORBX_OK;
}

extern "C" int jpgx_set_stream(jpgx_handle h, void* cuda_stream)
{
    ORBX_REQUIRE(h != nullptr, "jpgx_set_stream: NULL handle");
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return ORBX_OK;
}

extern "C" int jpgx_get_stream(jpgx_handle h, void** cuda_stream)
{
    ORBX_REQUIRE(h != nullptr && cuda_stream != nullptr, "jpgx_get_stream: NULL argument");
    *cuda_stream = h->stream == h->own_stream ? nullptr : (void*)h->stream;
    return ORBX_OK;
}

extern "C" int jpgx_synchronize(jpgx_handle h)
{
    ORBX_REQUIRE(h != nullptr, "jpgx_synchronize: NULL handle");
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int jpgx_probe(const uint8_t* file, size_t size, int32_t* info)
{
    ORBX_REQUIRE(file != nullptr && info != nullptr, "jpgx_probe: NULL argument");
    HostFile hf;
    const char* why = "";
    const int rc = parse_jpeg(file, size, hf, &why);
    if (rc) { set_error("jpgx_probe: %s", why); return rc; }
    const int mx = (hf.width + 8 * hf.hs - 1) / (8 * hf.hs), my = (hf.height + 8 * hf.vs - 1) / (8 * hf.vs);
    info[0] = hf.width; info[1] = hf.height; info[2] = hf.restart;
    info[3] = mx * my * (hf.ncomp == 1 ? 1 : hf.hs * hf.vs + 2);
    info[4] = hf.ncomp; info[5] = hf.hs * 16 + hf.vs;
    return ORBX_OK;
}

// ncomp: 1 = grey files into one-byte pixels, 3 = YCbCr files into BGR pixels
static int decode_batch_dev(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int hh, uint8_t* d_frames,
                            size_t frame_pitch, size_t stride, int ncomp)
{
    ORBX_REQUIRE(h != nullptr, "jpgx_decode_batch: NULL handle");
    ORBX_REQUIRE(nfiles >= 0 && w >= 1 && hh >= 1 && w <= 65535 && hh <= 65535, "jpgx_decode_batch: nfiles %d / size %dx%d out of range", nfiles, w, hh);
    if (nfiles == 0) return ORBX_OK;
    ORBX_REQUIRE(files && sizes && d_frames, "jpgx_decode_batch: NULL pointer");
    ORBX_REQUIRE(stride >= (size_t)w * ncomp && frame_pitch >= stride * (size_t)(hh - 1) + (size_t)w * ncomp, "jpgx_decode_batch: stride %zu / frame pitch %zu too small for %dx%d x %d", stride, frame_pitch, w, hh, ncomp);
    ORBX_CUDA(cudaSetDevice(h->device));
    // this call's buffer set was last used two calls ago: its uploads must have left the pinned buffers before they are
    // overwritten (the device copy is protected in stream order, below)
    JpSet& S = h->set[h->next_set];
    h->next_set ^= 1;
    if (S.uploaded_pending) { ORBX_CUDA(cudaEventSynchronize(S.uploaded)); S.uploaded_pending = false; }
    // 1. headers: where every file's bytes, intervals and stripped copy go follows from them alone
    JpGeom g;
    memset(&g, 0, sizeof(g));
    int mcus_y = 0;
    uint32_t blocks = 0;
    std::vector<HostFile> hf((size_t)nfiles);
    std::vector<size_t> off((size_t)nfiles + 1, 0), first_iv((size_t)nfiles + 1, 0), scr((size_t)nfiles + 1, 0);
    for (int i = 0; i < nfiles; i++) {
        ORBX_REQUIRE(files[i] != nullptr, "jpgx_decode_batch: file %d is NULL", i);
        const char* why = "";
        const int rc = parse_jpeg(files[i], sizes[i], hf[(size_t)i], &why);
        if (rc) { set_error("jpgx_decode_batch: file %d: %s", i, why); return rc; }
        ORBX_REQUIRE(hf[(size_t)i].width == w && hf[(size_t)i].height == hh, "jpgx_decode_batch: file %d is %dx%d, the batch is %dx%d", i,
                     hf[(size_t)i].width, hf[(size_t)i].height, w, hh);
        if (hf[(size_t)i].ncomp != ncomp) {
            set_error("jpgx_decode_batch: file %d has %d component(s); %s", i, hf[(size_t)i].ncomp,
                      ncomp == 1 ? "colour files go through jpgx_decode_bgr_batch" : "grey-scale files go through jpgx_decode_gray_batch");
            return ORBX_E_UNSUPPORTED;
        }
        ORBX_REQUIRE(hf[(size_t)i].hs == hf[0].hs && hf[(size_t)i].vs == hf[0].vs, "jpgx_decode_batch: file %d has another chroma sampling than file 0", i);
        if (i == 0) {
            // the geometry every file of the batch shares
            const HostFile& f0 = hf[0];
            g.ncomp = ncomp;
            g.bpm = ncomp == 1 ? 1 : f0.hs * f0.vs + 2;
            g.mcus_x = (w + 8 * f0.hs - 1) / (8 * f0.hs);
            mcus_y = (hh + 8 * f0.vs - 1) / (8 * f0.vs);
            int slot = 0;
            uint32_t offb = 0;
            for (int c = 0; c < 3; c++) {
                g.comp_h[c] = c == 0 ? f0.hs : 1; g.comp_v[c] = c == 0 ? f0.vs : 1;
                g.comp_bw[c] = g.mcus_x * g.comp_h[c];
                g.comp_off[c] = offb;
                if (c < ncomp) {
                    offb += (uint32_t)g.comp_bw[c] * (uint32_t)(mcus_y * g.comp_v[c]);
                    for (int by = 0; by < g.comp_v[c]; by++)
                        for (int bx = 0; bx < g.comp_h[c]; bx++, slot++) { g.slot_comp[slot] = c; g.slot_bx[slot] = bx; g.slot_by[slot] = by; }
                }
            }
            for (; slot < JP_MAX_SLOTS; slot++) { g.slot_comp[slot] = 0; g.slot_bx[slot] = g.slot_by[slot] = 0; }
            g.blocks_per_file = offb;
            blocks = (uint32_t)g.mcus_x * (uint32_t)mcus_y * (uint32_t)g.bpm;         // == offb
        }
        const uint32_t nmcus = blocks / (uint32_t)g.bpm;
        const uint32_t ri = hf[(size_t)i].restart ? (uint32_t)hf[(size_t)i].restart : nmcus;
        const size_t niv = (nmcus + ri - 1) / ri;
        off[(size_t)i + 1] = off[(size_t)i] + align_up(hf[(size_t)i].scan_len, 4);
        first_iv[(size_t)i + 1] = first_iv[(size_t)i] + niv;
        scr[(size_t)i + 1] = scr[(size_t)i] + align_up(hf[(size_t)i].scan_len, 4) + 12 * niv + 16;      // interval j at align4(its start) + 12 j
    }
    const size_t total = off[(size_t)nfiles], nint = first_iv[(size_t)nfiles];
    ORBX_REQUIRE(scr[(size_t)nfiles] < (1ull << 32), "jpgx_decode_batch: %zu bytes of compressed data in one batch", total);
    int rc = jp_grow_host(&S.h_stream, &S.h_stream_bytes, total + 64, true);
    if (!rc) rc = jp_grow_host(&S.h_intervals, &S.h_intervals_n, nint);
    if (!rc) rc = jp_grow_host(&S.h_file_tables, &S.h_file_tables_n, (size_t)nfiles);
    if (!rc) rc = jp_grow(&S.d_stream, &S.d_stream_bytes, total + 64);
    if (!rc) rc = jp_grow(&h->d_scratch, &h->d_scratch_bytes, scr[(size_t)nfiles] + 64);
    if (!rc) rc = jp_grow(&S.d_intervals, &S.d_intervals_bytes, nint * sizeof(JpInterval));
    if (!rc) rc = jp_grow(&h->d_nwords, &h->d_nwords_bytes, nint * sizeof(uint32_t));
    if (!rc) rc = jp_grow(&S.d_file_tables, &S.d_file_tables_bytes, (size_t)nfiles * sizeof(int32_t));
    if (!rc) rc = jp_grow(&h->d_coefs, &h->d_coefs_bytes, (size_t)nfiles * blocks * 64 * sizeof(int16_t));
    if (rc) return rc;
    // 2a. table sets, shared by the files that carry the same ones
    std::vector<JpTables> sets;
    std::vector<const HostFile*> set_owner;
    for (int i = 0; i < nfiles; i++) {
        const HostFile& f = hf[(size_t)i];
        int set = -1;
        for (size_t s = 0; s < set_owner.size(); s++) {
            const HostFile& o = *set_owner[s];
            bool same = !memcmp(o.quant, f.quant, sizeof(f.quant));
            for (int c = 0; c < 3 && same; c++)
                same = !memcmp(o.dc[c].bits, f.dc[c].bits, 17) && !memcmp(o.dc[c].vals, f.dc[c].vals, 256) && !memcmp(o.ac[c].bits, f.ac[c].bits, 17) &&
                       !memcmp(o.ac[c].vals, f.ac[c].vals, 256);
            if (same) { set = (int)s; break; }
        }
        if (set < 0) {
            sets.emplace_back();
            JpTables& t = sets.back();
            for (int c = 0; c < 3; c++)
                if (!build_table(f.dc[c], t.dc[c]) || !build_table(f.ac[c], t.ac[c])) { set_error("jpgx_decode_batch: file %d: over-subscribed Huffman table", i); return ORBX_E_INVALID; }
            memcpy(t.quant, f.quant, sizeof(t.quant));
            set = (int)sets.size() - 1;
            set_owner.push_back(&f);
        }
        S.h_file_tables[i] = set;
    }
    // 2b. per file, independent of the others (host threads when there is enough to copy): its scan into the pinned stream,
    // cut at its restart markers
    auto stage_file = [&](int i) {
        const HostFile& f = hf[(size_t)i];
        const uint8_t* scan = files[i] + f.scan;
        memcpy(S.h_stream + off[(size_t)i], scan, f.scan_len);
        const uint32_t nmcus = blocks / (uint32_t)g.bpm;
        const uint32_t ri = f.restart ? (uint32_t)f.restart : nmcus;
        size_t p = 0, j = 0;
        for (uint32_t m0 = 0; m0 < nmcus; m0 += ri, j++) {
            // the interval ends at the next marker: FF followed by anything but 00 (RSTn between intervals, EOI after the last)
            size_t q = p;
            for (;;) {
                const uint8_t* hit = q < f.scan_len ? (const uint8_t*)memchr(scan + q, 0xFF, f.scan_len - q) : nullptr;
                if (!hit) { q = f.scan_len; break; }
                q = (size_t)(hit - scan);
                if (q + 1 >= f.scan_len) { q = f.scan_len; break; }
                if (scan[q + 1] != 0x00) break;
                q += 2;
            }
            JpInterval& iv = S.h_intervals[first_iv[(size_t)i] + j];
            iv.src = (uint32_t)(off[(size_t)i] + p);
            iv.src_len = (uint32_t)(q - p);
            iv.scratch = (uint32_t)(scr[(size_t)i] + align_up(p, 4) + 12 * j);
            iv.first_block = m0 * (uint32_t)g.bpm;
            iv.nblocks = std::min(ri, nmcus - m0) * (uint32_t)g.bpm;
            iv.file = (uint32_t)i;
            p = q;
            if (p + 1 < f.scan_len) {              // skip the marker
                while (p < f.scan_len && scan[p] == 0xFF) p++;
                if (p < f.scan_len) p++;
            }
        }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int nthreads = total < (2u << 20) ? 1 : (int)std::min<unsigned>(std::min<unsigned>(hw, 8u), (unsigned)nfiles);
    if (nthreads <= 1) {
        for (int i = 0; i < nfiles; i++) stage_file(i);
    } else {
        std::atomic<int> next(0);
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++)
            pool.emplace_back([&] { for (int i = next.fetch_add(1); i < nfiles; i = next.fetch_add(1)) stage_file(i); });
        for (std::thread& t : pool) t.join();
    }
    const size_t ni = nint;
    // subsequences of the self-synchronising path: ceil(bits / sub_bits) per interval (from the stuffed length: a few may be
    // empty).  JP_SUB_BITS bits each, fewer when the longest interval would then not fill a warp (block rows: ~18 x 1024 bits)
    uint32_t longest = 0;
    for (size_t i = 0; i < ni; i++) longest = std::max(longest, S.h_intervals[i].src_len);
    uint32_t sub_bits = JP_SUB_BITS;
    if ((size_t)longest * 8 < 32u * JP_SUB_BITS) sub_bits = std::max(256u, (uint32_t)(((size_t)longest * 8 + 31) / 32 + 31) / 32 * 32);
    size_t nsubs = 0;
    uint32_t max_nsub = 1;
    bool packable = true;
    for (size_t i = 0; i < ni; i++) {
        JpInterval& iv = S.h_intervals[i];
        iv.nsub = std::max<uint32_t>(1u, (uint32_t)(((size_t)iv.src_len * 8 + sub_bits - 1) / sub_bits));
        iv.first_sub = (uint32_t)nsubs;
        nsubs += iv.nsub;
        max_nsub = std::max(max_nsub, iv.nsub);
        packable = packable && (size_t)iv.src_len * 8 + 64 < JP_MAX_INTERVAL_BITS;      // (bit positions are 32-bit)
    }
    // short intervals (a restart marker every few blocks) are decoded one lane each, sequentially: nothing to synchronise
    const bool sync_path = (size_t)longest * 8 > 4 * JP_SUB_BITS && packable;
    // a few very long intervals (a small batch of files without restart markers): a cluster of CTAs per interval, so that the
    // batch covers the SMs
    int cl = 1;
    while (cl < 8 && ni * (size_t)(cl * 2) <= 2 * (size_t)h->sm_count && max_nsub >= 1024u * (unsigned)(cl * 2)) cl *= 2;
    const unsigned sync_threads = std::min(1024u, ((max_nsub + (unsigned)cl - 1) / (unsigned)cl + 31u) / 32u * 32u);
    const size_t list_entries = nsubs + ni * (size_t)cl * sync_threads;      // the CTAs' work lists (k_jpeg_sync)
    if (sync_path) {
        rc = jp_grow(&h->d_subw, &h->d_subw_bytes, nsubs * 4 * sizeof(unsigned long long) + list_entries * sizeof(uint32_t));
        if (rc) return rc;
    }
    if ((int)sets.size() > S.tables_cap) {
        cudaFreeHost(S.h_tables); S.h_tables = nullptr;
        cudaFree(S.d_tables); S.d_tables = nullptr;
        S.tables_cap = (int)sets.size() * 2;
        ORBX_CUDA(cudaMallocHost((void**)&S.h_tables, sizeof(JpTables) * (size_t)S.tables_cap));
        ORBX_CUDA(cudaMalloc((void**)&S.d_tables, sizeof(JpTables) * (size_t)S.tables_cap));
    }
    memcpy(S.h_tables, sets.data(), sizeof(JpTables) * sets.size());

    // 3. upload, decode
    // uploads on the copy stream, behind the kernels that last read this set's device copy; the caller's stream waits for them
    if (S.decoded_pending) ORBX_CUDA(cudaStreamWaitEvent(h->copy_stream, S.decoded, 0));
    ORBX_CUDA(cudaMemcpyAsync(S.d_stream, S.h_stream, total, cudaMemcpyHostToDevice, h->copy_stream));
    ORBX_CUDA(cudaMemcpyAsync(S.d_intervals, S.h_intervals, ni * sizeof(JpInterval), cudaMemcpyHostToDevice, h->copy_stream));
    ORBX_CUDA(cudaMemcpyAsync(S.d_file_tables, S.h_file_tables, (size_t)nfiles * sizeof(int32_t), cudaMemcpyHostToDevice, h->copy_stream));
    ORBX_CUDA(cudaMemcpyAsync(S.d_tables, S.h_tables, sizeof(JpTables) * sets.size(), cudaMemcpyHostToDevice, h->copy_stream));
    ORBX_CUDA(cudaEventRecord(S.uploaded, h->copy_stream));
    S.uploaded_pending = true;
    ORBX_CUDA(cudaStreamWaitEvent(h->stream, S.uploaded, 0));
    ORBX_CUDA(cudaMemsetAsync(h->d_coefs, 0, (size_t)nfiles * blocks * 64 * sizeof(int16_t), h->stream));
    const size_t max_chunks = ((size_t)longest + JP_CHUNK - 1) / JP_CHUNK;
    if (longest > 16384 && ni <= 65535 && ni < 4 * (size_t)h->sm_count && ni * max_chunks <= (1u << 18)) {
        // few long intervals (files without restart markers): several CTAs per interval
        rc = jp_grow(&h->d_chunk_kept, &h->d_chunk_kept_bytes, ni * max_chunks * sizeof(uint32_t));
        if (rc) return rc;
        const dim3 cg((unsigned)max_chunks, (unsigned)ni);
        k_jpeg_unstuff_count<<<cg, JP_CHUNK_THREADS, 0, h->stream>>>(S.d_stream, S.d_intervals, (uint32_t)max_chunks, h->d_chunk_kept);
        k_jpeg_unstuff_chunk<<<cg, JP_CHUNK_THREADS, 0, h->stream>>>(S.d_stream, S.d_intervals, (uint32_t)max_chunks, h->d_chunk_kept, h->d_scratch, h->d_nwords);
    } else if (longest > 16384) {                   // long intervals: a CTA each
        k_jpeg_unstuff_block<<<(unsigned)ni, 1024, 0, h->stream>>>(S.d_stream, S.d_intervals, h->d_scratch, h->d_nwords);
    } else {
        const unsigned ub = (unsigned)((ni + JP_HUFF_THREADS / 32 - 1) / (JP_HUFF_THREADS / 32));
        k_jpeg_unstuff<<<ub, JP_HUFF_THREADS, 0, h->stream>>>(S.d_stream, S.d_intervals, (int)ni, h->d_scratch, h->d_nwords);
    }
    ORBX_CUDA(cudaGetLastError());
    int lanes = 1;                                  // active lanes per warp of the sequential kernel: enough warps to keep every scheduler busy first
    while (lanes < 32 && ni / (size_t)(lanes * 2) >= 1200) lanes *= 2;
    const int lane_step = 32 / lanes;
    const unsigned seq_blocks = (unsigned)((ni * (size_t)lane_step + JP_HUFF_THREADS - 1) / JP_HUFF_THREADS);
    if (!sync_path) {
        k_jpeg_huff<<<seq_blocks, JP_HUFF_THREADS, 0, h->stream>>>(S.d_intervals, (int)ni, S.d_tables, S.d_file_tables, h->d_scratch, h->d_nwords,
                                                                  h->d_coefs, g, lane_step);
    } else {
        unsigned long long* st = h->d_subw;       // [nsubs] exit states, [nsubs] states decoded from, [nsubs][4] counts, the work lists
        uint32_t* wl = reinterpret_cast<uint32_t*>(st + 4 * nsubs);
        const unsigned threads = sync_threads;
        if (cl == 1) {
            k_jpeg_sync<1><<<(unsigned)ni, threads, 0, h->stream>>>(S.d_intervals, S.d_tables, S.d_file_tables, h->d_scratch, h->d_nwords, st, st + nsubs,
                                                                   (int32_t*)(st + 2 * nsubs), wl, h->d_coefs, g, sub_bits);
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)ni * (unsigned)cl); cfg.blockDim = dim3(threads); cfg.stream = h->stream;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = (unsigned)cl; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            const JpInterval* a0 = S.d_intervals; const JpTables* a1 = S.d_tables; const int32_t* a2 = S.d_file_tables;
            const uint8_t* a3 = h->d_scratch; const uint32_t* a4 = h->d_nwords;
            unsigned long long* a5 = st; unsigned long long* a6 = st + nsubs; int32_t* a7 = (int32_t*)(st + 2 * nsubs);
            int16_t* a8 = h->d_coefs;
            if (cl == 2) ORBX_CUDA(cudaLaunchKernelEx(&cfg, k_jpeg_sync<2>, a0, a1, a2, a3, a4, a5, a6, a7, wl, a8, g, sub_bits));
            else if (cl == 4) ORBX_CUDA(cudaLaunchKernelEx(&cfg, k_jpeg_sync<4>, a0, a1, a2, a3, a4, a5, a6, a7, wl, a8, g, sub_bits));
            else ORBX_CUDA(cudaLaunchKernelEx(&cfg, k_jpeg_sync<8>, a0, a1, a2, a3, a4, a5, a6, a7, wl, a8, g, sub_bits));
        }
    }
    ORBX_CUDA(cudaGetLastError());
    JpOut o;
    memset(&o, 0, sizeof(o));
    size_t plane_off[3] = {0, 0, 0}, planes_pitch = 0;
    if (ncomp == 1) {
        o.base[0] = d_frames; o.pitch[0] = frame_pitch; o.stride[0] = stride; o.cw[0] = w; o.ch[0] = hh;
    } else {
        // the component planes, whole blocks each, one set per file
        for (int c = 0; c < 3; c++) {
            plane_off[c] = planes_pitch;
            o.stride[c] = (size_t)g.comp_bw[c] * 8;
            o.cw[c] = g.comp_bw[c] * 8; o.ch[c] = mcus_y * g.comp_v[c] * 8;
            planes_pitch += o.stride[c] * (size_t)o.ch[c];
        }
        rc = jp_grow(&h->d_planes, &h->d_planes_bytes, planes_pitch * (size_t)nfiles);
        if (rc) return rc;
        for (int c = 0; c < 3; c++) { o.base[c] = h->d_planes + plane_off[c]; o.pitch[c] = planes_pitch; }
    }
    k_jpeg_idct<<<dim3((blocks + JP_IDCT_THREADS - 1) / JP_IDCT_THREADS, (unsigned)nfiles), JP_IDCT_THREADS, 0, h->stream>>>(
        h->d_coefs, S.d_tables, S.d_file_tables, nfiles, g, o);
    ORBX_CUDA(cudaGetLastError());
    if (ncomp == 3) {
        const dim3 yg((unsigned)((w + 2047) / 2048), (unsigned)hh, (unsigned)nfiles);
        if (hf[0].hs == 2)
            k_jpeg_ycc<2><<<yg, 128, 0, h->stream>>>(h->d_planes, planes_pitch, plane_off[1], plane_off[2], (int)o.stride[0], (int)o.stride[1], w, hh, hf[0].vs,
                                                    d_frames, frame_pitch, stride);
        else
            k_jpeg_ycc<1><<<yg, 128, 0, h->stream>>>(h->d_planes, planes_pitch, plane_off[1], plane_off[2], (int)o.stride[0], (int)o.stride[1], w, hh, hf[0].vs,
                                                    d_frames, frame_pitch, stride);
        ORBX_CUDA(cudaGetLastError());
    }
    ORBX_CUDA(cudaEventRecord(S.decoded, h->stream));
    S.decoded_pending = true;
    return ORBX_OK;
}

extern "C" int jpgx_decode_gray_batch_dev(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int hh,
                                          uint8_t* d_frames, size_t frame_pitch, size_t stride)
{
    ORBX_NOTHROW(decode_batch_dev(h, files, sizes, nfiles, w, hh, d_frames, frame_pitch, stride, 1))
}

extern "C" int jpgx_decode_bgr_batch_dev(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int hh,
                                         uint8_t* d_frames, size_t frame_pitch, size_t stride)
{
    ORBX_NOTHROW(decode_batch_dev(h, files, sizes, nfiles, w, hh, d_frames, frame_pitch, stride, 3))
}

static int decode_batch_host(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int hh, uint8_t* frames,
                             size_t frame_pitch, size_t stride, int ncomp)
{
    ORBX_REQUIRE(h != nullptr, "jpgx_decode_batch: NULL handle");
    ORBX_REQUIRE(nfiles >= 0 && w >= 1 && hh >= 1, "jpgx_decode_batch: nfiles %d / size %dx%d out of range", nfiles, w, hh);
    if (nfiles == 0) return ORBX_OK;
    ORBX_REQUIRE(frames != nullptr, "jpgx_decode_batch: NULL pointer");
    ORBX_REQUIRE(stride >= (size_t)w * ncomp && frame_pitch >= stride * (size_t)(hh - 1) + (size_t)w * ncomp, "jpgx_decode_batch: stride %zu / frame pitch %zu too small for %dx%d x %d", stride, frame_pitch, w, hh, ncomp);
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t dpitch = align_up((size_t)w * ncomp, 16), dframe = dpitch * (size_t)hh;
    int rc = jp_grow(&h->d_frames, &h->d_frames_bytes, dframe * (size_t)nfiles);
    if (rc) return rc;
    rc = decode_batch_dev(h, files, sizes, nfiles, w, hh, h->d_frames, dframe, dpitch, ncomp);
    if (rc) return rc;
    for (int i = 0; i < nfiles; i++)
        ORBX_CUDA(cudaMemcpy2DAsync(frames + (size_t)i * frame_pitch, stride, h->d_frames + (size_t)i * dframe, dpitch, (size_t)w * ncomp, (size_t)hh,
                                    cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

extern "C" int jpgx_decode_gray_batch(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int hh, uint8_t* frames,
                                      size_t frame_pitch, size_t stride)
{
    ORBX_NOTHROW(decode_batch_host(h, files, sizes, nfiles, w, hh, frames, frame_pitch, stride, 1))
}

extern "C" int jpgx_decode_bgr_batch(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int hh, uint8_t* frames,
                                     size_t frame_pitch, size_t stride)
{
    ORBX_NOTHROW(decode_batch_host(h, files, sizes, nfiles, w, hh, frames, frame_pitch, stride, 3))
}

#ifdef JP_PROFILE
// {intervals, sum of rounds, maximum} of k_jpeg_sync since the last call (debug builds only: tools/jpeg_probe.py prints it)
extern "C" __attribute__((visibility("default"))) int jpgx_debug_rounds(unsigned int* out3)
{
    ORBX_CUDA(cudaDeviceSynchronize());
    ORBX_CUDA(cudaMemcpyFromSymbol(out3, g_jp_rounds, sizeof(g_jp_rounds)));
    unsigned int z[3] = {0, 0, 0};
    ORBX_CUDA(cudaMemcpyToSymbol(g_jp_rounds, z, sizeof(z)));
    return ORBX_OK;
}
#endif
