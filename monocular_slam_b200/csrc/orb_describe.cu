// orb_describe.cu -- K4 + K5: intensity-centroid orientation and rotated-BRIEF descriptors, one WARP per keypoint.
//
// Stages (viii)-(x) of OrbFeatureDetector::detect and the whole of OrbDescriptorExtractor::compute as called at
// reference src/FeatureExtractor.cpp:17,19 (OpenCV orb.cpp ICAngles / computeOrbDescriptors, core fastAtan2,
// imgproc 7x7 sigma-2 blur).
//
// A warp handles its keypoints in turn.  The 43x43 level patch around a keypoint (the rotated pattern reaches 18
// pixels, the blur 3 more) is fetched with aligned 16-byte loads (4 per patch row) INTO REGISTERS one keypoint ahead, so
// the global-load latency is hidden behind the previous keypoint's arithmetic, then stored to shared memory.  With only
// __syncwarp between the steps:
//   K4: integer moments m10 = sum u I, m01 = sum v I over the radius-15 disc (lane = column u; the disc is symmetric
//       under transposition, so column u owns rows |v| <= umax[|u|]), shuffle-reduced exactly; angle = fastAtan2(m01,
//       m10): OpenCV's 7th-order polynomial in degrees, every multiply and add rounded separately (no FMA contraction)
//       so the float result is bit-identical to the CPU.
//   K5: the 37x37 neighbourhood is blurred on the fly instead of blurring whole levels.  Row pass
//       r = k0*S0; r = fma(kj, Sj, r) (left to right), column pass c = k3*R0; c = fma(k3+j, R+j + R-j, c),
//       round-half-even to u8 -- the exact operation order of the CPU path.  Both passes run on packed fp32x2
//       (FFMA2 / FADD2: one instruction, two IEEE-exact operations): a lane filters two rows (resp. two columns) at
//       once with a sliding 7-tap window; the rounding is the add of 1.5 * 2^23 whose low result byte is the integer.
//       Lane l then evaluates BRIEF tests l, l+32, ...: both pattern points are
//       rotated in float32 (x = px*a - py*b, unfused), rounded half-even, looked up in the blurred patch, compared,
//       and a warp ballot assembles 32 tests into one little-endian word of the descriptor.
// OpenCV blurs only the level image itself, not the reflected border it keeps around it; sample positions that fall
// outside the level (possible only for caller-provided keypoints, never for detect's own) therefore read the
// un-blurred BORDER_REFLECT_101 pixel, and so does this kernel.
#include "common.cuh"

namespace orbx {

// 256 tests x (x0, y0, x1, y1); read-only, one coalesced 4-byte load per thread
__device__ __align__(16) const int8_t d_pattern[256 * 4] = {
#include "brief_pattern.inc"
};

namespace {

constexpr int OD_WARPS = 4;
#ifndef OD_KPW_N
#define OD_KPW_N 8
#endif
constexpr int OD_KPW = OD_KPW_N;         // keypoints per warp in the detect path
constexpr int OD_THREADS = OD_WARPS * 32;
constexpr int OD_R = 21;                 // raw patch radius
constexpr int OD_P = 2 * OD_R + 1;       // 43
constexpr int OD_RP = 68;                // raw patch pitch: 64-byte aligned window + 4 (17 words: rows fall in distinct banks)
constexpr int OD_BR = 18;                // blurred patch radius
constexpr int OD_B = 2 * OD_BR + 1;      // 37
constexpr int OD_BP = 40;                // blurred patch pitch
constexpr int OD_RUN = 19;               // outputs per lane and run in the blur passes (two overlapping runs cover 37)

constexpr int OD_BRP = 38;               // pitch of the row-pass result in floats (even: column pairs are read as float2)

struct __align__(16) WarpPatch {
    float row[OD_P * OD_BRP];            // first: 8-byte aligned for the float2 accesses
    uint8_t raw[OD_P * OD_RP];
    uint8_t val[OD_B * OD_BP];
};

__constant__ int c_umax[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// cv::fastAtan2 (scalar path), degrees in [0, 360)
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / 3.141592653589793238462643383279502884);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.141592653589793238462643383279502884);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.141592653589793238462643383279502884);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.141592653589793238462643383279502884);
    const float eps = 2.2204460492503131e-16f;   // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// The 43 x 64-byte window of an interior patch as 16-byte vectors held in registers (<= 6 per lane): issued for the
// NEXT keypoint before the current one is processed, stored to shared memory when its turn comes, so the global-load
// latency of the patch (the dominant stall of this kernel) is off the critical path.
constexpr int OD_VEC = OD_P * 4;                       // 172 vectors
constexpr int OD_VPL = (OD_VEC + 31) / 32;             // 6 per lane
struct PatchRegs { uint4 q[OD_VPL]; };

__device__ __forceinline__ bool patch_interior(int xi, int yi, int w, int h)
{
    return xi - OD_R >= 0 && yi - OD_R >= 0 && xi + OD_R < w && yi + OD_R < h;
}
__device__ __forceinline__ void patch_issue(PatchRegs& P, const uint8_t* __restrict__ img, int pitch, int xi, int yi)
{
    const int lane = threadIdx.x & 31;
    const int x0 = xi - OD_R, y0 = yi - OD_R;
    const uint8_t* src = img + (size_t)y0 * pitch + (x0 & ~15);
#pragma unroll
    for (int k = 0; k < OD_VPL; k++) {
        const int i = lane + 32 * k;
        if (i < OD_VEC) P.q[k] = *reinterpret_cast<const uint4*>(src + (size_t)(i >> 2) * pitch + 16 * (i & 3));
    }
}
__device__ __forceinline__ void patch_commit(uint8_t* raw, const PatchRegs& P)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < OD_VPL; k++) {
        const int i = lane + 32 * k;
        if (i < OD_VEC) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(raw + (i >> 2) * OD_RP + 16 * (i & 3));
            dst[0] = P.q[k].x; dst[1] = P.q[k].y; dst[2] = P.q[k].z; dst[3] = P.q[k].w;
        }
    }
}

// One warp, one keypoint.  (xi, yi): integer keypoint position on its level.  mode: ORBX_DO_ANGLE computes the angle,
// else angle_in is used; ORBX_DO_DESC writes the 32 descriptor bytes.  Returns the angle (valid on every lane).
__device__ __forceinline__ float orient_describe_warp(WarpPatch& S, const char4* __restrict__ s_pat, const uint8_t* __restrict__ img,
                                                      int w, int h, int pitch, int xi, int yi, int mode, float angle_in,
                                                      uint8_t* __restrict__ desc_row, bool staged = false)
{
    // staged: the caller has already put the (interior) patch into S.raw with patch_commit
    const int lane = threadIdx.x & 31;
    const int x0 = xi - OD_R, y0 = yi - OD_R;
    const bool interior = patch_interior(xi, yi, w, h);
    int a = 0;   // smem column of patch column 0
    if (staged) {
        a = x0 & 15;
    } else if (interior) {
        a = x0 & 15;
        const uint8_t* src = img + (size_t)y0 * pitch + (x0 - a);     // 16-byte aligned; the 64-byte window holds a + 43 <= 58 bytes
        for (int i = lane; i < OD_P * 4; i += 32) {
            const int r = i >> 2, v = i & 3;
            const uint4 q = *reinterpret_cast<const uint4*>(src + (size_t)r * pitch + 16 * v);
            uint32_t* dst = reinterpret_cast<uint32_t*>(S.raw + r * OD_RP + 16 * v);
            dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
        }
    } else {
        for (int i = lane; i < OD_P * OD_P; i += 32) {
            const int r = i / OD_P, c = i - r * OD_P;
            S.raw[r * OD_RP + c] = img[(size_t)reflect101(y0 + r, h) * pitch + reflect101(x0 + c, w)];
        }
    }
    __syncwarp();

    float angle = angle_in;
    if (mode & ORBX_DO_ANGLE) {
        int m10 = 0, m01 = 0;
        const int u = lane - 15;
        if (lane < 31) {
            // the disc is symmetric under transposition (|u| <= umax[|v|]  <=>  |v| <= umax[|u|]), so column u owns the rows
            // |v| <= vm: one table lookup per lane, and m10 = u * sum(val) needs no multiply inside the loop
            const uint8_t* col = S.raw + OD_R * OD_RP + a + OD_R + u;
            const int vm = c_umax[u < 0 ? -u : u];
            int sum = col[0];
#pragma unroll
            for (int v = 1; v <= 15; v++) {
                if (v <= vm) {
                    const int lo = col[-v * OD_RP], hi = col[v * OD_RP];
                    sum += lo + hi;
                    m01 += v * (hi - lo);
                }
            }
            m10 = u * sum;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
            m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
        }
        angle = fast_atan2_deg((float)m01, (float)m10);   // every lane holds the same sums
    }
    if (!(mode & ORBX_DO_DESC)) return angle;

    // one double-precision cos/sin per keypoint: a = (float)cos((double)ang), as the CPU path does
    float ca = 0.f, sa = 0.f;
    if (lane == 0) {
        const float ang = __fmul_rn(angle, (float)(3.141592653589793238462643383279502884 / 180.f));
        ca = (float)cos((double)ang);
        sa = (float)sin((double)ang);
    }
    ca = __shfl_sync(0xFFFFFFFFu, ca, 0);
    sa = __shfl_sync(0xFFFFFFFFu, sa, 0);

    // cv::getGaussianKernel(7, 2, CV_32F)
    const float k0 = 0x1.1f5f62p-4f, k1 = 0x1.0c70fcp-3f, k2 = 0x1.869472p-3f, k3 = 0x1.ba95c0p-3f;
    if (interior) {
        // ---- blur on packed fp32x2 (FFMA2: one instruction, two IEEE-exact FMAs).  Row pass: lane p < 22 filters rows 2p and
        // 2p+1 together, streaming a 7-tap window along the 43 columns; column pass: lane q < 19 filters columns 2q and 2q+1
        // together down the 43 rows.  Same operation order per element as the scalar path below (and as OpenCV).
        const float2 K0 = make_float2(k0, k0), K1 = make_float2(k1, k1), K2 = make_float2(k2, k2), K3 = make_float2(k3, k3);
        if (lane < (OD_P + 1) / 2) {
            const int r0 = 2 * lane, r1 = min(r0 + 1, OD_P - 1);
            const uint8_t* s0 = S.raw + r0 * OD_RP + a;
            const uint8_t* s1 = S.raw + r1 * OD_RP + a;
            float* o0 = S.row + r0 * OD_BRP;
            float* o1 = S.row + r1 * OD_BRP;
            float2 w[7];            // sliding window over the columns (the unrolled loop renames instead of moving)
#pragma unroll
            for (int i = 0; i < 6; i++) w[i] = make_float2((float)s0[i], (float)s1[i]);
#pragma unroll
            for (int j = 0; j < OD_B; j++) {
                w[(j + 6) % 7] = make_float2((float)s0[j + 6], (float)s1[j + 6]);
                float2 acc = __fmul2_rn(K0, w[j % 7]);
                acc = __ffma2_rn(K1, w[(j + 1) % 7], acc);
                acc = __ffma2_rn(K2, w[(j + 2) % 7], acc);
                acc = __ffma2_rn(K3, w[(j + 3) % 7], acc);
                acc = __ffma2_rn(K2, w[(j + 4) % 7], acc);
                acc = __ffma2_rn(K1, w[(j + 5) % 7], acc);
                acc = __ffma2_rn(K0, w[(j + 6) % 7], acc);
                o0[j] = acc.x;
                o1[j] = acc.y;      // lane 21: r1 == r0, the same value to the same address
            }
        }
        __syncwarp();
        if (lane < (OD_B + 1) / 2) {
            const float2* p = reinterpret_cast<const float2*>(S.row) + lane;      // columns 2q, 2q+1; row stride OD_BRP / 2
            // round half to even by adding 1.5 * 2^23: the low byte of the sum's bit pattern is the integer (0 <= acc <= 255.0003)
            const float2 MAGIC = make_float2(12582912.f, 12582912.f);
            float2 w[7];
#pragma unroll
            for (int i = 0; i < 6; i++) w[i] = p[i * (OD_BRP / 2)];
#pragma unroll
            for (int j = 0; j < OD_B; j++) {
                w[(j + 6) % 7] = p[(j + 6) * (OD_BRP / 2)];
                float2 acc = __fmul2_rn(K3, w[(j + 3) % 7]);
                acc = __ffma2_rn(K2, __fadd2_rn(w[(j + 4) % 7], w[(j + 2) % 7]), acc);
                acc = __ffma2_rn(K1, __fadd2_rn(w[(j + 5) % 7], w[(j + 1) % 7]), acc);
                acc = __ffma2_rn(K0, __fadd2_rn(w[(j + 6) % 7], w[j % 7]), acc);
                const float2 t = __fadd2_rn(acc, MAGIC);
                // column 37 (lane 18's second one) does not exist: its byte lands in the padding of the row (pitch 40)
                *reinterpret_cast<uint16_t*>(S.val + j * OD_BP + 2 * lane) =
                    (uint16_t)__byte_perm(__float_as_uint(t.x), __float_as_uint(t.y), 0x0040);
            }
        }
        __syncwarp();
    } else {
    // ---- scalar path for patches that leave the level (caller-provided keypoints near the border only)
    // row pass: 43 rows x 37 columns as 86 runs of 19 outputs (columns 0..18 and 18..36)
    for (int id = lane; id < OD_P * 2; id += 32) {
        const int r = id >> 1, c0 = (id & 1) * (OD_B - OD_RUN);
        const uint8_t* s = S.raw + r * OD_RP + a + c0;      // taps c .. c+6 == patch columns (c+3) +- 3
        float f[OD_RUN + 6];
#pragma unroll
        for (int j = 0; j < OD_RUN + 6; j++) f[j] = (float)s[j];
        float* o = S.row + r * OD_BRP + c0;
#pragma unroll
        for (int j = 0; j < OD_RUN; j++) {
            float acc = __fmul_rn(k0, f[j]);
            acc = __fmaf_rn(k1, f[j + 1], acc);
            acc = __fmaf_rn(k2, f[j + 2], acc);
            acc = __fmaf_rn(k3, f[j + 3], acc);
            acc = __fmaf_rn(k2, f[j + 4], acc);
            acc = __fmaf_rn(k1, f[j + 5], acc);
            acc = __fmaf_rn(k0, f[j + 6], acc);
            o[j] = acc;
        }
    }
    __syncwarp();
    // column pass, symmetric form: 37 columns x 2 runs of 19 rows; positions outside the level keep the raw pixel
    for (int id = lane; id < OD_B * 2; id += 32) {
        const int half = id >= OD_B ? 1 : 0;
        const int c = id - half * OD_B, r0 = half * (OD_B - OD_RUN);
        const float* p = S.row + r0 * OD_BRP + c;
        float f[OD_RUN + 6];
#pragma unroll
        for (int j = 0; j < OD_RUN + 6; j++) f[j] = p[j * OD_BRP];
#pragma unroll
        for (int j = 0; j < OD_RUN; j++) {
            float acc = __fmul_rn(k3, f[j + 3]);
            acc = __fmaf_rn(k2, __fadd_rn(f[j + 4], f[j + 2]), acc);
            acc = __fmaf_rn(k1, __fadd_rn(f[j + 5], f[j + 1]), acc);
            acc = __fmaf_rn(k0, __fadd_rn(f[j + 6], f[j]), acc);
            int v = __float2int_rn(acc);
            // no clamp: the kernel's weights sum to 1 within 1e-6, so 0 <= acc <= 255.0003 and v is already a valid u8
            const int r = r0 + j;
            const int gy = yi - OD_BR + r, gx = xi - OD_BR + c;
            if (gx < 0 || gx >= w || gy < 0 || gy >= h) v = S.raw[(r + 3) * OD_RP + a + c + 3];
            S.val[r * OD_BP + c] = (uint8_t)v;
        }
    }
    __syncwarp();
    }

    // ---- rBRIEF: lane l evaluates tests l, l + 32, ...; word j of the descriptor is the ballot of round j
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const char4 pt = s_pat[32 * j + lane];
        const float px0 = (float)pt.x, py0 = (float)pt.y, px1 = (float)pt.z, py1 = (float)pt.w;
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(px0, ca), __fmul_rn(py0, sa)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(px0, sa), __fmul_rn(py0, ca)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(px1, ca), __fmul_rn(py1, sa)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(px1, sa), __fmul_rn(py1, ca)));
        const int v0 = S.val[(OD_BR + iy0) * OD_BP + (OD_BR + ix0)];
        const int v1 = S.val[(OD_BR + iy1) * OD_BP + (OD_BR + ix1)];
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, v0 < v1);
        if (lane == j) mine = word;
    }
    if (lane < 8) reinterpret_cast<uint32_t*>(desc_row)[lane] = mine;
    __syncwarp();   // the patch buffers are reused by this warp's next keypoint
    return angle;
}

__device__ __forceinline__ void load_pattern(char4* s_pat)
{
    for (int i = threadIdx.x; i < 256; i += OD_THREADS) s_pat[i] = __ldg(reinterpret_cast<const char4*>(d_pattern) + i);
    __syncthreads();
}

// detect / detect+compute: warps stride over the frame's keypoints; keypoint kidx is located through the per-level
// selected counts (levels are concatenated in order, each already sorted by (y, x)).
__global__ void __launch_bounds__(OD_THREADS)
k_orient_describe(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slots, size_t slot_stride,
                  const Sel* __restrict__ sel, size_t sel_stride, FrameCounters* __restrict__ ctr,
                  orbx_keypoint* __restrict__ out, uint8_t* __restrict__ desc, int cap, int32_t* __restrict__ counts, int mode)
{
    __shared__ __align__(16) WarpPatch s_patch[OD_WARPS];
    __shared__ char4 s_pat[256];
    load_pattern(s_pat);
    const int frame = blockIdx.y, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FrameCounters& C = ctr[frame];
    int total = 0;
    for (int l = 0; l < g.nlevels; l++) total += C.nsel[l];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[frame] = total;
        C.total = total;
        if (total > cap) atomicOr(&C.overflow, 4);
    }
    const int limit = min(total, cap);
    auto locate = [&](int kidx, int& level, int& rank) {
        int off = 0;
        level = 0; rank = 0;
        for (int l = 0; l < g.nlevels; l++) {
            const int n = C.nsel[l];
            if (kidx >= off && kidx < off + n) { level = l; rank = kidx - off; }
            off += n;
        }
    };
    // software pipeline over the warp's keypoints: the patch of keypoint i+1 travels from global memory into registers
    // while keypoint i is processed out of shared memory
    const int kstep = gridDim.x * OD_WARPS;
    int kidx = blockIdx.x * OD_WARPS + wid;
    PatchRegs P;
    int level = 0, rank = 0;
    Sel s = { 0u, 0.f };
    bool staged = false;
    if (kidx < limit) {
        locate(kidx, level, rank);
        s = sel[frame * sel_stride + g.lv[level].sel_off + rank];
        const LevelGeom& L = g.lv[level];
        staged = patch_interior((int)(s.xy & 0xFFFFu), (int)(s.xy >> 16), L.w, L.h);
        if (staged) patch_issue(P, slots + frame * slot_stride + L.img_off, L.pitch, (int)(s.xy & 0xFFFFu), (int)(s.xy >> 16));
    }
    for (; kidx < limit; kidx += kstep) {
        const LevelGeom& L = g.lv[level];
        const Sel cur = s;
        const int cur_level = level;
        const bool cur_staged = staged;
        if (cur_staged) patch_commit(s_patch[wid].raw, P);
        __syncwarp();
        if (kidx + kstep < limit) {         // issue the next keypoint's patch
            locate(kidx + kstep, level, rank);
            s = sel[frame * sel_stride + g.lv[level].sel_off + rank];
            const LevelGeom& NL = g.lv[level];
            staged = patch_interior((int)(s.xy & 0xFFFFu), (int)(s.xy >> 16), NL.w, NL.h);
            if (staged) patch_issue(P, slots + frame * slot_stride + NL.img_off, NL.pitch, (int)(s.xy & 0xFFFFu), (int)(s.xy >> 16));
        }
        const int xi = (int)(cur.xy & 0xFFFFu), yi = (int)(cur.xy >> 16);
        const uint8_t* img = slots + frame * slot_stride + L.img_off;
        uint8_t* drow = desc ? desc + ((size_t)frame * cap + kidx) * 32 : nullptr;
        const float angle = orient_describe_warp(s_patch[wid], s_pat, img, L.w, L.h, L.pitch, xi, yi, mode, -1.f, drow, cur_staged);
        if (lane == 0) {
            orbx_keypoint k;
            k.x = __fmul_rn((float)xi, L.scale);
            k.y = __fmul_rn((float)yi, L.scale);
            k.size = __fmul_rn(31.f, L.scale);
            k.angle = angle;
            k.response = cur.response;
            k.octave = cur_level;
            k.class_id = -1;
            out[(size_t)frame * cap + kidx] = k;
        }
    }
}

// compute with caller-provided keypoints (already border-filtered and grouped by octave on the host).
__global__ void __launch_bounds__(OD_THREADS)
k_describe_given(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slot, const orbx_keypoint* __restrict__ kps, int n,
                 uint8_t* __restrict__ desc)
{
    __shared__ __align__(16) WarpPatch s_patch[OD_WARPS];
    __shared__ char4 s_pat[256];
    load_pattern(s_pat);
    const int wid = threadIdx.x >> 5;
    for (int i = blockIdx.x * OD_WARPS + wid; i < n; i += gridDim.x * OD_WARPS) {
        const orbx_keypoint k = kps[i];
        const LevelGeom& L = g.lv[k.octave];
        const int xi = __float2int_rn(__fmul_rn(k.x, L.inv_scale)), yi = __float2int_rn(__fmul_rn(k.y, L.inv_scale));
        orient_describe_warp(s_patch[wid], s_pat, slot + L.img_off, L.w, L.h, L.pitch, xi, yi, ORBX_DO_DESC, k.angle, desc + (size_t)i * 32);
    }
}

}  // namespace

cudaError_t launch_orient_describe(const FrameGeom& g, const uint8_t* slots, size_t slot_stride, const Sel* sel,
                                   size_t sel_stride, FrameCounters* ctr, orbx_keypoint* out, uint8_t* desc, int cap,
                                   int32_t* counts, int nframes, int mode, cudaStream_t s)
{
    // a warp handles OD_KPW keypoints in turn (the next one's patch is prefetched while the current one is processed)
    int blocks = (cap + OD_WARPS * OD_KPW - 1) / (OD_WARPS * OD_KPW);
    if (blocks > 2048) blocks = 2048;
    dim3 grid(blocks, nframes);
    k_orient_describe<<<grid, OD_THREADS, 0, s>>>(g, slots, slot_stride, sel, sel_stride, ctr, out, desc, cap, counts, mode);
    return cudaGetLastError();
}

cudaError_t launch_describe_given(const FrameGeom& g, const uint8_t* slot, const orbx_keypoint* kps, int n, uint8_t* desc,
                                  cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    int blocks = (n + OD_WARPS - 1) / OD_WARPS;
    if (blocks > 4096) blocks = 4096;
    k_describe_given<<<blocks, OD_THREADS, 0, s>>>(g, slot, kps, n, desc);
    return cudaGetLastError();
}

}  // namespace orbx
