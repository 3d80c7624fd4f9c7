// orb_describe.cu -- K4 + K5: intensity-centroid orientation and rotated-BRIEF descriptors, one CTA per keypoint.
//
// Stages (viii)-(x) of OrbFeatureDetector::detect and the whole of OrbDescriptorExtractor::compute as called at
// reference src/FeatureExtractor.cpp:17,19 (OpenCV orb.cpp ICAngles / computeOrbDescriptors, core fastAtan2,
// imgproc 7x7 sigma-2 blur).
//
// A CTA of 256 threads stages the 43x43 level patch around the keypoint in shared memory (the rotated pattern reaches
// 18 pixels, the blur 3 more), then
//   K4: integer moments m10 = sum u I, m01 = sum v I over the radius-15 disc (umax table), block-reduced exactly;
//       angle = fastAtan2(m01, m10): OpenCV's 7th-order polynomial in degrees, every multiply and add rounded
//       separately (no FMA contraction) so the float result is bit-identical to the CPU.
//   K5: the 37x37 neighbourhood is blurred on the fly instead of blurring whole levels: row pass
//       r = k0*S0; r = fma(kj, Sj, r) (left to right), column pass c = k3*R0; c = fma(k3+j, R+j + R-j, c),
//       round-half-even to u8 -- the exact operation order of the CPU path.  Thread t then evaluates BRIEF test t:
//       both pattern points are rotated in float32 (x = px*a - py*b, unfused), rounded half-even, looked up in the
//       blurred patch, compared, and a warp ballot assembles 32 tests into one little-endian word of the descriptor.
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

constexpr int OD_THREADS = 256;
constexpr int OD_R = 21;                 // raw patch radius
constexpr int OD_P = 2 * OD_R + 1;       // 43
constexpr int OD_PP = 44;                // raw patch pitch
constexpr int OD_BR = 18;                // blurred patch radius
constexpr int OD_B = 2 * OD_BR + 1;      // 37
constexpr int OD_BP = 40;                // blurred patch pitch

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

__device__ const int d_umax[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };

// Shared body.  (xi, yi): integer keypoint position on its level.  mode: ORBX_DO_ANGLE computes the angle, else
// angle_in is used; ORBX_DO_DESC writes the 32 descriptor bytes.  Returns the angle (valid on every thread).
__device__ __forceinline__ float orient_describe_body(const uint8_t* __restrict__ img, int w, int h, int pitch, int xi, int yi,
                                                      int mode, float angle_in, uint8_t* __restrict__ desc_row)
{
    __shared__ uint8_t s_raw[OD_P * OD_PP];
    __shared__ float s_row[OD_P * OD_B];
    __shared__ uint8_t s_val[OD_B * OD_BP];
    __shared__ int s_m[2][OD_THREADS / 32];
    __shared__ float s_angle, s_cos, s_sin;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool interior = xi >= OD_R && yi >= OD_R && xi + OD_R < w && yi + OD_R < h;

    for (int i = tid; i < OD_P * OD_P; i += OD_THREADS) {
        const int r = i / OD_P, c = i - r * OD_P;
        int gy = yi - OD_R + r, gx = xi - OD_R + c;
        if (!interior) { gy = reflect101(gy, h); gx = reflect101(gx, w); }
        s_raw[r * OD_PP + c] = img[(size_t)gy * pitch + gx];
    }
    __syncthreads();

    float angle = angle_in;
    if (mode & ORBX_DO_ANGLE) {
        int m10 = 0, m01 = 0;
        for (int i = tid; i < 31 * 31; i += OD_THREADS) {
            const int v = i / 31 - 15, u = i - (v + 15) * 31 - 15;
            if (abs(u) <= d_umax[abs(v)]) {
                const int val = s_raw[(OD_R + v) * OD_PP + (OD_R + u)];
                m10 += u * val;
                m01 += v * val;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
            m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
        }
        if (lane == 0) { s_m[0][wid] = m10; s_m[1][wid] = m01; }
        __syncthreads();
        if (tid == 0) {
            int a10 = 0, a01 = 0;
            for (int k = 0; k < OD_THREADS / 32; k++) { a10 += s_m[0][k]; a01 += s_m[1][k]; }
            s_angle = fast_atan2_deg((float)a01, (float)a10);
        }
        __syncthreads();
        angle = s_angle;
    }
    if (!(mode & ORBX_DO_DESC)) return angle;
    if (tid == 0) {   // one double-precision cos/sin per keypoint: a = (float)cos((double)ang), as the CPU path does
        const float ang = __fmul_rn(angle, (float)(3.141592653589793238462643383279502884 / 180.f));
        s_cos = (float)cos((double)ang);
        s_sin = (float)sin((double)ang);
    }

    // ---- blur: row pass over 43 rows x 37 columns
    // cv::getGaussianKernel(7, 2, CV_32F)
    const float k0 = 0x1.1f5f62p-4f, k1 = 0x1.0c70fcp-3f, k2 = 0x1.869472p-3f, k3 = 0x1.ba95c0p-3f, k4 = k2, k5 = k1, k6 = k0;
    for (int i = tid; i < OD_P * OD_B; i += OD_THREADS) {
        const int r = i / OD_B, c = i - r * OD_B;
        const uint8_t* s = s_raw + r * OD_PP + c;     // taps c .. c+6 == patch columns (c+3) +- 3
        float acc = __fmul_rn(k0, (float)s[0]);
        acc = __fmaf_rn(k1, (float)s[1], acc);
        acc = __fmaf_rn(k2, (float)s[2], acc);
        acc = __fmaf_rn(k3, (float)s[3], acc);
        acc = __fmaf_rn(k4, (float)s[4], acc);
        acc = __fmaf_rn(k5, (float)s[5], acc);
        acc = __fmaf_rn(k6, (float)s[6], acc);
        s_row[i] = acc;
    }
    __syncthreads();
    // ---- column pass, symmetric form; positions outside the level keep the raw reflected pixel
    for (int i = tid; i < OD_B * OD_B; i += OD_THREADS) {
        const int r = i / OD_B, c = i - r * OD_B;
        const float* p = s_row + (r + 3) * OD_B + c;
        float acc = __fmul_rn(k3, p[0]);
        acc = __fmaf_rn(k4, __fadd_rn(p[OD_B], p[-OD_B]), acc);
        acc = __fmaf_rn(k5, __fadd_rn(p[2 * OD_B], p[-2 * OD_B]), acc);
        acc = __fmaf_rn(k6, __fadd_rn(p[3 * OD_B], p[-3 * OD_B]), acc);
        int v = __float2int_rn(acc);
        v = min(max(v, 0), 255);
        if (!interior) {
            const int gy = yi - OD_BR + r, gx = xi - OD_BR + c;
            if (gx < 0 || gx >= w || gy < 0 || gy >= h) v = s_raw[(r + 3) * OD_PP + (c + 3)];
        }
        s_val[r * OD_BP + c] = (uint8_t)v;
    }
    __syncthreads();

    // ---- rBRIEF: thread t evaluates test t
    const float a = s_cos, b = s_sin;   // written before the two barriers of the blur passes
    const char4 pt = __ldg(reinterpret_cast<const char4*>(d_pattern) + tid);
    const float x0 = (float)pt.x, y0 = (float)pt.y, x1 = (float)pt.z, y1 = (float)pt.w;
    const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int v0 = s_val[(OD_BR + iy0) * OD_BP + (OD_BR + ix0)];
    const int v1 = s_val[(OD_BR + iy1) * OD_BP + (OD_BR + ix1)];
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, v0 < v1);
    if (lane == 0) reinterpret_cast<uint32_t*>(desc_row)[wid] = word;
    return angle;
}

// detect / detect+compute: CTAs stride over the frame's keypoints; keypoint kidx is located through the per-level
// selected counts (levels are concatenated in order, each already sorted by (y, x)).
__global__ void __launch_bounds__(OD_THREADS)
k_orient_describe(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slots, size_t slot_stride,
                  const Sel* __restrict__ sel, size_t sel_stride, FrameCounters* __restrict__ ctr,
                  orbx_keypoint* __restrict__ out, uint8_t* __restrict__ desc, int cap, int32_t* __restrict__ counts, int mode)
{
    const int frame = blockIdx.y;
    FrameCounters& C = ctr[frame];
    int total = 0;
    for (int l = 0; l < g.nlevels; l++) total += C.nsel[l];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[frame] = total;
        C.total = total;
        if (total > cap) atomicOr(&C.overflow, 4);
    }
    const int limit = min(total, cap);
    for (int kidx = blockIdx.x; kidx < limit; kidx += gridDim.x) {
        int off = 0, level = 0, rank = 0;
        for (int l = 0; l < g.nlevels; l++) {
            const int n = C.nsel[l];
            if (kidx >= off && kidx < off + n) { level = l; rank = kidx - off; }
            off += n;
        }
        const LevelGeom& L = g.lv[level];
        const Sel s = sel[frame * sel_stride + L.sel_off + rank];
        const int xi = (int)(s.xy & 0xFFFFu), yi = (int)(s.xy >> 16);
        const uint8_t* img = slots + frame * slot_stride + L.img_off;
        uint8_t* drow = desc ? desc + ((size_t)frame * cap + kidx) * 32 : nullptr;
        const float angle = orient_describe_body(img, L.w, L.h, L.pitch, xi, yi, mode, -1.f, drow);
        if (threadIdx.x == 0) {
            orbx_keypoint k;
            k.x = __fmul_rn((float)xi, L.scale);
            k.y = __fmul_rn((float)yi, L.scale);
            k.size = __fmul_rn(31.f, L.scale);
            k.angle = angle;
            k.response = s.response;
            k.octave = level;
            k.class_id = -1;
            out[(size_t)frame * cap + kidx] = k;
        }
        __syncthreads();   // the shared patch buffers are reused by the next keypoint of this CTA
    }
}

// compute with caller-provided keypoints (already border-filtered and grouped by octave on the host).
__global__ void __launch_bounds__(OD_THREADS)
k_describe_given(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ slot, const orbx_keypoint* __restrict__ kps,
                 uint8_t* __restrict__ desc)
{
    const orbx_keypoint k = kps[blockIdx.x];
    const LevelGeom& L = g.lv[k.octave];
    const int xi = __float2int_rn(__fmul_rn(k.x, L.inv_scale)), yi = __float2int_rn(__fmul_rn(k.y, L.inv_scale));
    orient_describe_body(slot + L.img_off, L.w, L.h, L.pitch, xi, yi, ORBX_DO_DESC, k.angle, desc + (size_t)blockIdx.x * 32);
}

}  // namespace

cudaError_t launch_orient_describe(const FrameGeom& g, const uint8_t* slots, size_t slot_stride, const Sel* sel,
                                   size_t sel_stride, FrameCounters* ctr, orbx_keypoint* out, uint8_t* desc, int cap,
                                   int32_t* counts, int nframes, int mode, cudaStream_t s)
{
    dim3 grid(cap < 4096 ? cap : 4096, nframes);
    k_orient_describe<<<grid, OD_THREADS, 0, s>>>(g, slots, slot_stride, sel, sel_stride, ctr, out, desc, cap, counts, mode);
    return cudaGetLastError();
}

cudaError_t launch_describe_given(const FrameGeom& g, const uint8_t* slot, const orbx_keypoint* kps, int n, uint8_t* desc,
                                  cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_describe_given<<<n, OD_THREADS, 0, s>>>(g, slot, kps, desc);
    return cudaGetLastError();
}

}  // namespace orbx
