/*
 * orb_oracle.c -- CPU restatement of the reference's ORB + brute-force Hamming path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under monocular_slam_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / reported baseline.
 *
 * What it restates.  The reference (eastgeneral2007/Monocular_SLAM) owns 6 lines of the
 * hot path: `detector.detect` / `extractor.compute` (src/FeatureExtractor.cpp:17,19) on
 * default-constructed cv::ORB objects (src/FeatureExtractor.h:23-24) and
 * `BFMatcher(NORM_HAMMING,false).knnMatch(..,2)` + Lowe ratio
 * (src/CameraPoseEstimator.cpp:200-213).  The arithmetic lives in the un-vendored
 * third-party dependency OpenCV (pinned 2.4.13 in cmake_modules/superbuild.cmake2:30-35;
 * the only runnable build here is cv2 4.13.0).  Every function below restates the
 * published OpenCV algorithm stage by stage in plain scalar C.
 *
 * Parity pinning.  The reference has no tests or golden vectors for this path
 * (SURVEY.md section 4), so with respect to the reference's own test-suite parity is
 * UNPINNED.  The oracle is pinned instead against cv2 4.13.0 run in the build
 * container: tests/golden/ holds cv2's outputs (keypoints, descriptors, matches) with
 * the generating script tests/golden/make_golden.py, and tests/test_oracle_vs_golden.py
 * requires this file to reproduce them bit for bit.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off (see oracle/Makefile).  FMA is used
 * only where written explicitly with fmaf() -- the blur -- because that is what the
 * oracle host's OpenCV does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_MAX_LEVELS 16

typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint; /* == cv::KeyPoint field order (28 bytes) */

typedef struct {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels, edge_threshold, first_level, wta_k, score_type /*0 HARRIS, 1 FAST*/, patch_size, fast_threshold;
} orc_params; /* cv::ORB constructor arguments; reference uses the defaults, src/FeatureExtractor.h:23-24 */

static const int g_pattern[256 * 4] = {
#include "brief_pattern.inc"
};

static inline int rne_f(float v) { return (int)lrintf(v); }   /* cvRound: round half to even */
static inline int rne_d(double v) { return (int)lrint(v); }

/* ---- A0: per-level scale, size and keypoint quota (OpenCV orb.cpp getScale / detectAndCompute) ---- */
float orc_level_scale(const orc_params* p, int level)
{
    return (float)pow((double)p->scale_factor, (double)(level - p->first_level));
}

void orc_level_size(const orc_params* p, int w, int h, int level, int* lw, int* lh)
{
    float s = orc_level_scale(p, level);
    *lw = rne_f((float)w / s);
    *lh = rne_f((float)h / s);
}

void orc_level_quotas(const orc_params* p, int* q)
{
    float factor = (float)(1.0 / (double)p->scale_factor);
    float nd = (float)p->nfeatures * (1.f - factor) / (1.f - (float)pow((double)factor, (double)p->nlevels));
    int sum = 0;
    for (int l = 0; l < p->nlevels - 1; l++) {
        q[l] = rne_f(nd);
        sum += q[l];
        nd *= factor;
    }
    int last = p->nfeatures - sum;
    q[p->nlevels - 1] = last > 0 ? last : 0;
}

/* ---- A0': gray conversion.  cv::ORB starts with `if (image.type() != CV_8UC1) cvtColor(image, gray, COLOR_BGR2GRAY)`
 * (the reference loads frames with CV_LOAD_IMAGE_UNCHANGED, src/FrameLoader.cpp:62, so 3-channel frames reach
 * detect()/compute() at src/FeatureExtractor.cpp:17,19).  OpenCV 4.x 8-bit path: 15-bit fixed point,
 * gray = (B*3735 + G*19235 + R*9798 + 2^14) >> 15 (imgproc color_rgb.simd.hpp RGB2Gray<uchar>); checked against
 * cv2.cvtColor in tests/golden/make_golden.py. */
void orc_bgr2gray(const uint8_t* bgr, int w, int h, size_t stride, uint8_t* gray /* w*h */)
{
    for (int y = 0; y < h; y++) {
        const uint8_t* s = bgr + (size_t)y * stride;
        uint8_t* d = gray + (size_t)y * w;
        for (int x = 0; x < w; x++)
            d[x] = (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + (1 << 14)) >> 15);
    }
}

/* ---- A1: INTER_LINEAR_EXACT resize, 8-bit, 1 channel (OpenCV imgproc resize.cpp, fixed-point path) ---- */
static void linear_coeffs(int n, int m, int* ofs, int* c1)
{
    double s = 1.0 / ((double)m / (double)n);
    for (int d = 0; d < m; d++) {
        double f = s * (d + 0.5) - 0.5;
        int i = (int)floor(f);
        if (i < 0) { ofs[d] = 0; c1[d] = 0; }
        else if (i >= n - 1) { ofs[d] = n - 1; c1[d] = 0; }
        else { ofs[d] = i; c1[d] = rne_d((f - i) * 256.0); }
    }
}

void orc_resize_linear_exact(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride)
{
    int* ox = (int*)malloc(sizeof(int) * dw), *cx = (int*)malloc(sizeof(int) * dw);
    int* oy = (int*)malloc(sizeof(int) * dh), *cy = (int*)malloc(sizeof(int) * dh);
    linear_coeffs(sw, dw, ox, cx);
    linear_coeffs(sh, dh, oy, cy);
    for (int y = 0; y < dh; y++) {
        const uint8_t* r0 = src + (size_t)oy[y] * sstride;
        const uint8_t* r1 = src + (size_t)(oy[y] + 1 < sh ? oy[y] + 1 : sh - 1) * sstride;
        int c1y = cy[y], c0y = 256 - c1y;
        for (int x = 0; x < dw; x++) {
            int o = ox[x], o1 = o + 1 < sw ? o + 1 : sw - 1;
            int c1x = cx[x], c0x = 256 - c1x;
            uint32_t t0 = (uint32_t)(c0x * r0[o] + c1x * r0[o1]);
            uint32_t t1 = (uint32_t)(c0x * r1[o] + c1x * r1[o1]);
            dst[(size_t)y * dstride + x] = (uint8_t)((c0y * t0 + c1y * t1 + 32768u) >> 16);
        }
    }
    free(ox); free(cx); free(oy); free(cy);
}

/* ---- A2: FAST-9/16 score, 3x3 strict NMS (OpenCV features2d fast.cpp / fast_score.cpp) ---- */
static const int ring_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int ring_dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

/* m = max over the 16 arcs of 9 contiguous ring pixels of max(min(d), min(-d)); corner iff m > t; score m-1 */
static int fast_arc_strength(const uint8_t* img, int stride, int x, int y)
{
    int d[16];
    int c = img[(size_t)y * stride + x];
    for (int k = 0; k < 16; k++) d[k] = c - img[(size_t)(y + ring_dy[k]) * stride + (x + ring_dx[k])];
    int best = -256;
    for (int k = 0; k < 16; k++) {
        int mn = 255, mx = -255;
        for (int j = 0; j < 9; j++) {
            int v = d[(k + j) & 15];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best;
}

/* score map: 0 for non-corners and the 3-pixel frame, else m-1 (1..254) */
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, int threshold, uint8_t* score /* w*h */)
{
    memset(score, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int m = fast_arc_strength(img, stride, x, y);
            if (m > threshold) score[(size_t)y * w + x] = (uint8_t)(m - 1);
        }
}

/* raster-ordered corners after NMS and the border filter [border, w-border) x [border, h-border) */
int orc_fast_nms(const uint8_t* score, int w, int h, int border, int32_t* xs, int32_t* ys, int32_t* sc, int cap)
{
    int n = 0;
    int b = border < 3 ? 3 : border;   /* FAST itself never looks at the outer 3 pixels */
    if (h <= 2 * border || w <= 2 * border) return 0;
    for (int y = b; y < h - b; y++)
        for (int x = b; x < w - b; x++) {
            int s = score[(size_t)y * w + x];
            if (!s) continue;
            const uint8_t* p = score + (size_t)y * w + x;
            if (s > p[-1] && s > p[1] && s > p[-w - 1] && s > p[-w] && s > p[-w + 1] && s > p[w - 1] && s > p[w] && s > p[w + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; sc[n] = s; }
                n++;
            }
        }
    return n;
}

/* ---- A4: Harris response on a 7x7 block (OpenCV orb.cpp HarrisResponses, blockSize 7, k 0.04) ---- */
float orc_harris(const uint8_t* img, int stride, int x, int y)
{
    int a = 0, b = 0, c = 0;
    for (int j = -3; j <= 3; j++)
        for (int i = -3; i <= 3; i++) {
            const uint8_t* p = img + (size_t)(y + j) * stride + (x + i);
            int Ix = (p[1] - p[-1]) * 2 + (p[-stride + 1] - p[-stride - 1]) + (p[stride + 1] - p[stride - 1]);
            int Iy = (p[stride] - p[-stride]) * 2 + (p[stride - 1] - p[-stride - 1]) + (p[stride + 1] - p[-stride + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    float scale = 1.f / ((1 << 2) * 7 * 255.f);
    float s4 = scale * scale * scale * scale;
    float fa = (float)a, fb = (float)b, fc = (float)c;
    return (fa * fb - fc * fc - 0.04f * (fa + fb) * (fa + fb)) * s4;
}

/* ---- A5: intensity-centroid orientation (OpenCV orb.cpp ICAngles + core fastAtan2 scalar) ---- */
static const int g_umax[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };

float orc_fast_atan2(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / 3.141592653589793238462643383279502884);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.141592653589793238462643383279502884);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.141592653589793238462643383279502884);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.141592653589793238462643383279502884);
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

void orc_ic_moments(const uint8_t* img, int stride, int x, int y, int* m01, int* m10)
{
    const uint8_t* c = img + (size_t)y * stride + x;
    int s01 = 0, s10 = 0;
    for (int u = -15; u <= 15; u++) s10 += u * c[u];
    for (int v = 1; v <= 15; v++) {
        int vs = 0, d = g_umax[v];
        for (int u = -d; u <= d; u++) {
            int vp = c[u + v * stride], vm = c[u - v * stride];
            vs += vp - vm;
            s10 += u * (vp + vm);
        }
        s01 += v * vs;
    }
    *m01 = s01; *m10 = s10;
}

float orc_ic_angle(const uint8_t* img, int stride, int x, int y)
{
    int m01, m10;
    orc_ic_moments(img, stride, x, y, &m01, &m10);
    return orc_fast_atan2((float)m01, (float)m10);
}

/* ---- A6: 7x7 sigma=2 float blur of a level, BORDER_REFLECT_101, rounded back to u8 ---- */
static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

void orc_gauss_kernel7(float* k)
{
    /* cv::getGaussianKernel(7, 2.0, CV_32F): exp(-(i-3)^2/(2 sigma^2)) normalised in double, stored as float */
    double t[7], sum = 0;
    for (int i = 0; i < 7; i++) { double x = i - 3; t[i] = exp(-0.5 * x * x / 4.0); sum += t[i]; }
    for (int i = 0; i < 7; i++) k[i] = (float)(t[i] / sum);
}

void orc_blur7(const uint8_t* img, int w, int h, int stride, uint8_t* out /* w*h */)
{
    float k[7];
    orc_gauss_kernel7(k);
    float* rows = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* r = img + (size_t)y * stride;
            float acc = k[0] * (float)r[reflect101(x - 3, w)];
            for (int j = 1; j < 7; j++) acc = fmaf(k[j], (float)r[reflect101(x - 3 + j, w)], acc);
            rows[(size_t)y * w + x] = acc;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float acc = k[3] * rows[(size_t)y * w + x];
            for (int j = 1; j <= 3; j++) {
                float s = rows[(size_t)reflect101(y + j, h) * w + x] + rows[(size_t)reflect101(y - j, h) * w + x];
                acc = fmaf(k[3 + j], s, acc);
            }
            int v = rne_f(acc);
            out[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(rows);
}

/* rotated BRIEF, WTA_K = 2 (OpenCV orb.cpp computeOrbDescriptors).  OpenCV blurs the level image in place inside a
 * larger buffer whose 32-pixel BORDER_REFLECT_101 frame stays un-blurred, so a sample that lands outside the level
 * (only possible for caller-provided keypoints close to the border of a coarse level) reads the raw reflected pixel. */
void orc_describe(const uint8_t* blur, const uint8_t* raw, int w, int h, int stride, int xi, int yi, float angle_deg, uint8_t* desc32)
{
    float ang = angle_deg * (float)(3.141592653589793238462643383279502884 / 180.f);
    float a = (float)cos((double)ang), b = (float)sin((double)ang);
    for (int i = 0; i < 32; i++) {
        int byte = 0;
        for (int k = 0; k < 8; k++) {
            const int* pt = g_pattern + (i * 8 + k) * 4;
            int v[2];
            for (int e = 0; e < 2; e++) {
                float px = (float)pt[2 * e], py = (float)pt[2 * e + 1];
                float xr = px * a - py * b;
                float yr = px * b + py * a;
                int sx = xi + rne_f(xr), sy = yi + rne_f(yr);
                if (sx >= 0 && sx < w && sy >= 0 && sy < h) v[e] = blur[(size_t)sy * stride + sx];
                else v[e] = raw[(size_t)reflect101(sy, h) * stride + reflect101(sx, w)];
            }
            byte |= (v[0] < v[1]) << k;
        }
        desc32[i] = (uint8_t)byte;
    }
}

/* ---- A3: retainBest(n): keep everything with response >= n-th largest (OpenCV keypoint.cpp) ---- */
static int cmp_float_desc(const void* a, const void* b)
{
    float x = *(const float*)a, y = *(const float*)b;
    return (x < y) - (x > y);
}

static int retain_best(float* resp, int* keep, int count, int n)
{
    /* keep[i] must be 1 for live entries on entry; returns number kept */
    int live = 0;
    for (int i = 0; i < count; i++) live += keep[i];
    if (n < 0 || live <= n) return live;
    if (n == 0) { memset(keep, 0, sizeof(int) * count); return 0; }
    float* tmp = (float*)malloc(sizeof(float) * live);
    int m = 0;
    for (int i = 0; i < count; i++) if (keep[i]) tmp[m++] = resp[i];
    qsort(tmp, m, sizeof(float), cmp_float_desc);
    float thr = tmp[n - 1];
    free(tmp);
    int kept = 0;
    for (int i = 0; i < count; i++) { if (keep[i] && !(resp[i] >= thr)) keep[i] = 0; kept += keep[i]; }
    return kept;
}

/* ---- whole pyramid ---- */
typedef struct {
    int nlevels;
    int w[ORC_MAX_LEVELS], h[ORC_MAX_LEVELS];
    float scale[ORC_MAX_LEVELS];
    uint8_t* img[ORC_MAX_LEVELS];   /* tightly packed, stride == w */
} orc_pyramid;

static void pyramid_build(const orc_params* p, const uint8_t* gray, int w, int h, size_t stride, orc_pyramid* pyr)
{
    pyr->nlevels = p->nlevels;
    for (int l = 0; l < p->nlevels; l++) {
        pyr->scale[l] = orc_level_scale(p, l);
        orc_level_size(p, w, h, l, &pyr->w[l], &pyr->h[l]);
        pyr->img[l] = (uint8_t*)malloc((size_t)pyr->w[l] * pyr->h[l] + 1);
        if (l == p->first_level) {
            for (int y = 0; y < h; y++) memcpy(pyr->img[l] + (size_t)y * pyr->w[l], gray + (size_t)y * stride, (size_t)w);
        } else {
            /* chained: level l is resized from level l-1 (first_level == 0 is the only supported configuration) */
            orc_resize_linear_exact(pyr->img[l - 1], pyr->w[l - 1], pyr->h[l - 1], pyr->w[l - 1],
                                    pyr->img[l], pyr->w[l], pyr->h[l], pyr->w[l]);
        }
    }
}

static void pyramid_free(orc_pyramid* pyr)
{
    for (int l = 0; l < pyr->nlevels; l++) free(pyr->img[l]);
}

/* export one pyramid level for stage-wise tests; returns 0 or -1 */
int orc_pyramid_level(const orc_params* p, const uint8_t* gray, int w, int h, size_t stride, int level, uint8_t* out)
{
    if (level < 0 || level >= p->nlevels || p->nlevels > ORC_MAX_LEVELS || p->first_level != 0) return -1;
    orc_pyramid pyr;
    pyramid_build(p, gray, w, h, stride, &pyr);
    memcpy(out, pyr.img[level], (size_t)pyr.w[level] * pyr.h[level]);
    pyramid_free(&pyr);
    return 0;
}

/*
 * detect (OpenCV orb.cpp computeKeyPoints): per level FAST -> border -> retainBest(2q | q) -> [Harris -> retainBest(q)]
 * -> octave/size -> IC angle -> pt *= scale.  Output in canonical order (octave, y, x).
 * Returns the number of keypoints found (may exceed cap; only the first cap are written), <0 on bad arguments.
 */
static int detect_on_pyramid(const orc_params* p, const orc_pyramid* pyr, orc_keypoint* out, int cap)
{
    int q[ORC_MAX_LEVELS];
    orc_level_quotas(p, q);
    int total = 0;
    for (int l = 0; l < p->nlevels; l++) {
        int w = pyr->w[l], h = pyr->h[l];
        const uint8_t* img = pyr->img[l];
        if (w < 7 || h < 7) continue;
        uint8_t* score = (uint8_t*)malloc((size_t)w * h);
        orc_fast_score_map(img, w, h, w, p->fast_threshold, score);
        int maxc = (w / 2 + 1) * (h / 2 + 1);
        int32_t* xs = (int32_t*)malloc(sizeof(int32_t) * maxc), *ys = (int32_t*)malloc(sizeof(int32_t) * maxc),
               *sc = (int32_t*)malloc(sizeof(int32_t) * maxc);
        int n = orc_fast_nms(score, w, h, p->edge_threshold, xs, ys, sc, maxc);
        float* resp = (float*)malloc(sizeof(float) * (n + 1));
        int* keep = (int*)malloc(sizeof(int) * (n + 1));
        for (int i = 0; i < n; i++) { resp[i] = (float)sc[i]; keep[i] = 1; }
        if (p->score_type == 0) {
            retain_best(resp, keep, n, 2 * q[l]);
            for (int i = 0; i < n; i++) if (keep[i]) resp[i] = orc_harris(img, w, xs[i], ys[i]);
            retain_best(resp, keep, n, q[l]);
        } else {
            retain_best(resp, keep, n, q[l]);
        }
        float sf = pyr->scale[l];
        for (int i = 0; i < n; i++) {     /* raster order == (y, x) order */
            if (!keep[i]) continue;
            if (total < cap) {
                orc_keypoint* k = &out[total];
                k->angle = orc_ic_angle(img, w, xs[i], ys[i]);
                k->x = (float)xs[i] * sf;
                k->y = (float)ys[i] * sf;
                k->size = (float)p->patch_size * sf;
                k->response = resp[i];
                k->octave = l;
                k->class_id = -1;
            }
            total++;
        }
        free(score); free(xs); free(ys); free(sc); free(resp); free(keep);
    }
    return total;
}

/*
 * compute (OpenCV orb.cpp detectAndCompute, useProvidedKeypoints): border filter on full-resolution coordinates,
 * stable regroup by octave, blur each level, rBRIEF with the keypoint's own angle.  kps is rewritten in place with the
 * surviving, regrouped keypoints; returns their number.
 */
static int compute_on_pyramid(const orc_params* p, const orc_pyramid* pyr, int w, int h, orc_keypoint* kps, int n, uint8_t* desc)
{
    int b = p->edge_threshold;
    orc_keypoint* tmp = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (n + 1));
    int m = 0;
    if (!(h <= 2 * b || w <= 2 * b)) {
        for (int l = 0; l < p->nlevels; l++)
            for (int i = 0; i < n; i++) {
                const orc_keypoint* k = &kps[i];
                if (k->octave != l) continue;
                /* cv::Rect(int).contains(Point2f -> Point): coordinates are cvRound-ed first */
                int xi = rne_f(k->x), yi = rne_f(k->y);
                if (xi >= b && xi < w - b && yi >= b && yi < h - b) tmp[m++] = *k;
            }
    }
    memcpy(kps, tmp, sizeof(orc_keypoint) * m);
    free(tmp);
    uint8_t* blur[ORC_MAX_LEVELS] = { 0 };
    for (int i = 0; i < m; i++) {
        int l = kps[i].octave;
        if (!blur[l]) {
            blur[l] = (uint8_t*)malloc((size_t)pyr->w[l] * pyr->h[l]);
            orc_blur7(pyr->img[l], pyr->w[l], pyr->h[l], pyr->w[l], blur[l]);
        }
        float scale = 1.f / pyr->scale[l];
        int xi = rne_f(kps[i].x * scale), yi = rne_f(kps[i].y * scale);
        orc_describe(blur[l], pyr->img[l], pyr->w[l], pyr->h[l], pyr->w[l], xi, yi, kps[i].angle, desc + (size_t)i * 32);
    }
    for (int l = 0; l < ORC_MAX_LEVELS; l++) free(blur[l]);
    return m;
}

static int params_ok(const orc_params* p)
{
    return p->nlevels >= 1 && p->nlevels <= ORC_MAX_LEVELS && p->first_level == 0 && p->wta_k == 2 && p->patch_size == 31 &&
           p->edge_threshold >= 31 && p->nfeatures >= 0 && p->scale_factor > 1.0f;
}

int orc_detect(const orc_params* p, const uint8_t* gray, int w, int h, size_t stride, orc_keypoint* out, int cap)
{
    if (!params_ok(p)) return -1;
    orc_pyramid pyr;
    pyramid_build(p, gray, w, h, stride, &pyr);
    int n = detect_on_pyramid(p, &pyr, out, cap);
    pyramid_free(&pyr);
    return n;
}

int orc_compute(const orc_params* p, const uint8_t* gray, int w, int h, size_t stride, orc_keypoint* kps, int n, uint8_t* desc)
{
    if (!params_ok(p)) return -1;
    for (int i = 0; i < n; i++) if (kps[i].octave < 0 || kps[i].octave >= p->nlevels) return -2;
    orc_pyramid pyr;
    pyramid_build(p, gray, w, h, stride, &pyr);   /* the reference path rebuilds the pyramid too */
    int m = compute_on_pyramid(p, &pyr, w, h, kps, n, desc);
    pyramid_free(&pyr);
    return m;
}

/* stage-wise helpers for kernel-level parity tests */
int orc_level_fast(const orc_params* p, const uint8_t* gray, int w, int h, size_t stride, int level,
                   int32_t* xs, int32_t* ys, int32_t* sc, int cap)
{
    if (!params_ok(p) || level < 0 || level >= p->nlevels) return -1;
    orc_pyramid pyr;
    pyramid_build(p, gray, w, h, stride, &pyr);
    int lw = pyr.w[level], lh = pyr.h[level];
    uint8_t* score = (uint8_t*)malloc((size_t)lw * lh);
    orc_fast_score_map(pyr.img[level], lw, lh, lw, p->fast_threshold, score);
    int n = orc_fast_nms(score, lw, lh, p->edge_threshold, xs, ys, sc, cap);
    free(score);
    pyramid_free(&pyr);
    return n;
}

/* ---- A7: brute-force Hamming kNN(k=2), ties -> lower train index (OpenCV batchDistance K=2; FORB.cpp:81-101) ---- */
static inline int hamming256(const uint8_t* a, const uint8_t* b)
{
    const uint64_t* x = (const uint64_t*)a;
    const uint64_t* y = (const uint64_t*)b;
    uint64_t w0, w1, w2, w3, v0, v1, v2, v3;
    memcpy(&w0, x, 8); memcpy(&w1, x + 1, 8); memcpy(&w2, x + 2, 8); memcpy(&w3, x + 3, 8);
    memcpy(&v0, y, 8); memcpy(&v1, y + 1, 8); memcpy(&v2, y + 2, 8); memcpy(&v3, y + 3, 8);
    return __builtin_popcountll(w0 ^ v0) + __builtin_popcountll(w1 ^ v1) + __builtin_popcountll(w2 ^ v2) + __builtin_popcountll(w3 ^ v3);
}

/* idx/dist are nq x 2; missing entries (nt < 2) are idx -1, dist -1.  Sequential strict-< insertion in train order. */
void orc_knn2(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int32_t* idx, int32_t* dist)
{
    for (int64_t i = 0; i < nq; i++) {
        int d0 = INT32_MAX, d1 = INT32_MAX, i0 = -1, i1 = -1;
        const uint8_t* qi = q + i * 32;
        for (int64_t j = 0; j < nt; j++) {
            int d = hamming256(qi, t + j * 32);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = (int)j; }
            else if (d < d1) { d1 = d; i1 = (int)j; }
        }
        idx[2 * i] = i0; dist[2 * i] = i0 < 0 ? -1 : d0;
        idx[2 * i + 1] = i1; dist[2 * i + 1] = i1 < 0 ? -1 : d1;
    }
}

/* Lowe ratio test exactly as src/CameraPoseEstimator.cpp:208-212 (float multiply, strict <); rows with <2 neighbours
 * are skipped (the reference would read out of bounds there).  Returns number of accepted queries (ascending order). */
int64_t orc_ratio_test(const int32_t* idx, const int32_t* dist, int64_t nq, float ratio, int32_t* good_q, int32_t* good_t, int32_t* good_d)
{
    int64_t n = 0;
    for (int64_t i = 0; i < nq; i++) {
        if (idx[2 * i] < 0 || idx[2 * i + 1] < 0) continue;
        float d0 = (float)dist[2 * i], d1 = (float)dist[2 * i + 1];
        if (d0 < d1 * ratio) { good_q[n] = (int32_t)i; good_t[n] = idx[2 * i]; good_d[n] = dist[2 * i]; n++; }
    }
    return n;
}
