/* CPU restatement of the decoder behind the reference's frame loader for baseline JPEG files (grey-scale first; YCbCr colour in
 * the second half of this file).
 *
 * TEST INFRASTRUCTURE ONLY (see orb_oracle.c).  The reference reads every frame with cv::imread (src/FrameLoader.cpp:62); for
 * .jpg files that is OpenCV's JPEG decoder, i.e. libjpeg -- a dependency that is neither in /root/reference nor vendored
 * there.  The runnable instance of that call in this container is cv2 4.13.0 built against libjpeg-turbo 3.1.2
 * (cv2.getBuildInformation()), default settings: DCT method JDCT_ISLOW.  This file restates the published baseline algorithm
 * that path executes for a one-component, 8-bit, Huffman-coded sequential file:
 *   marker parsing                        ITU-T T.81 Annex B; libjpeg jdmarker.c (get_sof, get_dht, get_dqt, get_dri, get_sos)
 *   Huffman decoding of a block           T.81 F.2.2; libjpeg jdhuff.c decode_mcu_slow (HUFF_EXTEND, jpeg_natural_order)
 *   restart intervals                     T.81 F.2.2.4 / E.2.4; jdhuff.c process_restart (DC predictor reset, byte alignment)
 *   dequantisation + inverse DCT          libjpeg jidctint.c jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2)
 *   range limiting                        libjpeg jdmaster.c prepare_range_limit_table, indexed with & RANGE_MASK
 * and is pinned against cv2.imdecode(..., IMREAD_UNCHANGED) on the committed files of tests/golden/jpeg_cases.npz
 * (tests/golden/make_golden_jpeg.py; tests/test_oracle_jpeg.py: every pixel equal).
 *
 * Anything else (progressive, arithmetic coding, 12-bit, samplings other than 4:2:0 / 4:2:2 / 4:4:4) returns
 * ORC_JPEG_UNSUPPORTED: the caller keeps using its CPU decoder for those files. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_JPEG_OK = 0, ORC_JPEG_UNSUPPORTED = -1, ORC_JPEG_CORRUPT = -2 };

static const uint8_t jpeg_natural_order[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

typedef struct {
    int present;
    uint8_t bits[17];       /* bits[l] = number of codes of length l */
    uint8_t vals[256];
    /* canonical decoding (T.81 F.2.2.3): mincode / maxcode / valptr per length */
    int32_t mincode[17], maxcode[18], valptr[17];
} huff_table;

typedef struct {
    int width, height, restart_interval;
    uint16_t quant[4][64];  /* natural order */
    int quant_present[4];
    huff_table dc[4], ac[4];
    int tq, td, ta;         /* table selectors of the single component */
    const uint8_t* scan;    /* entropy-coded data */
    size_t scan_len;
} jpeg_info;

static void huff_build(huff_table* t)
{
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        t->valptr[l] = k;
        t->mincode[l] = code;
        code += t->bits[l];
        k += t->bits[l];
        t->maxcode[l] = t->bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    t->maxcode[17] = 0x7fffffff;
}

/* Parse up to the start of the entropy-coded segment.  Returns ORC_JPEG_OK for a file this decoder handles. */
static int jpeg_parse(const uint8_t* f, size_t n, jpeg_info* ji)
{
    memset(ji, 0, sizeof(*ji));
    if (n < 4 || f[0] != 0xFF || f[1] != 0xD8) return ORC_JPEG_CORRUPT;
    size_t p = 2;
    int have_sof = 0;
    for (;;) {
        if (p + 4 > n) return ORC_JPEG_CORRUPT;
        if (f[p] != 0xFF) return ORC_JPEG_CORRUPT;
        while (p < n && f[p] == 0xFF) p++;          /* fill bytes */
        if (p >= n) return ORC_JPEG_CORRUPT;
        const int m = f[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return ORC_JPEG_CORRUPT;     /* EOI before SOS */
        if (p + 2 > n) return ORC_JPEG_CORRUPT;
        const size_t len = ((size_t)f[p] << 8) | f[p + 1];
        if (len < 2 || p + len > n) return ORC_JPEG_CORRUPT;
        const uint8_t* s = f + p + 2;
        const size_t sl = len - 2;
        if (m == 0xC0 || m == 0xC1) {               /* baseline / extended sequential, Huffman */
            if (sl < 6) return ORC_JPEG_CORRUPT;
            if (s[0] != 8) return ORC_JPEG_UNSUPPORTED;
            ji->height = (s[1] << 8) | s[2];
            ji->width = (s[3] << 8) | s[4];
            if (s[5] != 1) return ORC_JPEG_UNSUPPORTED;          /* one component only */
            if (sl < 9 || ji->width == 0 || ji->height == 0) return ORC_JPEG_CORRUPT;
            ji->tq = s[8] & 3;
            have_sof = 1;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return ORC_JPEG_UNSUPPORTED;            /* progressive, lossless, arithmetic, hierarchical */
        } else if (m == 0xC4) {                     /* DHT, possibly several tables */
            size_t q = 0;
            while (q < sl) {
                if (q + 17 > sl) return ORC_JPEG_CORRUPT;
                const int tc = s[q] >> 4, th = s[q] & 15;
                if (tc > 1 || th > 3) return ORC_JPEG_CORRUPT;
                huff_table* t = tc ? &ji->ac[th] : &ji->dc[th];
                int count = 0;
                t->bits[0] = 0;
                for (int l = 1; l <= 16; l++) { t->bits[l] = s[q + l]; count += t->bits[l]; }
                q += 17;
                if (count > 256 || q + (size_t)count > sl) return ORC_JPEG_CORRUPT;
                memset(t->vals, 0, sizeof(t->vals));
                memcpy(t->vals, s + q, (size_t)count);
                q += (size_t)count;
                t->present = 1;
                huff_build(t);
            }
        } else if (m == 0xDB) {                     /* DQT */
            size_t q = 0;
            while (q < sl) {
                const int pq = s[q] >> 4, tq = s[q] & 15;
                if (tq > 3 || pq > 1) return ORC_JPEG_CORRUPT;
                q++;
                if (q + (size_t)(pq ? 128 : 64) > sl) return ORC_JPEG_CORRUPT;
                for (int i = 0; i < 64; i++) {
                    const int v = pq ? ((s[q] << 8) | s[q + 1]) : s[q];
                    q += pq ? 2 : 1;
                    ji->quant[tq][jpeg_natural_order[i]] = (uint16_t)v;
                }
                ji->quant_present[tq] = 1;
            }
        } else if (m == 0xDD) {                     /* DRI */
            if (sl < 2) return ORC_JPEG_CORRUPT;
            ji->restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {                     /* SOS */
            if (!have_sof) return ORC_JPEG_CORRUPT;
            if (sl < 6 || s[0] != 1) return ORC_JPEG_UNSUPPORTED;
            ji->td = s[2] >> 4;
            ji->ta = s[2] & 15;
            if (s[3] != 0 || s[4] != 63 || s[5] != 0) return ORC_JPEG_UNSUPPORTED;
            if (ji->td > 3 || ji->ta > 3 || !ji->dc[ji->td].present || !ji->ac[ji->ta].present || !ji->quant_present[ji->tq]) return ORC_JPEG_CORRUPT;
            ji->scan = f + p + len;
            ji->scan_len = n - (p + len);
            return ORC_JPEG_OK;
        }
        p += len;
    }
}

/* bit reader over one restart interval: stuffed zero bytes removed, zero bits past the end (libjpeg's behaviour on a short segment) */
typedef struct { const uint8_t* p; const uint8_t* end; uint32_t buf; int nbits; } bitreader;

static void br_fill(bitreader* b)
{
    while (b->nbits <= 24) {
        uint32_t byte = 0;
        if (b->p < b->end) {
            byte = *b->p++;
            if (byte == 0xFF) {
                if (b->p < b->end && *b->p == 0x00) b->p++;         /* stuffed byte */
                else { b->p = b->end; byte = 0; }                   /* a marker: no more data in this interval */
            }
        }
        b->buf |= byte << (24 - b->nbits);
        b->nbits += 8;
    }
}
static int br_get(bitreader* b, int n)
{
    if (n == 0) return 0;
    br_fill(b);
    const int v = (int)(b->buf >> (32 - n));
    b->buf <<= n;
    b->nbits -= n;
    return v;
}
static int huff_decode(bitreader* b, const huff_table* t)
{
    br_fill(b);
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | (int)(b->buf >> 31);
        b->buf <<= 1;
        b->nbits--;
        if (code <= t->maxcode[l]) return t->vals[(t->valptr[l] + code - t->mincode[l]) & 255];
        if (b->nbits < 8) br_fill(b);
    }
    return 0;       /* a code longer than 16 bits: corrupt data, libjpeg substitutes zero */
}
#define HUFF_EXTEND(x, s) ((x) < (1 << ((s) - 1)) ? (x) + (int)((~0u) << (s)) + 1 : (x))

/* jidctint.c jpeg_idct_islow */
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

static uint8_t range_limit(int x)
{
    x &= 1023;                                      /* RANGE_MASK; the table of prepare_range_limit_table past CENTERJSAMPLE */
    if (x < 128) return (uint8_t)(x + 128);
    if (x < 512) return 255;
    if (x < 896) return 0;
    return (uint8_t)(x - 896);
}

static void idct_islow(const int16_t* coef, const uint16_t* quant, uint8_t* out, int stride, int cols, int rows)
{
    int ws[64];
    for (int c = 0; c < 8; c++) {
        const int in0 = coef[c] * quant[c], in1 = coef[8 + c] * quant[8 + c], in2 = coef[16 + c] * quant[16 + c], in3 = coef[24 + c] * quant[24 + c];
        const int in4 = coef[32 + c] * quant[32 + c], in5 = coef[40 + c] * quant[40 + c], in6 = coef[48 + c] * quant[48 + c], in7 = coef[56 + c] * quant[56 + c];
        int z2 = in2, z3 = in6;
        int z1 = (z2 + z3) * FIX_0_541196100;
        int tmp2 = z1 + z3 * (-FIX_1_847759065);
        int tmp3 = z1 + z2 * FIX_0_765366865;
        z2 = in0; z3 = in4;
        int tmp0 = (int)((unsigned)(z2 + z3) << 13);
        int tmp1 = (int)((unsigned)(z2 - z3) << 13);
        const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = in7; tmp1 = in5; tmp2 = in3; tmp3 = in1;
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int z4 = tmp1 + tmp3;
        const int z5 = (z3 + z4) * FIX_1_175875602;
        tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        ws[c] = DESCALE(tmp10 + tmp3, 11);      ws[56 + c] = DESCALE(tmp10 - tmp3, 11);
        ws[8 + c] = DESCALE(tmp11 + tmp2, 11);  ws[48 + c] = DESCALE(tmp11 - tmp2, 11);
        ws[16 + c] = DESCALE(tmp12 + tmp1, 11); ws[40 + c] = DESCALE(tmp12 - tmp1, 11);
        ws[24 + c] = DESCALE(tmp13 + tmp0, 11); ws[32 + c] = DESCALE(tmp13 - tmp0, 11);
    }
    for (int r = 0; r < 8; r++) {
        const int* w = ws + 8 * r;
        int z2 = w[2], z3 = w[6];
        int z1 = (z2 + z3) * FIX_0_541196100;
        int tmp2 = z1 + z3 * (-FIX_1_847759065);
        int tmp3 = z1 + z2 * FIX_0_765366865;
        int tmp0 = (int)((unsigned)(w[0] + w[4]) << 13);
        int tmp1 = (int)((unsigned)(w[0] - w[4]) << 13);
        const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int z4 = tmp1 + tmp3;
        const int z5 = (z3 + z4) * FIX_1_175875602;
        tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int o[8] = {DESCALE(tmp10 + tmp3, 18), DESCALE(tmp11 + tmp2, 18), DESCALE(tmp12 + tmp1, 18), DESCALE(tmp13 + tmp0, 18),
                          DESCALE(tmp13 - tmp0, 18), DESCALE(tmp12 - tmp1, 18), DESCALE(tmp11 - tmp2, 18), DESCALE(tmp10 - tmp3, 18)};
        if (r < rows)
            for (int c = 0; c < cols; c++) out[(size_t)r * stride + c] = range_limit(o[c]);
    }
}

/* info[4] = {width, height, restart interval in MCUs (0: none), number of 8x8 blocks} */
int orc_jpeg_probe(const uint8_t* file, size_t size, int32_t* info)
{
    jpeg_info ji;
    const int rc = jpeg_parse(file, size, &ji);
    if (rc) return rc;
    info[0] = ji.width; info[1] = ji.height; info[2] = ji.restart_interval;
    info[3] = ((ji.width + 7) / 8) * ((ji.height + 7) / 8);
    return ORC_JPEG_OK;
}

/* Decode into out (rows of `stride` bytes, at least width x height). */
int orc_jpeg_decode_gray(const uint8_t* file, size_t size, uint8_t* out, int stride)
{
    jpeg_info ji;
    const int rc = jpeg_parse(file, size, &ji);
    if (rc) return rc;
    const int bw = (ji.width + 7) / 8, bh = (ji.height + 7) / 8, nblocks = bw * bh;
    const int ri = ji.restart_interval ? ji.restart_interval : nblocks;
    const uint8_t* p = ji.scan;
    const uint8_t* end = ji.scan + ji.scan_len;
    const huff_table* dct = &ji.dc[ji.td];
    const huff_table* act = &ji.ac[ji.ta];
    for (int b0 = 0; b0 < nblocks; b0 += ri) {
        /* this interval's bytes end at the next marker (RSTn, or EOI for the last one) */
        const uint8_t* q = p;
        while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00)) q++;
        if (q + 1 >= end) q = end;
        bitreader br = {p, q, 0, 0};
        int dc = 0;
        const int b1 = b0 + ri < nblocks ? b0 + ri : nblocks;
        for (int b = b0; b < b1; b++) {
            int16_t coef[64];
            memset(coef, 0, sizeof(coef));
            int s = huff_decode(&br, dct);
            if (s) { const int r = br_get(&br, s & 15); s = HUFF_EXTEND(r, s & 15); }
            dc += s;
            coef[0] = (int16_t)dc;
            for (int k = 1; k < 64; k++) {
                const int rs = huff_decode(&br, act);
                const int r = rs >> 4, sz = rs & 15;
                if (sz) {
                    k += r;
                    const int v = br_get(&br, sz);
                    if (k > 63) break;              /* corrupt data */
                    coef[jpeg_natural_order[k]] = (int16_t)HUFF_EXTEND(v, sz);
                } else {
                    if (r != 15) break;
                    k += 15;
                }
            }
            const int bx = b % bw, by = b / bw;
            const int cols = ji.width - bx * 8 < 8 ? ji.width - bx * 8 : 8, rows = ji.height - by * 8 < 8 ? ji.height - by * 8 : 8;
            idct_islow(coef, ji.quant[ji.tq], out + (size_t)by * 8 * stride + bx * 8, stride, cols, rows);
        }
        /* skip the marker that ended the interval */
        p = q;
        if (p + 1 < end && p[0] == 0xFF) {
            while (p < end && *p == 0xFF) p++;
            if (p < end) p++;
        }
    }
    return ORC_JPEG_OK;
}


/* ---- colour: three-component YCbCr files (JFIF), interleaved scan, luma sampling 2x2 (4:2:0), 2x1 (4:2:2) or 1x1 (4:4:4) with
 * 1x1 chroma.  What OpenCV returns for them (imread IMREAD_UNCHANGED: BGR) is libjpeg's default output path:
 *   interleaved MCU decoding              T.81 A.2.3; jdhuff.c decode_mcu (one DC predictor per component), jdcoefct.c
 *   chroma upsampling                     jdsample.c h2v2_fancy_upsample / h2v1_fancy_upsample (do_fancy_upsampling, the default);
 *                                         context rows at the top and bottom replicate the first / last real row (jdmainct.c)
 *   colour conversion                     jdcolor.c ycc_rgb_convert (16-bit fixed point tables, SCALEBITS 16), range-limited
 * pinned against cv2.imdecode on the colour files of tests/golden/jpeg_cases.npz.
 * orc_jpeg_decode_planes: the component planes after the inverse DCT, each padded to whole MCUs; planes[c] must hold the
 * pw[c] * ph[c] bytes orc_jpeg_probe_colour reports. */
typedef struct { int h, v, tq, td, ta; } jcomp;
typedef struct {
    int width, height, restart_interval, ncomp, hmax, vmax;
    jcomp comp[3];
    uint16_t quant[4][64];
    huff_table dc[4], ac[4];
    const uint8_t* scan; size_t scan_len;
} jpeg_cinfo;

static int jpeg_parse_colour(const uint8_t* f, size_t n, jpeg_cinfo* ji)
{
    memset(ji, 0, sizeof(*ji));
    if (n < 4 || f[0] != 0xFF || f[1] != 0xD8) return ORC_JPEG_CORRUPT;
    size_t p = 2;
    int have_sof = 0, qp[4] = {0, 0, 0, 0}, cid[3] = {0, 0, 0};
    for (;;) {
        if (p + 4 > n || f[p] != 0xFF) return ORC_JPEG_CORRUPT;
        while (p < n && f[p] == 0xFF) p++;
        if (p >= n) return ORC_JPEG_CORRUPT;
        const int m = f[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return ORC_JPEG_CORRUPT;
        if (p + 2 > n) return ORC_JPEG_CORRUPT;
        const size_t len = ((size_t)f[p] << 8) | f[p + 1];
        if (len < 2 || p + len > n) return ORC_JPEG_CORRUPT;
        const uint8_t* s = f + p + 2;
        const size_t sl = len - 2;
        if (m == 0xC0 || m == 0xC1) {
            if (sl < 6 || s[0] != 8) return ORC_JPEG_UNSUPPORTED;
            ji->height = (s[1] << 8) | s[2];
            ji->width = (s[3] << 8) | s[4];
            ji->ncomp = s[5];
            if (ji->ncomp != 3 || sl < 6 + 9 || ji->width == 0 || ji->height == 0) return ORC_JPEG_UNSUPPORTED;
            for (int c = 0; c < 3; c++) {
                cid[c] = s[6 + 3 * c];
                ji->comp[c].h = s[7 + 3 * c] >> 4; ji->comp[c].v = s[7 + 3 * c] & 15; ji->comp[c].tq = s[8 + 3 * c] & 3;
                if (ji->comp[c].h > ji->hmax) ji->hmax = ji->comp[c].h;
                if (ji->comp[c].v > ji->vmax) ji->vmax = ji->comp[c].v;
            }
            have_sof = 1;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return ORC_JPEG_UNSUPPORTED;
        } else if (m == 0xC4) {
            size_t q = 0;
            while (q < sl) {
                if (q + 17 > sl) return ORC_JPEG_CORRUPT;
                const int tc = s[q] >> 4, th = s[q] & 15;
                if (tc > 1 || th > 3) return ORC_JPEG_CORRUPT;
                huff_table* t = tc ? &ji->ac[th] : &ji->dc[th];
                int count = 0;
                t->bits[0] = 0;
                for (int l = 1; l <= 16; l++) { t->bits[l] = s[q + l]; count += t->bits[l]; }
                q += 17;
                if (count > 256 || q + (size_t)count > sl) return ORC_JPEG_CORRUPT;
                memset(t->vals, 0, sizeof(t->vals));
                memcpy(t->vals, s + q, (size_t)count);
                q += (size_t)count;
                t->present = 1;
                huff_build(t);
            }
        } else if (m == 0xDB) {
            size_t q = 0;
            while (q < sl) {
                const int pq = s[q] >> 4, tq = s[q] & 15;
                if (tq > 3 || pq > 1) return ORC_JPEG_CORRUPT;
                q++;
                if (q + (size_t)(pq ? 128 : 64) > sl) return ORC_JPEG_CORRUPT;
                for (int i = 0; i < 64; i++) {
                    const int v = pq ? ((s[q] << 8) | s[q + 1]) : s[q];
                    q += pq ? 2 : 1;
                    ji->quant[tq][jpeg_natural_order[i]] = (uint16_t)v;
                }
                qp[tq] = 1;
            }
        } else if (m == 0xDD) {
            if (sl < 2) return ORC_JPEG_CORRUPT;
            ji->restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {
            if (!have_sof) return ORC_JPEG_CORRUPT;
            if (sl < 4 + 6 || s[0] != 3) return ORC_JPEG_UNSUPPORTED;
            for (int c = 0; c < 3; c++) {
                if (s[1 + 2 * c] != cid[c]) return ORC_JPEG_UNSUPPORTED;
                ji->comp[c].td = s[2 + 2 * c] >> 4; ji->comp[c].ta = s[2 + 2 * c] & 15;
                if (ji->comp[c].td > 3 || ji->comp[c].ta > 3 || !ji->dc[ji->comp[c].td].present || !ji->ac[ji->comp[c].ta].present || !qp[ji->comp[c].tq]) return ORC_JPEG_CORRUPT;
            }
            if (s[7] != 0 || s[8] != 63 || s[9] != 0) return ORC_JPEG_UNSUPPORTED;
            if (ji->comp[1].h != 1 || ji->comp[1].v != 1 || ji->comp[2].h != 1 || ji->comp[2].v != 1) return ORC_JPEG_UNSUPPORTED;
            if (!((ji->comp[0].h == 2 && ji->comp[0].v == 2) || (ji->comp[0].h == 2 && ji->comp[0].v == 1) || (ji->comp[0].h == 1 && ji->comp[0].v == 1)))
                return ORC_JPEG_UNSUPPORTED;
            ji->scan = f + p + len;
            ji->scan_len = n - (p + len);
            return ORC_JPEG_OK;
        }
        p += len;
    }
}

/* info[12] = {width, height, restart interval, hmax, vmax, then per component: plane width, plane height} (+ h0 v0 of luma at [11]?) */
int orc_jpeg_probe_colour(const uint8_t* file, size_t size, int32_t* info)
{
    jpeg_cinfo ji;
    const int rc = jpeg_parse_colour(file, size, &ji);
    if (rc) return rc;
    const int mx = (ji.width + 8 * ji.hmax - 1) / (8 * ji.hmax), my = (ji.height + 8 * ji.vmax - 1) / (8 * ji.vmax);
    info[0] = ji.width; info[1] = ji.height; info[2] = ji.restart_interval; info[3] = ji.hmax; info[4] = ji.vmax;
    for (int c = 0; c < 3; c++) { info[5 + 2 * c] = mx * ji.comp[c].h * 8; info[6 + 2 * c] = my * ji.comp[c].v * 8; }
    info[11] = ji.comp[0].h * 16 + ji.comp[0].v;
    return ORC_JPEG_OK;
}

int orc_jpeg_decode_planes(const uint8_t* file, size_t size, uint8_t* p0, uint8_t* p1, uint8_t* p2)
{
    jpeg_cinfo ji;
    const int rc = jpeg_parse_colour(file, size, &ji);
    if (rc) return rc;
    uint8_t* planes[3] = {p0, p1, p2};
    const int mx = (ji.width + 8 * ji.hmax - 1) / (8 * ji.hmax), my = (ji.height + 8 * ji.vmax - 1) / (8 * ji.vmax);
    const int nmcu = mx * my, ri = ji.restart_interval ? ji.restart_interval : nmcu;
    const uint8_t* p = ji.scan;
    const uint8_t* end = ji.scan + ji.scan_len;
    for (int m0 = 0; m0 < nmcu; m0 += ri) {
        const uint8_t* q = p;
        while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00)) q++;
        if (q + 1 >= end) q = end;
        bitreader br = {p, q, 0, 0};
        int dc[3] = {0, 0, 0};
        const int m1 = m0 + ri < nmcu ? m0 + ri : nmcu;
        for (int m = m0; m < m1; m++) {
            const int mcx = m % mx, mcy = m / mx;
            for (int c = 0; c < 3; c++) {
                const int stride = mx * ji.comp[c].h * 8;
                for (int by = 0; by < ji.comp[c].v; by++)
                    for (int bx = 0; bx < ji.comp[c].h; bx++) {
                        int16_t coef[64];
                        memset(coef, 0, sizeof(coef));
                        int s = huff_decode(&br, &ji.dc[ji.comp[c].td]);
                        if (s) { const int r = br_get(&br, s & 15); s = HUFF_EXTEND(r, s & 15); }
                        dc[c] += s;
                        coef[0] = (int16_t)dc[c];
                        for (int k = 1; k < 64; k++) {
                            const int rs = huff_decode(&br, &ji.ac[ji.comp[c].ta]);
                            const int r = rs >> 4, sz = rs & 15;
                            if (sz) {
                                k += r;
                                const int v = br_get(&br, sz);
                                if (k > 63) break;
                                coef[jpeg_natural_order[k]] = (int16_t)HUFF_EXTEND(v, sz);
                            } else {
                                if (r != 15) break;
                                k += 15;
                            }
                        }
                        const int X = (mcx * ji.comp[c].h + bx) * 8, Y = (mcy * ji.comp[c].v + by) * 8;
                        idct_islow(coef, ji.quant[ji.comp[c].tq], planes[c] + (size_t)Y * stride + X, stride, 8, 8);
                    }
            }
        }
        p = q;
        if (p + 1 < end && p[0] == 0xFF) {
            while (p < end && *p == 0xFF) p++;
            if (p < end) p++;
        }
    }
    return ORC_JPEG_OK;
}

static uint8_t clamp255(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

/* one upsampled chroma sample at output position (x, y); cw x ch = the component's real (downsampled) size */
static int chroma_at(const uint8_t* pl, int stride, int cw, int ch, int x, int y, int hs, int vs)
{
    if (hs == 1 && vs == 1) return pl[(size_t)y * stride + x];
    const int cx = x >> 1;
    /* jdsample.c jinit_upsampler: the fancy filters are chosen only for components more than two samples wide; narrower
     * ones are replicated (h2v1_upsample / h2v2_upsample) */
    if (cw <= 2) return pl[(size_t)(vs == 2 ? y >> 1 : y) * stride + cx];
    if (vs == 1) {                                  /* h2v1_fancy_upsample */
        const uint8_t* r = pl + (size_t)y * stride;
        if (x & 1) return cx == cw - 1 ? r[cx] : (3 * r[cx] + r[cx + 1] + 2) >> 2;
        return cx == 0 ? r[0] : (3 * r[cx] + r[cx - 1] + 1) >> 2;
    }
    /* h2v2_fancy_upsample: the nearer row weighs 3, the row above (even output rows) or below (odd) 1; then the same across */
    const int cy = y >> 1;
    int oy = (y & 1) ? cy + 1 : cy - 1;
    if (oy < 0) oy = 0;
    if (oy > ch - 1) oy = ch - 1;
    const uint8_t* r0 = pl + (size_t)cy * stride;
    const uint8_t* r1 = pl + (size_t)oy * stride;
    const int cur = 3 * r0[cx] + r1[cx];
    if (x & 1) {
        if (cx == cw - 1) return (cur * 4 + 7) >> 4;
        return (3 * cur + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
    }
    if (cx == 0) return (cur * 4 + 8) >> 4;
    return (3 * cur + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

/* out: height rows of `stride` bytes, 3 bytes (B, G, R) per pixel */
int orc_jpeg_decode_bgr(const uint8_t* file, size_t size, uint8_t* out, int stride)
{
    jpeg_cinfo ji;
    int rc = jpeg_parse_colour(file, size, &ji);
    if (rc) return rc;
    int32_t info[12];
    orc_jpeg_probe_colour(file, size, info);
    uint8_t* pl[3];
    for (int c = 0; c < 3; c++) pl[c] = (uint8_t*)malloc((size_t)info[5 + 2 * c] * info[6 + 2 * c]);
    rc = orc_jpeg_decode_planes(file, size, pl[0], pl[1], pl[2]);
    if (!rc) {
        const int hs = ji.comp[0].h, vs = ji.comp[0].v;
        const int cw = (ji.width + hs - 1) / hs, ch = (ji.height + vs - 1) / vs;
        for (int y = 0; y < ji.height; y++)
            for (int x = 0; x < ji.width; x++) {
                const int Y = pl[0][(size_t)y * info[5] + x];
                const int cb = chroma_at(pl[1], info[7], cw, ch, x, y, hs, vs) - 128, cr = chroma_at(pl[2], info[9], cw, ch, x, y, hs, vs) - 128;
                /* jdcolor.c build_ycc_rgb_table: FIX(x) = x * 65536 + 0.5 */
                const int r = Y + ((91881 * cr + 32768) >> 16);
                const int b = Y + ((116130 * cb + 32768) >> 16);
                const int g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
                uint8_t* o = out + (size_t)y * stride + 3 * x;
                o[0] = clamp255(b); o[1] = clamp255(g); o[2] = clamp255(r);
            }
    }
    for (int c = 0; c < 3; c++) free(pl[c]);
    return rc;
}
