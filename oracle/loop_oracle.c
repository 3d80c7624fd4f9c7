/*
 * loop_oracle.c -- CPU restatement of the reference's loop-closure candidate scoring.  TEST INFRASTRUCTURE ONLY (see the
 * header of orb_oracle.c): the checker of hamx_nbest_dev / hamx_loop_score_dev, never part of the product.
 *
 * What it follows (reference file:line):
 *   LoopCloser::NBestMatches   src/LoopCloser.cpp:53-105   for every row of descriptors1 the n best rows of descriptors2:
 *                              a sorted list filled by insertion -- a candidate is placed before the first entry it is
 *                              strictly smaller than and pushes the rest down (:86-101), so equal distances keep the lower
 *                              train index first.
 *   LoopCloser::DetectLoop     src/LoopCloser.cpp:19-51    the current frame against every stored frame: count the n-best
 *                              distances below dist_thr (:34-41), keep the frame with the strictly largest count (:42-46).
 *
 * Two defects of the reference are NOT reproduced, because they make the routine meaningless on ORB data (SURVEY.md 8f
 * rank 4, "as the authors intended"): NBestMatches reads the CV_8U descriptor rows as float (:77) and takes a Euclidean
 * norm; DetectLoop bounds its inner loop with distances.at(i) instead of distances.at(di) (:38) and indexes frames[size]
 * (:29).  The distance here is the one the reference defines for ORB descriptors, ThirdParty/DBoW2/DBoW2/FORB.cpp:81-101
 * (256-bit XOR + popcount), and the threshold is an integer number of differing bits.
 */
#include <stdint.h>
#include <string.h>

static int ham256(const uint8_t* a, const uint8_t* b)
{
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

/* dist / idx are nq x n, ascending; entries beyond the train set stay at dist = -1, idx = -1 (FLT_MAX / -1 in the reference). */
void orc_nbest(const uint8_t* q, int nq, const uint8_t* t, int nt, int n, int32_t* dist, int32_t* idx)
{
    for (int i = 0; i < nq; i++) {
        int32_t* cd = dist + (size_t)i * n;
        int32_t* ci = idx + (size_t)i * n;
        for (int k = 0; k < n; k++) { cd[k] = INT32_MAX; ci[k] = -1; }
        for (int j = 0; j < nt; j++) {
            int32_t cur_d = ham256(q + (size_t)i * 32, t + (size_t)j * 32), cur_i = j;
            for (int k = 0; k < n; k++)
                if (cur_d < cd[k]) {       /* :88-100: take the slot, carry the displaced entry down */
                    const int32_t d2 = cd[k], i2 = ci[k];
                    cd[k] = cur_d; ci[k] = cur_i;
                    cur_d = d2; cur_i = i2;
                }
        }
        for (int k = 0; k < n; k++) if (ci[k] < 0) cd[k] = -1;
    }
}

/* frames: nframes x cap x 32 bytes, frame f holding counts[f] descriptors.  scores[f] = number of n-best distances of the
 * current frame's descriptors against frame f that are below thr.  Returns the first frame with the strictly largest
 * score, or -1 when no frame scores above zero (count_max starts at 0, :26). */
int orc_loop_score(const uint8_t* q, int nq, const uint8_t* frames, const int32_t* counts, int nframes, int cap, int n, int thr,
                   int32_t* scores)
{
    int best = -1, count_max = 0;
    int32_t dist[64], idx[64];
    for (int f = 0; f < nframes; f++) {
        int count = 0;
        for (int i = 0; i < nq; i++) {
            orc_nbest(q + (size_t)i * 32, 1, frames + (size_t)f * cap * 32, counts[f], n, dist, idx);
            for (int k = 0; k < n; k++) if (idx[k] >= 0 && dist[k] < thr) count++;
        }
        scores[f] = count;
        if (count > count_max) { best = f; count_max = count; }
    }
    return best;
}
