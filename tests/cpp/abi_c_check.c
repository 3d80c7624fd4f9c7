/* include/orbx.h must be a plain C header (the boundary is a C ABI: cgo / JNI / ctypes bind it), and the call sequences
 * INTEGRATION.md shows must match its prototypes.  Compiled with `gcc -std=c99 -pedantic -Wall -Werror -c` by
 * tests/test_abi.py; never run. */
#include "orbx.h"

int integration_md_sequences(const uint8_t* const* frame_ptrs, int n, size_t stride, orbx_keypoint* kps, uint8_t* desc, int cap, int32_t* counts,
                             orbx_dmatch* good, int64_t* ngood, uint8_t* status, double* F, int32_t* ninl, orbx_dmatch* good5, int64_t* ngood5,
                             uint8_t* status5, double* F5, int32_t* ninl5)
{
    orbx_handle orb;
    hamx_handle bf;
    fmx_handle fm;
    trx_handle tx;
    orbx_params p;
    int rc;
    orbx_default_params(&p);
    p.nfeatures = 2000;
    rc = orbx_create(&orb, &p, 0, 1920, 1080, 64);
    rc |= hamx_create(&bf, 0);
    rc |= fmx_create(&fm, 0);
    rc |= trx_create(&tx, 0);
    /* section 2: blocking sequence mode */
    rc |= orbx_extract_batch(orb, frame_ptrs, n, 1920, 1080, stride, kps, desc, cap, counts);
    rc |= orbx_match_consecutive(orb, bf, 0.8f, n, cap, good, ngood);
    rc |= orbx_filter_consecutive(orb, fm, 3., 0.85, n, cap, status, F, ninl);
    rc |= orbx_match_back(orb, bf, 5, 0.8f, n, cap, good5, ngood5);
    rc |= orbx_filter_back(orb, fm, 3., 0.85, n, 5, cap, status5, F5, ninl5);
    /* pipelined */
    rc |= orbx_submit_batch(orb, bf, frame_ptrs, n, 1920, 1080, stride, 0.8f, kps, desc, cap, counts, good, ngood);
    rc |= orbx_wait_batch(orb);
    rc |= orbx_submit_batch_filtered(orb, bf, fm, frame_ptrs, n, 1920, 1080, stride, 0.8f, kps, desc, cap, counts, good, ngood, 3., 0.85, status, F, ninl);
    rc |= orbx_wait_batch(orb);
    rc |= orbx_submit_batch_back(orb, bf, fm, 5, frame_ptrs, n, 1920, 1080, stride, 0.8f, kps, desc, cap, counts, good5, ngood5, 3., 0.85, status5, F5, ninl5);
    rc |= orbx_wait_batch(orb);
    rc |= hamx_set_kernel(bf, HAMX_KERNEL_AUTO);
    rc |= hamx_reserve(bf, 2000, 200000, 1);
    rc |= trx_destroy(tx);
    rc |= fmx_destroy(fm);
    rc |= hamx_destroy(bf);
    rc |= orbx_destroy(orb);
    return rc;
}

int consumers(trx_handle tx, hamx_handle bf, const orbx_dmatch* d_good, const int64_t* d_ngood, const uint8_t* d_status, int32_t* d_premap,
              const int32_t* d_ncur, int nframes, int cap, int32_t* d_cur_map, int32_t* d_assoc_q, int32_t* d_assoc_mp, int32_t* d_nassoc,
              const int32_t* d_next_id, uint8_t* d_accept, int32_t* d_nnew, const orbx_keypoint* d_kps, const orbx_keypoint* d_hist_kps, int nhist,
              const trx_cameras* d_cams, double* d_X, uint8_t* d_front, int32_t* d_nfront, const uint8_t* cur_desc, int ncur, const uint8_t* stored,
              const int32_t* counts, int32_t* scores)
{
    int32_t best_frame;
    int rc = trx_associate_dev(tx, d_good, d_ngood, d_status, d_premap, d_ncur, nframes, 5, cap, d_cur_map, d_assoc_q, d_assoc_mp, d_nassoc);
    rc |= trx_select_new_dev(tx, d_good, d_ngood, d_status, d_premap, d_cur_map, d_ncur, d_next_id, nframes, 5, cap, d_accept, d_nnew);
    rc |= trx_triangulate_back_dev(tx, d_kps, nframes, cap, 5, d_hist_kps, nhist, d_good, d_ngood, d_accept, d_cams, d_X, d_front, d_nfront);
    rc |= hamx_loop_score(bf, cur_desc, ncur, stored, counts, nframes, cap, 10, 40, scores, &best_frame);
    return rc + best_frame;
}

/* INTEGRATION.md "The vendored DBoW2": the vocabulary's text-file columns in, whole batches transformed and scored on the device */
int bag_of_words(int k, int L, int nnodes, const int32_t* parent, const uint8_t* leaf, const uint8_t* desc, const double* weight,
                 const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, uint32_t* d_words, double* d_vals, int32_t* d_nbow,
                 uint32_t* d_fv_nodes, int32_t* d_fv_offsets, uint32_t* d_fv_feats, int32_t* d_nfv, const int64_t* d_start, double* d_scores,
                 int q, int nbow_q)
{
    bowx_handle bw;
    int32_t info[6], stopped;
    uint32_t node;
    double w;
    int rc = bowx_create(&bw, 0);
    rc |= bowx_set_vocabulary(bw, k, L, BOWX_L1_NORM, BOWX_TF_IDF, nnodes, parent, leaf, desc, weight);
    rc |= bowx_vocabulary_info(bw, info);
    rc |= bowx_transform_batch_dev(bw, d_desc, d_counts, nframes, cap, 4, d_words, d_vals, d_nbow, d_fv_nodes, d_fv_offsets, d_fv_feats, d_nfv);
    rc |= bowx_score_batch_dev(bw, d_words + (size_t)q * cap, d_vals + (size_t)q * cap, nbow_q, d_start, d_nbow, d_words, d_vals, nframes, d_scores);
    rc |= bowx_stop_words(bw, 0.5, &stopped);
    rc |= bowx_parent_node(bw, 0, 2, &node);
    rc |= bowx_word_weight(bw, 0, &w);
    rc |= bowx_synchronize(bw);
    rc |= bowx_destroy(bw);
    return rc + info[5] + stopped + (int)node + (w > 0);
}

/* INTEGRATION.md "Frame ingest from JPEG files" */
int jpeg_ingest(orbx_handle orb, const uint8_t* const* files, const size_t* sizes, int n, int w, int h, uint8_t* d_frames, orbx_keypoint* d_kps,
                uint8_t* d_desc, int cap, int32_t* d_counts)
{
    jpgx_handle jp;
    int32_t info[6];
    int rc = jpgx_create(&jp, 0);
    if (jpgx_probe(files[0], sizes[0], info) == ORBX_E_UNSUPPORTED) return 1;      /* keep imread for this file */
    rc |= jpgx_decode_gray_batch_dev(jp, files, sizes, n, w, h, d_frames, (size_t)w * h, (size_t)w);
    rc |= orbx_extract_batch_dev(orb, d_frames, (size_t)w * h, n, w, h, (size_t)w, d_kps, d_desc, cap, d_counts);
    rc |= jpgx_synchronize(jp);
    rc |= jpgx_destroy(jp);
    return rc + info[0];
}
