/*
 * orbx.h -- C ABI of liborbx.so: a B200 (sm_100a) ORB front-end for the Monocular_SLAM pipeline.
 *
 * It replaces, and only replaces, the reference's per-frame feature extraction and descriptor matching:
 *
 *   reference call site                                                    replaced by
 *   ---------------------------------------------------------------------  ---------------------------------
 *   detector.detect(frameBuffer, keypoints)    src/FeatureExtractor.cpp:17  orbx_detect
 *   extractor.compute(frameBuffer, kps, desc)  src/FeatureExtractor.cpp:19  orbx_compute
 *   (the two back to back, FeatureExtractor::process :13-31)                orbx_detect_and_compute / orbx_extract_batch
 *   BFMatcher(NORM_HAMMING,false).knnMatch(d1,d2,raw,2)
 *                                      src/CameraPoseEstimator.cpp:202-204  hamx_knn2
 *   matchFeatures(d1,d2,matches,ratio) src/CameraPoseEstimator.cpp:200-213  hamx_match_ratio
 *   the distance itself   ThirdParty/DBoW2/DBoW2/FORB.cpp:81-101            (256-bit XOR + popcount inside hamx_*)
 *
 * The ORB objects in the reference are default-constructed cv::ORB (src/FeatureExtractor.h:23-24); their
 * constructor arguments are the fields of orbx_params with the same defaults.  Results follow OpenCV's ORB /
 * BFMatcher bit for bit (see DESIGN.md "Parity"); keypoints are returned in canonical order (octave, y, x) and
 * descriptor row i always belongs to keypoint i, which is the index contract the reference relies on
 * (src/CameraPoseEstimator.cpp:434-436,558-559).
 *
 * Conventions: plain C, no exceptions or STL across the boundary; every function returns 0 on success or a
 * negative orbx_status; orbx_last_error() gives the message for the calling thread.  Pointers are HOST pointers
 * unless the function name ends in _dev.  A handle is bound to one device and one stream and is not thread-safe
 * (the reference runs its nodes on a single thread, src/main.cpp:49-51).  There is no CPU fallback: without a
 * CUDA device every call fails with ORBX_E_CUDA.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ORBX_API __attribute__((visibility("default")))
#else
#define ORBX_API
#endif

typedef enum {
    ORBX_OK = 0,
    ORBX_E_INVALID = -1,   /* bad argument (null pointer, size out of range, unsupported parameter) */
    ORBX_E_CUDA = -2,      /* CUDA runtime error, or no usable device */
    ORBX_E_CAPACITY = -3,  /* more keypoints than the caller's buffer (or the handle's internal lists) can hold */
    ORBX_E_ALLOC = -4,     /* out of host or device memory */
    ORBX_E_ALIGN = -5,     /* a device pointer is not 16-byte aligned */
    ORBX_E_UNSUPPORTED = -6 /* a valid input of a kind this library does not handle (jpgx_*: a JPEG that is not one-component baseline) */
} orbx_status;

/* == cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle (degrees), response, octave, class_id (-1) */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orbx_keypoint;
/* == cv::DMatch (16 bytes) */
typedef struct { int32_t query_idx, train_idx, img_idx; float distance; } orbx_dmatch;
/* per-query best two neighbours; absent entries have idx = -1, dist = -1 */
typedef struct { int32_t dist0, idx0, dist1, idx1; } hamx_top2;
/* one (query set, train set) problem of a batched launch; q and t are DEVICE pointers, 16-byte aligned */
typedef struct { const uint8_t* q; const uint8_t* t; int32_t nq, nt; } hamx_pair;

#define ORBX_HARRIS_SCORE 0
#define ORBX_FAST_SCORE 1

/* cv::ORB constructor arguments.  Only nfeatures, nlevels, scale_factor, score_type and fast_threshold may differ
 * from the defaults; edge_threshold 31, first_level 0, wta_k 2, patch_size 31 are required. */
typedef struct {
    int32_t nfeatures;       /* 500  */
    float scale_factor;      /* 1.2f */
    int32_t nlevels;         /* 8    */
    int32_t edge_threshold;  /* 31   */
    int32_t first_level;     /* 0    */
    int32_t wta_k;           /* 2    */
    int32_t score_type;      /* ORBX_HARRIS_SCORE */
    int32_t patch_size;      /* 31   */
    int32_t fast_threshold;  /* 20   */
} orbx_params;

typedef struct orbx_context* orbx_handle;
typedef struct hamx_context* hamx_handle;

ORBX_API const char* orbx_last_error(void);
ORBX_API const char* orbx_version(void);
ORBX_API int orbx_device_count(void);
/* Page-locked host memory for frames and result buffers (callers that do not link the CUDA runtime themselves): copies
 * from / to such memory are asynchronous, which is what lets orbx_submit_batch overlap them with the kernels. */
ORBX_API int orbx_host_alloc(size_t bytes, void** out);
/* The same, write-combined: for buffers the host only WRITES (frames on their way to the device); reading them back on the
 * host is very slow.  Freed with orbx_host_free. */
ORBX_API int orbx_host_alloc_wc(size_t bytes, void** out);
ORBX_API int orbx_host_free(void* p);
ORBX_API void orbx_default_params(orbx_params* p);

/* ------------------------------------------------------------------ extraction (FeatureExtractor) */

/* Persistent state for frames up to max_w x max_h, up to max_batch frames in flight per call.
 * Mirrors ProcessingNode::init()/destroy() (src/ProcessingNode.h:19-21). */
ORBX_API int orbx_create(orbx_handle* out, const orbx_params* params, int device, int max_w, int max_h, int max_batch);
ORBX_API int orbx_destroy(orbx_handle h);
/* Run on a caller-provided cudaStream_t (e.g. the framework's current stream); NULL restores the handle's own. */
ORBX_API int orbx_set_stream(orbx_handle h, void* cuda_stream);
ORBX_API int orbx_synchronize(orbx_handle h);
/* Input pixel format of every image pointer passed afterwards: 1 = CV_8UC1 gray (default), 3 = CV_8UC3 BGR interleaved
 * (stride in bytes, >= 3 * w).  cv::ORB converts non-gray input with cvtColor(BGR2GRAY) before anything else -- the
 * reference loads frames with CV_LOAD_IMAGE_UNCHANGED (src/FrameLoader.cpp:62) and hands them to detect()/compute()
 * as they are (src/FeatureExtractor.cpp:17,19) -- and so does this library, on the device, bit-exactly. */
ORBX_API int orbx_set_input_channels(orbx_handle h, int channels);
/* Largest number of keypoints one frame can return with the handle's parameters (ties included). */
ORBX_API int orbx_max_keypoints(orbx_handle h);
/* Level geometry and quotas as the handle computes them (arrays of nlevels entries; any pointer may be NULL). */
ORBX_API int orbx_level_info(orbx_handle h, int w, int h_, int32_t* widths, int32_t* heights, float* scales, int32_t* quotas);

/* OrbFeatureDetector::detect(image, keypoints): gray is h x w, 8-bit, row stride in bytes. */
ORBX_API int orbx_detect(orbx_handle h, const uint8_t* gray, int w, int h_, size_t stride, orbx_keypoint* out, int cap, int* n);
/* OrbDescriptorExtractor::compute(image, keypoints, descriptors): keypoints are filtered (border) and stably
 * regrouped by octave in place, *n is updated, desc receives *n rows of 32 bytes.  Angles are taken from the
 * keypoints, as OpenCV does. */
ORBX_API int orbx_compute(orbx_handle h, const uint8_t* gray, int w, int h_, size_t stride, orbx_keypoint* kps, int* n, uint8_t* desc);
/* detect followed by compute with one pyramid; results identical to the two separate calls. */
ORBX_API int orbx_detect_and_compute(orbx_handle h, const uint8_t* gray, int w, int h_, size_t stride,
                            orbx_keypoint* out, uint8_t* desc, int cap, int* n);
/* nframes (<= max_batch) frames of one size in one submission.  Frame f writes out[f*cap ..], desc[f*cap*32 ..],
 * counts[f].  This is the throughput entry point used for sequences (frames are independent, src/main.cpp:36-51). */
ORBX_API int orbx_extract_batch(orbx_handle h, const uint8_t* const* frames, int nframes, int w, int h_, size_t stride,
                       orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts);
/* Same with everything resident on the handle's device: frames at d_frames + f*frame_pitch_bytes. Asynchronous on
 * the handle's stream; d_counts[f] is the keypoint count of frame f.  Overflow is reported by orbx_check_dev. */
ORBX_API int orbx_extract_batch_dev(orbx_handle h, const uint8_t* d_frames, size_t frame_pitch_bytes, int nframes, int w, int h_,
                           size_t stride, orbx_keypoint* d_out, uint8_t* d_desc, int cap, int32_t* d_counts);
/* Synchronises and returns ORBX_E_CAPACITY if any _dev call since the last check overflowed a list. */
ORBX_API int orbx_check_dev(orbx_handle h);

/* Sequence mode on host buffers: match every frame of the batch last extracted with orbx_extract_batch against its
 * predecessor (frame 0 against the last frame of the previous batch; orbx_reset_sequence forgets it, and so does a batch
 * extracted with another cap) without moving the descriptors back to the device.  The caller states the geometry of its
 * buffers: good is [nframes][cap], ngood [nframes]; nframes must be the size of that batch and cap at least
 * min(its cap, orbx_max_keypoints()), otherwise ORBX_E_INVALID (nothing is written). */
ORBX_API int orbx_match_consecutive(orbx_handle h, hamx_handle m, float ratio, int nframes, int cap, orbx_dmatch* good, int64_t* ngood);
ORBX_API int orbx_reset_sequence(orbx_handle h);

/* Pipelined sequence mode on host buffers: up to orbx_pipeline_depth() (3) batches in flight.  orbx_submit_batch enqueues, without
 * waiting, (1) the upload of the frames, (2) FeatureExtractor::process of every frame (src/FeatureExtractor.cpp:13-31)
 * and, when m != NULL, matchFeatures(frame f, frame f-1, ratio) (src/CameraPoseEstimator.cpp:200-213,409; frame 0 against
 * the last frame of the previous submission), (3) the copy of keypoints [nframes][cap], descriptors [nframes][cap][32]
 * and accepted matches [nframes][cap] back into the caller's buffers (pinned memory keeps the copies asynchronous).
 * The uploads of the next batches and the download of batch k-1 overlap the kernels of batch k.  orbx_wait_batch blocks until
 * the OLDEST submitted batch is complete, then fills its counts[] / ngood[] and reports overflows; the buffers passed
 * to orbx_submit_batch must stay valid until then.  cap <= orbx_max_keypoints().  Other entry points of the handle
 * must not be called while a batch is in flight (they return ORBX_E_INVALID). */
ORBX_API int orbx_submit_batch(orbx_handle h, hamx_handle m, const uint8_t* const* frames, int nframes, int w, int h_, size_t stride,
                      float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts, orbx_dmatch* good, int64_t* ngood);
ORBX_API int orbx_wait_batch(orbx_handle h);
ORBX_API int orbx_batches_in_flight(orbx_handle h);
ORBX_API int orbx_pipeline_depth(orbx_handle h);

/* Per-stage device times (CUDA events on the handle's stream around each stage of every batch while enabled).
 * stage_ms receives ORBX_NSTAGES averages per batch: pyramid, FAST, FAST-score cut, Harris+selection, orient+describe. */
#define ORBX_NSTAGES 5
ORBX_API int orbx_set_profiling(orbx_handle h, int enabled);
ORBX_API int orbx_read_profile(orbx_handle h, float* stage_ms, int* nbatches);

/* Stage-level taps for parity tests and profiling (device work + copy back). */
ORBX_API int orbx_debug_pyramid_level(orbx_handle h, const uint8_t* gray, int w, int h_, size_t stride, int level, uint8_t* out);
ORBX_API int orbx_debug_fast_level(orbx_handle h, const uint8_t* gray, int w, int h_, size_t stride, int level,
                          int32_t* xs, int32_t* ys, int32_t* scores, int cap, int* n);

/* ------------------------------------------------------------------ matching (CameraPoseEstimator::matchFeatures) */

ORBX_API int hamx_create(hamx_handle* out, int device);
ORBX_API int hamx_destroy(hamx_handle h);
/* The matcher's kernels run on this cudaStream_t (NULL: the handle's own).  A handle's workspaces are shared by all its
 * calls, so one handle must not be used on two streams concurrently.  The orbx_* sequence entry points that take a matcher
 * run it on the extractor's stream for the duration of the call and put the stream installed here back afterwards. */
ORBX_API int hamx_set_stream(hamx_handle h, void* cuda_stream);
ORBX_API int hamx_get_stream(hamx_handle h, void** cuda_stream);
ORBX_API int hamx_synchronize(hamx_handle h);
/* Workspaces grow on demand, and growing frees and allocates device memory, which synchronises the whole device.
 * hamx_reserve sizes them once for every later call with at most nq queries, nt train rows and npairs batched pairs, so
 * that the _dev entry points stay asynchronous. */
ORBX_API int hamx_reserve(hamx_handle h, int64_t nq, int64_t nt, int npairs);

/* BFMatcher(NORM_HAMMING,false).knnMatch(q, t, out, 2): q is nq x 32 bytes, t is nt x 32 bytes.
 * out holds nq*2 entries; out_counts[i] = min(nt, 2) entries of row i are valid, sorted by (distance, trainIdx). */
ORBX_API int hamx_knn2(hamx_handle h, const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, orbx_dmatch* out, int32_t* out_counts);
/* matchFeatures(): knnMatch k=2 + Lowe ratio `d0 < d1 * ratio` in float, accepted matches in ascending query order.
 * Queries with fewer than two neighbours are dropped (the reference would index out of range there). */
ORBX_API int hamx_match_ratio(hamx_handle h, const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, float ratio,
                     orbx_dmatch* good /* cap nq */, int64_t* ngood);

/* Device-resident pieces (asynchronous on the handle's stream; pointers must be 16-byte aligned). */
ORBX_API int hamx_knn2_dev(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int64_t nt, int64_t train_offset,
                  hamx_top2* d_out);
/* Two kernels compute the same bit-identical result.  HAMX_KERNEL_TENSOR: the distance matrix as an int8 contraction of
 * +-1-expanded descriptors on the tensor cores (tcgen05.mma, accumulators in tensor memory), the integer pipes only fold it
 * into the top-2 -- 5.5x the throughput on large problems.  HAMX_KERNEL_INTEGER: XOR + POPC on the integer pipes, no set-up
 * cost -- faster on small ones and the only kernel behind the batched frame-pair entry points.  HAMX_KERNEL_AUTO (default)
 * picks per call by problem size. */
#define HAMX_KERNEL_AUTO 0
#define HAMX_KERNEL_INTEGER 1
#define HAMX_KERNEL_TENSOR 2
ORBX_API int hamx_set_kernel(hamx_handle h, int mode);
/* Merge nparts per-shard results laid out [part][nq] (e.g. the output of an all-gather) into d_out[nq]. */
ORBX_API int hamx_merge_top2_dev(hamx_handle h, const hamx_top2* d_parts, int nparts, int64_t nq, hamx_top2* d_out);
/* Ratio test + ordered compaction: d_good gets the accepted matches, *d_ngood their number. */
ORBX_API int hamx_ratio_dev(hamx_handle h, const hamx_top2* d_top2, int64_t nq, float ratio, orbx_dmatch* d_good, int64_t* d_ngood);

/* npairs independent matchFeatures() problems in one launch (d_pairs is a device-resident table; max_nq / max_nt bound
 * the pairs' sizes).  Pair p writes its accepted matches to d_good + p*good_stride and their number to d_ngood[p]. */
ORBX_API int hamx_match_pairs_dev(hamx_handle h, const hamx_pair* d_pairs, int npairs, int max_nq, int max_nt, float ratio,
                         orbx_dmatch* d_good, size_t good_stride, int64_t* d_ngood);

/* Consecutive-frame matching of a batch that orbx_extract_batch_dev left on the device (descriptors [nframes][cap][32],
 * counts [nframes]): frame f is the query set, frame f-1 the train set -- matchFeatures(desc_cur, desc_prev) of
 * src/CameraPoseEstimator.cpp:409; frame 0 is matched against d_prev_desc / *d_prev_count (the last frame of the
 * previous batch) or, when those are NULL, yields no matches.  d_good is [nframes][cap], d_ngood [nframes]. */
ORBX_API int hamx_match_consecutive_dev(hamx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap,
                               const uint8_t* d_prev_desc, const int32_t* d_prev_count, float ratio,
                               orbx_dmatch* d_good, int64_t* d_ngood);

/* Steady-state pattern of CameraPoseEstimator::pnpPoseEstimation (src/CameraPoseEstimator.cpp:388,405-409): every frame
 * is matched, as the query set, against each of its `back` (numBackTraverse = 5) predecessors.  Pair (f, j), j = 1..back,
 * is matchFeatures(desc[f], desc[f-j], ratio); its accepted matches go to d_good[(f*back + j-1) * cap ..] and their number
 * to d_ngood[f*back + j-1].  Predecessors before the batch come from the history [nhist][cap][32] (entry 0 = the frame
 * just before the batch); pairs without a predecessor yield 0 matches.  hamx_update_history_dev writes the history that
 * follows the batch (the `back` most recent frames, most recent first) into a second buffer. */
ORBX_API int hamx_match_back_dev(hamx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int back,
                        const uint8_t* d_hist_desc, const int32_t* d_hist_counts, int nhist, float ratio,
                        orbx_dmatch* d_good, int64_t* d_ngood);
ORBX_API int hamx_update_history_dev(hamx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int back,
                            const uint8_t* d_old_desc, const int32_t* d_old_counts, int nold, uint8_t* d_new_desc, int32_t* d_new_counts);
/* Host-buffer form for the batch last extracted with orbx_extract_batch: good is [nframes][back][cap], ngood
 * [nframes][back] (nframes / cap: the caller's buffer geometry, checked like orbx_match_consecutive's); the handle keeps
 * the history across batches (orbx_reset_sequence forgets it).  back <= 8. */
ORBX_API int orbx_match_back(orbx_handle h, hamx_handle m, int back, float ratio, int nframes, int cap, orbx_dmatch* good, int64_t* ngood);

/* Train-sharded matching over peer memory (BASELINE configs 4 and 5; one process per GPU of one NVLink box).  The
 * reference has no counterpart (it is single-device); the result is bit-identical to hamx_knn2_dev over the union of
 * the shards.  Setup, once: every rank calls hamx_p2p_export (allocates its gather buffer and returns a 64-byte
 * cudaIpc handle), the ranks exchange the handles (e.g. torch.distributed.all_gather), every rank calls hamx_p2p_import
 * with the world's handles in rank order, then a barrier.  hamx_p2p_import_ptrs is the same-process variant (raw
 * device pointers of the other ranks' local_base, e.g. several handles in one test process).
 * Per query batch (collective: same nq on every rank, same number of calls): hamx_knn2_p2p_dev computes the local
 * top-2 over this rank's train rows [train_offset, train_offset + nt), the matching kernel itself stores them into
 * every rank's gather buffer over NVLink and publishes a flag, and a merge kernel waits for all ranks' flags and
 * reduces with the (distance, trainIdx) rule.  No NCCL call and no host synchronisation is involved.  The two halves are
 * also exported separately (scatter, merge). */
ORBX_API int hamx_p2p_export(hamx_handle h, int64_t nq_max, int world, int rank, uint8_t* ipc_handle /* 64 bytes, may be NULL */,
                    void** local_base /* may be NULL */);
ORBX_API int hamx_p2p_import(hamx_handle h, const uint8_t* ipc_handles /* world * 64 bytes, rank order */);
ORBX_API int hamx_p2p_import_ptrs(hamx_handle h, void* const* peer_bases /* world entries; own entry ignored */);
ORBX_API int hamx_p2p_close(hamx_handle h);
ORBX_API int hamx_knn2_p2p_dev(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int64_t nt, int64_t train_offset,
                      hamx_top2* d_out);
ORBX_API int hamx_knn2_p2p_scatter_dev(hamx_handle h, const uint8_t* d_q, int64_t nq, const uint8_t* d_t, int64_t nt, int64_t train_offset);
ORBX_API int hamx_p2p_merge_dev(hamx_handle h, int64_t nq, hamx_top2* d_out);

/* Loop-closure candidate scoring (reference src/LoopCloser.cpp:19-105, "as the authors intended": the
 * distance is the 256-bit Hamming distance the reference defines for ORB descriptors, ThirdParty/DBoW2/DBoW2/FORB.cpp:81-101,
 * instead of a float norm over CV_8U rows, and thr is a number of bits).
 *   hamx_nbest[_dev]      = LoopCloser::NBestMatches(descriptors1, descriptors2, n, distances, indices) (:53-105): for every
 *                           query row the n (<= 16) best train rows, ascending, filled by the reference's own insertion rule
 *                           (ties: see csrc/hamming.cu k_nbest); dist / idx are [nq][n], absent entries -1.
 *   hamx_loop_score[_dev] = the loop of LoopCloser::DetectLoop (:30-47): the current frame's descriptors against nframes
 *                           stored frames ([nframes][cap][32] bytes, counts[f] rows valid) in ONE launch; scores[f] = number
 *                           of n-best distances below thr, best = the first frame with the strictly largest non-zero score
 *                           (-1: none).  d_best (may be NULL) receives {frame, score}.
 *   hamx_loop_best_dev    = that arg-max alone, e.g. over scores gathered from several GPUs (stored frames sharded by frame). */
ORBX_API int hamx_nbest(hamx_handle h, const uint8_t* q, int nq, const uint8_t* t, int nt, int n, int32_t* dist, int32_t* idx);
ORBX_API int hamx_nbest_dev(hamx_handle h, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int n, int32_t* d_dist, int32_t* d_idx);
ORBX_API int hamx_loop_score(hamx_handle h, const uint8_t* q, int nq, const uint8_t* frames, const int32_t* counts, int nframes, int cap, int n,
                    int thr, int32_t* scores, int32_t* best_frame);
ORBX_API int hamx_loop_score_dev(hamx_handle h, const uint8_t* d_q, int nq, const uint8_t* d_frames, const int32_t* d_counts, int nframes, int cap,
                        int n, int thr, int32_t* d_scores, int32_t* d_best);
ORBX_API int hamx_loop_best_dev(hamx_handle h, const int32_t* d_scores, int nframes, int32_t* d_best);

/* Register-only popcount microbenchmark: the measured integer-pipe peak used as the matcher's roofline denominator.
 * gpopc_per_s = 32-bit POPC results per second / 1e9, over the whole device. */
ORBX_API int hamx_popc_peak(int device, double* gpopc_per_s, double* elapsed_ms);

/* ------------------------------------------------------------------ outlier filter (computeFundamentalMatrix)
 * The step after matchFeatures for every matched frame pair (src/CameraPoseEstimator.cpp:545-586, called at :291, :419):
 *     F = findFundamentalMat(inputs1, inputs2, CV_FM_RANSAC, MAX_DISTANCE, CONFIDENCE, status);   (:563)
 *     F = findFundamentalMat(inliers1, inliers2, CV_FM_8POINT);                                     (:585)
 * One CUDA block per pair reproduces OpenCV's sequential RANSAC (same random sample sequence, same candidate order,
 * same best-update and iteration-budget rule), so `status` is the mask OpenCV returns; F is the 8-point matrix of the
 * inliers, scaled to F[8] = 1 (row-major double[9]; all zeros when there is no result).  With 8..14
 * correspondences OpenCV runs least-median-of-squares instead of RANSAC (same sampler, fixed iteration count); so does
 * the kernel -- identical to OpenCV with 14, a valid LMedS answer with 8..13 (there OpenCV's own winner is decided by
 * 1e-25-level rounding noise).  Fewer than 8 correspondences report no model.  max_distance <= 0 means 3, confidence outside (0, 1) means 0.99, as in OpenCV; the reference passes
 * MAX_DISTANCE = 3 and CONFIDENCE = 0.85 (src/ParamConfig.h:24-25). */
typedef struct fmx_context* fmx_handle;
ORBX_API int fmx_create(fmx_handle* out, int device);
ORBX_API int fmx_destroy(fmx_handle h);
ORBX_API int fmx_set_stream(fmx_handle h, void* cuda_stream);     /* same rules as hamx_set_stream */
ORBX_API int fmx_get_stream(fmx_handle h, void** cuda_stream);
ORBX_API int fmx_synchronize(fmx_handle h);
/* The reference's own signature for one pair: keypoints of both frames and the DMatch list (queryIdx -> kps1, trainIdx -> kps2).
 * status gets nm bytes (0/1), F 9 doubles, *ninliers the number of set status bytes. */
ORBX_API int fmx_compute_fundamental(fmx_handle h, const orbx_keypoint* kps1, int n1, const orbx_keypoint* kps2, int n2,
                            const orbx_dmatch* matches, int nm, double max_distance, double confidence,
                            uint8_t* status, double* F, int32_t* ninliers);
/* npairs independent pairs in one launch.  pts1 / pts2: [npairs][cap][2] float (x, y) correspondences, counts[p] of them
 * valid; status [npairs][cap], F [npairs][9], ninliers [npairs]. */
ORBX_API int fmx_fundamental_batch(fmx_handle h, const float* pts1, const float* pts2, const int32_t* counts, int npairs, int cap,
                          double max_distance, double confidence, uint8_t* status, double* F, int32_t* ninliers);
/* Per pair {inliers, RANSAC iterations run, candidate matrices scored, 1 if F was produced} of the last fmx_fundamental_batch. */
ORBX_API int fmx_last_info(fmx_handle h, int npairs, int32_t* info /* npairs * 4 */);
/* Device-resident form (asynchronous on the handle's stream); d_info is [npairs][4] as fmx_last_info returns it.  Batches of more
 * pairs than the device has CTA slots for (2 per SM) keep 704 bytes per pair in a grow-only workspace of the handle: the first such
 * call, and any later one with more pairs, allocates (a device-wide synchronisation), the others do not. */
ORBX_API int fmx_fundamental_batch_dev(fmx_handle h, const float* d_pts1, const float* d_pts2, const int32_t* d_counts, int npairs, int cap,
                              double max_distance, double confidence, uint8_t* d_status, double* d_F, int32_t* d_info);
/* Sequence mode on the device: pair f = (frame f, frame f-1) of a batch whose keypoints [nframes][cap] and consecutive-frame
 * match lists [nframes][cap] / d_ngood[nframes] are device-resident (hamx_match_consecutive_dev); frame 0 pairs with
 * d_prev_kps (the last frame of the previous batch) or, when that is NULL, reports no model.  status byte i of pair f
 * belongs to d_good[f*cap + i]. */
ORBX_API int fmx_filter_consecutive_dev(fmx_handle h, const orbx_keypoint* d_kps, const orbx_keypoint* d_prev_kps, int nframes, int cap,
                               const orbx_dmatch* d_good, const int64_t* d_ngood, double max_distance, double confidence,
                               uint8_t* d_status, double* d_F, int32_t* d_info);
/* The same for the steady-state pattern of hamx_match_back_dev: pair (f, j) = (frame f, frame f-j), j = 1..back, at index
 * f*back + j-1 of d_good [nframes*back][cap] / d_ngood / d_status / d_F / d_info; predecessors before the batch come from the
 * keypoint history d_hist_kps [nhist][cap] (entry 0 = the frame just before the batch). */
ORBX_API int fmx_filter_back_dev(fmx_handle h, const orbx_keypoint* d_kps, int nframes, int cap, int back, const orbx_keypoint* d_hist_kps,
                        int nhist, const orbx_dmatch* d_good, const int64_t* d_ngood, double max_distance, double confidence,
                        uint8_t* d_status, double* d_F, int32_t* d_info);
/* Host-buffer form for the batch last passed to orbx_extract_batch + orbx_match_consecutive on `h`: status [nframes][cap]
 * (the caller's buffer geometry, checked like orbx_match_consecutive's), F [nframes][9], ninliers [nframes].  The matches
 * never leave the device in between. */
ORBX_API int orbx_filter_consecutive(orbx_handle h, fmx_handle fm, double max_distance, double confidence, int nframes, int cap,
                            uint8_t* status, double* F, int32_t* ninliers);
/* The filter for the pairs orbx_match_back just matched: status [nframes][back][cap], F [nframes][back][9], ninliers [nframes][back]
 * (computeFundamentalMatrix as called in the loop at src/CameraPoseEstimator.cpp:405-419). */
ORBX_API int orbx_filter_back(orbx_handle h, fmx_handle fm, double max_distance, double confidence, int nframes, int back, int cap,
                     uint8_t* status, double* F, int32_t* ninliers);
/* orbx_submit_batch with the outlier filter appended to the batch's device work (fm may be NULL: plain orbx_submit_batch):
 * status [nframes][cap] and F [nframes][9] are written before the matching orbx_wait_batch returns, ninliers [nframes] by it. */
ORBX_API int orbx_submit_batch_filtered(orbx_handle h, hamx_handle m, fmx_handle fm, const uint8_t* const* frames, int nframes, int w, int h_,
                               size_t stride, float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts,
                               orbx_dmatch* good, int64_t* ngood, double max_distance, double confidence, uint8_t* status,
                               double* F, int32_t* ninliers);
/* The pipelined form of the steady-state loop (src/CameraPoseEstimator.cpp:405-419): every frame of the batch against its `back`
 * (<= 8) predecessors, matchFeatures and -- when fm is not NULL -- computeFundamentalMatrix per pair.  good / status are
 * [nframes][back][cap], ngood / ninliers [nframes][back], F [nframes][back][9]; pair (f, j) at index f*back + j-1 as in
 * hamx_match_back_dev.  The history of the frames before the batch lives on the device (orbx_reset_sequence forgets it). */
ORBX_API int orbx_submit_batch_back(orbx_handle h, hamx_handle m, fmx_handle fm, int back, const uint8_t* const* frames, int nframes, int w,
                           int h_, size_t stride, float ratio, orbx_keypoint* out, uint8_t* desc, int cap, int32_t* counts,
                           orbx_dmatch* good, int64_t* ngood, double max_distance, double confidence, uint8_t* status, double* F,
                           int32_t* ninliers);

/* ------------------------------------------------------------------ triangulation and map-point association
 * The consumers of the filtered match lists, batched on the device so that the lists need not leave it:
 *   TriangulateSinglePointFromTwoView / TriangulateMultiplePointsFromTwoView     src/CameraPoseEstimator.cpp:86-152
 *   the bootstrap's test of the four [R|t] candidates by points in front         src/CameraPoseEstimator.cpp:334-349
 *   the association loop of pnpPoseEstimation                                    src/CameraPoseEstimator.cpp:402-455
 *   the new-map-point loop                                                       src/CameraPoseEstimator.cpp:488-512
 * X is the dehomogenised null direction of the reference's 4x4 system (cv::SVD in the reference, a one-sided Jacobi SVD in
 * double here: X agrees with OpenCV to ~1e-14 relative on well-posed pairs; the parity tests ask for 1e-9); the
 * front-of-both-cameras flags and their counts are identical.  All matrices are row-major doubles. */
typedef struct trx_context* trx_handle;
/* the Rt1, Rt2 (3x4), K1, K2 (3x3) arguments of :86-88 for one problem */
typedef struct { double Rt1[12], Rt2[12], K1[9], K2[9]; } trx_cameras;

ORBX_API int trx_create(trx_handle* out, int device);
ORBX_API int trx_destroy(trx_handle h);
ORBX_API int trx_set_stream(trx_handle h, void* cuda_stream);     /* same rules as hamx_set_stream */
ORBX_API int trx_get_stream(trx_handle h, void** cuda_stream);
ORBX_API int trx_synchronize(trx_handle h);
/* TriangulateMultiplePointsFromTwoView(pts1, pts2, Rt1, Rt2, K1, K2, result, countFront = true): pts are n x 2 doubles
 * (vector<Point2d>), X receives n x 3 doubles (vector<Point3d>), front n bytes (may be NULL), *nfront the return value. */
ORBX_API int trx_triangulate(trx_handle h, const double* pts1, const double* pts2, int n, const double* Rt1, const double* Rt2, const double* K1,
                    const double* K2, double* X, uint8_t* front, int32_t* nfront);
/* The loop at :334-349: the same correspondences under nhyp candidate [R|t] for the second view (Rts: nhyp x 12).  counts[i] =
 * points in front under candidate i, *best = the first maximum (`if (maxCount < count)`), X (may be NULL) = the points of the
 * winner. */
ORBX_API int trx_triangulate_hypotheses(trx_handle h, const double* pts1, const double* pts2, int n, const double* Rt1, const double* Rts, int nhyp,
                               const double* K1, const double* K2, double* X, int32_t* counts, int32_t* best);
/* nprob problems in one launch, nhyp camera hypotheses each (cams is [nprob][nhyp]).  pts1 / pts2: [nprob][cap][2] float
 * (keypoint positions are float; the reference widens them to double, src/FeatureExtractor.cpp:20-22), counts[p] of them
 * valid; select (may be NULL) [nprob][cap]: only entries with a non-zero byte are triangulated (e.g. the RANSAC status).
 * X [nprob][nhyp][cap][3], front [nprob][nhyp][cap], nfront [nprob][nhyp]; best (may be NULL) [nprob].  Entries that are
 * not triangulated get X = 0, front = 0. */
ORBX_API int trx_triangulate_batch(trx_handle h, const float* pts1, const float* pts2, const int32_t* counts, const uint8_t* select, int nprob,
                          int cap, const trx_cameras* cams, int nhyp, double* X, uint8_t* front, int32_t* nfront, int32_t* best);
ORBX_API int trx_triangulate_batch_dev(trx_handle h, const float* d_pts1, const float* d_pts2, const int32_t* d_counts, const uint8_t* d_select,
                              int nprob, int cap, const trx_cameras* d_cams, int nhyp, double* d_X, uint8_t* d_front, int32_t* d_nfront,
                              int32_t* d_best);
/* Sequence mode on the device, fed by hamx_match_back_dev / fmx_filter_back_dev: pair (f, j) at index f*back + j-1 joins,
 * for every match of its list with a non-zero d_select byte (NULL: all), keypoint train_idx of frame f-j as view 1 and
 * keypoint query_idx of frame f as view 2 (the argument order of :504-506); d_cams[f*back + j-1] holds the two cameras.
 * X [nframes*back][cap][3], front [nframes*back][cap], nfront [nframes*back]. */
ORBX_API int trx_triangulate_back_dev(trx_handle h, const orbx_keypoint* d_kps, int nframes, int cap, int back, const orbx_keypoint* d_hist_kps,
                             int nhist, const orbx_dmatch* d_good, const int64_t* d_ngood, const uint8_t* d_select, const trx_cameras* d_cams,
                             double* d_X, uint8_t* d_front, int32_t* d_nfront);
/* The association loop (:402-455) for nprob independent current frames.  Problem p has `back` match lists: list l (the
 * matches against predecessor l, most recent first) at d_good + (p*back + l)*cap with d_ngood[p*back + l] entries, of which
 * only those with a non-zero d_status byte take part (NULL: all; the FILTERING_WITH_F block :412-424 drops the others), and
 * d_premap[(p*back + l)*cap + t] = map-point index of the predecessor's feature t, -1 for none.  d_ncur[p] = features of the
 * current frame.  Outputs: d_cur_map [nprob][cap] = the map point each current feature inherits (-1: none),
 * d_assoc_q / d_assoc_mp [nprob][cap] = the associations (current feature, map point) in the order the reference makes
 * them -- the order of mapPoints / imagePoints handed to solvePnPRansac (:447-449, :472) -- and d_nassoc[p] their number. */
ORBX_API int trx_associate_dev(trx_handle h, const orbx_dmatch* d_good, const int64_t* d_ngood, const uint8_t* d_status, const int32_t* d_premap,
                      const int32_t* d_ncur, int nprob, int back, int cap, int32_t* d_cur_map, int32_t* d_assoc_q, int32_t* d_assoc_mp,
                      int32_t* d_nassoc);
/* The new-map-point loop (:488-512) over the same lists: d_accept [nprob][back][cap] marks the matches the reference
 * triangulates (both features without a map point when the sequential walk reaches them); d_premap and d_cur_map are
 * updated in place the way registerNewMapPoint does (:235-243), new indices counted from d_next_id[p] (NULL: 0) in the order
 * of the walk; d_nnew[p] = number of new points.  Feed d_accept to trx_triangulate_back_dev as d_select. */
ORBX_API int trx_select_new_dev(trx_handle h, const orbx_dmatch* d_good, const int64_t* d_ngood, const uint8_t* d_status, int32_t* d_premap,
                       int32_t* d_cur_map, const int32_t* d_ncur, const int32_t* d_next_id, int nprob, int back, int cap,
                       uint8_t* d_accept, int32_t* d_nnew);

/* ------------------------------------------------------------------ bag of words (vendored DBoW2: loop closing)
 * The reference vendors DBoW2 with the ORB descriptor class (ThirdParty/DBoW2/DBoW2: TemplatedVocabulary.h, FORB.cpp,
 * BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp) for the loop closer its authors sketched (src/LoopCloser.cpp).  This
 * family replaces TemplatedVocabulary<FORB::TDescriptor, FORB>: transform() of whole batches of frames straight from the
 * extractor's device-resident descriptors, and score() of one bag-of-words vector against a database of them.
 * Results are the reference's doubles bit for bit (sums run in std::map order); only KL scores, which go through log(),
 * may differ in the last place.  One defined deviation: when a descent ends at a childless node above level L - levelsup
 * the reference leaves the feature-vector node id unwritten (an uninitialised local); here it is the node the descent
 * ended at. */
typedef struct bowx_context* bowx_handle;
enum { BOWX_TF_IDF = 0, BOWX_TF = 1, BOWX_IDF = 2, BOWX_BINARY = 3 };                    /* WeightingType, BowVector.h:36-42 */
enum { BOWX_L1_NORM = 0, BOWX_L2_NORM = 1, BOWX_CHI_SQUARE = 2, BOWX_KL = 3, BOWX_BHATTACHARYYA = 4, BOWX_DOT_PRODUCT = 5 };  /* ScoringType, :45-53 */

ORBX_API int bowx_create(bowx_handle* out, int device);
ORBX_API int bowx_destroy(bowx_handle h);
ORBX_API int bowx_set_stream(bowx_handle h, void* cuda_stream);    /* same rules as hamx_set_stream */
ORBX_API int bowx_get_stream(bowx_handle h, void** cuda_stream);
ORBX_API int bowx_synchronize(bowx_handle h);
/* The vocabulary as loadFromTextFile reads it (TemplatedVocabulary.h:1333-1416): node 0 is the root, node i >= 1 is line i
 * of the file with parent[i] < i, leaf[i] its isLeaf column (flagged nodes get word ids in node order), desc[i] its 32
 * descriptor bytes, weight[i] its weight; k, L, scoring, weighting the header line.  Children keep node order; a descent
 * ends at the first node without children, as isLeaf() does.  (A file written by saveToTextFile ends with a newline, which
 * the reference's loader turns into one more child of the root with an uninitialised descriptor; pass the nodes the file
 * lists.)  Blocking; replaces any previous vocabulary. */
ORBX_API int bowx_set_vocabulary(bowx_handle h, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent,
                                 const uint8_t* leaf, const uint8_t* desc, const double* weight);
/* info[6] = {k, L, scoring, weighting, nodes, words}: getBranchingFactor, getDepthLevels, getScoringType, getWeightingType, size */
ORBX_API int bowx_vocabulary_info(bowx_handle h, int32_t* info);
/* stopWords(minWeight) (:1316-1329): words lighter than min_weight get weight 0 and are ignored from then on; *count of them */
ORBX_API int bowx_stop_words(bowx_handle h, double min_weight, int32_t* count);
/* getParentNode(wid, levelsup) (:1264-1275) and getWordWeight(wid) (:1042-1045) */
ORBX_API int bowx_parent_node(bowx_handle h, uint32_t word, int levelsup, uint32_t* node);
ORBX_API int bowx_word_weight(bowx_handle h, uint32_t word, double* weight);
/* transform(feature, word, weight, nid, levelsup) (:1218-1260) for n descriptors [n][32]: the word each one falls into, the
 * weight of the node its descent ended at, and the node `levelsup` levels above the words.  Blocking, host buffers. */
ORBX_API int bowx_transform_features(bowx_handle h, const uint8_t* desc, int n, int levelsup, uint32_t* word, double* weight, uint32_t* node);
/* the same for the first d_counts[f] descriptors of every frame f of a device-resident [nframes][cap][32] array (the layout of
 * orbx_extract_batch_dev); outputs [nframes][cap]; asynchronous on the handle's stream */
ORBX_API int bowx_transform_features_dev(bowx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int levelsup,
                                         uint32_t* d_word, double* d_weight, uint32_t* d_node);
/* transform(features, BowVector&, FeatureVector&, levelsup) (:1128-1199) for nframes frames: desc [nframes][cap][32] with
 * counts[f] descriptors each (cap <= 16384).  Frame f's bag-of-words vector is bow_words / bow_vals [f*cap .. f*cap + nbow[f])
 * in ascending word order (the iteration order of the std::map); its feature vector has nfv[f] nodes fv_nodes[f*cap + g] in
 * ascending order, node g owning the feature indices fv_feats[f*cap + fv_offsets[f*(cap+1) + g] .. f*cap + fv_offsets[f*(cap+1) + g+1])
 * in the order they were added.  The four fv_* pointers may all be NULL (the overload without a feature vector, :1064-1122).
 * Blocking, host buffers. */
ORBX_API int bowx_transform_batch(bowx_handle h, const uint8_t* desc, const int32_t* counts, int nframes, int cap, int levelsup,
                                  uint32_t* bow_words, double* bow_vals, int32_t* nbow, uint32_t* fv_nodes, int32_t* fv_offsets,
                                  uint32_t* fv_feats, int32_t* nfv);
/* device-resident form, asynchronous on the handle's stream */
ORBX_API int bowx_transform_batch_dev(bowx_handle h, const uint8_t* d_desc, const int32_t* d_counts, int nframes, int cap, int levelsup,
                                      uint32_t* d_bow_words, double* d_bow_vals, int32_t* d_nbow, uint32_t* d_fv_nodes,
                                      int32_t* d_fv_offsets, uint32_t* d_fv_feats, int32_t* d_nfv);
/* score(v1, v2) of the vocabulary's scoring object (ScoringObject.cpp) for two vectors in ascending word order */
ORBX_API int bowx_score(bowx_handle h, const uint32_t* words1, const double* vals1, int n1, const uint32_t* words2, const double* vals2,
                        int n2, double* score);
/* One query vector (as v1) against nentries stored vectors (each as v2): entry e is db_words / db_vals
 * [db_start[e] .. db_start[e] + db_count[e]) of arrays db_len long -- a packed database or the padded output of
 * bowx_transform_batch alike.  scores [nentries].  nq <= 8192. */
ORBX_API int bowx_score_batch(bowx_handle h, const uint32_t* qwords, const double* qvals, int nq, const int64_t* db_start,
                              const int32_t* db_count, const uint32_t* db_words, const double* db_vals, int64_t db_len, int nentries,
                              double* scores);
ORBX_API int bowx_score_batch_dev(bowx_handle h, const uint32_t* d_qwords, const double* d_qvals, int nq, const int64_t* d_db_start,
                                  const int32_t* d_db_count, const uint32_t* d_db_words, const double* d_db_vals, int nentries,
                                  double* d_scores);

/* ------------------------------------------------------------------ frame ingest: baseline JPEG files
 * The reference reads every frame with cv::imread(path, CV_LOAD_IMAGE_UNCHANGED) (src/FrameLoader.cpp:62); for a .jpg that is
 * libjpeg behind OpenCV (default DCT method JDCT_ISLOW).  This family decodes a whole batch of such files on the GPU, straight
 * into the device-resident frames orbx_extract_batch_dev reads, bit for bit what cv2.imdecode returns (tests/golden/
 * jpeg_cases.npz): only the compressed bytes cross PCIe.  Handled: 8 bits, Huffman, baseline or extended sequential (SOF0 /
 * SOF1), one component (grey: one byte per pixel) or three (YCbCr with 4:2:0, 4:2:2 or 4:4:4 sampling, one interleaved scan:
 * B, G, R bytes per pixel as imread returns them -- libjpeg's fancy chroma upsampling and fixed-point colour conversion), any
 * quantisation / Huffman tables, with or without restart markers (both decode in parallel: self-synchronising subsequences
 * inside every restart interval).  Everything else -- progressive, arithmetic, 12-bit, other samplings, CMYK -- returns
 * ORBX_E_UNSUPPORTED and the caller keeps its CPU decoder for that file; damaged headers ORBX_E_INVALID. */
typedef struct jpgx_context* jpgx_handle;
ORBX_API int jpgx_create(jpgx_handle* out, int device);
ORBX_API int jpgx_destroy(jpgx_handle h);
ORBX_API int jpgx_set_stream(jpgx_handle h, void* cuda_stream);    /* same rules as hamx_set_stream */
ORBX_API int jpgx_get_stream(jpgx_handle h, void** cuda_stream);
ORBX_API int jpgx_synchronize(jpgx_handle h);
/* Headers only (no GPU work): info[6] = {width, height, restart interval in MCUs (0: none), 8x8 blocks, components (1 or 3),
 * luma sampling h*16 + v (0x11 for grey)}. */
ORBX_API int jpgx_probe(const uint8_t* file, size_t size, int32_t* info);
/* nfiles encoded files in host memory, all w x h, into device frames: frame i at d_frames + i*frame_pitch, rows `stride` bytes
 * apart.  The files are copied before the call returns; the decode is asynchronous on the handle's stream. */
ORBX_API int jpgx_decode_gray_batch_dev(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int h_,
                                        uint8_t* d_frames, size_t frame_pitch, size_t stride);
/* the same into host memory (blocking) */
ORBX_API int jpgx_decode_gray_batch(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int h_, uint8_t* frames,
                                    size_t frame_pitch, size_t stride);
/* three-component files into BGR frames (3 bytes per pixel, rows `stride` >= 3 w bytes apart): what orbx_set_input_channels(h, 3)
 * makes the extractor read.  All files of a batch share the chroma sampling.  A grey file here, or a colour file in the calls
 * above, returns ORBX_E_UNSUPPORTED. */
ORBX_API int jpgx_decode_bgr_batch_dev(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int h_,
                                       uint8_t* d_frames, size_t frame_pitch, size_t stride);
ORBX_API int jpgx_decode_bgr_batch(jpgx_handle h, const uint8_t* const* files, const size_t* sizes, int nfiles, int w, int h_, uint8_t* frames,
                                   size_t frame_pitch, size_t stride);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
