"""ctypes binding of liborbx.so (include/orbx.h).  No CPU fallback: a missing library or device raises."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborbx.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]
)
DMATCH_DTYPE = np.dtype([("query_idx", "<i4"), ("train_idx", "<i4"), ("img_idx", "<i4"), ("distance", "<f4")])
TOP2_DTYPE = np.dtype([("dist0", "<i4"), ("idx0", "<i4"), ("dist1", "<i4"), ("idx1", "<i4")])
CAMERAS_DTYPE = np.dtype([("Rt1", "<f8", (3, 4)), ("Rt2", "<f8", (3, 4)), ("K1", "<f8", (3, 3)), ("K2", "<f8", (3, 3))])   # trx_cameras
assert KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16 and TOP2_DTYPE.itemsize == 16 and CAMERAS_DTYPE.itemsize == 336

HARRIS_SCORE, FAST_SCORE = 0, 1
KERNEL_AUTO, KERNEL_INTEGER, KERNEL_TENSOR = 0, 1, 2     # hamx_set_kernel
OK, E_INVALID, E_CUDA, E_CAPACITY, E_ALLOC, E_ALIGN, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6

# every symbol include/orbx.h declares (tests/test_abi.py checks the header against this list and the built library)
SYMBOLS = [
    "orbx_last_error", "orbx_version", "orbx_device_count", "orbx_default_params", "orbx_create", "orbx_destroy",
    "orbx_set_stream", "orbx_synchronize", "orbx_max_keypoints", "orbx_level_info", "orbx_detect", "orbx_compute",
    "orbx_detect_and_compute", "orbx_extract_batch", "orbx_extract_batch_dev", "orbx_check_dev",
    "orbx_debug_pyramid_level", "orbx_debug_fast_level", "hamx_create", "hamx_destroy", "hamx_set_stream",
    "hamx_synchronize", "hamx_knn2", "hamx_match_ratio", "hamx_knn2_dev", "hamx_merge_top2_dev", "hamx_ratio_dev",
    "hamx_popc_peak", "hamx_match_pairs_dev", "hamx_match_consecutive_dev", "orbx_match_consecutive", "orbx_reset_sequence",
    "orbx_set_profiling", "orbx_read_profile", "orbx_submit_batch", "orbx_wait_batch", "orbx_batches_in_flight",
    "hamx_p2p_export", "hamx_p2p_import", "hamx_p2p_import_ptrs", "hamx_p2p_close", "hamx_knn2_p2p_dev",
    "hamx_knn2_p2p_scatter_dev", "hamx_p2p_merge_dev", "orbx_set_input_channels", "orbx_pipeline_depth", "hamx_match_back_dev", "hamx_update_history_dev", "orbx_match_back", "orbx_host_alloc", "orbx_host_alloc_wc", "orbx_host_free",
    "fmx_create", "fmx_destroy", "fmx_set_stream", "fmx_synchronize", "fmx_compute_fundamental", "fmx_fundamental_batch",
    "fmx_last_info", "fmx_fundamental_batch_dev", "fmx_filter_consecutive_dev", "orbx_filter_consecutive", "orbx_submit_batch_filtered", "fmx_filter_back_dev", "orbx_filter_back", "orbx_submit_batch_back",
    "hamx_get_stream", "fmx_get_stream", "hamx_reserve", "hamx_set_kernel", "hamx_nbest", "hamx_nbest_dev", "hamx_loop_score", "hamx_loop_score_dev", "hamx_loop_best_dev",
    "trx_create", "trx_destroy", "trx_set_stream", "trx_get_stream", "trx_synchronize", "trx_triangulate", "trx_triangulate_hypotheses",
    "trx_triangulate_batch", "trx_triangulate_batch_dev", "trx_triangulate_back_dev", "trx_associate_dev", "trx_select_new_dev",
    "bowx_create", "bowx_destroy", "bowx_set_stream", "bowx_get_stream", "bowx_synchronize", "bowx_set_vocabulary", "bowx_vocabulary_info",
    "bowx_stop_words", "bowx_parent_node", "bowx_word_weight", "bowx_transform_features", "bowx_transform_features_dev",
    "bowx_transform_batch", "bowx_transform_batch_dev", "bowx_score", "bowx_score_batch", "bowx_score_batch_dev",
    "jpgx_create", "jpgx_destroy", "jpgx_set_stream", "jpgx_get_stream", "jpgx_synchronize", "jpgx_probe", "jpgx_decode_gray_batch_dev",
    "jpgx_decode_gray_batch", "jpgx_decode_bgr_batch_dev", "jpgx_decode_bgr_batch",
]
NSTAGES = 5
STAGE_NAMES = ("pyramid", "fast", "select", "harris_select", "orient_describe")


class OrbxError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("liborbx status %d: %s" % (status, message))
        self.status = status


class Params(C.Structure):
    """cv::ORB constructor arguments (reference defaults: src/FeatureExtractor.h:23-24)."""

    _fields_ = [
        ("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32), ("edge_threshold", C.c_int32),
        ("first_level", C.c_int32), ("wta_k", C.c_int32), ("score_type", C.c_int32), ("patch_size", C.c_int32),
        ("fast_threshold", C.c_int32),
    ]


def build(force=False, verbose=False):
    """Compile liborbx.so in-tree with nvcc for sm_100a (monocular_slam_b200/csrc/build.sh)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc")) if f.endswith((".cu", ".cuh", ".inc", ".sh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "orbx.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    out = subprocess.run(["sh", os.path.join(_HERE, "csrc", "build.sh")], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise RuntimeError("building liborbx.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load liborbx.so; raises if it has not been built (there is deliberately no fallback implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrbxError(E_CUDA, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(monocular_slam_b200 has no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    ip, fp, dp = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_version.restype = C.c_char_p
    L.orbx_default_params.argtypes = [C.POINTER(Params)]
    L.orbx_default_params.restype = None
    L.orbx_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_set_stream.argtypes = [vp, vp]
    L.orbx_synchronize.argtypes = [vp]
    L.orbx_max_keypoints.argtypes = [vp]
    L.orbx_level_info.argtypes = [vp, C.c_int, C.c_int, i32p, i32p, fp, i32p]
    L.orbx_detect.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, ip]
    L.orbx_compute.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, ip, vp]
    L.orbx_detect_and_compute.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, C.c_int, ip]
    L.orbx_extract_batch.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_size_t, vp, vp, C.c_int, i32p]
    L.orbx_extract_batch_dev.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_size_t, vp, vp, C.c_int, vp]
    L.orbx_check_dev.argtypes = [vp]
    L.orbx_debug_pyramid_level.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, C.c_int, vp]
    L.orbx_debug_fast_level.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, C.c_int, i32p, i32p, i32p, C.c_int, ip]
    L.hamx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.hamx_destroy.argtypes = [vp]
    L.hamx_set_stream.argtypes = [vp, vp]
    L.trx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.trx_destroy.argtypes = [vp]
    L.trx_set_stream.argtypes = [vp, vp]
    L.trx_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.trx_synchronize.argtypes = [vp]
    L.trx_triangulate.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, i32p]
    L.trx_triangulate_hypotheses.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, i32p]
    L.trx_triangulate_batch.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp]
    L.trx_triangulate_batch_dev.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp]
    L.trx_triangulate_back_dev.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.trx_associate_dev.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.trx_select_new_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.jpgx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.jpgx_destroy.argtypes = [vp]
    L.jpgx_set_stream.argtypes = [vp, vp]
    L.jpgx_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.jpgx_synchronize.argtypes = [vp]
    L.jpgx_probe.argtypes = [vp, C.c_size_t, i32p]
    L.jpgx_decode_gray_batch_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t]
    L.jpgx_decode_gray_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t]
    L.jpgx_decode_bgr_batch_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t]
    L.jpgx_decode_bgr_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t]
    L.bowx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.bowx_destroy.argtypes = [vp]
    L.bowx_set_stream.argtypes = [vp, vp]
    L.bowx_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.bowx_synchronize.argtypes = [vp]
    L.bowx_set_vocabulary.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.bowx_vocabulary_info.argtypes = [vp, i32p]
    L.bowx_stop_words.argtypes = [vp, C.c_double, i32p]
    L.bowx_parent_node.argtypes = [vp, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
    L.bowx_word_weight.argtypes = [vp, C.c_uint32, dp]
    L.bowx_transform_features.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.bowx_transform_features_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.bowx_transform_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.bowx_transform_batch_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.bowx_score.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, dp]
    L.bowx_score_batch.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int64, C.c_int, vp]
    L.bowx_score_batch_dev.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp]
    L.hamx_nbest.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
    L.hamx_nbest_dev.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
    L.hamx_loop_score.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, i32p]
    L.hamx_loop_score_dev.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.hamx_loop_best_dev.argtypes = [vp, vp, C.c_int, vp]
    L.hamx_set_kernel.argtypes = [vp, C.c_int]
    L.hamx_reserve.argtypes = [vp, C.c_int64, C.c_int64, C.c_int]
    L.hamx_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.fmx_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.hamx_synchronize.argtypes = [vp]
    L.hamx_knn2.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, vp, i32p]
    L.hamx_match_ratio.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_float, vp, i64p]
    L.hamx_knn2_dev.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_int64, vp]
    L.hamx_merge_top2_dev.argtypes = [vp, vp, C.c_int, C.c_int64, vp]
    L.hamx_ratio_dev.argtypes = [vp, vp, C.c_int64, C.c_float, vp, vp]
    L.hamx_popc_peak.argtypes = [C.c_int, dp, dp]
    L.hamx_match_pairs_dev.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, C.c_size_t, vp]
    L.hamx_match_consecutive_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_float, vp, vp]
    L.orbx_match_consecutive.argtypes = [vp, vp, C.c_float, C.c_int, C.c_int, vp, i64p]
    L.orbx_reset_sequence.argtypes = [vp]
    L.orbx_submit_batch.argtypes = [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_float, vp, vp, C.c_int, vp, vp, vp]
    L.orbx_wait_batch.argtypes = [vp]
    L.orbx_batches_in_flight.argtypes = [vp]
    L.orbx_set_input_channels.argtypes = [vp, C.c_int]
    L.orbx_pipeline_depth.argtypes = [vp]
    L.orbx_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.orbx_host_alloc_wc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.orbx_host_free.argtypes = [vp]
    L.hamx_match_back_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_float, vp, vp]
    L.hamx_update_history_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    L.orbx_match_back.argtypes = [vp, vp, C.c_int, C.c_float, C.c_int, C.c_int, vp, i64p]
    L.hamx_p2p_export.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.hamx_p2p_import.argtypes = [vp, vp]
    L.hamx_p2p_import_ptrs.argtypes = [vp, C.POINTER(vp)]
    L.hamx_p2p_close.argtypes = [vp]
    L.hamx_knn2_p2p_dev.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_int64, vp]
    L.hamx_knn2_p2p_scatter_dev.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.c_int64]
    L.hamx_p2p_merge_dev.argtypes = [vp, C.c_int64, vp]
    L.fmx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.fmx_destroy.argtypes = [vp]
    L.fmx_set_stream.argtypes = [vp, vp]
    L.fmx_synchronize.argtypes = [vp]
    L.fmx_compute_fundamental.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_double, C.c_double, vp, dp, i32p]
    L.fmx_fundamental_batch.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp]
    L.fmx_last_info.argtypes = [vp, C.c_int, vp]
    L.fmx_fundamental_batch_dev.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp]
    L.fmx_filter_consecutive_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_double, C.c_double, vp, vp, vp]
    L.orbx_filter_consecutive.argtypes = [vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp, vp]
    L.orbx_filter_back.argtypes = [vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.orbx_submit_batch_back.argtypes = [vp, vp, vp, C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_float, vp, vp, C.c_int, vp,
                                         vp, vp, C.c_double, C.c_double, vp, vp, vp]
    L.fmx_filter_back_dev.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_double, C.c_double, vp, vp, vp]
    L.orbx_submit_batch_filtered.argtypes = [vp, vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_float, vp, vp, C.c_int, vp, vp,
                                             vp, C.c_double, C.c_double, vp, vp, vp]
    L.orbx_set_profiling.argtypes = [vp, C.c_int]
    L.orbx_read_profile.argtypes = [vp, fp, ip]
    _lib = L
    return L


def check(status):
    if status != 0:
        raise OrbxError(status, lib().orbx_last_error().decode("utf-8", "replace"))
