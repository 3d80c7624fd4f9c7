"""Python mirror of the reference's call surface for the ORB front-end, over the C ABI of liborbx.so.

Names and argument meaning follow the reference (and the OpenCV 2.4 objects it uses):

* ``OrbFeatureDetector.detect`` / ``OrbDescriptorExtractor.compute``  -- src/FeatureExtractor.cpp:17,19
  (members declared at src/FeatureExtractor.h:23-24); both are the same ``ORB`` class, as in OpenCV.
* ``BFMatcher(NORM_HAMMING, False).knnMatch(q, t, 2)``               -- src/CameraPoseEstimator.cpp:202-204
* ``match_features(d1, d2, ratio)``                                   -- matchFeatures, src/CameraPoseEstimator.cpp:200-213
* ``FeatureExtractor`` (init / process / destroy)                     -- ProcessingNode, src/ProcessingNode.h:16-32

Keypoints are numpy structured arrays with cv::KeyPoint's fields (``KEYPOINT_DTYPE``), descriptors ``N x 32 uint8``,
matches structured arrays with cv::DMatch's fields (``DMATCH_DTYPE``).  Everything runs on the GPU through the
library; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DMATCH_DTYPE, FAST_SCORE, HARRIS_SCORE, KEYPOINT_DTYPE, TOP2_DTYPE, OrbxError, Params, check

NORM_HAMMING = 6   # cv::NORM_HAMMING


def _gray(image):
    """An h x w (CV_8UC1) or h x w x 3 (CV_8UC3, BGR) uint8 image with contiguous rows; 3-channel frames are converted to
    gray on the device, as cv::ORB does on the CPU."""
    img = np.asarray(image)
    if img.dtype != np.uint8 or not (img.ndim == 2 or (img.ndim == 3 and img.shape[2] == 3)):
        raise ValueError("expected an h x w (gray) or h x w x 3 (BGR) uint8 image, got %s %s" % (img.dtype, img.shape))
    if img.ndim == 2 and (img.strides[1] != 1 or img.strides[0] < img.shape[1]):
        img = np.ascontiguousarray(img)
    if img.ndim == 3 and (img.strides[2] != 1 or img.strides[1] != 3 or img.strides[0] < 3 * img.shape[1]):
        img = np.ascontiguousarray(img)
    return img


class ORB:
    """cv::ORB with the reference's defaults; only ``nfeatures``, ``scoreType``, ``nlevels``, ``scaleFactor`` and
    ``fastThreshold`` may be changed (see include/orbx.h)."""

    def __init__(self, nfeatures=500, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                 scoreType=HARRIS_SCORE, patchSize=31, fastThreshold=20, device=0, max_size=(1920, 1080), max_batch=1):
        self._h = C.c_void_p()
        self.params = Params(nfeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K, scoreType, patchSize,
                             fastThreshold)
        self.device = device
        self.max_batch = max_batch
        check(_lib.lib().orbx_create(C.byref(self._h), C.byref(self.params), device, int(max_size[0]), int(max_size[1]),
                                     max_batch))
        self.max_keypoints = _lib.lib().orbx_max_keypoints(self._h)
        self.default_cap = min(self.max_keypoints, nfeatures + max(nfeatures // 4, 512))

    def _set_channels(self, img):
        ch = 3 if img.ndim == 3 else 1
        if ch != getattr(self, "_channels", 1):
            check(_lib.lib().orbx_set_input_channels(self._h, ch))
            self._channels = ch

    def set_input_channels(self, channels):
        """For the raw-pointer (_dev) entry points: 1 = gray, 3 = BGR interleaved."""
        check(_lib.lib().orbx_set_input_channels(self._h, int(channels)))
        self._channels = int(channels)

    # -- lifetime (ProcessingNode::destroy)
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().orbx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(_lib.lib().orbx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(_lib.lib().orbx_synchronize(self._h))

    def level_info(self, w, h):
        n = self.params.nlevels
        ws, hs, qs = (np.zeros(n, np.int32) for _ in range(3))
        sc = np.zeros(n, np.float32)
        i32p = C.POINTER(C.c_int32)
        check(_lib.lib().orbx_level_info(self._h, w, h, ws.ctypes.data_as(i32p), hs.ctypes.data_as(i32p),
                                         sc.ctypes.data_as(C.POINTER(C.c_float)), qs.ctypes.data_as(i32p)))
        return ws, hs, sc, qs

    def _retry(self, fn, cap):
        """Call fn(cap); on ORBX_E_CAPACITY (ties can return more than nfeatures) retry once with the maximum."""
        try:
            return fn(cap)
        except OrbxError as e:
            if e.status != _lib.E_CAPACITY or cap >= self.max_keypoints:
                raise
            return fn(self.max_keypoints)

    # -- OrbFeatureDetector::detect(image, keypoints)
    def detect(self, image, cap=None):
        img = _gray(image)
        self._set_channels(img)
        self._last_batch = None

        def run(cap):
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            n = C.c_int(0)
            check(_lib.lib().orbx_detect(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                         kps.ctypes.data, cap, C.byref(n)))
            return kps[:n.value].copy()

        return self._retry(run, cap or self.default_cap)

    # -- OrbDescriptorExtractor::compute(image, keypoints, descriptors)
    def compute(self, image, keypoints):
        img = _gray(image)
        self._set_channels(img)
        self._last_batch = None
        kps = np.array(keypoints, dtype=KEYPOINT_DTYPE, copy=True)
        n = C.c_int(len(kps))
        desc = np.zeros((max(len(kps), 1), 32), np.uint8)
        check(_lib.lib().orbx_compute(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                      kps.ctypes.data, C.byref(n), desc.ctypes.data))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def detectAndCompute(self, image, cap=None):
        img = _gray(image)
        self._set_channels(img)
        self._last_batch = None

        def run(cap):
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = C.c_int(0)
            check(_lib.lib().orbx_detect_and_compute(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                                     kps.ctypes.data, desc.ctypes.data, cap, C.byref(n)))
            return kps[:n.value].copy(), desc[:n.value].copy()

        return self._retry(run, cap or self.default_cap)

    # -- a batch of frames of one size (sequence extraction; frames are independent, src/main.cpp:36-51)
    def extract_batch(self, frames, cap=None, out=None):
        """frames: sequence of h x w uint8 arrays (or one n x h x w array).  Returns (kps[n, cap], desc[n, cap, 32], counts[n]).
        ``out`` may carry preallocated (kps, desc, counts) buffers, e.g. pinned memory."""
        frames = [_gray(f) for f in frames]
        self._set_channels(frames[0])
        n = len(frames)
        h, w = frames[0].shape[:2]
        stride = frames[0].strides[0]
        if any(f.shape != frames[0].shape or f.strides[0] != stride for f in frames):
            raise ValueError("all frames of a batch must share one shape and stride")
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])

        def run(cap):
            if out is not None and out[0].shape[1] == cap:
                kps, desc, counts = out
            else:
                kps = np.zeros((n, cap), KEYPOINT_DTYPE)
                desc = np.zeros((n, cap, 32), np.uint8)
                counts = np.zeros(n, np.int32)
            self._last_batch = None
            check(_lib.lib().orbx_extract_batch(self._h, ptrs, n, w, h, stride, kps.ctypes.data, desc.ctypes.data, cap,
                                                counts.ctypes.data_as(C.POINTER(C.c_int32))))
            # the geometry the sequence calls below size their buffers from: _retry may have raised the capacity
            self._last_batch = (n, cap)
            return kps, desc, counts

        return self._retry(run, cap or self.default_cap)

    def _seq_geometry(self, cap, nframes, what):
        """(nframes, cap) of the batch the handle holds; the optional caller values must agree with it (the library checks
        the same through its own arguments, so a stale value fails instead of overflowing a buffer)."""
        last = getattr(self, "_last_batch", None)
        if last is None:
            raise ValueError("%s: no batch has been extracted with extract_batch on this ORB object" % what)
        n, c = last
        if nframes is not None and nframes != n:
            raise ValueError("%s: nframes = %d, but the last extract_batch held %d frames" % (what, nframes, n))
        if cap is not None and cap != c:
            raise ValueError("%s: cap = %d, but the last extract_batch ran with cap %d (a capacity retry raises it)" % (what, cap, c))
        return n, c

    def extract_batch_dev(self, d_frames_ptr, frame_pitch, nframes, w, h, stride, d_kps_ptr, d_desc_ptr, cap, d_counts_ptr):
        """Device-resident variant (raw device pointers, asynchronous on the handle's stream)."""
        check(_lib.lib().orbx_extract_batch_dev(self._h, d_frames_ptr, frame_pitch, nframes, w, h, stride, d_kps_ptr,
                                                d_desc_ptr, cap, d_counts_ptr))

    def check_dev(self):
        check(_lib.lib().orbx_check_dev(self._h))

    # -- sequence mode: consecutive-frame matching of the batch last extracted, descriptors stay on the device
    def match_consecutive(self, matcher, ratio, cap=None, nframes=None, out=None):
        """matchFeatures(desc[f], desc[f-1], ratio) for every frame f of the last extract_batch (frame 0 against the last
        frame of the previous batch).  Returns (good[nframes, cap] DMATCH_DTYPE, ngood[nframes]); the buffers are sized from
        the geometry of that extract_batch, ``cap`` / ``nframes`` are only cross-checked against it."""
        n, c = self._seq_geometry(cap, nframes, "match_consecutive")
        if out is not None:
            good, ngood = out
            if good.shape != (n, c) or good.dtype != DMATCH_DTYPE or ngood.shape != (n,) or ngood.dtype != np.int64 \
                    or not good.flags.c_contiguous:
                raise ValueError("match_consecutive: out must be (good[%d, %d] DMATCH_DTYPE, ngood[%d] int64)" % (n, c, n))
        else:
            good = np.zeros((n, c), DMATCH_DTYPE)
            ngood = np.zeros(n, np.int64)
        check(_lib.lib().orbx_match_consecutive(self._h, matcher._h, float(ratio), n, c, good.ctypes.data,
                                                ngood.ctypes.data_as(C.POINTER(C.c_int64))))
        return good, ngood

    def filter_consecutive(self, fundamental, cap=None, nframes=None, max_distance=3.0, confidence=0.85):
        """computeFundamentalMatrix (src/CameraPoseEstimator.cpp:545-586) for every (frame f, frame f-1) pair of the batch
        match_consecutive just ran on; the match lists stay on the device in between.
        Returns (status[nframes, cap] uint8 aligned with good[f], F[nframes, 3, 3], ninliers[nframes])."""
        n, c = self._seq_geometry(cap, nframes, "filter_consecutive")
        status = np.zeros((n, c), np.uint8)
        F = np.zeros((n, 3, 3), np.float64)
        ninl = np.zeros(n, np.int32)
        check(_lib.lib().orbx_filter_consecutive(self._h, fundamental._h, float(max_distance), float(confidence), n, c,
                                                 status.ctypes.data, F.ctypes.data, ninl.ctypes.data))
        return status, F, ninl

    def filter_back(self, fundamental, back, cap=None, nframes=None, max_distance=3.0, confidence=0.85):
        """computeFundamentalMatrix for every (frame f, frame f-j) pair match_back just matched (the loop at
        src/CameraPoseEstimator.cpp:405-419).  Returns (status[nframes, back, cap], F[nframes, back, 3, 3], ninliers[nframes, back])."""
        n, c = self._seq_geometry(cap, nframes, "filter_back")
        status = np.zeros((n, back, c), np.uint8)
        F = np.zeros((n, back, 3, 3), np.float64)
        ninl = np.zeros((n, back), np.int32)
        check(_lib.lib().orbx_filter_back(self._h, fundamental._h, float(max_distance), float(confidence), n, int(back), c,
                                          status.ctypes.data, F.ctypes.data, ninl.ctypes.data))
        return status, F, ninl

    # -- pipelined sequence mode: pipeline_depth() batches in flight (upload / kernels / download overlap across batches)
    def submit_batch(self, frames, matcher, ratio, out, fundamental=None, max_distance=3.0, confidence=0.85, back=0):
        """Enqueue extraction (+ consecutive-frame matching when ``matcher`` is given) of a batch and return at once.
        ``out`` = (kps[n, cap], desc[n, cap, 32], counts[n] int32, good[n, cap], ngood[n] int64): caller-owned buffers
        (pinned for real overlap) that are complete after the matching ``wait_batch()``.  With ``fundamental`` (a
        FundamentalFilter) the outlier filter of every (frame, predecessor) pair runs in the same submission and ``out``
        carries three more buffers: status[n, cap] uint8, F[n, 3, 3] float64, ninliers[n] int32.  With ``back`` >= 1 every
        frame is paired with its ``back`` predecessors instead (the numBackTraverse loop, src/CameraPoseEstimator.cpp:405-419):
        good[n, back, cap], ngood[n, back], status[n, back, cap], F[n, back, 3, 3], ninliers[n, back]."""
        frames = [_gray(f) for f in frames]
        self._set_channels(frames[0])
        n = len(frames)
        h, w = frames[0].shape[:2]
        stride = frames[0].strides[0]
        if any(f.shape != frames[0].shape or f.strides[0] != stride for f in frames):
            raise ValueError("all frames of a batch must share one shape and stride")
        kps, desc, counts, good, ngood = out[:5]
        cap = kps.shape[1]
        if kps.shape[0] < n or desc.shape[:2] != kps.shape[:2] or counts.dtype != np.int32 or ngood.dtype != np.int64:
            raise ValueError("bad output buffers")
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        self._last_batch = None          # match_consecutive / match_back pair with extract_batch only
        self._inflight = getattr(self, "_inflight", [])
        if back:
            if good.shape[1:] != (back, cap) or ngood.shape[1:] != (back,):
                raise ValueError("bad match output buffers for back = %d" % back)
            st = F = ni = None
            if fundamental is not None:
                st, F, ni = out[5:8]
                if st.shape[1:] != (back, cap) or st.dtype != np.uint8 or F.dtype != np.float64 or ni.dtype != np.int32:
                    raise ValueError("bad filter output buffers")
            check(_lib.lib().orbx_submit_batch_back(self._h, matcher._h, fundamental._h if fundamental is not None else None, int(back),
                                                    ptrs, n, w, h, stride, float(ratio), kps.ctypes.data, desc.ctypes.data, cap,
                                                    counts.ctypes.data, good.ctypes.data, ngood.ctypes.data, float(max_distance),
                                                    float(confidence), st.ctypes.data if st is not None else None,
                                                    F.ctypes.data if F is not None else None, ni.ctypes.data if ni is not None else None))
        elif fundamental is not None:
            status, F, ninl = out[5:8]
            if status.shape[:2] != kps.shape[:2] or status.dtype != np.uint8 or F.dtype != np.float64 or ninl.dtype != np.int32:
                raise ValueError("bad filter output buffers")
            check(_lib.lib().orbx_submit_batch_filtered(self._h, matcher._h, fundamental._h, ptrs, n, w, h, stride, float(ratio),
                                                        kps.ctypes.data, desc.ctypes.data, cap, counts.ctypes.data, good.ctypes.data,
                                                        ngood.ctypes.data, float(max_distance), float(confidence), status.ctypes.data,
                                                        F.ctypes.data, ninl.ctypes.data))
        else:
            check(_lib.lib().orbx_submit_batch(self._h, matcher._h if matcher is not None else None, ptrs, n, w, h, stride,
                                               float(ratio), kps.ctypes.data, desc.ctypes.data, cap, counts.ctypes.data,
                                               good.ctypes.data, ngood.ctypes.data))
        self._inflight.append((frames, out))     # keep the buffers alive until the batch is collected

    def wait_batch(self):
        """Block until the oldest submitted batch is complete; returns its ``out`` tuple."""
        status = _lib.lib().orbx_wait_batch(self._h)
        done = self._inflight.pop(0)[1] if getattr(self, "_inflight", None) else None
        check(status)
        return done

    def pipeline_depth(self):
        """How many submitted batches may be in flight before wait_batch() has to be called."""
        return _lib.lib().orbx_pipeline_depth(self._h)

    def batches_in_flight(self):
        return _lib.lib().orbx_batches_in_flight(self._h)

    def match_back(self, matcher, back, ratio, cap=None, nframes=None):
        """matchFeatures(desc[f], desc[f-j], ratio) for j = 1..back and every frame f of the last extract_batch -- the
        steady-state loop of CameraPoseEstimator::pnpPoseEstimation (numBackTraverse = 5, src/CameraPoseEstimator.cpp:405-409);
        frames before the batch come from the handle's history.  Returns (good[nframes, back, cap], ngood[nframes, back])."""
        n, c = self._seq_geometry(cap, nframes, "match_back")
        good = np.zeros((n, back, c), DMATCH_DTYPE)
        ngood = np.zeros((n, back), np.int64)
        check(_lib.lib().orbx_match_back(self._h, matcher._h, int(back), float(ratio), n, c, good.ctypes.data,
                                         ngood.ctypes.data_as(C.POINTER(C.c_int64))))
        return good, ngood

    def reset_sequence(self):
        check(_lib.lib().orbx_reset_sequence(self._h))

    def set_profiling(self, enabled):
        check(_lib.lib().orbx_set_profiling(self._h, int(bool(enabled))))

    def read_profile(self):
        """Average device milliseconds per batch of each stage since profiling was enabled: dict name -> ms, and #batches."""
        ms = (C.c_float * _lib.NSTAGES)()
        n = C.c_int(0)
        check(_lib.lib().orbx_read_profile(self._h, ms, C.byref(n)))
        return dict(zip(_lib.STAGE_NAMES, [float(v) for v in ms])), n.value

    # -- stage taps for parity tests
    def debug_pyramid_level(self, image, level):
        img = _gray(image)
        self._set_channels(img)
        ws, hs, _, _ = self.level_info(img.shape[1], img.shape[0])
        out = np.zeros((hs[level], ws[level]), np.uint8)
        check(_lib.lib().orbx_debug_pyramid_level(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], level,
                                                  out.ctypes.data))
        return out

    def debug_fast_level(self, image, level):
        img = _gray(image)
        self._set_channels(img)
        ws, hs, _, _ = self.level_info(img.shape[1], img.shape[0])
        cap = (int(ws[level]) // 2 + 1) * (int(hs[level]) // 2 + 1)
        xs, ys, sc = (np.zeros(cap, np.int32) for _ in range(3))
        n = C.c_int(0)
        i32p = C.POINTER(C.c_int32)
        check(_lib.lib().orbx_debug_fast_level(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], level,
                                               xs.ctypes.data_as(i32p), ys.ctypes.data_as(i32p), sc.ctypes.data_as(i32p), cap,
                                               C.byref(n)))
        return xs[:n.value].copy(), ys[:n.value].copy(), sc[:n.value].copy()


# The reference declares one object of each type (src/FeatureExtractor.h:23-24); in OpenCV 2.4 both are typedefs of cv::ORB.
OrbFeatureDetector = ORB
OrbDescriptorExtractor = ORB


def ORB_create(nfeatures=500, scoreType=HARRIS_SCORE, **kw):
    return ORB(nfeatures=nfeatures, scoreType=scoreType, **kw)


class BFMatcher:
    """cv::BFMatcher(NORM_HAMMING, crossCheck=false) as constructed at src/CameraPoseEstimator.cpp:202."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False, device=0):
        if normType != NORM_HAMMING or crossCheck:
            raise ValueError("only BFMatcher(NORM_HAMMING, False) -- the reference's configuration -- is provided")
        self._h = C.c_void_p()
        self.device = device
        check(_lib.lib().hamx_create(C.byref(self._h), device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().hamx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(_lib.lib().hamx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(_lib.lib().hamx_synchronize(self._h))

    def reserve(self, nq, nt, npairs=1):
        """Size the workspaces once so that later *_dev calls of up to this shape never allocate (hamx_reserve)."""
        check(_lib.lib().hamx_reserve(self._h, int(nq), int(nt), int(npairs)))

    @staticmethod
    def _desc(d):
        d = np.ascontiguousarray(d, dtype=np.uint8)
        if d.ndim != 2 or d.shape[1] != 32:
            if d.size == 0:
                return d.reshape(0, 32)
            raise ValueError("descriptors must be N x 32 uint8 (CV_8U, 256-bit ORB), got %s" % (d.shape,))
        return d

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2):
        """Returns (matches[nq, 2] DMATCH_DTYPE, counts[nq]): row i holds counts[i] = min(nt, 2) valid entries sorted by
        (distance, trainIdx), exactly like vector<vector<DMatch>> from knnMatch(q, t, out, 2)."""
        if k != 2:
            raise ValueError("the reference calls knnMatch with k = 2 only")
        q, t = self._desc(queryDescriptors), self._desc(trainDescriptors)
        out = np.zeros((len(q), 2), DMATCH_DTYPE)
        counts = np.zeros(len(q), np.int32)
        check(_lib.lib().hamx_knn2(self._h, q.ctypes.data, len(q), t.ctypes.data, len(t), out.ctypes.data,
                                   counts.ctypes.data_as(C.POINTER(C.c_int32))))
        return out, counts

    def match_ratio(self, d1, d2, ratio):
        q, t = self._desc(d1), self._desc(d2)
        good = np.zeros(max(len(q), 1), DMATCH_DTYPE)
        n = C.c_int64(0)
        check(_lib.lib().hamx_match_ratio(self._h, q.ctypes.data, len(q), t.ctypes.data, len(t), float(ratio), good.ctypes.data,
                                          C.byref(n)))
        return good[:n.value].copy()

    # -- loop-closure candidate scoring (src/LoopCloser.cpp:19-105 with the Hamming distance)
    def NBestMatches(self, descriptors1, descriptors2, n=10):
        """LoopCloser::NBestMatches: (distances[nq, n] int32, indices[nq, n] int32), ascending, absent entries -1."""
        q, t = self._desc(descriptors1), self._desc(descriptors2)
        dist = np.full((len(q), n), -1, np.int32)
        idx = np.full((len(q), n), -1, np.int32)
        check(_lib.lib().hamx_nbest(self._h, q.ctypes.data, len(q), t.ctypes.data, len(t), int(n), dist.ctypes.data, idx.ctypes.data))
        return dist, idx

    def loop_score(self, descriptors1, frames_desc, frame_counts, n=10, thr=40):
        """The loop of LoopCloser::DetectLoop over all stored frames in one launch: frames_desc uint8 [nframes, cap, 32],
        frame_counts [nframes].  Returns (scores[nframes] int32, best frame or -1)."""
        q = self._desc(descriptors1)
        fd = np.ascontiguousarray(frames_desc, np.uint8)
        if fd.ndim != 3 or fd.shape[2] != 32:
            raise ValueError("frames_desc must be [nframes, cap, 32] uint8")
        fc = np.ascontiguousarray(frame_counts, np.int32)
        if fc.shape != (fd.shape[0],):
            raise ValueError("frame_counts must be [nframes]")
        scores = np.zeros(max(fd.shape[0], 1), np.int32)
        best = C.c_int32(-1)
        check(_lib.lib().hamx_loop_score(self._h, q.ctypes.data, len(q), fd.ctypes.data, fc.ctypes.data, fd.shape[0], max(fd.shape[1], 1), int(n),
                                         int(thr), scores.ctypes.data, C.byref(best)))
        return scores[:fd.shape[0]], int(best.value)

    def loop_score_dev(self, d_q, nq, d_frames, d_counts, nframes, cap, n, thr, d_scores, d_best=0):
        check(_lib.lib().hamx_loop_score_dev(self._h, d_q, nq, d_frames, d_counts, nframes, cap, int(n), int(thr), d_scores, d_best or None))

    def loop_best_dev(self, d_scores, nframes, d_best):
        check(_lib.lib().hamx_loop_best_dev(self._h, d_scores, nframes, d_best))

    def nbest_dev(self, d_q, nq, d_t, nt, n, d_dist, d_idx):
        check(_lib.lib().hamx_nbest_dev(self._h, d_q, nq, d_t, nt, int(n), d_dist, d_idx))

    # device-resident pieces (raw device pointers; asynchronous on the handle's stream)
    def knn2_dev(self, d_q, nq, d_t, nt, train_offset, d_out):
        check(_lib.lib().hamx_knn2_dev(self._h, d_q, nq, d_t, nt, train_offset, d_out))

    def set_kernel(self, mode):
        """_lib.KERNEL_AUTO (by problem size), KERNEL_INTEGER (XOR + POPC) or KERNEL_TENSOR (tcgen05 int8 contraction); the
        results are identical."""
        check(_lib.lib().hamx_set_kernel(self._h, int(mode)))

    # -- train-sharded matching over peer memory (include/orbx.h "hamx_p2p_*")
    def p2p_export(self, nq_max, world, rank):
        """Allocate this rank's gather buffer; returns (64-byte cudaIpc handle as bytes, local device base address)."""
        buf = (C.c_uint8 * 64)()
        base = C.c_void_p()
        check(_lib.lib().hamx_p2p_export(self._h, int(nq_max), int(world), int(rank), buf, C.byref(base)))
        return bytes(buf), base.value

    def p2p_import(self, handles):
        """handles: the world's 64-byte cudaIpc handles in rank order (other processes' buffers)."""
        blob = b"".join(handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(_lib.lib().hamx_p2p_import(self._h, arr))

    def p2p_import_ptrs(self, bases):
        """Same-process variant: device base addresses of the world's buffers in rank order."""
        arr = (C.c_void_p * len(bases))(*[C.c_void_p(b) for b in bases])
        check(_lib.lib().hamx_p2p_import_ptrs(self._h, arr))

    def p2p_close(self):
        check(_lib.lib().hamx_p2p_close(self._h))

    def knn2_p2p_dev(self, d_q, nq, d_t, nt, train_offset, d_out):
        check(_lib.lib().hamx_knn2_p2p_dev(self._h, d_q, nq, d_t, nt, train_offset, d_out))

    def knn2_p2p_scatter_dev(self, d_q, nq, d_t, nt, train_offset):
        check(_lib.lib().hamx_knn2_p2p_scatter_dev(self._h, d_q, nq, d_t, nt, train_offset))

    def p2p_merge_dev(self, nq, d_out):
        check(_lib.lib().hamx_p2p_merge_dev(self._h, nq, d_out))

    def merge_top2_dev(self, d_parts, nparts, nq, d_out):
        check(_lib.lib().hamx_merge_top2_dev(self._h, d_parts, nparts, nq, d_out))

    def ratio_dev(self, d_top2, nq, ratio, d_good, d_ngood):
        check(_lib.lib().hamx_ratio_dev(self._h, d_top2, nq, float(ratio), d_good, d_ngood))

    def match_pairs_dev(self, d_pairs, npairs, max_nq, max_nt, ratio, d_good, good_stride, d_ngood):
        check(_lib.lib().hamx_match_pairs_dev(self._h, d_pairs, npairs, max_nq, max_nt, float(ratio), d_good, good_stride, d_ngood))

    def match_consecutive_dev(self, d_desc, d_counts, nframes, cap, d_prev_desc, d_prev_count, ratio, d_good, d_ngood):
        check(_lib.lib().hamx_match_consecutive_dev(self._h, d_desc, d_counts, nframes, cap, d_prev_desc or None,
                                                    d_prev_count or None, float(ratio), d_good, d_ngood))


_default_matcher = {}


def match_features(descriptors1, descriptors2, ratio=0.8, device=0):
    """matchFeatures(descriptors1, descriptors2, matches, ratio = 0.8) of src/CameraPoseEstimator.cpp:200-213:
    BFMatcher(NORM_HAMMING).knnMatch(k=2), keep raw[i][0] iff raw[i][0].distance < raw[i][1].distance * ratio."""
    m = _default_matcher.get(device)
    if m is None:
        m = _default_matcher[device] = BFMatcher(NORM_HAMMING, False, device=device)
    return m.match_ratio(descriptors1, descriptors2, ratio)


MAX_DISTANCE, CONFIDENCE = 3.0, 0.85      # src/ParamConfig.h:24-25


class FundamentalFilter:
    """The outlier filter CameraPoseEstimator runs after matchFeatures: findFundamentalMat(FM_RANSAC) for the status mask,
    then findFundamentalMat(FM_8POINT) on the inliers (computeFundamentalMatrix, src/CameraPoseEstimator.cpp:545-586)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self.device = device
        check(_lib.lib().fmx_create(C.byref(self._h), device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().fmx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(_lib.lib().fmx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        check(_lib.lib().fmx_synchronize(self._h))

    def compute_fundamental(self, keypoints1, keypoints2, matches, max_distance=MAX_DISTANCE, confidence=CONFIDENCE):
        """One pair, the reference's arguments: KEYPOINT_DTYPE arrays of both frames and a DMATCH_DTYPE list
        (queryIdx -> keypoints1, trainIdx -> keypoints2).  Returns (F 3x3 (zeros: none), status uint8[len(matches)], ninliers)."""
        k1 = np.ascontiguousarray(keypoints1, _lib.KEYPOINT_DTYPE)
        k2 = np.ascontiguousarray(keypoints2, _lib.KEYPOINT_DTYPE)
        m = np.ascontiguousarray(matches, DMATCH_DTYPE)
        status = np.zeros(max(len(m), 1), np.uint8)
        F = np.zeros(9, np.float64)
        ninl = C.c_int32(0)
        check(_lib.lib().fmx_compute_fundamental(self._h, k1.ctypes.data, len(k1), k2.ctypes.data, len(k2), m.ctypes.data, len(m),
                                                 float(max_distance), float(confidence), status.ctypes.data,
                                                 F.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ninl)))
        return F.reshape(3, 3), status[:len(m)], int(ninl.value)

    def find_batch(self, pts1, pts2, counts, max_distance=MAX_DISTANCE, confidence=CONFIDENCE):
        """npairs independent pairs: pts1 / pts2 float32 [npairs, cap, 2], counts int32 [npairs].
        Returns (status[npairs, cap] uint8, F[npairs, 3, 3], ninliers[npairs] int32)."""
        p1 = np.ascontiguousarray(pts1, np.float32)
        p2 = np.ascontiguousarray(pts2, np.float32)
        cnt = np.ascontiguousarray(counts, np.int32)
        npairs, cap = p1.shape[0], p1.shape[1]
        if p1.shape != p2.shape or p1.ndim != 3 or p1.shape[2] != 2 or cnt.shape != (npairs,):
            raise ValueError("pts1 / pts2 must be [npairs, cap, 2] and counts [npairs]")
        status = np.zeros((npairs, cap), np.uint8)
        F = np.zeros((npairs, 3, 3), np.float64)
        ninl = np.zeros(npairs, np.int32)
        check(_lib.lib().fmx_fundamental_batch(self._h, p1.ctypes.data, p2.ctypes.data, cnt.ctypes.data, npairs, cap,
                                               float(max_distance), float(confidence), status.ctypes.data, F.ctypes.data,
                                               ninl.ctypes.data))
        return status, F, ninl

    def last_info(self, npairs):
        """Per pair of the last find_batch: [inliers, RANSAC iterations run, candidate matrices scored, F produced]."""
        info = np.zeros((npairs, 4), np.int32)
        check(_lib.lib().fmx_last_info(self._h, npairs, info.ctypes.data))
        return info

    def find_batch_dev(self, d_pts1, d_pts2, d_counts, npairs, cap, max_distance, confidence, d_status, d_F, d_info):
        check(_lib.lib().fmx_fundamental_batch_dev(self._h, d_pts1, d_pts2, d_counts, npairs, cap, float(max_distance),
                                                   float(confidence), d_status, d_F, d_info))


class Triangulator:
    """The consumers of the match lists in CameraPoseEstimator, on the device (include/orbx.h "trx_*"):
    TriangulateMultiplePointsFromTwoView (src/CameraPoseEstimator.cpp:134-152), the bootstrap's four-hypothesis test
    (:334-349), the association loop (:402-455) and the new-map-point loop (:488-512)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self.device = device
        check(_lib.lib().trx_create(C.byref(self._h), device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().trx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(_lib.lib().trx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(_lib.lib().trx_synchronize(self._h))

    @staticmethod
    def _mat(a, shape):
        a = np.ascontiguousarray(a, np.float64)
        if a.shape != shape:
            raise ValueError("expected a %s matrix, got %s" % (shape, a.shape))
        return a

    def TriangulateMultiplePointsFromTwoView(self, pts1, pts2, Rt1, Rt2, K1, K2, countFront=False):
        """The reference's routine: pts n x 2 (Point2d), Rt 3x4, K 3x3.  Returns (result[n, 3] Point3d, count) with count = 0
        unless countFront, exactly like :134-152."""
        X, front, n = self.triangulate(pts1, pts2, Rt1, Rt2, K1, K2)
        return X, (n if countFront else 0)

    def triangulate(self, pts1, pts2, Rt1, Rt2, K1, K2):
        """(X[n, 3], front[n] uint8, number in front of both cameras)."""
        p1 = np.ascontiguousarray(pts1, np.float64).reshape(-1, 2)
        p2 = np.ascontiguousarray(pts2, np.float64).reshape(-1, 2)
        if p1.shape != p2.shape:
            raise ValueError("pts1 and pts2 must have the same length")
        n = len(p1)
        X = np.zeros((n, 3), np.float64)
        front = np.zeros(max(n, 1), np.uint8)
        nf = C.c_int32(0)
        check(_lib.lib().trx_triangulate(self._h, p1.ctypes.data, p2.ctypes.data, n, self._mat(Rt1, (3, 4)).ctypes.data,
                                         self._mat(Rt2, (3, 4)).ctypes.data, self._mat(K1, (3, 3)).ctypes.data,
                                         self._mat(K2, (3, 3)).ctypes.data, X.ctypes.data, front.ctypes.data, C.byref(nf)))
        return X, front[:n], int(nf.value)

    def triangulate_hypotheses(self, pts1, pts2, Rt1, Rts, K1, K2):
        """The loop at :334-349: (best index, counts[nhyp], X[n, 3] of the winner)."""
        p1 = np.ascontiguousarray(pts1, np.float64).reshape(-1, 2)
        p2 = np.ascontiguousarray(pts2, np.float64).reshape(-1, 2)
        R = np.ascontiguousarray(Rts, np.float64)
        if R.ndim != 3 or R.shape[1:] != (3, 4):
            raise ValueError("Rts must be [nhyp, 3, 4]")
        n, nh = len(p1), R.shape[0]
        X = np.zeros((max(n, 1), 3), np.float64)
        counts = np.zeros(nh, np.int32)
        best = C.c_int32(-1)
        check(_lib.lib().trx_triangulate_hypotheses(self._h, p1.ctypes.data, p2.ctypes.data, n, self._mat(Rt1, (3, 4)).ctypes.data,
                                                    R.ctypes.data, nh, self._mat(K1, (3, 3)).ctypes.data, self._mat(K2, (3, 3)).ctypes.data,
                                                    X.ctypes.data, counts.ctypes.data, C.byref(best)))
        return int(best.value), counts, X[:n]

    def triangulate_batch(self, pts1, pts2, counts, cams, select=None, want_best=False):
        """nprob problems x nhyp hypotheses: pts float32 [nprob, cap, 2], counts int32 [nprob], cams CAMERAS_DTYPE [nprob, nhyp]
        (or [nprob]), select uint8 [nprob, cap] or None.  Returns (X[nprob, nhyp, cap, 3], front[nprob, nhyp, cap], nfront[nprob, nhyp]
        [, best[nprob]])."""
        p1 = np.ascontiguousarray(pts1, np.float32)
        p2 = np.ascontiguousarray(pts2, np.float32)
        cnt = np.ascontiguousarray(counts, np.int32)
        cm = np.ascontiguousarray(cams, _lib.CAMERAS_DTYPE)
        nprob, cap = p1.shape[0], p1.shape[1]
        if cm.ndim == 1:
            cm = cm.reshape(nprob, 1)
        nh = cm.shape[1]
        if p1.shape != p2.shape or p1.ndim != 3 or p1.shape[2] != 2 or cnt.shape != (nprob,) or cm.shape[0] != nprob:
            raise ValueError("pts1 / pts2 must be [nprob, cap, 2], counts [nprob], cams [nprob(, nhyp)]")
        sel = None
        if select is not None:
            sel = np.ascontiguousarray(select, np.uint8)
            if sel.shape != (nprob, cap):
                raise ValueError("select must be [nprob, cap]")
        X = np.zeros((nprob, nh, cap, 3), np.float64)
        front = np.zeros((nprob, nh, cap), np.uint8)
        nfront = np.zeros((nprob, nh), np.int32)
        best = np.zeros(nprob, np.int32) if want_best else None
        check(_lib.lib().trx_triangulate_batch(self._h, p1.ctypes.data, p2.ctypes.data, cnt.ctypes.data, sel.ctypes.data if sel is not None else None,
                                               nprob, cap, cm.ctypes.data, nh, X.ctypes.data, front.ctypes.data, nfront.ctypes.data,
                                               best.ctypes.data if want_best else None))
        return (X, front, nfront, best) if want_best else (X, front, nfront)

    # device-resident pieces (raw device pointers; asynchronous on the handle's stream)
    def triangulate_batch_dev(self, d_pts1, d_pts2, d_counts, d_select, nprob, cap, d_cams, nhyp, d_X, d_front, d_nfront, d_best=None):
        check(_lib.lib().trx_triangulate_batch_dev(self._h, d_pts1, d_pts2, d_counts, d_select or None, nprob, cap, d_cams, nhyp, d_X, d_front,
                                                   d_nfront, d_best or None))

    def triangulate_back_dev(self, d_kps, nframes, cap, back, d_hist_kps, nhist, d_good, d_ngood, d_select, d_cams, d_X, d_front, d_nfront):
        check(_lib.lib().trx_triangulate_back_dev(self._h, d_kps, nframes, cap, back, d_hist_kps or None, nhist, d_good, d_ngood,
                                                  d_select or None, d_cams, d_X, d_front, d_nfront))

    def associate_dev(self, d_good, d_ngood, d_status, d_premap, d_ncur, nprob, back, cap, d_cur_map, d_assoc_q, d_assoc_mp, d_nassoc):
        check(_lib.lib().trx_associate_dev(self._h, d_good, d_ngood, d_status or None, d_premap, d_ncur, nprob, back, cap, d_cur_map,
                                           d_assoc_q, d_assoc_mp, d_nassoc))

    def select_new_dev(self, d_good, d_ngood, d_status, d_premap, d_cur_map, d_ncur, d_next_id, nprob, back, cap, d_accept, d_nnew):
        check(_lib.lib().trx_select_new_dev(self._h, d_good, d_ngood, d_status or None, d_premap, d_cur_map, d_ncur, d_next_id or None,
                                            nprob, back, cap, d_accept, d_nnew))


def popc_peak(device=0):
    """Measured POPC throughput of the device in Gpopc/s (the matcher's roofline denominator)."""
    g, ms = C.c_double(0), C.c_double(0)
    check(_lib.lib().hamx_popc_peak(device, C.byref(g), C.byref(ms)))
    return g.value, ms.value


# ------------------------------------------------------------------------------------------------ pipeline node
class Features:
    """struct Features of src/Frame.h:22-34."""

    def __init__(self):
        self.positions = np.zeros((0, 2), np.float64)       # vector<Point2d>
        self.descriptors = np.zeros((0, 32), np.uint8)      # Mat N x 32 CV_8U
        self.mapPointsIndices = np.zeros(0, np.int32)       # vector<int>
        self.scales = np.zeros(0, np.float64)               # vector<double>


class Frame:
    """The part of src/Frame.h:36-72 the front-end touches."""

    def __init__(self, frameBuffer):
        self.frameBuffer = frameBuffer
        self.features = Features()


class DataManager:
    """src/DataManager.h:23-36 (frames only)."""

    def __init__(self, frames=()):
        self.frames = [f if isinstance(f, Frame) else Frame(f) for f in frames]


class FeatureExtractor:
    """ProcessingNode "FeatureExtractor" (src/FeatureExtractor.{h,cpp}): process(data, frameIdx) fills
    frame.features.{descriptors, positions, scales, mapPointsIndices} exactly as lines 13-31 do."""

    name = "FeatureExtractor"

    def __init__(self, nfeatures=500, device=0, max_size=(1920, 1080), **orb_kw):
        self._args = dict(nfeatures=nfeatures, device=device, max_size=max_size, **orb_kw)
        self.detector = None
        self.extractor = None

    def init(self):
        self.detector = OrbFeatureDetector(**self._args)
        self.extractor = self.detector   # one pyramid/handle serves both, results identical to two objects

    def destroy(self):
        if self.detector is not None:
            self.detector.close()
        self.detector = self.extractor = None

    def validationCheck(self, data, frameIdx):
        return True

    def process(self, data, frameIdx):
        if self.detector is None:
            self.init()
        frame = data.frames[frameIdx]
        img = frame.frameBuffer
        if img.ndim != 2:
            # frames may be 3-channel in the reference (FrameLoader.cpp:62; cv::ORB converts internally); colour
            # conversion belongs to frame ingest, a "next" row of the scope table, and is not done on the CPU here
            raise ValueError("FeatureExtractor expects 8-bit grayscale frames")
        keypoints = self.detector.detect(img)
        keypoints, descriptor = self.extractor.compute(img, keypoints)
        frame.features.descriptors = descriptor
        frame.features.positions = np.stack([keypoints["x"], keypoints["y"]], axis=1).astype(np.float64)
        frame.features.scales = keypoints["size"].astype(np.float64)
        frame.features.mapPointsIndices = np.full(len(keypoints), -1, np.int32)
