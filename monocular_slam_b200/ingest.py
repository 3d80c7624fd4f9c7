"""Frame ingest in front of the extractor: encoded images -> host decode workers -> pinned slots -> orbx_submit_batch.

The reference reads every frame with ``imread(path, CV_LOAD_IMAGE_UNCHANGED)`` on the host before anything else happens
(src/FrameLoader.cpp:36-67, the call at :62).  PNG / JPEG entropy decoding is a serial bit stream per image and stays on the
host here too (OpenCV's decoder, the one the reference calls); what this module adds is the plumbing that keeps the GPU fed:
``workers`` threads decode straight into page-locked slots (cv2 releases the GIL while decoding), a full slot goes to the
pipelined ``ORB.submit_batch`` (upload, extraction and matching overlap the decoding of the next slot), and gray conversion
of colour frames happens on the device (orbx_set_input_channels), as cv::ORB does it on the CPU.  ``IngestRing.stats`` says
how many frames per second one decode worker sustains, which is the number that tells how many host cores a given GPU rate
needs (bench.py ``ingest``).
"""
import ctypes as C
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from ._lib import DMATCH_DTYPE, KEYPOINT_DTYPE, check


def _pinned(shape, dtype):
    """A numpy array over page-locked memory from orbx_host_alloc (freed by IngestRing.close)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(_lib.lib().orbx_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape), p


def imread_unchanged(source):
    """The reference's loader call: a path (imread) or encoded bytes (imdecode), flag IMREAD_UNCHANGED."""
    import cv2
    if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
        img = cv2.imdecode(np.frombuffer(source, np.uint8) if not isinstance(source, np.ndarray) else source, cv2.IMREAD_UNCHANGED)
    else:
        img = cv2.imread(str(source), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise ValueError("cannot decode %r" % (source if not isinstance(source, (bytes, bytearray, memoryview, np.ndarray)) else "<bytes>"))
    return img


class JpegDecoder:
    """Baseline JPEG files decoded on the GPU (include/orbx.h "jpgx_*"): the pixels cv2.imdecode / the reference's imread would
    return, for a whole batch at once -- grey files as [h][w], YCbCr files (4:2:0, 4:2:2, 4:4:4) as [h][w][3] BGR.
    ``decode(files)`` returns host frames; ``decode_dev(files, w, h, d_frames, frame_pitch, stride, channels)`` writes device
    frames (asynchronous on the handle's stream) for ORB.extract_batch_dev.  A file of another kind raises OrbxError with status
    E_UNSUPPORTED -- decode that one with ``imread_unchanged``."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self.device = device
        check(_lib.lib().jpgx_create(C.byref(self._h), device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().jpgx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(_lib.lib().jpgx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(_lib.lib().jpgx_synchronize(self._h))

    @staticmethod
    def probe(data):
        """(width, height, restart interval in MCUs, blocks, components, luma sampling h*16+v) from the headers; raises OrbxError
        for files this decoder refuses."""
        buf = (C.c_uint8 * len(data)).from_buffer_copy(bytes(data))
        info = (C.c_int32 * 6)()
        check(_lib.lib().jpgx_probe(buf, len(data), info))
        return tuple(info)

    @staticmethod
    def _args(files):
        keep = [np.frombuffer(bytes(f), np.uint8) if not isinstance(f, np.ndarray) else np.ascontiguousarray(f, np.uint8).reshape(-1) for f in files]
        ptrs = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
        sizes = (C.c_size_t * len(keep))(*[k.size for k in keep])
        return keep, ptrs, sizes

    def decode_dev(self, files, w, h, d_frames_ptr, frame_pitch, stride, channels=1):
        keep, ptrs, sizes = self._args(files)
        fn = _lib.lib().jpgx_decode_gray_batch_dev if channels == 1 else _lib.lib().jpgx_decode_bgr_batch_dev
        check(fn(self._h, ptrs, sizes, len(keep), int(w), int(h), d_frames_ptr, int(frame_pitch), int(stride)))

    def decode(self, files, out=None):
        files = list(files)
        if not files:
            return np.zeros((0, 0, 0), np.uint8)
        w, h, _, _, ncomp, _ = self.probe(files[0])
        if out is None:
            out = np.zeros((len(files), h, w) if ncomp == 1 else (len(files), h, w, 3), np.uint8)
        keep, ptrs, sizes = self._args(files)
        fn = _lib.lib().jpgx_decode_gray_batch if ncomp == 1 else _lib.lib().jpgx_decode_bgr_batch
        check(fn(self._h, ptrs, sizes, len(keep), w, h, out.ctypes.data_as(C.c_void_p), h * w * ncomp, w * ncomp))
        return out


class IngestRing:
    """Decode -> pinned slot -> ``orb.submit_batch``.  ``channels`` = 1 (gray files) or 3 (BGR files, converted on the device).

    ``run(sources)`` yields, per batch and in order, ``(first_index, n, kps, desc, counts, good, ngood)`` -- the arrays are
    the ring's own pinned buffers and are overwritten ``pipeline_depth`` batches later, so consume or copy them."""

    def __init__(self, orb, matcher, width, height, batch=None, workers=4, ratio=0.8, channels=1, decode=imread_unchanged):
        self.orb, self.matcher, self.ratio = orb, matcher, float(ratio)
        self.w, self.h, self.ch = int(width), int(height), int(channels)
        self.batch = int(batch or orb.max_batch)
        self.workers = int(workers)
        self.decode = decode
        self.depth = orb.pipeline_depth()
        cap = self.cap = orb.default_cap
        self._ptrs = []
        shape = (self.batch, self.h, self.w) if self.ch == 1 else (self.batch, self.h, self.w, self.ch)
        self.slots, self.outs = [], []
        for _ in range(self.depth + 1):        # one slot is being filled while `depth` batches are in flight
            a, p = _pinned(shape, np.uint8)
            self._ptrs.append(p)
            self.slots.append(a)
            o = []
            for shp, dt in (((self.batch, cap), KEYPOINT_DTYPE), ((self.batch, cap, 32), np.uint8), ((self.batch, cap), DMATCH_DTYPE)):
                a, p = _pinned(shp, dt)
                self._ptrs.append(p)
                o.append(a)
            self.outs.append((o[0], o[1], np.zeros(self.batch, np.int32), o[2], np.zeros(self.batch, np.int64)))
        self.pool = ThreadPoolExecutor(self.workers)
        self.stats = {"frames": 0, "decode_seconds": 0.0, "wall_seconds": 0.0, "workers": self.workers}

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None
        while self.orb.batches_in_flight():
            self.orb.wait_batch()
        for p in self._ptrs:
            _lib.lib().orbx_host_free(p)
        self._ptrs = []

    def _decode_into(self, source, dst):
        t0 = time.perf_counter()
        img = self.decode(source)
        if img.shape[:2] != (self.h, self.w) or (img.ndim == 3) != (self.ch == 3) or img.dtype != np.uint8:
            raise ValueError("decoded frame is %s %s, the ring was built for %dx%d with %d channel(s)" % (img.dtype, img.shape, self.w, self.h, self.ch))
        np.copyto(dst, img)
        return time.perf_counter() - t0

    def run(self, sources):
        sources = list(sources)
        orb = self.orb
        orb.reset_sequence()
        pending = []                                  # (first index, n, slot index) of the batches in flight
        t_start = time.perf_counter()
        nslots = len(self.slots)

        def collect():
            first, n, si = pending.pop(0)
            kps, desc, counts, good, ngood = orb.wait_batch()
            return first, n, kps, desc, counts, good, ngood

        for b, first in enumerate(range(0, len(sources), self.batch)):
            n = min(self.batch, len(sources) - first)
            si = b % nslots
            slot = self.slots[si]
            futs = [self.pool.submit(self._decode_into, sources[first + i], slot[i]) for i in range(n)]
            # the oldest batch is collected while the workers decode (its slot is the one that will be filled next)
            if len(pending) == self.depth:
                yield collect()
            self.stats["decode_seconds"] += sum(f.result() for f in futs)
            orb.submit_batch([slot[i] for i in range(n)], self.matcher, self.ratio, self.outs[si])
            pending.append((first, n, si))
            self.stats["frames"] += n
        while pending:
            yield collect()
        self.stats["wall_seconds"] += time.perf_counter() - t_start


class JpegIngest:
    """JPEG files (grey, or colour with ``channels=3``) -> features, with the entropy decoder on the GPU: compressed bytes over PCIe -> ``JpegDecoder`` on its
    own stream -> ``ORB.extract_batch_dev`` + consecutive-frame matching on the extractor's stream -> results in pinned host
    memory.  Two frame buffers: batch k+1 is decoded while batch k is extracted.  torch supplies the device buffers, streams and
    events (plumbing only).  A batch holding a file the decoder refuses (colour, progressive ...) raises OrbxError with status
    E_UNSUPPORTED: feed such files to ``IngestRing`` instead.

    ``run(files)`` yields, per batch and in order, ``(first_index, n, kps, desc, counts, good, ngood)`` -- pinned buffers that are
    overwritten two batches later.  While it runs the extractor and the matcher work on this object's stream; afterwards they are
    back on their own."""

    def __init__(self, orb, matcher, width, height, batch=None, ratio=0.8, channels=1):
        import torch
        self.torch = torch
        self.orb, self.matcher, self.ratio = orb, matcher, float(ratio)
        self.w, self.h, self.ch = int(width), int(height), int(channels)
        orb.set_input_channels(self.ch)
        self.batch = int(batch or orb.max_batch)
        dev = torch.device("cuda", orb.device)
        self.dev = dev
        self.dec = JpegDecoder(device=orb.device)
        self.s_dec, self.s_orb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.dec.set_stream(self.s_dec.cuda_stream)
        cap = self.cap = orb.default_cap
        B = self.batch
        self.frames = [torch.zeros((B, self.h, self.w * self.ch), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.d = [{"kps": torch.empty((B, cap, 7), dtype=torch.float32, device=dev), "desc": torch.empty((B, cap, 32), dtype=torch.uint8, device=dev),
                   "cnt": torch.zeros(B, dtype=torch.int32, device=dev), "good": torch.empty((B, cap, 4), dtype=torch.int32, device=dev),
                   "ngood": torch.zeros(B, dtype=torch.int64, device=dev)} for _ in range(2)]
        self.hst = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in d.items()} for d in self.d]
        self.prev_desc = torch.zeros((cap, 32), dtype=torch.uint8, device=dev)
        self.prev_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ev_dec = [torch.cuda.Event() for _ in range(2)]
        self.ev_ext = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]

    def close(self):
        self.torch.cuda.synchronize(self.dev)
        self.dec.close()

    def run(self, files):
        torch = self.torch
        files = list(files)
        orb, m, B, cap = self.orb, self.matcher, self.batch, self.cap
        orb.set_stream(self.s_orb.cuda_stream)
        m.set_stream(self.s_orb.cuda_stream)
        have_prev = False
        pending = []

        def collect():
            first, n, b = pending.pop(0)
            self.ev_done[b].synchronize()
            h = self.hst[b]
            return (first, n, h["kps"].numpy().view(KEYPOINT_DTYPE).reshape(B, cap), h["desc"].numpy(), h["cnt"].numpy(),
                    h["good"].numpy().view(DMATCH_DTYPE).reshape(B, cap), h["ngood"].numpy())
        try:
            for k, first in enumerate(range(0, len(files), B)):
                n = min(B, len(files) - first)
                b = k & 1
                if len(pending) == 2:
                    yield collect()                          # frees buffer set b
                self.s_dec.wait_event(self.ev_ext[b])
                self.dec.decode_dev(files[first:first + n], self.w, self.h, self.frames[b].data_ptr(), self.w * self.h * self.ch, self.w * self.ch, self.ch)
                self.ev_dec[b].record(self.s_dec)
                d, h = self.d[b], self.hst[b]
                with torch.cuda.stream(self.s_orb):
                    self.s_orb.wait_event(self.ev_dec[b])
                    orb.extract_batch_dev(self.frames[b].data_ptr(), self.w * self.h * self.ch, n, self.w, self.h, self.w * self.ch, d["kps"].data_ptr(), d["desc"].data_ptr(),
                                          cap, d["cnt"].data_ptr())
                    self.ev_ext[b].record(self.s_orb)
                    m.match_consecutive_dev(d["desc"].data_ptr(), d["cnt"].data_ptr(), n, cap, self.prev_desc.data_ptr() if have_prev else 0,
                                            self.prev_cnt.data_ptr() if have_prev else 0, self.ratio, d["good"].data_ptr(), d["ngood"].data_ptr())
                    self.prev_desc.copy_(d["desc"][n - 1], non_blocking=True)
                    self.prev_cnt.copy_(d["cnt"][n - 1:n], non_blocking=True)
                    have_prev = True
                    for key in d:
                        h[key].copy_(d[key], non_blocking=True)
                    self.ev_done[b].record(self.s_orb)
                pending.append((first, n, b))
            while pending:
                yield collect()
            orb.check_dev()
        finally:
            torch.cuda.synchronize(self.dev)
            orb.set_stream(0)                                # both handles are left on their own streams
            m.set_stream(0)
