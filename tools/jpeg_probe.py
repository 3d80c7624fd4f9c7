#!/usr/bin/env python
"""Timing of the GPU JPEG decoder: 64 grey 1080p frames as JPEG files in memory (quality 90), with a restart marker per block row
and without any; wall time of the call (host parsing + staging), device time of its kernels, and cv2.imdecode on one core.
Usage: python tools/jpeg_probe.py [nframes]"""
import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monocular_slam_b200 import JpegDecoder, _lib  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def rounds():
    """Rounds of k_jpeg_sync since the last call, if the library was built with -DJP_PROFILE (ORBX_EXTRA_FLAGS)."""
    import ctypes as C
    L = _lib.lib()
    if not hasattr(L, "jpgx_debug_rounds"):
        return ""
    out = (C.c_uint * 3)()
    L.jpgx_debug_rounds(out)
    return "; sync rounds per interval: mean %.1f, max %d" % (out[1] / max(out[0], 1), out[2]) if out[0] else ""


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    W, H = 1920, 1080
    base = [syn.natural_frame(i, W, H) for i in range(8)]
    frames = [base[i % 8] for i in range(B)]
    dev = torch.device("cuda:0")
    dec = JpegDecoder()
    stream = torch.cuda.Stream()
    dec.set_stream(stream.cuda_stream)
    d_frames = torch.zeros((B, H, W), dtype=torch.uint8, device=dev)
    for name, rst in (("restart marker every block row", (W + 7) // 8), ("restart marker every 16 blocks", 16), ("no restart markers", 0)):
        params = [cv2.IMWRITE_JPEG_QUALITY, 90] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
        files = [cv2.imencode(".jpg", f, params)[1].tobytes() for f in frames]
        nbytes = sum(len(f) for f in files)
        t0 = time.perf_counter()
        ref = [cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED) for f in files[:8]]
        cpu = 8 / (time.perf_counter() - t0)
        for _ in range(2):
            dec.decode_dev(files, W, H, d_frames.data_ptr(), W * H, W)
        stream.synchronize()
        assert all(np.array_equal(d_frames[i].cpu().numpy(), ref[i]) for i in range(8))
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = 0.0
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(reps):
            t1 = time.perf_counter()
            dec.decode_dev(files, W, H, d_frames.data_ptr(), W * H, W)
            host += time.perf_counter() - t1
        e1.record(stream)
        stream.synchronize()
        wall = time.perf_counter() - t0
        print(rounds()[2:])
        print("%-32s %5.0f KB per file: call returns after %.2f ms (host parse + staging), batch done after %.2f ms wall = %.0f frames/s "
              "(device span %.2f ms); cv2.imdecode on one core: %.0f frames/s" % (name, nbytes / B / 1e3, host / reps * 1e3, wall / reps * 1e3, B * reps / wall,
                                                                                  e0.elapsed_time(e1) / reps, cpu))
    # colour frames (4:2:0, the libjpeg default): BGR out, what the extractor reads with three input channels
    cframes = [syn.bgr_frame(i, W, H) for i in range(8)]
    d_bgr = torch.zeros((B, H, W, 3), dtype=torch.uint8, device=dev)
    for name, rst in (("colour 4:2:0, a marker per MCU row", (W + 15) // 16), ("colour 4:2:0, no restart markers", 0)):
        params = [cv2.IMWRITE_JPEG_QUALITY, 90] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
        files = [cv2.imencode(".jpg", cframes[i % 8], params)[1].tobytes() for i in range(B)]
        t0 = time.perf_counter()
        ref = [cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED) for f in files[:8]]
        cpu = 8 / (time.perf_counter() - t0)
        for _ in range(2):
            dec.decode_dev(files, W, H, d_bgr.data_ptr(), W * H * 3, W * 3, channels=3)
        stream.synchronize()
        assert all(np.array_equal(d_bgr[i].cpu().numpy(), ref[i]) for i in range(8))
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            dec.decode_dev(files, W, H, d_bgr.data_ptr(), W * H * 3, W * 3, channels=3)
        stream.synchronize()
        wall = time.perf_counter() - t0
        print(rounds()[2:])
        print("%-36s %5.0f KB per file: %.2f ms per batch = %.0f frames/s; cv2.imdecode on one core: %.0f frames/s"
              % (name, sum(len(f) for f in files) / B / 1e3, wall / reps * 1e3, B * reps / wall, cpu))
    dec.close()


if __name__ == "__main__":
    main()
