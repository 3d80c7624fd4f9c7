#!/usr/bin/env python
"""Per-phase cycle counts of k_fm_ransac (critical path of one pair, measured by thread 0 of every CTA with clock64).
Needs a debug build of the library:  ORBX_EXTRA_FLAGS=-DFM_PROFILE sh monocular_slam_b200/csrc/build.sh
(the product build has no such counters and does not export fmx_debug_profile).
Usage: python tools/fmat_phases.py <npairs> <matches per pair> <inlier ratio>"""
import ctypes as C, sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monocular_slam_b200 import FundamentalFilter, _lib, synthetic as syn
npairs, n, inl = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
p1 = np.zeros((npairs, n, 2), np.float32); p2 = np.zeros((npairs, n, 2), np.float32)
for i in range(npairs): p1[i], p2[i] = syn.two_view_matches(100 + i, n, inl, 0.5)
fm = FundamentalFilter(); L = _lib.lib()
out = (C.c_ulonglong * 8)()
fm.find_batch(p1, p2, np.full(npairs, n, np.int32))
L.fmx_debug_profile(out, 1)
fm.find_batch(p1, p2, np.full(npairs, n, np.int32))
L.fmx_debug_profile(out, 1)
v = np.array(list(out), np.float64); names = ["draw", "collinear(+redo)", "solve", "score", "replay", "mask+8pt sums", "jacobi+rank2", "-"]
tot = v.sum()
info = fm.last_info(npairs)
print("pairs %d n %d inl %.2f: iterations mean %.1f; per-pair kcycles (thread 0) total %.1f" % (npairs, n, inl, info[:,1].mean(), tot / npairs / 1e3))
for k in range(7): print("  %-18s %8.1f kcycles/pair  %5.1f %%" % (names[k], v[k] / npairs / 1e3, 100 * v[k] / tot))
# per-pair latency: %globaltimer at the start and at the end of every pair's CTA in the last launch
t = np.zeros((npairs, 2), np.uint64)
L.fmx_debug_pair_times(t.ctypes.data_as(C.POINTER(C.c_ulonglong)), npairs)
lat = (t[:, 1] - t[:, 0]).astype(np.float64) / 1e3
start = (t[:, 0] - t[:, 0].min()).astype(np.float64) / 1e3
end = (t[:, 1] - t[:, 0].min()).astype(np.float64) / 1e3
it = info[:, 1]
print("pair latency us: p50 %.0f  p90 %.0f  p99 %.0f  max %.0f   (iterations of the slowest pair: %d; correlation with iterations %.2f)"
      % (np.percentile(lat, 50), np.percentile(lat, 90), np.percentile(lat, 99), lat.max(), it[lat.argmax()], np.corrcoef(lat, it)[0, 1]))
print("pair start us:   p50 %.0f  p99 %.0f  max %.0f   (pairs that started after the first one ended: %d)" % (np.percentile(start, 50), np.percentile(start, 99), start.max(), int((start > end.min()).sum())))
print("pair end us:     p50 %.0f  p99 %.0f  max %.0f = the launch" % (np.percentile(end, 50), np.percentile(end, 99), end.max()))
