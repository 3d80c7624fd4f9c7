"""The C++ host side (include/orbx_shim.hpp): compiles against the C ABI everywhere; on a GPU it must reproduce the
oracle when driven like the reference's pipeline (FeatureExtractor::process per frame, matchFeatures per pair)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from monocular_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_demo(tmpdir, name="shim_demo"):
    _lib.build()
    exe = os.path.join(str(tmpdir), name)
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L", libdir, "-lorbx",
                           "-Wl,-rpath," + libdir])
    return exe


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = _build_demo(tmp_path)
    if _lib.lib().orbx_device_count() > 0:
        pytest.skip("a GPU is present; covered by the gpu test")
    raw = tmp_path / "f.raw"
    raw.write_bytes(bytes(64 * 64))
    r = subprocess.run([exe, "64", "64", "1", str(raw), str(tmp_path / "o.bin"), "500", "0.8"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_pipeline_demo_compiles(tmp_path):
    _build_demo(tmp_path, "pipeline_demo")


@pytest.mark.gpu
@pytest.mark.parametrize("demo,n,extra", [("shim_demo", 3, []), ("pipeline_demo", 9, ["2"])])
def test_shim_matches_oracle(tmp_path, demo, n, extra):
    """shim_demo: the reference's call-by-call surface; pipeline_demo: SequenceFrontEnd (batches of 2 through the pipelined
    C ABI, two process() calls).  Both must reproduce the oracle frame by frame and match by match."""
    import oracle
    from monocular_slam_b200 import synthetic as syn
    exe = _build_demo(tmp_path, demo)
    w, h, nf, ratio = 800, 600, 700, 0.8
    seq = syn.sequence(n, w, h, seed=33)
    raw = tmp_path / "frames.raw"
    raw.write_bytes(seq.tobytes())
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, str(w), str(h), str(n), str(raw), str(out), str(nf), str(ratio)] + extra)
    buf = out.read_bytes()
    off = 0
    P = oracle.Params(nfeatures=nf)
    prev = None
    for i in range(n):
        ok, od = oracle.detect_and_compute(seq[i], P)
        (cnt,) = struct.unpack_from("<i", buf, off); off += 4
        assert cnt == len(ok)
        pos = np.frombuffer(buf, "<f8", cnt * 2, off).reshape(cnt, 2); off += cnt * 16
        scales = np.frombuffer(buf, "<f8", cnt, off); off += cnt * 8
        mpi = np.frombuffer(buf, "<i4", cnt, off); off += cnt * 4
        desc = np.frombuffer(buf, np.uint8, cnt * 32, off).reshape(cnt, 32); off += cnt * 32
        assert np.array_equal(pos, np.stack([ok["x"], ok["y"]], 1).astype(np.float64))
        assert np.array_equal(scales, ok["size"].astype(np.float64)) and (mpi == -1).all()
        assert np.array_equal(desc, od)
        if i > 0:
            (m,) = struct.unpack_from("<i", buf, off); off += 4
            dm = np.frombuffer(buf, _lib.DMATCH_DTYPE, m, off); off += m * 16
            gq, gt, gd = oracle.match_features(od, prev, ratio)
            assert m == len(gq) and np.array_equal(dm["query_idx"], gq) and np.array_equal(dm["train_idx"], gt)
            assert np.array_equal(dm["distance"].astype(np.int32), gd)
        prev = od
    assert off == len(buf)
