"""The C++ host side (include/orbx_shim.hpp): compiles against the C ABI everywhere; on a GPU it must reproduce the
oracle when driven like the reference's pipeline (FeatureExtractor::process per frame, matchFeatures per pair)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from monocular_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_demo(tmpdir, name="shim_demo", extra_includes=()):
    _lib.build()
    exe = os.path.join(str(tmpdir), name)
    libdir = os.path.dirname(_lib.LIB_PATH)
    inc = []
    for d in (os.path.join(ROOT, "include"),) + tuple(extra_includes):
        inc += ["-I", d]
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-Werror"] + inc +
                          [os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L", libdir, "-lorbx", "-Wl,-rpath," + libdir])
    return exe


STUB_OPENCV = os.path.join(ROOT, "tests", "cpp", "stub_opencv")


def test_opencv_branch_of_the_shim_compiles(tmp_path):
    """The ORBX_SHIM_USE_OPENCV form of the shim (what a maintainer of the reference builds) against a stub of the OpenCV 2.4
    headers; without a GPU the program must fail loudly instead of falling back."""
    exe = _build_demo(tmp_path, "shim_opencv_demo", (STUB_OPENCV,))
    if _lib.lib().orbx_device_count() > 0:
        pytest.skip("a GPU is present; covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_opencv_branch_of_the_shim_runs(tmp_path):
    exe = _build_demo(tmp_path, "shim_opencv_demo", (STUB_OPENCV,))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok:"), r.stdout + r.stderr


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


def test_fmat_demo_compiles(tmp_path):
    _build_demo(tmp_path, "fmat_demo")


@pytest.mark.gpu
def test_shim_compute_fundamental_matrix_matches_oracle(tmp_path):
    """computeFundamentalMatrix through the C++ shim: status, aligned inlier positions and F as the oracle computes them."""
    import oracle
    from monocular_slam_b200 import synthetic as syn
    exe = _build_demo(tmp_path, "fmat_demo")
    n = 700
    p1, p2 = syn.two_view_matches(21, n, 0.65, 0.5)
    pts = tmp_path / "pts.bin"
    pts.write_bytes(np.c_[p1, p2].astype(np.float64).tobytes())
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, str(n), str(pts), str(out)])
    buf = out.read_bytes()
    (ni,) = struct.unpack_from("<i", buf, 0)
    status = np.frombuffer(buf, np.uint8, n, 4)
    in1 = np.frombuffer(buf, "<f8", ni * 2, 4 + n).reshape(ni, 2)
    in2 = np.frombuffer(buf, "<f8", ni * 2, 4 + n + ni * 16).reshape(ni, 2)
    F = np.frombuffer(buf, "<f8", 9, 4 + n + ni * 32).reshape(3, 3)
    _, mo, _ = oracle.fm_ransac(p1, p2, 3.0, 0.85)
    assert np.array_equal(status, mo) and ni == mo.sum()
    assert np.array_equal(in1, p1[mo > 0].astype(np.float64)) and np.array_equal(in2, p2[mo > 0].astype(np.float64))
    Fo = oracle.fm_8point(p1[mo > 0], p2[mo > 0])
    assert np.linalg.norm(F - Fo) / np.linalg.norm(Fo) <= 1e-6


@pytest.mark.gpu
def test_sequence_front_end_with_filter(tmp_path):
    """SequenceFrontEnd::enableFilter: per-frame status / F out of the pipelined C++ path equal fmx_compute_fundamental on the
    keypoints and matches the same run reports (the filter's own parity with the oracle is covered in test_gpu_fmat.py)."""
    from monocular_slam_b200 import FundamentalFilter, KEYPOINT_DTYPE
    from monocular_slam_b200 import synthetic as syn
    exe = _build_demo(tmp_path, "pipeline_demo")
    w, h, nf, ratio, n = 800, 600, 700, 0.8, 7
    seq = syn.sequence(n, w, h, seed=35)
    raw = tmp_path / "frames.raw"
    raw.write_bytes(seq.tobytes())
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, str(w), str(h), str(n), str(raw), str(out), str(nf), str(ratio), "3", "filter"])
    buf = out.read_bytes()
    off = 0
    fm = FundamentalFilter()
    prev = None
    for i in range(n):
        (cnt,) = struct.unpack_from("<i", buf, off); off += 4
        pos = np.frombuffer(buf, "<f8", cnt * 2, off).reshape(cnt, 2); off += cnt * 16
        off += cnt * 8 + cnt * 4 + cnt * 32
        kps = np.zeros(cnt, KEYPOINT_DTYPE)
        kps["x"], kps["y"] = pos[:, 0], pos[:, 1]
        if i > 0:
            (m,) = struct.unpack_from("<i", buf, off); off += 4
            dm = np.frombuffer(buf, _lib.DMATCH_DTYPE, m, off); off += m * 16
            status = np.frombuffer(buf, np.uint8, m, off); off += m
            F = np.frombuffer(buf, "<f8", 9, off).reshape(3, 3); off += 72
            Fh, sh, nh = fm.compute_fundamental(kps, prev, dm)
            assert m > 100 and np.array_equal(status, sh) and np.array_equal(F, Fh)
        prev = kps
    assert off == len(buf)
    fm.close()


def test_bow_demo_compiles(tmp_path):
    _build_demo(tmp_path, "bow_demo")


@pytest.mark.gpu
def test_shim_vocabulary_matches_oracle(tmp_path):
    """DBoW2::OrbVocabulary of the C++ shim: a text vocabulary loaded from disk, BowVector / FeatureVector per frame and scores
    as the oracle computes them, bit for bit."""
    import oracle
    from monocular_slam_b200 import synthetic as syn
    exe = _build_demo(tmp_path, "bow_demo")
    va = syn.vocabulary(61, k=9, L=3)
    vpath = tmp_path / "voc.txt"
    syn.write_vocabulary_text(str(vpath), va, trailing_newline=True)       # as saveToTextFile ends its files
    nframes, n, lu = 3, 400, 2
    desc = np.stack([syn.vocabulary_features(70 + f, va, n, pool=150) for f in range(nframes)])
    dpath = tmp_path / "desc.bin"
    dpath.write_bytes(desc.tobytes())
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, str(vpath), str(dpath), str(nframes), str(n), str(lu), str(out)])
    buf = out.read_bytes()
    ov = oracle.BowVocabulary(va)
    pos, bows = 0, []
    for f in range(nframes):
        w, v, nodes, offs, feats = ov.transform(desc[f], lu)
        (nb,) = struct.unpack_from("<i", buf, pos); pos += 4
        rec = np.frombuffer(buf, np.dtype([("w", "<u4"), ("v", "<f8")]), nb, pos); pos += nb * 12
        assert np.array_equal(rec["w"], w) and np.array_equal(rec["v"], v)
        (nf,) = struct.unpack_from("<i", buf, pos); pos += 4
        assert nf == len(nodes)
        for g in range(nf):
            node, c = struct.unpack_from("<Ii", buf, pos); pos += 8
            fe = np.frombuffer(buf, "<u4", c, pos); pos += 4 * c
            assert node == nodes[g] and np.array_equal(fe, feats[offs[g]:offs[g + 1]])
        bows.append((w, v))
    s = np.frombuffer(buf, "<f8", 2 * nframes, pos)
    expect = np.array([ov.score(bows[0], b) for b in bows])
    assert np.array_equal(s[:nframes], expect) and np.array_equal(s[nframes:], expect)


def test_jpeg_demo_compiles(tmp_path):
    _build_demo(tmp_path, "jpeg_demo")


@pytest.mark.gpu
def test_shim_jpeg_decoder_matches_golden(tmp_path):
    """orbx_shim::JpegDecoder on committed files read from disk: the pixels cv2.imdecode returned (grey and colour)."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "jpeg_cases.npz"))
    exe = _build_demo(tmp_path, "jpeg_demo")
    for keys in (["tex333_q90", "tex333_q50rst8", "tex333_q75opt"], ["bgr71_420q90", "bgr71_420q30rst3"], ["bgr71_444q100opt"]):
        paths = []
        for k in keys:
            p = tmp_path / (k + ".jpg")
            p.write_bytes(G[k + "_file"].tobytes())
            paths.append(str(p))
        out = tmp_path / "frames.bin"
        subprocess.check_call([exe, str(out)] + paths)
        buf = out.read_bytes()
        rows, cols, ch = struct.unpack_from("<iii", buf, 0)
        frames = np.frombuffer(buf, np.uint8, len(keys) * rows * cols * ch, 12).reshape((len(keys), rows, cols) + ((ch,) if ch > 1 else ()))
        for i, k in enumerate(keys):
            assert np.array_equal(frames[i], G[k + "_pixels"]), k
    prog = tmp_path / "prog.jpg"
    prog.write_bytes(G["refuse_progressive_file"].tobytes())
    assert subprocess.call([exe, str(tmp_path / "x.bin"), str(prog)], stderr=subprocess.DEVNULL) == 4
