"""GPU parity of the grey-scale JPEG decoder (csrc/jpeg.cu through the C ABI, `jpgx_*`): every committed file of
tests/golden/jpeg_cases.npz must decode to the pixels cv2.imdecode returned for it (tests/golden/make_golden_jpeg.py), larger
seeded frames to the oracle's pixels, and files of another kind must be refused."""
import os

import numpy as np
import pytest

import oracle
from monocular_slam_b200 import ORB, JpegDecoder, OrbxError, _lib
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_cases.npz"))


def test_golden_files_decode_to_cv2s_pixels():
    dec = JpegDecoder()
    by_shape = {}
    for k in G["names"]:
        by_shape.setdefault(G[k + "_pixels"].shape, []).append(k)
    assert len(by_shape) == 6
    for shape, keys in by_shape.items():
        files = [G[k + "_file"].tobytes() for k in keys]
        got = dec.decode(files)                        # one batch per image size: 5 files with different tables / restart intervals
        for i, k in enumerate(keys):
            assert np.array_equal(got[i], G[k + "_pixels"]), k
        for f, k in zip(files, keys):                  # and one at a time
            assert np.array_equal(dec.decode([f])[0], G[k + "_pixels"]), k
    dec.close()


def test_golden_colour_files_decode_to_cv2s_pixels():
    dec = JpegDecoder()
    for k in G["colour_names"]:
        f, ref = G[k + "_file"].tobytes(), G[k + "_pixels"]
        assert dec.probe(f)[4] == 3
        got = dec.decode([f, f])                           # a batch of two
        assert got.shape == (2,) + ref.shape
        assert np.array_equal(got[0], ref) and np.array_equal(got[1], ref), k
        assert np.array_equal(got[0], oracle.jpeg_decode_bgr(f))
    # grey and colour files do not mix in a call
    with pytest.raises(OrbxError) as e:
        dec.decode([G["bgr71_420q90_file"].tobytes(), G["tex333_q90_file"].tobytes()])
    assert e.value.status in (_lib.E_UNSUPPORTED, _lib.E_INVALID)
    dec.close()


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
def test_colour_frames_feed_the_extractor_as_bgr(sampling):
    """1080p-class colour files with and without restart markers: BGR frames identical to cv2's, and ORB on them (three input
    channels, gray conversion on the device) identical to ORB on cv2's frames."""
    import torch
    cv2 = pytest.importorskip("cv2")
    W, H, B = 1001, 701, 2
    frames = [syn.bgr_frame(30 + i, W, H) for i in range(B)]
    sf = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444}[sampling]
    dec = JpegDecoder()
    for rst in (0, 9):
        params = [cv2.IMWRITE_JPEG_QUALITY, 88, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
        files = [cv2.imencode(".jpg", f, params)[1].tobytes() for f in frames]
        ref = np.stack([cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED) for f in files])
        got = dec.decode(files)
        assert np.array_equal(got, ref), (sampling, rst)
    orb = ORB(nfeatures=400, max_size=(W, H), max_batch=B)
    orb.set_input_channels(3)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        orb.set_stream(stream.cuda_stream); dec.set_stream(stream.cuda_stream)
        cap = orb.default_cap
        d_frames = torch.zeros((B, H, W, 3), dtype=torch.uint8, device="cuda")
        d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda"); d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        dec.decode_dev(files, W, H, d_frames.data_ptr(), W * H * 3, W * 3, channels=3)
        orb.extract_batch_dev(d_frames.data_ptr(), W * H * 3, B, W, H, W * 3, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
        orb.check_dev()
        stream.synchronize()
    assert np.array_equal(d_frames.cpu().numpy(), ref)
    orb.set_stream(0)
    kps, desc, counts = orb.extract_batch(list(ref))
    cnt = d_cnt.cpu().numpy()
    assert np.array_equal(cnt, counts) and cnt.min() > 100
    for f in range(B):
        assert np.array_equal(d_desc[f, :cnt[f]].cpu().numpy(), desc[f, :cnt[f]])
    dec.close(); orb.close()


def test_unsupported_and_damaged_files_are_refused():
    dec = JpegDecoder()
    for k in ("refuse_progressive_file", "refuse_411_file"):
        with pytest.raises(OrbxError) as e:
            dec.decode([G[k].tobytes()])
        assert e.value.status == _lib.E_UNSUPPORTED
    with pytest.raises(OrbxError) as e:
        dec.probe(b"\\x89PNG\\r\\n\\x1a\\n" + bytes(64))
    assert e.value.status == _lib.E_INVALID
    good = G["tex333_q90_file"].tobytes()
    with pytest.raises(OrbxError):
        dec.decode([good, G["noise64_q90_file"].tobytes()])       # two sizes in one batch
    # a file cut in the middle of its scan: zero bits past the end, as libjpeg pads -- the rows before the cut are intact
    ref = G["tex333_q90_pixels"]
    cut = dec.decode([good[:len(good) // 2]])[0]
    assert np.array_equal(cut[:64], ref[:64]) and np.array_equal(cut, oracle.jpeg_decode_gray(good[:len(good) // 2]))
    dec.close()


@pytest.mark.parametrize("w,h,rst,quality", [(640, 480, 80, 90), (1920, 1080, 240, 90), (1920, 1080, 0, 75), (1001, 701, 7, 97)])
def test_larger_frames_match_the_oracle_and_feed_the_extractor(w, h, rst, quality):
    cv2 = pytest.importorskip("cv2")
    frames = [syn.frame(40 + i, w, h) if i % 2 else syn.natural_frame(40 + i, w, h) for i in range(3)]
    params = [cv2.IMWRITE_JPEG_QUALITY, quality] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
    files = [cv2.imencode(".jpg", f, params)[1].tobytes() for f in frames]
    dec = JpegDecoder()
    assert dec.probe(files[0])[:3] == (w, h, rst) and dec.probe(files[0])[4:] == (1, 0x11)
    got = dec.decode(files)
    for i, f in enumerate(files):
        assert np.array_equal(got[i], oracle.jpeg_decode_gray(f)), i
        assert np.array_equal(got[i], cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED)), i
    dec.close()


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
@pytest.mark.parametrize("w,h", [(16, 16), (32, 8), (48, 33), (640, 480), (1008, 350), (2064, 40)])
def test_colour_frames_with_16_byte_aligned_rows(sampling, w, h):
    """Widths that are multiples of 16: every row of the tight BGR output is 16-byte aligned, so the colour conversion takes its
    sixteen-pixels-per-thread form (edge columns included: the first and the last thread of a row); the larger sizes, without restart
    markers, are unstuffed by several CTAs per file.  Pixels identical to cv2.imdecode and to the oracle."""
    cv2 = pytest.importorskip("cv2")
    sf = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444}[sampling]
    rng = np.random.default_rng(w * 131 + h)
    frames = [syn.bgr_frame(70 + i, w, h) if min(w, h) >= 64 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for i in range(3)]
    frames[2] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)          # noise: many FF bytes in the entropy-coded segment
    dec = JpegDecoder()
    for rst in (0, 1000):
        params = [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
        files = [cv2.imencode(".jpg", f, params)[1].tobytes() for f in frames]
        got = dec.decode(files)
        for i, f in enumerate(files):
            assert np.array_equal(got[i], cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED)), (rst, i)
        assert np.array_equal(got[2], oracle.jpeg_decode_bgr(files[2]))
    dec.close()


@pytest.mark.parametrize("n,w,h,rst", [(1, 1920, 1080, 0), (5, 800, 600, 0), (3, 1600, 900, 5000), (2, 3840, 2160, 0), (3, 3840, 2160, 480)])
def test_long_intervals_are_unstuffed_by_several_ctas(n, w, h, rst):
    """Grey files whose restart intervals are far longer than 16 KB (none, or a marker every few thousand blocks -- intervals that
    start at any byte offset): the chunked unstuffing.  Noise frames: an FF byte every ~200 bytes of the scan.  The last case -- 4K
    frames with a marker per block row, 810 intervals of up to ~27 KB -- has too many intervals for it and takes the CTA-per-interval form."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(n * 7 + w)
    frames = [rng.integers(0, 256, (h, w), dtype=np.uint8) if i % 2 == 0 else syn.frame(80 + i, w, h) for i in range(n)]
    params = [cv2.IMWRITE_JPEG_QUALITY, 92] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
    files = [cv2.imencode(".jpg", f, params)[1].tobytes() for f in frames]
    dec = JpegDecoder()
    got = dec.decode(files)
    for i, f in enumerate(files):
        assert np.array_equal(got[i], cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED)), i
    # a file cut in the middle of its scan (zero bits past the end, as libjpeg pads) and one cut on a chunk boundary of the scan
    for cut in (len(files[0]) // 2, len(files[0]) - (len(files[0]) % 8192)):
        assert np.array_equal(dec.decode([files[0][:cut]])[0], oracle.jpeg_decode_gray(files[0][:cut])), cut
    dec.close()


def test_device_frames_go_straight_into_extraction():
    """jpgx_decode_gray_batch_dev -> orbx_extract_batch_dev on one stream: the keypoints of the decoded frames."""
    import torch
    cv2 = pytest.importorskip("cv2")
    B, W, H = 3, 640, 480
    seq = syn.sequence(B, W, H, seed=6)
    files = [cv2.imencode(".jpg", seq[i], [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_RST_INTERVAL, 80])[1].tobytes() for i in range(B)]
    decoded = np.stack([cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED) for f in files])
    orb = ORB(nfeatures=500, max_size=(W, H), max_batch=B)
    dec = JpegDecoder()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        orb.set_stream(stream.cuda_stream)
        dec.set_stream(stream.cuda_stream)
        cap = orb.default_cap
        d_frames = torch.zeros((B, H, W), dtype=torch.uint8, device="cuda")
        d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda")
        d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        for _ in range(2):
            dec.decode_dev(files, W, H, d_frames.data_ptr(), W * H, W)
            orb.extract_batch_dev(d_frames.data_ptr(), W * H, B, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
        orb.check_dev()
        stream.synchronize()
    assert np.array_equal(d_frames.cpu().numpy(), decoded)
    kps, desc, counts = orb.extract_batch(list(decoded))
    cnt = d_cnt.cpu().numpy()
    assert np.array_equal(cnt, counts)
    for f in range(B):
        assert np.array_equal(d_desc[f, :cnt[f]].cpu().numpy(), desc[f, :cnt[f]])
    dec.close(); orb.close()


def test_jpeg_ingest_equals_the_blocking_path_on_decoded_frames():
    """JpegIngest (decoder and extractor on two streams, two buffer sets): keypoints, descriptors and consecutive-frame matches of
    10 files in batches of 4 equal those of the blocking calls on the frames cv2 decodes."""
    cv2 = pytest.importorskip("cv2")
    from monocular_slam_b200 import BFMatcher, JpegIngest
    n, W, H = 10, 640, 480
    seq = syn.sequence(n, W, H, seed=11)
    files = [cv2.imencode(".jpg", seq[i], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 80])[1].tobytes() for i in range(n)]
    decoded = [cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED) for f in files]
    orb = ORB(nfeatures=400, max_size=(W, H), max_batch=4)
    m = BFMatcher()
    ing = JpegIngest(orb, m, W, H, batch=4, ratio=0.8)
    got = {}
    for first, nb, kps, desc, counts, good, ngood in ing.run(files):
        for i in range(nb):
            c = int(counts[i])
            got[first + i] = (kps[i, :c].copy(), desc[i, :c].copy(), good[i, :int(ngood[i])].copy())
    ing.close()
    assert sorted(got) == list(range(n))
    prev = None
    for f in range(n):
        k, d = orb.detectAndCompute(decoded[f])
        gk, gd, gg = got[f]
        assert np.array_equal(gk, k) and np.array_equal(gd, d), f
        if prev is not None:
            want = m.match_ratio(d, prev, 0.8)
            assert np.array_equal(gg["query_idx"], want["query_idx"]) and np.array_equal(gg["train_idx"], want["train_idx"]), f
        else:
            assert len(gg) == 0
        prev = d
    m.close(); orb.close()


def test_jpeg_ingest_colour_files():
    """JpegIngest with channels=3: colour files -> BGR frames on the device -> gray conversion + ORB; equals the blocking path on
    cv2's BGR frames."""
    cv2 = pytest.importorskip("cv2")
    from monocular_slam_b200 import BFMatcher, JpegIngest
    n, W, H = 5, 640, 480
    frames = [syn.bgr_frame(50 + i, W, H) for i in range(n)]
    files = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for f in frames]          # 4:2:0, no restart markers
    decoded = [cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED) for f in files]
    orb = ORB(nfeatures=300, max_size=(W, H), max_batch=2)
    m = BFMatcher()
    ing = JpegIngest(orb, m, W, H, batch=2, ratio=0.8, channels=3)
    got = {}
    for first, nb, kps, desc, counts, good, ngood in ing.run(files):
        for i in range(nb):
            got[first + i] = (kps[i, :int(counts[i])].copy(), desc[i, :int(counts[i])].copy())
    ing.close()
    for f in range(n):
        k, d = orb.detectAndCompute(decoded[f])
        assert len(k) > 50 and np.array_equal(got[f][0], k) and np.array_equal(got[f][1], d), f
    m.close(); orb.close()
