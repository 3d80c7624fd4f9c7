"""oracle/jpeg_oracle.c against tests/golden/jpeg_cases.npz: what cv2.imdecode (libjpeg-turbo, JDCT_ISLOW) returns for the
committed grey-scale files (tests/golden/make_golden_jpeg.py), pixel for pixel."""
import os

import numpy as np
import pytest

import oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_cases.npz"))


def test_every_golden_file_decodes_to_cv2s_pixels():
    assert len(G["names"]) == 30
    seen_rst = 0
    for k in G["names"]:
        data, ref = G[k + "_file"].tobytes(), G[k + "_pixels"]
        w, h, rst, blocks = oracle.jpeg_probe(data)
        assert (h, w) == ref.shape and blocks == ((w + 7) // 8) * ((h + 7) // 8)
        seen_rst += rst > 0
        assert np.array_equal(oracle.jpeg_decode_gray(data), ref), k
    assert seen_rst == 18          # three of the five settings carry restart markers


def test_colour_files_decode_to_cv2s_pixels():
    assert len(G["colour_names"]) == 20
    for k in G["colour_names"]:
        assert np.array_equal(oracle.jpeg_decode_bgr(G[k + "_file"].tobytes()), G[k + "_pixels"]), k


def test_unsupported_files_are_refused_not_guessed():
    for k in ("refuse_progressive_file", "refuse_411_file"):
        for fn in (oracle.jpeg_probe, oracle.jpeg_decode_bgr):
            with pytest.raises(ValueError) as e:
                fn(G[k].tobytes())
            assert e.value.args[0] == oracle.JPEG_UNSUPPORTED
    with pytest.raises(ValueError) as e:
        oracle.jpeg_probe(G["bgr71_420q90_file"].tobytes())          # the grey-scale entry point refuses a colour file
    assert e.value.args[0] == oracle.JPEG_UNSUPPORTED
    with pytest.raises(ValueError) as e:
        oracle.jpeg_probe(b"\x89PNG\r\n\x1a\n" + bytes(64))
    assert e.value.args[0] == oracle.JPEG_CORRUPT


def test_truncated_file_is_zero_padded_like_libjpeg():
    # a file cut in the middle of the scan still decodes (libjpeg pads the bit stream with zeros): the rows before the cut are intact
    k = "tex333_q90"
    data, ref = G[k + "_file"].tobytes(), G[k + "_pixels"]
    got = oracle.jpeg_decode_gray(data[:len(data) // 2])
    assert got.shape == ref.shape and np.array_equal(got[:64], ref[:64])
