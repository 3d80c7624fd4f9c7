"""Host-side model of the chunked byte unstuffing of csrc/jpeg.cu (k_jpeg_unstuff_count + k_jpeg_unstuff_chunk): an interval is cut
into chunks of JP_CHUNK source bytes; a byte is dropped iff it is 00 and its predecessor INSIDE the interval is FF (T.81 B.1.1.5), which
depends on the predecessor alone, so every chunk counts and compacts on its own and lands at the sum of the counts in front of it.  The
model restates that arithmetic in numpy (four bytes per thread, the predecessor read across thread and chunk boundaries) and is checked
against a sequential unstuffing on the byte patterns that straddle the boundaries."""
import numpy as np
import pytest

JP_CHUNK = 8192


def sequential(src):
    out, prev_ff = bytearray(), False
    for b in src:
        if not (b == 0 and prev_ff):
            out.append(b)
        prev_ff = b == 0xFF          # (a dropped 00 is not FF, so FF 00 00 keeps the second 00)
    return bytes(out)


def keep4(src, i):
    """jp_keep4: keep mask of the four bytes at offset i (i % 4 == 0) of the interval."""
    n = len(src)
    before = src[i - 1] if i else 0
    keep = []
    for k in range(4):
        if i + k >= n:
            keep.append(False)
            continue
        keep.append(not (src[i + k] == 0 and before == 0xFF))
        before = src[i + k]
    return keep


def chunked(src):
    n = len(src)
    nchunks = max(1, -(-n // JP_CHUNK))
    kept = []                                            # pass 1: k_jpeg_unstuff_count
    for c in range(nchunks):
        kept.append(sum(sum(keep4(src, i)) for i in range(c * JP_CHUNK, min(n, (c + 1) * JP_CHUNK), 4)))
    out = bytearray(sum(kept))
    for c in range(nchunks):                             # pass 2: k_jpeg_unstuff_chunk, every chunk on its own
        at = sum(kept[:c])
        for i in range(c * JP_CHUNK, min(n, (c + 1) * JP_CHUNK), 4):
            for k, keep in enumerate(keep4(src, i)):
                if keep:
                    out[at] = src[i + k]
                    at += 1
        assert at == sum(kept[:c + 1])
    return bytes(out)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, JP_CHUNK - 1, JP_CHUNK, JP_CHUNK + 1, 2 * JP_CHUNK, 3 * JP_CHUNK + 7])
def test_random_streams_with_many_ff_bytes(n):
    r = np.random.default_rng(n)
    src = r.choice(np.array([0x00, 0xFF, 0x12, 0xFE], np.uint8), size=n, p=[0.35, 0.35, 0.2, 0.1]).tobytes()
    assert chunked(src) == sequential(src)


@pytest.mark.parametrize("at", [JP_CHUNK - 2, JP_CHUNK - 1, JP_CHUNK, 4 * 511 + 3, 4 * 512])
@pytest.mark.parametrize("pattern", [b"\xff\x00", b"\xff\x00\x00", b"\xff\xff\x00", b"\xff\x00\xff\x00", b"\x00\xff", b"\xff"])
def test_patterns_across_thread_and_chunk_boundaries(at, pattern):
    src = bytearray(b"\x55" * (2 * JP_CHUNK + 5))
    src[at:at + len(pattern)] = pattern
    src = bytes(src)
    assert chunked(src) == sequential(src)


def test_interval_that_ends_in_ff_and_one_that_starts_with_00():
    assert chunked(b"\x01\xff") == b"\x01\xff"
    assert chunked(b"\x00\x01") == b"\x00\x01"           # the byte in front of the interval does not count
