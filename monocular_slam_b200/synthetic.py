"""Deterministic synthetic inputs for tests and bench.py (numpy only; no OpenCV needed on the GPU box).

The reference ships no frames and no descriptor fixtures, and there is no
network, so every workload in BASELINE.json is synthesised here from fixed seeds:

* ``frame``      -- corner-dense grayscale frame: smoothed noise + random flat rectangles
                    (numpy restatement of the SURVEY.md Appendix B generator).
* ``natural_frame`` -- a frame with camera-image statistics (0.25 % FAST corners instead of 18 %).
* ``sequence``   -- consecutive frames of a monocular sequence: one large texture viewed through a
                    window that translates a few pixels per frame, so frame i matches frame i-1.
* ``descriptors``/``planted_queries`` -- uniform random 256-bit descriptors, and queries that are
                    near-duplicates of train rows (config 4 of BASELINE.json).
"""
import numpy as np


def _gauss_blur(img, sigma):
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    f = img.astype(np.float32)
    p = np.pad(f, ((0, 0), (r, r)), mode="reflect")
    acc = np.zeros_like(f)
    for i, kv in enumerate(k):
        acc += np.float32(kv) * p[:, i:i + f.shape[1]]
    p = np.pad(acc, ((r, r), (0, 0)), mode="reflect")
    out = np.zeros_like(f)
    for i, kv in enumerate(k):
        out += np.float32(kv) * p[i:i + f.shape[0], :]
    return out


def frame(seed, w, h, nrect=300, sigma=1.5):
    """One corner-dense frame (h x w uint8)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    f = _gauss_blur(img, sigma)
    lo, hi = float(f.min()), float(f.max())
    img = np.clip(np.rint((f - lo) * (255.0 / max(hi - lo, 1e-6))), 0, 255).astype(np.uint8)
    for _ in range(nrect):
        x = int(rng.integers(0, max(w - 40, 1)))
        y = int(rng.integers(0, max(h - 40, 1)))
        rw, rh = (int(v) for v in rng.integers(5, 40, 2))
        img[y:y + rh, x:x + rw] = int(rng.integers(0, 256))
    return img


def textured_frame(seed, w, h):
    """A less adversarial frame: low-frequency shading + mid-frequency texture + a few hundred rectangles."""
    rng = np.random.default_rng(seed)
    coarse = _gauss_blur(rng.integers(0, 256, (h, w), dtype=np.uint8), 6.0)
    fine = _gauss_blur(rng.integers(0, 256, (h, w), dtype=np.uint8), 1.0)
    f = 0.6 * (coarse - coarse.mean()) * 6.0 + 0.4 * (fine - fine.mean()) * 1.5 + 128.0
    img = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    for _ in range(200):
        x = int(rng.integers(0, max(w - 60, 1)))
        y = int(rng.integers(0, max(h - 60, 1)))
        rw, rh = (int(v) for v in rng.integers(8, 60, 2))
        img[y:y + rh, x:x + rw] = int(rng.integers(0, 256))
    return img


def natural_frame(seed, w, h):
    """A frame with the statistics of a camera image rather than of a stress test: 1/f-like shading (blurred noise at
    three scales), a few hundred flat objects with their own brightness (edges, corners at their vertices), some textured
    patches, lens blur (sigma 0.9) and sensor noise (sigma 1.2).  At 1080p about 0.25 % of the pixels pass FAST-9 at
    threshold 20 (``frame``: 18 %, ``textured_frame``: 4 %), and ORB still finds its 2000 keypoints."""
    rng = np.random.default_rng(seed)
    f = np.zeros((h, w), np.float32)
    for sigma, amp in ((24.0, 60.0), (8.0, 25.0), (3.0, 8.0)):
        n = _gauss_blur(rng.integers(0, 256, (h, w), dtype=np.uint8), sigma)
        n = (n - n.mean()) / max(float(n.std()), 1e-6)
        f += np.float32(amp) * n
    f += 128
    for _ in range(260):
        x = int(rng.integers(0, max(w - 120, 1)))
        y = int(rng.integers(0, max(h - 120, 1)))
        rw, rh = (int(v) for v in rng.integers(12, 120, 2))
        f[y:y + rh, x:x + rw] = 0.35 * f[y:y + rh, x:x + rw] + float(rng.integers(20, 235)) * 0.65
    for _ in range(24):
        x = int(rng.integers(0, max(w - 160, 1)))
        y = int(rng.integers(0, max(h - 160, 1)))
        rw, rh = (int(v) for v in rng.integers(40, 160, 2))
        patch = f[y:y + rh, x:x + rw]
        patch += rng.normal(0, 22, patch.shape).astype(np.float32)
    f = _gauss_blur(np.clip(f, 0, 255), 0.9)
    f += rng.normal(0, 1.2, (h, w)).astype(np.float32)
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def sequence(nframes, w, h, seed=0, step=(3, 5), generator=frame):
    """``nframes`` consecutive h x w views of one texture, translating ``step`` = (dy, dx) pixels per frame."""
    dy, dx = step
    big = generator(seed, w + abs(dx) * nframes, h + abs(dy) * nframes)
    out = np.empty((nframes, h, w), np.uint8)
    for i in range(nframes):
        out[i] = big[i * abs(dy):i * abs(dy) + h, i * abs(dx):i * abs(dx) + w]
    return out


def descriptors(seed, n):
    """n uniform-random 256-bit descriptors (n x 32 uint8)."""
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def planted_queries(seed, train, nq, frac=0.5, max_flips=20):
    """Queries of which ``frac`` are copies of random train rows with <= max_flips bits flipped (so the ratio test passes)."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    nplant = int(nq * frac)
    rows = rng.choice(nq, nplant, replace=False)
    src = rng.integers(0, len(train), nplant)
    q[rows] = train[src]
    for r in rows:
        nflip = int(rng.integers(0, max_flips + 1))
        bits = rng.choice(256, nflip, replace=False)
        for b in bits:
            q[r, b >> 3] ^= np.uint8(1 << (b & 7))
    return q


def bgr_frame(seed, w, h):
    """A 3-channel (BGR, 8UC3) frame: three differently seeded gray frames as channels, so the gray conversion matters."""
    return np.ascontiguousarray(np.stack([frame(seed + 101 * c, w, h, nrect=120) for c in range(3)], axis=2))


def two_view_matches(seed, n, inlier_ratio=0.7, noise=0.5, size=(1920, 1080)):
    """Matched keypoint positions of a synthetic two-view pair: ``n`` 3-D points seen by two cameras (small rotation +
    translation, focal 900 px), Gaussian pixel noise, and a fraction ``1 - inlier_ratio`` of the second view replaced by
    uniform positions (wrong matches).  Returns float32 (pts1[n, 2], pts2[n, 2]) -- the `inputs1` / `inputs2` arrays
    computeFundamentalMatrix builds (src/CameraPoseEstimator.cpp:555-560)."""
    r = np.random.default_rng(seed)
    w, h = size
    X = np.c_[r.uniform(-4, 4, n), r.uniform(-3, 3, n), r.uniform(4, 12, n)]
    K = np.array([[900.0, 0, w / 2], [0, 900.0, h / 2], [0, 0, 1]])
    a = 0.05
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([0.4, 0.05, 0.1])
    p1 = (K @ X.T).T
    p1 = p1[:, :2] / p1[:, 2:]
    p2 = (K @ (R @ X.T + t[:, None])).T
    p2 = p2[:, :2] / p2[:, 2:]
    p1 += r.normal(0, noise, p1.shape)
    p2 += r.normal(0, noise, p2.shape)
    bad = r.random(n) > inlier_ratio
    p2[bad] = np.c_[r.uniform(0, w, bad.sum()), r.uniform(0, h, bad.sum())]
    return p1.astype(np.float32), p2.astype(np.float32)


def layered_views(seed, w, h, nviews, nlayers=4, motion=(2, 5)):
    """``nviews`` views of the scene of ``layered_pair`` along the same sideways translation: in view v every layer is shifted
    by v times its own multiple of ``motion``.  View 0 and view 1 are exactly ``layered_pair(seed, w, h, nlayers, motion)``
    when nviews == 2."""
    r = np.random.default_rng(seed)
    dy, dx = abs(int(motion[0])), abs(int(motion[1]))
    big = frame(seed, w + (nviews - 1) * nlayers * dx + 8, h + (nviews - 1) * nlayers * dy + 8)
    edges = np.linspace(0, h, nlayers + 1).astype(int)
    mult = r.permutation(np.arange(1, nlayers + 1))
    out = np.empty((nviews, h, w), np.uint8)
    out[0] = big[:h, :w]
    for v in range(1, nviews):
        for layer in range(nlayers):
            sy, sx = v * dy * int(mult[layer]), v * dx * int(mult[layer])
            y0, y1 = int(edges[layer]), int(edges[layer + 1])
            out[v, y0:y1] = big[y0 + sy:y1 + sy, sx:sx + w]
    return out


def layered_pair(seed, w, h, nlayers=4, motion=(2, 5)):
    """Two views (h x w uint8 each) of a scene of fronto-parallel layers at different depths under a sideways camera
    translation: every layer is a horizontal band of one big texture and shifts by its own multiple of ``motion`` (dy, dx).
    The flow vectors are parallel but of different lengths, i.e. a non-planar scene with the epipole at infinity -- unlike
    ``sequence`` (one translating plane), its fundamental matrix is well defined."""
    r = np.random.default_rng(seed)
    dy, dx = abs(int(motion[0])), abs(int(motion[1]))
    big = frame(seed, w + nlayers * dx + 8, h + nlayers * dy + 8)
    a = big[:h, :w].copy()
    b = np.empty_like(a)
    edges = np.linspace(0, h, nlayers + 1).astype(int)
    mult = r.permutation(np.arange(1, nlayers + 1))
    for layer in range(nlayers):
        sy, sx = dy * int(mult[layer]), dx * int(mult[layer])
        y0, y1 = int(edges[layer]), int(edges[layer + 1])
        b[y0:y1] = big[y0 + sy:y1 + sy, sx:sx + w]
    return a, b


# ---- bag-of-words vocabulary (DBoW2, ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h): there is no network to fetch a trained
# ORB vocabulary from, so the tests and the bench use a synthetic tree of the same shape

def vocabulary(seed, k=10, L=3, ragged=False, stop_frac=0.03):
    """A k-ary vocabulary tree of depth L as the arrays the text format holds, in the order HKmeansStep creates nodes (the k
    children of a node get consecutive ids, then each child is expanded): ``parent`` [N] int32 (node 0 is the root),
    ``leaf`` [N] uint8 (the file's isLeaf column: flagged nodes get word ids in node order), ``desc`` [N][32] uint8,
    ``weight`` [N] float64.  A child is its parent with fewer and fewer bits flipped per level, so descents are decided by a
    few bits and ties between children are common.  ``ragged``: some nodes have fewer than k children, some branches end
    early, and one childless node is not flagged as a word (the reference then ends the descent there with word id 0)."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]
    unflagged_done = not ragged

    def flips(d, nbits):
        bits = np.unpackbits(d)
        idx = rng.choice(256, size=nbits, replace=False)
        bits[idx] ^= 1
        return np.packbits(bits)

    def expand(node, level):
        nonlocal unflagged_done
        nchild = k if not ragged else int(rng.integers(max(1, k - 3), k + 1))
        first = len(parent)
        for _ in range(nchild):
            base = desc[node] if level > 1 else rng.integers(0, 256, 32, dtype=np.uint8)
            parent.append(node)
            leaf.append(0)
            desc.append(flips(base, max(2, 96 >> (level - 1))) if level > 1 else base)
            weight.append(0.0)
        for c in range(first, first + nchild):
            ends_early = ragged and level >= 2 and rng.random() < 0.15
            if level == L or ends_early:
                if ragged and not unflagged_done and level == L:
                    unflagged_done = True                 # childless, not a word: weight from the file, word id 0
                    weight[c] = float(rng.uniform(0.5, 9.0))
                    continue
                leaf[c] = 1
                weight[c] = 0.0 if rng.random() < stop_frac else float(rng.uniform(0.5, 9.0))
            else:
                expand(c, level + 1)

    expand(0, 1)
    return {"k": int(k), "L": int(L), "parent": np.asarray(parent, np.int32), "leaf": np.asarray(leaf, np.uint8),
            "desc": np.stack(desc).astype(np.uint8), "weight": np.asarray(weight, np.float64)}


def vocabulary_features(seed, voc, n, pool=None, max_flips=6):
    """n descriptors near the vocabulary's words: each is a word's descriptor (drawn from a pool of ``pool`` words, so words
    repeat within a frame) with up to ``max_flips`` bits flipped."""
    rng = np.random.default_rng(seed)
    words = np.flatnonzero(voc["leaf"])
    if pool is not None and pool < len(words):
        words = rng.choice(words, size=pool, replace=False)
    out = np.empty((n, 32), np.uint8)
    for i in range(n):
        bits = np.unpackbits(voc["desc"][rng.choice(words)])
        nf = int(rng.integers(0, max_flips + 1))
        if nf:
            bits[rng.choice(256, size=nf, replace=False)] ^= 1
        out[i] = np.packbits(bits)
    return out


def write_vocabulary_text(path, voc, scoring=0, weighting=0, trailing_newline=False):
    """saveToTextFile's format (TemplatedVocabulary.h:1426-1448): "k L  scoring weighting", then one line per node after the
    root: parent, isLeaf, the 32 descriptor bytes, the weight.  The reference writes a newline after the last node, and its
    loader (:1329-1416, `while(!f.eof())`) then reads one more, empty, line as an extra child of the root whose descriptor
    is never initialised; ``trailing_newline=False`` leaves that newline out so the file loads as what it says."""
    lines = ["%d %d  %d %d" % (voc["k"], voc["L"], scoring, weighting)]
    for i in range(1, len(voc["parent"])):
        lines.append("%d %d %s %s" % (voc["parent"][i], 1 if voc["leaf"][i] else 0, " ".join(str(int(b)) for b in voc["desc"][i]) + " ",
                                      repr(float(voc["weight"][i]))))
    with open(path, "w") as f:
        f.write("\n".join(lines) + ("\n" if trailing_newline else ""))


def read_vocabulary_text(path):
    """The arrays of `vocabulary` plus ``scoring`` / ``weighting`` from a file in loadFromTextFile's format (empty lines are
    skipped: the phantom node of a trailing newline is not reproduced, see write_vocabulary_text)."""
    with open(path) as f:
        head = f.readline().split()
        k, L, scoring, weighting = (int(x) for x in head[:4])
        parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]
        for line in f:
            t = line.split()
            if not t:
                continue
            parent.append(int(t[0]))
            leaf.append(1 if int(t[1]) > 0 else 0)
            desc.append(np.asarray([int(x) for x in t[2:34]], np.uint8))
            weight.append(float(t[34]))
    return {"k": k, "L": L, "scoring": scoring, "weighting": weighting, "parent": np.asarray(parent, np.int32),
            "leaf": np.asarray(leaf, np.uint8), "desc": np.stack(desc), "weight": np.asarray(weight, np.float64)}


def vocabulary_large(seed, k=10, L=6, stop_frac=0.01):
    """A complete k-ary tree of depth L in breadth-first node order, generated level by level with array operations (the
    ORB-SLAM vocabulary shape k = 10, L = 6 has 1.1 M nodes): level-1 descriptors are random, a child differs from its parent
    in a random subset of bits that halves with every level (the AND of `level` random byte arrays: 64, 32, 16 ... bits on
    average).  Same dictionary as `vocabulary`."""
    rng = np.random.default_rng(seed)
    counts = [k ** lvl for lvl in range(L + 1)]
    n = sum(counts)
    parent = np.zeros(n, np.int32)
    desc = np.zeros((n, 32), np.uint8)
    first = np.cumsum([0] + counts)           # first node id of every level
    for lvl in range(1, L + 1):
        lo, hi = first[lvl], first[lvl + 1]
        par = first[lvl - 1] + np.arange(counts[lvl], dtype=np.int64) // k
        parent[lo:hi] = par
        if lvl == 1:
            desc[lo:hi] = rng.integers(0, 256, (counts[lvl], 32), dtype=np.uint8)
        else:
            mask = rng.integers(0, 256, (counts[lvl], 32), dtype=np.uint8)
            for _ in range(lvl - 1):
                mask &= rng.integers(0, 256, (counts[lvl], 32), dtype=np.uint8)
            desc[lo:hi] = desc[par] ^ mask
    leaf = np.zeros(n, np.uint8)
    leaf[first[L]:] = 1
    weight = np.zeros(n, np.float64)
    w = rng.uniform(0.5, 9.0, counts[L])
    w[rng.random(counts[L]) < stop_frac] = 0.0
    weight[first[L]:] = w
    return {"k": int(k), "L": int(L), "parent": parent, "leaf": leaf, "desc": desc, "weight": weight}
