#!/usr/bin/env python
"""Golden vectors for the bag-of-words path (tests/golden/bow_cases.npz), produced by the REFERENCE'S OWN CODE.

The reference vendors DBoW2 (ThirdParty/DBoW2: TemplatedVocabulary.h, FORB.cpp, BowVector.cpp, FeatureVector.cpp,
ScoringObject.cpp).  Unlike src/, those files need nothing from OpenCV but cv::Mat as a 1x32 byte row, so `make -C oracle
dbow_ref` compiles them where they lie under /root/reference against a stand-in header (oracle/dbow_ref/) into
oracle/_ref/libdbow_ref.so with oracle/dbow_ref_harness.cpp in front.  This script writes synthetic vocabularies in the
reference's text format (without the trailing newline: see synthetic.write_vocabulary_text), has the reference load them
(loadFromTextFile), and stores what its transform / score / stopWords / getParentNode return.

Cases: a regular 10-ary tree of depth 3 and a ragged 7-ary tree of depth 4 (missing children, branches that end early, a
childless node that is not a word); frames of 300 / 57 / 1 / 0 features with repeated words; every weighting x scoring
for the bag-of-words vectors and their pairwise scores; feature vectors at every levelsup for which the reference defines
them (a descent that ends above level L - levelsup leaves the node id unwritten there).

Run (build container only; needs /root/reference):  python tests/golden/make_golden_bow.py
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from monocular_slam_b200 import synthetic as syn  # noqa: E402

u8p, u32p, i32p, f64p, ip = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_int32, C.c_double, C.c_int))


def ref_lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "dbow_ref"])
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdbow_ref.so"))
    L.dbowref_load_text.restype = C.c_void_p
    L.dbowref_load_text.argtypes = [C.c_char_p]
    L.dbowref_free.argtypes = [C.c_void_p]
    L.dbowref_info.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
    L.dbowref_transform.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, u32p, f64p, ip, u32p, i32p, u32p, ip]
    L.dbowref_transform_bow.argtypes = [C.c_void_p, u8p, C.c_int, u32p, f64p, ip]
    L.dbowref_words.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, u32p, f64p, u32p]
    L.dbowref_score.restype = C.c_double
    L.dbowref_score.argtypes = [C.c_void_p, u32p, f64p, C.c_int, u32p, f64p, C.c_int]
    L.dbowref_stop_words.argtypes = [C.c_void_p, C.c_double]
    return L


def p(a, t):
    return a.ctypes.data_as(t)


def ref_transform(L, h, desc, levelsup):
    n = len(desc)
    m = max(n, 1)
    words, vals, nb = np.zeros(m, np.uint32), np.zeros(m, np.float64), C.c_int(0)
    nodes, offs, feats, nf = np.zeros(m, np.uint32), np.zeros(m + 1, np.int32), np.zeros(m, np.uint32), C.c_int(0)
    d = np.ascontiguousarray(desc, np.uint8)
    L.dbowref_transform(h, p(d, u8p), n, levelsup, p(words, u32p), p(vals, f64p), C.byref(nb), p(nodes, u32p), p(offs, i32p), p(feats, u32p),
                        C.byref(nf))
    k = nf.value
    return words[:nb.value].copy(), vals[:nb.value].copy(), nodes[:k].copy(), offs[:k + 1].copy(), feats[:offs[k]].copy()


def ref_transform_bow(L, h, desc):
    n = len(desc)
    m = max(n, 1)
    words, vals, nb = np.zeros(m, np.uint32), np.zeros(m, np.float64), C.c_int(0)
    d = np.ascontiguousarray(desc, np.uint8)
    L.dbowref_transform_bow(h, p(d, u8p), n, p(words, u32p), p(vals, f64p), C.byref(nb))
    return words[:nb.value].copy(), vals[:nb.value].copy()


def main():
    L = ref_lib()
    out = {}
    tmp = tempfile.mkdtemp()
    vocs = {"A": syn.vocabulary(11, k=10, L=3), "B": syn.vocabulary(12, k=7, L=4, ragged=True)}
    fv_levels = {"A": [0, 1, 2, 3, 4], "B": [2, 3, 4, 5]}         # B's branches end from level 2 on: L - levelsup <= 2 is defined
    for name, voc in vocs.items():
        for key in ("parent", "leaf", "desc", "weight"):
            out["%s_%s" % (name, key)] = voc[key]
        out["%s_kL" % name] = np.array([voc["k"], voc["L"]], np.int32)
        frames = [syn.vocabulary_features(100 + i, voc, n, pool=pool) for i, (n, pool) in enumerate([(300, 80), (57, 30), (1, None)])]
        frames.append(np.zeros((0, 32), np.uint8))
        for i, f in enumerate(frames):
            out["%s_frame%d" % (name, i)] = f
        combos = [(s, w) for s in range(6) for w in range(4)] if name == "B" else [(0, 0), (5, 1), (1, 2)]
        for scoring, weighting in combos:
            path = os.path.join(tmp, "%s_%d_%d.txt" % (name, scoring, weighting))
            syn.write_vocabulary_text(path, voc, scoring, weighting)
            h = L.dbowref_load_text(path.encode())
            assert h
            k, Lv, nw, sc, we = (C.c_int() for _ in range(5))
            L.dbowref_info(h, k, Lv, nw, sc, we)
            assert (k.value, Lv.value, nw.value, sc.value, we.value) == (voc["k"], voc["L"], int(voc["leaf"].sum()), scoring, weighting)
            tag = "%s_s%d_w%d" % (name, scoring, weighting)
            bows = []
            for i, f in enumerate(frames):
                w_, v_ = ref_transform_bow(L, h, f)
                out["%s_f%d_words" % (tag, i)], out["%s_f%d_vals" % (tag, i)] = w_, v_
                bows.append((w_, v_))
            sm = np.zeros((len(frames), len(frames)), np.float64)
            for a in range(len(frames)):
                for b in range(len(frames)):
                    (w1, v1), (w2, v2) = bows[a], bows[b]
                    sm[a, b] = L.dbowref_score(h, p(w1, u32p), p(v1, f64p), len(w1), p(w2, u32p), p(v2, f64p), len(w2))
            out["%s_scores" % tag] = sm
            if (scoring, weighting) == (0, 0):
                for lu in fv_levels[name]:
                    for i, f in enumerate(frames):
                        w_, v_, nodes, offs, feats = ref_transform(L, h, f, lu)
                        assert np.array_equal(w_, bows[i][0]) and np.array_equal(v_, bows[i][1])
                        out["%s_f%d_l%d_nodes" % (tag, i, lu)], out["%s_f%d_l%d_offs" % (tag, i, lu)] = nodes, offs
                        out["%s_f%d_l%d_feats" % (tag, i, lu)] = feats
                f = frames[0]
                word, weight, par = np.zeros(len(f), np.uint32), np.zeros(len(f), np.float64), np.zeros(len(f), np.uint32)
                for lu in (0, 1, 2, 7):
                    L.dbowref_words(h, p(np.ascontiguousarray(f), u8p), len(f), lu, p(word, u32p), p(weight, f64p), p(par, u32p))
                    out["%s_words_l%d_parent" % (tag, lu)] = par.copy()
                out["%s_words_word" % tag], out["%s_words_weight" % tag] = word.copy(), weight.copy()
                # stopWords(2.0), then the same frame again
                out["%s_stop_count" % tag] = np.array([L.dbowref_stop_words(h, 2.0)], np.int32)
                w_, v_, nodes, offs, feats = ref_transform(L, h, frames[0], fv_levels[name][0])
                out["%s_stop_words" % tag], out["%s_stop_vals" % tag] = w_, v_
                out["%s_stop_nodes" % tag], out["%s_stop_offs" % tag], out["%s_stop_feats" % tag] = nodes, offs, feats
            L.dbowref_free(h)
    path = os.path.join(HERE, "bow_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %.0f KB" % (path, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
