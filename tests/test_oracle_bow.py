"""oracle/bow_oracle.c against tests/golden/bow_cases.npz, which the reference's own DBoW2 produced
(tests/golden/make_golden_bow.py): bag-of-words vectors and feature vectors bit for bit, scores bit for bit except KL
(log() of the C library vs itself: still exact here, 1e-12 allowed)."""
import os

import numpy as np
import pytest

import oracle
from monocular_slam_b200 import synthetic as syn

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bow_cases.npz"))
NFRAMES = 4


def voc_arrays(name):
    k, L = (int(x) for x in G["%s_kL" % name])
    return {"k": k, "L": L, "parent": G["%s_parent" % name], "leaf": G["%s_leaf" % name], "desc": G["%s_desc" % name],
            "weight": G["%s_weight" % name]}


def combos(name):
    return sorted({(int(k.split("_")[1][1:]), int(k.split("_")[2][1:])) for k in G.files if k.startswith(name + "_s") and k.endswith("_scores")})


def test_golden_generators_are_reproducible():
    # the vocabularies and frames in the fixture are what the seeded generators produce today
    a = syn.vocabulary(11, k=10, L=3)
    assert all(np.array_equal(a[k], G["A_" + k]) for k in ("parent", "leaf", "desc", "weight"))
    assert np.array_equal(syn.vocabulary_features(100, a, 300, pool=80), G["A_frame0"])


@pytest.mark.parametrize("name", ["A", "B"])
def test_bow_vectors_and_scores_all_weightings_and_scorings(name):
    va = voc_arrays(name)
    assert len(combos(name)) == (24 if name == "B" else 3)
    for scoring, weighting in combos(name):
        voc = oracle.BowVocabulary(va, scoring, weighting)
        assert voc.size() == int(va["leaf"].sum())
        tag = "%s_s%d_w%d" % (name, scoring, weighting)
        bows = []
        for i in range(NFRAMES):
            w, v = voc.transform(G["%s_frame%d" % (name, i)])
            assert np.array_equal(w, G["%s_f%d_words" % (tag, i)]), (tag, i)
            assert np.array_equal(v, G["%s_f%d_vals" % (tag, i)]), (tag, i)
            bows.append((w, v))
        ref = G["%s_scores" % tag]
        for a in range(NFRAMES):
            for b in range(NFRAMES):
                s = voc.score(bows[a], bows[b])
                if scoring == oracle.KL:
                    assert abs(s - ref[a, b]) <= 1e-12 * max(1.0, abs(ref[a, b])), (tag, a, b)
                else:
                    assert s == ref[a, b], (tag, a, b, s, ref[a, b])


@pytest.mark.parametrize("name,levels", [("A", [0, 1, 2, 3, 4]), ("B", [2, 3, 4, 5])])
def test_feature_vectors(name, levels):
    voc = oracle.BowVocabulary(voc_arrays(name))
    tag = "%s_s0_w0" % name
    for lu in levels:
        for i in range(NFRAMES):
            w, v, nodes, offs, feats = voc.transform(G["%s_frame%d" % (name, i)], lu)
            assert np.array_equal(w, G["%s_f%d_words" % (tag, i)]) and np.array_equal(v, G["%s_f%d_vals" % (tag, i)])
            assert np.array_equal(nodes, G["%s_f%d_l%d_nodes" % (tag, i, lu)]), (name, lu, i)
            assert np.array_equal(offs, G["%s_f%d_l%d_offs" % (tag, i, lu)])
            assert np.array_equal(feats, G["%s_f%d_l%d_feats" % (tag, i, lu)])


@pytest.mark.parametrize("name", ["A", "B"])
def test_words_parents_and_stop_words(name):
    voc = oracle.BowVocabulary(voc_arrays(name))
    tag = "%s_s0_w0" % name
    f = G["%s_frame0" % name]
    word, weight, _ = voc.transform_features(f)
    assert np.array_equal(word, G["%s_words_word" % tag])
    # getWordWeight(word): the weight of the word's node (differs from the descent's weight only for the unflagged node)
    assert np.array_equal(np.array([voc.word_weight(w) for w in word]), G["%s_words_weight" % tag])
    for lu in (0, 1, 2, 7):
        assert np.array_equal(np.array([voc.parent_node(w, lu) for w in word], np.uint32), G["%s_words_l%d_parent" % (tag, lu)])
    assert voc.stop_words(2.0) == int(G["%s_stop_count" % tag][0])
    lu = 0 if name == "A" else 2
    w, v, nodes, offs, feats = voc.transform(f, lu)
    for got, key in ((w, "words"), (v, "vals"), (nodes, "nodes"), (offs, "offs"), (feats, "feats")):
        assert np.array_equal(got, G["%s_stop_%s" % (tag, key)]), key


def test_text_format_round_trip(tmp_path):
    voc = syn.vocabulary(5, k=4, L=3, ragged=True)
    path = str(tmp_path / "voc.txt")
    syn.write_vocabulary_text(path, voc, scoring=2, weighting=1)
    back = syn.read_vocabulary_text(path)
    assert (back["k"], back["L"], back["scoring"], back["weighting"]) == (4, 3, 2, 1)
    assert all(np.array_equal(voc[k], back[k]) for k in ("parent", "leaf", "desc", "weight"))
    # the reference's own writer ends the file with a newline: tolerated (the phantom node is not reproduced)
    syn.write_vocabulary_text(path, voc, trailing_newline=True)
    assert len(syn.read_vocabulary_text(path)["parent"]) == len(voc["parent"])
