"""GPU parity of the bag-of-words path (csrc/bow.cu through the C ABI, `bowx_*`): against the golden vectors the reference's
own DBoW2 produced (tests/golden/bow_cases.npz) and, on larger seeded cases, against oracle/bow_oracle.c.  Words, node ids,
bag-of-words values, feature vectors and scores are compared bit for bit; KL scores (log()) to 1e-12 relative."""
import os

import numpy as np
import pytest

import oracle
from monocular_slam_b200 import ORB, OrbxError, Vocabulary
from monocular_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bow_cases.npz"))
NFRAMES = 4


def voc_arrays(name):
    k, L = (int(x) for x in G["%s_kL" % name])
    return {"k": k, "L": L, "parent": G["%s_parent" % name], "leaf": G["%s_leaf" % name], "desc": G["%s_desc" % name],
            "weight": G["%s_weight" % name]}


def combos(name):
    return sorted({(int(k.split("_")[1][1:]), int(k.split("_")[2][1:])) for k in G.files if k.startswith(name + "_s") and k.endswith("_scores")})


def close(s, ref, scoring):
    if scoring == oracle.KL:
        return abs(s - ref) <= 1e-12 * max(1.0, abs(ref))
    return s == ref


@pytest.mark.parametrize("name", ["A", "B"])
def test_golden_bow_vectors_and_scores(name):
    va = voc_arrays(name)
    voc = Vocabulary()
    for scoring, weighting in combos(name):
        voc.set(va, scoring, weighting)
        assert (voc.size(), voc.getBranchingFactor(), voc.getDepthLevels(), voc.getScoringType(), voc.getWeightingType()) == \
            (int(va["leaf"].sum()), va["k"], va["L"], scoring, weighting)
        tag = "%s_s%d_w%d" % (name, scoring, weighting)
        bows = []
        for i in range(NFRAMES):
            w, v = voc.transform(G["%s_frame%d" % (name, i)])
            assert np.array_equal(w, G["%s_f%d_words" % (tag, i)]), (tag, i)
            assert np.array_equal(v, G["%s_f%d_vals" % (tag, i)]), (tag, i)
            bows.append((w, v))
        ref = G["%s_scores" % tag]
        for a in range(NFRAMES):
            row = voc.score_batch(bows[a], bows)
            for b in range(NFRAMES):
                assert close(row[b], ref[a, b], scoring), (tag, a, b, row[b], ref[a, b])
                assert close(voc.score(bows[a], bows[b]), ref[a, b], scoring)
    voc.close()


@pytest.mark.parametrize("name,levels", [("A", [0, 1, 2, 3, 4]), ("B", [2, 3, 4, 5])])
def test_golden_feature_vectors(name, levels):
    voc = Vocabulary(voc_arrays(name))
    tag = "%s_s0_w0" % name
    for lu in levels:
        for i in range(NFRAMES):
            (w, v), (nodes, offs, feats) = voc.transform(G["%s_frame%d" % (name, i)], lu)
            assert np.array_equal(w, G["%s_f%d_words" % (tag, i)]) and np.array_equal(v, G["%s_f%d_vals" % (tag, i)])
            assert np.array_equal(nodes, G["%s_f%d_l%d_nodes" % (tag, i, lu)]), (name, lu, i)
            assert np.array_equal(offs, G["%s_f%d_l%d_offs" % (tag, i, lu)])
            assert np.array_equal(feats, G["%s_f%d_l%d_feats" % (tag, i, lu)])
    voc.close()


@pytest.mark.parametrize("name", ["A", "B"])
def test_golden_words_parents_and_stop_words(name):
    voc = Vocabulary(voc_arrays(name))
    tag = "%s_s0_w0" % name
    f = G["%s_frame0" % name]
    word, weight, _ = voc.transform_features(f)
    assert np.array_equal(word, G["%s_words_word" % tag])
    assert np.array_equal(np.array([voc.getWordWeight(w) for w in word]), G["%s_words_weight" % tag])
    assert all(voc.transform_word(f[i]) == word[i] for i in range(0, len(f), 37))
    for lu in (0, 1, 2, 7):
        assert np.array_equal(np.array([voc.getParentNode(w, lu) for w in word], np.uint32), G["%s_words_l%d_parent" % (tag, lu)])
    assert voc.stopWords(2.0) == int(G["%s_stop_count" % tag][0])
    lu = 0 if name == "A" else 2
    (w, v), (nodes, offs, feats) = voc.transform(f, lu)
    for got, key in ((w, "words"), (v, "vals"), (nodes, "nodes"), (offs, "offs"), (feats, "feats")):
        assert np.array_equal(got, G["%s_stop_%s" % (tag, key)]), key
    voc.close()


# ---- larger seeded cases against the oracle

SHAPES = [
    dict(k=10, L=5),                    # 111 k nodes, the shape of an ORB vocabulary one level short
    dict(k=16, L=3),                    # the widest node a 16-lane group holds in one pass
    dict(k=19, L=3, ragged=True),       # 17..19 children: 32 lanes per descriptor
    dict(k=45, L=2, ragged=True),       # more children than lanes: several passes per level
    dict(k=2, L=9),                     # deep and narrow
    dict(k=1, L=4),                     # a chain
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "k%d_L%d%s" % (s["k"], s["L"], "_ragged" if s.get("ragged") else ""))
def test_descent_matches_oracle(shape):
    va = syn.vocabulary(21, **shape)
    feats = np.concatenate([syn.vocabulary_features(3, va, 700, max_flips=40), syn.descriptors(8, 300)])
    ov, gv = oracle.BowVocabulary(va), Vocabulary(va)
    for lu in (0, 1, shape["L"] - 1, shape["L"], shape["L"] + 3):
        w0, wt0, n0 = ov.transform_features(feats, lu)
        w1, wt1, n1 = gv.transform_features(feats, lu)
        assert np.array_equal(w0, w1) and np.array_equal(wt0, wt1) and np.array_equal(n0, n1), (shape, lu)
    gv.close()


@pytest.mark.parametrize("scoring,weighting", [(0, 0), (1, 0), (5, 1), (3, 2), (2, 3), (4, 0)])
def test_batch_matches_oracle(scoring, weighting):
    va = syn.vocabulary(31, k=10, L=4, ragged=True)
    ov, gv = oracle.BowVocabulary(va, scoring, weighting), Vocabulary(va, scoring, weighting)
    cap = 1500
    counts = np.array([1500, 0, 1, 777, 1499, 32, 33, 1024], np.int32)
    desc = np.zeros((len(counts), cap, 32), np.uint8)
    for f, n in enumerate(counts):
        desc[f, :n] = syn.vocabulary_features(50 + f, va, n, pool=200 if f % 2 else None)
        desc[f, n:] = 0xAB                       # rows past counts[f] must be ignored
    lu = 2                                      # branches end from level 2 on: level L - 2 is reached by every descent
    got = gv.transform_batch(desc, counts, lu)
    bows = []
    for f, n in enumerate(counts):
        w, v, nodes, offs, fe = ov.transform(desc[f, :n], lu)
        (gw, gvals), (gn, go, gf) = got[f]
        assert np.array_equal(w, gw) and np.array_equal(v, gvals), f
        assert np.array_equal(nodes, gn) and np.array_equal(offs, go) and np.array_equal(fe, gf), f
        bows.append((w, v))
    # without the feature vector: same vectors
    for f, bow in enumerate(gv.transform_batch(desc, counts)):
        assert np.array_equal(bow[0], bows[f][0]) and np.array_equal(bow[1], bows[f][1])
    # every frame against the database of all of them
    for a in range(len(counts)):
        row = gv.score_batch(bows[a], bows)
        for b in range(len(counts)):
            assert close(row[b], ov.score(bows[a], bows[b]), scoring), (a, b)
    assert len(gv.score_batch(bows[0], [])) == 0
    gv.close()


def test_device_chain_from_the_extractor():
    """Descriptors straight from orbx_extract_batch_dev into bowx_transform_batch_dev, scores on the device."""
    import torch
    B, W, H = 4, 640, 480
    seq = syn.sequence(B, W, H, seed=4)
    va = syn.vocabulary(41, k=10, L=4)
    orb = ORB(nfeatures=800, max_size=(W, H), max_batch=B)
    voc = Vocabulary(va)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        orb.set_stream(stream.cuda_stream)
        voc.set_stream(stream.cuda_stream)
        cap = orb.default_cap
        d_frames = torch.from_numpy(seq).cuda()
        d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda")
        d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        d_w = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
        d_v = torch.zeros((B, cap), dtype=torch.float64, device="cuda")
        d_nb = torch.zeros(B, dtype=torch.int32, device="cuda")
        d_nodes = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
        d_offs = torch.zeros((B, cap + 1), dtype=torch.int32, device="cuda")
        d_feats = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
        d_nfv = torch.zeros(B, dtype=torch.int32, device="cuda")
        d_scores = torch.zeros(B, dtype=torch.float64, device="cuda")
        d_start = torch.arange(B, dtype=torch.int64, device="cuda") * cap
        for _ in range(2):
            orb.extract_batch_dev(d_frames.data_ptr(), W * H, B, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
            voc.transform_batch_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, cap, 2, d_w.data_ptr(), d_v.data_ptr(), d_nb.data_ptr(),
                                    d_nodes.data_ptr(), d_offs.data_ptr(), d_feats.data_ptr(), d_nfv.data_ptr())
        orb.check_dev()
        stream.synchronize()
        nb = d_nb.cpu().numpy()
        # frame 1 as the query against the padded output as the database
        voc.score_batch_dev(d_w[1].data_ptr(), d_v[1].data_ptr(), int(nb[1]), d_start.data_ptr(), d_nb.data_ptr(), d_w.data_ptr(), d_v.data_ptr(), B,
                            d_scores.data_ptr())
        stream.synchronize()
    cnt, desc = d_cnt.cpu().numpy(), d_desc.cpu().numpy()
    ov = oracle.BowVocabulary(va)
    bows = []
    for f in range(B):
        w, v, nodes, offs, fe = ov.transform(desc[f, :cnt[f]], 2)
        assert cnt[f] > 300 and len(w) > 50
        assert np.array_equal(d_w[f, :nb[f]].cpu().numpy().view(np.uint32), w) and np.array_equal(d_v[f, :nb[f]].cpu().numpy(), v)
        g = int(d_nfv[f])
        assert np.array_equal(d_nodes[f, :g].cpu().numpy().view(np.uint32), nodes)
        assert np.array_equal(d_offs[f, :g + 1].cpu().numpy(), offs)
        assert np.array_equal(d_feats[f, :offs[-1]].cpu().numpy().view(np.uint32), fe)
        bows.append((w, v))
    s = d_scores.cpu().numpy()
    assert s[1] == ov.score(bows[1], bows[1]) and abs(s[1] - 1.0) < 1e-12
    assert all(s[f] == ov.score(bows[1], bows[f]) for f in range(B))
    voc.close(); orb.close()


def test_empty_and_errors():
    voc = Vocabulary()
    with pytest.raises(OrbxError):
        voc.transform(syn.descriptors(1, 10))                  # no vocabulary yet
    va = syn.vocabulary(5, k=4, L=2)
    with pytest.raises(OrbxError):
        bad = dict(va); bad["parent"] = va["parent"].copy(); bad["parent"][3] = 7
        voc.set(bad)                                           # a parent after its child
    none = dict(va); none["leaf"] = np.zeros_like(va["leaf"])
    voc.set(none)                                              # no words: empty(), transform() returns empty vectors
    assert voc.empty() and voc.transform_word(syn.descriptors(1, 1)[0]) == 0
    (w, v), (nodes, offs, feats) = voc.transform(syn.descriptors(1, 10), 1)
    assert len(w) == 0 and len(nodes) == 0 and len(feats) == 0
    voc.set(va)
    w, v = voc.transform(np.zeros((0, 32), np.uint8))
    assert len(w) == 0
    assert voc.score((w, v), (w, v)) == 0.0
    with pytest.raises(OrbxError):
        voc.transform_batch(np.zeros((1, 20000, 32), np.uint8))   # more features per frame than one CTA sorts
    with pytest.raises(OrbxError):
        voc.getParentNode(voc.size(), 0)
    voc.close()
