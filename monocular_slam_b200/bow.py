"""Bag-of-words vocabulary on the device: the mirror of the reference's vendored DBoW2 class
``TemplatedVocabulary<FORB::TDescriptor, FORB>`` (ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h) over include/orbx.h "bowx_*".

Method names follow the reference (``transform``, ``score``, ``size``, ``getParentNode``, ``getWordWeight``, ``stopWords``,
``loadFromTextFile`` ...).  A ``BowVector`` (std::map<WordId, WordValue>) is a pair of arrays ``(words ascending, values)``;
a ``FeatureVector`` (std::map<NodeId, vector<unsigned>>) is ``(nodes ascending, offsets, feature indices)``.  There is no
CPU implementation behind any of it: without the CUDA library the constructor raises.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = 0, 1, 2, 3, 4, 5


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Vocabulary:
    """``Vocabulary(voc)`` with ``voc`` the arrays of a text vocabulary (``synthetic.read_vocabulary_text`` /
    ``synthetic.vocabulary``: k, L, parent, leaf, desc, weight and optionally scoring, weighting), or ``Vocabulary()`` followed
    by ``loadFromTextFile(path)``."""

    def __init__(self, voc=None, scoring=L1_NORM, weighting=TF_IDF, device=0):
        self._h = C.c_void_p()
        self.device = device
        check(_lib.lib().bowx_create(C.byref(self._h), device))
        if voc is not None:
            self.set(voc, scoring, weighting)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().bowx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(_lib.lib().bowx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(_lib.lib().bowx_synchronize(self._h))

    # ---- the vocabulary itself
    def set(self, voc, scoring=L1_NORM, weighting=TF_IDF):
        parent = np.ascontiguousarray(voc["parent"], np.int32)
        leaf = np.ascontiguousarray(voc["leaf"], np.uint8)
        desc = np.ascontiguousarray(voc["desc"], np.uint8)
        weight = np.ascontiguousarray(voc["weight"], np.float64)
        n = len(parent)
        if leaf.shape != (n,) or desc.shape != (n, 32) or weight.shape != (n,):
            raise ValueError("vocabulary arrays disagree on the number of nodes")
        check(_lib.lib().bowx_set_vocabulary(self._h, int(voc["k"]), int(voc["L"]), int(voc.get("scoring", scoring)), int(voc.get("weighting", weighting)),
                                             n, _p(parent), _p(leaf), _p(desc), _p(weight)))

    def loadFromTextFile(self, path):
        """TemplatedVocabulary.h:1333-1416 (the nodes the file lists; see synthetic.write_vocabulary_text on its last line)."""
        from . import synthetic
        self.set(synthetic.read_vocabulary_text(path))
        return True

    def _info(self):
        info = np.zeros(6, np.int32)
        check(_lib.lib().bowx_vocabulary_info(self._h, info.ctypes.data_as(C.POINTER(C.c_int32))))
        return info

    def getBranchingFactor(self):
        return int(self._info()[0])

    def getDepthLevels(self):
        return int(self._info()[1])

    def getScoringType(self):
        return int(self._info()[2])

    def getWeightingType(self):
        return int(self._info()[3])

    def size(self):
        return int(self._info()[5])

    def empty(self):
        return self.size() == 0

    def stopWords(self, minWeight):
        c = C.c_int32(0)
        check(_lib.lib().bowx_stop_words(self._h, float(minWeight), C.byref(c)))
        return c.value

    def getParentNode(self, wid, levelsup):
        node = C.c_uint32(0)
        check(_lib.lib().bowx_parent_node(self._h, int(wid), int(levelsup), C.byref(node)))
        return node.value

    def getWordWeight(self, wid):
        w = C.c_double(0)
        check(_lib.lib().bowx_word_weight(self._h, int(wid), C.byref(w)))
        return w.value

    # ---- transform
    def transform_features(self, desc, levelsup=0):
        """transform(feature, id, weight, nid, levelsup) for every row of desc [n][32]: (word ids, weights, node ids)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word, weight, node = np.zeros(n, np.uint32), np.zeros(n, np.float64), np.zeros(n, np.uint32)
        check(_lib.lib().bowx_transform_features(self._h, _p(desc), n, int(levelsup), _p(word), _p(weight), _p(node)))
        return word, weight, node

    def transform_word(self, feature):
        """WordId transform(const TDescriptor& feature) (:1049-1059)."""
        if self.empty():
            return 0
        return int(self.transform_features(np.asarray(feature, np.uint8).reshape(1, 32))[0][0])

    def transform_batch(self, desc, counts=None, levelsup=None):
        """desc [nframes][cap][32], counts [nframes] (default: all cap).  Returns per frame a BowVector ``(words, values)``; with
        ``levelsup`` a pair ``(BowVector, FeatureVector)`` with FeatureVector = ``(nodes, offsets, features)``."""
        desc = np.ascontiguousarray(desc, np.uint8)
        if desc.ndim != 3 or desc.shape[2] != 32:
            raise ValueError("desc must be [nframes][cap][32]")
        nframes, cap = desc.shape[:2]
        cap_ = max(cap, 1)
        if cap == 0:
            desc = np.zeros((nframes, 1, 32), np.uint8)
        counts = np.full(nframes, cap, np.int32) if counts is None else np.ascontiguousarray(counts, np.int32)
        words, vals, nbow = np.zeros((nframes, cap_), np.uint32), np.zeros((nframes, cap_), np.float64), np.zeros(nframes, np.int32)
        fv = levelsup is not None
        nodes = np.zeros((nframes, cap_), np.uint32) if fv else None
        offs = np.zeros((nframes, cap_ + 1), np.int32) if fv else None
        feats = np.zeros((nframes, cap_), np.uint32) if fv else None
        nfv = np.zeros(nframes, np.int32) if fv else None
        check(_lib.lib().bowx_transform_batch(self._h, _p(desc), _p(counts), nframes, cap_, int(levelsup or 0), _p(words), _p(vals), _p(nbow),
                                              _p(nodes) if fv else None, _p(offs) if fv else None, _p(feats) if fv else None,
                                              _p(nfv) if fv else None))
        out = []
        for f in range(nframes):
            bow = (words[f, :nbow[f]].copy(), vals[f, :nbow[f]].copy())
            if fv:
                g = nfv[f]
                out.append((bow, (nodes[f, :g].copy(), offs[f, :g + 1].copy(), feats[f, :offs[f, g]].copy())))
            else:
                out.append(bow)
        return out

    def transform(self, features, levelsup=None):
        """transform(features, v) / transform(features, v, fv, levelsup) for one frame's descriptors [n][32]."""
        features = np.ascontiguousarray(features, np.uint8).reshape(-1, 32)
        return self.transform_batch(features[None], None, levelsup)[0]

    def transform_batch_dev(self, d_desc, d_counts, nframes, cap, levelsup, d_bow_words, d_bow_vals, d_nbow, d_fv_nodes=0, d_fv_offsets=0,
                            d_fv_feats=0, d_nfv=0):
        """Device pointers in and out (bowx_transform_batch_dev), asynchronous on the handle's stream."""
        check(_lib.lib().bowx_transform_batch_dev(self._h, d_desc, d_counts, int(nframes), int(cap), int(levelsup), d_bow_words, d_bow_vals, d_nbow,
                                                  d_fv_nodes or None, d_fv_offsets or None, d_fv_feats or None, d_nfv or None))

    def transform_features_dev(self, d_desc, d_counts, nframes, cap, levelsup, d_word, d_weight, d_node):
        check(_lib.lib().bowx_transform_features_dev(self._h, d_desc, d_counts, int(nframes), int(cap), int(levelsup), d_word, d_weight, d_node))

    # ---- score
    def score(self, a, b):
        """score(v1, v2) of the vocabulary's scoring object."""
        w1, v1 = np.ascontiguousarray(a[0], np.uint32), np.ascontiguousarray(a[1], np.float64)
        w2, v2 = np.ascontiguousarray(b[0], np.uint32), np.ascontiguousarray(b[1], np.float64)
        s = C.c_double(0)
        check(_lib.lib().bowx_score(self._h, _p(w1), _p(v1), len(w1), _p(w2), _p(v2), len(w2), C.byref(s)))
        return s.value

    def score_batch(self, query, database):
        """score(query, entry) for every entry of ``database`` (a list of BowVectors): float64 [len(database)]."""
        qw, qv = np.ascontiguousarray(query[0], np.uint32), np.ascontiguousarray(query[1], np.float64)
        n = len(database)
        count = np.array([len(e[0]) for e in database], np.int32)
        start = np.zeros(n, np.int64)
        if n:
            start[1:] = np.cumsum(count[:-1], dtype=np.int64)
        words = np.concatenate([np.asarray(e[0], np.uint32) for e in database]) if n else np.zeros(0, np.uint32)
        vals = np.concatenate([np.asarray(e[1], np.float64) for e in database]) if n else np.zeros(0, np.float64)
        words, vals = np.ascontiguousarray(words), np.ascontiguousarray(vals)
        scores = np.zeros(n, np.float64)
        check(_lib.lib().bowx_score_batch(self._h, _p(qw), _p(qv), len(qw), _p(start), _p(count), _p(words), _p(vals), len(words), n, _p(scores)))
        return scores

    def score_batch_dev(self, d_qwords, d_qvals, nq, d_db_start, d_db_count, d_db_words, d_db_vals, nentries, d_scores):
        check(_lib.lib().bowx_score_batch_dev(self._h, d_qwords, d_qvals, int(nq), d_db_start, d_db_count, d_db_words, d_db_vals, int(nentries),
                                              d_scores))
