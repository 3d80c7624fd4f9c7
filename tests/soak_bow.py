#!/usr/bin/env python
"""Random soak of the bag-of-words kernels against oracle/bow_oracle.c (run by hand on a GPU box; not collected by pytest):
random tree shapes (k 1-40, L 1-6, ragged or complete), scoring / weighting types, frames with repeated words, stop words.
Usage: python tests/soak_bow.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from monocular_slam_b200 import Vocabulary  # noqa: E402
from monocular_slam_b200 import synthetic as syn  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    voc = Vocabulary()
    for c in range(cases):
        k = int(rng.choice([1, 2, 3, 5, 8, 10, 16, 17, 20, 33, 40]))
        L = int(rng.integers(1, 7))
        while k ** L > 60000:
            L -= 1
        ragged = bool(rng.random() < 0.5)
        va = syn.vocabulary(int(rng.integers(1 << 30)), k=k, L=L, ragged=ragged, stop_frac=float(rng.choice([0.0, 0.03, 0.5])))
        scoring, weighting = int(rng.integers(0, 6)), int(rng.integers(0, 4))
        ov = oracle.BowVocabulary(va, scoring, weighting)
        voc.set(va, scoring, weighting)
        if rng.random() < 0.3:
            mw = float(rng.uniform(0.5, 6))
            assert voc.stopWords(mw) == ov.stop_words(mw)
        nframes, cap = int(rng.integers(1, 6)), int(rng.choice([1, 7, 64, 300, 1100]))
        counts = rng.integers(0, cap + 1, nframes).astype(np.int32)
        desc = np.zeros((nframes, cap, 32), np.uint8)
        for f in range(nframes):
            if counts[f] and not va["leaf"].any():
                desc[f, :counts[f]] = syn.descriptors(int(rng.integers(1 << 30)), int(counts[f]))       # a vocabulary without words: empty()
            elif counts[f]:
                near = syn.vocabulary_features(int(rng.integers(1 << 30)), va, int(counts[f]), pool=int(rng.choice([3, 50, 100000])), max_flips=int(rng.integers(0, 60)))
                desc[f, :counts[f]] = near
        # the node level is defined by the reference only where every descent reaches it: ragged trees end from level 2 on
        lu = int(rng.integers(max(0, L - 2) if ragged else 0, L + 3))
        got = voc.transform_batch(desc, counts, lu)
        bows = []
        for f in range(nframes):
            w, v, nodes, offs, fe = ov.transform(desc[f, :counts[f]], lu)
            (gw, gv), (gn, go, gf) = got[f]
            ok = np.array_equal(w, gw) and np.array_equal(v, gv) and np.array_equal(nodes, gn) and np.array_equal(offs, go) and np.array_equal(fe, gf)
            if not ok:
                bad += 1
                print("case %d frame %d: transform MISMATCH (k=%d L=%d ragged=%s scoring=%d weighting=%d n=%d levelsup=%d)" % (c, f, k, L, ragged, scoring, weighting, counts[f], lu))
            bows.append((w, v))
        q = bows[int(rng.integers(nframes))]
        s = voc.score_batch(q, bows)
        for f in range(nframes):
            want = ov.score(q, bows[f])
            if not (s[f] == want or (scoring == oracle.KL and abs(s[f] - want) <= 1e-12 * max(1.0, abs(want))) or (np.isnan(s[f]) and np.isnan(want))):
                bad += 1
                print("case %d entry %d: score MISMATCH %r vs %r (scoring %d)" % (c, f, s[f], want, scoring))
    voc.close()
    print("soak_bow: %d cases, %d mismatches" % (cases, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
