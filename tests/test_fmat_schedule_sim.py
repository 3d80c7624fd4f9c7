"""CPU: the scheduling logic of the filter kernel (monocular_slam_b200/csrc/fmat.cu), checked by simulation against OpenCV's
sequential loop

    for (iter = 0; iter < niters; iter++)                      // budget checked once per iteration
        for each model k of iteration iter:
            if (count[iter][k] > max(maxGood, 6)) { best = (iter, k); maxGood = count; niters = min(niters, r(count)); }

(a) step 4 of a round: the warp-parallel evaluation (exclusive prefix maximum over candidates -> improvements, exclusive
    prefix minimum of their budgets over ITERATIONS -> how far the loop gets) must give the same best candidate, budget and
    iteration counter as the loop;
(b) step 3: warps take candidates in order and finish out of order; before taking one, a warp reads the best exact count
    completed so far (a candidate that cannot exceed it is abandoned: its count is recorded as 0) and the smallest budget implied
    by a completed candidate together with the iteration it came from (a candidate of a LATER iteration at or beyond that budget
    is skipped).  Whatever the interleaving, step 4 on the recorded counts must equal the loop on the true counts.

r(count) is any non-increasing function of the count (the kernel's is RANSACUpdateNumIters' closed form)."""
import numpy as np
import pytest

INT_MAX = 2 ** 31 - 1


def sequential(counts, nmodels, iter0, n0, floor0, r):
    """OpenCV's loop over one round: returns (best candidate index or -1, maxgood, niters, iterations executed)."""
    niters, maxgood, best, i = n0, floor0, -1, 0
    while i < len(nmodels) and iter0 + i < niters:
        for k in range(nmodels[i]):
            c = counts[i][k]
            if c > maxgood:
                maxgood, best = c, 3 * i + k
                niters = min(niters, r(c))
        i += 1
    return best, maxgood, niters, i


def warp_step4(counts, nmodels, iter0, n0, floor0, r):
    """Step 4 as the kernel does it: 32 lanes x 4 iterations x 3 candidates, two warp scans."""
    chunk = len(nmodels)
    cnt = np.zeros((32, 12), np.int64)
    for lane in range(32):
        for j in range(4):
            i = lane * 4 + j
            for k in range(3):
                cnt[lane, 3 * j + k] = counts[i][k] if i < chunk and k < nmodels[i] else 0
    lmax = cnt.max(1)
    excl = np.concatenate([[0], np.maximum.accumulate(lmax)[:-1]])
    lim = np.full((32, 12), INT_MAX, np.int64)
    lmin = np.full(32, INT_MAX, np.int64)
    for lane in range(32):
        run = max(int(excl[lane]), floor0)
        for e in range(12):
            if cnt[lane, e] > run:
                run = int(cnt[lane, e])
                lim[lane, e] = min(n0, r(run))
                lmin[lane] = min(lmin[lane], lim[lane, e])
            else:
                cnt[lane, e] = -1
    pexcl = np.concatenate([[INT_MAX], np.minimum.accumulate(lmin)[:-1]])
    last, last_cnt, last_n = -1, 0, 0
    for lane in range(32):
        P = min(int(pexcl[lane]), n0)
        Q = P
        for e in range(12):
            if e % 3 == 0:
                P = Q
            if cnt[lane, e] >= 0:
                if iter0 + lane * 4 + e // 3 < P:
                    last, last_cnt, last_n = lane * 12 + e, int(cnt[lane, e]), min(Q, int(lim[lane, e]))
                Q = min(Q, int(lim[lane, e]))
    nf = last_n if last >= 0 else n0
    done = min(chunk, max(last // 3 + 1 if last >= 0 else 0, nf - iter0))
    return last, (last_cnt if last >= 0 else floor0), nf, done


def make_round(rng, n=1000):
    chunk = int(rng.integers(1, 129))
    nmodels = rng.choice([0, 1, 1, 3, 3, 2], chunk)
    style = rng.integers(0, 3)
    counts = []
    for i in range(chunk):
        if style == 0:
            row = rng.integers(0, 40, 3)                           # many ties, small counts
        elif style == 1:
            row = rng.integers(0, n + 1, 3)
        else:
            row = np.where(rng.random(3) < 0.1, rng.integers(n // 2, n + 1, 3), rng.integers(0, 60, 3))
        counts.append([int(v) for v in row])
    return counts, [int(v) for v in nmodels]


def budget(n, scale):
    return lambda c: max(0, int(scale * (n - c) ** 2 / n))          # non-increasing in the count, 0 when every point is an inlier


@pytest.mark.parametrize("seed", range(6))
def test_warp_evaluation_of_the_update_rule_equals_the_loop(seed):
    rng = np.random.default_rng(seed)
    for _ in range(400):
        counts, nmodels = make_round(rng)
        iter0 = int(rng.integers(0, 900))
        n0 = iter0 + len(nmodels) + int(rng.integers(0, 200))       # the round never exceeds the remaining budget
        floor0 = int(rng.choice([6, 6, 30, 500]))
        r = budget(1000, float(rng.choice([0.001, 0.01, 0.1, 1.0])))
        want = sequential(counts, nmodels, iter0, n0, floor0, r)
        got = warp_step4(counts, nmodels, iter0, n0, floor0, r)
        assert got == want, (counts, nmodels, iter0, n0, floor0)


@pytest.mark.parametrize("seed", range(6))
def test_out_of_order_scoring_with_running_bound_and_budget(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(300):
        counts, nmodels = make_round(rng)
        chunk = len(nmodels)
        iter0 = int(rng.integers(0, 900))
        n0 = iter0 + chunk + int(rng.integers(0, 200))
        floor0 = int(rng.choice([6, 6, 30, 500]))
        r = budget(1000, float(rng.choice([0.001, 0.01, 0.1, 1.0])))
        cands = [(i, k) for i in range(chunk) for k in range(3)]       # handed out in this order; k >= nmodels[i] are empty slots
        nwarps = int(rng.integers(1, 9))
        run_max, limit = floor0, (n0, 255)                             # (budget, iteration it came from)
        recorded = [[0, 0, 0] for _ in range(chunk)]
        busy = []                                                      # (finish time, i, k, bound read when taken)
        t, nxt = 0.0, 0
        while nxt < len(cands) or busy:
            if nxt < len(cands) and len(busy) < nwarps:
                i, k = cands[nxt]
                nxt += 1
                if k >= nmodels[i]:
                    continue
                bound, (lim, src) = run_max, limit                     # read BEFORE the candidate is taken
                if src < i and iter0 + i >= lim:
                    continue                                           # skipped: recorded count stays 0
                busy.append((t + float(rng.exponential(1.0)), i, k, bound))
                continue
            busy.sort()
            t, i, k, bound = busy.pop(0)
            c = counts[i][k]
            if c <= bound:
                continue                                               # abandoned: recorded count stays 0
            recorded[i][k] = c
            if c > run_max:
                run_max = c
                limit = min(limit, (min(n0, r(c)), i))
        want = sequential(counts, nmodels, iter0, n0, floor0, r)
        got = sequential(recorded, nmodels, iter0, n0, floor0, r)
        assert got == want, (counts, nmodels, iter0, n0, floor0, nwarps)
        assert warp_step4(recorded, nmodels, iter0, n0, floor0, r) == want
