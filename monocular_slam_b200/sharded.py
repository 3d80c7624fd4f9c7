"""Multi-GPU sharding of the ORB front-end: one process per GPU, torch.distributed for the plumbing.

* Extraction shards by FRAME and needs no collective: frames are independent (the reference loads every frame before
  processing, src/main.cpp:36-51, and FeatureExtractor::process touches only frames[frameIdx], FeatureExtractor.cpp:14).
  ``frame_block`` gives rank r a contiguous block so consecutive-frame matching stays local; the block's first frame is
  matched against the previous rank's last frame, which the rank simply extracts itself (one extra frame, no exchange).
* Large map-vs-frame / loop-closure matching shards the TRAIN set: every rank computes the per-query top-2 over its
  contiguous train range with global indices and the 16-byte candidates (nq x 16 B per rank) are exchanged and merged.
  Top-2 under the lexicographic (distance, index) order is associative and commutative, so the merged result is
  bit-identical to a single-device pass.  Two exchanges exist: ``p2p=True`` (product path on an NVLink box) -- the
  matching kernel stores its results straight into every rank's gather buffer over peer memory and a merge kernel waits
  on device-side flags, no NCCL call on the data path; ``p2p=False`` -- ONE NCCL all-gather followed by a merge kernel
  (the baseline the fused path is compared with, and the path the CPU/gloo tests exercise with injected stand-ins).
"""
import numpy as np


def bind_to_gpu_numa(device):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (NVML's ideal affinity), so that the pinned
    frame buffers it allocates afterwards are first-touched on that node and the H2D copies do not cross sockets.  With
    one process per GPU on a two-socket box this is what lets the end-to-end path scale.  Returns the CPU list, or None
    when NVML or the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def shard_bounds(n, world):
    """Contiguous, near-equal ranges: rank r owns [bounds[r], bounds[r+1])."""
    base, extra = divmod(int(n), int(world))
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def frame_block(nframes, rank, world):
    """Frames [lo, hi) extracted by ``rank`` and the frame its first one is matched against (``lo - 1``, or None)."""
    b = shard_bounds(nframes, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    return lo, hi, (lo - 1 if lo > 0 else None)


class ShardedSequence:
    """Frame-sharded extraction + consecutive-frame matching of ONE sequence over the ranks (BASELINE.json configs 2 and 3).

    Rank r owns the contiguous block ``frame_block(nframes, r, world)``.  Frames are independent, so there is no exchange:
    to match its first frame against the previous rank's last frame, the rank simply extracts that one extra frame
    itself (1 / block-length of redundant work instead of a 64 KB peer copy and a dependency between ranks).

    ``run(frames)`` walks the block in batches through the pipelined host path (``ORB.submit_batch`` / ``wait_batch``) and
    yields ``(frame_index, keypoints, descriptors, matches_against_previous_frame)`` in frame order for the rank's own
    frames.  With ``fundamental`` (a FundamentalFilter) the outlier filter runs in the same submissions and two more items are
    yielded per frame: the RANSAC status of the matches and the 8-point F (computeFundamentalMatrix,
    src/CameraPoseEstimator.cpp:545-586).  ``extract_match(frames, first_is_lead_in)`` can be injected to exercise the block
    logic without a GPU.
    """

    def __init__(self, orb=None, matcher=None, ratio=0.8, rank=0, world=1, batch=None, extract_match=None, fundamental=None,
                 max_distance=3.0, confidence=0.85):
        self.fundamental, self.max_distance, self.confidence = fundamental, float(max_distance), float(confidence)
        self.orb, self.matcher, self.ratio = orb, matcher, float(ratio)
        self.rank, self.world = int(rank), int(world)
        self.batch = int(batch or (orb.max_batch if orb is not None else 16))
        self._extract_match = extract_match
        if orb is None and extract_match is None:
            raise ValueError("ShardedSequence needs an ORB + BFMatcher (CUDA) or an injected extract_match")

    def block(self, nframes):
        return frame_block(nframes, self.rank, self.world)

    def run(self, frames):
        lo, hi, prev = self.block(len(frames))
        if hi <= lo:
            return
        start = lo if prev is None else prev          # lead-in frame: extracted, matched against nothing, not reported
        if self._extract_match is not None:
            chunks = [(b, min(b + self.batch, hi)) for b in range(start, hi, self.batch)]
            for b, e in chunks:
                for off, (k, d, m) in enumerate(self._extract_match(frames[b:e], b == start)):
                    if b + off >= lo:
                        yield b + off, k, d, m
            return
        import numpy as np
        from ._lib import DMATCH_DTYPE, KEYPOINT_DTYPE
        orb, cap, depth = self.orb, self.orb.default_cap, self.orb.pipeline_depth()
        orb.reset_sequence()
        pending = []

        def collect():
            b, n = pending.pop(0)
            res = orb.wait_batch()
            kps, desc, counts, good, ngood = res[:5]
            for i in range(n):
                f = b + i
                if f >= lo:
                    item = (f, kps[i, :counts[i]].copy(), desc[i, :counts[i]].copy(), good[i, :ngood[i]].copy())
                    if self.fundamental is not None:
                        item += (res[5][i, :ngood[i]].copy(), res[6][i].copy())
                    yield item

        for b in range(start, hi, self.batch):
            n = min(self.batch, hi - b)
            if len(pending) == depth:
                yield from collect()
            out = (np.zeros((n, cap), KEYPOINT_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32),
                   np.zeros((n, cap), DMATCH_DTYPE), np.zeros(n, np.int64))
            if self.fundamental is not None:
                out += (np.zeros((n, cap), np.uint8), np.zeros((n, 3, 3), np.float64), np.zeros(n, np.int32))
            orb.submit_batch(list(frames[b:b + n]), self.matcher, self.ratio, out, fundamental=self.fundamental,
                             max_distance=self.max_distance, confidence=self.confidence)
            pending.append((b, n))
        while pending:
            yield from collect()


class ShardedMatcher:
    """Train-sharded brute-force Hamming kNN(k=2) (BASELINE.json configs 4 and 5).

    ``local_top2(q, t_shard, offset) -> tensor[nq, 4] int32 (dist0, idx0, dist1, idx1)`` and
    ``merge(parts[world, nq, 4]) -> tensor[nq, 4]`` default to the CUDA kernels of a ``BFMatcher``; the host-side logic
    (ranges, offsets, the single all-gather) can be exercised on CPU by injecting stand-ins (tests do, over gloo).
    """

    def __init__(self, matcher=None, group=None, local_top2=None, merge=None, p2p=False, nq_max=0):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.matcher = matcher
        self._local = local_top2 or self._local_cuda
        self._merge = merge or self._merge_cuda
        if matcher is None and (local_top2 is None or merge is None):
            raise ValueError("ShardedMatcher needs a BFMatcher (CUDA); there is no CPU implementation in this package")
        self.p2p = bool(p2p)
        self.nq_max = int(nq_max)
        if self.p2p:
            self._setup_p2p()

    def _setup_p2p(self):
        """Export this rank's gather buffer, exchange the cudaIpc handles (plumbing: one all_gather_object), map the peers."""
        if self.matcher is None or self.nq_max < 1:
            raise ValueError("p2p=True needs a BFMatcher and nq_max >= 1")
        handle, _ = self.matcher.p2p_export(self.nq_max, self.world, self.rank)
        handles = [handle]
        if self.world > 1:
            handles = [None] * self.world
            self.dist.all_gather_object(handles, handle, group=self.group)
        self.matcher.p2p_import(handles)
        if self.world > 1:
            self.dist.barrier(group=self.group)     # nobody scatters before every rank has mapped every buffer

    def close(self):
        if self.p2p and self.matcher is not None:
            if self.world > 1:
                self.dist.barrier(group=self.group)  # peers may still be writing into this rank's buffer
            self.matcher.p2p_close()
            self.p2p = False

    # ---- CUDA implementations (hamx_knn2_dev / hamx_merge_top2_dev on the current torch stream)
    def _use_current_stream(self):
        import torch
        s = torch.cuda.current_stream().cuda_stream
        if s == 0:
            raise RuntimeError("run ShardedMatcher under a non-default torch stream (stream handle 0 selects the library's own stream)")
        self.matcher.set_stream(s)

    def _local_cuda(self, q, t_shard, offset):
        import torch
        self._use_current_stream()
        out = torch.empty((q.shape[0], 4), dtype=torch.int32, device=q.device)
        self.matcher.knn2_dev(q.data_ptr(), q.shape[0], t_shard.data_ptr(), t_shard.shape[0], int(offset), out.data_ptr())
        return out

    def _merge_cuda(self, parts):
        import torch
        self._use_current_stream()
        out = torch.empty((parts.shape[1], 4), dtype=torch.int32, device=parts.device)
        self.matcher.merge_top2_dev(parts.data_ptr(), parts.shape[0], parts.shape[1], out.data_ptr())
        return out

    # ---- the sharded operation
    def knn2(self, q, t_shard, train_offset):
        """q: [nq, 32] uint8, replicated on every rank; t_shard: this rank's rows [train_offset, train_offset + len)."""
        import torch
        if self.p2p:
            if q.shape[0] > self.nq_max:
                raise ValueError("%d queries exceed nq_max=%d of the peer-memory buffers" % (q.shape[0], self.nq_max))
            self._use_current_stream()
            out = torch.empty((q.shape[0], 4), dtype=torch.int32, device=q.device)
            self.matcher.knn2_p2p_dev(q.data_ptr(), q.shape[0], t_shard.data_ptr(), t_shard.shape[0], int(train_offset), out.data_ptr())
            return out
        local = self._local(q, t_shard, train_offset)
        if self.world == 1:
            return local
        flat = torch.empty((self.world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(flat, local.contiguous(), group=self.group)   # rank-major concatenation
        return self._merge(flat.view(self.world, local.shape[0], local.shape[1]))

    def knn2_from_full(self, q, t_full):
        """Convenience for tests: every rank holds the full train set and takes its own slice."""
        b = shard_bounds(t_full.shape[0], self.world)
        lo, hi = int(b[self.rank]), int(b[self.rank + 1])
        return self.knn2(q, t_full[lo:hi].contiguous(), lo)


class ShardedLoopScorer:
    """Loop-closure candidate scoring (LoopCloser::DetectLoop, src/LoopCloser.cpp:19-51) with the stored frames sharded BY
    FRAME over the ranks: every rank scores its own contiguous block of frames against the (replicated) descriptors of the
    current frame in one launch, the per-frame scores -- 4 bytes per stored frame -- are exchanged with one all-gather, and
    every rank takes the same arg-max (first frame with the strictly largest non-zero score).  The scores of different
    frames are independent, so the result equals a single-device pass over all frames.

    ``local_scores(q, frames, counts) -> int32 tensor [nlocal]`` and ``best(scores) -> (frame, score)`` default to the CUDA
    kernels of a ``BFMatcher``; stand-ins can be injected to exercise the exchange on CPU (gloo)."""

    def __init__(self, matcher=None, n=10, thr=40, group=None, local_scores=None, best=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.matcher, self.n, self.thr = matcher, int(n), int(thr)
        self._local = local_scores or self._local_cuda
        self._best = best or self._best_cuda
        if matcher is None and (local_scores is None or best is None):
            raise ValueError("ShardedLoopScorer needs a BFMatcher (CUDA); there is no CPU implementation in this package")

    def _stream(self):
        import torch
        s = torch.cuda.current_stream().cuda_stream
        if s == 0:
            raise RuntimeError("run ShardedLoopScorer under a non-default torch stream")
        self.matcher.set_stream(s)

    def _local_cuda(self, q, frames, counts):
        import torch
        self._stream()
        scores = torch.empty(max(frames.shape[0], 1), dtype=torch.int32, device=q.device)
        self.matcher.loop_score_dev(q.data_ptr(), q.shape[0], frames.data_ptr(), counts.data_ptr(), frames.shape[0], frames.shape[1],
                                    self.n, self.thr, scores.data_ptr())
        return scores[:frames.shape[0]]

    def _best_cuda(self, scores):
        import torch
        self._stream()
        out = torch.empty(2, dtype=torch.int32, device=scores.device)
        self.matcher.loop_best_dev(scores.data_ptr(), scores.shape[0], out.data_ptr())
        return out

    def score(self, q, frames_shard, counts_shard, nframes_total):
        """q [nq, 32] uint8 (replicated); frames_shard [nlocal, cap, 32] / counts_shard [nlocal]: this rank's block
        ``shard_bounds(nframes_total, world)``.  Returns (scores[nframes_total] int32 tensor, best[2] = {frame, score})."""
        import torch
        b = shard_bounds(nframes_total, self.world)
        lo, hi = int(b[self.rank]), int(b[self.rank + 1])
        if frames_shard.shape[0] != hi - lo:
            raise ValueError("rank %d owns frames [%d, %d) but was given %d" % (self.rank, lo, hi, frames_shard.shape[0]))
        local = self._local(q, frames_shard, counts_shard)
        if self.world == 1:
            return local, self._best(local)
        width = int((b[1:] - b[:-1]).max())
        padded = torch.zeros(width, dtype=torch.int32, device=local.device)
        padded[:hi - lo] = local
        if self.dist.get_backend(self.group) == "gloo" and padded.is_cuda:
            # plumbing fallback (several ranks on one GPU in the tests: NCCL refuses that): 4 bytes per frame through the host
            parts = [torch.empty(width, dtype=torch.int32) for _ in range(self.world)]
            self.dist.all_gather(parts, padded.cpu(), group=self.group)
            flat = torch.cat(parts).to(local.device)
        else:
            flat = torch.empty(self.world * width, dtype=torch.int32, device=local.device)
            self.dist.all_gather_into_tensor(flat, padded, group=self.group)
        scores = torch.cat([flat[r * width:r * width + int(b[r + 1] - b[r])] for r in range(self.world)])
        return scores, self._best(scores)


class ShardedBowDatabase:
    """score(query, entry) of the vendored DBoW2's scoring objects (ThirdParty/DBoW2/DBoW2/ScoringObject.cpp) against a
    database of bag-of-words vectors sharded BY ENTRY over the ranks: every rank scores its own contiguous block
    ``shard_bounds(nentries_total, world)`` in one launch (bowx_score_batch_dev), the 8-byte scores are exchanged with one
    all-gather, and every rank takes the same best entry (largest score, the first one on ties).  Entries are independent, so
    the scores equal a single-device pass.

    ``local_scores(qwords, qvals, start, count, words, vals) -> float64 tensor [nlocal]`` defaults to the CUDA kernel of a
    ``Vocabulary``; a stand-in can be injected to exercise the exchange on CPU (gloo)."""

    def __init__(self, vocabulary=None, group=None, local_scores=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.voc = vocabulary
        self._local = local_scores or self._local_cuda
        if vocabulary is None and local_scores is None:
            raise ValueError("ShardedBowDatabase needs a Vocabulary (CUDA); there is no CPU implementation in this package")

    def _local_cuda(self, qwords, qvals, start, count, words, vals):
        import torch
        s = torch.cuda.current_stream().cuda_stream
        if s == 0:
            raise RuntimeError("run ShardedBowDatabase under a non-default torch stream")
        self.voc.set_stream(s)
        n = count.shape[0]
        scores = torch.empty(max(n, 1), dtype=torch.float64, device=qwords.device)
        self.voc.score_batch_dev(qwords.data_ptr(), qvals.data_ptr(), qwords.shape[0], start.data_ptr(), count.data_ptr(), words.data_ptr(),
                                 vals.data_ptr(), n, scores.data_ptr())
        return scores[:n]

    def score(self, qwords, qvals, start_shard, count_shard, words, vals, nentries_total):
        """The query (replicated) against this rank's block of entries: entry e of the block is words / vals
        [start_shard[e], start_shard[e] + count_shard[e]).  Returns (scores[nentries_total] float64 tensor, best entry or -1)."""
        import torch
        b = shard_bounds(nentries_total, self.world)
        lo, hi = int(b[self.rank]), int(b[self.rank + 1])
        if count_shard.shape[0] != hi - lo:
            raise ValueError("rank %d owns entries [%d, %d) but was given %d" % (self.rank, lo, hi, count_shard.shape[0]))
        local = self._local(qwords, qvals, start_shard, count_shard, words, vals)
        if self.world == 1:
            scores = local
        else:
            width = int((b[1:] - b[:-1]).max())
            padded = torch.zeros(width, dtype=torch.float64, device=local.device)
            padded[:hi - lo] = local
            if self.dist.get_backend(self.group) == "gloo" and padded.is_cuda:
                parts = [torch.empty(width, dtype=torch.float64) for _ in range(self.world)]     # plumbing fallback, as ShardedLoopScorer
                self.dist.all_gather(parts, padded.cpu(), group=self.group)
                flat = torch.cat(parts).to(local.device)
            else:
                flat = torch.empty(self.world * width, dtype=torch.float64, device=local.device)
                self.dist.all_gather_into_tensor(flat, padded, group=self.group)
            scores = torch.cat([flat[r * width:r * width + int(b[r + 1] - b[r])] for r in range(self.world)])
        best = int(torch.argmax(scores)) if nentries_total else -1
        return scores, best
