#!/usr/bin/env python
"""bench.py -- ORB extract+match frames/s @1080p / 2000 kp (BASELINE.json metric), plus Hamming Gcmp/s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference [...]                          # the reference's CPU path on the host cores

A "step" is one pass of the hot path over one batch of synthetic frames: ORB extraction (pyramid, FAST+NMS, retention,
Harris, orientation, rBRIEF) of every frame and matchFeatures(frame i, frame i-1, ratio 0.75) for consecutive frames.
`value` is timed with the frames resident in HBM; `e2e` goes through the host-buffer C ABI (pinned host frames in,
keypoints / descriptors / matches back in host memory) with the copies inside the timed region.  Under torchrun every
rank extracts and matches its own block of frames (frame sharding, no collective: weak scaling); the secondary Hamming
measurement shards the train set over the ranks and merges per-query top-2 after one all-gather.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ORB extract+match frames/s @1080p 2k kp; Hamming Gcmp/s at 1/2/4/8 GPUs"
W, H, NFEAT, RATIO = 1920, 1080, 2000, 0.75
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml, else nvidia-smi)."""

    def __init__(self, index, period=0.002):
        self.index, self.samples, self.reasons, self.max_mhz, self.period = index, [], set(), None, period
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "note": "no samples (nvml unavailable)"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ shared config
def level_pixels():
    """Pixels of all 8 pyramid levels of a W x H frame (cv::ORB's level sizes: round(W / scale_l), float32 scale)."""
    px = 0
    for l in range(8):
        sc = np.float32(np.float64(np.float32(1.2)) ** l)
        px += int(np.rint(np.float32(W) / sc)) * int(np.rint(np.float32(H) / sc))
    return px


def workload_config(world, B):
    """`config` of the JSON line, identical for both arms (the reference arm processes the same frames on the host)."""
    level_px = level_pixels()
    cfg = {"workload": "synthetic %dx%d monocular sequence, %d kp/frame, consecutive-frame matching, ratio 0.75 (BASELINE.json configs[%d])"
                       % (W, H, NFEAT, 1 if (W, H, NFEAT) == (1920, 1080, 2000) else 2),
           "frame_w": W, "frame_h": H, "nfeatures": NFEAT, "nlevels": 8, "scale_factor": 1.2, "score_type": "HARRIS",
           "batch_frames_per_gpu": B, "parallelism": "frames sharded over %d GPU(s), no collective" % world,
           "frames_seed": "synthetic.sequence(B, W, H, seed=100 + rank)"}
    if True:
        cfg["l2"] = ("inputs larger than L2: %d frames x %.1f MB = %.0f MB of frames (+%.0f MB of pyramid levels) per step vs 126 MB L2"
                     % (B, W * H / 1e6, B * W * H / 1e6, B * (level_px - W * H) / 1e6))
    return cfg


# ------------------------------------------------------------------------------------------------ CPU reference path
_W = {}


def _cpu_worker_init(frames, lead_in, nfeat, use_cv2, threads=1):
    """Runs in a worker process: this worker's frames of the step's sequence (+ the frame before them, to match the
    first one against) and its extractor."""
    _W["frames"], _W["lead_in"], _W["use_cv2"] = frames, lead_in, use_cv2
    if use_cv2:
        import cv2
        cv2.setNumThreads(threads)
        _W["cv2"] = cv2
        _W["orb"] = cv2.ORB_create(nfeatures=nfeat)
        _W["bf"] = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    else:
        import oracle
        _W["oracle"] = oracle
        _W["params"] = oracle.Params(nfeatures=nfeat)
    _W["prev"] = None
    if lead_in is not None:                     # its descriptors are what the previous worker computes for its last frame
        _W["prev"] = _cpu_extract(lead_in)[1]
    _W["prev0"] = _W["prev"]
    return True


def _cpu_extract(img):
    """The reference's call order: detect, compute (src/FeatureExtractor.cpp:17,19)."""
    if _W["use_cv2"]:
        kp = _W["orb"].detect(img, None)
        return _W["orb"].compute(img, kp)
    return _W["oracle"].detect_and_compute(img, _W["params"])


def _cpu_worker_step(_):
    """One bounded sample: extract every frame of the worker's block and match it against its predecessor, exactly in
    the reference's call order (knnMatch k=2 + ratio: src/CameraPoseEstimator.cpp:200-213)."""
    t0 = time.perf_counter()
    prev, nkp, nmatch, npairs = _W["prev0"], 0, 0, 0
    for img in _W["frames"]:
        kp, des = _cpu_extract(img)
        nkp += len(kp)
        if prev is not None and des is not None:
            if _W["use_cv2"]:
                raw = _W["bf"].knnMatch(des, prev, 2)
                good = [r[0] for r in raw if len(r) == 2 and r[0].distance < r[1].distance * np.float32(RATIO)]
                nmatch += len(good)
            else:
                nmatch += len(_W["oracle"].match_features(des, prev, RATIO)[0])
            npairs += 1
        prev = des
    return len(_W["frames"]), time.perf_counter() - t0, nkp, nmatch, npairs


class CpuReference:
    """The reference's CPU extract+match path on all host cores: one process per core, each running OpenCV's ORB /
    BFMatcher single-threaded (cv2 is the library the reference calls; if it cannot be imported the C oracle port is
    timed instead and `kind` says "port").  A step is the SAME batch of frames the GPU arm's rank 0 processes
    (synthetic.sequence(B, W, H, seed=100)), dealt to the workers in contiguous blocks."""

    def __init__(self, nframes, max_workers=128, threads=1, workers=None):
        import multiprocessing as mp
        from monocular_slam_b200 import synthetic as syn
        try:
            import cv2  # noqa: F401
            self.use_cv2 = True
        except Exception:
            self.use_cv2 = False
            import oracle
            oracle.build()
        try:
            ncpu = len(os.sched_getaffinity(0))
        except Exception:
            ncpu = os.cpu_count() or 1
        self.workers = workers or max(1, min(ncpu, max_workers, nframes))
        self.threads = threads
        self.nframes = nframes
        seq = syn.sequence(nframes, W, H, seed=100)
        bounds = np.linspace(0, nframes, self.workers + 1).astype(int)
        ctx = mp.get_context("spawn")
        self.pools = []
        # one single-process pool per worker so that each keeps its own frames and every step reaches every worker
        for i in range(self.workers):
            lo, hi = int(bounds[i]), int(bounds[i + 1])
            self.pools.append(ctx.Pool(1, initializer=_cpu_worker_init,
                                       initargs=(seq[lo:hi].copy(), seq[lo - 1].copy() if lo > 0 else None, NFEAT, self.use_cv2, threads)))
        self.stats = None
        self.step()   # first touch: imports, page-in, first matches

    def step(self):
        t0 = time.perf_counter()
        res = [p.apply_async(_cpu_worker_step, (0,)) for p in self.pools]
        res = [r.get() for r in res]
        n = sum(r[0] for r in res)
        self.stats = {"keypoints_per_frame": sum(r[2] for r in res) / max(n, 1),
                      "matches_per_frame": sum(r[3] for r in res) / max(sum(r[4] for r in res), 1)}
        return n, time.perf_counter() - t0

    def close(self):
        for p in self.pools:
            p.terminate()

    def describe(self):
        return {"kind": "reference" if self.use_cv2 else "port", "cores": self.workers * self.threads,
                "sample": "%d frames of %dx%d per step, %d worker process(es) x %d thread(s), %s; detect+compute+knnMatch(k=2)+ratio %.2f against the previous frame"
                          % (self.nframes, W, H, self.workers, self.threads,
                             "cv2 %s ORB/BFMatcher" % __import__("cv2").__version__ if self.use_cv2 else "oracle/orb_oracle.c", RATIO)}


def cpu_thread_settings(nframes=2):
    """the same calls with cv2.setNumThreads(1) and cv2.setNumThreads(all cores) in ONE process
    (frames/s each), reported beside the process pool."""
    out = {}
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    for name, thr in (("one_process_1_thread", 1), ("one_process_all_threads", ncpu)):
        try:
            ref = CpuReference(nframes, threads=thr, workers=1)
            best = min(ref.step()[1] for _ in range(2))
            ref.close()
            out[name] = nframes / best
        except Exception:
            out[name] = None
    out["threads_all"] = ncpu
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    ref = CpuReference(args.batch)
    for _ in range(max(args.warmup - 1, 0)):
        ref.step()
    frames, elapsed = 0, 0.0
    for _ in range(args.steps):
        n, dt = ref.step()
        frames += n
        elapsed += dt
    ref.close()
    value = frames / elapsed
    d = ref.describe()
    d["value"] = value
    d["unit"] = "frames/s"
    cfg = workload_config(args.gpus, args.batch)
    cfg.update(ref.stats)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
            "cpu_baseline": d,
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU path
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def run_cfg3(torch, syn, ORB, matcher, dev, local_rank, rank, world, stream, timed, max_over_ranks, barrier, args):
    """BASELINE.json configs[2]: 3840x2160 frames, 8000 keypoints per frame, frames sharded over the ranks (every rank its own
    block of 16 frames, consecutive-frame matching inside the block).  Device-resident and end-to-end (pipelined host path)
    frames/s, a few steps each."""
    from monocular_slam_b200 import DMATCH_DTYPE, KEYPOINT_DTYPE
    w3, h3, nf3, b3, steps3 = 3840, 2160, 8000, 16, 5
    seq3 = syn.sequence(b3, w3, h3, seed=500 + rank)
    h_fr = torch.from_numpy(seq3).pin_memory()
    d_fr = h_fr.to(dev, non_blocking=True)
    orb3 = ORB(nfeatures=nf3, max_size=(w3, h3), max_batch=b3, device=local_rank)
    orb3.set_stream(stream.cuda_stream)
    cap = orb3.default_cap
    d_kps = torch.empty((b3, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.empty((b3, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(b3, dtype=torch.int32, device=dev)
    d_good = torch.empty((b3, cap, 4), dtype=torch.int32, device=dev)
    d_ngood = torch.zeros(b3, dtype=torch.int64, device=dev)

    def step_dev():
        orb3.extract_batch_dev(d_fr.data_ptr(), w3 * h3, b3, w3, h3, w3, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
        matcher.match_consecutive_dev(d_desc.data_ptr(), d_cnt.data_ptr(), b3, cap, 0, 0, RATIO, d_good.data_ptr(), d_ngood.data_ptr())
    ms, _ = timed(step_dev, steps3, 3)
    orb3.check_dev()
    ms = max_over_ranks(ms)
    counts, ngood = d_cnt.cpu().numpy(), d_ngood.cpu().numpy()
    depth = orb3.pipeline_depth()
    frames_np = [h_fr[i].numpy() for i in range(b3)]
    outs = [(torch.empty((b3, cap, 7), dtype=torch.float32).pin_memory().numpy().view(KEYPOINT_DTYPE).reshape(b3, cap),
             torch.empty((b3, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(b3, np.int32),
             torch.empty((b3, cap, 4), dtype=torch.int32).pin_memory().numpy().view(DMATCH_DTYPE).reshape(b3, cap),
             np.zeros(b3, np.int64)) for _ in range(depth)]
    k = [0]

    def step_pipe():
        if orb3.batches_in_flight() == depth:
            orb3.wait_batch()
        orb3.submit_batch(frames_np, matcher, RATIO, outs[k[0] % depth])
        k[0] += 1

    def drain():
        last = None
        while orb3.batches_in_flight():
            last = orb3.wait_batch()
        return last
    orb3.reset_sequence()
    for _ in range(3):
        step_pipe()
    drain()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps3):
        step_pipe()
    last = drain()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    assert np.array_equal(last[2], counts), "config 3: host and device paths disagree on keypoint counts"
    assert np.array_equal(last[4][1:], ngood[1:]), "config 3: host and device paths disagree on match counts"
    orb3.close()
    return {"workload": "synthetic %dx%d sequence, %d kp/frame, %d frames per GPU and step, consecutive-frame matching (BASELINE.json configs[2])"
                        % (w3, h3, nf3, b3),
            "value": world * b3 * steps3 / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms / steps3, "steps": steps3,
            "keypoints_per_frame": float(counts.mean()), "matches_per_frame": float(ngood[1:].mean()),
            "e2e": {"value": world * b3 * steps3 / (e2e_ms * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms / steps3,
                    "h2d_bytes_per_step": b3 * w3 * h3, "d2h_bytes_per_step": b3 * cap * 76 + b3 * 8 + b3 * 200}}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from monocular_slam_b200 import ORB, BFMatcher, DMATCH_DTYPE, KEYPOINT_DTYPE, popc_peak
    from monocular_slam_b200 import synthetic as syn
    from monocular_slam_b200.sharded import ShardedMatcher, shard_bounds

    from monocular_slam_b200.sharded import bind_to_gpu_numa
    numa_cpus = None
    if world > 1 and os.environ.get("ORBX_NUMA_BIND", "1")[:1] != "0":
        numa_cpus = bind_to_gpu_numa(local_rank)     # before any pinned allocation: host frames live next to their GPU
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    # every rank owns its own block of the sequence (frame sharding): B consecutive frames, seeded by rank
    seq = syn.sequence(B, W, H, seed=100 + rank)
    if args.wc_frames:
        # write-combined page-locked frames (orbx_host_alloc_wc): the host only writes them
        import ctypes as _C
        from monocular_slam_b200 import _lib as _LL
        _p = _C.c_void_p()
        _LL.check(_LL.lib().orbx_host_alloc_wc(seq.nbytes, _C.byref(_p)))
        _wc = np.frombuffer((_C.c_uint8 * seq.nbytes).from_address(_p.value), np.uint8).reshape(seq.shape)
        np.copyto(_wc, seq)
        h_frames = torch.from_numpy(_wc)
    else:
        h_frames = torch.from_numpy(seq).pin_memory()
    d_frames = h_frames.to(dev, non_blocking=True)
    orb = ORB(nfeatures=NFEAT, max_size=(W, H), max_batch=B, device=local_rank)
    matcher = BFMatcher(device=local_rank)
    orb.set_stream(stream.cuda_stream)
    matcher.set_stream(stream.cuda_stream)
    cap = orb.default_cap
    d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    d_good = torch.empty((B, cap, 4), dtype=torch.int32, device=dev)
    d_ngood = torch.zeros(B, dtype=torch.int64, device=dev)
    d_prev = torch.zeros((cap, 32), dtype=torch.uint8, device=dev)
    d_prevn = torch.zeros(1, dtype=torch.int32, device=dev)
    state = {"have_prev": False}

    def step_device():
        orb.extract_batch_dev(d_frames.data_ptr(), W * H, B, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
        matcher.match_consecutive_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, cap,
                                      d_prev.data_ptr() if state["have_prev"] else 0, d_prevn.data_ptr() if state["have_prev"] else 0,
                                      RATIO, d_good.data_ptr(), d_ngood.data_ptr())
        d_prev.copy_(d_desc[B - 1], non_blocking=True)
        d_prevn.copy_(d_cnt[B - 1:B], non_blocking=True)
        state["have_prev"] = True

    # host-buffer path: pinned buffers for everything that crosses PCIe
    h_kps = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
    h_cnt = np.zeros(B, np.int32)
    h_good = torch.empty((B, cap, 4), dtype=torch.int32).pin_memory()
    h_ngood = np.zeros(B, np.int64)
    kps_np = h_kps.numpy().view(KEYPOINT_DTYPE).reshape(B, cap)
    desc_np = h_desc.numpy()
    good_np = h_good.numpy().view(DMATCH_DTYPE).reshape(B, cap)
    frames_np = [h_frames[i].numpy() for i in range(B)]

    def step_host():
        orb.extract_batch(frames_np, cap=cap, out=(kps_np, desc_np, h_cnt))
        orb.match_consecutive(matcher, RATIO, cap, B, out=(good_np, h_ngood))

    def timed(step, steps, warmup, sampler=None):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.__enter__()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if sampler:
            sampler.__exit__()
        barrier()
        return e0.elapsed_time(e1), wall * 1e3

    # ---- headline: device-resident
    sampler = ClockSampler(local_rank)
    dev_ms, _ = timed(step_device, args.steps, args.warmup, sampler)
    orb.check_dev()
    # per-stage device times: the same steps again with the library's stage events on
    orb.set_profiling(True)
    step_device()
    orb.read_profile()     # drop the first batch
    prof_ms, _ = timed(step_device, args.steps, 0)
    stages, nb = orb.read_profile()
    orb.set_profiling(False)
    orb.check_dev()
    dev_ms = max_over_ranks(dev_ms)
    counts = d_cnt.cpu().numpy()
    ngood_dev = d_ngood.cpu().numpy()
    value = world * B * args.steps / (dev_ms * 1e-3)

    # matching kernels alone (same stream, CUDA events)
    def step_match():
        matcher.match_consecutive_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, cap, d_prev.data_ptr(), d_prevn.data_ptr(), RATIO,
                                      d_good.data_ptr(), d_ngood.data_ptr())
    match_ms, _ = timed(step_match, args.steps, 2)
    match_ms /= args.steps

    # ---- end to end through host buffers: pinned host frames in, keypoints / descriptors / matches back in pinned host
    # memory.  (a) pipelined: orbx_submit_batch / orbx_wait_batch, orbx_pipeline_depth() batches in flight, every step still uploads its own
    # frames and downloads its own results inside the timed region; (b) blocking: orbx_extract_batch + orbx_match_consecutive.
    def pinned_out():
        return (torch.empty((B, cap, 7), dtype=torch.float32).pin_memory().numpy().view(KEYPOINT_DTYPE).reshape(B, cap),
                torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(B, np.int32),
                torch.empty((B, cap, 4), dtype=torch.int32).pin_memory().numpy().view(DMATCH_DTYPE).reshape(B, cap),
                np.zeros(B, np.int64))
    depth = orb.pipeline_depth()
    outs = [pinned_out() for _ in range(depth)]
    pstate = {"k": 0, "last": None, "outs": outs, "fm": None, "back": 0}

    def step_pipe():
        if orb.batches_in_flight() == depth:
            pstate["last"] = orb.wait_batch()
        orb.submit_batch(frames_np, matcher, RATIO, pstate["outs"][pstate["k"] % depth], fundamental=pstate["fm"], back=pstate["back"])
        pstate["k"] += 1

    def drain():
        while orb.batches_in_flight():
            pstate["last"] = orb.wait_batch()

    def timed_pipe(steps, warmup):
        for _ in range(warmup):
            step_pipe()
        drain()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_pipe()
        drain()                      # every timed step's results are in host memory before the clock stops
        wall = time.perf_counter() - t0
        barrier()
        return wall * 1e3

    orb.reset_sequence()
    e2e_ms = max_over_ranks(timed_pipe(args.steps, max(args.warmup, 2)))
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    p_cnt, p_ngood = pstate["last"][2].copy(), pstate["last"][4].copy()
    orb.reset_sequence()
    _, blk_wall_ms = timed(step_host, args.steps, max(args.warmup, 1))
    blk_ms = max_over_ranks(blk_wall_ms)
    blk_value = world * B * args.steps / (blk_ms * 1e-3)
    assert np.array_equal(p_cnt, h_cnt), "pipelined and blocking host paths disagree on keypoint counts"
    assert np.array_equal(p_ngood[1:], h_ngood[1:]), "pipelined and blocking host paths disagree on match counts"
    h2d = B * W * H
    d2h = B * cap * (28 + 32 + 16) + B * 8 + B * 200   # keypoint, descriptor and match arrays [B][cap], match counts, the leading 200 B of the per-frame counters
    # parity of the two paths inside the bench: same counts, same number of accepted matches
    assert np.array_equal(h_cnt, counts), "host and device paths disagree on keypoint counts"
    assert np.array_equal(h_ngood[1:], ngood_dev[1:]), "host and device paths disagree on match counts"

    # ---- the floor under the end-to-end step: the same bytes as plain copies, timed in the same run with every rank copying
    # at once (pinned host <-> device, one cudaMemcpyAsync each way per step, CUDA events on this rank's stream)
    d2h_dev = torch.empty(d2h, dtype=torch.uint8, device=dev)
    d2h_host = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    d_floor = torch.empty_like(d_frames)

    def step_h2d():
        d_floor.copy_(h_frames, non_blocking=True)

    def step_d2h():
        d2h_host.copy_(d2h_dev, non_blocking=True)
    h2d_floor_ms = max_over_ranks(timed(step_h2d, args.steps, 2)[0]) / args.steps
    d2h_floor_ms = max_over_ranks(timed(step_d2h, args.steps, 2)[0]) / args.steps
    # both directions at once, as the pipelined path runs them (upload of step k+1 beside the download of step k-1): the
    # download on a side stream that `timed`'s events wait for through a join at the end of every step
    side = torch.cuda.Stream(device=dev)

    def step_duplex():
        with torch.cuda.stream(side):
            d2h_host.copy_(d2h_dev, non_blocking=True)
        d_floor.copy_(h_frames, non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)
    duplex_floor_ms = max_over_ranks(timed(step_duplex, args.steps, 2)[0]) / args.steps
    del d_floor, d2h_dev, d2h_host, side

    # ---- sustained: >= 2 s of back-to-back device-resident steps (the headline above is a 50 ms burst)
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(2.2e3 / max(dev_ms / args.steps, 1e-3)))
        sus_sampler = ClockSampler(local_rank, period=0.01)
        sus_ms, _ = timed(step_device, n_sus, 0, sus_sampler)
        orb.check_dev()
        sus_ms = max_over_ranks(sus_ms)
        sustained = {"value": world * B * n_sus / (sus_ms * 1e-3), "unit": "frames/s", "steps": n_sus, "seconds": sus_ms * 1e-3,
                     "ms_per_step": sus_ms / n_sus, "clocks": sus_sampler.summary()}

    # ---- roofline of the dominant kernel (FAST): algorithmic bytes = every level pixel read once
    ws, hs, _, _ = orb.level_info(W, H)
    level_px = int((ws.astype(np.int64) * hs).sum())
    peak, peak_src = hbm_peak()
    fast_ms = stages["fast"]
    achieved = level_px * B / (fast_ms * 1e-3) / 1e9
    traffic, issue = None, None
    tp = os.path.join(ROOT, "profiles", "fast_traffic.json")
    if os.path.exists(tp) and (W, H, NFEAT, B) == (1920, 1080, 2000, 64):
        try:
            prof = json.load(open(tp))
            traffic = prof.get("dram_bytes_per_launch")
            winst = prof.get("warp_instructions_per_launch")
            if winst:
                # what actually bounds k_fast: warp instructions per launch (ncu smsp__inst_executed.sum of the same launch,
                # profiles/) over the live launch time, against 4 schedulers x 1 warp instruction per clock per SM
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
                issue = {"warp_instructions_per_launch": winst, "achieved_gwinst_s": winst / (stages["fast"] * 1e-3) / 1e9,
                         "peak_gwinst_s": sms * 4 * clk / 1e9, "frac": winst / (stages["fast"] * 1e-3) / (sms * 4 * clk)}
        except Exception:
            traffic, issue = None, None
    roofline = {"kernel": "k_fast", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "launch_ms": fast_ms,
                "algorithmic_bytes_per_launch": level_px * B,
                "issue_frac": issue["frac"] if issue else None,
                "issue": issue,
                "note": "FAST reads each pyramid-level byte once (3.096 x W x H per frame); it is bound by instruction issue "
                        "(about 80 instructions per pixel on the integer pipes), not by HBM: see roofline.issue"}

    # the HBM-bound stage of the path, for the same roofline arithmetic: the pyramid reads levels 0..6 and writes levels 1..7
    pyr_bytes = int(B * (2 * level_px - W * H - int(ws[-1]) * int(hs[-1])))
    roofline_pyramid = {"kernel": "k_pyr_down (7 launches)", "bound": "hbm", "achieved": pyr_bytes / (stages["pyramid"] * 1e-3) / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": pyr_bytes / (stages["pyramid"] * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes_per_step": pyr_bytes, "ms_per_step": stages["pyramid"],
                        "note": "bit-exact INTER_LINEAR_EXACT chain; ncu shows it bound by L1 / shared-memory throughput (70 %), not DRAM (14 %)"}

    # ---- the same step on frames with camera-image statistics (synthetic.natural_frame: 0.25 % FAST corners instead of the
    # 18 % of the Appendix-B stress generator the headline uses)
    natural = None
    if not args.no_natural:
        nat = torch.from_numpy(syn.sequence(B, W, H, seed=300 + rank, generator=syn.natural_frame)).pin_memory()
        d_keep = d_frames
        d_frames = nat.to(dev, non_blocking=True)
        state["have_prev"] = False
        nat_ms, _ = timed(step_device, args.steps, args.warmup)
        orb.check_dev()
        nat_ms = max_over_ranks(nat_ms)
        orb.set_profiling(True)
        step_device()
        orb.read_profile()
        timed(step_device, max(args.steps // 2, 2), 0)
        nat_stages, _ = orb.read_profile()
        orb.set_profiling(False)
        orb.check_dev()
        nat_counts, nat_good = d_cnt.cpu().numpy(), d_ngood.cpu().numpy()
        natural = {"value": world * B * args.steps / (nat_ms * 1e-3), "unit": "frames/s", "ms_per_step": nat_ms / args.steps,
                   "generator": "synthetic.natural_frame (1/f shading, flat objects, textured patches, lens blur, sensor noise)",
                   "stages_ms_per_step": nat_stages, "keypoints_per_frame": float(nat_counts.mean()),
                   "matches_per_frame": float(nat_good[1:].mean())}
        d_frames = d_keep
        state["have_prev"] = False
        del nat

    # ---- BASELINE.json configs[2]: 3840x2160 frames, 8000 keypoints, frame-sharded over the ranks (a short leg)
    cfg3 = None
    if not args.no_cfg3 and (W, H, NFEAT) == (1920, 1080, 2000):
        cfg3 = run_cfg3(torch, syn, ORB, matcher, dev, local_rank, rank, world, stream, timed, max_over_ranks, barrier, args)

    # ---- the reference's unmodified call-by-call loop, one frame at a time through the drop-in calls: detect, compute
    # (src/FeatureExtractor.cpp:17,19), then matchFeatures + computeFundamentalMatrix against the 5 previous frames
    # (src/CameraPoseEstimator.cpp:405-419); host wall clock per frame, results in host memory after every call
    single = None
    if not args.no_single:
        from monocular_slam_b200 import FundamentalFilter
        orb1 = ORB(nfeatures=NFEAT, max_size=(W, H), max_batch=1, device=local_rank)
        fm1 = FundamentalFilter(device=local_rank)
        hist = []
        nfr = min(B, 24)

        def one_frame(img):
            kp = orb1.detect(img)
            kp, des = orb1.compute(img, kp)
            for pk, pd in hist[-5:][::-1]:
                good = matcher.match_ratio(des, pd, 0.8)
                fm1.compute_fundamental(kp, pk, good)
            hist.append((kp, des))
        for i in range(6):
            one_frame(seq[i])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(6, nfr):
            one_frame(seq[i])
        single_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / (nfr - 6))
        single = {"ms_per_frame": single_ms, "frames": nfr - 6,
                  "calls_per_frame": "orbx_detect + orbx_compute + 5 x hamx_match_ratio + 5 x fmx_compute_fundamental, blocking, pageable host buffers"}
        fm1.close()
        orb1.close()

    # ---- secondary metric: train-sharded Hamming kNN2, Gcmp/s over all ranks
    hamming = None
    if not args.no_hamming:
        nq, nt_shard = args.ham_nq, args.ham_nt
        gq = torch.Generator(device=dev); gq.manual_seed(7)
        q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=gq)
        gt = torch.Generator(device=dev); gt.manual_seed(60 + rank)
        t_shard = torch.randint(0, 256, (nt_shard, 32), dtype=torch.uint8, device=dev, generator=gt)
        # product exchange: peer memory (the matching kernel scatters its top-2 into every rank's gather buffer, flag-wait
        # merge); with one rank there is nothing to exchange and the plain kernel runs
        matcher.reserve(nq, max(nt_shard * world, 200000), 1)     # no allocation (device-wide sync) inside the timed calls
        sm = ShardedMatcher(matcher, p2p=world > 1, nq_max=nq)
        out = {}

        def check_sharded(mm, qs, shard, shard_rows, seed0, what):
            """The sharded result must equal a single-device pass over the concatenation of every rank's shard (regenerated
            here from the ranks' seeds); with one rank, a torch XOR + popcount reference on a few queries."""
            got = mm.knn2(qs, shard, rank * shard_rows)
            if world > 1:
                parts = []
                for r in range(world):
                    g = torch.Generator(device=dev); g.manual_seed(seed0 + r)
                    parts.append(torch.randint(0, 256, (shard_rows, 32), dtype=torch.uint8, device=dev, generator=g))
                full = torch.cat(parts)
                want = torch.empty_like(got)
                matcher.set_stream(stream.cuda_stream)
                matcher.knn2_dev(qs.data_ptr(), qs.shape[0], full.data_ptr(), full.shape[0], 0, want.data_ptr())
                ok = bool(torch.equal(got, want))
            else:
                lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=dev)
                sub = qs[:32]
                dmat = lut[(sub[:, None, :] ^ shard[None, :, :]).long()].sum(-1)                       # [32, nt]
                key = dmat.long() * (1 << 32) + torch.arange(shard.shape[0], device=dev)[None, :]
                k2 = torch.topk(key, 2, dim=1, largest=False).values
                want = torch.stack([k2[:, 0] >> 32, k2[:, 0] & 0xFFFFFFFF, k2[:, 1] >> 32, k2[:, 1] & 0xFFFFFFFF], 1).int()
                ok = bool(torch.equal(got[:32], want))
            t = torch.tensor([1 if ok else 0], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) != 1:
                print("bench.py: %s: the train-sharded top-2 differs from the single-device pass (rank %d ok=%s)" % (what, rank, ok), file=sys.stderr)
                if world > 1:
                    dist.barrier()
                sys.exit(3)
            return True
        parity_ok = check_sharded(sm, q[:4096].contiguous(), t_shard, nt_shard, 60, "config 5 (%d x %d per GPU)" % (nq, nt_shard))

        def step_ham():
            out["r"] = sm.knn2(q, t_shard, rank * nt_shard)
        from monocular_slam_b200 import _lib as _K
        ham_steps = max(2, min(args.steps, 3))
        # the integer-pipe kernel (XOR + POPC: what the roofline of SURVEY.md 8d describes) beside the product's choice, the
        # tensor-core kernel; same call, same exchange, identical results
        matcher.set_kernel(_K.KERNEL_INTEGER)
        int_ms = max_over_ranks(timed(step_ham, 2, 1)[0]) / 2
        r_int = out["r"].clone()
        matcher.set_kernel(_K.KERNEL_AUTO)
        ham_sampler = ClockSampler(local_rank, period=0.01)
        ham_ms, _ = timed(step_ham, ham_steps * 3, 2, ham_sampler)
        ham_ms = max_over_ranks(ham_ms) / (ham_steps * 3)
        assert torch.equal(out["r"], r_int), "tensor-core and integer-pipe kernels disagree"
        gpopc, _ = popc_peak(local_rank)
        gcmp = world * nq * nt_shard / (ham_ms * 1e-3) / 1e9
        gcmp_int = world * nq * nt_shard / (int_ms * 1e-3) / 1e9
        ham_clk = (ham_sampler.summary().get("sm_mhz") or 1965.0) * 1e6
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        tmem_peak = world * sms * 64 * ham_clk / 4 / 1e9        # accumulators per second readable from tensor memory (64 B / clk / SM)
        int8_peak = 4500.0 * world      # nominal dense int8 / fp8 TOP/s of a B200 (B200_PROFILING.md); no measured int8 peak exists
        hamming = {"value": gcmp, "unit": "Gcmp/s", "nq": nq, "nt_per_gpu": nt_shard, "ms_per_step": ham_ms, "clocks": ham_sampler.summary(),
                   "kernel": "tensor cores: +-1 int8 contraction (tcgen05.mma kind::i8, accumulators in tensor memory), integer pipes fold the top-2",
                   "workload": "%d queries x %d train rows per GPU (train-sharded; %s)" % (
                       nq, nt_shard, "top-2 exchanged over peer memory inside the matching kernel + flag-wait merge" if world > 1 else "single GPU, no exchange"),
                   "integer_pipe_kernel": {"value": gcmp_int, "unit": "Gcmp/s", "ms_per_step": int_ms, "frac_popc_pipe": gcmp_int * 5 / (gpopc * world),
                                           "frac_algorithmic": gcmp_int * 8 / (gpopc * world)},
                   "roofline": {"bound": "int_popc", "achieved": gcmp * 8, "peak": gpopc * world, "unit": "Gpopc/s",
                                "frac": gcmp * 8 / (gpopc * world),
                                "note": "ALGORITHMIC work of 8 POPC per 256-bit comparison (SURVEY.md 8d) against the register-only POPC "
                                        "microbenchmark (hamx_popc_peak) per GPU x n_gpus.  The product kernel issues no POPC at all (the distances "
                                        "come out of the tensor cores), hence frac >> 1; what bounds it is the tensor-memory read path: see roofline_tmem"},
                   "roofline_tmem": {"bound": "tensor memory read", "achieved": gcmp, "peak": tmem_peak, "unit": "G accumulators/s", "frac": gcmp / tmem_peak,
                                     "bytes_per_clk_per_sm": gcmp * 1e9 * 4 / (world * sms * ham_clk),
                                     "note": "every comparison is one 32-bit accumulator that has to leave tensor memory; peak = the 64 B/clk/SM measured "
                                             "for plain tcgen05.ld in B300_MICROARCH.md.  The kernel reads with .pack::16b (two 16-bit keys per register) "
                                             "and exceeds that figure, so it is a reference line, not a ceiling; the accumulator stage of a tile takes "
                                             "longer to read than to compute, which is what keeps the tensor pipe at ~68 % (profiles/r2_hamming_tc.txt)"},
                   "roofline_tensor": {"bound": "tensor", "achieved": gcmp * 2 * 288 / 1e3, "peak": int8_peak, "unit": "TOP/s (int8, 288-byte K incl. the index step)",
                                       "frac": (gcmp * 2 * 288 / 1e3 / int8_peak) if int8_peak else None,
                                       "peak_source": "nominal 4.5 POP/s dense int8 per B200 (B200_PROFILING.md table); ncu's tensor-pipe-active of the same kernel agrees"}}
        # BASELINE.json configs[3]: map-vs-frame tracking match, 200k map descriptors x 2000 queries, train set sharded over the ranks
        mq = q[:2000].contiguous()
        m_nt = 200000 // world
        gm = torch.Generator(device=dev); gm.manual_seed(90 + rank)
        m_shard = torch.randint(0, 256, (m_nt, 32), dtype=torch.uint8, device=dev, generator=gm)
        variants = {"peer_memory" if world > 1 else "single_gpu": sm}
        if world > 1:
            variants["nccl_all_gather"] = ShardedMatcher(matcher, p2p=False)
        mvf = {"nq": 2000, "nt_total": m_nt * world, "unit": "us per call (device time, max over ranks)"}
        for name, mm in variants.items():
            parity_ok = check_sharded(mm, mq, m_shard, m_nt, 90, "config 4 (%s)" % name) and parity_ok
            def step_mvf():
                out["m"] = mm.knn2(mq, m_shard, rank * m_nt)
            mvf_ms, _ = timed(step_mvf, 50, 5)
            mvf[name] = max_over_ranks(mvf_ms) / 50 * 1e3
        hamming["map_vs_frame"] = mvf
        hamming["parity_ok"] = parity_ok
        hamming["parity_check"] = ("sharded top-2 (4096 / 2000 queries) == hamx_knn2_dev over the concatenation of all ranks' shards, regenerated "
                                   "from their seeds on every rank" if world > 1 else "first 32 queries == torch XOR + popcount + topk over the shard")
        # flat scalars for the driver's record (it keeps scalars inside `roofline` only)
        roofline.update({"hamming_gcmp_s": gcmp, "hamming_ms_per_step": ham_ms, "hamming_frac_algorithmic": gcmp * 8 / (gpopc * world),
                         "hamming_frac_tmem_read": gcmp / tmem_peak, "hamming_frac_tensor": hamming["roofline_tensor"]["frac"],
                         "hamming_int_kernel_gcmp_s": gcmp_int, "hamming_int_kernel_frac_popc_pipe": gcmp_int * 5 / (gpopc * world),
                         "hamming_kernels_agree": 1, "hamming_popc_peak_gpopc_s_per_gpu": gpopc,
                         "hamming_sm_mhz": ham_sampler.summary().get("sm_mhz"),
                         "hamming_parity_ok": 1 if parity_ok else 0,
                         "mvf_us_peer": mvf.get("peer_memory", mvf.get("single_gpu")), "mvf_us_nccl": mvf.get("nccl_all_gather")})
        sm.close()

    # ---- the next stage of the reference after matching: computeFundamentalMatrix for every matched pair,
    # RANSAC status + 8-point F, one CUDA block per pair.  Workload: B frames x numBackTraverse = 5 predecessors per step.
    fundamental = None
    if not args.no_fundamental:
        from monocular_slam_b200 import FundamentalFilter
        fpairs, fn = B * 5, 1000
        p1 = np.zeros((fpairs, fn, 2), np.float32)
        p2 = np.zeros((fpairs, fn, 2), np.float32)
        for i in range(fpairs):
            p1[i], p2[i] = syn.two_view_matches(1000 * rank + i, fn, 0.7, 0.5, (W, H))
        fcounts = np.full(fpairs, fn, np.int32)
        fm = FundamentalFilter(device=local_rank)
        fm.set_stream(stream.cuda_stream)
        fd1, fd2 = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
        fdc = torch.from_numpy(fcounts).to(dev)
        fds = torch.zeros((fpairs, fn), dtype=torch.uint8, device=dev)
        fdF = torch.zeros((fpairs, 9), dtype=torch.float64, device=dev)
        fdi = torch.zeros((fpairs, 4), dtype=torch.int32, device=dev)

        def step_fm():
            fm.find_batch_dev(fd1.data_ptr(), fd2.data_ptr(), fdc.data_ptr(), fpairs, fn, 3.0, 0.85, fds.data_ptr(), fdF.data_ptr(), fdi.data_ptr())
        fm_ms, _ = timed(step_fm, args.steps, args.warmup)
        fm_ms = max_over_ranks(fm_ms) / args.steps
        finfo = fdi.cpu().numpy()
        # host-buffer call (fmx_fundamental_batch) on pinned memory, like the main e2e
        hp1, hp2 = torch.from_numpy(p1).pin_memory().numpy(), torch.from_numpy(p2).pin_memory().numpy()
        hst = torch.zeros((fpairs, fn), dtype=torch.uint8).pin_memory().numpy()
        hF = torch.zeros((fpairs, 9), dtype=torch.float64).pin_memory().numpy()
        hni = np.zeros(fpairs, np.int32)
        from monocular_slam_b200 import _lib as _L

        def host_call():
            _L.check(_L.lib().fmx_fundamental_batch(fm._h, hp1.ctypes.data, hp2.ctypes.data, fcounts.ctypes.data, fpairs, fn, 3.0, 0.85,
                                                    hst.ctypes.data, hF.ctypes.data, hni.ctypes.data))
        host_call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            host_call()
        fm_host_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 5
        assert np.array_equal(hni, finfo[:, 0]), "host and device paths of the fundamental filter disagree"
        fundamental = {"value": world * fpairs / (fm_ms * 1e-3), "unit": "pairs/s", "ms_per_step": fm_ms,
                       "workload": "%d pairs per GPU and step (%d frames x 5 predecessors) x %d matches, 70 %% inliers, 0.5 px noise; "
                                   "findFundamentalMat(FM_RANSAC, 3, 0.85) status + FM_8POINT on the inliers" % (fpairs, B, fn),
                       "l2": "the correspondences of a step (%.1f MB) are L2-resident between launches; the kernel keeps them in shared memory and "
                             "is bound by per-pair latency, not by memory (profiles/r1g_fmat.txt: 7.6 MB of DRAM traffic per launch)" % ((p1.nbytes + p2.nbytes) / 1e6),
                       "ransac_iterations_mean": float(finfo[:, 1].mean()), "candidates_scored_mean": float(finfo[:, 2].mean()),
                       "inliers_mean": float(finfo[:, 0].mean()),
                       "e2e": {"value": world * fpairs / (fm_host_ms * 1e-3), "unit": "pairs/s", "ms_per_step": fm_host_ms,
                               "h2d_bytes_per_step": int(p1.nbytes + p2.nbytes + fcounts.nbytes),
                               "d2h_bytes_per_step": int(fpairs * fn + fpairs * 72 + fpairs * 16),
                               "timing": "host wall clock around fmx_fundamental_batch (pinned host buffers), max over ranks"}}
        # the sequence pipeline with the filter appended to every batch (orbx_submit_batch_filtered): frames in, keypoints,
        # descriptors, consecutive-frame matches, their RANSAC status and F out
        pstate["outs"] = [o + (torch.empty((B, cap), dtype=torch.uint8).pin_memory().numpy(),
                               torch.zeros((B, 3, 3), dtype=torch.float64).pin_memory().numpy(), np.zeros(B, np.int32)) for o in outs]
        pstate["fm"] = fm
        orb.reset_sequence()
        pf_ms = max_over_ranks(timed_pipe(args.steps, max(args.warmup, 2)))
        pstate["fm"] = None
        last = pstate["last"]
        fundamental["sequence_pipeline"] = {
            "value": world * B * args.steps / (pf_ms * 1e-3), "unit": "frames/s", "ms_per_step": pf_ms / args.steps,
            "what": "e2e loop of this bench with computeFundamentalMatrix of every (frame, predecessor) pair added on the device "
                    "(orbx_submit_batch_filtered); compare with e2e.value",
            "matches_per_pair": float(last[4][1:].mean()), "inliers_per_pair": float(last[7][1:].mean()),
            "note": "the synthetic sequence is a translating texture, i.e. a planar scene: F is degenerate there, the numbers time the "
                    "filter on real match lists but say nothing about pose quality"}
        # the reference's steady-state loop (src/CameraPoseEstimator.cpp:405-419) through orbx_submit_batch_back: every frame
        # extracted, matched against its 5 predecessors and each of the 5 match lists filtered, host buffers in and out
        nb_ = 5

        def pinned(shape, dtype):
            return torch.zeros(shape, dtype=dtype).pin_memory().numpy()
        pstate["outs"] = [o[:3] + (pinned((B, nb_, cap, 4), torch.int32).view(DMATCH_DTYPE).reshape(B, nb_, cap), np.zeros((B, nb_), np.int64),
                                   pinned((B, nb_, cap), torch.uint8), pinned((B, nb_, 3, 3), torch.float64), np.zeros((B, nb_), np.int32))
                          for o in outs]
        pstate["fm"], pstate["back"] = fm, nb_
        orb.reset_sequence()
        ss_ms = max_over_ranks(timed_pipe(args.steps, max(args.warmup, 2)))
        pstate["fm"], pstate["back"] = None, 0
        last = pstate["last"]
        fundamental["steady_state_pipeline"] = {
            "value": world * B * args.steps / (ss_ms * 1e-3), "unit": "frames/s", "ms_per_step": ss_ms / args.steps,
            "what": "per frame: extraction + matchFeatures against the 5 previous frames + computeFundamentalMatrix of the 5 match lists "
                    "(orbx_submit_batch_back, 3 batches in flight, pinned host buffers in and out)",
            "pairs_per_step": B * nb_, "matches_per_pair": float(last[4].mean()), "inliers_per_pair": float(last[7].mean()),
            "d2h_bytes_per_step": int(B * cap * 60 + B * nb_ * cap * 17 + B * nb_ * 88)}
        if world == 1 and not args.no_cpu:
            try:
                import oracle
                ns = 16
                t0 = time.perf_counter()
                for i in range(ns):
                    _, m, _ = oracle.fm_ransac(p1[i], p2[i], 3.0, 0.85)
                    oracle.fm_8point(p1[i][m > 0], p2[i][m > 0])
                o_ms = (time.perf_counter() - t0) * 1e3 / ns
                fundamental["cpu_baseline"] = {"value": 1e3 / o_ms, "unit": "pairs/s", "cores": 1, "kind": "port",
                                               "sample": "oracle/fmat_oracle.c on the first %d pairs of the step, one thread" % ns}
                try:
                    import cv2
                    cv2.setNumThreads(1)
                    t0 = time.perf_counter()
                    for i in range(ns):
                        a, b = p1[i].astype(np.float64), p2[i].astype(np.float64)
                        _, m = cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, 3.0, 0.85)
                        cv2.findFundamentalMat(a[m.ravel() > 0], b[m.ravel() > 0], cv2.FM_8POINT)
                    c_ms = (time.perf_counter() - t0) * 1e3 / ns
                    fundamental["cpu_baseline_cv2"] = {"value": 1e3 / c_ms, "unit": "pairs/s", "cores": 1, "kind": "reference",
                                                       "sample": "cv2 %s findFundamentalMat x2 on the same %d pairs, one thread" % (cv2.__version__, ns)}
                except ImportError:
                    pass
            except Exception as e:
                fundamental["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (e,)}
        # ---- the consumer of the filter's status masks: triangulation of every inlier of every pair,
        # TriangulateSinglePointFromTwoView (src/CameraPoseEstimator.cpp:86-132) one thread per correspondence
        if not args.no_triangulation:
            from monocular_slam_b200 import CAMERAS_DTYPE, Triangulator
            tri = Triangulator(device=local_rank)
            tri.set_stream(stream.cuda_stream)
            a = 0.05
            K = np.array([[900.0, 0, W / 2], [0, 900.0, H / 2], [0, 0, 1]])       # the cameras synthetic.two_view_matches projects with
            R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
            cam = np.zeros(fpairs, CAMERAS_DTYPE)
            cam["Rt1"], cam["Rt2"], cam["K1"], cam["K2"] = np.c_[np.eye(3), np.zeros(3)], np.c_[R, np.array([0.4, 0.05, 0.1])], K, K
            d_cam = torch.from_numpy(cam.view(np.float64).reshape(fpairs, 42)).to(dev)
            tX = torch.zeros((fpairs, fn, 3), dtype=torch.float64, device=dev)
            tF = torch.zeros((fpairs, fn), dtype=torch.uint8, device=dev)
            tN = torch.zeros(fpairs, dtype=torch.int32, device=dev)

            def step_tri():
                tri.triangulate_batch_dev(fd1.data_ptr(), fd2.data_ptr(), fdc.data_ptr(), fds.data_ptr(), fpairs, fn, d_cam.data_ptr(), 1,
                                          tX.data_ptr(), tF.data_ptr(), tN.data_ptr())
            tri_ms = max_over_ranks(timed(step_tri, args.steps, args.warmup)[0]) / args.steps
            npts = int(fds.sum().item())
            nfront = int(tN.sum().item())
            hsel = fds.cpu().numpy()

            def tri_host():
                return tri.triangulate_batch(hp1, hp2, fcounts, cam, select=hsel)
            tri_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                hX, hFr, hN = tri_host()
            tri_host_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 3
            assert int(hN.sum()) == nfront, "host and device paths of the triangulation disagree"
            triangulation = {"value": world * npts / (tri_ms * 1e-3) / 1e6, "unit": "Mpoints/s", "ms_per_step": tri_ms,
                             "workload": "the %d pairs of the fundamental leg: every RANSAC inlier triangulated (%d points per GPU and step, %d in front of both cameras)"
                                         % (fpairs, npts, nfront),
                             "e2e": {"value": world * npts / (tri_host_ms * 1e-3) / 1e6, "unit": "Mpoints/s", "ms_per_step": tri_host_ms,
                                     "h2d_bytes_per_step": int(2 * p1.nbytes + fcounts.nbytes + hsel.nbytes + cam.nbytes),
                                     "d2h_bytes_per_step": int(fpairs * fn * 25 + fpairs * 4),
                                     "timing": "host wall clock around trx_triangulate_batch (pageable numpy buffers), max over ranks"}}
            if world == 1 and not args.no_cpu:
                try:
                    import oracle
                    ns = 16
                    t0 = time.perf_counter()
                    for i in range(ns):
                        sel = hsel[i, :fn] > 0
                        oracle.triangulate(p1[i][sel].astype(np.float64), p2[i][sel].astype(np.float64), cam[i]["Rt1"], cam[i]["Rt2"], K, K)
                    o_s = time.perf_counter() - t0
                    pts = int(hsel[:ns].sum())
                    triangulation["cpu_baseline"] = {"value": pts / o_s / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
                                                     "sample": "oracle/tri_oracle.c on the inliers of the first %d pairs, one thread" % ns}
                    try:
                        import cv2
                        cv2.setNumThreads(1)
                        P1, P2 = K @ cam[0]["Rt1"], K @ cam[0]["Rt2"]
                        t0 = time.perf_counter()
                        for i in range(ns):
                            sel = hsel[i, :fn] > 0
                            cv2.triangulatePoints(P1, P2, p1[i][sel].astype(np.float64).T.copy(), p2[i][sel].astype(np.float64).T.copy())
                        c_s = time.perf_counter() - t0
                        triangulation["cpu_baseline_cv2"] = {"value": pts / c_s / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "reference",
                                                             "sample": "cv2 %s triangulatePoints (the same per-point 4x4 cv::SVD) on the same points, one thread" % cv2.__version__}
                    except ImportError:
                        pass
                except Exception as e:
                    triangulation["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (e,)}
            tri.close()
            fundamental["triangulation"] = triangulation
            roofline.update({"tri_mpoints_s": triangulation["value"], "tri_ms_per_step": tri_ms})
        fm.close()

    # ---- loop-closure candidate scoring: the current frame's descriptors against 1000 stored frames of
    # 2000 descriptors in one launch (LoopCloser::DetectLoop, src/LoopCloser.cpp:19-51); stored frames sharded by frame over
    # the ranks (fixed total: strong scaling), per-frame scores exchanged with one all-gather
    loop = None
    if not args.no_loop:
        from monocular_slam_b200.sharded import ShardedLoopScorer
        lf_total, lcap, lnq, ln, lthr = 1000, 2000, 2000, 10, 40
        lb = shard_bounds(lf_total, world)
        llo, lhi = int(lb[rank]), int(lb[rank + 1])
        gl = torch.Generator(device=dev); gl.manual_seed(17)
        lq = torch.randint(0, 256, (lnq, 32), dtype=torch.uint8, device=dev, generator=gl)
        gs = torch.Generator(device=dev); gs.manual_seed(1700 + rank)
        lframes = torch.randint(0, 256, (lhi - llo, lcap, 32), dtype=torch.uint8, device=dev, generator=gs)
        lcounts = torch.full((lhi - llo,), lcap, dtype=torch.int32, device=dev)
        revisit = 617                                   # one stored frame is a noisy copy of the current one
        if llo <= revisit < lhi:
            lframes[revisit - llo, :lnq] = lq
            lframes[revisit - llo, :lnq, 5] ^= 0x0F
        scorer = ShardedLoopScorer(matcher, n=ln, thr=lthr)
        lout = {}

        def step_loop():
            lout["r"] = scorer.score(lq, lframes, lcounts, lf_total)
        loop_ms = max_over_ranks(timed(step_loop, 5, 2)[0]) / 5
        lscores, lbest = lout["r"]
        assert int(lbest[0]) == revisit and int(lbest[1]) == lnq, "loop scoring did not find the revisited frame"
        loop = {"value": lnq * lcap * lf_total / (loop_ms * 1e-3) / 1e9, "unit": "Gcmp/s", "ms_per_query_frame": loop_ms,
                "workload": "%d descriptors of the current frame x %d stored frames x %d descriptors, n-best = %d, threshold %d bits; frames "
                            "sharded over %d GPU(s), 4-byte scores all-gathered" % (lnq, lf_total, lcap, ln, lthr, world),
                "best_frame": int(lbest[0]), "stored_descriptor_mb": lf_total * lcap * 32 / 1e6}
        if world == 1 and not args.no_cpu:
            try:
                import oracle
                ns = 2
                t0 = time.perf_counter()
                oracle.loop_score(lq.cpu().numpy(), lframes[:ns].cpu().numpy(), np.full(ns, lcap, np.int32), ln, lthr)
                o_s = time.perf_counter() - t0
                loop["cpu_baseline"] = {"value": lnq * lcap * ns / o_s / 1e9, "unit": "Gcmp/s", "cores": 1, "kind": "port",
                                        "sample": "oracle/loop_oracle.c (NBestMatches lists + count) on the first %d stored frames, one thread" % ns}
            except Exception as e:
                loop["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (e,)}
        roofline.update({"loop_gcmp_s": loop["value"], "loop_ms_per_query_frame": loop_ms})
        del lframes

    # ---- bag of words (the DBoW2 half: ThirdParty/DBoW2): the batch's descriptors, still on the device,
    # through a k = 10, L = 6 vocabulary (the ORB-SLAM shape, 1.1 M nodes, synthetic) into one BowVector + FeatureVector per
    # frame, then one frame's vector scored against a database of stored vectors.  Frames are independent: sharded like the
    # extraction, no collective.
    bow = None
    if not args.no_bow:
        from monocular_slam_b200 import Vocabulary
        vk, vL, levelsup, ndb = 10, 6, 4, 10000
        va = syn.vocabulary_large(77, vk, vL)
        voc = Vocabulary(va, device=local_rank)
        voc.set_stream(stream.cuda_stream)
        orb.extract_batch_dev(d_frames.data_ptr(), W * H, B, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
        bw = torch.zeros((B, cap), dtype=torch.int32, device=dev); bv = torch.zeros((B, cap), dtype=torch.float64, device=dev)
        bn = torch.zeros(B, dtype=torch.int32, device=dev); fn_ = torch.zeros(B, dtype=torch.int32, device=dev)
        fnod = torch.zeros((B, cap), dtype=torch.int32, device=dev); foff = torch.zeros((B, cap + 1), dtype=torch.int32, device=dev)
        ffe = torch.zeros((B, cap), dtype=torch.int32, device=dev)

        def step_bow():
            voc.transform_batch_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, cap, levelsup, bw.data_ptr(), bv.data_ptr(), bn.data_ptr(),
                                    fnod.data_ptr(), foff.data_ptr(), ffe.data_ptr(), fn_.data_ptr())
        reps = 10
        bow_ms = max_over_ranks(timed(step_bow, reps, 3)[0]) / reps
        nfeat = int(d_cnt.sum().item())
        nbw = bn.cpu().numpy()
        # the database: ndb stored vectors, packed rows as large as ndb distinct vectors would be (so that the reads come from
        # HBM, not from L2): the vectors of the batch's other frames over and over, and frame 0 itself once, at entry `revisit_e`
        per = int(nbw.max())
        revisit_e = 6173
        src = (1 + torch.arange(ndb, device=dev) % max(B - 1, 1)) % B
        src[revisit_e] = 0
        pw, pv, dcount = bw[src, :per].contiguous(), bv[src, :per].contiguous(), bn[src].contiguous()
        pstart = torch.arange(ndb, dtype=torch.int64, device=dev) * per
        dscore = torch.zeros(ndb, dtype=torch.float64, device=dev)

        def step_score():
            voc.score_batch_dev(bw[0].data_ptr(), bv[0].data_ptr(), int(nbw[0]), pstart.data_ptr(), dcount.data_ptr(), pw.data_ptr(), pv.data_ptr(), ndb,
                                dscore.data_ptr())
        score_ms = max_over_ranks(timed(step_score, reps, 3)[0]) / reps
        sc = dscore.cpu().numpy()
        assert abs(sc[revisit_e] - 1.0) < 1e-9 and (B < 2 or int(sc.argmax()) == revisit_e), "the stored copy of the query must score 1 and win"
        db_bytes = int(dcount.sum().item()) * 4 + int(nbw[0]) * 8      # every stored word id once, the values of the common words
        bow = {"value": world * nfeat / (bow_ms * 1e-3) / 1e6, "unit": "Mfeatures/s", "transform_ms_per_step": bow_ms,
               "workload": "%d frames x %d descriptors through a k=%d L=%d vocabulary (%d nodes, %.0f MB of node records), levelsup %d: "
                           "BowVector + FeatureVector per frame" % (B, cap, vk, vL, len(va["parent"]), len(va["parent"]) * 48 / 1e6, levelsup),
               "words_per_frame": float(nbw.mean()),
               "record_bytes_per_step": nfeat * vL * vk * 48, "record_gbs": nfeat * vL * vk * 48 / (bow_ms * 1e-3) / 1e9,
               "score_ms": score_ms, "score_entries": ndb, "score_mentries_s": world * ndb / (score_ms * 1e-3) / 1e6,
               "score_db_bytes": db_bytes, "score_gbs": db_bytes / (score_ms * 1e-3) / 1e9, "score_frac_hbm": db_bytes / (score_ms * 1e-3) / 1e9 / peak}
        if world == 1 and not args.no_cpu:
            try:
                import oracle
                ov = oracle.BowVocabulary(va)
                hd, hc = d_desc[:4].cpu().numpy(), d_cnt[:4].cpu().numpy()
                t0 = time.perf_counter()
                obows = [ov.transform(hd[f, :hc[f]], levelsup) for f in range(4)]
                o_s = time.perf_counter() - t0
                for f in range(4):     # the oracle is the checker here too
                    assert np.array_equal(obows[f][0], bw[f, :nbw[f]].cpu().numpy().view(np.uint32)) and np.array_equal(obows[f][1], bv[f, :nbw[f]].cpu().numpy())
                t0 = time.perf_counter()
                os_ = [ov.score(obows[0][:2], obows[e % 4][:2]) for e in range(400)]
                s_s = time.perf_counter() - t0
                assert os_[0] == sc[revisit_e] and (B < 4 or (os_[1] == sc[0] and os_[2] == sc[1]))      # entries 0, 1 hold frames 1, 2
                bow["cpu_baseline"] = {"value": int(hc.sum()) / o_s / 1e6, "unit": "Mfeatures/s", "cores": 1, "kind": "port",
                                       "sample": "oracle/bow_oracle.c transform() of the first 4 frames, one thread",
                                       "score_mentries_s": 400 / s_s / 1e6}
            except Exception as e:
                bow["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (e,)}
        roofline.update({"bow_mfeatures_s": bow["value"], "bow_transform_ms": bow_ms, "bow_score_ms": score_ms, "bow_score_frac_hbm": bow["score_frac_hbm"]})
        voc.close()
        del pw, pv

    # ---- frame ingest: the reference decodes every frame on the host (imread, src/FrameLoader.cpp:62).  The
    # same frames as PNG / JPEG files in memory -> decode workers -> pinned slots -> the pipelined extract + match path: frames/s
    # one worker sustains, frames/s of the whole ring on this box's cores, and the cores a GPU-bound rate would need
    ingest = None
    if world == 1 and not args.no_ingest:
        try:
            import cv2
            from monocular_slam_b200.ingest import IngestRing, imread_unchanged
            try:
                ncpu = len(os.sched_getaffinity(0))
            except Exception:
                ncpu = os.cpu_count() or 1
            ingest = {"host_cores": ncpu, "frame": "%dx%d gray" % (W, H), "gpu_bound_fps": e2e_value}
            for fmt, ext, params in (("png", ".png", [cv2.IMWRITE_PNG_COMPRESSION, 3]), ("jpeg", ".jpg", [cv2.IMWRITE_JPEG_QUALITY, 90])):
                enc = [cv2.imencode(ext, seq[i], params)[1].tobytes() for i in range(B)]
                t0 = time.perf_counter()
                for e in enc[:8]:
                    imread_unchanged(e)
                per_core = 8 / (time.perf_counter() - t0)
                ring = IngestRing(orb, matcher, W, H, batch=B, workers=ncpu, ratio=RATIO)
                nb = 0
                for res in ring.run(enc * 2):       # first pass: warm-up
                    nb += 1
                ring.stats.update({"frames": 0, "decode_seconds": 0.0, "wall_seconds": 0.0})
                reps = max(2, int(1.5 * per_core * ncpu / B))
                kp_total = 0
                for first, n, kps_, desc_, counts_, good_, ngood_ in ring.run(enc * reps):
                    kp_total += int(counts_[:n].sum())
                st = dict(ring.stats)
                ring.close()
                ingest[fmt] = {"bytes_per_frame": int(np.mean([len(e) for e in enc])), "decode_fps_per_core": per_core,
                               "ring_fps": st["frames"] / st["wall_seconds"], "ring_workers": ncpu, "frames": st["frames"],
                               "decode_fps_per_worker_in_ring": st["frames"] / max(st["decode_seconds"], 1e-9),
                               "keypoints_per_frame": kp_total / max(st["frames"], 1),
                               "cores_for_gpu_bound_rate": e2e_value / per_core}
            # the same JPEG frames with the entropy decoder on the GPU (jpgx_*): files in host memory -> compressed bytes over PCIe
            # -> restart-interval parallel Huffman + IDCT -> the extractor's device frames -> extraction + matching -> results in
            # pinned host memory.  `rows`: one restart marker per block row (what a recorder under our control writes);
            # `none`: no restart markers (one warp per file: only large batches pay)
            from monocular_slam_b200 import JpegDecoder
            jdec = JpegDecoder(device=local_rank)
            jstream = torch.cuda.Stream(device=dev)       # the decoder's own stream: batch k+1 is decoded while batch k is extracted
            jdec.set_stream(jstream.cuda_stream)
            jbuf = [torch.zeros((B, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
            ev_dec = [torch.cuda.Event() for _ in range(2)]
            ev_ext = [torch.cuda.Event() for _ in range(2)]
            h_kps = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory(); h_desc = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
            h_good = torch.empty((B, cap, 4), dtype=torch.int32).pin_memory(); h_cnt = torch.empty(B, dtype=torch.int32).pin_memory()
            ingest["jpeg_gpu"] = {}
            jstate = {"k": 0}
            for label, rst in (("rows", (W + 7) // 8), ("none", 0)):
                enc = [cv2.imencode(".jpg", seq[i], [cv2.IMWRITE_JPEG_QUALITY, 90] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else []))[1].tobytes()
                       for i in range(B)]

                def step_jpeg():
                    b = jstate["k"] & 1
                    jstate["k"] += 1
                    jstream.wait_event(ev_ext[b])                       # the extraction that read this buffer two steps ago
                    jdec.decode_dev(enc, W, H, jbuf[b].data_ptr(), W * H, W)
                    ev_dec[b].record(jstream)
                    stream.wait_event(ev_dec[b])
                    orb.extract_batch_dev(jbuf[b].data_ptr(), W * H, B, W, H, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_cnt.data_ptr())
                    ev_ext[b].record(stream)
                    matcher.match_consecutive_dev(d_desc.data_ptr(), d_cnt.data_ptr(), B, cap, 0, 0, RATIO, d_good.data_ptr(), d_ngood.data_ptr())
                    h_kps.copy_(d_kps, non_blocking=True); h_desc.copy_(d_desc, non_blocking=True)
                    h_good.copy_(d_good, non_blocking=True); h_cnt.copy_(d_cnt, non_blocking=True)
                jreps = 8 if rst else 2
                _, jwall = timed(step_jpeg, jreps, 2)
                orb.check_dev()
                ref0 = cv2.imdecode(np.frombuffer(enc[0], np.uint8), cv2.IMREAD_UNCHANGED)
                assert np.array_equal(jbuf[(jstate["k"] - 1) & 1][0].cpu().numpy(), ref0), "GPU-decoded frame differs from cv2.imdecode"
                # the decoder alone (files in host memory -> device frames), on its own stream
                def step_decode():
                    jdec.decode_dev(enc, W, H, jbuf[0].data_ptr(), W * H, W)
                t0 = time.perf_counter()
                for _ in range(jreps):
                    step_decode()
                jstream.synchronize()
                dwall = time.perf_counter() - t0
                ingest["jpeg_gpu"][label] = {"fps": B * jreps / (jwall * 1e-3), "bytes_per_frame": int(np.mean([len(e) for e in enc])),
                                             "restart_interval_blocks": rst, "keypoints_per_frame": float(h_cnt.numpy().mean()),
                                             "decode_only_fps": B * jreps / dwall,
                                             "decode_only_compressed_gbs": sum(len(e) for e in enc) * jreps / dwall / 1e9}
            jdec.close()
            e2e_extra = {"ingest_png_fps": ingest["png"]["ring_fps"], "ingest_jpeg_fps": ingest["jpeg"]["ring_fps"],
                         "ingest_png_decode_fps_per_core": ingest["png"]["decode_fps_per_core"],
                         "ingest_jpeg_decode_fps_per_core": ingest["jpeg"]["decode_fps_per_core"],
                         "ingest_jpeg_gpu_decode_fps": ingest["jpeg_gpu"]["rows"]["fps"],
                         "ingest_jpeg_gpu_decode_no_restart_fps": ingest["jpeg_gpu"]["none"]["fps"]}
        except Exception as e:
            ingest = {"failed": repr(e)}

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample on the host cores -- the same B frames the GPU step
    # processes, dealt to one single-threaded cv2 process per core; plus the two thread settings BASELINE.md asks for
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            ref = CpuReference(B)
            n, dt = 0, 0.0
            for _ in range(2):
                a, b = ref.step()
                n += a
                dt += b
            ref.close()
            cpu = ref.describe()
            cpu.update({"value": n / dt, "unit": "frames/s"})
            cpu.update({"ref_" + k: v for k, v in ref.stats.items()})
            cpu.update(cpu_thread_settings())
        except Exception as e:   # the baseline is reporting only; never fail the GPU measurement for it
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}

    clocks = sampler.summary()
    # kernels of liborbx.so per timed step: ingest; per part of the batch (ORBX_SPLIT=n cuts batches in n parts on n
    # streams; default 1): 7 pyramid levels, FAST, score cut, Harris selection, orient+describe; then pair table, train-set
    # expansion, tensor-core kNN2, ratio test (profiles/r2_launches.csv)
    try:
        halves = max(1, min(int(os.environ.get("ORBX_SPLIT", "1")), B // 8))
    except ValueError:
        halves = 1
    launches_per_step = 1 + halves * ((len(ws) - 1) + 4) + 4
    if rank == 0:
        assert level_px == level_pixels()
        cfg = workload_config(world, B)
        cfg.update({"keypoints_per_frame": float(counts.mean()), "matches_per_frame": float(ngood_dev[1:].mean())})
        e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms / args.steps,
               "timing": "host wall clock around orbx_submit_batch / orbx_wait_batch (%d batches in flight, drained before the clock stops)" % depth + ", max over ranks",
               "blocking_value": blk_value, "blocking_ms_per_step": blk_ms / args.steps,
               "blocking_note": "orbx_extract_batch + orbx_match_consecutive, each call returns with its results in host memory",
               # the same bytes as plain pinned copies in the same run, every rank at once: what the box's PCIe path allows
               "h2d_floor_ms": h2d_floor_ms, "d2h_floor_ms": d2h_floor_ms, "h2d_floor_gbs_per_gpu": h2d / (h2d_floor_ms * 1e-3) / 1e9,
               "over_h2d_floor": (e2e_ms / args.steps) / h2d_floor_ms,
               # upload and download running together (two streams), every rank at once: the comparator for the pipelined step
               "duplex_floor_ms": duplex_floor_ms, "over_duplex_floor": (e2e_ms / args.steps) / duplex_floor_ms,
               "numa_bound_cpus_rank0": (len(numa_cpus) if numa_cpus else None)}
        if single:
            e2e["single_frame_ms"] = single["ms_per_frame"]
        if ingest and "png" in ingest:
            e2e.update(e2e_extra)
        if cfg3:
            e2e["cfg3_e2e_fps"] = cfg3["e2e"]["value"]
            roofline["cfg3_fps"] = cfg3["value"]
        if sustained:
            roofline.update({"sustained_fps": sustained["value"], "sustained_seconds": sustained["seconds"],
                             "sustained_sm_mhz": sustained["clocks"].get("sm_mhz")})
        if natural:
            nat_fast = natural["stages_ms_per_step"]["fast"]
            nat_traffic = None
            try:
                nat_traffic = json.load(open(tp)).get("natural", {}).get("dram_bytes_per_launch") if (W, H, NFEAT, B) == (1920, 1080, 2000, 64) else None
            except Exception:
                pass
            natural["roofline"] = {"kernel": "k_fast", "bound": "hbm", "achieved": level_px * B / (nat_fast * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": level_px * B / (nat_fast * 1e-3) / 1e9 / peak, "traffic": nat_traffic, "launch_ms": nat_fast}
            roofline.update({"natural_fps": natural["value"], "natural_fast_ms": nat_fast, "dense_fast_ms": stages["fast"],
                             "natural_fast_frac": natural["roofline"]["frac"], "natural_fast_traffic": nat_traffic})
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": cfg,
                "clocks": clocks,
                "e2e": e2e,
                "gpu_launches": launches_per_step * args.steps,
                "roofline": roofline,
                "roofline_pyramid": roofline_pyramid,
                "stages_ms_per_step": dict(stages, match=match_ms, profiled_step=prof_ms / args.steps),
                "cpu_baseline": cpu,
                "sustained": sustained,
                "natural_images": natural,
                "config3": cfg3,
                "single_frame": single,
                "hamming": hamming,
                "fundamental": fundamental,
                "loop_closure": loop,
                "bag_of_words": bow,
                "ingest": ingest}
        emit(line)
    matcher.close()
    orb.close()
    if world > 1:
        dist.destroy_process_group()


def _protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there when NCCL_DEBUG is set)
    write to file descriptor 1 directly, so point fd 1 at stderr for the whole run and keep the real stdout aside."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


_REAL_STDOUT = sys.stdout


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU and step")
    ap.add_argument("--ham-nq", type=int, default=1 << 20)
    ap.add_argument("--ham-nt", type=int, default=125000, help="train rows per GPU")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-fundamental", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-natural", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--no-triangulation", action="store_true")
    ap.add_argument("--no-loop", action="store_true")
    ap.add_argument("--no-bow", action="store_true")
    ap.add_argument("--no-ingest", action="store_true")
    ap.add_argument("--wc-frames", action="store_true", help="keep the host frames in write-combined page-locked memory")
    ap.add_argument("--frame", default="1920x1080", help="frame size WxH (default: BASELINE.json configs[1]; 3840x2160 with --nfeatures 8000 is configs[2])")
    ap.add_argument("--nfeatures", type=int, default=2000)
    args = ap.parse_args()
    global W, H, NFEAT
    W, H = (int(v) for v in args.frame.lower().split("x"))
    NFEAT = args.nfeatures
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        print("bench.py: --gpus %d needs torchrun (one process per GPU); running 1 GPU" % args.gpus, file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
