#!/usr/bin/env python
"""bench.py -- TES acquisition throughput on B200 (BASELINE.json metric: acq evals/s,
1 eval = one (candidate, maximizer-sample) pair).

A "step" is ONE full acquisition step (stages A-D of SURVEY.md section 8) over one batch of synthetic
input.  Default workload = config C3 of BASELINE.json (the largest single-GPU configuration):
Hartmann-6 hyper-parameters, 10^6 candidates per GPU (the 10-per-axis mesh of functions.get_meshgrid; with N ranks the
global mesh has 10 N points along its slowest axis and every rank holds one 10^6-point slab), S = 1024 RFF function samples with F = 1024 features each, n = 64
observations, TES_ep (criterion ftl, mode empirical, deterministic) on the <= t_max most frequent
trusted maximizers.  The three global-RNG draws of the RFF posterior weight draw (optfunc.py:329, :334, :344) are
host-supplied inputs of the step; the draw itself (theta/W/b, optfunc.py:303-385) runs inside the timed step.

  value    evals/s with every input already resident in HBM (device-timed, max over ranks)
  e2e      the same step through the reference-facing Python API with HOST (pinned) buffers: H2D copies
           of candidates / observations / RFF weights and the D2H read of the result are inside the
           timed region
  roofline the dominant kernel (stage A: tcgen05 screen + fp64 refine; MUFU-bound) against the in-run measured peak
  cpu_baseline / --impl reference: the CPU oracle port of the same step on the box's host cores, on a
           bounded sub-sample of the same workload

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c3u|c3s|c3m|c2|c4|c4m|c4s]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the CPU arm uses every host core, whatever the launcher exported (torch.distributed.run
# sets OMP_NUM_THREADS=1 for its workers).  Must happen before numpy / OpenMP load.
CPU_THREADS = os.cpu_count() or 1
if "--impl" in sys.argv and "reference" in sys.argv:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(CPU_THREADS)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HART6 = dict(l=np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949]), sigma=1.423, sigma0=1e-4)  # functions.py:540-543


def hartmann6(x):
    """functions.negative_rescaled_hartmann6d (functions.py:500-545), restated for the harness."""
    A = np.array([[10., 3., 17., 3.5, 1.7, 8.], [0.05, 10., 17., 0.1, 8., 14.], [3., 3.5, 1.7, 10., 17., 8.],
                  [17., 8., 0.05, 10., 0.1, 14.]])
    alpha = np.array([1., 1.2, 3., 3.2])
    P = 1e-4 * np.array([[1312., 1696., 5569., 124., 8283., 5886.], [2329., 4135., 8307., 3736., 1004., 9991.],
                         [2348., 1451., 3522., 2883., 3047., 6650.], [4047., 8828., 8732., 5743., 1091., 381.]])
    x = x.reshape(-1, 1, 6)
    return (2.58 + np.sum(alpha * np.exp(-np.sum(A * (x - P) ** 2, axis=2)), axis=1)) / 1.94


def meshgrid(xmin, xmax, nx, xdim, slab=0, nslab=1):
    """functions.get_meshgrid (functions.py:135-141).  nslab > 1: the mesh is refined to nx * nslab points along its
    slowest axis (the second coordinate, np.meshgrid's 'xy' order) and slab `slab` of nx of them is returned, i.e.
    the contiguous shard of a rank when the global mesh is split across nslab GPUs."""
    x1d = np.linspace(xmin, xmax, nx)
    axes = [x1d] * xdim
    if nslab > 1:
        axes = list(axes)
        axes[1] = np.linspace(xmin, xmax, nx * nslab)[slab * nx:(slab + 1) * nx]
    xds = np.meshgrid(*axes)
    return np.concatenate([xd.reshape(-1, 1) for xd in xds], axis=1)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE rff_screen_tc5_kernel launch, from the ncu --set full capture
# summarised in profiles/ (filled per workload once measured; None = not captured)
TC5_TRAFFIC_BYTES = {"c3": 465.8e6, "c3u": 465.8e6}   # profiles/r01_ncu_rff_screen_tc5.md (algorithmic: 115 MB)
# same for ONE launch of the product-path GEMM kernel (stage A on meshes)
PROD_TRAFFIC_BYTES = {"c3": 715.1e6}  # profiles/r02_ncu_prod_gemm_tc8.md (read 677.9 + write 37.2 MB per 37-sample launch)

WORKLOADS = {
    # d, N per GPU, S function samples, F features, n observations, t_max test points; criterion/mode/q and the
    # Monte-Carlo sizes of the criterion (I = niteration, Y = nysample)
    "c3": dict(desc="C3 Hartmann-6 TES_ep empirical, 1M candidates x 1024 maximizer samples per GPU",
               d=6, N=10 ** 6, S=1024, F=1024, n=64, t_max=128),
    "c3u": dict(desc="C3 shapes on UNSTRUCTURED (uniform random) candidates: the general MUFU-bound stage A", d=6,
                N=10 ** 6, S=1024, F=1024, n=64, t_max=128, uniform=True),
    "c3s": dict(desc="C3 shapes scaled down (smoke)", d=6, N=50_000, S=128, F=256, n=32, t_max=32),
    "c3m": dict(desc="C3 shapes, mid size (profiling)", d=6, N=200_000, S=256, F=1024, n=64, t_max=64),
    "c2": dict(desc="C2 Brooms-Barn pH hyper-parameters, 2-D 20x20 grid, TES_ep (ftl, ep), S=32 (launch-latency case)", d=2,
               N=400, S=32, F=300, n=12, t_max=32, mode="ep"),
    # C4: 10-D GP-prior objective, uniform candidates on [0,10]^10, 2^24 candidates over 8 GPUs = 2^21 per
    # GPU grouped consecutively into batches of q = 8, S = 1024, F = 4096, n = 256, batch TES_sp with T <= 32, I = 1, Y = 8
    "c4": dict(desc="C4 synthetic 10-D GP objective, batch TES_sp q=8, F=4096, 2^21 candidates x 1024 maximizer samples "
                    "per GPU (2^24 over 8 GPUs)", d=10, N=2 ** 21, S=1024, F=4096, n=256, t_max=32, criterion="sftl",
               q=8, niteration=1, nysample=8, ntestsample=1000, xmax=10.0),
    "c4m": dict(desc="C4 shapes, mid size (profiling)", d=10, N=2 ** 18, S=256, F=1024, n=256, t_max=32, criterion="sftl",
                q=8, niteration=1, nysample=8, ntestsample=1000, xmax=10.0),
    "c4s": dict(desc="C4 shapes scaled down (smoke)", d=10, N=2 ** 15, S=64, F=512, n=32, t_max=8, criterion="sftl",
                q=8, niteration=1, nysample=4, ntestsample=500, xmax=10.0),
}
for _w in WORKLOADS.values():
    _w.setdefault("criterion", "ftl")
    _w.setdefault("mode", "empirical")
    _w.setdefault("q", 1)
    _w.setdefault("niteration", 10)
    _w.setdefault("nysample", 10)
    _w.setdefault("ntestsample", max(1000, 64 * _w["t_max"]))
    _w.setdefault("xmax", 1.0)


def gp_prior_objective(xdim, l, sigma, seed, n_feats=10000):
    """functions.func_gp_prior (functions.py:163-177), restated for the harness: one RFF prior sample."""
    rs = np.random.RandomState(seed)
    W = rs.randn(n_feats, xdim) * np.sqrt(l)
    b = 2.0 * np.pi * rs.rand(n_feats, 1)
    theta = rs.randn(n_feats, 1)
    return lambda x: (np.sqrt(2.0 * sigma / n_feats) * theta.T.dot(np.cos(W.dot(x.reshape(-1, xdim).T) + b))).reshape(-1)


def config_of(wl):
    c = {"workload": wl["desc"], "criterion": wl["criterion"], "mode": "sample" if wl["criterion"] == "sftl" else wl["mode"],
         "deterministic": True, "N_per_gpu": wl["N"], "S": wl["S"], "F": wl["F"], "n_obs": wl["n"], "d": wl["d"],
         "t_max": wl["t_max"]}
    if wl["criterion"] == "sftl":
        c.update(q=wl["q"], niteration=wl["niteration"], nysample=wl["nysample"], ntestsample=wl["ntestsample"])
    return c


def step_kwargs(wl):
    return dict(criterion=wl["criterion"], mode=wl["mode"], deterministic=True, ntestsample=wl["ntestsample"],
                t_max=wl["t_max"], q=wl["q"], niteration=wl["niteration"], nysample=wl["nysample"])


def make_workload(name, rank, world, host_only=False):
    wl = dict(WORKLOADS[name])
    d, N, S, F, n = wl["d"], wl["N"], wl["S"], wl["F"], wl["n"]
    rs = np.random.RandomState(1)
    X = rs.rand(n, d) * wl["xmax"]
    if d == 6:
        hyp = HART6
        f = hartmann6(X)
        Y = (f - f.mean()).reshape(-1, 1)
    elif d == 10:
        hyp = dict(l=np.full(10, 1.0), sigma=2.0, sigma0=1e-3)   # functions.func_gp_prior(10, l=1.0, sigma=2.0, seed=1)
        Y = gp_prior_objective(10, 1.0, 2.0, 1)(X).reshape(-1, 1)
    else:
        hyp = dict(l=np.array([177.9418631280815, 309.4630784148164]), sigma=0.29444226420446995,
                   sigma0=0.06476473751595126)
        Y = rs.randn(n, 1) * 0.5
    if d == 6 and N == 10 ** 6 and not wl.get("uniform"):
        cand = meshgrid(0.0, 1.0, 10, 6, slab=rank, nslab=world)      # every rank: a 10^6-point slab of the global mesh
    elif d == 2 and N == 400:
        cand = meshgrid(0.0, 1.0, 20, 2)
    else:
        cand = np.random.RandomState(7 + rank).rand(N, d) * wl["xmax"]
    np.random.seed(11)  # identical draws on every rank
    # The three global-RNG draws of the RFF posterior weight draw (optfunc.py:329, :334, :344), in the reference's
    # order: the INPUTS of a step.  Stage A0 (theta/W/b from them, optfunc.py:303-385) runs INSIDE the timed step, on
    # the device for the b200 arm and as the oracle's numpy restatement (where the reference runs it) for the CPU arm.
    # The CPU arm only ever uses the first `ns` function samples (cpu_sample_sizes), so it only draws those.
    if host_only:
        S = min(S, cpu_sample_sizes(wl)[1])
    randn_W, rand_b, noise = np.random.randn(S, F, d), np.random.rand(S, F, 1), np.random.randn(S, F, 1)
    wl.update(name=name, hyp=hyp, X=X, Y=Y, cand=np.ascontiguousarray(cand), randn_W=randn_W, rand_b=rand_b,
              noise=noise)
    return wl


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w": float(np.median(pw)) if pw else None}


def cpu_step(wl, n_cand, n_samp, seed=0):
    """The CPU oracle port of one step on a bounded sub-sample: first n_cand candidates x first n_samp
    function samples of the same workload (stage A: oracle/tes_oracle.c with OpenMP on every host core;
    stages B-D: numpy oracle with BLAS threads).  Returns (pairs, seconds, info)."""
    from oracle import pipeline_np
    cand = wl["cand"][:n_cand]
    t0 = time.perf_counter()
    kw = step_kwargs(wl)
    kw["ntestsample"] = min(kw["ntestsample"], max(1000, 64 * min(wl["t_max"], n_samp)))
    out = pipeline_np.tes_step(cand, wl["X"], wl["Y"], wl["hyp"]["l"], wl["hyp"]["sigma"], wl["hyp"]["sigma0"],
                               None, None, None, seed=seed,
                               draws=(wl["randn_W"][:n_samp], wl["rand_b"][:n_samp], wl["noise"][:n_samp]), **kw)
    dt = time.perf_counter() - t0
    return cand.shape[0] * n_samp, dt, out


def cpu_sample_sizes(wl):
    # ~10-30 s of CPU work on a few cores: F*(d+13) fp64 ops per pair at ~1e8 feature-evals/s/core
    if wl.get("name", "") == "c3" or (wl["d"] == 6 and wl["N"] == 10 ** 6):
        return 16384, 128
    if wl["q"] > 1:  # the oracle materialises the (Y, S', S', nx, K, q) tensor of evaluate_batch_sample_mp.py:347-365
        return min(wl["N"], 256), min(wl["S"], 64)
    return min(wl["N"], 4096), min(wl["S"], 64)


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    nc, ns = cpu_sample_sizes(wl)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_step(wl, nc, ns)
    times, pairs = [], 0
    for s in range(args.steps):
        pairs, dt, info = cpu_step(wl, nc, ns, seed=s)
        times.append(dt)
    tmean = float(np.mean(times))
    val = pairs / tmean
    sample = ("first %d candidates x first %d function samples of %s (F=%d, n=%d), full step A0-D (weight draw included); "
              "threads: OMP_NUM_THREADS=%s OPENBLAS_NUM_THREADS=%s os.cpu_count()=%d" % (
                  nc, ns, wl["name"], wl["F"], wl["n"], os.environ.get("OMP_NUM_THREADS"),
                  os.environ.get("OPENBLAS_NUM_THREADS"), os.cpu_count()))
    line = {
        "impl": "reference", "metric": "acq evals/s (candidates x maximizer samples)", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tmean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wl),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": CPU_THREADS, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


_REAL_STDOUT_FD = None


def _stdout_to_stderr():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on the first
    communicator when NCCL_DEBUG is set): file descriptor 1 points at stderr until the line is printed."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(_REAL_STDOUT_FD, 1)
    sys.stdout.write(json.dumps(line) + "\n")
    sys.stdout.flush()


STAGES = [("draw", "A0_draw_ms"), ("rff_topk", "A_maximizers_ms"), ("stage_b", "B_posterior_ms"),
          ("stage_c", "C_maximizer_stats_ms"), ("criterion", "D_criterion_ms"), ("criterion_screen", "D_screen_ms")]
REPLICATED = ["X", "Y", "randn_W", "rand_b", "noise"]   # identical on every rank; "cand" is the rank's own shard
DRAWS = ["randn_W", "rand_b", "noise"]                  # ... of which every rank consumes the slice of its S / world samples


class Bench:
    """One process per GPU.  Holds the distributed context and runs timed loops of acquisition.tes_step."""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def gather(self, vec):
        """(k,) floats of this rank -> (world, k) numpy on every rank."""
        t = self.torch.tensor(vec, dtype=self.torch.float64, device="cuda")
        if self.world == 1:
            return t.cpu().numpy()[None]
        out = self.torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device="cuda")
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()

    def upload(self, host, names):
        """Pinned host buffers -> device.  The rank's candidate shard comes over its own PCIe link; the REPLICATED
        operands cross PCIe once (rank 0) and reach the other GPUs through one NCCL broadcast over NVLink."""
        torch = self.torch
        out = {}
        for k in names:
            if k == "cand" or k in DRAWS:
                # pinned host tensors: tes_step uploads the candidates on a side stream under the weight draw, and of the
                # draws only the slice of function samples this rank draws the weights of (S / world)
                out[k] = host[k]
                continue
            if self.world > 1 and k in REPLICATED:
                t = host[k].cuda(non_blocking=True) if self.rank == 0 else \
                    torch.empty(host[k].shape, dtype=host[k].dtype, device="cuda")
                self.dist.broadcast(t, 0)
                out[k] = t
            else:
                out[k] = host[k].cuda(non_blocking=True)
        return out

    def run(self, wl, steps, warmup, deterministic=True, want_e2e=True, sample_clocks=False):
        """Timed loop over one workload -> dict (value, ms, stages, e2e, launches, last step's out)."""
        torch = self.torch
        from tes_b200 import _lib, acquisition, ops
        hyp, kw = wl["hyp"], step_kwargs(wl)
        kw["deterministic"] = deterministic
        shard = acquisition.Shard(wl["N"])
        names = ["cand"] + REPLICATED
        host = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).pin_memory() for k in names}
        devt = {k: host[k].cuda(non_blocking=True) for k in names}
        split = self.world > 1 and wl["S"] % self.world == 0
        h2d_bytes = sum(host[k].numel() * 8 // (self.world if (split and k in DRAWS) else 1) for k in names)

        def step(d_, seed, timers=None, **extra):
            k2 = dict(kw, **extra)
            return acquisition.tes_step(d_["cand"], d_["X"], d_["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"], None, None,
                                        None, seed=seed, shard=shard, timers=timers,
                                        draws=(d_["randn_W"], d_["rand_b"], d_["noise"]), **k2)

        for w in range(warmup):
            self.flush.zero_()
            out = step(devt, w)
        self.barrier()
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        timers = {}
        ops.STATS["product_flagged"] = 0
        ops.STATS["product_unresolved"] = 0
        launches0 = _lib.lib().tes_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for s in range(steps):
            self.flush.zero_()  # L2 flush between timed iterations (its ~0.05 ms is inside the timed region)
            out = step(devt, 1000 + s, timers)
        e1.record()
        self.barrier()
        launches = _lib.lib().tes_launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        dev_s = self.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
        pairs_step = shard.n_global * wl["S"]
        res = {"value": pairs_step * steps / dev_s, "ms_per_step": dev_s / steps * 1e3, "steps": steps, "out": out,
               "launches": int(launches), "clocks": clocks, "shard": shard, "devt": devt, "step": step,
               "flagged_samples_per_step": ops.STATS["product_flagged"] / steps,
               "unresolved_samples_per_step": ops.STATS.get("product_unresolved", 0) / steps, "timers": timers}
        # per-stage milliseconds (CUDA events on the launching stream), mean over steps; min / max over ranks
        mine = [float(np.mean([a.elapsed_time(b_) for a, b_ in timers[k]])) if timers.get(k) else float("nan")
                for k, _ in STAGES]
        allr = self.gather(mine + [clocks["power_w"] if clocks and clocks.get("power_w") else float("nan")])
        res["stages"] = {lbl: {"rank0": allr[0, i], "min": float(np.nanmin(allr[:, i])), "max": float(np.nanmax(allr[:, i]))}
                         if self.world > 1 else allr[0, i] for i, (_, lbl) in enumerate(STAGES)
                         if not np.all(np.isnan(allr[:, i]))}
        res["t_stage_a"] = mine[1] * 1e-3
        res["t_criterion"] = mine[4] * 1e-3
        if not want_e2e:
            return res

        # ---- e2e: the same step from pinned HOST buffers, H2D of every input + D2H of the result inside the region
        def step_e2e(seed):
            d_ = self.upload(host, names)
            o = step(d_, seed)
            _ = o["argmax_idx"].cpu()  # D2H: the S trusted-maximizer indices (+ query idx/value/point already on host)
            return o
        step_e2e(0)
        self.barrier()
        t0 = time.perf_counter()
        e0.record()
        for s in range(steps):
            self.flush.zero_()
            step_e2e(2000 + s)
        e1.record()
        self.barrier()
        e2e_s = self.max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0))
        res["e2e"] = {"value": pairs_step * steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(h2d_bytes),
                      "d2h_bytes_per_step": int(wl["S"] * 8 + 8 + 8 + wl["d"] * wl["q"] * 8),
                      "ms_per_step": e2e_s / steps * 1e3,
                      "replicated_inputs": "observations: rank 0 H2D + NCCL broadcast; draws: every rank uploads the slice of "
                                            "its S / N function samples (the weight draw is sharded by sample, theta / W / b "
                                            "all-gathered)" if self.world > 1 else "H2D"}
        return res


def roofline_of(wl, res, product, peaks, fp64_peak, fp32_peak, mufu_peak):
    """Roofline object of the dominant kernel group (stage A) from the CUDA-event time of the timed region."""
    t_rff, t_crit = res["t_stage_a"], res["t_criterion"]
    d, F = wl["d"], wl["F"]
    local_pairs = wl["N"] * wl["S"]
    algo_excl_cos = local_pairs * F * (2 * d + 3) / t_rff * 1e-12  # SURVEY 8(d) figure, cos not counted
    screened = wl["N"] >= 32768 and d <= 10
    common = {"launch_ms": t_rff * 1e3, "share_of_step": t_rff * 1e3 / res["ms_per_step"],
              "criterion_ms": t_crit * 1e3, "algorithmic_excl_cos_tflops": algo_excl_cos,
              "fp64_fma_peak_tflops": fp64_peak, "mufu_peak_tcos": mufu_peak}
    if product:
        # stage A on a product-structured candidate set = one GEMM per function sample (csrc/rffprod.cu), run as a
        # 3xFP16 screen on the tensor cores (tcgen05.mma kind::f16 + TMEM) followed by the exact fp64 re-evaluation
        # of the k + 8 best candidates per sample.  algorithmic: 2 * N * 2F flops per sample (the fp64 GEMM being
        # emulated); executed: 3x that in fp16.
        bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
        gflops = 2.0 * local_pairs * 2 * F
        achieved = 3.0 * gflops / t_rff * 1e-12
        return dict(common, **{
            "bound": "tensor",
            "pipe": "5th-gen tensor cores: tcgen05.mma.cta_group::2 kind::f16 (M256 N256 K16 per CTA pair), accumulators in "
                    "TMEM, operands by 1-D bulk async copies (TMA engine) in a 3-stage mbarrier ring",
            "kernel": "stage-A group: operand build + prod_gemm_topk_tc8_kernel (GEMM + top-k epilogue) + list merge + "
                      "exact fp64 refine",
            "kernel_alone": "455.6 us per 37-sample launch = 998 TFLOP/s executed = 0.75 of the peak, tensor pipe 52 % busy "
                            "(ncu, profiles/r02_ncu_prod_gemm_tc8.md)",
            "achieved": achieved, "peak": bf16, "unit": "TFLOP/s", "frac": achieved / bf16,
            "achieved_is": "EXECUTED 3xFP16 tensor flops (12F per pair) / stage-A time (CUDA events, timed region)",
            "frac_algorithmic": gflops / t_rff * 1e-12 / bf16,
            "algorithmic_is": "4F flops per pair (the fp64 GEMM the 3xFP16 split emulates) against the same tensor peak; "
                              "SURVEY 8d's F*(2d+3) flops + F cos per pair is algorithmically different work (this path "
                              "replaces the cosines by the angle-addition GEMM) and is reported as algorithmic_excl_cos_tflops",
            "traffic": PROD_TRAFFIC_BYTES.get(wl["name"]),
            "peak_source": "bf16_tflops_sustained of MEASURED_PEAKS.json (cuBLAS, %s)" % (
                "measured" if peaks else "fallback 1400"),
            "algorithmic_fp64_gemm_tflops": gflops / t_rff * 1e-12,
            "equivalent_tcos": local_pairs * F / t_rff * 1e-12,
            "per_unit": "product-structured candidates (slow columns [%d, %d), nv=%d): 4F GEMM flops per (candidate, "
                        "sample) pair (12F executed as 3xFP16) instead of F cosines" % product})
    if screened and d >= 5:
        cos_rate = local_pairs * F / t_rff * 1e-12
        return dict(common, **{
            "bound": "mufu", "kernel": "rff_screen_tc5_kernel (+ pack/xmax/bound/refine/fallback/merge)",
            "achieved": cos_rate, "peak": mufu_peak, "unit": "Tcos/s", "frac": cos_rate / mufu_peak,
            "frac_algorithmic": cos_rate / mufu_peak,
            "traffic": TC5_TRAFFIC_BYTES.get(wl["name"]),
            "peak_source": "measured in this run: tes_mufu_cos_burn (__cosf = FMUL.RZ + MUFU.COS, 8 chains/thread)",
            "fp32_peak_tflops": fp32_peak, "algorithmic_vs_fp64_peak": algo_excl_cos / fp64_peak,
            "per_unit": "F cosines (7/8 MUFU.COS, 1/8 FMA-pipe polynomial) + F*(3d+2) TF32 tcgen05 MACs per (candidate, "
                        "sample) pair; algorithmic = F*(2d+3) flops + F cos (SURVEY 8d)"})
    if screened:
        ops32 = local_pairs * F * (d + 5)
        achieved = 2.0 * ops32 / t_rff * 1e-12
        return dict(common, **{
            "bound": "fp32", "kernel": "rff_screen_kernel (+ pack/xmax/bound/refine/fallback/merge)",
            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": None,
            "peak_source": "measured in this run: tes_fp32_fma_burn (packed FFMA2, 8 chains/thread, 1 s)",
            "algorithmic_vs_fp64_peak": algo_excl_cos / fp64_peak, "mufu_cos_per_s": local_pairs * F / t_rff,
            "per_unit": "F*(d+5) fp32-pipe lane ops + F MUFU.COS per (candidate, sample) pair; "
                        "algorithmic = F*(2d+3) flops + F cos (SURVEY 8d)"})
    instr = local_pairs * F * (d + 13)           # fp64-pipe instructions per launch (tes_spec.h sequence)
    achieved = 2.0 * instr / t_rff * 1e-12       # FMA-equivalent TFLOP/s, same convention as the DFMA burn
    return dict(common, **{
        "bound": "fp64", "kernel": "rff_topk_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved / fp64_peak, "traffic": None,
        "peak_source": "measured in this run: tes_fp64_fma_burn (8 DFMA chains/thread, 1 s sustained)",
        "per_unit": "F*(d+13) fp64-pipe instr per (candidate, sample) pair = F*(2d+3) flops + F cos"})


def parity_check(bench, wl, res):
    """Outside the timed region: the screened step against the all-fp64 step, and the product-path stage A against
    the general kernel on the first samples -- the fast paths must select the same things."""
    torch = bench.torch
    from tes_b200 import ops, optfunc
    out = {}
    devt, step = res["devt"], res["step"]
    if wl["criterion"] == "ftl":
        a = step(devt, 4242)
        b = step(devt, 4242, screen=False)
        same = (a["query_idx"] == b["query_idx"] and a["query_val"] == b["query_val"]
                and np.array_equal(a["query_x"], b["query_x"]))
        out["screened_step_vs_fp64_step"] = {"identical_query": bool(same), "query_idx": int(a["query_idx"]),
                                             "query_val": a["query_val"], "screen_kept": a.get("screen_kept"),
                                             "trusted_maximizers_identical": bool(torch.equal(a["argmax_idx"], b["argmax_idx"]))}
    product = ops.detect_product_structure(devt["cand"]) if wl["N"] >= ops.PRODUCT_MIN_N else None
    if product:
        ns = min(16, wl["S"])
        theta, W, b_ = optfunc.draw_random_init_weights_features(
            wl["d"], ns, wl["F"], devt["X"], devt["Y"], wl["hyp"]["l"], wl["hyp"]["sigma"], wl["hyp"]["sigma0"],
            draws=(devt["randn_W"][:ns], devt["rand_b"][:ns], devt["noise"][:ns]))
        i1, v1 = ops.rff_topk(devt["cand"], W, b_, theta, wl["hyp"]["sigma"], k=3, product=product)
        i2, v2 = ops.rff_topk(devt["cand"], W, b_, theta, wl["hyp"]["sigma"], k=3, product=False)
        out["product_stage_a_vs_general_kernel"] = {"samples": ns, "k": 3, "indices_identical": bool(torch.equal(i1, i2)),
                                                    "values_bit_identical": bool(torch.equal(v1, v2))}
    return out


def c5_sweep(bench, sizes=(64, 128, 256, 512, 1024, 2048), batch=1024, reps=2):
    """BASELINE config 5: batched maximizer-conditioned chol2inv (utils.py:1859-1883), batch sharded over the ranks
    without any collective.  One record per m (scripts/bench_chol_sweep.py holds the stand-alone version)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import bench_chol_sweep as cs
    return cs.sweep(bench.torch, bench.dist, bench.rank, bench.world, list(sizes), batch, reps, cpu=False,
                    flush=bench.flush)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--stochastic", action="store_true", help="TES_ep stochastic estimator (the reference's default, "
                    "bo_discrete.py --deterministic 0) instead of the deterministic one")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads reported beside C3")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check after the timed region (profiling runs)")
    args = ap.parse_args()
    _stdout_to_stderr()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        wl = make_workload(args.workload, 0, 1, host_only=True) if rank == 0 else None
        run_reference_arm(args, wl, rank)
        return

    import torch
    from tes_b200 import ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    bench = Bench(rank, world, local_rank)
    wl = make_workload(args.workload, rank, world)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # ---- in-run pipe peaks (roofline denominators that MEASURED_PEAKS.json does not carry)
    fp64_peak = ops.fp64_fma_peak(1.0)
    fp32_peak = ops.fp32_fma_peak(1.0)
    mufu_peak = ops.mufu_cos_peak(1.0)

    res = bench.run(wl, args.steps, args.warmup, deterministic=not args.stochastic, sample_clocks=True)
    out = res["out"]
    product = ops.detect_product_structure(res["devt"]["cand"]) if wl["N"] >= ops.PRODUCT_MIN_N else None
    roofline = roofline_of(wl, res, product, peaks, fp64_peak, fp32_peak, mufu_peak)
    parity = None if args.no_parity else parity_check(bench, wl, res)

    line = None
    if rank == 0:
        cfg = config_of(wl)
        cfg["deterministic"] = not args.stochastic
        line = {
            "metric": "acq evals/s (candidates x maximizer samples)", "value": res["value"], "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(cfg, T=out.get("T"), Sp=out.get("Sp"), K=out.get("K"),
                           step="A0 weight draw (optfunc.py:303-385) + A maximizers + B posteriors + C maximizer "
                                "statistics + D criterion and arg-max; inputs = candidates, observations and the three "
                                "global-RNG draws",
                           criterion_screen_kept=out.get("screen_kept"),
                           n_unique_maximizers=out.get("n_unique_maximizers"),
                           l2="flushed between timed iterations (256 MiB memset)",
                           candidates=("product set U x V detected (slow columns [%d, %d), nv=%d): stage A = GEMM path "
                                       "(3xFP16 tensor-core screen + fp64 refine)" % product)
                           if product else "unstructured: stage A = tcgen05 screen + fp64 refine (MUFU-bound)",
                           sharding="candidates contiguous per rank; weight draw sharded by function sample (theta / W / b all-gathered); NCCL all-gather of (value,index) partials"),
            "clocks": res["clocks"],
            "stage_ms": res["stages"],
            "stage_a_flagged_samples_per_step": res["flagged_samples_per_step"],
            "stage_a_unresolved_samples_per_step": res["unresolved_samples_per_step"],
            "e2e": res["e2e"],
            "gpu_launches": res["launches"],
            "roofline": roofline,
            "parity_check": parity,
        }
    del res
    torch.cuda.empty_cache()
    secondary = {}
    if args.workload == "c3" and not args.no_secondary and not args.stochastic:
        # ---- BASELINE config 4 (north star): batch TES_sp q=8, F=4096, 2^21 candidates per GPU (2^24 at 8 GPUs)
        w4 = make_workload("c4", rank, world)
        r4 = bench.run(w4, 3, 1, want_e2e=False)
        if rank == 0:
            secondary["c4"] = {"workload": w4["desc"], "value": r4["value"], "unit": "evals/s",
                               "ms_per_step": r4["ms_per_step"], "steps": 3, "warmup": 1, "n_gpus": world,
                               "global_candidates": r4["shard"].n_global, "stage_ms": r4["stages"],
                               "T": r4["out"].get("T"), "K": r4["out"].get("K"),
                               "roofline": {k: v for k, v in roofline_of(w4, r4, None, peaks, fp64_peak, fp32_peak,
                                                                         mufu_peak).items()
                                            if k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "share_of_step")}}
        del w4, r4
        torch.cuda.empty_cache()
        # ---- C3 with the reference's DEFAULT estimator (bo_discrete.py --deterministic 0: evaluate_mp.py:220-295)
        r3 = bench.run(wl, 2, 1, deterministic=False, want_e2e=False)
        if rank == 0:
            secondary["c3_stochastic"] = {"workload": wl["desc"] + " -- stochastic estimator, I=%d Y=%d" % (
                wl["niteration"], wl["nysample"]), "value": r3["value"], "unit": "evals/s",
                "ms_per_step": r3["ms_per_step"], "steps": 2, "warmup": 1, "stage_ms": r3["stages"]}
        del r3
        torch.cuda.empty_cache()
        # ---- the general case: C3 shapes on UNSTRUCTURED candidates (no product path: tcgen05 screen, MUFU-bound)
        wu = make_workload("c3u", rank, world)
        ru = bench.run(wu, 3, 1, want_e2e=False)
        if rank == 0:
            secondary["c3u"] = {"workload": wu["desc"], "value": ru["value"], "unit": "evals/s",
                                "ms_per_step": ru["ms_per_step"], "steps": 3, "warmup": 1, "n_gpus": world,
                                "stage_ms": ru["stages"],
                                "roofline": {k: v for k, v in roofline_of(wu, ru, None, peaks, fp64_peak, fp32_peak,
                                                                          mufu_peak).items()
                                             if k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "share_of_step")}}
        del wu, ru
        torch.cuda.empty_cache()
        if world > 1:
            # ---- strong scaling: the SAME global 10^6-point mesh split over the ranks
            ws = make_workload("c3", 0, 1)
            per = ws["N"] // world
            ws["cand"] = np.ascontiguousarray(ws["cand"][rank * per:(rank + 1) * per])
            ws["N"] = per
            rs_ = bench.run(ws, 5, 2, want_e2e=False)
            if rank == 0:
                secondary["c3_strong"] = {"workload": "C3 mesh, 10^6 candidates in total split over %d GPUs" % world,
                                          "scaling": "strong", "value": rs_["value"], "unit": "evals/s",
                                          "ms_per_step": rs_["ms_per_step"], "steps": 5, "stage_ms": rs_["stages"]}
            del ws, rs_
            torch.cuda.empty_cache()
        # ---- BASELINE config 5: posterior scaling sweep
        c5 = c5_sweep(bench)
        if rank == 0:
            secondary["c5"] = c5
    if rank == 0 and world == 1 and args.workload == "c3" and not args.no_secondary and not args.stochastic:
        # BASELINE config 2 (the 400-point pH grid, TES_ep mode ep): a launch-latency case
        w2 = make_workload("c2", 0, 1)
        r2 = bench.run(w2, 20, 3, want_e2e=False)
        t0 = time.perf_counter()
        cpu_step(w2, w2["N"], w2["S"])
        cpu2 = time.perf_counter() - t0
        secondary["c2"] = {"workload": w2["desc"], "ms_per_step": r2["ms_per_step"], "value": r2["value"],
                           "unit": "evals/s", "steps": 20, "T": r2["out"].get("T"),
                           "gpu_launches_per_step": r2["launches"] / 20, "stage_ms": r2["stages"],
                           "cpu_ms_per_step": cpu2 * 1e3, "cpu_cores": os.cpu_count(),
                           "note": "whole step incl. weight draw and EP on the device; host-latency bound"}
    if rank == 0 and secondary:
        line["secondary"] = secondary
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nc, ns = cpu_sample_sizes(wl)
        pairs, dt, info = cpu_step(wl, nc, ns)
        line["cpu_baseline"] = {"value": pairs / dt, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "first %d candidates x first %d function samples, full step A0-D, %.1f s"
                                          % (nc, ns, dt)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        _emit(line)
    if world > 1:
        bench.dist.destroy_process_group()


if __name__ == "__main__":
    main()
