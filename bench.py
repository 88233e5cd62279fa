#!/usr/bin/env python
"""bench.py -- TES acquisition throughput on B200 (BASELINE.json metric: acq evals/s,
1 eval = one (candidate, maximizer-sample) pair).

A "step" is ONE full acquisition step (stages A-D of SURVEY.md section 8) over one batch of synthetic
input.  Default workload = config C3 of BASELINE.json (the largest single-GPU configuration):
Hartmann-6 hyper-parameters, 10^6 candidates per GPU (the 10-per-axis mesh of functions.get_meshgrid; with N ranks the
global mesh has 10 N points along its slowest axis and every rank holds one 10^6-point slab), S = 1024 RFF function samples with F = 1024 features each, n = 64
observations, TES_ep (criterion ftl, mode empirical, deterministic) on the <= t_max most frequent
trusted maximizers.  theta/W/b are host-supplied draws (optfunc.py:303-385), inputs of the step.

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
TC5_TRAFFIC_BYTES = {"c3": 465.8e6}   # profiles/r01_ncu_rff_screen_tc5.md (algorithmic: 115 MB)

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
    # the RFF posterior weight draw (optfunc.py:303-385): on the device for the b200 arm; the CPU reference arm runs
    # the oracle's numpy restatement (where the reference runs it).  Same np.random draws, same order, either way.
    # The CPU arm only ever uses the first `ns` function samples (cpu_sample_sizes), so it only draws those.
    if host_only:
        from oracle import tes_oracle_np as onp
        S = min(S, cpu_sample_sizes(wl)[1])
        randn_W, rand_b, noise = np.random.randn(S, F, d), np.random.rand(S, F, 1), np.random.randn(S, F, 1)
        theta, W, b = onp.draw_random_init_weights_features(d, S, F, X, Y, hyp["l"].reshape(1, d), hyp["sigma"],
                                                            hyp["sigma0"], randn_W, rand_b, noise)
    else:
        from tes_b200 import optfunc
        randn_W, rand_b, noise = np.random.randn(S, F, d), np.random.rand(S, F, 1), np.random.randn(S, F, 1)
        theta, W, b = optfunc.draw_random_init_weights_features_np(d, S, F, X, Y, hyp["l"].reshape(1, d), hyp["sigma"],
                                                                   hyp["sigma0"], draws=(randn_W, rand_b, noise))
    wl.update(name=name, hyp=hyp, X=X, Y=Y, cand=np.ascontiguousarray(cand),
              theta=np.ascontiguousarray(theta.reshape(-1, F)), W=np.ascontiguousarray(W),
              b=np.ascontiguousarray(b.reshape(-1, F)))
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


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
                               wl["theta"][:n_samp].reshape(n_samp, wl["F"], 1), wl["W"][:n_samp],
                               wl["b"][:n_samp].reshape(n_samp, wl["F"], 1), seed=seed, **kw)
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
    cores = os.cpu_count()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_step(wl, nc, ns)
    times, pairs = [], 0
    for s in range(args.steps):
        pairs, dt, info = cpu_step(wl, nc, ns, seed=s)
        times.append(dt)
    tmean = float(np.mean(times))
    val = pairs / tmean
    sample = "first %d candidates x first %d function samples of %s (F=%d, n=%d), full step A-D" % (
        nc, ns, wl["name"], wl["F"], wl["n"])
    line = {
        "impl": "reference", "metric": "acq evals/s (candidates x maximizer samples)", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tmean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wl),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 latency case reported beside C3")
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
    import torch.distributed as dist
    from tes_b200 import _lib, acquisition, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = make_workload(args.workload, rank, world)
    hyp = wl["hyp"]
    kw = step_kwargs(wl)
    shard = acquisition.Shard(wl["N"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident inputs (for `value`) and pinned host inputs (for `e2e`)
    names = ["cand", "X", "Y", "theta", "W", "b"]
    host = {k: torch.from_numpy(wl[k]).pin_memory() for k in names}
    devt = {k: host[k].cuda(non_blocking=True) for k in names}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    h2d_bytes = sum(host[k].numel() * 8 for k in names)

    def step_device(seed, timers=None):
        return acquisition.tes_step(devt["cand"], devt["X"], devt["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"],
                                    devt["theta"], devt["W"], devt["b"], seed=seed, shard=shard, timers=timers, **kw)

    def step_e2e(seed):
        d_ = {k: host[k].cuda(non_blocking=True) for k in names}  # pinned host -> device, every step
        out = acquisition.tes_step(d_["cand"], d_["X"], d_["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"], d_["theta"],
                                   d_["W"], d_["b"], seed=seed, shard=shard, **kw)
        _ = out["argmax_idx"].cpu()  # D2H: the S trusted-maximizer indices (+ query idx/value/point already on host)
        return out

    # ---- fp64 pipe peak (roofline denominator; not in MEASURED_PEAKS.json)
    fp64_peak = ops.fp64_fma_peak(1.0)
    fp32_peak = ops.fp32_fma_peak(1.0)
    mufu_peak = ops.mufu_cos_peak(1.0)

    for w in range(args.warmup):
        flush.zero_()
        out = step_device(w)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    timers = {}
    launches0 = _lib.lib().tes_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush_ms = 0.0
    barrier()
    e0.record()
    for s in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (its ~0.05 ms is inside the timed region)
        out = step_device(1000 + s, timers)
    e1.record()
    barrier()
    launches = _lib.lib().tes_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    dev_s = e0.elapsed_time(e1) * 1e-3
    tt = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_s = float(tt[0])
    pairs_step = shard.n_global * wl["S"]
    value = pairs_step * args.steps / dev_s

    # ---- e2e with host buffers
    step_e2e(0)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for s in range(args.steps):
        flush.zero_()
        out2 = step_e2e(2000 + s)
    e1.record()
    barrier()
    e2e_s = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)
    tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s = float(tt[0])
    e2e_val = pairs_step * args.steps / e2e_s
    d2h_bytes = wl["S"] * 8 + 8 + 8 + wl["d"] * 8

    # ---- roofline of the dominant kernel (stage A: rff_pack + rff_topk + merge), CUDA events per launch
    ra = [a.elapsed_time(b_) * 1e-3 for a, b_ in timers.get("rff_topk", [])]
    rc = [a.elapsed_time(b_) * 1e-3 for a, b_ in timers.get("criterion", [])]
    t_rff = float(np.mean(ra)) if ra else float("nan")
    d, F = wl["d"], wl["F"]
    local_pairs = wl["N"] * wl["S"]
    algo_excl_cos = local_pairs * F * (2 * d + 3) / t_rff * 1e-12  # SURVEY 8(d) figure, cos not counted
    screened = wl["N"] >= 32768 and d <= 10
    product = ops.detect_product_structure(devt["cand"]) if wl["N"] >= ops.PRODUCT_MIN_N else None
    if product:
        # stage A on a product-structured candidate set = one GEMM per function sample (csrc/rffprod.cu), run as a
        # 3xFP16 screen on the tensor cores (tcgen05.mma kind::f16 + TMEM, hi/lo fp16 operands, fp32 segments of 256 k
        # drained from TMEM into fp64 sums)
        # followed by the exact fp64 re-evaluation of the k + 8 best candidates per sample.
        # algorithmic: 2 * N * 2F flops per sample (the fp64 GEMM); executed: 3x that in fp16
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
        tf32_peak = bf16                # fp16 dense tensor rate = the bf16 rate
        gflops = 2.0 * local_pairs * 2 * F
        achieved = 3.0 * gflops / t_rff * 1e-12
        roofline = {"bound": "tensor", "pipe": "5th-gen tensor cores: tcgen05.mma kind::f16 (M128 N128 K16), accumulators in TMEM, operands by "
                            "1-D bulk async copies (TMA engine) in a 3-stage mbarrier ring",
                    "kernel": "prod_gemm_topk_tc_kernel (+ operand build, list merge, exact fp64 refine)",
                    "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                    "traffic": 775.6e6 if wl["name"] == "c3" else None,   # profiles/r01_ncu_prod_gemm_tcgen05.md, per launch (read 754.5 + write 21.1 MB)
                    "kernel_alone": "745 us per 45-sample launch = 742 TFLOP/s = 0.56 of the peak (ncu, profiles/)",
                    "peak_source": "bf16_tflops_sustained of MEASURED_PEAKS.json (cuBLAS, %s)" % (
                        "measured" if peaks else "fallback 1400"),
                    "algorithmic_fp64_gemm_tflops": gflops / t_rff * 1e-12,
                    "fp64_fma_peak_tflops": fp64_peak, "mufu_peak_tcos": mufu_peak,
                    "algorithmic_excl_cos_tflops": algo_excl_cos,
                    "equivalent_tcos": local_pairs * F / t_rff * 1e-12,
                    "per_unit": "product-structured candidates (slow columns [%d, %d), nv=%d): 4F GEMM flops per (candidate, "
                                "sample) pair (12F executed as 3xFP16) instead of F cosines; SURVEY 8d algorithmic = "
                                "F*(2d+3) flops + F cos" % product,
                    "launch_ms": t_rff * 1e3, "share_of_step": t_rff * args.steps / dev_s,
                    "criterion_ms": float(np.mean(rc)) * 1e3 if rc else None}
    elif screened and d >= 5:
        # stage A = tcgen05/TMEM screen (3xTF32 dot products on the 5th-gen tensor cores, cosine on the MUFU with 1 of
        # every 8 pairs on the FMA pipe) + exact fp64 refine: one cosine per (candidate, sample, feature) on the XU
        # pipe (16 lanes/SM/clk) is the binding resource
        cos_rate = local_pairs * F / t_rff * 1e-12
        roofline = {"bound": "mufu", "kernel": "rff_screen_tc5_kernel (+ pack/xmax/bound/refine/fallback/merge)",
                    "achieved": cos_rate, "peak": mufu_peak, "unit": "Tcos/s", "frac": cos_rate / mufu_peak,
                    "traffic": TC5_TRAFFIC_BYTES.get(wl["name"]),
                    "peak_source": "measured in this run: tes_mufu_cos_burn (__cosf = FMUL.RZ + MUFU.COS, 8 chains/thread)",
                    "fp64_peak_tflops": fp64_peak, "fp32_peak_tflops": fp32_peak,
                    "algorithmic_excl_cos_tflops": algo_excl_cos,
                    "algorithmic_vs_fp64_peak": algo_excl_cos / fp64_peak,
                    "per_unit": "F cosines (7/8 MUFU.COS, 1/8 FMA-pipe polynomial) + F*(3d+2) TF32 tcgen05 MACs per (candidate, sample) pair; "
                                "algorithmic = F*(2d+3) flops + F cos (SURVEY 8d)",
                    "launch_ms": t_rff * 1e3, "share_of_step": t_rff * args.steps / dev_s,
                    "criterion_ms": float(np.mean(rc)) * 1e3 if rc else None}
    elif screened:
        # stage A = fp32 screen (all pairs, packed FFMA2 + MUFU) + exact fp64 refine of the kept candidates.
        # fp32-pipe lane operations per (pair, feature): d (dot) + 3 (2*pi reduction) + 1 (theta) packed FMA
        # lanes + 1 FMUL inside __cosf = d + 5, plus 1 MUFU.COS (XU pipe, reported separately).
        ops32 = local_pairs * F * (d + 5)
        achieved = 2.0 * ops32 / t_rff * 1e-12
        roofline = {"bound": "fp32", "kernel": "rff_screen_kernel (+ pack/xmax/bound/refine/fallback/merge)",
                    "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                    "traffic": None,
                    "peak_source": "measured in this run: tes_fp32_fma_burn (packed FFMA2, 8 chains/thread, 1 s)",
                    "fp64_peak_tflops": fp64_peak,
                    "algorithmic_excl_cos_tflops": algo_excl_cos,
                    "algorithmic_vs_fp64_peak": algo_excl_cos / fp64_peak,
                    "mufu_cos_per_s": local_pairs * F / t_rff,
                    "per_unit": "F*(d+5) fp32-pipe lane ops + F MUFU.COS per (candidate, sample) pair; "
                                "algorithmic = F*(2d+3) flops + F cos (SURVEY 8d)",
                    "launch_ms": t_rff * 1e3, "share_of_step": t_rff * args.steps / dev_s,
                    "criterion_ms": float(np.mean(rc)) * 1e3 if rc else None}
    else:
        instr = local_pairs * F * (d + 13)           # fp64-pipe instructions per launch (tes_spec.h sequence)
        achieved = 2.0 * instr / t_rff * 1e-12       # FMA-equivalent TFLOP/s, same convention as the DFMA burn
        roofline = {"bound": "fp64", "kernel": "rff_topk_kernel", "achieved": achieved, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": None,
                    "peak_source": "measured in this run: tes_fp64_fma_burn (8 DFMA chains/thread, 1 s sustained)",
                    "algorithmic_excl_cos_tflops": algo_excl_cos,
                    "per_unit": "F*(d+13) fp64-pipe instr per (candidate, sample) pair = F*(2d+3) flops + F cos",
                    "launch_ms": t_rff * 1e3, "share_of_step": t_rff * args.steps / dev_s,
                    "criterion_ms": float(np.mean(rc)) * 1e3 if rc else None}

    line = None
    if rank == 0:
        line = {
            "metric": "acq evals/s (candidates x maximizer samples)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_of(wl), T=out.get("T"), Sp=out.get("Sp"), K=out.get("K"),
                           criterion_screen_kept=out.get("screen_kept"),
                           criterion_screen_ms=(float(np.mean([a.elapsed_time(b_) for a, b_ in timers["criterion_screen"]]))
                                                if timers.get("criterion_screen") else None),
                           n_unique_maximizers=out.get("n_unique_maximizers"),
                           l2="flushed between timed iterations (256 MiB memset)",
                           candidates=("product set U x V detected (slow columns [%d, %d), nv=%d): stage A = GEMM path (3xFP16 tensor-core screen + fp64 refine)" % product)
                           if product else "unstructured: stage A = tcgen05 screen + fp64 refine (MUFU-bound)",
                           sharding="candidates contiguous per rank; NCCL all-gather of (value,index) partials only"),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
    if rank == 0 and world == 1 and args.workload == "c3" and not args.no_secondary:
        # BASELINE config 2 (the 400-point pH grid, TES_ep mode ep) in the same run: a launch-latency case, reported
        # beside the headline so that both single-GPU configurations are on the line
        w2 = make_workload("c2", 0, 1)
        d2 = {k: torch.from_numpy(w2[k]).cuda() for k in names}
        kw2, sh2 = step_kwargs(w2), acquisition.Shard(w2["N"])
        run2 = lambda sd: acquisition.tes_step(d2["cand"], d2["X"], d2["Y"], w2["hyp"]["l"], w2["hyp"]["sigma"],
                                               w2["hyp"]["sigma0"], d2["theta"], d2["W"], d2["b"], seed=sd, shard=sh2, **kw2)
        for w in range(3):
            o2 = run2(w)
        torch.cuda.synchronize()
        n2, l0 = 20, _lib.lib().tes_launch_count()
        e0.record()
        for s in range(n2):
            flush.zero_()
            o2 = run2(100 + s)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n2
        t0 = time.perf_counter()
        cpu_step(w2, w2["N"], w2["S"])
        cpu2 = time.perf_counter() - t0
        line["secondary"] = {"c2": {"workload": w2["desc"], "ms_per_step": ms2, "value": w2["N"] * w2["S"] / (ms2 * 1e-3),
                                    "unit": "evals/s", "steps": n2, "T": o2.get("T"),
                                    "gpu_launches_per_step": (_lib.lib().tes_launch_count() - l0) / n2,
                                    "cpu_ms_per_step": cpu2 * 1e3, "cpu_cores": os.cpu_count(),
                                    "note": "whole step incl. EP on the device; host-latency bound (launches + the bucket-count sync)"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nc, ns = cpu_sample_sizes(wl)
        pairs, dt, info = cpu_step(wl, nc, ns)
        line["cpu_baseline"] = {"value": pairs / dt, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "first %d candidates x first %d function samples, full step A-D, %.1f s"
                                          % (nc, ns, dt)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
