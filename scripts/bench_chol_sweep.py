#!/usr/bin/env python
"""C5 of BASELINE.json -- posterior scaling sweep: `batch` maximizer-conditioned matrices
K([x*_j; X]) + sigma0 diag(0, I) of size m x m (utils.eval_invKmaxsams, utils.py:1859-1883), inverted through
tes_batched_chol2inv_f64 (utils.chol2inv, utils.py:657-674), for m in {64 ... 2048}.

One JSON line per m: matrices/s, the algorithmic rates of SURVEY 8(d) (m^3/3 + m^3 flops and 2 m^2 8 bytes per
matrix) against the measured fp64 / HBM peaks, and a CPU baseline (the numpy oracle's chol2inv on a few matrices of
the same stack, all host cores through BLAS).  The batch is sharded across ranks without any collective
(torchrun --nproc-per-node N); the reported rate is the whole job's (max time over ranks).

  python scripts/bench_chol_sweep.py [--batch 1024] [--sizes 64 128 ...] [--reps 3] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HART6_L = np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949])   # functions.py:540-543
SIGMA, SIGMA0 = 1.423, 1e-2                                             # sigma0 raised for conditioning


def build_stack(torch, m, batch, first, device):
    """(batch, m, m): X ~ U[0,1]^6 shared ((m-1) rows), x*_j the j-th maximizer; K = sigma exp(-1/2 sum l (x-x')^2)."""
    g = torch.Generator(device="cpu").manual_seed(1234 + m)
    X = torch.rand(m - 1, 6, dtype=torch.float64, generator=g).to(device)
    xs = torch.rand(first + batch, 6, dtype=torch.float64, generator=g)[first:].to(device)
    l = torch.from_numpy(HART6_L).to(device)

    def k(a, b):
        d = ((a[:, None, :] - b[None, :, :]) ** 2 * l).sum(-1)
        return SIGMA * torch.exp(-0.5 * d)
    A = torch.empty(batch, m, m, dtype=torch.float64, device=device)
    A[:, 1:, 1:] = k(X, X) + SIGMA0 * torch.eye(m - 1, dtype=torch.float64, device=device)
    kx = k(xs, X)
    A[:, 0, 1:] = kx
    A[:, 1:, 0] = kx
    A[:, 0, 0] = SIGMA
    return A


def sweep(torch, dist, rank, world, sizes, batch, reps, cpu=True, flush=None, emit=None):
    """-> list of one record per m (rank 0; [] elsewhere).  `emit(record)` is called as soon as a size is done."""
    from tes_b200 import _lib, ops
    dev = torch.device("cuda", torch.cuda.current_device())
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6535.1))
    fp64_peak = ops.fp64_fma_peak(0.5)
    if flush is None:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    records = []
    for m in sizes:
        local = batch // world + (1 if rank < batch % world else 0)
        first = rank * (batch // world) + min(rank, batch % world)
        sub = max(1, min(local, int(16e9 // (m * m * 8))))           # <= 16 GB per array and call
        chunks = [(first + o, min(sub, local - o)) for o in range(0, local, sub)]
        stacks = [build_stack(torch, m, nb, f0, dev) for f0, nb in chunks] if len(chunks) == 1 else None
        times, worst, launches = [], 0.0, 0
        inner = 16 if m <= 256 else 1
        for rep in range(reps + 1):                                  # first repetition = warm-up
            tot = 0.0
            for ci, (f0, nb) in enumerate(chunks):
                A = stacks[ci] if stacks else build_stack(torch, m, nb, f0, dev)
                flush.zero_()
                n0 = _lib.lib().tes_launch_count()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(inner):       # small m: several calls back to back so the device never waits on Python
                    inv, info = ops.batched_chol2inv(A, return_info=True)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1) * 1e-3 / inner
                launches = (_lib.lib().tes_launch_count() - n0) // inner
                if rep == 0 and ci == 0:                             # residual check on the first matrices
                    assert int(info.abs().sum()) == 0
                    r = (A[:4] @ inv[:4] - torch.eye(m, dtype=torch.float64, device=dev)).abs().max()
                    worst = float(r)
                del inv, info
                if not stacks:
                    del A
            if rep:
                times.append(tot)
        t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = float(t[0])
        flops = batch * (m ** 3 / 3.0 + m ** 3)
        byts = batch * 2.0 * m * m * 8
        line = {"metric": "C5 batched chol2inv", "m": m, "batch": batch, "n_gpus": world,
                "value": batch / t, "unit": "matrices/s", "ms": t * 1e3, "calls_per_rank": len(chunks), "calls_per_timing": inner,
                "launches_per_call": int(launches), "max_residual_A_Ainv_minus_I": worst,
                "roofline": {"bound": "hbm" if m <= 128 else "fp64",
                             "algorithmic_tflops": flops / t * 1e-12, "fp64_peak_tflops": fp64_peak * world,
                             "frac_fp64": flops / t * 1e-12 / (fp64_peak * world),
                             "algorithmic_gbs": byts / t * 1e-9, "hbm_peak_gbs": hbm_peak * world,
                             "frac_hbm": byts / t * 1e-9 / (hbm_peak * world),
                             "per_unit": "m^3/3 + m^3 flops, 2 m^2 8 bytes per matrix (SURVEY 8d)"}}
        if rank == 0 and cpu:
            from oracle import tes_oracle_np as onp
            nb = max(1, min(8, int(2e9 // (m ** 3))))
            Ah = build_stack(torch, m, nb, 0, dev).cpu().numpy()
            t0 = time.perf_counter()
            onp.multichol2inv(Ah)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nb / dt, "unit": "matrices/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "%d matrices of the same stack through the numpy oracle's multichol2inv" % nb}
        if rank == 0:
            records.append(line)
            if emit:
                emit(line)
        del stacks
        torch.cuda.empty_cache()
    return records


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--sizes", dest="m", type=int, nargs="*", default=[64, 128, 256, 512, 1024, 2048])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    sweep(torch, dist, rank, world, args.m, args.batch, args.reps, cpu=not args.no_cpu,
          emit=lambda line: print(json.dumps(line), flush=True))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
