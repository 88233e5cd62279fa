"""Dev: where the C3 step spends the time that no stage timer covers (host-side wall clock between stages)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tes_b200 import acquisition, ops

wl = bench.make_workload("c3", 0, 1)
dv = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).cuda() for k in ("cand", "X", "Y", "randn_W", "rand_b", "noise")}
hyp, kw = wl["hyp"], bench.step_kwargs(wl)
orig = {}
acc = {}
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(mod, name, g)
for mod, name in ((acquisition, "select_test_points"), (acquisition, "trusted_maximizers"), (ops, "argmax"), (ops, "topk_merge"),
                  (ops, "ep_acq_screen"), (ops, "detect_product_structure"), (ops, "dedup")):
    wrap(mod, name)
def step(seed):
    return acquisition.tes_step(dv["cand"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"], None, None, None, seed=seed,
                                draws=(dv["randn_W"], dv["rand_b"], dv["noise"]), **kw)
for w in range(3):
    step(w)
acc.clear()
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 5
for s in range(n):
    step(100 + s)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print("step (with the extra syncs of this script) %.2f ms" % (tot / n * 1e3))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("  %-28s %.3f ms" % (k, v / n * 1e3))
