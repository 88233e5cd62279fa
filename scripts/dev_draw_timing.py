"""Dev: time tes_rff_draw_f64 at C3 sizes (S=1024, F=1024, n=64, d=6); run under ncu for the per-kernel list."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tes_b200 import optfunc

wl = bench.make_workload("c3", 0, 1)
dv = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).cuda() for k in ("X", "Y", "randn_W", "rand_b", "noise")}
h = wl["hyp"]
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    optfunc.draw_random_init_weights_features(wl["d"], wl["S"], wl["F"], dv["X"], dv["Y"], h["l"], h["sigma"], h["sigma0"],
                                              draws=(dv["randn_W"], dv["rand_b"], dv["noise"]))
    e1.record(); torch.cuda.synchronize()
    print("draw %.3f ms" % e0.elapsed_time(e1))
