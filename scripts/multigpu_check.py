"""Run under torchrun with N >= 2 GPUs: the candidate-sharded TES step must give the SAME trusted maximizers,
acquisition arg-max and value as the unsharded step (computed redundantly on every rank).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      scripts/multigpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from tes_b200 import acquisition, optfunc

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rs = np.random.RandomState(0)
d, N, S, F, n = 6, 120_001, 64, 256, 24
l = np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949])
sigma, sigma0 = 1.423, 1e-3
X = rs.rand(n, d)
Y = rs.randn(n, 1) * 0.5
cand = rs.rand(N, d)
np.random.seed(3)
theta, W, b = optfunc.draw_random_init_weights_features_np(d, S, F, X, Y, l.reshape(1, d), sigma, sigma0)
kw = dict(criterion="ftl", mode="empirical", deterministic=False, ntestsample=2000, nysample=4, niteration=2, seed=9,
          t_max=24, return_acq=True)
# unsharded reference on every rank (a private 1-rank view)
solo = acquisition.Shard.__new__(acquisition.Shard)
solo.dist, solo.group, solo.world, solo.rank, solo.offset, solo.n_global = None, None, 1, 0, 0, N
ref = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, shard=solo, **kw)
# sharded
cuts = np.linspace(0, N, world + 1).astype(int)
lo, hi = cuts[rank], cuts[rank + 1]
shard = acquisition.Shard(hi - lo)
out = acquisition.tes_step(cand[lo:hi], X, Y, l, sigma, sigma0, theta, W, b, shard=shard, **kw)
ok = (torch.equal(out["argmax_idx"], ref["argmax_idx"]) and torch.equal(out["argmax_val"], ref["argmax_val"])
      and out["query_idx"] == ref["query_idx"] and out["query_val"] == ref["query_val"]
      and np.array_equal(out["query_x"], ref["query_x"])
      and torch.equal(out["acq"], ref["acq"][lo:hi]))   # Philox draws keyed on the GLOBAL candidate index
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_CHECK world=%d T=%d Sp=%d query_idx=%d ok=%s" % (world, out["T"], out["Sp"], out["query_idx"],
                                                                    bool(flag.item())))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
