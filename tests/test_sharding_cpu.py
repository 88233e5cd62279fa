"""N > 1 host logic on CPU: two gloo ranks shard the candidate set, compute their partial top-k with the
oracle (standing in for the CUDA kernel, which needs a GPU), all-gather the partials and merge -- the result
must equal the unsharded oracle answer bit for bit and be identical on both ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle as co
    from tes_b200 import acquisition
    rs = np.random.RandomState(0)
    N, d, S, F, k = 1501, 3, 6, 40, 3
    cand = rs.rand(N, d)
    W = rs.randn(S, F, d) * 2
    b = 2 * np.pi * rs.rand(S, F, 1)
    theta = rs.randn(S, F, 1)
    cuts = [0, 700, N]                       # ragged shards
    lo, hi = cuts[rank], cuts[rank + 1]
    shard = acquisition.Shard(hi - lo)
    assert (shard.world, shard.rank, shard.offset, shard.n_global) == (world, rank, lo, N)
    idx, val = co.rff_topk(cand[lo:hi], W, b, theta, 1.3, k, idx_offset=shard.offset)
    mi, mv = acquisition.gather_and_merge(shard, torch.from_numpy(idx), torch.from_numpy(val))
    ridx, rval = co.rff_topk(cand, W, b, theta, 1.3, k)
    ok = np.array_equal(mi.numpy(), ridx) and np.array_equal(mv.numpy(), rval)
    # stage-D style arg-max merge: one (value, index) pair per rank, first maximum wins
    acq = np.sin(np.arange(N) * 0.37)
    acq[[100, 900]] = 2.0                    # exact tie across the two shards -> lower global index
    lv = torch.tensor([[[acq[lo:hi].max()]]])
    li = torch.tensor([[[int(np.argmax(acq[lo:hi])) + lo]]])
    qi, qv = acquisition.merge_topk_host(shard.all_gather(lv.reshape(1, 1)).reshape(-1, 1, 1),
                                         shard.all_gather(li.reshape(1, 1)).reshape(-1, 1, 1), 1)
    ok = ok and int(qi[0, 0]) == 100 and float(qv[0, 0]) == 2.0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_topk_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_merge_topk_host_semantics():
    from tes_b200 import acquisition
    vals = torch.tensor([[[3.0, 1.0]], [[3.0, 2.0]], [[float("-inf"), float("-inf")]]])
    idx = torch.tensor([[[5, 9]], [[2, 7]], [[-1, -1]]])
    oi, ov = acquisition.merge_topk_host(vals, idx, 3)
    assert oi.tolist() == [[2, 5, 7]] and ov.tolist() == [[3.0, 3.0, 2.0]]


def test_optimize_continuous_function_host_logic():
    """utils_for_continuous.optimize_continuous_function (:20-87) re-hosted: candidate scoring, top-k start points
    (value desc, index asc), TF-1 Adam with the clip constraint on central-difference gradients.  The objective here is
    a plain numpy function (the device criterion is exercised by tests/test_bo_loop_gpu.py)."""
    import numpy as np
    from tes_b200 import utils_for_continuous as ufc
    calls = []

    def f(x, step):
        calls.append((x.shape[0], step))
        return -((x - np.array([0.3, 0.8])) ** 2).sum(1)
    rs = np.random.RandomState(0)
    cand = rs.rand(12, 2)
    mx, val, xs, fs = ufc.optimize_continuous_function(2, f, cand, 3, 0.0, 1.0, ntrain=40, rand_candidate_xs=rs.rand(3, 2),
                                                       debug=True)
    assert calls[0] == (15, 0) and calls[1] == (3 * 4, 1) and calls[-1] == (3, 41)
    # Adam moves every coordinate by ~lr per step towards the optimum while the gradient keeps its sign
    assert mx.shape == (1, 2) and val == fs.max() and np.all(xs >= 0.0) and np.all(xs <= 1.0)
    assert val > np.sort(f(cand, -1))[-1] - 1e-12
    # exact TF-1 Adam on a linear objective: g = -1 => x_t = x_0 + sum_t lr_t m_t/(sqrt(v_t)+eps) = x_0 + t*lr (to 1e-8)
    lin = lambda x, step: x.sum(1)
    x, fx = ufc.adam_ascent(lin, np.array([[0.1, 0.2]]), -np.inf, np.inf, 10)
    np.testing.assert_allclose(x, np.array([[0.11, 0.21]]), atol=1e-7)
    x, fx = ufc.adam_ascent(lin, np.array([[0.995, 0.2]]), 0.0, 1.0, 10)
    assert x[0, 0] == 1.0 and abs(x[0, 1] - 0.21) < 1e-7
