"""Parity AT THE BENCHMARKED SHAPES (bench.py's own workload builder): the fast paths the bench runs must give the
same selections as the plain fp64 kernels / the C oracle on the full-size inputs.

  C3  10^6-point Hartmann-6 mesh, S = 1024, F = 1024, t_max = 128:
        screened step == all-fp64 step (query index / value / point, trusted maximizers)
        product-path stage A (tcgen05 GEMM) == general tcgen05 screen kernel for ALL 1024 samples, k = 3
        full-N stage A, S = 8, bit-exact vs the C oracle (oracle/tes_oracle.c)
  C3 mesh + 5 extra uniform rows (utils_for_continuous.py:36-38 appends top_init_k random candidates)
  C4  one 2^18-candidate slice, d = 10, F = 4096, q = 8: whitened batch TES_sp kernel == direct kernel,
        stage-A screen == fp64 kernel
"""
import os

import numpy as np
import pytest

import bench

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    import torch
    from tes_b200 import optfunc
    wl = bench.make_workload("c3", 0, 1)
    dv = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).cuda() for k in ("cand", "X", "Y", "randn_W", "rand_b", "noise")}
    hyp = wl["hyp"]
    theta, W, b = optfunc.draw_random_init_weights_features(
        wl["d"], wl["S"], wl["F"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"],
        draws=(dv["randn_W"], dv["rand_b"], dv["noise"]))
    return wl, dv, (theta, W, b)


def test_c3_screened_step_equals_fp64_step(c3):
    import torch
    from tes_b200 import acquisition
    wl, dv, _ = c3
    hyp, kw = wl["hyp"], bench.step_kwargs(wl)
    run = lambda **extra: acquisition.tes_step(  # noqa: E731
        dv["cand"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"], None, None, None, seed=5,
        draws=(dv["randn_W"], dv["rand_b"], dv["noise"]), **dict(kw, **extra))
    a = run()
    b = run(screen=False)
    assert a["screen_kept"] is not None and a["screen_kept"] < wl["N"] // 8   # the screen actually pruned
    assert b.get("screen_kept") is None
    assert a["T"] == b["T"] and a["T"] > 100                                   # the 68 KB M-tile regime of the bench
    assert torch.equal(a["argmax_idx"], b["argmax_idx"])
    assert a["query_idx"] == b["query_idx"]
    assert a["query_val"] == b["query_val"]
    np.testing.assert_array_equal(a["query_x"], b["query_x"])


def test_c3_product_stage_a_equals_general_kernel_all_samples(c3):
    import torch
    from tes_b200 import ops
    wl, dv, (theta, W, b) = c3
    product = ops.detect_product_structure(dv["cand"])
    assert product is not None
    i1, v1 = ops.rff_topk(dv["cand"], W, b, theta, wl["hyp"]["sigma"], k=3, product=product)
    i2, v2 = ops.rff_topk(dv["cand"], W, b, theta, wl["hyp"]["sigma"], k=3, product=False, method=ops.RFF_SCREEN_TC5)
    assert torch.equal(i1, i2)
    assert torch.equal(v1, v2)          # both end in the exact spec sequence: bit-identical values


def test_c3_full_mesh_stage_a_bit_exact_vs_c_oracle(c3):
    from oracle import c_oracle as co
    from tes_b200 import ops
    wl, dv, (theta, W, b) = c3
    S8 = 8
    idx, val = ops.rff_topk(dv["cand"], W[:S8], b[:S8], theta[:S8], wl["hyp"]["sigma"], k=3)
    ridx, rval = co.rff_topk(wl["cand"], W[:S8].cpu().numpy(), b[:S8].cpu().numpy(), theta[:S8].cpu().numpy(),
                             wl["hyp"]["sigma"], 3)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    np.testing.assert_array_equal(val.cpu().numpy(), rval)


def test_c3_mesh_plus_extra_rows(c3):
    """The continuous drivers append top_init_k uniform candidates to the grid (utils_for_continuous.py:36-38)."""
    import torch
    from tes_b200 import ops
    wl, dv, (theta, W, b) = c3
    S = 64
    extra = torch.from_numpy(np.random.RandomState(3).rand(5, wl["d"])).cuda()
    cand = torch.cat([dv["cand"], extra], 0).contiguous()
    pre = ops.detect_product_prefix(cand)
    assert pre is not None and pre[3] == dv["cand"].shape[0]          # the mesh is found as the product prefix
    n0 = ops.STATS.get("product_prefix", 0)
    i1, v1 = ops.rff_topk(cand, W[:S], b[:S], theta[:S], wl["hyp"]["sigma"], k=3)
    assert ops.STATS.get("product_prefix", 0) == n0 + 1                # ... and the GEMM path ran on it
    i2, v2 = ops.rff_topk(cand, W[:S], b[:S], theta[:S], wl["hyp"]["sigma"], k=3, product=False, method=ops.RFF_FP64)
    assert torch.equal(i1, i2) and torch.equal(v1, v2)
    # ... and an extra row that wins must be found: plant the maximizer of sample 0 (perturbed) as an extra row
    best = cand[i2[0, 0]].clone()
    cand2 = torch.cat([dv["cand"], best[None], extra], 0).contiguous()
    j1, w1 = ops.rff_topk(cand2, W[:S], b[:S], theta[:S], wl["hyp"]["sigma"], k=3)
    j2, w2 = ops.rff_topk(cand2, W[:S], b[:S], theta[:S], wl["hyp"]["sigma"], k=3, product=False, method=ops.RFF_FP64)
    assert torch.equal(j1, j2) and torch.equal(w1, w2)
    assert int(j1[0, 1]) == dv["cand"].shape[0]      # the planted duplicate ties with the mesh point: lower index first


def test_c3_slab_of_the_8_gpu_mesh_resolves_near_ties_without_the_general_kernel():
    """Rank 5's slab of the 8-GPU weak-scaling mesh (80 points along the slowest axis): a few function samples are flat
    along that axis and have more near-ties than the k + 8 survivors; the resolve pass settles them on the GEMM path.
    Stage A == the general tcgen05 screen kernel for all 1024 samples."""
    import torch
    from tes_b200 import ops, optfunc
    wl = bench.make_workload("c3", 5, 8)
    dv = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).cuda() for k in ("cand", "X", "Y", "randn_W", "rand_b", "noise")}
    hyp = wl["hyp"]
    theta, W, b = optfunc.draw_random_init_weights_features(
        wl["d"], wl["S"], wl["F"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"],
        draws=(dv["randn_W"], dv["rand_b"], dv["noise"]))
    before = dict(ops.STATS)
    i1, v1 = ops.rff_topk(dv["cand"], W, b, theta, hyp["sigma"], k=3, idx_offset=5 * 10 ** 6)
    flagged = ops.STATS["product_flagged"] - before["product_flagged"]
    unresolved = ops.STATS.get("product_unresolved", 0) - before.get("product_unresolved", 0)
    assert flagged >= 1 and unresolved == 0
    i2, v2 = ops.rff_topk(dv["cand"], W, b, theta, hyp["sigma"], k=3, idx_offset=5 * 10 ** 6, product=False,
                          method=ops.RFF_SCREEN_TC5)
    assert torch.equal(i1, i2) and torch.equal(v1, v2)


@pytest.fixture(scope="module")
def c4slice():
    import torch
    from tes_b200 import optfunc
    wl = dict(bench.WORKLOADS["c4"])
    wl.update(N=2 ** 18, S=64)
    bench.WORKLOADS["_c4slice"] = wl
    try:
        wl = bench.make_workload("_c4slice", 0, 1)
    finally:
        del bench.WORKLOADS["_c4slice"]
    dv = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).cuda() for k in ("cand", "X", "Y", "randn_W", "rand_b", "noise")}
    hyp = wl["hyp"]
    theta, W, b = optfunc.draw_random_init_weights_features(
        wl["d"], wl["S"], wl["F"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"],
        draws=(dv["randn_W"], dv["rand_b"], dv["noise"]))
    return wl, dv, (theta, W, b)


def test_c4_slice_stage_a_screen_equals_fp64(c4slice):
    import torch
    from tes_b200 import ops
    wl, dv, (theta, W, b) = c4slice
    S = 16
    i1, v1 = ops.rff_topk(dv["cand"], W[:S], b[:S], theta[:S], wl["hyp"]["sigma"], k=1)
    i2, v2 = ops.rff_topk(dv["cand"], W[:S], b[:S], theta[:S], wl["hyp"]["sigma"], k=1, method=ops.RFF_FP64)
    assert torch.equal(i1, i2) and torch.equal(v1, v2)


def test_c4_slice_batch_tes_sp_whitened_equals_direct(c4slice):
    """q = 8, F = 4096 step on 2^18 candidates (2^15 queries): the whitened-residual kernel (sp2.cuh) against the
    direct kernel (y = mu + L z, triangular solve per log-density) with the same counter-based draws."""
    from tes_b200 import acquisition
    wl, dv, (theta, W, b) = c4slice
    hyp, kw = wl["hyp"], bench.step_kwargs(wl)
    run = lambda: acquisition.tes_step(dv["cand"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"], theta,  # noqa: E731
                                       W, b, seed=9, return_acq=True, **kw)
    a = run()
    os.environ["TES_SP_METHOD"] = "1"
    try:
        b_ = run()
    finally:
        del os.environ["TES_SP_METHOD"]
    assert a["query_idx"] == b_["query_idx"]
    np.testing.assert_allclose(a["acq"].cpu().numpy(), b_["acq"].cpu().numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(a["query_x"], b_["query_x"])
