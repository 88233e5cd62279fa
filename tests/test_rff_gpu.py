"""Stage A parity: CUDA rff_topk / rff_eval (through the C ABI) vs the oracle.

Bar: function-sample values BIT-EXACT vs oracle/tes_oracle.c (same fp64 operation sequence,
include/tes_spec.h); top-k indices bit-exact; values within 1e-12 of the reference's numpy path
(golden fixtures from optfunc.make_function_sample_np)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(seed, N, d, S, F, shared=False, scale=3.0):
    rs = np.random.RandomState(seed)
    cand = rs.rand(N, d)
    l = rs.rand(d) * scale + 0.3
    W = rs.randn(1 if shared else S, F, d) * np.sqrt(l)
    b = 2 * np.pi * rs.rand(1 if shared else S, F, 1)
    theta = rs.randn(S, F, 1)
    return cand, W, b, theta, 0.7 + rs.rand()


@pytest.mark.parametrize("tag", ["nltF", "ngeF", "d6"])
def test_golden_reference_values(golden, tag):
    from tes_b200 import ops
    g = golden("rff")
    sigma = float(g[tag + "_sigma"])
    f = ops.rff_eval(g[tag + "_xs"], g[tag + "_W"], g[tag + "_b"], g[tag + "_theta"], sigma).cpu().numpy()
    np.testing.assert_allclose(f, g[tag + "_fvals"], rtol=0, atol=1e-12)
    idx, val = ops.rff_topk(g[tag + "_xs"], g[tag + "_W"], g[tag + "_b"], g[tag + "_theta"], sigma, k=3)
    np.testing.assert_array_equal(idx[:, 0].cpu().numpy(), g[tag + "_argmax"])


@pytest.mark.parametrize("N,d,S,F,k,shared", [
    (400, 2, 30, 300, 3, False),     # C1 Branin shapes (script_batch_branin.sh)
    (400, 2, 5, 300, 1, False),      # C2 pH shapes
    (5000, 6, 16, 257, 5, False),    # ragged F (not a multiple of the 128-feature stage)
    (1537, 10, 7, 64, 3, False),     # C4 dimensionality, ragged tile
    (3000, 3, 9, 100, 16, False),    # odd d (padded record), max k
    (2048, 6, 12, 128, 3, True),     # shared W (north-star GEMM case), same formula
    (700, 12, 4, 50, 2, False),      # d > 10: generic path
    (1, 2, 3, 8, 3, False),          # N < k: padded with (-inf, -1)
])
def test_topk_bit_exact_vs_c_oracle(N, d, S, F, k, shared):
    from oracle import c_oracle as co
    from tes_b200 import ops
    cand, W, b, theta, sigma = _problem(N + d, N, d, S, F, shared)
    f_gpu = ops.rff_eval(cand, W, b, theta, sigma).cpu().numpy()
    f_cpu = co.rff_eval(cand, W, b, theta, sigma)
    np.testing.assert_array_equal(f_gpu, f_cpu)  # bit-exact values
    idx, val = ops.rff_topk(cand, W, b, theta, sigma, k=k, idx_offset=1000)
    ridx, rval = co.rff_topk(cand, W, b, theta, sigma, k, idx_offset=1000)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    np.testing.assert_array_equal(val.cpu().numpy(), rval)


def test_ties_resolve_to_lowest_index():
    from tes_b200 import ops
    # duplicated candidates -> exactly equal function values; the lower index must win
    cand, W, b, theta, sigma = _problem(3, 600, 2, 6, 40)
    cand2 = np.concatenate([cand, cand[::-1].copy(), cand], axis=0)
    idx, val = ops.rff_topk(cand2, W, b, theta, sigma, k=4)
    idx1, val1 = ops.rff_topk(cand, W, b, theta, sigma, k=2)
    idx, idx1 = idx.cpu().numpy(), idx1.cpu().numpy()
    N = cand.shape[0]
    for j in range(6):
        best = idx1[j, 0]
        # the three copies of the best candidate sit at best, 2N-1-best, 2N+best
        assert list(idx[j, :3]) == sorted([best, 2 * N - 1 - best, 2 * N + best])
    np.testing.assert_array_equal(val[:, 0].cpu().numpy(), val1[:, 0].cpu().numpy())


def test_sharded_partials_merge_to_single_gpu_result():
    """Candidate sharding: per-shard top-k with idx_offset + merge == unsharded."""
    import torch
    from tes_b200 import ops
    cand, W, b, theta, sigma = _problem(11, 4099, 6, 10, 96)
    full_idx, full_val = ops.rff_topk(cand, W, b, theta, sigma, k=3)
    cuts = [0, 1000, 1001, 3000, 4099]
    pv, pi = [], []
    for a, e in zip(cuts[:-1], cuts[1:]):
        i, v = ops.rff_topk(cand[a:e], W, b, theta, sigma, k=3, idx_offset=a)
        pv.append(v)
        pi.append(i)
    mi, mv = ops.topk_merge(torch.stack(pv), torch.stack(pi))
    assert torch.equal(mi, full_idx) and torch.equal(mv, full_val)


def test_large_property_argmax_is_max_of_dense_values():
    """Size-independent property at a larger size: the reported top-1 equals the arg-max of the
    dense evaluation (first maximum), and the top-k values are sorted."""
    import torch
    from tes_b200 import ops
    cand, W, b, theta, sigma = _problem(5, 200_000, 6, 8, 256)
    idx, val = ops.rff_topk(cand, W, b, theta, sigma, k=3)
    f = ops.rff_eval(cand, W, b, theta, sigma)
    assert torch.equal(idx[:, 0], torch.argmax(f, dim=1))
    assert torch.equal(val[:, 0], f.max(dim=1).values)
    assert bool((val[:, 0] >= val[:, 1]).all()) and bool((val[:, 1] >= val[:, 2]).all())


def test_bad_arguments_raise():
    from tes_b200 import TesError, ops
    cand, W, b, theta, sigma = _problem(1, 10, 2, 2, 4)
    with pytest.raises((TesError, ValueError)):
        ops.rff_topk(cand, W, b, theta, sigma, k=0)
    with pytest.raises((TesError, ValueError)):
        ops.rff_topk(cand, W, b, theta, -1.0, k=1)
    with pytest.raises(ValueError):
        ops.rff_topk(cand, W[:1].repeat(3, 0), b, theta, sigma, k=1)
