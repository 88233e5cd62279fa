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


# ---- fp32 screen + fp64 refine: must be BIT-IDENTICAL to the pure fp64 kernel and to the oracle -------------
@pytest.mark.parametrize("N,d,S,F,k,shared,scale", [
    (40_000, 6, 24, 200, 1, False, 3.0),
    (70_001, 6, 16, 257, 3, False, 3.0),     # ragged tile + ragged feature chunk
    (33_000, 2, 40, 300, 5, False, 30.0),    # Branin/pH-like large inverse length-scales (big arguments)
    (50_000, 10, 12, 128, 8, False, 1.0),    # max k for the screen
    (36_000, 3, 9, 100, 2, False, 3.0),      # odd d (padded record)
    (40_000, 6, 12, 160, 3, True, 3.0),      # shared W
    (3_000, 6, 7, 64, 3, False, 3.0),        # small N, screen forced
])
@pytest.mark.parametrize("method", [2, 3, 4, 5])
def test_screened_topk_bit_identical(N, d, S, F, k, shared, scale, method):
    import torch
    from oracle import c_oracle as co
    from tes_b200 import ops
    cand, W, b, theta, sigma = _problem(N + d + k, N, d, S, F, shared, scale)
    i1, v1 = ops.rff_topk(cand, W, b, theta, sigma, k=k, idx_offset=7, method=ops.RFF_FP64)
    i2, v2 = ops.rff_topk(cand, W, b, theta, sigma, k=k, idx_offset=7, method=method)
    assert torch.equal(i1, i2) and torch.equal(v1, v2)
    if N * S * F <= 4e9:
        ridx, rval = co.rff_topk(cand, W, b, theta, sigma, k, idx_offset=7)
        np.testing.assert_array_equal(i2.cpu().numpy(), ridx)
        np.testing.assert_array_equal(v2.cpu().numpy(), rval)


@pytest.mark.parametrize("method", [2, 3, 4, 5])
def test_screened_topk_handles_exact_ties_and_overflow(method):
    """(a) duplicated candidates: exact fp64 ties, the lower index must win; (b) a constant function sample
    (theta = 0) makes EVERY candidate tie, the kept list overflows its 16 slots and the flag-gated fp64
    fallback must produce the plain fp64 answer (indices 0..k-1)."""
    import torch
    from tes_b200 import ops
    cand, W, b, theta, sigma = _problem(9, 20_000, 6, 6, 96)
    cand2 = np.concatenate([cand, cand], axis=0)
    theta = theta.copy()
    theta[2] = 0.0
    i1, v1 = ops.rff_topk(cand2, W, b, theta, sigma, k=4, method=ops.RFF_FP64)
    i2, v2 = ops.rff_topk(cand2, W, b, theta, sigma, k=4, method=method)
    assert torch.equal(i1, i2) and torch.equal(v1, v2)
    i2 = i2.cpu().numpy()
    assert list(i2[2]) == [0, 1, 2, 3]
    for j in (0, 1, 3, 4, 5):
        assert i2[j, 1] == i2[j, 0] + 20_000 and i2[j, 3] == i2[j, 2] + 20_000


@pytest.mark.parametrize("method", [2, 3, 4, 5])
def test_screened_topk_near_ties_many_seeds(method):
    """Stress: dense candidate clouds around one point make thousands of near-ties within the fp32 bound."""
    import torch
    from tes_b200 import ops
    for seed in range(6):
        rs = np.random.RandomState(100 + seed)
        d, S, F, N = 6, 8, 128, 40_000
        center = rs.rand(1, d)
        cand = center + rs.randn(N, d) * 10.0 ** (-2 - seed)   # spreads 1e-2 ... 1e-7
        W = rs.randn(S, F, d) * 2.0
        b = 2 * np.pi * rs.rand(S, F, 1)
        theta = rs.randn(S, F, 1)
        i1, v1 = ops.rff_topk(cand, W, b, theta, 1.3, k=3, method=ops.RFF_FP64)
        i2, v2 = ops.rff_topk(cand, W, b, theta, 1.3, k=3, method=method)
        assert torch.equal(i1, i2) and torch.equal(v1, v2), seed


def test_cosf_error_model_exhaustive():
    """The fp32 screen relies on |__cosf(x) - cos(x)| <= 3.0e-7 + 1.7e-7 |x| for |x| <= 512 (rff.cu SCR_E0/E1).
    Verified here over EVERY float in (0, 512] (both signs) on the device that runs the kernels; the
    measured residual on B200 is 1.50e-7, asserted below 2.0e-7 (so E0 keeps a 1.5x margin)."""
    import torch
    from tes_b200 import _lib
    from tes_b200.device import ptr, stream_ptr
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    for which in (0, 1):   # 0: __cosf (MUFU), 1: the packed FMA-pipe polynomial of the tensor-core screen
        out.zero_()
        rc = _lib.lib().tes_cosf_error_probe(0.0, 512.0, 1.7e-7, which, ptr(out), stream_ptr())
        _lib.check(rc, "tes_cosf_error_probe")
        worst = float(out.cpu()[0])
        assert worst < 2.0e-7, (which, worst)
    # and the documented small-argument behaviour, [-pi, pi]: 2^-21.19 (CUDA programming guide) with slack
    out.zero_()
    _lib.check(_lib.lib().tes_cosf_error_probe(0.0, 3.1415927, 0.0, 0, ptr(out), stream_ptr()), "probe")
    assert float(out.cpu()[0]) < 4.5e-7


@pytest.mark.parametrize("method", [2, 3, 4, 5])
def test_screened_topk_large_arguments_use_reduction_path(method):
    """|W x + b| far beyond the verified __cosf domain: the per-sample mode switches to the explicit 2*pi
    reduction; results stay bit-identical."""
    import torch
    from tes_b200 import ops
    rs = np.random.RandomState(4)
    N, d, S, F = 40_000, 2, 6, 96
    cand = rs.rand(N, d) * 10.0
    W = rs.randn(S, F, d) * 40.0          # |h| up to ~1e3 rad
    W[:3] *= 0.02                          # first three samples stay in the direct domain
    b = 2 * np.pi * rs.rand(S, F, 1)
    theta = rs.randn(S, F, 1)
    i1, v1 = ops.rff_topk(cand, W, b, theta, 0.9, k=2, method=ops.RFF_FP64)
    i2, v2 = ops.rff_topk(cand, W, b, theta, 0.9, k=2, method=method)
    assert torch.equal(i1, i2) and torch.equal(v1, v2)


@pytest.mark.parametrize("d", [1, 2, 3, 6, 7, 10])
def test_mma_split_precision_argument_error(d):
    """The tensor-core screen's bound assumes |h_mma - h| <= 6e-6 A (rff.cu MMA_HARG).  Measured here on random
    operands through the same fragment code; asserted below 3e-6 (2x margin)."""
    import torch
    from tes_b200 import _lib
    from tes_b200.device import dev, ptr, stream_ptr
    worst = 0.0
    for seed in range(4):
        rs = np.random.RandomState(seed)
        n8 = 64
        x = dev(rs.rand(16, d) * 10 ** rs.uniform(-1, 1))
        w = dev(rs.randn(8 * n8, d) * 10 ** rs.uniform(-1, 1.5))
        b = dev(rs.rand(8 * n8) * 6.283)
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().tes_mma_arg_error_probe(ptr(x), ptr(w), ptr(b), d, n8, ptr(out), stream_ptr()), "probe")
        worst = max(worst, float(out.cpu()[0]))
    assert 0.0 < worst < 3.0e-6, worst


# ---- Adam refinement of the trusted maximizers (csrc/rffopt.cu, utils_for_continuous.py:201-232) -----------------
@pytest.mark.parametrize("S,k,d,F,ntrain,shared", [(5, 3, 2, 300, 50, False), (4, 5, 6, 128, 120, False),
                                                   (3, 2, 10, 64, 30, True), (2, 1, 1, 32, 10, False)])
def test_adam_refinement_matches_oracle(S, k, d, F, ntrain, shared):
    from oracle import tes_oracle_np as onp
    from tes_b200 import ops
    rs = np.random.RandomState(S * 100 + d)
    x0 = rs.rand(S, k, d)
    x0[0, 0] = 0.999                                   # starts next to the upper bound: exercises the clip constraint
    W = rs.randn(1 if shared else S, F, d) * 2.0
    b = rs.rand(1 if shared else S, F, 1) * 2 * np.pi
    theta = rs.randn(S, F, 1)
    rx, rf = onp.adam_refine_maximizers(x0, theta, W, b, 1.3, 0.0, 1.0, ntrain)
    gx, gf = ops.rff_adam_refine(x0, W, b, theta, 1.3, 0.0, 1.0, ntrain)
    # fp64 both sides, different summation order over the F features: 1e-9 after `ntrain` steps (tolerance of the
    # refinement: the step size itself is 1e-3)
    np.testing.assert_allclose(gx.cpu().numpy(), rx, rtol=0, atol=1e-9)
    np.testing.assert_allclose(gf.cpu().numpy(), rf, rtol=0, atol=1e-9 * max(1.0, np.abs(rf).max()))
    assert gx.min() >= 0.0 and gx.max() <= 1.0
    # ascent: the refined values are not below the starting values (Adam with lr 1e-3 on a smooth function)
    f0 = onp.adam_refine_maximizers(x0, theta, W, b, 1.3, 0.0, 1.0, 0)[1]
    assert np.all(gf.cpu().numpy() >= f0 - 1e-6)


def test_sample_xmaxs_fmaxs_with_refinement_end_to_end():
    """Top-k over the initialisers + Adam refinement + arg-max over k through the reference-named entry point."""
    from oracle import tes_oracle_np as onp
    from tes_b200 import utils_for_continuous as ufc
    rs = np.random.RandomState(3)
    d, S, F, k = 2, 6, 100, 3
    init = rs.rand(500, d)
    W, b, theta = rs.randn(1, S, F, d) * 3.0, rs.rand(1, S, F, 1) * 6.28, rs.randn(1, S, F, 1)
    inits, max_xs, max_fs, top_idx, xs = ufc.sample_xmaxs_fmaxs(d, 1, S, F, None, np.array([0.8]), None, theta, W, b, init,
                                                               k, xmin=0.0, xmax=1.0, get_xs=True, ntrain=80)
    ridx, _ = onp.rff_topk(init, theta[0], W[0], b[0], 0.8, k)
    np.testing.assert_array_equal(top_idx[0], ridx)
    rx, rf = onp.adam_refine_maximizers(init[ridx], theta[0], W[0], b[0], 0.8, 0.0, 1.0, 80)
    best = np.argmax(rf, axis=1)
    np.testing.assert_allclose(xs[0], rx, atol=1e-9)
    np.testing.assert_allclose(max_xs[0], rx[np.arange(S), best], atol=1e-9)
    np.testing.assert_allclose(max_fs[0], rf[np.arange(S), best], atol=1e-9)


# ---- product-structured candidate sets (csrc/rffprod.cu) -------------------------------------------------------------
def _mesh(axes):
    """functions.get_meshgrid order (np.meshgrid 'xy', flattened C-order)."""
    xds = np.meshgrid(*axes)
    return np.ascontiguousarray(np.concatenate([xd.reshape(-1, 1) for xd in xds], axis=1))


@pytest.mark.parametrize("axes_n,S,F,k,shared", [((7, 5, 6, 4), 9, 96, 1, False), ((10, 10, 10, 10, 10), 5, 257, 3, False),
                                                 ((33, 41), 6, 130, 8, False), ((12, 9, 5), 7, 64, 2, True),
                                                 ((6, 6, 6, 6, 6, 6), 4, 300, 3, False)])
def test_product_path_is_bit_identical_to_the_general_kernels(axes_n, S, F, k, shared):
    """A mesh is detected as U x V and scored by one DMMA GEMM per sample + exact refine: indices AND values are the
    bits of the C oracle (= of the general kernels)."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    rs = np.random.RandomState(sum(axes_n) + S)
    d = len(axes_n)
    cand = _mesh([np.sort(rs.rand(n)) * 2.0 - 0.5 for n in axes_n])
    Wn = rs.randn(1 if shared else S, F, d) * 2.0
    bn = 2 * np.pi * rs.rand(1 if shared else S, F, 1)
    th = rs.randn(S, F, 1)
    det = ops.detect_product_structure(ops.dev(cand))
    assert det is not None and cand.shape[0] % det[2] == 0
    idx, val = ops.rff_topk(cand, Wn, bn, th, 1.3, k=k, idx_offset=1000, method=ops.RFF_PRODUCT)
    ridx, rval = co.rff_topk(cand, Wn, bn, th, 1.3, k, idx_offset=1000)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    np.testing.assert_array_equal(val.cpu().numpy(), rval)
    # every admissible split point gives the same answer (meshgrid 'xy' order: columns 1, 0, 2, 3, ... slowest to fastest)
    for du in range(2, d):
        nv = int(np.prod(axes_n[du:]))
        for tf32 in (8, 6, 5, 4, 3, 2, 1, 0):   # 3xFP16 on tcgen05 (P generated on the SM / streamed) / mma.sync, 3xTF32 mma.sync, fp64 DMMA: same bits after the refine
            i2, v2, fl = ops.rff_topk_product(cand, Wn, bn, th, 1.3, 0, du, nv, k=k, idx_offset=1000, tf32=tf32)
            assert int(fl.sum()) == 0
            np.testing.assert_array_equal(i2.cpu().numpy(), ridx)
            np.testing.assert_array_equal(v2.cpu().numpy(), rval)
    s_lo, s_hi, nv = det
    for tf32 in (8, 6, 5, 4, 3, 2, 1, 0):
        i2, v2, fl = ops.rff_topk_product(cand, Wn, bn, th, 1.3, s_lo, s_hi, nv, k=k, idx_offset=1000, tf32=tf32)
        good = (fl == 0).cpu().numpy()
        assert good.mean() > 0.5
        np.testing.assert_array_equal(i2.cpu().numpy()[good], ridx[good])
        np.testing.assert_array_equal(v2.cpu().numpy()[good], rval[good])


@pytest.mark.parametrize("axes_n,k", [((20, 20, 21, 19), 1), ((30, 17, 23, 9), 8), ((9, 8, 130), 3)])
def test_product_path_two_level_merge_equals_single_cta_merge(axes_n, k, monkeypatch):
    """The partial lists of the GEMM screens are merged by prod_merge_regs_kernel (one level when they fit one CTA's
    registers, else per-block + per-sample); TES_PROD_MERGE=0 keeps the one-CTA-per-sample kernel: same survivors, so
    the same indices, values and flags, and both equal the C oracle."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    rs = np.random.RandomState(sum(axes_n) + k)
    d = len(axes_n)
    cand = _mesh([np.sort(rs.rand(n)) * 1.5 for n in axes_n])
    S, F = 11, 160
    Wn = rs.randn(S, F, d) * 1.7
    bn = 2 * np.pi * rs.rand(S, F, 1)
    th = rs.randn(S, F, 1)
    det = ops.detect_product_structure(ops.dev(cand))
    assert det is not None
    ridx, rval = co.rff_topk(cand, Wn, bn, th, 0.8, k)
    for tf32 in (8, 6, 5, 4, 3, 2):
        out = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("TES_PROD_MERGE", mode)
            i2, v2, fl = ops.rff_topk_product(cand, Wn, bn, th, 0.8, det[0], det[1], det[2], k=k, tf32=tf32)
            out[mode] = (i2.cpu().numpy(), v2.cpu().numpy(), fl.cpu().numpy())
        for a, b_ in zip(out["1"], out["0"]):
            np.testing.assert_array_equal(a, b_)
        good = out["1"][2] == 0
        assert good.mean() > 0.5
        np.testing.assert_array_equal(out["1"][0][good], ridx[good])
        np.testing.assert_array_equal(out["1"][1][good], rval[good])


def test_product_path_flags_ties_and_falls_back():
    """Exact ties (a constant sample; duplicated axis values) cannot be ranked by the GEMM values: the sample is flagged
    and recomputed by the general kernel -> lowest-index tie rule preserved."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    rs = np.random.RandomState(3)
    ax = np.sort(rs.rand(40))
    ax2 = ax.copy()
    ax2[7] = ax2[6]                                  # duplicated axis value -> duplicated candidates
    cand = _mesh([ax2, ax, ax[:20]])
    S, F, d = 5, 80, 3
    Wn, bn, th = rs.randn(S, F, d), 2 * np.pi * rs.rand(S, F, 1), rs.randn(S, F, 1)
    th[2] = 0.0                                      # a constant (zero) function sample: every candidate ties
    det = ops.detect_product_structure(ops.dev(cand))
    assert det is not None
    for tf32 in (8, 6, 5, 4, 3, 2, 1, 0):
        _, _, flag = ops.rff_topk_product(cand, Wn, bn, th, 0.9, det[0], det[1], det[2], k=3, tf32=tf32)
        assert int(flag[2]) == 1
    idx, val = ops.rff_topk(cand, Wn, bn, th, 0.9, k=3, method=ops.RFF_PRODUCT)
    ridx, rval = co.rff_topk(cand, Wn, bn, th, 0.9, 3)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    np.testing.assert_array_equal(val.cpu().numpy(), rval)
    assert idx.cpu().numpy()[2].tolist() == [0, 1, 2]


def test_product_structure_detection():
    import torch
    from tes_b200 import ops
    rs = np.random.RandomState(0)
    assert ops.detect_product_structure(ops.dev(rs.rand(5000, 4))) is None
    m = _mesh([np.linspace(0, 1, 10)] * 6)
    assert ops.detect_product_structure(ops.dev(m)) == (0, 3, 1000)
    assert ops.detect_product_structure(ops.dev(m[300_000:600_000])) is not None      # a slab of the mesh (a rank's shard)
    m2 = m.copy()
    m2[123457, 4] += 1e-9
    assert ops.detect_product_structure(ops.dev(m2)) is None
    # large N: the automatic path takes it and agrees with the general kernel forced explicitly
    th, Wn, bn = rs.randn(3, 64, 6), rs.randn(3, 64, 6), rs.rand(3, 64, 1)
    a = ops.rff_topk(m, Wn, bn, th[:, :, :1], 1.0, k=2)
    b = ops.rff_topk(m, Wn, bn, th[:, :, :1], 1.0, k=2, product=False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_product_detection_kernels_against_numpy():
    """tes_product_runlen_f64 / tes_product_check_f64 (the two passes behind ops.detect_product_structure) against
    their numpy definitions, incl. a constant column, a NaN (never equal: run length 0 when it sits in row 0) and a
    column that is neither slow nor fast."""
    import torch
    from tes_b200 import _lib, ops
    from tes_b200.device import ptr, stream_ptr
    rs = np.random.RandomState(5)
    u, v = np.sort(rs.rand(6, 2)), np.sort(rs.rand(35, 3))
    cand = np.concatenate([np.repeat(u, 35, axis=0), np.tile(v, (6, 1))], axis=1)     # slow columns 0-1, fast 2-4
    cand = np.concatenate([cand, np.full((210, 1), 0.25), rs.rand(210, 1)], axis=1)   # constant, unstructured
    for variant in range(3):
        c = cand.copy()
        if variant == 1:
            c[0, 2] = np.nan
        if variant == 2:
            c[57, 0] = np.nan
        N, d = c.shape
        cd = ops.dev(c)
        r = torch.empty((d,), dtype=torch.int64, device=cd.device)
        assert _lib.lib().tes_product_runlen_f64(ptr(cd), N, d, ptr(r), stream_ptr()) == 0
        diff = ~(c == c[0])
        want_r = [int(np.argmax(diff[:, k])) if diff[:, k].any() else N for k in range(d)]
        assert r.cpu().tolist() == want_r
        for nv in (35, 7, 210, 1):
            bad = torch.empty((2 * d,), dtype=torch.int32, device=cd.device)
            assert _lib.lib().tes_product_check_f64(ptr(cd), N, d, nv, ptr(bad), stream_ptr()) == 0
            i = np.arange(N)
            slow_bad = [int((~(c[:, k] == c[(i // nv) * nv, k])).any()) for k in range(d)]
            fast_bad = [int((~(c[:, k] == c[i % nv, k])).any()) for k in range(d)]
            assert bad.cpu().tolist() == slow_bad + fast_bad
    assert ops.detect_product_structure(ops.dev(cand[:, :5])) == (0, 2, 35)
    assert ops.detect_product_structure(ops.dev(cand)) is None



@pytest.mark.parametrize("eps", [1e-13, 1e-9, 1e-6, 1e-4])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_product_path_near_ties(eps, seed):
    """Axis values eps apart give candidates whose function values differ by ~eps -- inside the screen's error bound for
    the 3xTF32 GEMM (and, for 1e-13, the fp64 one): either both near-tied candidates reach the exact refine or the sample
    is flagged; the answer is the oracle's in every case."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    rs = np.random.RandomState(100 + seed)
    ax = np.sort(rs.rand(24))
    ax2 = ax.copy()
    ax2[1::2] = ax2[0::2] + eps                       # every axis value has a twin eps away
    cand = _mesh([ax2, ax2, ax[:12], ax2[:16]])
    S, F, d = 6, 128, 4
    Wn, bn, th = rs.randn(S, F, d) * 1.5, 2 * np.pi * rs.rand(S, F, 1), rs.randn(S, F, 1)
    ridx, rval = co.rff_topk(cand, Wn, bn, th, 1.1, 3)
    idx, val = ops.rff_topk(cand, Wn, bn, th, 1.1, k=3, method=ops.RFF_PRODUCT)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    np.testing.assert_array_equal(val.cpu().numpy(), rval)
    det = ops.detect_product_structure(ops.dev(cand))
    for tf32 in (0, 1, 2, 3, 4, 5, 6, 8):
        i2, v2, fl = ops.rff_topk_product(cand, Wn, bn, th, 1.1, det[0], det[1], det[2], k=3, tf32=tf32)
        good = (fl == 0).cpu().numpy()
        np.testing.assert_array_equal(i2.cpu().numpy()[good], ridx[good])
        np.testing.assert_array_equal(v2.cpu().numpy()[good], rval[good])


@pytest.mark.parametrize("k,flat_axes,expect_resolved", [(3, 1, True), (1, 1, True), (8, 1, True), (3, 2, False)])
def test_product_path_resolve_pass_for_near_tie_clusters(k, flat_axes, expect_resolved):
    """A mesh that is fine along a direction the function sample is (nearly) flat in: dozens of candidates within the
    screen's error bound of the best, more than the k + 8 survivors can hold, so every sample is flagged.  The resolve
    pass (threshold epilogue of the same tcgen05 GEMM + exact refine of everything above vk - 2B) gives the oracle's
    bits without the general kernel; with a cluster larger than its 2048-entry list it hands the sample on."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    rs = np.random.RandomState(17 + k)
    ax_fine = np.linspace(0.0, 1.0, 48)
    cand = _mesh([np.sort(rs.rand(40)), ax_fine, ax_fine[:44], np.sort(rs.rand(30))])   # columns: 1 = slowest ('xy' order)
    S, F, d = 7, 160, 4
    Wn, bn, th = rs.randn(S, F, d) * 1.5, 2 * np.pi * rs.rand(S, F, 1), rs.randn(S, F, 1)
    Wn[:, :, 1] *= 1e-9                      # flat along the slowest mesh axis (column 1): 48-fold near-ties
    if flat_axes == 2:
        Wn[:, :, 0] *= 1e-9                  # ... and along the next one: 48 x 40 = 1920-fold, x more within the bound
        Wn[:, :, 2] *= 1e-7
    ridx, rval = co.rff_topk(cand, Wn, bn, th, 1.1, k, idx_offset=77)
    det = ops.detect_product_structure(ops.dev(cand))
    assert det is not None
    _, _, fl = ops.rff_topk_product(cand, Wn, bn, th, 1.1, det[0], det[1], det[2], k=k, idx_offset=77)
    assert int(fl.sum()) == S                # every sample has more near-ties than survivors
    ri, rv, rfl = ops.rff_topk_product_resolve(cand, Wn, bn, th, 1.1, det[0], det[1], det[2], k=k, idx_offset=77)
    good = (rfl == 0).cpu().numpy()
    assert bool(good.all()) == expect_resolved
    np.testing.assert_array_equal(ri.cpu().numpy()[good], ridx[good])
    np.testing.assert_array_equal(rv.cpu().numpy()[good], rval[good])
    before = dict(ops.STATS)
    idx, val = ops.rff_topk(cand, Wn, bn, th, 1.1, k=k, idx_offset=77, method=ops.RFF_PRODUCT)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    np.testing.assert_array_equal(val.cpu().numpy(), rval)
    assert ops.STATS["product_flagged"] == before["product_flagged"] + S
    unresolved = ops.STATS.get("product_unresolved", 0) - before.get("product_unresolved", 0)
    assert (unresolved == 0) == expect_resolved


def test_product_path_resolve_pass_shared_weights_and_nan():
    """w_shared layout through the resolve pass; a NaN weight makes the threshold NaN: the sample stays flagged."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    rs = np.random.RandomState(5)
    ax = np.linspace(0.0, 1.0, 40)
    cand = _mesh([np.sort(rs.rand(36)), ax, np.sort(rs.rand(33))])
    S, F, d = 5, 96, 3
    Wn, bn, th = rs.randn(1, F, d), 2 * np.pi * rs.rand(1, F, 1), rs.randn(S, F, 1)
    Wn[:, :, 1] *= 1e-9
    det = ops.detect_product_structure(ops.dev(cand))
    ridx, rval = co.rff_topk(cand, Wn, bn, th, 0.7, 3)
    ri, rv, rfl = ops.rff_topk_product_resolve(cand, Wn, bn, th, 0.7, det[0], det[1], det[2], k=3)
    assert int(rfl.sum()) == 0
    np.testing.assert_array_equal(ri.cpu().numpy(), ridx)
    np.testing.assert_array_equal(rv.cpu().numpy(), rval)
    th2 = th.copy()
    th2[3, 5, 0] = np.nan
    _, _, rfl = ops.rff_topk_product_resolve(cand, Wn, bn, th2, 0.7, det[0], det[1], det[2], k=3)
    assert int(rfl[3]) == 1 and int(rfl.sum()) == 1
