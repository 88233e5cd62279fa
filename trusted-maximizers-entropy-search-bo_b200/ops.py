"""Thin typed wrappers over the C ABI (include/tes_b200.h) for CUDA torch tensors."""
import numpy as np
import torch

from . import _lib
from .device import WORKSPACE, dev, ptr, stream_ptr


def _rff_args(cand, W, b, theta):
    cand = dev(cand)
    N, d = cand.shape
    theta = dev(theta)
    theta = theta.reshape(theta.shape[0], -1)
    S, F = theta.shape
    W = dev(W).reshape(-1, F, d)
    b = dev(b).reshape(-1, F)
    if W.shape[0] not in (1, S) or b.shape[0] != W.shape[0]:
        raise ValueError("W must be (S,F,d) or (1,F,d) and b (S,F[,1]) or (1,F[,1]) to match theta (S,F[,1])")
    w_shared = int(W.shape[0] == 1 and S > 1)
    return cand, W, b, theta, N, d, S, F, w_shared


RFF_AUTO, RFF_FP64, RFF_SCREEN, RFF_SCREEN_MMA, RFF_SCREEN_TC5, RFF_SCREEN_TC5T = 0, 1, 2, 3, 4, 5


STATS = {"product_flagged": 0}   # samples the product path handed back to the general kernel (diagnostics)

RFF_PRODUCT = 6          # product-structured candidate sets: one fp64 tensor-pipe GEMM per sample (csrc/rffprod.cu)
PRODUCT_MIN_N = 1 << 15  # below this the general kernels are launch-bound anyway
RESOLVE_FLAGGED = __import__("os").environ.get("TES_PRODUCT_RESOLVE", "1") != "0"   # resolve pass for near-tie samples
PRODUCT_MODE = int(__import__("os").environ.get("TES_PRODUCT_MODE", "8"))   # GEMM engine of the product path (8: CTA pairs, 256 x 256 tiles)


def detect_product_structure(cand):
    """Is the candidate set a product U x V in row-major order -- the coordinates in columns [s_lo, s_hi) depend on
    i // nv only and the other columns on i % nv only (functions.get_meshgrid, or any slab of such a mesh)?
    -> (s_lo, s_hi, nv) or None; among the admissible splits the most balanced one (|nu - nv| smallest).
    Two kinds of pass over the candidate array on the device (tes_product_runlen_f64 once, tes_product_check_f64 once
    per distinct block length), d and 2 d integers back to the host."""
    N, d = cand.shape
    if d < 2 or N < 4 or d > 64:
        return None
    cand = cand.contiguous()
    L = _lib.lib()
    r_d = torch.empty((d,), dtype=torch.int64, device=cand.device)
    _lib.check(L.tes_product_runlen_f64(ptr(cand), N, d, ptr(r_d), stream_ptr()), "tes_product_runlen_f64")
    r = r_d.cpu().tolist()           # r[k]: first row whose coordinate k differs from row 0 (N: constant column)
    checked = {}

    def violations(nv):
        if nv not in checked:
            bad = torch.empty((2 * d,), dtype=torch.int32, device=cand.device)
            _lib.check(L.tes_product_check_f64(ptr(cand), N, d, nv, ptr(bad), stream_ptr()), "tes_product_check_f64")
            checked[nv] = bad.cpu().tolist()
        return checked[nv]

    best = None
    # slow = a prefix [0, du) (meshes of >= 3 dimensions) or a suffix [du, d) (the 2-D mesh, whose first column is fast)
    for s_lo, s_hi in [(0, du) for du in range(1, d)] + [(du, d) for du in range(1, d)]:
        nv = min(r[s_lo:s_hi])       # run length of the first row's slow coordinates
        if nv >= N:                  # the slow columns never change
            continue
        if nv < 2 or N % nv:
            continue
        nu = N // nv
        if best is not None and abs(nu - nv) >= abs(best[3] - best[2]):
            continue
        bad = violations(nv)
        if not any(bad[s_lo:s_hi]) and not any(bad[d + k] for k in range(d) if not (s_lo <= k < s_hi)):
            best = (s_lo, s_hi, nv, nu)
    return None if best is None else best[:3]


def detect_product_prefix(cand, max_tail=None):
    """"Mesh + tail" (utils_for_continuous.py:36-38 appends top_init_k uniform rows to the grid): the longest prefix of
    the candidate set that is a product U x V in row-major order -> (s_lo, s_hi, nv, n_prefix) or None.  The tail
    (N - n_prefix rows) must not exceed `max_tail` (default max(4096, N / 16))."""
    N, d = cand.shape
    if d < 2 or N < 8 or d > 32:
        return None
    max_tail = max(4096, N // 16) if max_tail is None else max_tail
    cand = cand.contiguous()
    L = _lib.lib()
    r_d = torch.empty((d,), dtype=torch.int64, device=cand.device)
    _lib.check(L.tes_product_runlen_f64(ptr(cand), N, d, ptr(r_d), stream_ptr()), "tes_product_runlen_f64")
    r = r_d.cpu().tolist()
    first = {}

    def firsts(nv):
        if nv not in first:
            out = torch.empty((2 * d,), dtype=torch.int64, device=cand.device)
            _lib.check(L.tes_product_prefix_f64(ptr(cand), N, d, nv, ptr(out), stream_ptr()), "tes_product_prefix_f64")
            first[nv] = out.cpu().tolist()
        return first[nv]

    best = None
    for s_lo, s_hi in [(0, du) for du in range(1, d)] + [(du, d) for du in range(1, d)]:
        nv = min(r[s_lo:s_hi])
        if nv >= N or nv < 2:
            continue
        f = firsts(nv)
        bad = min(min(f[s_lo:s_hi]), min(f[d + k] for k in range(d) if not (s_lo <= k < s_hi)))
        n_pre = (min(bad, N) // nv) * nv
        nu = n_pre // nv
        if nu < 2 or N - n_pre > max_tail:
            continue
        key = (n_pre, -abs(nu - nv))
        if best is None or key > best[0]:
            best = (key, (s_lo, s_hi, nv, n_pre))
    return None if best is None else best[1]


def rff_topk_product(cand, W, b, theta, sigma, s_lo, s_hi, nv, k=1, idx_offset=0, tf32=None):
    """tes_rff_topk_product_f64 -> (idx (S,k), val (S,k), flag (S,) int32).  flag[j] = 1: recompute sample j with the
    general kernel (rff_topk does that).  tf32: 2 = 3xFP16 tensor-core screen (default), 1 = 3xTF32, 0 / False = fp64
    (DMMA), 3 = 3xFP16 on tcgen05/TMEM with both operands streamed, 4 / 5 = the same with the P operand (half of it) generated
    on the SM, 6 = CTA pairs (cta_group::2) on 256 x 128 tiles, 8 = CTA pairs on 256 x 256 tiles (default: half the
    operand bytes per MMA); identical results in every mode.  None: PRODUCT_MODE (env TES_PRODUCT_MODE)."""
    if tf32 is None:
        tf32 = PRODUCT_MODE
    cand, W, b, theta, N, d, S, F, ws_ = _rff_args(cand, W, b, theta)
    L = _lib.lib()
    need = L.tes_rff_topk_product_workspace_bytes(N, d, int(nv), S, F, int(k), int(tf32))
    if need < 0:
        raise ValueError("invalid sizes for tes_rff_topk_product_f64: N=%d d=%d nv=%d S=%d F=%d k=%d" % (N, d, nv, S, F, k))
    ws = WORKSPACE.get(need, cand.device)
    out_idx = torch.empty((S, k), dtype=torch.int64, device=cand.device)
    out_val = torch.empty((S, k), dtype=torch.float64, device=cand.device)
    flag = torch.empty((S,), dtype=torch.int32, device=cand.device)
    rc = L.tes_rff_topk_product_f64(ptr(cand), N, d, int(s_lo), int(s_hi), int(nv), ptr(W), ptr(b), ptr(theta), S, F, ws_,
                                    float(sigma), int(k), int(idx_offset), ptr(out_idx), ptr(out_val), ptr(flag),
                                    int(tf32), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_rff_topk_product_f64")
    return out_idx, out_val, flag


def rff_topk_product_resolve(cand, W, b, theta, sigma, s_lo, s_hi, nv, k=1, idx_offset=0):
    """Samples the product path flagged for near-ties: threshold pass of the same tcgen05 GEMM + exact refine of every
    candidate above the threshold (tes_rff_topk_product_resolve_f64) -> (idx, val, flag); flag = 1: still unresolved."""
    cand, W, b, theta, N, d, S, F, ws_ = _rff_args(cand, W, b, theta)
    L = _lib.lib()
    need = L.tes_rff_topk_product_resolve_workspace_bytes(N, d, int(nv), S, F, int(k))
    if need < 0:
        raise ValueError("invalid sizes for tes_rff_topk_product_resolve_f64: N=%d d=%d nv=%d S=%d F=%d k=%d"
                         % (N, d, nv, S, F, k))
    ws = WORKSPACE.get(need, cand.device)
    out_idx = torch.empty((S, k), dtype=torch.int64, device=cand.device)
    out_val = torch.empty((S, k), dtype=torch.float64, device=cand.device)
    flag = torch.empty((S,), dtype=torch.int32, device=cand.device)
    rc = L.tes_rff_topk_product_resolve_f64(ptr(cand), N, d, int(s_lo), int(s_hi), int(nv), ptr(W), ptr(b), ptr(theta), S, F,
                                            ws_, float(sigma), int(k), int(idx_offset), ptr(out_idx), ptr(out_val),
                                            ptr(flag), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_rff_topk_product_resolve_f64")
    return out_idx, out_val, flag


def rff_topk(cand, W, b, theta, sigma, k=1, idx_offset=0, method=RFF_AUTO, product=None):
    """Top-k candidates of every RFF function sample -> (idx (S,k) int64, val (S,k)) CUDA tensors.
    utils_for_continuous.py:172-187.  `method`: RFF_AUTO / RFF_FP64 / RFF_SCREEN... (identical results).
    `product`: (s_lo, s_hi, nv) if the caller knows the candidates are a product set, False to skip the check, None (default)
    to detect it (RFF_AUTO, N >= 32768, k <= 8).  The product path returns the same bits as the general kernels; the
    samples it flags as unprovable (exact ties, NaN) are recomputed by the general kernel."""
    cand, W, b, theta, N, d, S, F, ws_ = _rff_args(cand, W, b, theta)
    if method in (RFF_AUTO, RFF_PRODUCT) and product is not False and k <= 8 and d >= 2 and \
            (N >= PRODUCT_MIN_N or method == RFF_PRODUCT):
        if product is None:
            product = detect_product_structure(cand)
        if product:
            idx, val, flag = rff_topk_product(cand, W, b, theta, sigma, product[0], product[1], product[2], k=k,
                                              idx_offset=idx_offset)
            bad = torch.nonzero(flag).reshape(-1)
            if bad.numel():
                STATS["product_flagged"] += int(bad.numel())
                if RESOLVE_FLAGGED and d <= 16:
                    # near-ties: second pass of the GEMM with a threshold epilogue on the flagged samples only
                    ri, rv, rflag = rff_topk_product_resolve(cand, W if ws_ else W[bad].contiguous(),
                                                             b if ws_ else b[bad].contiguous(), theta[bad].contiguous(),
                                                             sigma, product[0], product[1], product[2], k=k,
                                                             idx_offset=idx_offset)
                    idx[bad], val[bad] = ri, rv
                    bad = bad[torch.nonzero(rflag).reshape(-1)]
                if bad.numel():
                    STATS["product_unresolved"] = STATS.get("product_unresolved", 0) + int(bad.numel())
                    gi, gv = rff_topk(cand, W if ws_ else W[bad], b if ws_ else b[bad], theta[bad], sigma, k=k,
                                      idx_offset=idx_offset, product=False)
                    idx[bad], val[bad] = gi, gv
            return idx, val
        if product is None and method == RFF_AUTO and N >= PRODUCT_MIN_N:
            # mesh + a few extra rows: product path on the prefix, general kernel on the tail, merge the two lists
            pre = detect_product_prefix(cand)
            if pre is not None and pre[3] >= PRODUCT_MIN_N and pre[3] < N:
                s_lo, s_hi, nv, n_pre = pre
                STATS["product_prefix"] = STATS.get("product_prefix", 0) + 1
                i1, v1 = rff_topk(cand[:n_pre], W, b, theta, sigma, k=k, idx_offset=idx_offset, product=(s_lo, s_hi, nv))
                i2, v2 = rff_topk(cand[n_pre:], W, b, theta, sigma, k=min(k, N - n_pre), idx_offset=idx_offset + n_pre,
                                  product=False)
                if i2.shape[1] < k:    # fewer tail rows than k: pad the list with empty entries
                    padi = torch.full((S, k - i2.shape[1]), -1, dtype=torch.int64, device=cand.device)
                    padv = torch.full((S, k - v2.shape[1]), float("-inf"), dtype=torch.float64, device=cand.device)
                    i2, v2 = torch.cat([i2, padi], 1), torch.cat([v2, padv], 1)
                return topk_merge(torch.stack([v1, v2]), torch.stack([i1, i2]))
        if method == RFF_PRODUCT:
            raise ValueError("RFF_PRODUCT: the candidate set is not a product U x V in row-major order")
    if method == RFF_PRODUCT:
        raise ValueError("RFF_PRODUCT needs k <= 8 and d >= 2")
    L = _lib.lib()
    need = L.tes_rff_topk_workspace_bytes(N, d, S, F, k)
    if need < 0:
        raise ValueError("invalid sizes for tes_rff_topk_f64: N=%d d=%d S=%d F=%d k=%d" % (N, d, S, F, k))
    ws = WORKSPACE.get(need, cand.device)
    out_idx = torch.empty((S, k), dtype=torch.int64, device=cand.device)
    out_val = torch.empty((S, k), dtype=torch.float64, device=cand.device)
    rc = L.tes_rff_topk_method_f64(ptr(cand), N, d, ptr(W), ptr(b), ptr(theta), S, F, ws_, float(sigma), int(k),
                                   int(idx_offset), ptr(out_idx), ptr(out_val), ptr(ws), ws.numel(), stream_ptr(),
                                   int(method))
    _lib.check(rc, "tes_rff_topk_f64")
    return out_idx, out_val


def rff_eval(cand, W, b, theta, sigma):
    """All function-sample values (S, N).  optfunc.py:389-402 / utils_for_continuous.py:257-296."""
    cand, W, b, theta, N, d, S, F, ws_ = _rff_args(cand, W, b, theta)
    out = torch.empty((S, N), dtype=torch.float64, device=cand.device)
    rc = _lib.lib().tes_rff_eval_f64(ptr(cand), N, d, ptr(W), ptr(b), ptr(theta), S, F, ws_, float(sigma),
                                     ptr(out), stream_ptr())
    _lib.check(rc, "tes_rff_eval_f64")
    return out


def topk_merge(vals, idx):
    """(P,S,k) partial lists -> (S,k) under (value desc, index asc)."""
    vals = dev(vals)
    idx = dev(idx, dtype=torch.int64)
    P, S, k = vals.shape
    out_val = torch.empty((S, k), dtype=torch.float64, device=vals.device)
    out_idx = torch.empty((S, k), dtype=torch.int64, device=vals.device)
    rc = _lib.lib().tes_topk_merge_f64(ptr(vals), ptr(idx), P, S, k, ptr(out_val), ptr(out_idx), stream_ptr())
    _lib.check(rc, "tes_topk_merge_f64")
    return out_idx, out_val


def fp64_fma_peak(seconds=0.5, device=None, fp32=False, mufu=False, mma=False):
    """Sustained DFMA (fp32=True: packed-FFMA2) rate in TFLOP/s, or (mufu=True) __cosf evaluations in Tcos/s, of
    the current device, CUDA-event timed."""
    import ctypes
    L = _lib.lib()
    dev_ = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    nsm = torch.cuda.get_device_properties(dev_).multi_processor_count
    blocks = nsm * 8
    sink = torch.empty(blocks * 256, dtype=torch.float32 if (fp32 or mufu) else torch.float64, device=dev_)
    flops = ctypes.c_double(0.0)
    iters = 1 << 16
    burn = L.tes_mufu_cos_burn if mufu else (L.tes_fp32_fma_burn if fp32 else
                                             (L.tes_fp64_mma_burn if mma else L.tes_fp64_fma_burn))

    def run(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(burn(blocks, it, ptr(sink), ctypes.byref(flops), stream_ptr()), "burn")
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3, flops.value

    run(iters)
    t, fl = run(iters)
    it2 = max(iters, int(iters * seconds / max(t, 1e-6)))
    t, fl = run(it2)
    return fl / t * 1e-12


def fp64_mma_peak(seconds=0.5, device=None):
    """Sustained DMMA.8x8x4 rate in TFLOP/s (the fp64 tensor pipe the GEMM core runs on)."""
    return fp64_fma_peak(seconds, device, mma=True)


def fp32_fma_peak(seconds=0.5, device=None):
    return fp64_fma_peak(seconds, device, fp32=True)


def mufu_cos_peak(seconds=0.5, device=None):
    return fp64_fma_peak(seconds, device, mufu=True)


# ------------------------------------------------------------------------------------- stage B
def _l(l, d, device):
    t = dev(l, device).reshape(-1)
    if t.numel() != d:
        raise ValueError("length-scale vector must have %d entries, got %d" % (d, t.numel()))
    return t


def se_ard_cross(x, xb, l, sigma):
    """utils.computeKnm (utils.py:703-719) -> (N, m)."""
    x = dev(x)
    xb = dev(xb, x.device)
    N, d = x.shape
    m = xb.shape[0]
    if xb.shape[1] != d:
        raise ValueError("x and xb must share the input dimension")
    l = _l(l, d, x.device)
    out = torch.empty((N, m), dtype=torch.float64, device=x.device)
    rc = _lib.lib().tes_se_ard_cross_f64(ptr(x), N, ptr(xb), m, d, ptr(l), float(sigma), ptr(out), stream_ptr())
    _lib.check(rc, "tes_se_ard_cross_f64")
    return out


def gemm(A, B, transb=False):
    A = dev(A)
    B = dev(B, A.device)
    M, K = A.shape
    N = B.shape[0] if transb else B.shape[1]
    if (B.shape[1] if transb else B.shape[0]) != K:
        raise ValueError("inner dimensions differ")
    C = torch.empty((M, N), dtype=torch.float64, device=A.device)
    rc = _lib.lib().tes_gemm_f64(ptr(A), A.stride(0), ptr(B), B.stride(0), int(transb), ptr(C), N, M, N, K,
                                 stream_ptr())
    _lib.check(rc, "tes_gemm_f64")
    return C


def pnk(test_xs, X, l, sigma, sigma0):
    """pNK = K([test;X]) + sigma0 diag(0_T, I_n)  (criteria/evaluate_mp.py:36-46)."""
    test_xs = dev(test_xs)
    X = dev(X, test_xs.device)
    T, d = test_xs.shape
    n = X.shape[0]
    l = _l(l, d, X.device)
    L = _lib.lib()
    need = L.tes_pnk_workspace_bytes(T, n, d)
    ws = WORKSPACE.get(need, X.device)
    out = torch.empty((T + n, T + n), dtype=torch.float64, device=X.device)
    rc = L.tes_pnk_f64(ptr(test_xs), T, ptr(X), n, d, ptr(l), float(sigma), float(sigma0), ptr(out), ptr(ws),
                       ws.numel(), stream_ptr())
    _lib.check(rc, "tes_pnk_f64")
    return out


def batched_chol2inv(A, return_info=False):
    """utils.chol2inv / multichol2inv (utils.py:657-700) over a (batch, m, m) stack (or one (m, m))."""
    A = dev(A)
    single = A.dim() == 2
    A3 = A.reshape(-1, A.shape[-2], A.shape[-1])
    batch, m, m2 = A3.shape
    if m != m2:
        raise ValueError("matrices must be square")
    L = _lib.lib()
    need = L.tes_batched_chol_workspace_bytes(batch, m, 1)
    ws = WORKSPACE.get(need, A.device)
    out = torch.empty_like(A3)
    info = torch.empty((batch,), dtype=torch.int32, device=A.device)
    rc = L.tes_batched_chol2inv_f64(ptr(A3), batch, m, ptr(out), ptr(info), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_batched_chol2inv_f64")
    out = out[0] if single else out.reshape(A.shape)
    return (out, info) if return_info else out


def chol2inv_checked(A, what):
    """batched_chol2inv that raises where the reference's tf.cholesky would (InvalidArgumentError: 'Cholesky
    decomposition was not successful'): utils.chol2inv / multichol2inv / precomputeInvKs / eval_invKmaxsams."""
    out, info = batched_chol2inv(A, return_info=True)
    nbad = int((info != 0).sum())
    if nbad:
        raise _lib.TesError("%s: %d of %d matrices have no Cholesky factor (not positive definite)"
                            % (what, nbad, info.numel()))
    return out


def gj_inverse(A):
    """General inverse (Gauss-Jordan, partial pivoting) of (batch, n, n) or (n, n): tf.linalg.inv / np.linalg.inv."""
    A = dev(A)
    single = A.dim() == 2
    A3 = A.reshape(-1, A.shape[-2], A.shape[-1]).contiguous()
    batch, n, n2 = A3.shape
    if n != n2:
        raise ValueError("matrices must be square")
    out = torch.empty_like(A3)
    rc = _lib.lib().tes_batched_gj_inverse_f64(ptr(A3), batch, n, ptr(out), stream_ptr())
    _lib.check(rc, "tes_batched_gj_inverse_f64")
    return out[0] if single else out.reshape(A.shape)


def batched_potrf(A, return_info=False):
    A = dev(A)
    single = A.dim() == 2
    A3 = A.reshape(-1, A.shape[-2], A.shape[-1])
    batch, m, _ = A3.shape
    L = _lib.lib()
    need = L.tes_batched_chol_workspace_bytes(batch, m, 0)
    ws = WORKSPACE.get(need, A.device)
    out = torch.empty_like(A3)
    info = torch.empty((batch,), dtype=torch.int32, device=A.device)
    rc = L.tes_batched_potrf_f64(ptr(A3), batch, m, ptr(out), ptr(info), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_batched_potrf_f64")
    out = out[0] if single else out.reshape(A.shape)
    return (out, info) if return_info else out


def posterior_stats(x, test_xs, X, Y, l, sigma, invpNK):
    """-> M (N,T), b (N), Kq (N).  evaluate_mp.py:76-90 / evaluate_sample_mp.py:77-124."""
    x = dev(x)
    dv = x.device
    test_xs, X, Y, invpNK = dev(test_xs, dv), dev(X, dv), dev(Y, dv).reshape(-1), dev(invpNK, dv)
    N, d = x.shape
    T, n = test_xs.shape[0], X.shape[0]
    if invpNK.shape != (T + n, T + n):
        raise ValueError("invpNK must be (T+n, T+n)")
    l = _l(l, d, dv)
    L = _lib.lib()
    need = L.tes_posterior_stats_workspace_bytes(N, T, n, d)
    ws = WORKSPACE.get(need, dv)
    M = torch.empty((N, T), dtype=torch.float64, device=dv)
    b = torch.empty((N,), dtype=torch.float64, device=dv)
    Kq = torch.empty((N,), dtype=torch.float64, device=dv)
    rc = L.tes_posterior_stats_f64(ptr(x), N, d, ptr(test_xs), T, ptr(X), n, ptr(Y), ptr(l), float(sigma),
                                   ptr(invpNK), ptr(M), ptr(b), ptr(Kq), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_posterior_stats_f64")
    return M, b, Kq


def gp_mean_var(x, X, Y, l, sigma, NKinv):
    """utils.compute_mean_var_f(fullcov=False) (utils.py:786-815) -> mean (N), var (N)."""
    x = dev(x)
    dv = x.device
    X, Y, NKinv = dev(X, dv), dev(Y, dv).reshape(-1), dev(NKinv, dv)
    N, d = x.shape
    n = X.shape[0]
    l = _l(l, d, dv)
    L = _lib.lib()
    ws = WORKSPACE.get(L.tes_gp_mean_var_workspace_bytes(N, n), dv)
    mean = torch.empty((N,), dtype=torch.float64, device=dv)
    var = torch.empty((N,), dtype=torch.float64, device=dv)
    rc = L.tes_gp_mean_var_f64(ptr(x), N, d, ptr(X), n, ptr(Y), ptr(l), float(sigma), ptr(NKinv), ptr(mean),
                               ptr(var), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_gp_mean_var_f64")
    return mean, var


# ------------------------------------------------------------------------------------- stage D
EP_DETERMINISTIC, EP_STOCHASTIC, EP_LITE, EP_MOMENTS = 0, 1, 2, 3


def ep_acq(mode, x, test_xs, X, Y, l, sigma, sigma0, invpNK, invKobs, p, post_mean, post_cov, niter=10,
           nysample=10, z=None, seed=0, n_global=None, x_offset=0, want_moments=False):
    """TES_ep acquisition for one hyper-parameter sample.  evaluate_mp.py:116-295, evaluate_mp_lite.py."""
    x = dev(x)
    dv = x.device
    test_xs, X, Y, invpNK = dev(test_xs, dv), dev(X, dv), dev(Y, dv).reshape(-1), dev(invpNK, dv)
    p, post_mean, post_cov = dev(p, dv).reshape(-1), dev(post_mean, dv), dev(post_cov, dv)
    N, d = x.shape
    T, n, Sp = test_xs.shape[0], X.shape[0], p.numel()
    if post_mean.shape != (Sp, T) or post_cov.shape != (Sp, T, T):
        raise ValueError("post_mean must be (Sp,T) and post_cov (Sp,T,T)")
    if invpNK.shape != (T + n, T + n):
        raise ValueError("invpNK must be (T+n, T+n)")
    invKobs_t = dev(invKobs, dv) if invKobs is not None else None
    l = _l(l, d, dv)
    zt = None
    if z is not None:
        zt = dev(z, dv)
        if tuple(zt.shape) != (niter, nysample, Sp, N):
            raise ValueError("z must be (niter, nysample, Sp, N) = %s, got %s"
                             % ((niter, nysample, Sp, N), tuple(zt.shape)))
    L = _lib.lib()
    ws = WORKSPACE.get(L.tes_ep_acq_workspace_bytes(N, T, n, d, Sp), dv)
    acq = torch.empty((N,), dtype=torch.float64, device=dv) if mode != EP_MOMENTS else None
    mom = want_moments or mode == EP_MOMENTS
    omean = torch.empty((N, Sp), dtype=torch.float64, device=dv) if mom else None
    ovar = torch.empty((N, Sp), dtype=torch.float64, device=dv) if mom else None
    rc = L.tes_ep_acq_f64(int(mode), ptr(x), N, d, ptr(test_xs), T, ptr(X), n, ptr(Y), ptr(l), float(sigma),
                          float(sigma0), ptr(invpNK), ptr(invKobs_t), ptr(p), Sp, ptr(post_mean), ptr(post_cov),
                          int(niter), int(nysample), ptr(zt), int(seed), int(N if n_global is None else n_global),
                          int(x_offset), ptr(acq), ptr(omean), ptr(ovar), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_ep_acq_f64")
    if mode == EP_MOMENTS:
        return omean, ovar
    return (acq, omean, ovar) if want_moments else acq


def ep_acq_screen(x, test_xs, X, Y, l, sigma, sigma0, invpNK, invKobs, p, post_cov):
    """Tensor-core screen of the deterministic TES_ep criterion -> (acq~ (N,), eps (N,)) with |acq~ - acq| <= eps proven
    (tes_ep_acq_screen_f64; the quadratic forms as a 3xFP16 tcgen05 GEMM).  p / post_cov: maximizers with p > 0 only."""
    x = dev(x)
    dv = x.device
    test_xs, X, Y, invpNK = dev(test_xs, dv), dev(X, dv), dev(Y, dv).reshape(-1), dev(invpNK, dv)
    p, post_cov, invKobs = dev(p, dv).reshape(-1), dev(post_cov, dv), dev(invKobs, dv)
    N, d = x.shape
    T, n, Sp = test_xs.shape[0], X.shape[0], p.numel()
    if post_cov.shape != (Sp, T, T) or invpNK.shape != (T + n, T + n):
        raise ValueError("post_cov must be (Sp,T,T) and invpNK (T+n, T+n)")
    L = _lib.lib()
    need = L.tes_ep_acq_screen_workspace_bytes(N, T, n, d, Sp)
    if need < 0:
        raise ValueError("tes_ep_acq_screen_f64 needs Sp <= 128 and n >= 1")
    ws = WORKSPACE.get(need, dv)
    acq = torch.empty((N,), dtype=torch.float64, device=dv)
    eps = torch.empty((N,), dtype=torch.float64, device=dv)
    rc = L.tes_ep_acq_screen_f64(ptr(x), N, d, ptr(test_xs), T, ptr(X), n, ptr(Y), ptr(_l(l, d, dv)), float(sigma),
                                 float(sigma0), ptr(invpNK), ptr(invKobs), ptr(p), Sp, ptr(post_cov), ptr(acq), ptr(eps),
                                 ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_ep_acq_screen_f64")
    return acq, eps


def sp_acq(x, test_xs, X, Y, l, sigma, sigma0, invpNK, p, post_samples, post_mask, niter=10, nysample=10, z=None,
           seed=0, n_global=None, x_offset=0, check_info=True):
    """TES_sp acquisition; x (nx,d) or (nx,q,d).  evaluate_sample_mp.py / evaluate_batch_sample_mp.py."""
    x = dev(x)
    dv = x.device
    if x.dim() == 2:
        x = x.unsqueeze(1)
    nx, q, d = x.shape
    test_xs, X, Y, invpNK = dev(test_xs, dv), dev(X, dv), dev(Y, dv).reshape(-1), dev(invpNK, dv)
    p = dev(p, dv).reshape(-1)
    P = dev(post_samples, dv)
    mask = dev(post_mask, dv, dtype=torch.uint8)
    T, n, Sp = test_xs.shape[0], X.shape[0], p.numel()
    if P.dim() != 3 or P.shape[0] != Sp or P.shape[1] != T:
        raise ValueError("post_samples must be (Sp,T,K)")
    K = P.shape[2]
    if tuple(mask.shape) != (Sp, K):
        raise ValueError("post_mask must be (Sp,K)")
    l = _l(l, d, dv)
    zt = None
    if z is not None:
        zt = dev(z, dv)
        want = (niter, nysample, Sp, nx, K, q)
        if zt.numel() != niter * nysample * Sp * nx * K * q:
            raise ValueError("z must have shape %s" % (want,))
    L = _lib.lib()
    need = L.tes_sp_acq_workspace_bytes(nx, q, T, n, d, Sp, K)
    if need < 0:
        raise ValueError("invalid sizes for tes_sp_acq_f64")
    ws = WORKSPACE.get(need, dv)
    acq = torch.empty((nx,), dtype=torch.float64, device=dv)
    rc = L.tes_sp_acq_f64(ptr(x), nx, q, d, ptr(test_xs), T, ptr(X), n, ptr(Y), ptr(l), float(sigma), float(sigma0),
                          ptr(invpNK), ptr(p), Sp, ptr(P), ptr(mask), K, int(niter), int(nysample), ptr(zt),
                          int(seed), int(nx if n_global is None else n_global), int(x_offset), ptr(acq), ptr(ws),
                          ws.numel(), stream_ptr())
    _lib.check(rc, "tes_sp_acq_f64")
    if check_info:
        nbad = int(ws[need - 256:need - 252].view(torch.int32)[0])
        if nbad:
            raise _lib.TesError("tes_sp_acq_f64: the predictive covariance of %d of %d queries has no Cholesky factor "
                                "(tf.cholesky raises in the reference, evaluate_batch_sample_mp.py:306-310)" % (nbad, nx))
    return acq


def batch_ep_acq(p, x):
    """Batch TES_ep (evaluate_batch_mp.py:145-331): -(sum p) log(sum p) per query, NaN for a non-finite query.
    x (nx, q, d)."""
    p = dev(p).reshape(-1)
    x = dev(x, p.device)
    if x.dim() != 3:
        raise ValueError("x must be (nx, batchsize, xdim)")
    nx, q, d = x.shape
    out = torch.empty((nx,), dtype=torch.float64, device=p.device)
    rc = _lib.lib().tes_batch_ep_acq_f64(ptr(p), p.numel(), ptr(x), nx, q, d, ptr(out), stream_ptr())
    _lib.check(rc, "tes_batch_ep_acq_f64")
    return out


def argmax(vals):
    """First-maximum arg-max of a vector -> (index int, value float tensors on device).  bo_discrete.py:806."""
    v = dev(vals).reshape(-1)
    L = _lib.lib()
    ws = WORKSPACE.get(L.tes_argmax_workspace_bytes(v.numel()), v.device)
    oi = torch.empty((1,), dtype=torch.int64, device=v.device)
    ov = torch.empty((1,), dtype=torch.float64, device=v.device)
    rc = L.tes_argmax_f64(ptr(v), v.numel(), ptr(ov), ptr(oi), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_argmax_f64")
    return oi, ov


# ------------------------------------------------------------------------------------- stage C
STREAM_JOINT_Z = 5  # include/tes_spec.h TES_STREAM_JOINT_Z


def normal_fill(seed, iteration, stream_id, count, first_elem=0, device=None):
    """Counter-based N(0,1) draws of include/tes_spec.h -> (count,) CUDA tensor."""
    out = torch.empty((int(count),), dtype=torch.float64,
                      device=device or torch.device("cuda", torch.cuda.current_device()))
    rc = _lib.lib().tes_normal_fill_f64(int(seed), int(iteration), int(stream_id), int(first_elem), int(count),
                                        ptr(out), stream_ptr())
    _lib.check(rc, "tes_normal_fill_f64")
    return out


def dedup(xs, resolution):
    """utils.py:1554-1573 -> (mask (S,) int32 with 1 = duplicate, leader (S,) int64)."""
    xs = dev(xs)
    S, d = xs.shape
    mask = torch.empty((S,), dtype=torch.int32, device=xs.device)
    leader = torch.empty((S,), dtype=torch.int64, device=xs.device)
    rc = _lib.lib().tes_dedup_f64(ptr(xs), S, d, float(resolution), ptr(mask), ptr(leader), stream_ptr())
    _lib.check(rc, "tes_dedup_f64")
    return mask, leader


def sym_eig(A, return_sweeps=False):
    """Batched symmetric eigen-decomposition (one-sided Jacobi): A (batch,n,n) or (n,n) -> evals (..,n),
    evecs (..,n,n) with eigenvectors as columns."""
    A = dev(A)
    single = A.dim() == 2
    A3 = A.reshape(-1, A.shape[-2], A.shape[-1])
    batch, n, n2 = A3.shape
    if n != n2:
        raise ValueError("matrices must be square")
    L = _lib.lib()
    need = L.tes_sym_eig_workspace_bytes(batch, n)
    if need < 0:
        raise ValueError("invalid sizes for tes_sym_eig_f64: batch=%d n=%d" % (batch, n))
    ws = WORKSPACE.get(need, A.device)
    evals = torch.empty((batch, n), dtype=torch.float64, device=A.device)
    evecs = torch.empty((batch, n, n), dtype=torch.float64, device=A.device)
    sweeps = torch.empty((batch,), dtype=torch.int32, device=A.device)
    rc = L.tes_sym_eig_f64(ptr(A3), batch, n, ptr(evals), ptr(evecs), ptr(sweeps), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_sym_eig_f64")
    if single:
        evals, evecs = evals[0], evecs[0]
    return (evals, evecs, sweeps) if return_sweeps else (evals, evecs)


def color_factor(cov):
    """utils.py:1668-1670: A = V diag(sqrt(clip(e, 0))) of a symmetric (n,n) covariance, on the device."""
    cov = dev(cov)
    n = cov.shape[0]
    e, v = sym_eig(cov)
    out = torch.empty((n, n), dtype=torch.float64, device=cov.device)
    rc = _lib.lib().tes_color_factor_f64(ptr(e), ptr(v), n, ptr(out), stream_ptr())
    _lib.check(rc, "tes_color_factor_f64")
    return out


def joint_maxidx(samples, prior_mean=None, remove_correlated=True):
    """utils.py:1675-1705 on samples (ns,T) = z A^T (modified in place: + mean(prior_mean)).
    -> masked (T,) int32, maxidx (ns,) int64, counts (T,) int32."""
    ns, T = samples.shape
    dv = samples.device
    L = _lib.lib()
    ws = WORKSPACE.get(L.tes_joint_maxidx_workspace_bytes(ns, T), dv)
    masked = torch.empty((T,), dtype=torch.int32, device=dv)
    maxidx = torch.empty((ns,), dtype=torch.int64, device=dv)
    counts = torch.empty((T,), dtype=torch.int32, device=dv)
    pm = None if prior_mean is None else dev(prior_mean, dv).reshape(-1)
    rc = L.tes_joint_maxidx_f64(ptr(samples), ns, T, ptr(pm), int(bool(remove_correlated)), ptr(masked), ptr(maxidx),
                                ptr(counts), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_joint_maxidx_f64")
    return masked, maxidx, counts


def bucket_moments(samples, order, starts, cols):
    """empirical_approximation.py:86-107 for every bucket -> mean (nb,nk), cov (nb,nk,nk)."""
    ns, T = samples.shape
    nb, nk = starts.numel() - 1, cols.numel()
    mean = torch.empty((nb, nk), dtype=torch.float64, device=samples.device)
    cov = torch.empty((nb, nk, nk), dtype=torch.float64, device=samples.device)
    rc = _lib.lib().tes_bucket_moments_f64(ptr(samples), ns, T, ptr(order), ptr(starts), nb, ptr(cols), nk, ptr(mean),
                                           ptr(cov), stream_ptr())
    _lib.check(rc, "tes_bucket_moments_f64")
    return mean, cov


def bucket_samples(samples, order, starts, cols, K, prior_mean):
    """utils.py:1716-1727 -> post_samples (nb,nk,K), mask (nb,K) bool."""
    ns, T = samples.shape
    nb, nk = starts.numel() - 1, cols.numel()
    out = torch.empty((nb, nk, K), dtype=torch.float64, device=samples.device)
    mask = torch.empty((nb, K), dtype=torch.uint8, device=samples.device)
    rc = _lib.lib().tes_bucket_samples_f64(ptr(samples), ns, T, ptr(order), ptr(starts), nb, ptr(cols), nk, int(K),
                                           ptr(dev(prior_mean, samples.device).reshape(-1)), ptr(out), ptr(mask),
                                           stream_ptr())
    _lib.check(rc, "tes_bucket_samples_f64")
    return out, mask.bool()


STREAM_RFF_W, STREAM_RFF_B, STREAM_RFF_E = 6, 7, 8  # include/tes_spec.h


def uniform_fill(seed, iteration, stream_id, count, first_elem=0, device=None):
    """Counter-based uniform draws on (0,1) of include/tes_spec.h -> (count,) CUDA tensor."""
    out = torch.empty((int(count),), dtype=torch.float64,
                      device=device or torch.device("cuda", torch.cuda.current_device()))
    rc = _lib.lib().tes_uniform_fill_f64(int(seed), int(iteration), int(stream_id), int(first_elem), int(count),
                                         ptr(out), stream_ptr())
    _lib.check(rc, "tes_uniform_fill_f64")
    return out


def rff_draw(X, Y, l, sigma, sigma0, randn_W, rand_b, randn_noise):
    """optfunc.py:303-385 on the device -> theta (S,F), W (S,F,d), b (S,F) CUDA tensors."""
    X = dev(X)
    dv = X.device
    n, d = X.shape
    Y = dev(Y, dv).reshape(-1)
    randn_W = dev(randn_W, dv)
    S, F = randn_W.shape[0], randn_W.shape[1]
    randn_W = randn_W.reshape(S, F, d)
    rand_b = dev(rand_b, dv).reshape(S, F)
    randn_noise = dev(randn_noise, dv).reshape(S, F)
    if Y.numel() != n:
        raise ValueError("Y must have one entry per observation")
    l = _l(l, d, dv)
    L = _lib.lib()
    need = L.tes_rff_draw_workspace_bytes(S, F, n, d)
    if need < 0:
        raise ValueError("invalid sizes for tes_rff_draw_f64: S=%d F=%d n=%d d=%d" % (S, F, n, d))
    ws = WORKSPACE.get(need, dv)
    theta = torch.empty((S, F), dtype=torch.float64, device=dv)
    W = torch.empty((S, F, d), dtype=torch.float64, device=dv)
    b = torch.empty((S, F), dtype=torch.float64, device=dv)
    rc = L.tes_rff_draw_f64(ptr(X), ptr(Y), n, d, ptr(l), float(sigma), float(sigma0), ptr(randn_W), ptr(rand_b),
                            ptr(randn_noise), S, F, ptr(theta), ptr(W), ptr(b), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_rff_draw_f64")
    return theta, W, b


def ep_moments(mean, cov, resolution=1e-9, max_niter=100, return_niter=False):
    """ep.approximate_EP_np (ep.py:73-204) for every maximizer at once: mean (nhyp,T), cov (nhyp,T,T) ->
    post_mean (nhyp,T,T), post_cov (nhyp,T,T,T) CUDA tensors."""
    cov = dev(cov)
    dv = cov.device
    T = cov.shape[-1]
    cov = cov.reshape(-1, T, T)
    nhyp = cov.shape[0]
    mean = dev(mean, dv).reshape(nhyp, T)
    L = _lib.lib()
    need = L.tes_ep_moments_workspace_bytes(nhyp, T)
    if need < 0:
        raise ValueError("invalid sizes for tes_ep_moments_f64: nhyp=%d T=%d" % (nhyp, T))
    ws = WORKSPACE.get(need, dv)
    pm = torch.empty((nhyp, T, T), dtype=torch.float64, device=dv)
    pc = torch.empty((nhyp, T, T, T), dtype=torch.float64, device=dv)
    nit = torch.empty((nhyp, T), dtype=torch.int32, device=dv)
    rc = L.tes_ep_moments_f64(ptr(mean), ptr(cov), nhyp, T, float(resolution), int(max_niter), ptr(pm), ptr(pc),
                              ptr(nit), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_ep_moments_f64")
    return (pm, pc, nit) if return_niter else (pm, pc)


def rff_adam_refine(x0, W, b, theta, sigma, xmin, xmax, ntrain, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """utils_for_continuous.py:201-232: `ntrain` TF-1 Adam steps on every (sample, start point) -> refined points
    (S,k,d) and their function-sample values (S,k), CUDA tensors."""
    x0 = dev(x0)
    S, k, d = x0.shape
    dv = x0.device
    theta = dev(theta, dv).reshape(S, -1)
    F = theta.shape[1]
    W = dev(W, dv).reshape(-1, F, d)
    b = dev(b, dv).reshape(-1, F)
    if W.shape[0] not in (1, S) or b.shape[0] != W.shape[0]:
        raise ValueError("W must be (S,F,d) or (1,F,d) and b must match")
    lo = dev(np.broadcast_to(np.asarray(xmin, dtype=np.float64), (d,)).copy(), dv)
    hi = dev(np.broadcast_to(np.asarray(xmax, dtype=np.float64), (d,)).copy(), dv)
    out_x = torch.empty_like(x0)
    out_f = torch.empty((S, k), dtype=torch.float64, device=dv)
    rc = _lib.lib().tes_rff_adam_refine_f64(ptr(x0), S, k, d, ptr(W), ptr(b), ptr(theta), F,
                                            int(W.shape[0] == 1 and S > 1), float(sigma), ptr(lo), ptr(hi), int(ntrain),
                                            float(lr), float(beta1), float(beta2), float(epsilon), ptr(out_x),
                                            ptr(out_f), stream_ptr())
    _lib.check(rc, "tes_rff_adam_refine_f64")
    return out_x, out_f
