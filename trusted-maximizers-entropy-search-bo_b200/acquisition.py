"""One TES acquisition step, end to end -- the sequence the reference's drivers run per BO iteration
(bo_discrete.py:504-647 `get_placeholder_values` + :1032-1055; bo_batch.py:550-767 + :994-1061):

  A  evaluate the S RFF posterior function samples on the candidate set, take each sample's arg-max
     (the trusted maximizers), de-duplicate them                      utils_for_continuous.py:172-187, utils.py:1576
  B  GP posterior mean/cov at the maximizers, invpNK                   utils.py:786-815, evaluate_mp.py:12-53
  C  joint samples -> arg-max buckets -> max_probs and per-maximizer sample / empirical / EP statistics
                                                                      utils.py:1733-1833 (host numpy, as in the reference)
  D  the TES criterion on every candidate and its arg-max              criteria/*.py, bo_discrete.py:794-807

Candidates may be sharded across ranks (one process per GPU, torch.distributed): each rank holds a
contiguous slice; only the per-sample (value, index) partials of stage A and one (value, index) pair
per rank of stage D cross NVLink (NCCL all-gather); results are independent of the number of ranks.
"""
import numpy as np
import torch

from . import ops, utils
from .device import dev


class _Timed:
    """Records a CUDA-event pair around a region into timers[name] (no-op when timers is None)."""

    def __init__(self, timers, name):
        self.timers, self.name = timers, name

    def __enter__(self):
        if self.timers is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.timers is not None:
            self.e1.record()
            self.timers.setdefault(self.name, []).append((self.e0, self.e1))
        return False


class Shard:
    """Describes this rank's contiguous slice of the global candidate set."""

    def __init__(self, n_local, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        if self.world > 1:
            if device is None:  # NCCL moves device tensors, gloo (CPU tests of the host logic) host tensors
                device = "cuda" if self.dist.get_backend(group) == "nccl" else "cpu"
            counts = torch.zeros(self.world, dtype=torch.int64, device=device)
            counts[self.rank] = n_local
            self.dist.all_reduce(counts, group=group)
            c = counts.cpu().numpy()
            self.offset = int(c[:self.rank].sum())
            self.n_global = int(c.sum())
        else:
            self.offset, self.n_global = 0, int(n_local)

    def all_gather(self, t):
        if self.world == 1:
            return t.unsqueeze(0)
        flat = t.contiguous().reshape(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.reshape((self.world,) + tuple(t.shape))

    def sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return t


def merge_topk_host(vals, idx, k):
    """Reference semantics of tes_topk_merge_f64 on host tensors: (P,S,kin) -> (S,k) by (value desc, index asc),
    entries with idx < 0 ignored, duplicated (value, index) pairs collapsed.  Used by the CPU (gloo) tests of
    the sharding logic; the product path merges on the device."""
    P, S, kin = vals.shape
    ov = torch.full((S, k), float("-inf"), dtype=torch.float64)
    oi = torch.full((S, k), -1, dtype=torch.int64)
    for j in range(S):
        ent = sorted({(-float(vals[p, j, q]), int(idx[p, j, q])) for p in range(P) for q in range(kin)
                      if int(idx[p, j, q]) >= 0})
        for r, (nv, i) in enumerate(ent[:k]):
            ov[j, r], oi[j, r] = -nv, i
    return oi, ov


def gather_and_merge(shard, idx, val):
    """All-gather the per-rank (S,k) partial lists and merge them (device kernel on CUDA tensors)."""
    if shard.world == 1:
        return idx, val
    gv, gi = shard.all_gather(val), shard.all_gather(idx)
    if val.is_cuda:
        return ops.topk_merge(gv, gi)
    return merge_topk_host(gv, gi, val.shape[1])


def trusted_maximizers(cand, W, b, theta, sigma, shard=None, k=1, timers=None):
    """Stage A.  Returns (idx (S,k) global indices, val (S,k), pts (S,k,d) coordinates)."""
    cand = dev(cand)
    shard = shard or Shard(cand.shape[0])
    with _Timed(timers, "rff_topk"):
        idx, val = ops.rff_topk(cand, W, b, theta, sigma, k=k, idx_offset=shard.offset)
    idx, val = gather_and_merge(shard, idx, val)
    local = idx - shard.offset
    mine = (local >= 0) & (local < cand.shape[0]) & (idx >= 0)
    pts = torch.zeros(idx.shape + (cand.shape[1],), dtype=torch.float64, device=cand.device)
    pts[mine] = cand[local[mine]]
    pts = shard.sum(pts)  # every winner is owned by exactly one rank
    return idx, val, pts


def select_test_points(max_xs, resolution=1e-3, t_max=None):
    """De-duplicate the maximizers at `resolution` (bo.py:830-836, bo_discrete.py:553) keeping first
    occurrences; optionally keep only the `t_max` most frequent ones (ties -> earlier), in original order.
    Runs on the device (tes_dedup_f64: same distance arithmetic as numpy => same decisions); numpy in ->
    numpy out, CUDA tensor in -> CUDA tensor out."""
    xs = dev(max_xs)
    mask, leader = ops.dedup(xs, resolution)
    keep = torch.nonzero(mask == 0).reshape(-1)
    if t_max is not None and keep.numel() > t_max:
        counts = torch.bincount(leader, minlength=xs.shape[0])[keep]
        order = torch.argsort(-counts, stable=True)[:t_max]      # most frequent first, ties -> earlier row
        keep = keep[torch.sort(order).values]
    out = xs[keep]
    return out if torch.is_tensor(max_xs) else out.cpu().numpy()


_COPY_STREAMS = {}


def _stage_upload(host_t):
    """Pinned host tensor -> (device tensor, event): the copy runs on a per-device side stream, ordered after the work
    already enqueued on the current stream (the allocator may hand out memory that work still reads)."""
    device = torch.device("cuda", torch.cuda.current_device())
    side = _COPY_STREAMS.get(device.index)
    if side is None:
        side = _COPY_STREAMS[device.index] = torch.cuda.Stream(device=device)
    out = torch.empty(host_t.shape, dtype=host_t.dtype, device=device)
    ev0 = torch.cuda.Event()
    ev0.record()
    side.wait_event(ev0)
    with torch.cuda.stream(side):
        out.copy_(host_t, non_blocking=True)
        ev1 = torch.cuda.Event()
        ev1.record(side)
    return out, ev1


def tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, criterion="ftl", mode="empirical", deterministic=True,
             ntestsample=1000, n_min_sample=2, nysample=10, niteration=10, z_joint=None, z=None, seed=0,
             duplicate_resolution=1e-3, t_max=None, shard=None, return_acq=False, timers=None, transform=None, q=1,
             screen=None, draws=None):
    """One acquisition step for one hyper-parameter sample.  Returns a dict with the query index /
    point / value and the intermediate sizes.  `criterion`: 'ftl' (TES_ep) or 'sftl' (TES_sp).
    `q` > 1 (criterion 'sftl' only): batch TES_sp (evaluate_batch_sample_mp.py) on the queries formed by every `q`
    consecutive candidates of this rank's slice (nx = N // q batches; the slice length must be a multiple of q);
    query_idx is then the global batch index and query_x the (q, d) batch.
    `screen` (deterministic TES_ep only; default: on when the acquisition vector is not requested, N >= 32768 and at most
    128 maximizers): score all candidates with the tensor-core screen (ops.ep_acq_screen: value + proven error bound),
    evaluate only the candidates that can still be the arg-max with the fp64 criterion -- same query index and value.
    out["screen_kept"] reports how many were re-evaluated (None: the fp64 criterion ran on everything).
    `draws` = (randn_W (S,F,d), rand_b (S,F,1), randn_noise (S,F,1)): the step starts at the RFF posterior weight
    draw (stage A0, optfunc.py:303-385, on the device) and `theta`, `W`, `b` are ignored (pass None)."""
    from . import optfunc
    from .criteria import evaluate_batch_sample_mp, evaluate_mp, evaluate_sample_mp
    if not hasattr(cand, "shape"):
        cand = np.asarray(cand, dtype=np.float64)
    N, d = int(cand.shape[0]), int(cand.shape[1])
    shard = shard or Shard(N)
    staged = None
    if draws is not None:
        # Stage A0.  With R ranks every rank draws the weights of S / R function samples (the draw is independent per
        # sample: its own W_j, b_j, noise, Gram matrix and eigen-decomposition -- same bits as the replicated draw) and
        # theta / W / b are all-gathered: stage A scores every sample against the rank's candidate slice.
        S_, F_ = int(draws[0].shape[0]), int(draws[0].shape[1])
        split = shard.world > 1 and S_ % shard.world == 0
        lo = shard.rank * (S_ // shard.world) if split else 0
        hi = lo + S_ // shard.world if split else S_
        draws_d = tuple(dev(t[lo:hi]) for t in draws)      # host draws: only this rank's slice crosses PCIe, first
        # a pinned host candidate array is uploaded on a side stream UNDER the weight draw (the only stage that does
        # not read the candidates); the compute stream waits for it before stage A
        if torch.is_tensor(cand) and not cand.is_cuda and cand.dtype == torch.float64 and cand.is_contiguous() \
                and cand.is_pinned():
            staged = _stage_upload(cand)
            cand = staged[0]
        else:
            cand = dev(cand)
        with _Timed(timers, "draw"):
            theta, W, b = optfunc.draw_random_init_weights_features(d, hi - lo, F_, X, Y, l, sigma, sigma0, draws=draws_d)
            if split:
                theta = shard.all_gather(theta).reshape(S_, F_, 1)
                W = shard.all_gather(W).reshape(S_, F_, d)
                b = shard.all_gather(b).reshape(S_, F_, 1)
        if staged is not None:
            torch.cuda.current_stream().wait_event(staged[1])
    else:
        cand = dev(cand)
    q = int(q)
    if q > 1 and (criterion != "sftl" or N % q or shard.offset % q):
        raise ValueError("q > 1 needs criterion='sftl' and candidate slices that are multiples of q")
    Xd, Yd = dev(X), dev(Y).reshape(-1, 1)
    n = Xd.shape[0]
    ls, sigmas, sigma0s = np.asarray(l, dtype=np.float64).reshape(1, d), np.array([sigma]), np.array([sigma0])
    # ---- A
    idx, val, pts = trusted_maximizers(cand, W, b, theta, sigma, shard, k=1, timers=timers)
    test_xs = select_test_points(pts[:, 0], duplicate_resolution, t_max)   # CUDA tensor
    out = {"n_unique_maximizers": int(test_xs.shape[0]), "argmax_idx": idx[:, 0], "argmax_val": val[:, 0]}
    if test_xs.shape[0] == 1:  # bo_discrete.py:557-559: a single maximizer is queried directly
        out.update(query_x=test_xs[0].cpu().numpy(), query_idx=int(idx[0, 0]), query_val=float("nan"), T=1, Sp=1)
        return out
    # ---- B
    tb = _Timed(timers, "stage_b")
    tb.__enter__()
    invK = utils.precomputeInvKs(d, 1, ls, sigmas, sigma0s, Xd)
    mean_t, cov_t = utils.compute_mean_var_f(dev(test_xs), Xd, Yd, ls[0], sigma, sigma0, invK[0], fullcov=True)
    tb.__exit__()
    # ---- C (device, replicated on every rank with identical draws; only the bucket counts visit the host)
    tc = _Timed(timers, "stage_c")
    tc.__enter__()
    T0 = test_xs.shape[0]
    st_mode = "sample" if criterion == "sftl" else mode
    test_k, max_probs, mean_k, cov_k, v0, v1 = utils.get_testidxs_stats(
        1, test_xs, mean_t.reshape(1, T0), cov_t.reshape(1, T0, T0), mode=st_mode, nsample=ntestsample,
        n_min_sample=n_min_sample, z=z_joint, transforms=None if transform is None else [transform],
        seed=None if z_joint is not None else seed)
    tc.__exit__()
    out.update(T=int(test_k.shape[0]), Sp=int((max_probs[0] > 0).sum()))
    if test_k.shape[0] == 1:
        out.update(query_x=test_k[0].cpu().numpy(), query_idx=-1, query_val=float("nan"))
        return out
    # ---- D (timer "criterion" = invpNK + criterion on every candidate)
    tm = _Timed(timers, "criterion")
    tm.__enter__()
    _, invpNK = evaluate_mp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, Xd, dev(test_k))
    sub = None   # candidate subset the exact criterion was evaluated on (screened path)
    if criterion == "ftl":
        pk = dev(max_probs).reshape(-1)
        nzp = torch.nonzero(pk > 0.0).reshape(-1)
        use_screen = (deterministic and not return_acq and N >= 32768 and nzp.numel() <= 128) if screen is None \
            else bool(screen and deterministic and nzp.numel() <= 128)
        if use_screen:
            with _Timed(timers, "criterion_screen"):
                acq_t, eps = ops.ep_acq_screen(cand, dev(test_k), Xd, Yd, ls[0], sigma, sigma0, dev(invpNK)[0],
                                               dev(invK)[0], pk[nzp], dev(v1)[0][nzp])
            fin = torch.isfinite(acq_t) & torch.isfinite(eps)
            lb = (acq_t - eps)[fin].max() if bool(fin.any()) else torch.tensor(float("-inf"), device=cand.device)
            keep = (~fin) | (acq_t + eps >= lb)
            sub = torch.nonzero(keep).reshape(-1)             # ascending: the first-maximum rule carries over
            if sub.numel() > max(4096, N // 8):
                sub = None                                    # bound too loose to pay off: fp64 on everything
            out["screen_kept"] = None if sub is None else int(sub.numel())
        xc = cand if sub is None else cand[sub].contiguous()
        acq = evaluate_mp.mp(xc, ls, sigmas, sigma0s, Xd, Yd, d, xc.shape[0], n, 1, nysample, dev(test_k),
                             dev(max_probs), dev(v0), dev(v1), invK, invpNK, stochastic=not deterministic,
                             niteration=niteration, z=z, seed=seed, n_global=shard.n_global, x_offset=shard.offset)
    elif criterion == "sftl" and q > 1:
        out["K"] = int(v0.shape[-1])
        acq = evaluate_batch_sample_mp.mp(cand.reshape(N // q, q, d), ls, sigmas, sigma0s, Xd, Yd, d, n, 1, nysample,
                                          dev(test_k), dev(max_probs), dev(v0), v1, invpNK, niteration=niteration,
                                          z=z, seed=seed, n_global=shard.n_global // q, x_offset=shard.offset // q)
    elif criterion == "sftl":
        out["K"] = int(v0.shape[-1])
        acq = evaluate_sample_mp.mp(cand, ls, sigmas, sigma0s, Xd, Yd, d, N, n, 1, nysample, dev(test_k),
                                    dev(max_probs), dev(v0), v1, invpNK, niteration=niteration, z=z, seed=seed,
                                    n_global=shard.n_global, x_offset=shard.offset)
    else:
        raise ValueError("criterion must be 'ftl' or 'sftl'")
    tm.__exit__()
    # arg-max over all candidates, first maximum wins (bo_discrete.py:806, :1053-1055)
    li, lv = ops.argmax(acq)
    if sub is not None:
        li = sub[li.reshape(-1)[0]].reshape(li.shape)
    pair_v = shard.all_gather(lv.reshape(1))
    pair_i = shard.all_gather((li + shard.offset // q).reshape(1))
    qi, qv = ops.topk_merge(pair_v.reshape(-1, 1, 1), pair_i.reshape(-1, 1, 1))
    qidx = int(qi[0, 0])
    loc = qidx - shard.offset // q
    qx = torch.zeros(d if q == 1 else (q, d), dtype=torch.float64, device=cand.device)
    if 0 <= loc < N // q:
        qx = cand[loc].clone() if q == 1 else cand[loc * q:(loc + 1) * q].clone()
    qx = shard.sum(qx)
    out.update(query_idx=qidx, query_val=float(qv[0, 0]), query_x=qx.cpu().numpy())
    if return_acq:
        out["acq"] = acq
    return out
