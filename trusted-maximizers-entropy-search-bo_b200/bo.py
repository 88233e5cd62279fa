"""The sequential continuous-domain BO loop of bo.py:1026-1244 for the TES criteria, re-hosted on the device kernels.
Same skeleton as bo_batch.py (tes_b200.bo_batch.run holds it) with this driver's differences: one query point per
iteration, trusted maximisers de-duplicated at 1e-3 (bo.py:830-836), criteria evaluate_sample_mp.mp ('sftl', TES_sp)
and evaluate_mp.mp ('ftl', TES_ep; `-d 1` deterministic, else the stochastic estimator), all nmax maximisers as start
points of the criterion's Adam refinement (bo.py:1151-1153), NaN-retry (bo.py:1190-1193)."""
from . import bo_batch


def run(f, xs, xmin, xmax, l, sigma, sigma0, criterion="ftl", mode="empirical", deterministic=0, **kw):
    kw.pop("batchsize", None)
    return bo_batch.run(f, xs, xmin, xmax, l, sigma, sigma0, criterion=criterion, mode=mode, batchsize=1, single=True,
                        deterministic=deterministic, **kw)


save = bo_batch.save
