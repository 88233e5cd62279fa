"""Development check: the fused (A generated on the SM) and unfused quadratic-form screens against the fp64 criterion.
Run as  TES_QUAD_FUSED=0|1 python scripts/dev_quad_fused.py  (the switch is read by the library at every call)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests._problems import HART6, make_problem  # noqa: E402
from tests.test_criteria_gpu import _ep_inputs  # noqa: E402

from tes_b200 import ops  # noqa: E402

pr = make_problem(seed=81, d=6, n=40, T=70, nx=200_000, hyp=HART6, mode="empirical", nsample=2800)
p, pm, pc = _ep_inputs(pr)
exact = ops.ep_acq(ops.EP_DETERMINISTIC, pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"],
                   pr["invpNK"][0], pr["invK"][0], p, pm, pc).cpu().numpy()
res = {}
for fused in ("0", "1"):
    os.environ["TES_QUAD_FUSED"] = fused
    acq, eps = ops.ep_acq_screen(pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"],
                                 pr["invpNK"][0], pr["invK"][0], p, pc)
    acq, eps = acq.cpu().numpy(), eps.cpu().numpy()
    fin = np.isfinite(eps) & np.isfinite(exact)
    err = np.abs(acq - exact)[fin]
    kept = int(np.sum(~fin | (acq + eps >= np.max((acq - eps)[fin]))))
    print("fused", fused, "err max %.3e median %.3e  eps median %.3e  kept %d" % (err.max(), np.median(err), np.median(eps[fin]), kept))
    res[fused] = acq
print("fused vs unfused max diff %.3e" % np.nanmax(np.abs(res["0"] - res["1"])))
