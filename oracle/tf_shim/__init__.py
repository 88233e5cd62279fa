"""TEST INFRASTRUCTURE ONLY -- an eager, numpy-backed stand-in for the TensorFlow 1.14 /
TensorFlow-Probability 0.7 symbols that the reference's TES hot path uses.

Why it exists: the reference's five TES criteria (criteria/evaluate_mp.py, evaluate_mp_lite.py,
evaluate_sample_mp.py, evaluate_batch_sample_mp.py, evaluate_batch_mp.py) and the TF halves of
utils.py are written against TF-1.14 graph ops.  TensorFlow is not installed here and there is no
network (requirements.txt:5-6 pin tensorflow-gpu 1.14.0 / tensorflow-probability 0.7.0; neither is
vendored under /root/reference).  Every op those files use is element-wise, a reshape/tile/gather,
a reduction, a small dense factorisation or a TFP distribution method -- so with `tensorflow` and
`tensorflow_probability` replaced by the two modules of this package the reference's files run
UNMODIFIED, eagerly, on numpy arrays.  tests/golden/make_golden.py does exactly that (through
oracle/ref_loader.py) and commits the outputs as fixtures; tests/test_oracle_golden.py pins the
numpy restatement oracle/tes_oracle_np.py on them, and the -m gpu tests compare the CUDA path with
the same fixtures.

What is restated here are TF's *op semantics*, not the reference's algorithm:
  * a Python float handed to a math op becomes a float32 tensor (tf.convert_to_tensor), numpy
    scalars/arrays keep their dtype -- this is why utils.evaluate_norm_entropy (utils.py:653-654)
    carries a float32-rounded 0.5*log(2*pi);
  * tf.reduce_logsumexp (TF 1.14 math_ops.py): shift by the max, a non-finite max replaced by 0;
  * tf.where with one argument = coordinates of the true entries, shape (n, rank);
  * tf.while_loop = a Python loop over the body;
  * tfp.distributions.Normal (TFP 0.7 normal.py): sample(n) = z*scale + loc with z ~ N(0,1) of
    shape (n,)+batch_shape; log_prob(x) = -0.5*squared_difference(x/scale, loc/scale)
    - (0.5*log(2*pi) + log(scale));
  * MultivariateNormalFullCovariance = MultivariateNormalTriL(cholesky(cov)): sample = loc + L z,
    log_prob(y) = sum_event N(0,1).log_prob(L^-1 (y-loc)) - sum(log|diag L|);
  * MultivariateNormalDiag(loc, scale_diag) likewise with a diagonal scale.

Random draws: every `sample()` / tf.random.* call takes its standard normals (or uniforms) from
`draws` (DrawSource): in RECORD mode they come from a seeded numpy RandomState and are logged in
call order, in REPLAY mode they are popped from a queue the test filled.  The log is what the
oracle and the CUDA path receive as their explicit `z` arguments.

Only tests/golden/make_golden.py (build container) imports this package.
"""
import sys
import types

from . import draws as _draws
from . import tensorflow as _tf
from . import tensorflow_probability as _tfp

draws = _draws.SOURCE


def install():
    """Make `import tensorflow` / `import tensorflow_probability` resolve to the shim."""
    sys.modules["tensorflow"] = _tf
    sys.modules["tensorflow_probability"] = _tfp
    return _tf, _tfp


def installed():
    return sys.modules.get("tensorflow") is _tf
