"""TEST INFRASTRUCTURE ONLY -- numpy-backed `tensorflow_probability` (0.7 API subset); see
__init__.py.  The formulas follow TFP 0.7's python sources (normal.py, mvn_linear_operator.py,
transformed_distribution.py, affine_linear_operator.py), which are not vendored in the reference.
"""
import math as _math
import types

import numpy as np

from .draws import SOURCE as _draws
from .tensorflow import _t

__version__ = "0.7.0-numpy-shim"

_HALF_LOG_2PI = 0.5 * _math.log(2.0 * _math.pi)  # normal.py: 0.5 * np.log(2. * np.pi)


class Normal:
    """tfp.distributions.Normal(loc, scale) -- TFP 0.7 normal.py."""

    def __init__(self, loc, scale, validate_args=False, allow_nan_stats=True, name="Normal"):
        self.loc = _t(loc)
        self.scale = _t(scale)
        self.dtype = np.result_type(self.loc, self.scale)

    @property
    def batch_shape(self):
        return np.broadcast_shapes(self.loc.shape, self.scale.shape)

    def sample(self, sample_shape=(), seed=None, name="sample"):
        # _sample_n: sampled = random_normal([n] + batch_shape); return sampled * scale + loc
        n = (int(sample_shape),) if np.isscalar(sample_shape) else tuple(int(s) for s in sample_shape)
        z = _draws.normal(n + tuple(self.batch_shape), self.dtype)
        return z * self.scale + self.loc

    def log_prob(self, value, name="log_prob"):
        x = _t(value)
        with np.errstate(all="ignore"):
            a = x / self.scale - self.loc / self.scale  # squared_difference(x/scale, loc/scale)
            return -0.5 * (a * a) - (_HALF_LOG_2PI + np.log(self.scale))

    def prob(self, value, name="prob"):
        return np.exp(self.log_prob(value))

    def cdf(self, value, name="cdf"):
        from scipy.special import ndtr
        return ndtr((_t(value) - self.loc) / self.scale)

    def log_cdf(self, value, name="log_cdf"):
        from scipy.special import log_ndtr
        return log_ndtr((_t(value) - self.loc) / self.scale)


class _MVNLinearOperator:
    """TransformedDistribution(Normal(0,1), AffineLinearOperator(shift=loc, scale)) with event
    rank 1 (mvn_linear_operator.py): sample = scale.matvec(z) + loc;
    log_prob(y) = sum_event N(0,1).log_prob(scale.solvevec(y - loc)) - log|det scale|."""

    def __init__(self, loc, batch_shape, event):
        self.loc = _t(loc)
        self._batch = tuple(batch_shape)
        self._event = int(event)
        self.dtype = self.loc.dtype

    @property
    def batch_shape(self):
        return self._batch

    def _matvec(self, z):
        raise NotImplementedError

    def _solvevec(self, r):
        raise NotImplementedError

    def _log_abs_det(self):
        raise NotImplementedError

    def sample(self, sample_shape=(), seed=None, name="sample"):
        n = (int(sample_shape),) if np.isscalar(sample_shape) else tuple(int(s) for s in sample_shape)
        z = _draws.normal(n + self._batch + (self._event,), self.dtype)
        return self._matvec(z) + self.loc

    def log_prob(self, value, name="log_prob"):
        y = _t(value)
        with np.errstate(all="ignore"):
            x = self._solvevec(y - self.loc)
            lp = np.sum(-0.5 * (x * x) - _HALF_LOG_2PI, axis=-1)
            return lp - self._log_abs_det()


class MultivariateNormalTriL(_MVNLinearOperator):
    def __init__(self, loc=None, scale_tril=None, validate_args=False, allow_nan_stats=True,
                 name="MultivariateNormalTriL"):
        self.scale_tril = _t(scale_tril)
        q = self.scale_tril.shape[-1]
        loc = np.zeros((q,), self.scale_tril.dtype) if loc is None else _t(loc)
        batch = np.broadcast_shapes(loc.shape[:-1], self.scale_tril.shape[:-2])
        super().__init__(loc, batch, q)

    def _matvec(self, z):
        return np.squeeze(self.scale_tril @ z[..., None], axis=-1)

    def _solvevec(self, r):
        # LinearOperatorLowerTriangular.solvevec = matrix_triangular_solve (forward substitution)
        L = self.scale_tril
        q = L.shape[-1]
        x = [None] * q
        for i in range(q):
            acc = r[..., i]
            for j in range(i):
                acc = acc - L[..., i, j] * x[j]
            x[i] = acc / L[..., i, i]
        return np.stack(x, axis=-1)

    def _log_abs_det(self):
        return np.sum(np.log(np.abs(np.diagonal(self.scale_tril, axis1=-2, axis2=-1))), axis=-1)


class MultivariateNormalFullCovariance(MultivariateNormalTriL):
    """mvn_full_covariance.py: scale_tril = tf.linalg.cholesky(covariance_matrix)."""

    def __init__(self, loc=None, covariance_matrix=None, validate_args=False, allow_nan_stats=True,
                 name="MultivariateNormalFullCovariance"):
        with np.errstate(all="ignore"):
            tril = np.linalg.cholesky(_t(covariance_matrix))
        super().__init__(loc=loc, scale_tril=tril)


class MultivariateNormalDiag(_MVNLinearOperator):
    def __init__(self, loc=None, scale_diag=None, scale_identity_multiplier=None,
                 validate_args=False, allow_nan_stats=True, name="MultivariateNormalDiag"):
        if scale_identity_multiplier is not None:
            raise NotImplementedError("scale_identity_multiplier")
        self.scale_diag = _t(scale_diag)
        loc = np.zeros_like(self.scale_diag) if loc is None else _t(loc)
        batch = np.broadcast_shapes(loc.shape[:-1], self.scale_diag.shape[:-1])
        super().__init__(loc, batch, self.scale_diag.shape[-1])

    def _matvec(self, z):
        return self.scale_diag * z

    def _solvevec(self, r):
        return r / self.scale_diag

    def _log_abs_det(self):
        return np.sum(np.log(np.abs(self.scale_diag)), axis=-1)


def _unsupported(name):
    def fail(*a, **k):
        raise NotImplementedError("tfp.distributions.%s is not on the TES hot path" % name)
    return fail


distributions = types.SimpleNamespace(
    Normal=Normal,
    MultivariateNormalTriL=MultivariateNormalTriL,
    MultivariateNormalFullCovariance=MultivariateNormalFullCovariance,
    MultivariateNormalDiag=MultivariateNormalDiag,
    TruncatedNormal=_unsupported("TruncatedNormal"),
)
