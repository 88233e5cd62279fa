"""TEST INFRASTRUCTURE ONLY -- numpy-backed eager `tensorflow` (1.14 API subset); see __init__.py.

Tensors are plain numpy arrays (operators `@ * + - / []` are numpy's, whose broadcasting and
batch-matmul rules coincide with TF's for every call site of the reference's hot path).
Anything not listed here raises AttributeError, so an unsupported reference code path fails
loudly instead of silently returning a mock.
"""
import types

import numpy as np

from .draws import SOURCE as _draws

__version__ = "1.14.0-numpy-shim"

float16 = np.float16
float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool = np.bool_  # noqa: A001  (tf.bool)


def _t(x):
    """tf.convert_to_tensor: bare Python floats/ints become float32/int32, numpy keeps its dtype."""
    if type(x) is float:
        return np.float32(x)
    if type(x) is int:
        return np.int32(x)
    if isinstance(x, (list, tuple)):
        return np.asarray([_t(v) for v in x]) if len(x) and not np.isscalar(x[0]) else np.asarray(x)
    return np.asarray(x)


def _ints(shape):
    if np.isscalar(shape):
        return (int(shape),)
    return tuple(int(s) for s in shape)


# ---- construction --------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    a = _t(value) if dtype is None else np.asarray(value, dtype=dtype)
    if shape is not None:
        a = np.broadcast_to(a, _ints(shape)).copy()
    return a


def convert_to_tensor(value, dtype=None, name=None):
    return constant(value, dtype=dtype)


def zeros(shape, dtype=float32, name=None):
    return np.zeros(_ints(shape), dtype=dtype)


def ones(shape, dtype=float32, name=None):
    return np.ones(_ints(shape), dtype=dtype)


def zeros_like(x, dtype=None, name=None):
    return np.zeros_like(_t(x), dtype=dtype)


def ones_like(x, dtype=None, name=None):
    return np.ones_like(_t(x), dtype=dtype)


def eye(num_rows, num_columns=None, batch_shape=None, dtype=float32, name=None):
    e = np.eye(int(num_rows), None if num_columns is None else int(num_columns), dtype=dtype)
    if batch_shape is not None:
        e = np.broadcast_to(e, _ints(batch_shape) + e.shape).copy()
    return e


def range(start, limit=None, delta=1, dtype=None, name=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return np.arange(int(start), int(limit), int(delta), dtype=dtype if dtype is not None else np.int32)


def cast(x, dtype, name=None):
    return np.asarray(_t(x)).astype(dtype)


# ---- shape manipulation --------------------------------------------------------------------
def shape(x, name=None, out_type=int32):
    return tuple(int(s) for s in np.shape(x))


def size(x, name=None, out_type=int32):
    return int(np.size(x))


def reshape(tensor, shape, name=None):  # noqa: A002
    return np.reshape(_t(tensor), _ints(shape))


def expand_dims(input, axis=None, name=None, dim=None):  # noqa: A002
    return np.expand_dims(_t(input), axis if axis is not None else dim)


def squeeze(input, axis=None, name=None):  # noqa: A002
    return np.squeeze(_t(input), axis=None if axis is None else tuple(np.atleast_1d(axis)))


def transpose(a, perm=None, name=None, conjugate=False):
    return np.transpose(_t(a), None if perm is None else tuple(int(p) for p in perm))


def tile(input, multiples, name=None):  # noqa: A002
    x = _t(input)
    multiples = _ints(multiples)
    if len(multiples) != x.ndim:
        raise ValueError("tf.tile: multiples %s do not match rank %d" % (multiples, x.ndim))
    return np.tile(x, multiples)


def concat(values, axis, name=None):
    return np.concatenate([_t(v) for v in values], axis=axis)


def stack(values, axis=0, name=None):
    return np.stack([_t(v) for v in values], axis=axis)


def pad(tensor, paddings, mode="CONSTANT", name=None, constant_values=0):
    if mode.upper() != "CONSTANT":
        raise NotImplementedError("tf.pad mode %s" % mode)
    pads = [(int(a), int(b)) for a, b in paddings]
    return np.pad(_t(tensor), pads, mode="constant", constant_values=constant_values)


def gather(params, indices, validate_indices=None, name=None, axis=0):
    return np.take(_t(params), np.asarray(indices), axis=axis)


def gather_nd(params, indices, name=None):
    idx = np.asarray(indices)
    return _t(params)[tuple(np.moveaxis(idx, -1, 0))]


def where(condition, x=None, y=None, name=None):
    if x is None and y is None:
        return np.argwhere(np.asarray(condition)).astype(np.int64)
    return np.where(np.asarray(condition), _t(x), _t(y))


def matrix_diag(diagonal, name=None):
    d = _t(diagonal)
    return d[..., :, None] * np.eye(d.shape[-1], dtype=d.dtype)


# ---- element-wise math ---------------------------------------------------------------------
def _unary(fn):
    def op(x, name=None):
        with np.errstate(all="ignore"):
            return fn(_t(x))
    return op


log = _unary(np.log)
exp = _unary(np.exp)
sqrt = _unary(np.sqrt)
square = _unary(np.square)
cos = _unary(np.cos)
sin = _unary(np.sin)
abs = _unary(np.abs)  # noqa: A001
reciprocal = _unary(lambda v: 1.0 / v)


def erf(x, name=None):
    from scipy.special import erf as _erf
    return _erf(_t(x))


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    return np.clip(_t(t), clip_value_min, clip_value_max)


def _cmp(fn):
    return lambda x, y, name=None: fn(_t(x), _t(y))


less = _cmp(np.less)
less_equal = _cmp(np.less_equal)
greater = _cmp(np.greater)
greater_equal = _cmp(np.greater_equal)
equal = _cmp(np.equal)
not_equal = _cmp(np.not_equal)
maximum = _cmp(np.maximum)
minimum = _cmp(np.minimum)


def squared_difference(x, y, name=None):
    d = _t(x) - _t(y)
    return d * d


# ---- reductions ----------------------------------------------------------------------------
def _reduce(fn):
    def op(input_tensor, axis=None, keepdims=False, name=None, keep_dims=None):
        if keep_dims is not None:
            keepdims = keep_dims
        if axis is not None and not np.isscalar(axis):
            axis = tuple(int(a) for a in axis)
        with np.errstate(all="ignore"):
            return fn(_t(input_tensor), axis=axis, keepdims=keepdims)
    return op


reduce_sum = _reduce(np.sum)
reduce_mean = _reduce(np.mean)
reduce_max = _reduce(np.max)
reduce_min = _reduce(np.min)
reduce_prod = _reduce(np.prod)


def reduce_logsumexp(input_tensor, axis=None, keepdims=False, name=None):
    """TF 1.14 math_ops.reduce_logsumexp: raw_max over axis (keepdims); a non-finite max is
    replaced by 0; log(sum(exp(x - max))) + max."""
    x = _t(input_tensor)
    with np.errstate(all="ignore"):
        raw_max = np.max(x, axis=axis, keepdims=True)
        my_max = np.where(np.isfinite(raw_max), raw_max, np.zeros_like(raw_max))
        result = np.log(np.sum(np.exp(x - my_max), axis=axis, keepdims=True)) + my_max
    if not keepdims:
        result = np.squeeze(result, axis=axis) if axis is not None else np.reshape(result, ())
    return result


def argmax(input, axis=None, name=None, dimension=None, output_type=int64):  # noqa: A002
    if axis is None:
        axis = dimension if dimension is not None else 0
    return np.argmax(_t(input), axis=axis).astype(output_type)


def _top_k(input, k=1, sorted=True, name=None):  # noqa: A002
    """tf.math.top_k on the last axis: values descending, ties -> lower index first."""
    x = _t(input)
    order = np.argsort(-x, axis=-1, kind="stable")[..., : int(k)]
    return np.take_along_axis(x, order, axis=-1), order.astype(np.int32)


# ---- linear algebra ------------------------------------------------------------------------
def matmul(a, b, transpose_a=False, transpose_b=False, adjoint_a=False, adjoint_b=False, name=None):
    a, b = _t(a), _t(b)
    if transpose_a or adjoint_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b or adjoint_b:
        b = np.swapaxes(b, -1, -2)
    return a @ b


def matrix_solve(matrix, rhs, adjoint=False, name=None):
    m = _t(matrix)
    if adjoint:
        m = np.swapaxes(m, -1, -2)
    return np.linalg.solve(m, _t(rhs))


def cholesky(input, name=None):  # noqa: A002
    return np.linalg.cholesky(_t(input))


def svd(tensor, full_matrices=False, compute_uv=True, name=None):
    """tf.svd returns (s, u, v) with v NOT transposed."""
    u, s, vt = np.linalg.svd(_t(tensor), full_matrices=full_matrices)
    return s, u, np.swapaxes(vt, -1, -2)


def _set_diag(input, diagonal, name=None):  # noqa: A002
    out = np.array(_t(input), copy=True)
    n = min(out.shape[-2:])
    idx = np.arange(n)
    out[..., idx, idx] = _t(diagonal)
    return out


linalg = types.SimpleNamespace(
    inv=lambda input, adjoint=False, name=None: np.linalg.inv(_t(input)),  # noqa: A002
    cholesky=cholesky,
    eigh=lambda tensor, name=None: tuple(np.linalg.eigh(_t(tensor))),
    eigvalsh=lambda tensor, name=None: np.linalg.eigvalsh(_t(tensor)),
    diag_part=lambda input, name=None: np.diagonal(_t(input), axis1=-2, axis2=-1).copy(),  # noqa: A002
    set_diag=_set_diag,
    matmul=matmul,
    solve=matrix_solve,
    svd=svd,
    det=lambda input, name=None: np.linalg.det(_t(input)),  # noqa: A002
)

math = types.SimpleNamespace(
    top_k=_top_k, reduce_logsumexp=reduce_logsumexp, log=log, exp=exp, sqrt=sqrt, square=square,
    squared_difference=squared_difference, reduce_sum=reduce_sum, reduce_mean=reduce_mean,
    reduce_max=reduce_max, erf=erf, argmax=argmax)

nn = types.SimpleNamespace(top_k=_top_k)


# ---- control flow --------------------------------------------------------------------------
def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10,
               back_prop=True, swap_memory=False, name=None, maximum_iterations=None,
               return_same_structure=False):
    state = tuple(loop_vars)
    it = 0
    while cond(*state):
        state = tuple(body(*state))
        it += 1
        if maximum_iterations is not None and it >= maximum_iterations:
            break
    return state


class TensorArray:
    def __init__(self, dtype, size=None, dynamic_size=None, **kw):
        self._items = {}

    def write(self, index, value, name=None):
        self._items[int(index)] = _t(value)
        return self

    def stack(self, name=None):
        return np.stack([self._items[i] for i in sorted(self._items)], axis=0)


# ---- random --------------------------------------------------------------------------------
def _random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    return _draws.normal(_ints(shape), dtype) * stddev + mean


def _random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):
    if maxval is None:
        maxval = 1
    return _draws.uniform(_ints(shape), dtype) * (maxval - minval) + minval


random = types.SimpleNamespace(normal=_random_normal, uniform=_random_uniform)
random_normal = _random_normal
random_uniform = _random_uniform


# ---- tf.distributions (TF 1.14 ships a Normal of its own; utils.py:650) --------------------
def _lazy_normal(*a, **k):
    from .tensorflow_probability import distributions as _d
    return _d.Normal(*a, **k)


distributions = types.SimpleNamespace(Normal=_lazy_normal)


def _graph_only(name):
    def fail(*a, **k):
        raise NotImplementedError(
            "tf.%s is graph-mode machinery (variables / optimizers / sessions); the eager numpy "
            "shim covers the criteria and the TF halves of utils.py only" % name)
    return fail


get_variable = _graph_only("get_variable")
assign = _graph_only("assign")
Session = _graph_only("Session")
placeholder = _graph_only("placeholder")
train = types.SimpleNamespace(AdamOptimizer=_graph_only("train.AdamOptimizer"))
random_normal_initializer = _graph_only("random_normal_initializer")
random_uniform_initializer = _graph_only("random_uniform_initializer")
