"""TEST INFRASTRUCTURE ONLY -- a lazy-graph `tensorflow` 1.14 stand-in on torch (CPU, float64) for the ONE piece of
reference graph code with variables and an optimizer on the TES hot path: utils_for_continuous.sample_xmaxs_fmaxs
(utils_for_continuous.py:138-252, the Adam refinement of the trusted maximizers) -- and, with the same machinery,
optimize_continuous_function (:20-87).

Semantics restated (TF 1.14, not vendored in the reference):
  * ops on concrete values run eagerly; anything downstream of a tf.Variable is a lazy Node evaluated by
    Session.run (memoised per run), so `assign` / `train` / fetches behave like graph ops;
  * tf.get_variable(..., constraint=c): the constraint is applied to the variable after every optimizer update
    (optimizer.py: `var.assign(var.constraint(var))` under a control dependency on the update op);
  * tf.train.AdamOptimizer (adam.py / training_ops ApplyAdam), defaults lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8:
        lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m <- beta1 m + (1 - beta1) g;  v <- beta2 v + (1 - beta2) g^2;
        var <- var - lr_t * m / (sqrt(v) + eps)
    with g = d loss / d var obtained here by torch autograd THROUGH THE REFERENCE'S OWN GRAPH-BUILDING CODE;
  * tf.math.top_k: descending values, ties -> lower index; tf.argmax: first maximum.

Only tests/golden/make_golden_adam.py (build container) imports this module.
"""
import sys
import types

import numpy as np
import torch

__version__ = "1.14.0-torch-graph-shim"

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool = torch.bool  # noqa: A001


# ------------------------------------------------------------------------------------------------ graph nodes
class Node:
    def __init__(self, fn, args=(), kwargs=None, name=None):
        self.fn, self.args, self.kwargs, self.name = fn, args, kwargs or {}, name

    def eval(self, cache):
        k = id(self)
        if k not in cache:
            cache[k] = self.fn(*[_ev(a, cache) for a in self.args], **{n: _ev(v, cache) for n, v in self.kwargs.items()})
        return cache[k]

    # operators build nodes
    def __add__(self, o): return _lift(torch.add)(self, o)
    def __radd__(self, o): return _lift(torch.add)(o, self)
    def __sub__(self, o): return _lift(torch.sub)(self, o)
    def __rsub__(self, o): return _lift(lambda a, b: b - a)(self, o)
    def __mul__(self, o): return _lift(torch.mul)(self, o)
    def __rmul__(self, o): return _lift(torch.mul)(o, self)
    def __truediv__(self, o): return _lift(torch.div)(self, o)
    def __neg__(self): return _lift(torch.neg)(self)
    def __matmul__(self, o): return _lift(torch.matmul)(self, o)
    def __getitem__(self, idx): return _lift(lambda t, i: t[i])(self, idx)


class Variable(Node):
    def __init__(self, shape, dtype, name, constraint):
        super().__init__(None, name=name)
        self.shape_, self.dtype, self.constraint = tuple(int(s) for s in shape), dtype, constraint
        self.value = None

    def eval(self, cache):
        if self.value is None:
            raise RuntimeError("variable %s read before its assign op ran" % self.name)
        return self.value


class Op(Node):
    """A node run for its side effect (assign / train)."""


def _has_node(x):
    if isinstance(x, Node):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_node(v) for v in x)
    return False


def _ev(x, cache):
    if isinstance(x, Node):
        return x.eval(cache)
    if isinstance(x, (list, tuple)):
        return type(x)(_ev(v, cache) for v in x)
    return x


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x))
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=torch.float64)
    if isinstance(x, (int, np.integer)):
        return x
    if isinstance(x, (list, tuple)):
        return type(x)(_t(v) for v in x)
    return x


def _lift(fn):
    def op(*args, **kwargs):
        if _has_node(args) or _has_node(tuple(kwargs.values())):
            return Node(lambda *a, **k: fn(*[_t(v) for v in a], **{n: _t(v) for n, v in k.items()}), args, kwargs)
        return fn(*[_t(v) for v in args], **{n: _t(v) for n, v in kwargs.items()})
    return op


def _ints(shape):
    return tuple(int(s) for s in (shape if isinstance(shape, (list, tuple)) else [shape]))


# ------------------------------------------------------------------------------------------------ ops
def shape(x, name=None, out_type=None):
    if isinstance(x, Variable):
        return x.shape_
    if isinstance(x, Node):
        raise NotImplementedError("tf.shape of a lazy tensor")
    if isinstance(x, torch.Tensor):
        return tuple(int(s) for s in x.shape)
    return tuple(int(s) for s in np.shape(x))


def constant(value, dtype=None, name=None):
    t = _t(np.asarray(value))
    return t.to(dtype) if dtype is not None else t


reshape = _lift(lambda t, shape, name=None: torch.reshape(t, _ints(shape)))
squeeze = _lift(lambda t, axis=None, name=None: torch.squeeze(t) if axis is None else torch.squeeze(t, axis))
expand_dims = _lift(lambda t, axis=None, name=None: torch.unsqueeze(t, axis))
transpose = _lift(lambda t, perm=None, name=None: t.t() if perm is None else t.permute(*perm))
sqrt = _lift(lambda t, name=None: torch.sqrt(t) if isinstance(t, torch.Tensor) else torch.sqrt(torch.tensor(float(t), dtype=torch.float64)))
cos = _lift(lambda t, name=None: torch.cos(t))
exp = _lift(lambda t, name=None: torch.exp(t))
log = _lift(lambda t, name=None: torch.log(t))
square = _lift(lambda t, name=None: t * t)
tile = _lift(lambda t, multiples, name=None: t.repeat(*_ints(multiples)))
concat = _lift(lambda values, axis, name=None: torch.cat(list(values), dim=axis))
stack = _lift(lambda values, axis=0, name=None: torch.stack(list(values), dim=axis))
clip_by_value = _lift(lambda t, lo, hi, name=None: torch.clamp(
    t, min=None if np.isneginf(np.asarray(lo)).all() else _as_bound(lo, t),
    max=None if np.isposinf(np.asarray(hi)).all() else _as_bound(hi, t)))


def _as_bound(b, like):
    b = np.asarray(b, dtype=np.float64)
    return torch.from_numpy(np.broadcast_to(b, tuple(like.shape)).copy()) if b.ndim else float(b)


def _matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


matmul = _lift(_matmul)


def _reduce(fn):
    def red(t, axis=None, keepdims=False, name=None):
        if axis is None:
            return fn(t)
        r = fn(t, dim=axis, keepdim=keepdims)
        return r.values if hasattr(r, "values") else r
    return _lift(red)


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)
reduce_max = _reduce(torch.max)
reduce_min = _reduce(torch.min)


def _top_k(t, k=1, sorted=True, name=None):  # noqa: A002
    order = torch.sort(-t, dim=-1, stable=True).indices[..., : int(k)]   # descending, ties -> lower index
    return torch.gather(t, -1, order), order


_top_k_l = _lift(_top_k)


def _gather(params, indices, validate_indices=None, name=None, axis=0):
    idx = indices if isinstance(indices, torch.Tensor) else torch.as_tensor(indices)
    return torch.index_select(params, axis, idx.reshape(-1).long()).reshape(
        tuple(params.shape[:axis]) + tuple(idx.shape) + tuple(params.shape[axis + 1:]))


gather = _lift(_gather)
argmax = _lift(lambda t, axis=None, name=None, dimension=None, output_type=None: torch.argmax(
    t, dim=axis if axis is not None else (dimension if dimension is not None else 0)))

math = types.SimpleNamespace(top_k=_top_k_l, reduce_logsumexp=None)
nn = types.SimpleNamespace(top_k=_top_k_l)
linalg = types.SimpleNamespace()


class TensorArray:
    def __init__(self, dtype, size=None, **kw):
        self._items = {}

    def write(self, index, value, name=None):
        self._items[int(index)] = value
        return self

    def stack(self, name=None):
        return stack([self._items[i] for i in sorted(self._items)], axis=0)


def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10, back_prop=True,
               swap_memory=False, name=None, maximum_iterations=None, return_same_structure=False):
    state = tuple(loop_vars)
    while cond(*state):
        state = tuple(body(*state))
    return state


_random = np.random.RandomState(0)


def _random_uniform(shape, minval=0, maxval=None, dtype=None, seed=None, name=None):
    u = _random.random_sample(_ints(shape))
    lo = np.asarray(minval, dtype=np.float64)
    hi = np.asarray(1.0 if maxval is None else maxval, dtype=np.float64)
    out = lo + u * (hi - lo)
    UNIFORM_LOG.append(out.copy())
    return torch.from_numpy(out)


UNIFORM_LOG = []
random = types.SimpleNamespace(uniform=_random_uniform)


def seed_uniform(seed):
    global _random
    _random = np.random.RandomState(seed)
    del UNIFORM_LOG[:]


# ------------------------------------------------------------------------------------------------ variables / training
VARIABLES = {}


def get_variable(name=None, shape=None, dtype=None, initializer=None, constraint=None, **kw):
    if name in VARIABLES:
        raise ValueError("Variable %s already exists" % name)   # TF 1.x without reuse
    v = Variable(shape, dtype, name, constraint)
    VARIABLES[name] = v
    return v


def reset_default_graph():
    VARIABLES.clear()


def assign(ref, value, name=None):
    def run(val):
        ref.value = _t(val).detach().clone().to(torch.float64)
        return ref.value
    return Op(run, (value,))


class _Adam:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, use_locking=False, name="Adam"):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def minimize(self, loss, var_list=None, **kw):
        st = {"t": 0, "m": None, "v": None}
        (var,) = var_list

        def run():
            x = var.value.detach().clone().requires_grad_(True)
            var.value = x
            val = loss.eval({})                      # fresh evaluation of the reference's graph at the current value
            (g,) = torch.autograd.grad(val, x)
            st["t"] += 1
            t = st["t"]
            if st["m"] is None:
                st["m"], st["v"] = torch.zeros_like(g), torch.zeros_like(g)
            lr_t = self.lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
            st["m"] = self.b1 * st["m"] + (1.0 - self.b1) * g
            st["v"] = self.b2 * st["v"] + (1.0 - self.b2) * g * g
            new = x.detach() - lr_t * st["m"] / (torch.sqrt(st["v"]) + self.eps)
            if var.constraint is not None:
                new = _ev(var.constraint(new), {})
            var.value = new.detach()
            return None
        return Op(run)


train = types.SimpleNamespace(AdamOptimizer=_Adam)


class Session:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        cache = {}

        def go(f):
            if isinstance(f, (list, tuple)):
                return [go(v) for v in f]
            r = f.eval(cache) if isinstance(f, Node) else f
            return r.detach().numpy() if isinstance(r, torch.Tensor) else r
        with torch.enable_grad():
            return go(fetches)


def install():
    """Make `import tensorflow` resolve to this module (fresh interpreter: before any reference module is loaded)."""
    sys.modules["tensorflow"] = sys.modules[__name__]
    if "tensorflow_probability" not in sys.modules:
        from . import tensorflow_probability as _tfp
        sys.modules["tensorflow_probability"] = _tfp
    return sys.modules[__name__]
