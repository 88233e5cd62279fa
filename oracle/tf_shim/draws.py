"""TEST INFRASTRUCTURE ONLY -- the random-draw source of the TF/TFP shim (see __init__.py)."""
import numpy as np


class DrawSource:
    """Standard normals / uniforms for the shim's sample() calls, in call order.

    record(seed): draws come from numpy.random.RandomState(seed) and are appended to `log`
                  as (kind, array) pairs.
    replay(list): draws are popped from the list; shape and kind must match.
    """

    def __init__(self):
        self.record(0)

    def record(self, seed):
        self.mode = "record"
        self.rs = np.random.RandomState(seed)
        self.log = []
        self.queue = []

    def replay(self, arrays):
        self.mode = "replay"
        self.queue = [(k, np.asarray(a)) for k, a in arrays]
        self.log = []

    def _next(self, kind, shape, dtype):
        shape = tuple(int(s) for s in shape)
        if self.mode == "record":
            a = self.rs.standard_normal(shape) if kind == "normal" else self.rs.random_sample(shape)
            a = a.astype(dtype)
        else:
            if not self.queue:
                raise RuntimeError("tf_shim draws: replay queue exhausted (%s %s)" % (kind, shape))
            k, a = self.queue.pop(0)
            if k != kind or tuple(a.shape) != shape:
                raise RuntimeError("tf_shim draws: expected %s %s, queue holds %s %s"
                                   % (kind, shape, k, a.shape))
            a = a.astype(dtype)
        self.log.append((kind, a))
        return a

    def normal(self, shape, dtype=np.float64):
        return self._next("normal", shape, dtype)

    def uniform(self, shape, dtype=np.float64):
        return self._next("uniform", shape, dtype)

    def normals_logged(self):
        return [a for k, a in self.log if k == "normal"]


SOURCE = DrawSource()
