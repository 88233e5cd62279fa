"""TEST INFRASTRUCTURE ONLY -- loader for the *reference's own* numpy functions.

The reference (ZhaoxuanWu/Trusted-Maximizers-Entropy-Search-BO, mounted read-only at
/root/reference in the build container) is pure Python on TensorFlow 1.14.  Its numpy
halves (optfunc.draw_random_init_weights_features_np, optfunc.make_function_sample_np,
utils.computeK*_np, utils.get_testidxs_stats, utils.sample_multivariate_normal_maxidx_np,
empirical_approximation.get_empirical_stat_from_samples, ep.approximate_EP_np, functions.*)
import and run unmodified once `matplotlib` and `keras` are pre-seeded in sys.modules as inert
stubs and `np.infty` (removed in numpy 2) is shimmed.  `tensorflow` and
`tensorflow_probability` resolve to oracle/tf_shim -- an eager numpy stand-in for the TF-1.14 /
TFP-0.7 symbols the hot path uses -- so the TF criteria (`criteria.evaluate_mp`, ...) and the TF
halves of utils.py (`computeKnm`, `chol2inv`, `compute_mean_var_f`, ...) run unmodified too:
`load("criteria.evaluate_mp").mp(...)`.

This module is used ONLY by tests/golden/make_golden.py (run in the build container to
produce the committed fixtures) -- /root/reference does not exist on the GPU box, and
nothing in the product package, bench.py or the -m gpu tests may import it.
"""
import contextlib
import importlib
import io
import os
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("TES_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils.py"))


GRAPH_MODE = False


def use_graph_shim():
    """Call BEFORE the first load(): `tensorflow` then resolves to oracle/tf_shim/graph.py (lazy graph on torch with
    variables, assign, AdamOptimizer and Session) -- what utils_for_continuous.sample_xmaxs_fmaxs needs."""
    global GRAPH_MODE
    if _cache:
        raise RuntimeError("use_graph_shim() must be called before any reference module is loaded")
    GRAPH_MODE = True


def _install_stubs():
    if not hasattr(np, "infty"):
        np.infty = np.inf  # utils.py:1669, empirical_approximation.py:66
    # tensorflow / tensorflow_probability: the eager numpy shim (oracle/tf_shim), so that the
    # reference's criteria/*.py and the TF halves of utils.py run unmodified; or the lazy graph shim
    if GRAPH_MODE:
        from oracle.tf_shim import graph
        if sys.modules.get("tensorflow") is not graph:
            graph.install()
    else:
        from oracle import tf_shim
        if not tf_shim.installed():
            tf_shim.install()
    for name in (
        "matplotlib",
        "matplotlib.pyplot",
        "keras",
        "keras.datasets",
        "keras.datasets.cifar10",
        "keras.models",
        "keras.layers",
        "keras.optimizers",
        "keras.utils",
        "keras.backend",
        "tfdeterminism",
        "GPy",
    ):
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__name__ = name
            m.__path__ = []
            sys.modules[name] = m


_cache = {}


def load(module_name):
    """Import `module_name` (e.g. 'utils', 'optfunc', 'ep') from the reference tree."""
    if module_name in _cache:
        return _cache[module_name]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    cwd = os.getcwd()
    sys.path.insert(0, REFERENCE_ROOT)
    os.chdir(REFERENCE_ROOT)  # functions.py reads pHdata/ by relative path
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            # make sure we get the reference's module, not a same-named one of ours
            for k in ("utils", "optfunc", "ep", "empirical_approximation", "functions", "criteria",
                      "utils_for_continuous"):
                mod = sys.modules.get(k)
                where = getattr(mod, "__file__", None) or "".join(getattr(mod, "__path__", []))
                if mod is not None and not where.startswith(REFERENCE_ROOT):
                    del sys.modules[k]
            mod = importlib.import_module(module_name)
            if module_name == "utils" and not hasattr(mod, "ep"):
                # utils.get_testidxs_stats(mode='ep') raises NameError as shipped
                # because utils.py never imports ep.  The oracle calls ep.approximate_EP_np
                # with the arguments utils.py:1820-1825 intends; we do NOT patch utils here.
                pass
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_ROOT)
    _cache[module_name] = mod
    return mod


@contextlib.contextmanager
def in_reference_cwd():
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
