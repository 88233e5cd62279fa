"""Mirrors of the reference's criteria/*.py for TES (same module and function names)."""
from . import evaluate_batch_mp, evaluate_batch_sample_mp, evaluate_mp, evaluate_mp_lite, evaluate_sample_mp  # noqa: F401
