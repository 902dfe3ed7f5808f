"""Alias of :mod:`torchoptics_b200.ray_tracing_lite`.

The reference's ``ray_tracing.py`` is the TensorFlow original of
``ray_tracing_lite.py`` with the same function names and signatures (minus
``default_device``); code written against either module imports this one."""
from .ray_tracing_lite import *            # noqa: F401,F403
from .ray_tracing_lite import RayTracer, compute_rms2d, trace_skew   # noqa: F401
