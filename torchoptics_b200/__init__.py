"""torchoptics_b200 -- the differentiable sequential ray-trace hot path of
TorchOptics/TorchLens as hand-written sm_100a CUDA kernels behind the reference's
Python API (``ray_tracing_lite`` / ``ray_tracing`` / ``lens_modeling``).

The CUDA library is a plain C-ABI shared object (include/torchoptics_b200.h) built
in-tree by ``python -m torchoptics_b200.build``; importing the package does not
need it, calling any traced function does (there is no CPU or eager fallback).
"""
from . import lens_modeling, ray_tracing_lite          # noqa: F401
from . import ray_tracing_lite as ray_tracing          # noqa: F401  same API as the TF original
from .lens_modeling import Lens, Specs, Structure      # noqa: F401
from .ray_tracing_lite import RayTracer, compute_rms2d, trace_skew   # noqa: F401
from .graph import GraphedSpotStep                      # noqa: F401

__version__ = '0.1.0'
