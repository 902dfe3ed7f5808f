"""Drop-in alias: the reference's own import lines

    import torchlens.ray_tracing_lite as rt
    import torchlens.lens_modeling as lm
    import torchlens.ray_tracing as rt            # the TensorFlow original's name for the same API

work unchanged against the B200 implementation.  The three names are the
very module objects of :mod:`torchoptics_b200` -- not copies -- so patches and private names agree.
(The reference itself is a namespace package without __init__.py; oracle/make_ref.py stages it under
oracle/_ref for the CPU arm of bench.py and loads it from THERE, ahead of this alias on sys.path.)
"""
import sys

from torchoptics_b200 import lens_modeling, ray_tracing_lite
from torchoptics_b200 import ray_tracing_lite as ray_tracing

for _name, _module in (('lens_modeling', lens_modeling), ('ray_tracing_lite', ray_tracing_lite),
                       ('ray_tracing', ray_tracing)):
    sys.modules[f'{__name__}.{_name}'] = _module
del _name, _module
