"""ctypes binding of ``libtorchoptics_b200.so`` (C ABI: include/torchoptics_b200.h).

There is no fallback: if the library has not been built, or a call fails, this
module raises.  Torch is used here only to hand over device pointers and the
current CUDA stream.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TL_LIB_OVERRIDE') or os.path.join(_PKG, 'libtorchoptics_b200.so')   # override: kernel experiments only

ABI_VERSION = 12
ARITH_GUARDED = 0
ARITH_EXACT = 1
AIM_REAL = 0
AIM_PARAXIAL = 1
PARAXIAL_FIRST_ORDER = 0
PARAXIAL_LAST_CURVATURE = 1
PARAXIAL_MAX_SLOTS = 64
MAX_SURFACES_FWD = 256
MAX_SURFACES_BWD = 32
MAX_SURFACES_SPOT = 32
MAX_SURFACES_GEN = 48
N_ASPHERE_TERMS = 7


class NativeLibraryError(RuntimeError):
    pass


class TlStrided(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('stride', ctypes.c_int64 * 4)]


class TlProblem(ctypes.Structure):
    _fields_ = [('x', TlStrided), ('y', TlStrided), ('z', TlStrided), ('cx', TlStrided),
                ('cy', TlStrided),
                ('c', ctypes.c_void_p), ('t', ctypes.c_void_p), ('mu', ctypes.c_void_p),
                ('live', ctypes.c_void_p),
                ('B', ctypes.c_int32), ('F', ctypes.c_int32), ('P', ctypes.c_int32),
                ('W', ctypes.c_int32), ('S', ctypes.c_int32),
                ('allow_backward_rays', ctypes.c_int32), ('arith', ctypes.c_int32),
                ('p_begin', ctypes.c_int32), ('p_end', ctypes.c_int32),
                ('xy_scale', ctypes.c_void_p),
                ('k', ctypes.c_void_p), ('a', ctypes.c_void_p), ('sd', ctypes.c_void_p),
                ('aim', ctypes.c_void_p), ('vig', ctypes.c_void_p)]


class TlTraceOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('x', 'y', 'cx', 'cy', 'ok', 'backward', 'opl',
                                               'z_relu', 'theta', 'theta_prime')]


class TlSeeds(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('gx', 'gy', 'gcx', 'gcy', 'gz_relu', 'gtheta', 'gtheta_prime',
                                               'gopl')]


class TlGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('gc', 'gt', 'gmu', 'gz_sum', 'gx', 'gy', 'gz', 'gcx', 'gcy',
                                               'gk', 'ga')]


class TlLens(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('c', 't', 'nd', 'v', 'mask', 'mask_g', 'stop_idx', 'hfov',
                                               'epd', 'rel_fields', 'wavelengths')] + \
               [(n, ctypes.c_int32) for n in ('B', 'L', 'F', 'W')]


class TlPenaltyOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('penalty', 'gc', 'gt', 'gmu', 'gz')]


class TlPsf(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('x', 'y', 'y_target', 'x_incr', 'y_incr', 'x_size', 'y_size')] + \
               [(n, ctypes.c_int32) for n in ('G', 'C', 'R', 'n_x_bins', 'n_y_bins')]


class TlParaxial(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('c', 't', 'n', 'live', 'glass')] + \
               [(n, ctypes.c_int32) for n in ('B', 'L', 'mode')]


class TlSpotOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('rms', 'rms_field', 'gc', 'gt', 'gmu', 'gz', 'gk', 'ga')]


EXPORTS = {
    # name: (restype, argtypes)
    'tl_abi_version': (ctypes.c_int, []),
    'tl_last_error': (ctypes.c_char_p, []),
    'tl_launch_count': (ctypes.c_int64, []),
    'tl_abi_describe': (ctypes.c_char_p, [ctypes.c_int32]),
    'tl_trace_fwd': (ctypes.c_int, [ctypes.POINTER(TlProblem), ctypes.POINTER(TlTraceOut), ctypes.c_void_p]),
    'tl_trace_bwd_workspace': (ctypes.c_size_t, [ctypes.POINTER(TlProblem)]),
    'tl_trace_bwd': (ctypes.c_int, [ctypes.POINTER(TlProblem), ctypes.POINTER(TlSeeds),
                                    ctypes.POINTER(TlGrads), ctypes.c_void_p, ctypes.c_size_t,
                                    ctypes.c_void_p]),
    'tl_rms_workspace': (ctypes.c_size_t, [ctypes.c_int32] * 4),
    'tl_rms_fwd': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int32] * 4 +
                   [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p]),
    'tl_rms_bwd': (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int32] * 4 +
                   [ctypes.c_void_p, ctypes.c_void_p]),
    'tl_spot_moment_count': (ctypes.c_int32, [ctypes.c_int32, ctypes.c_int32]),
    'tl_spot_moment_count_general': (ctypes.c_int32, [ctypes.c_int32, ctypes.c_int32]),
    'tl_spot_workspace': (ctypes.c_size_t, [ctypes.POINTER(TlProblem), ctypes.c_int32]),
    'tl_spot_kernel_name': (ctypes.c_char_p, [ctypes.POINTER(TlProblem), ctypes.c_int32]),
    'tl_spot_kernel_only': (ctypes.c_int, [ctypes.POINTER(TlProblem), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_void_p]),
    'tl_spot_accumulate': (ctypes.c_int, [ctypes.POINTER(TlProblem), ctypes.c_int32, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                          ctypes.c_void_p]),
    'tl_stage_fwd': (ctypes.c_int, [ctypes.POINTER(TlLens)] + [ctypes.c_void_p] * 5),
    'tl_stage_bwd': (ctypes.c_int, [ctypes.POINTER(TlLens)] + [ctypes.c_void_p] * 7),
    'tl_stage_ref': (ctypes.c_int, [ctypes.POINTER(TlLens), ctypes.POINTER(TlProblem)] + [ctypes.c_void_p] * 6 +
                     [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    'tl_spot_accumulate_ref': (ctypes.c_int, [ctypes.POINTER(TlProblem), ctypes.c_int32, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'tl_lens_spot_finalize': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(TlLens), ctypes.c_int64,
                                             ctypes.POINTER(TlSpotOut), ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    'tl_aim': (ctypes.c_int, [ctypes.POINTER(TlLens)] + [ctypes.c_void_p] * 5 +
               [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    'tl_spot_finalize': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int32] * 4 +
                         [ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(TlSpotOut), ctypes.c_void_p]),
    'tl_psf_workspace': (ctypes.c_size_t, [ctypes.POINTER(TlPsf)]),
    'tl_psf_bin': (ctypes.c_int, [ctypes.POINTER(TlPsf), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                  ctypes.c_void_p]),
    'tl_paraxial_fwd': (ctypes.c_int, [ctypes.POINTER(TlParaxial), ctypes.c_void_p, ctypes.c_void_p]),
    'tl_paraxial_bwd': (ctypes.c_int, [ctypes.POINTER(TlParaxial)] + [ctypes.c_void_p] * 5),
    'tl_penalty_moment_count': (ctypes.c_int32, [ctypes.c_int32]),
    'tl_penalty_workspace': (ctypes.c_size_t, [ctypes.POINTER(TlProblem)]),
    'tl_penalty_accumulate': (ctypes.c_int, [ctypes.POINTER(TlProblem), ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.c_void_p]),
    'tl_penalty_finalize': (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int32] * 4 +
                            [ctypes.c_double, ctypes.POINTER(TlPenaltyOut), ctypes.c_void_p]),
    'tl_peer_handle_bytes': (ctypes.c_size_t, []),
    'tl_peer_create': (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    'tl_peer_connect': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'tl_peer_allreduce_f64': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_void_p]),
    'tl_peer_status': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_uint32)]),
    'tl_peer_destroy': (ctypes.c_int, [ctypes.c_void_p]),
}

# struct ids of tl_abi_layout
LAYOUT_STRUCTS = (TlStrided, TlProblem, TlTraceOut, TlSeeds, TlGrads, TlSpotOut, TlPenaltyOut, TlLens, TlPsf, TlParaxial)

_lib = None


def describe(struct):
    """'Name:size;field@offset;...' of a ctypes struct, the format of tl_abi_describe."""
    return f'{struct.__name__}:{ctypes.sizeof(struct)}' + ''.join(
        f';{name}@{getattr(struct, name).offset}' for name, _ in struct._fields_)


def check_layout(lib, structs=None):
    """Every ctypes struct of this binding -- size, field NAMES in order and offsets -- against what
    the LOADED library reports (tl_abi_describe): a reordered, renamed, resized or re-typed field
    raises instead of corrupting a call."""
    for which, struct in enumerate(structs or LAYOUT_STRUCTS):
        got = lib.tl_abi_describe(which)
        got = got.decode() if got else None
        if got != describe(struct):
            raise NativeLibraryError(f'binding layout {describe(struct)!r} != library layout {got!r}; '
                                     'rebuild libtorchoptics_b200.so or fix _native.py')


def load():
    """Load the shared library (once) and declare every export of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f'{LIB_PATH} is missing: build it with `python -m torchoptics_b200.build` '
            '(or __graft_entry__.build()).  torchoptics_b200 has no CPU or eager fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.tl_abi_version() != ABI_VERSION:
        raise NativeLibraryError('libtorchoptics_b200.so: ABI version mismatch, rebuild it')
    check_layout(lib)
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().tl_last_error().decode(errors='replace')
        raise NativeLibraryError(f'{what} failed ({code}): {msg}')


def launch_count():
    return int(load().tl_launch_count())


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_ptr(device):
    """The current stream of ``device`` as a ``cudaStream_t``.  (``torch.cuda.current_stream(device)`` builds a
    Stream object behind a device-index lookup: ~4 us per call, four calls in an eager step of ~0.5 ms.)"""
    if _raw_stream is not None:
        idx = device.index if isinstance(device, torch.device) else torch.device(device).index
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device() if idx is None else idx))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class on_device:
    """``with torch.cuda.device(dev)`` that costs nothing when ``dev`` is already the current device."""
    __slots__ = ('idx', 'prev')

    def __init__(self, device):
        idx = device.index if isinstance(device, torch.device) else torch.device(device).index
        self.idx = idx
        self.prev = None

    def __enter__(self):
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
            self.prev = None
        return False


def require_cuda(t, name):
    if not t.is_cuda:
        raise NativeLibraryError(
            f'{name} is on {t.device}: the ray-trace kernels run on CUDA devices only '
            '(there is no CPU path; use the oracle under oracle/ for CPU checks)')


def strided(t, shape):
    """Broadcast view of ``t`` over ``shape`` as (pointer, element strides)."""
    v = torch.broadcast_to(t, shape)
    s = TlStrided()
    s.ptr = v.data_ptr()
    for i, st in enumerate(v.stride()):
        s.stride[i] = st
    return s
