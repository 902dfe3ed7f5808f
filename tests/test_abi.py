"""The C-ABI library loads on a CPU-only box and exports every function that
include/torchoptics_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from torchoptics_b200 import _native
from torchoptics_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'torchoptics_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tl_[a-z_0-9]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    build_library()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 16
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_native.EXPORTS), set(names) ^ set(_native.EXPORTS)


def test_abi_version_and_argument_checks():
    lib = _native.load()
    assert lib.tl_abi_version() == _native.ABI_VERSION == 12
    assert lib.tl_spot_moment_count(11, 1) == 6 * 11 + 5
    assert lib.tl_spot_moment_count(11, 0) == 3
    assert lib.tl_rms_workspace(1, 3, 64, 3) > 0
    # NULL problem -> error code and message, never a crash
    assert lib.tl_trace_fwd(None, None, None) == -1
    assert b'NULL' in lib.tl_last_error()


def test_struct_layout_matches_header():
    # TlStrided = pointer + 4 x int64; TlProblem packs 5 of them, 4 pointers, 9 int32
    assert ctypes.sizeof(_native.TlStrided) == 40
    assert ctypes.sizeof(_native.TlProblem) == 5 * 40 + 4 * 8 + 9 * 4 + 4 + 8 + 3 * 8 + 8 + 8
    assert ctypes.sizeof(_native.TlLens) == 11 * 8 + 4 * 4
    assert ctypes.sizeof(_native.TlGrads) == 11 * 8


def test_field_names_and_offsets_match_the_library():
    """Names and offsets, not just sizes: every field of every struct against what the library
    exports (a reordered field of equal size passes a size check -- and an offset-only check)."""
    import pytest
    lib = _native.load()
    _native.check_layout(lib)
    text = lib.tl_abi_describe(1).decode()
    assert text.startswith(f'TlProblem:{ctypes.sizeof(_native.TlProblem)};x@0;y@40;')
    assert text.count(';') == len(_native.TlProblem._fields_)
    assert lib.tl_abi_describe(99) is None
    # two pointer fields swapped: same size, same offsets by position -- caught through the names
    fields = list(_native.TlSpotOut._fields_)
    fields[2], fields[3] = fields[3], fields[2]
    swapped = type('TlSpotOut', (ctypes.Structure,), {'_fields_': fields})
    assert ctypes.sizeof(swapped) == ctypes.sizeof(_native.TlSpotOut)
    structs = list(_native.LAYOUT_STRUCTS)
    structs[5] = swapped
    with pytest.raises(_native.NativeLibraryError):
        _native.check_layout(lib, structs)
