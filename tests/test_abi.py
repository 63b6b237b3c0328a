"""The C-ABI library loads and exports every function include/smallz4_b200.h declares; without a
GPU its entry points fail loudly instead of falling back to a CPU implementation."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smallz4_b200.h")
LIB = os.path.join(ROOT, "smallz4_b200", "libsmallz4_b200.so")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sz4_[a-z0-9_]+)\s*\(", text)) - {"sz4_get_bytes", "sz4_send_bytes"})


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ["sz4_create", "sz4_destroy", "sz4_lz4", "sz4_compress_host", "sz4_compress_host_range", "sz4_compress_device", "sz4_compress_bound",
                 "sz4_version", "sz4_last_error", "sz4_frame_header", "sz4_frame_end", "sz4_set_option", "sz4_last_stats"]:
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_version_and_frame_constants(lib):
    lib.sz4_version.restype = ctypes.c_char_p
    assert lib.sz4_version().decode().startswith("1.5")           # reference: smallz4::getVersion() == "1.5"
    buf = ctypes.create_string_buffer(16)
    lib.sz4_frame_header.restype = ctypes.c_size_t
    lib.sz4_frame_end.restype = ctypes.c_size_t
    assert lib.sz4_frame_header(buf, 0) == 7 and buf.raw[:7] == bytes([0x04, 0x22, 0x4D, 0x18, 0x40, 0x70, 0xDF])   # smallz4.h:488-494
    assert lib.sz4_frame_header(buf, 1) == 4 and buf.raw[:4] == bytes([0x02, 0x21, 0x4C, 0x18])                     # smallz4.h:482
    assert lib.sz4_frame_end(buf, 0) == 4 and lib.sz4_frame_end(buf, 1) == 0
    lib.sz4_compress_bound.restype = ctypes.c_size_t
    lib.sz4_compress_bound.argtypes = [ctypes.c_size_t, ctypes.c_int]
    assert lib.sz4_compress_bound(0, 0) >= 11 and lib.sz4_compress_bound(1 << 20, 0) >= (1 << 20) + 15


def test_no_cpu_fallback_without_a_gpu(lib):
    """sz4_create must fail (SZ4_ERR_CUDA) when no CUDA device is usable -- nothing computes on the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    h = ctypes.c_void_p()
    assert lib.sz4_create(ctypes.byref(h), 0) == -1
    assert not h.value
