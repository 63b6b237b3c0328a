"""ctypes binding of libsmallz4_b200.so (include/smallz4_b200.h).

Mirrors the reference's public interface (smallz4.h:38-80): ``lz4(get_bytes, send_bytes,
max_chain_length, dictionary, use_legacy_format)``, ``get_version()``, the level thresholds, and
the CLI's level mapping (smallz4.cpp:175,233).  All compute happens in the CUDA library; if it is
missing or no GPU is present the calls raise -- there is no CPU path.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libsmallz4_b200.so")

SHORT_CHAINS_GREEDY = 3      # smallz4::ShortChainsGreedy
SHORT_CHAINS_LAZY = 6        # smallz4::ShortChainsLazy
MAX_CHAIN_LENGTH = 65535     # smallz4::MaxChainLength

GET_BYTES = ctypes.CFUNCTYPE(ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)
SEND_BYTES = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


class Sz4Error(RuntimeError):
    pass


def level_to_chain(level: int) -> int:
    """-0..-8 -> 0..8, -9 -> 65535 (smallz4.cpp:175,232-239)."""
    if not 0 <= level <= 9:
        raise ValueError("level must be 0..9")
    return MAX_CHAIN_LENGTH if level == 9 else level


def load_library(path=None):
    path = path or os.environ.get("SMALLZ4_B200_LIB", DEFAULT_LIB)
    if not os.path.exists(path):
        raise Sz4Error(f"{path} not found: build it with `python __graft_entry__.py build` (nvcc, sm_100a); "
                       "there is no CPU fallback")
    lib = ctypes.CDLL(path)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.sz4_create.argtypes = [ctypes.POINTER(vp), i32]; lib.sz4_create.restype = i32
    lib.sz4_destroy.argtypes = [vp]; lib.sz4_destroy.restype = None
    lib.sz4_last_error.argtypes = [vp]; lib.sz4_last_error.restype = ctypes.c_char_p
    lib.sz4_version.argtypes = []; lib.sz4_version.restype = ctypes.c_char_p
    lib.sz4_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_longlong]; lib.sz4_set_option.restype = i32
    lib.sz4_compress_bound.argtypes = [sz, i32]; lib.sz4_compress_bound.restype = sz
    lib.sz4_lz4.argtypes = [vp, GET_BYTES, SEND_BYTES, ctypes.c_ushort, vp, sz, i32, vp]; lib.sz4_lz4.restype = i32
    lib.sz4_compress_host.argtypes = [vp, vp, sz, vp, sz, ctypes.POINTER(sz), ctypes.c_ushort, vp, sz, i32]
    lib.sz4_compress_host.restype = i32
    lib.sz4_compress_device.argtypes = [vp, vp, sz, sz, i32, i32, vp, sz, ctypes.POINTER(sz), ctypes.c_ushort, i32, vp]
    lib.sz4_compress_device.restype = i32
    lib.sz4_compress_host_range.argtypes = [vp, vp, sz, sz, i32, i32, vp, sz, ctypes.POINTER(sz), ctypes.c_ushort, i32]
    lib.sz4_compress_host_range.restype = i32
    lib.sz4_frame_header.argtypes = [vp, i32]; lib.sz4_frame_header.restype = sz
    lib.sz4_frame_end.argtypes = [vp, i32]; lib.sz4_frame_end.restype = sz
    lib.sz4_last_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong)]
    lib.sz4_last_stats.restype = i32
    lib.sz4_last_phase_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]; lib.sz4_last_phase_ms.restype = i32
    lib.sz4_last_dp_redos.argtypes = [vp]; lib.sz4_last_dp_redos.restype = ctypes.c_longlong
    lib.sz4_last_path_redos.argtypes = [vp]; lib.sz4_last_path_redos.restype = ctypes.c_longlong
    lib.sz4_debug_fetch.argtypes = [vp, ctypes.c_char_p, vp, sz]; lib.sz4_debug_fetch.restype = i32
    return lib


def _u8(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8)
    return np.frombuffer(bytes(x), dtype=np.uint8)


class Compressor:
    """One compression context (one CUDA stream, reusable device buffers) on one GPU."""

    def __init__(self, device=-1, lib_path=None, **options):
        self.lib = load_library(lib_path)
        h = ctypes.c_void_p()
        rc = self.lib.sz4_create(ctypes.byref(h), device)
        if rc != 0:
            raise Sz4Error("sz4_create failed: no usable CUDA device (this library has no CPU path)")
        self.h = h
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sz4_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise Sz4Error(f"smallz4_b200 error {rc}: {self.lib.sz4_last_error(self.h).decode()}")

    def set_option(self, name, value):
        self._check(self.lib.sz4_set_option(self.h, name.encode(), int(value)))

    @staticmethod
    def get_version():
        return load_library().sz4_version().decode()

    def compress(self, data, level=9, dictionary=None, use_legacy_format=False, max_chain_length=None) -> bytes:
        """Host buffer in, complete .lz4 frame out (sz4_compress_host)."""
        src = _u8(data)
        d = _u8(dictionary) if dictionary is not None and len(dictionary) else None
        chain = level_to_chain(level) if max_chain_length is None else max_chain_length
        cap = self.lib.sz4_compress_bound(src.size, int(use_legacy_format)) + (src.size if d is not None else 0) + 64
        dst = np.empty(cap, dtype=np.uint8)
        n = ctypes.c_size_t(0)
        self._check(self.lib.sz4_compress_host(self.h, src.ctypes.data if src.size else None, src.size, dst.ctypes.data, cap,
                                               ctypes.byref(n), chain, d.ctypes.data if d is not None else None,
                                               d.size if d is not None else 0, int(use_legacy_format)))
        return dst[:n.value].tobytes()

    def compress_into(self, src_ptr, n, dst_ptr, cap, level=9, use_legacy_format=False):
        """Raw host pointers (e.g. pinned torch tensors): returns the frame length."""
        out = ctypes.c_size_t(0)
        self._check(self.lib.sz4_compress_host(self.h, src_ptr, n, dst_ptr, cap, ctypes.byref(out), level_to_chain(level),
                                               None, 0, int(use_legacy_format)))
        return out.value

    def compress_device(self, d_src_ptr, halo, n, d_dst_ptr, cap, level=9, first=True, last=True, use_legacy_format=False,
                        stream=0):
        """Device pointers: whole blocks with their halo in, [size][payload] block records out."""
        out = ctypes.c_size_t(0)
        self._check(self.lib.sz4_compress_device(self.h, d_src_ptr, halo, n, int(first), int(last), d_dst_ptr, cap,
                                                 ctypes.byref(out), level_to_chain(level), int(use_legacy_format), stream))
        return out.value

    def compress_range_into(self, src_ptr, halo, n, dst_ptr, cap, level=9, first=False, last=True, use_legacy_format=False):
        """Host pointers: `halo` bytes of history then `n` bytes of whole blocks in, block records out (host)."""
        out = ctypes.c_size_t(0)
        self._check(self.lib.sz4_compress_host_range(self.h, src_ptr, halo, n, int(first), int(last), dst_ptr, cap,
                                                     ctypes.byref(out), level_to_chain(level), int(use_legacy_format)))
        return out.value

    def lz4(self, get_bytes, send_bytes, max_chain_length=MAX_CHAIN_LENGTH, dictionary=None, use_legacy_format=False):
        """Drop-in for smallz4::lz4 with Python callables: get_bytes(n) -> bytes, send_bytes(b)."""
        def _get(buf, n, _user):
            chunk = get_bytes(n)
            ctypes.memmove(buf, chunk, len(chunk))
            return len(chunk)

        def _send(buf, n, _user):
            send_bytes(ctypes.string_at(buf, n))

        d = _u8(dictionary) if dictionary is not None and len(dictionary) else None
        g, s = GET_BYTES(_get), SEND_BYTES(_send)
        self._check(self.lib.sz4_lz4(self.h, g, s, max_chain_length, d.ctypes.data if d is not None else None,
                                     d.size if d is not None else 0, int(use_legacy_format), None))

    def last_stats(self):
        ms, launches = ctypes.c_double(0), ctypes.c_ulonglong(0)
        self.lib.sz4_last_stats(self.h, ctypes.byref(ms), ctypes.byref(launches))
        return ms.value, launches.value

    PHASES = ("sort", "chain", "search", "fixup", "dp", "path", "emit")

    def last_phase_ms(self):
        arr = (ctypes.c_double * 7)()
        self.lib.sz4_last_phase_ms(self.h, arr)
        return dict(zip(self.PHASES, list(arr)))

    def last_dp_redos(self):
        return self.lib.sz4_last_dp_redos(self.h)

    def last_path_redos(self):
        return self.lib.sz4_last_path_redos(self.h)

    def debug_fetch(self, what, count):
        dt = {"pe": np.uint16, "ph": np.uint16, "pe8": np.uint16, "jump": np.uint64, "len_found": np.uint32, "dist_found": np.uint16,
              "len_final": np.uint32, "dist_final": np.uint16, "cost": np.uint32}[what]
        out = np.zeros(count, dtype=dt)
        self._check(self.lib.sz4_debug_fetch(self.h, what.encode(), out.ctypes.data, count))
        return out
