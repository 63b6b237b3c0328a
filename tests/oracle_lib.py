"""ctypes access to the CPU checker (oracle/) -- test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LEVEL_CHAIN = {lvl: lvl for lvl in range(9)}
LEVEL_CHAIN[9] = 65535          # smallz4.cpp:175,233


class Opts(ctypes.Structure):
    _fields_ = [("max_chain", ctypes.c_uint32), ("legacy", ctypes.c_int), ("block_size", ctypes.c_uint32),
                ("dict", ctypes.c_void_p), ("dict_len", ctypes.c_size_t)]


class Stats(ctypes.Structure):
    _fields_ = [("blocks", ctypes.c_uint64), ("raw_blocks", ctypes.c_uint64),
                ("oob_first_reads", ctypes.c_uint64), ("selfmatch_skips", ctypes.c_uint64),
                ("chain_hops", ctypes.c_uint64), ("searches", ctypes.c_uint64)]


class Trace(ctypes.Structure):
    _fields_ = [("prev_exact", ctypes.c_void_p), ("len_found", ctypes.c_void_p), ("dist_found", ctypes.c_void_p),
                ("len_final", ctypes.c_void_p), ("cost", ctypes.c_void_p), ("skipped", ctypes.c_void_p), ("hops", ctypes.c_void_p)]


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


_ORACLE = None
_REF = None


def oracle():
    global _ORACLE
    if _ORACLE is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build_oracle()
        lib = ctypes.CDLL(so)
        lib.sz4o_bound.argtypes = [ctypes.c_size_t, ctypes.POINTER(Opts)]
        lib.sz4o_bound.restype = ctypes.c_size_t
        lib.sz4o_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Opts), ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.POINTER(Stats), ctypes.POINTER(Trace)]
        lib.sz4o_compress.restype = ctypes.c_int64
        lib.sz4o_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_size_t]
        lib.sz4o_decompress.restype = ctypes.c_int64
        _ORACLE = lib
    return _ORACLE


def reference():
    """The unmodified reference behind oracle/ref_shim.cpp, or None if oracle/_ref was never built."""
    global _REF
    if _REF is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libsmallz4ref.so")
        if not os.path.exists(so):
            if os.path.exists("/root/reference/smallz4.h"):
                build_oracle()
            if not os.path.exists(so):
                return None
        lib = ctypes.CDLL(so)
        lib.ref_smallz4_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_uint, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        lib.ref_smallz4_compress.restype = ctypes.c_longlong
        _REF = lib
    return _REF


def _u8(x):
    a = np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else np.ascontiguousarray(x, dtype=np.uint8)
    return a


def oracle_compress(data, level=9, legacy=False, dictionary=None, block_size=0, max_chain=None,
                    want_trace=False):
    src = _u8(data)
    d = _u8(dictionary) if dictionary is not None and len(dictionary) else None
    o = Opts(LEVEL_CHAIN[level] if max_chain is None else max_chain, int(legacy), block_size,
             d.ctypes.data if d is not None else None, d.size if d is not None else 0)
    cap = oracle().sz4o_bound(src.size, ctypes.byref(o))
    dst = np.empty(cap, dtype=np.uint8)
    st = Stats()
    tr = None
    arrays = {}
    if want_trace:
        total = src.size + (65535 if d is not None else 0)
        arrays = {"prev_exact": np.zeros(total, np.uint16), "len_found": np.zeros(total, np.uint32),
                  "dist_found": np.zeros(total, np.uint16), "len_final": np.zeros(total, np.uint32),
                  "cost": np.zeros(total, np.uint32), "skipped": np.zeros(total, np.uint8), "hops": np.zeros(total, np.uint32)}
        tr = Trace(*[arrays[k].ctypes.data for k in ("prev_exact", "len_found", "dist_found", "len_final", "cost", "skipped", "hops")])
    n = oracle().sz4o_compress(src.ctypes.data if src.size else None, src.size, ctypes.byref(o), dst.ctypes.data, cap,
                               ctypes.byref(st), ctypes.byref(tr) if tr is not None else None)
    assert n >= 0, "oracle compress failed"
    out = dst[:n].tobytes()
    stats = {f[0]: getattr(st, f[0]) for f in Stats._fields_}
    return (out, stats, arrays) if want_trace else (out, stats)


def oracle_decompress(frame, max_out, dictionary=None):
    src = _u8(frame)
    d = _u8(dictionary) if dictionary is not None and len(dictionary) else None
    dst = np.empty(max(max_out, 1), dtype=np.uint8)
    n = oracle().sz4o_decompress(src.ctypes.data, src.size, d.ctypes.data if d is not None else None,
                                 d.size if d is not None else 0, dst.ctypes.data, max_out)
    assert n >= 0, f"oracle decompress failed ({n})"
    return dst[:n].tobytes()


def reference_compress(data, level=9, legacy=False, dictionary=None, max_chain=None):
    lib = reference()
    assert lib is not None, "oracle/_ref not built"
    src = _u8(data)
    d = _u8(dictionary) if dictionary is not None and len(dictionary) else None
    cap = 2 * src.size + 4096
    dst = np.empty(cap, dtype=np.uint8)
    n = lib.ref_smallz4_compress(src.ctypes.data if src.size else None, src.size,
                                 d.ctypes.data if d is not None else None, d.size if d is not None else 0,
                                 LEVEL_CHAIN[level] if max_chain is None else max_chain, int(legacy),
                                 dst.ctypes.data, cap)
    assert n >= 0
    return dst[:n].tobytes()


def reference_cat():
    """Path of the reference's own decoder (oracle/_ref/smallz4cat, built from /root/reference/smallz4cat.c), or None."""
    exe = os.path.join(ORACLE_DIR, "_ref", "smallz4cat")
    if not os.path.exists(exe) and os.path.exists("/root/reference/smallz4cat.c"):
        build_oracle()
    return exe if os.path.exists(exe) else None


def reference_decompress(frame, dictionary=None):
    """Decode with the UNMODIFIED smallz4cat: frame on stdin, output on stdout (smallz4cat.c:362-420; with -D the
    file-name argument is not usable, so stdin it is)."""
    exe = reference_cat()
    assert exe is not None, "oracle/_ref/smallz4cat not built"
    cmd = [exe]
    tmp = None
    if dictionary is not None and len(dictionary):
        import tempfile
        tmp = tempfile.NamedTemporaryFile(suffix=".dict")
        tmp.write(bytes(dictionary)); tmp.flush()
        cmd += ["-D", tmp.name]
    r = subprocess.run(cmd, input=bytes(frame), capture_output=True)
    if tmp is not None:
        tmp.close()
    assert r.returncode == 0, f"smallz4cat failed: {r.stderr[:200]!r}"
    return r.stdout
