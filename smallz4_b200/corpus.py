"""Deterministic synthetic corpora (ctypes binding of csrc/corpus.c).

Every byte is a function of (kind, seed, absolute offset), so a rank can generate its own
shard plus the 64 KiB halo in front of it.  Kinds follow BASELINE.json: text-like Markov,
mixed binary, zeros/runs, incompressible random, and "mixed" (256 KiB regions of those).
"""
import ctypes
import os
import subprocess

import numpy as np

KINDS = {"text": 0, "binary": 1, "runs": 2, "zeros": 3, "random": 4, "mixed": 5}
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsz4corpus.so")
        src = os.path.join(_HERE, "csrc", "corpus.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", src, "-o", so])
        _LIB = ctypes.CDLL(so)
        _LIB.sz4_corpus_fill.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.c_int, ctypes.c_uint64]
        _LIB.sz4_corpus_fill.restype = None
    return _LIB


def fill(out: np.ndarray, kind: str, seed: int = 1, offset: int = 0) -> np.ndarray:
    """Fill a uint8 array with corpus bytes [offset, offset+len(out))."""
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
    _lib().sz4_corpus_fill(out.ctypes.data, offset, out.size, KINDS[kind], seed)
    return out


def make(kind: str, nbytes: int, seed: int = 1, offset: int = 0) -> np.ndarray:
    return fill(np.empty(nbytes, dtype=np.uint8), kind, seed, offset)
