#!/usr/bin/env python
"""Per-phase device times (CUDA events inside the library) for each corpus kind, plus the DP counters.

    python tools/profile_phases.py [--size-mb 256] [--level 9] [kinds ...]

Needs a GPU.  Prints one line per kind; used to decide what to optimise next (see profiles/)."""
import argparse
import ctypes
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from smallz4_b200 import corpus  # noqa: E402
from smallz4_b200.api import Compressor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kinds", nargs="*", default=["mixed", "text", "binary", "runs", "zeros", "random"])
    ap.add_argument("--size-mb", type=int, default=256)
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--opt", action="append", default=[], help="name=value for sz4_set_option (repeatable)")
    a = ap.parse_args()
    c = Compressor(device=0, profile=1)
    for o in a.opt:
        name, value = o.split("=")
        c.set_option(name, int(value))
    c.lib.sz4_debug_counters.restype = ctypes.POINTER(ctypes.c_uint)
    c.lib.sz4_debug_counters.argtypes = [ctypes.c_void_p]
    khz = 1965e3
    for kind in a.kinds:
        data = corpus.make(kind, a.size_mb << 20, 1)
        c.compress(data, level=a.level)
        frame = c.compress(data, level=a.level)
        ms, launches = c.last_stats()
        v = [c.lib.sz4_debug_counters(c.h)[i] for i in range(12)]
        print(f"{kind:7s} L{a.level} {ms:7.1f} ms  {data.size / ms / 1e6:6.3f} GB/s  ratio {data.size / len(frame):6.2f}  "
              f"sha {hashlib.sha256(frame).hexdigest()[:10]}  "
              f"{ {k: round(x, 1) for k, x in c.last_phase_ms().items()} }  dp redos {c.last_dp_redos()}, path redos {c.last_path_redos()}")
        print(f"        dp_spec longest task {v[4] * 1024 / khz:6.2f} ms, all tasks {v[5] * 1024 / khz:8.0f} ms-warp; "
              f"dp_verify slowest block {v[8] * 1024 / khz:6.2f} ms (redo {v[9] * 1024 / khz:6.2f})", flush=True)


if __name__ == "__main__":
    main()
