#!/usr/bin/env python
"""Debugging aid: duration of every k_search tile of one batch (how the tail of the kernel was found).

    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -DSZ4_TILE_STATS \\
         smallz4_b200/csrc/sz4_pipeline.cu -o variants/tilestats.so
    SMALLZ4_B200_LIB=variants/tilestats.so python tools/tile_stats.py mixed 256
"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smallz4_b200 import corpus
from smallz4_b200.api import Compressor
kind = sys.argv[1] if len(sys.argv) > 1 else "mixed"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c = Compressor(device=0)
d = corpus.make(kind, mb << 20, 1)
c.compress(d, level=9); c.compress(d, level=9)
n = min(1 << 16, (mb // 4) * (((4 << 20) + 11263) // 11264))
us = np.zeros(n, np.uint32); t0 = np.zeros(n, np.uint32)
c.lib.sz4_debug_tile_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint]
assert c.lib.sz4_debug_tile_stats(us.ctypes.data, t0.ctypes.data, n) == 0
t0 = (t0 - t0.min()).astype(np.int64); end = t0 + us
print(f"{kind}: {n} tiles, mean {us.mean():.0f} us, median {np.median(us):.0f}, p99 {np.percentile(us, 99):.0f}, max {us.max()} us; kernel {end.max() / 1000:.1f} ms")
print("sum of tile times / 148 SMs = %.1f ms" % (us.sum() / 148 / 1000))
order = np.argsort(-us.astype(np.int64))[:12]
tiles_per_block = ((4 << 20) + 11263) // 11264
for k in order:
    blk, t = divmod(int(k), tiles_per_block)
    pos = blk * (4 << 20) + t * 11264
    print(f"  tile {k:6d} (block {blk}, offset {pos >> 16} x 64 KiB) {us[k]:7d} us, started at {t0[k] / 1000:7.1f} ms, ended {end[k] / 1000:7.1f} ms")
late = np.argsort(-end)[:8]
print("last to finish:", [(int(k), int(us[k]), round(end[k] / 1000, 1)) for k in late])
