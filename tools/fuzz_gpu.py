#!/usr/bin/env python
"""Randomised differential test on a GPU: many small inputs (corpus kinds glued together at random), random level,
frame format, block size (131072 so that block borders and the twice-inserted positions are frequent), batch size and
hand-over threshold, now and then a dictionary, each frame compared with the oracle's.

    python tools/fuzz_gpu.py --seconds 300 [--seed 1]
"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import oracle_compress  # noqa: E402
from smallz4_b200 import corpus  # noqa: E402
from smallz4_b200.api import Compressor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rnd = random.Random(a.seed)
    c = Compressor(device=0)
    t0, n, total = time.time(), 0, 0
    while time.time() - t0 < a.seconds:
        parts = []
        for _ in range(rnd.randint(1, 4)):
            kind = rnd.choice(["text", "text", "binary", "binary", "runs", "zeros", "random", "mixed"])
            size = rnd.choice([rnd.randint(0, 40), rnd.randint(0, 3000), rnd.randint(0, 150_000)])
            piece = corpus.make(kind, size, rnd.randint(1, 50), offset=rnd.choice([0, 65536 * rnd.randint(0, 1000)])).tobytes()
            if rnd.random() < 0.2 and parts:
                piece = parts[0][: rnd.randint(0, 5000)] + piece          # far repeats of the beginning
            parts.append(piece)
        data = b"".join(parts)
        level = rnd.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9])
        legacy = rnd.random() < 0.15
        bs = rnd.choice([131072, 131072, 196608, 0])
        opts = {"block_size": bs, "batch_blocks": rnd.choice([1, 2, 3, 64]), "long_age": rnd.choice([0, 1, 8]), "tail_lanes": rnd.choice([0, 12])}
        for k, v in opts.items():
            c.set_option(k, v)
        dictionary = None
        if not legacy and rnd.random() < 0.12:                           # -D (round-1 path; long runs go through the opt-in replay)
            dictionary = corpus.make(rnd.choice(["text", "binary", "zeros"]), rnd.choice([1, 100, 5000, 65536, 70000]), rnd.randint(1, 9), offset=1 << 40).tobytes()
            c.set_option("allow_scalar_dict", 1)
            data = data[:200_000]
        got = c.compress(data, level=level, use_legacy_format=legacy, dictionary=dictionary)
        c.set_option("allow_scalar_dict", 0)
        want, _ = oracle_compress(data, level, legacy, dictionary, block_size=bs)
        if got != want:
            path = os.path.join(ROOT, "gpurun_out", f"fuzz_fail_{n}.bin")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            open(path, "wb").write(data)
            print(f"MISMATCH case {n}: {len(data)} bytes, level {level}, legacy {legacy}, {opts} -> {path}", flush=True)
            sys.exit(1)
        n += 1
        total += len(data)
    print(f"fuzz ok: {n} cases, {total / 1e6:.1f} MB, seed {a.seed}, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
