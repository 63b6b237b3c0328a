#!/usr/bin/env python
"""Per-block golden digests of the bench corpus at its FULL size, from the UNMODIFIED reference.

BASELINE.json configs[1]/[2]: 256 MB of the mixed corpus (seed 1), 4 MiB blocks, levels 1..9.  The
scalar reference needs about a minute per block and level-9 core, so the whole stream cannot be one
job; but a block record depends only on the block and the 64 KiB (+ 12 bytes) in front of it
(DESIGN.md "Independence of blocks"; smallz4.h:615-629, 798-804).  So block k is produced by the
reference run over the window of blocks [k-1, k] (for k = 0: block 0 alone) and the record of the
window's last block is kept.  `--prove` checks that claim against whole-stream reference runs.

Output: tests/golden/golden_blocks.json
  {"kind", "seed", "size", "block", "levels": {"9": [{"n": record bytes, "sha256": ...} x blocks]}}
A record is what the reference sends for one block: 4-byte size word + payload (smallz4.h:769-780).

Run in the build container (needs /root/reference -> oracle/_ref):
  python tests/golden/make_golden_blocks.py --levels 9,1,2,3,4,5,6,7,8 --workers 8
Finished jobs are appended to a partial file, so the script can be interrupted and resumed.
"""
import argparse
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

KIND, SEED, SIZE, BLOCK = "mixed", 1, 256 << 20, 4 << 20
OUT = os.path.join(HERE, "golden_blocks.json")
PARTIAL = os.path.join(HERE, "_golden_blocks_partial.jsonl")


def split_records(frame, legacy=False):
    """[bytes] -- the block records of a modern frame (header 7 bytes, end mark 4 zero bytes)."""
    at, out = 7, []
    while True:
        word = int.from_bytes(frame[at:at + 4], "little")
        if word == 0:
            break
        n = word & 0x7FFFFFFF
        out.append(frame[at:at + 4 + n])
        at += 4 + n
    assert at + 4 == len(frame)
    return out


def window_record(k, level, kind=KIND, seed=SEED, size=SIZE, block=BLOCK):
    """Record of block k of the stream, from the reference run over blocks [k-1, k]."""
    from oracle_lib import reference_compress
    from smallz4_b200 import corpus
    lo = max(k - 1, 0) * block
    hi = min((k + 1) * block, size)
    data = corpus.make(kind, hi - lo, seed, offset=lo)
    recs = split_records(reference_compress(data, level))
    assert len(recs) == (2 if k > 0 else 1)
    return recs[-1]


def _job(args):
    k, level = args
    t = time.perf_counter()
    rec = window_record(k, level)
    return {"level": level, "block": k, "n": len(rec), "sha256": hashlib.sha256(rec).hexdigest(),
            "seconds": round(time.perf_counter() - t, 1)}


def prove(levels, nblocks=3):
    """Window records == records of a whole-stream reference run (first `nblocks` blocks of the corpus, and a
    stretch that starts in the middle of the corpus is covered by the block digests themselves)."""
    from oracle_lib import reference_compress
    from smallz4_b200 import corpus
    size = nblocks * BLOCK - 777                     # last block short, like the end of a real stream
    data = corpus.make(KIND, size, SEED)
    for level in levels:
        whole = split_records(reference_compress(data, level))
        for k in range(nblocks):
            w = window_record(k, level, size=size)
            assert w == whole[k], f"level {level} block {k}: window record differs from the whole-stream record"
        print(f"level {level}: {nblocks} window records equal the whole-stream records", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", default="9,1,2,3,4,5,6,7,8")
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--prove", action="store_true")
    args = ap.parse_args()
    levels = [int(x) for x in args.levels.split(",")]
    if args.prove:
        prove(levels)
        return
    nblocks = SIZE // BLOCK
    done = {}
    if os.path.exists(PARTIAL):
        with open(PARTIAL) as f:
            for line in f:
                r = json.loads(line)
                done[(r["level"], r["block"])] = r
    jobs = [(k, lvl) for lvl in levels for k in range(nblocks) if (lvl, k) not in done]
    print(f"{len(done)} records present, {len(jobs)} to do", flush=True)
    if jobs:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(args.workers) as pool, open(PARTIAL, "a") as f:
            for r in pool.imap_unordered(_job, jobs, chunksize=1):
                done[(r["level"], r["block"])] = r
                f.write(json.dumps(r) + "\n")
                f.flush()
                print(r["level"], r["block"], r["n"], r["seconds"], flush=True)
    out = {"generator": "tests/golden/make_golden_blocks.py", "reference": "smalLZ4 1.5 (/root/reference), unmodified",
           "kind": KIND, "seed": SEED, "size": SIZE, "block": BLOCK, "levels": {}}
    for lvl in sorted({l for l, _ in done}):
        if all((lvl, k) in done for k in range(nblocks)):
            out["levels"][str(lvl)] = [{"n": done[(lvl, k)]["n"], "sha256": done[(lvl, k)]["sha256"]} for k in range(nblocks)]
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", OUT, "levels", sorted(out["levels"]), flush=True)


if __name__ == "__main__":
    main()
