#!/usr/bin/env python
"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libsmallz4ref.so).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Each case is (corpus kind, seed, size, level, legacy, dictionary spec); inputs are regenerated
from smallz4_b200.corpus, so only digests are stored.  A few tiny cases also store the full frame.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle_lib import reference_compress  # noqa: E402
from smallz4_b200 import corpus  # noqa: E402

MB = 1 << 20


def dict_bytes(spec):
    """spec = None | [kind, seed, size] -> dictionary bytes generated from the corpus at offset 1<<40."""
    if not spec:
        return None
    kind, seed, size = spec
    return corpus.make(kind, size, seed, offset=1 << 40).tobytes()


def cases():
    out = []
    # edge sizes around the reference's block-end rules (12 / 5 bytes) at several levels
    for n in [0, 1, 4, 5, 11, 12, 13, 14, 17, 18, 19, 31, 64, 255, 256, 1000]:
        for lvl in [1, 4, 9]:
            out.append(("text", 7, n, lvl, False, None))
    out.append(("text", 7, 12, 9, True, None))
    out.append(("zeros", 1, 13, 9, False, None))
    # every level on every kind, one block
    for kind in ["text", "binary", "mixed", "random"]:
        for lvl in range(10):
            out.append((kind, 11, 300_000, lvl, False, None))
    for lvl in [1, 3, 4, 6, 7, 9]:
        out.append(("runs", 11, 200_000, lvl, False, None))
        out.append(("zeros", 11, 400_000, lvl, False, None))
    # legacy frames
    for lvl in [1, 5, 9]:
        out.append(("text", 13, 250_000, lvl, True, None))
    out.append(("zeros", 13, 300_000, 9, True, None))
    # dictionaries: short, 65535, 65536 (CLI maximum) and longer than the window
    for dspec in [["text", 21, 1000], ["text", 21, 65535], ["text", 21, 65536], ["text", 21, 70000]]:
        for lvl in [1, 3, 5, 9]:
            out.append(("text", 21, 150_000, lvl, False, dspec))
    out.append(("zeros", 5, 200_000, 9, False, ["zeros", 5, 65536]))
    out.append(("binary", 5, 200_000, 9, False, ["binary", 5, 65536]))
    # more than one block (4 MiB modern, 8 MiB legacy), cheap levels plus one -9
    out.append(("text", 31, 4 * MB + 5, 1, False, None))
    out.append(("text", 31, 4 * MB + 12, 4, False, None))
    out.append(("text", 31, 9 * MB + 777, 2, False, None))
    out.append(("text", 31, 9 * MB + 777, 6, False, None))
    out.append(("binary", 31, 5 * MB, 9, False, None))
    out.append(("zeros", 31, 9 * MB + 1, 9, False, None))
    out.append(("zeros", 31, 4 * MB + 70_000, 3, False, None))
    out.append(("random", 31, 4 * MB + 100, 9, False, None))
    out.append(("text", 31, 17 * MB, 3, True, None))
    out.append(("text", 31, 5 * MB, 5, False, ["text", 31, 65536]))
    out.append(("zeros", 31, 5 * MB, 9, False, ["zeros", 31, 30000]))
    return out


def main():
    records = []
    for kind, seed, n, lvl, legacy, dspec in cases():
        data = corpus.make(kind, n, seed).tobytes()
        frame = reference_compress(data, lvl, legacy, dict_bytes(dspec))
        rec = {"kind": kind, "seed": seed, "size": n, "level": lvl, "legacy": legacy, "dict": dspec,
               "frame_size": len(frame), "sha256": hashlib.sha256(frame).hexdigest()}
        if len(frame) <= 64:
            rec["frame_hex"] = frame.hex()
        records.append(rec)
        print(rec["kind"], n, lvl, legacy, dspec, len(frame), flush=True)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "smalLZ4 1.5 (/root/reference)",
                   "cases": records}, f, indent=0)


if __name__ == "__main__":
    main()
