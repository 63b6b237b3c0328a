#!/usr/bin/env python
"""Bounds check of the kernels' index arithmetic without a GPU: the SIMT-emulated build of the product sources
(tests/emu) compiled with AddressSanitizer, run over a few small inputs (blocks, chunks, batches, runs, legacy,
streaming, every walk handed to k_long) and compared with the oracle.  compute-sanitizer is closed on this GPU pool.

    python tools/emu_asan_check.py          # re-executes itself with libasan preloaded; ~4 minutes
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/libsmallz4_emu_asan.so"
FLAGS = ["-DSZ4_EMU", "-DSZ4_DP_SEG=4096", "-DSZ4_DP_WARM=512", "-DSZ4_DP_SLACK=128", "-DSZ4_DP_RING=512", "-DSZ4_GREEDY_SEG=4096",
         "-DSZ4_GREEDY_WARM=256", "-DSZ4_PATH_SEG=4096", "-DSZ4_PATH_WARM=256", "-DSZ4_LSD_CHUNK=65536"]


def main():
    if os.environ.get("SZ4_ASAN_CHILD") != "1":
        emu, csrc = os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "smallz4_b200", "csrc")
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", *FLAGS, "-I" + emu, "-I" + csrc, "-fsanitize=address", "-fno-omit-frame-pointer",
                               "-w", "-x", "c++", os.path.join(csrc, "sz4_pipeline.cu"), "-x", "c++", os.path.join(emu, "cuda_emu.cpp"),
                               "-shared", "-fPIC", "-o", SO])
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, SZ4_ASAN_CHILD="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import oracle_compress
    from smallz4_b200 import corpus
    from smallz4_b200.api import Compressor
    bs = 131072
    ok = True
    for age in (8, 0):
        c = Compressor(lib_path=SO, block_size=bs, batch_blocks=2, long_age=age)
        cases = [("mixed", 3 * bs + 777, 9, False), ("binary", 90000, 9, False), ("text", 70000, 3, False), ("runs", 150000, 9, False),
                 ("zeros", 200000, 5, False), ("mixed", bs + 5, 6, True), ("text", 13, 9, False), ("text", 0, 9, False), ("mixed", 1_150_000, 1, False)]
        for kind, n, level, legacy in cases:
            data = corpus.make(kind, n, 31).tobytes()
            same = c.compress(data, level=level, use_legacy_format=legacy) == oracle_compress(data, level, legacy, block_size=bs)[0]
            print(age, kind, n, level, legacy, "equal" if same else "DIFF", flush=True)
            ok &= same
        if age == 8:
            d = corpus.make("text", 3000, 21, offset=1 << 40).tobytes()                # the dictionary path (round-1 kernels)
            data = corpus.make("text", 50_000, 21).tobytes()
            for level in (2, 9):
                same = c.compress(data, level=level, dictionary=d) == oracle_compress(data, level, False, d, block_size=bs)[0]
                print(age, "dictionary", level, "equal" if same else "DIFF", flush=True)
                ok &= same
        data = corpus.make("mixed", 5 * bs + 99, 33).tobytes()
        pos, out = [0], []

        def get(k):
            chunk = data[pos[0]: pos[0] + min(k, 30000)]
            pos[0] += len(chunk)
            return chunk

        c.set_option("stream_blocks", 2)
        c.lz4(get, out.append, max_chain_length=65535)
        same = b"".join(out) == oracle_compress(data, 9, block_size=bs)[0]
        print(age, "stream", "equal" if same else "DIFF", flush=True)
        ok &= same
        c.close()
    print("ALL OK, no AddressSanitizer report" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
