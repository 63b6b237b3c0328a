#!/usr/bin/env python
"""Drive sz4_lz4 (the drop-in for smallz4::lz4) with a generator callback: the input is produced 64 KiB at a time and
the frame is hashed record by record as it arrives, so nothing of the stream's size ever sits in host memory.

    python tools/stream_check.py --mb 1024 [--stream-blocks 16] [--level 9]

Prints one JSON line: records, sha256 of every block record, peak RSS of this process, callback statistics.
Needs a GPU.  Used by tests/test_gpu_parity.py::test_lz4_streams_with_bounded_memory."""
import argparse
import hashlib
import json
import os
import resource
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from smallz4_b200 import corpus  # noqa: E402
from smallz4_b200.api import Compressor  # noqa: E402


class RecordHasher:
    """Incremental splitter of a modern LZ4 frame into [size][payload] block records."""

    def __init__(self):
        self.buf = bytearray()
        self.state = "header"
        self.digests, self.sizes, self.calls, self.first_call_at = [], [], 0, None

    def push(self, chunk):
        self.calls += 1
        if self.first_call_at is None:
            self.first_call_at = time.perf_counter()
        self.buf += chunk
        while True:
            if self.state == "header":
                if len(self.buf) < 7:
                    return
                assert bytes(self.buf[:7]) == bytes([0x04, 0x22, 0x4D, 0x18, 0x40, 0x70, 0xDF])
                del self.buf[:7]
                self.state = "records"
            elif self.state == "records":
                if len(self.buf) < 4:
                    return
                word = int.from_bytes(self.buf[:4], "little")
                if word == 0:
                    del self.buf[:4]
                    self.state = "end"
                    continue
                n = 4 + (word & 0x7FFFFFFF)
                if len(self.buf) < n:
                    return
                self.digests.append(hashlib.sha256(self.buf[:n]).hexdigest())
                self.sizes.append(n)
                del self.buf[:n]
            else:
                assert not self.buf, "bytes behind the end mark"
                return


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--stream-blocks", type=int, default=16)
    ap.add_argument("--kind", default="mixed")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    total = a.mb << 20
    c = Compressor(device=0, stream_blocks=a.stream_blocks)
    pos = [0]
    last_get = [0.0]

    def get(n):
        nonlocal total
        take = min(n, total - pos[0])
        if take <= 0:
            return b""
        chunk = corpus.make(a.kind, take, a.seed, offset=pos[0]).tobytes()
        pos[0] += take
        last_get[0] = time.perf_counter()
        return chunk

    # warm-up: a stream of three batches allocates everything the call needs (CUDA context, device arrays, the four
    # pinned halves); what the big stream adds on top of that is what depends on its length
    warm_total = min(total, 3 * a.stream_blocks * (4 << 20))
    real_total, total = total, warm_total
    c.lz4(get, lambda b: None, max_chain_length=(65535 if a.level == 9 else a.level))
    rss_warm = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    total, pos[0] = real_total, 0
    h = RecordHasher()
    t0 = time.perf_counter()
    c.lz4(get, h.push, max_chain_length=(65535 if a.level == 9 else a.level))
    dt = time.perf_counter() - t0
    assert h.state == "end"
    print(json.dumps({"bytes": total, "records": len(h.digests), "sha256": h.digests, "seconds": dt,
                      "peak_rss_kb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss, "rss_after_warmup_kb": rss_warm,
                      "send_calls": h.calls,
                      # incremental output: the first records left before the last input was pulled
                      "first_send_before_last_get": h.first_call_at is not None and h.first_call_at < last_get[0]}))


if __name__ == "__main__":
    main()
