"""N>1 path on the CPU: two processes (gloo) each compress their range of whole blocks plus halo and
rank 0 assembles the frame; it must equal the single-stream frame.  The compressor behind the
C ABI is the emulated build here (host logic under test: smallz4_b200/shard.py); on the GPU box
bench.py drives the same code with the CUDA library."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BS = 131072


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, level, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu_lib import emu_compressor
    from smallz4_b200 import corpus, shard

    ranges = shard.plan(total, world, block=BS)
    begin, end = ranges[rank]
    halo = shard.halo_for(begin)
    buf = corpus.make("mixed", halo + (end - begin), seed=4, offset=begin - halo)      # shard + halo, generated locally
    out = np.empty(2 * (end - begin) + 4096, dtype=np.uint8)
    comp = emu_compressor(block_size=BS, batch_blocks=2)
    n = shard.compress_shard(comp, buf.ctypes.data, halo, end - begin, out.ctypes.data, out.size, level,
                             first=(begin == 0), last=(end == total))
    frame = shard.gather_frame(out[:n].tobytes(), dist)
    if rank == 0:
        with open(out_path, "wb") as f:
            f.write(frame)
    dist.destroy_process_group()


@pytest.mark.parametrize("level", [3, 9])
def test_two_ranks_equal_one_stream(tmp_path, level):
    from emu_lib import build_emu
    from oracle_lib import oracle_compress
    from smallz4_b200 import corpus
    build_emu()
    total = 5 * BS + 1234
    out_path = str(tmp_path / "frame.lz4")
    mp.spawn(_worker, args=(2, _free_port(), total, level, out_path), nprocs=2, join=True)
    want, _ = oracle_compress(corpus.make("mixed", total, seed=4).tobytes(), level, block_size=BS)
    assert open(out_path, "rb").read() == want


def test_plan_covers_the_stream_with_whole_blocks():
    from smallz4_b200 import shard
    for total in [0, 1, BS, BS + 1, 7 * BS + 5]:
        for world in [1, 2, 3, 8]:
            r = shard.plan(total, world, block=BS)
            assert r[0][0] == 0 and r[-1][1] == total
            for (b0, e0), (b1, e1) in zip(r, r[1:]):
                assert e0 == b1
            for b, e in r:
                if e > b:                                   # ranks beyond the last block get an empty range
                    assert b % BS == 0 and (e % BS == 0 or e == total)
