"""Sharding of one stream over several GPUs (one process per GPU).

Blocks of the LZ4 frame depend only on the 64 KiB in front of them (DESIGN.md "Independence of
blocks"), so every rank takes a contiguous range of whole blocks plus a read-only halo, compresses it
with sz4_compress_device and hands back the concatenated [size][payload] block records.  Rank 0
puts header + records in rank order + end mark together.  No collective touches the data path;
torch.distributed is only used to move the finished records to rank 0.
"""
import numpy as np

BLOCK = 4 << 20          # smallz4.h:124 MaxBlockSize
BLOCK_LEGACY = 8 << 20   # smallz4.h:127
HALO = 131072            # >= 65535 + 12 bytes of history, multiple of 16


def plan(total_bytes, world, block=BLOCK):
    """[(begin, end)] per rank: contiguous ranges of whole blocks (the last range may end mid-block)."""
    nblocks = (total_bytes + block - 1) // block
    per = (nblocks + world - 1) // world if world else 0
    out = []
    for r in range(world):
        b = min(r * per, nblocks) * block
        e = min(min((r + 1) * per, nblocks) * block, total_bytes)
        out.append((min(b, total_bytes), e))
    return out


def halo_for(begin, legacy=False):
    """History a rank needs in front of its first block."""
    return 0 if legacy else min(begin, HALO)


def compress_shard(comp, ptr, halo, nbytes, out_ptr, out_cap, level, first, last, legacy=False):
    """ptr -> `halo` bytes of history then `nbytes` of blocks (device memory).  Returns the record length."""
    if nbytes == 0:
        return 0
    return comp.compress_device(ptr, halo, nbytes, out_ptr, out_cap, level=level, first=first, last=last,
                                use_legacy_format=legacy)


def frame_header(legacy=False):
    return bytes([0x02, 0x21, 0x4C, 0x18]) if legacy else bytes([0x04, 0x22, 0x4D, 0x18, 1 << 6, 7 << 4, 0xDF])


def frame_end(legacy=False):
    return b"" if legacy else bytes(4)


def gather_frame(records: bytes, dist, legacy=False):
    """Collect every rank's block records on rank 0 and return the complete frame there (None elsewhere)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(records)], dtype=torch.int64))
    biggest = max(int(s.item()) for s in sizes)
    mine = torch.zeros(max(biggest, 1), dtype=torch.uint8)
    if records:
        mine[: len(records)] = torch.from_numpy(np.frombuffer(records, dtype=np.uint8).copy())
    parts = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank != 0:
        return None
    body = b"".join(parts[r][: int(sizes[r].item())].numpy().tobytes() for r in range(world))
    return frame_header(legacy) + body + frame_end(legacy)
