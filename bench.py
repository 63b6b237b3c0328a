#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: input GB/s at level -9 on the synthetic mixed corpus.

A "step" is one pass of the hot path (smallz4 -9, 4 MiB blocks) over the whole workload.

  python bench.py                       N=1: BASELINE configs[1], 256 MB of the mixed corpus on one B200
  torchrun ... bench.py --gpus N        BASELINE configs[3]: a fixed 8 GB corpus cut by smallz4_b200/shard.py into N
                                        contiguous ranges of whole blocks, each with its 128 KiB halo (strong scaling,
                                        no collective on the data path -- blocks only depend on their halo)

  value         whole-job input GB/s with the input already resident in HBM (sz4_compress_device)
  e2e           the same through the host API (sz4_compress_host): pinned host input -> H2D -> kernels -> D2H of the
                frame into pinned host memory, all inside the timed region
  parity        after the timed region the block records are hashed and compared with digests of the UNMODIFIED
                reference (tests/golden/golden_blocks.json: the first 256 MB of the corpus, all 64 blocks)
  roofline      the dominant kernel (longest phase, measured live with CUDA events on the library's stream) against
                the measured HBM copy bandwidth in MEASURED_PEAKS.json, with what ncu says bounds it
  cpu_baseline  the unmodified reference (oracle/_ref) on one host core, on contiguous 1 MiB pieces of the workload

`--impl reference` times the reference's own CPU code on all host cores (one process per core, one contiguous
1 MiB piece of the same corpus each per step -- BASELINE configs[0] is "1 MB") with the same config keys.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

MB = 1 << 20
BLOCK = 4 * MB
HALO = 131072
KIND, SEED = "mixed", 1
METRIC = "input GB/s at -9 (smallz4 optimal parse, 4 MiB blocks), byte-identical to the reference"
GOLDEN_BLOCKS = os.path.join(ROOT, "tests", "golden", "golden_blocks.json")
SELF_DIGESTS = os.path.join(ROOT, "tests", "golden", "blocks_8gb_selfcheck.json")
COUNTERS = os.path.join(ROOT, "profiles", "r2_counters.json")


def workload_name(total_mb):
    size = f"{total_mb // 1024} GB" if total_mb >= 1024 and total_mb % 1024 == 0 else f"{total_mb} MB"
    return f"{size} synthetic mixed corpus (text/binary/runs/random/zeros) at -9, maxChainLength 65535, 4 MiB blocks"


def fill_parallel(out, kind, seed, offset, threads=8):
    """corpus.fill over several threads (the generator is a pure function of the offset; ctypes drops the GIL)."""
    from smallz4_b200 import corpus
    n = out.size
    step = max(((n + threads - 1) // threads + 65535) // 65536 * 65536, 65536)
    ts = [threading.Thread(target=corpus.fill, args=(out[a:min(a + step, n)], kind, seed, offset + a)) for a in range(0, n, step)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


# --------------------------------------------------------------------------- reference arm (CPU)
def _ref_worker(args):
    """One process per core: compress one contiguous piece with the unmodified reference (or the oracle port)."""
    offset, nbytes, level, kind = args
    from oracle_lib import oracle_compress, reference, reference_compress
    from smallz4_b200 import corpus
    data = corpus.make(KIND, nbytes, SEED, offset=offset)
    t = time.perf_counter()
    if kind == "reference" and reference() is not None:
        reference_compress(data, level)
    else:
        oracle_compress(data, level)
    return time.perf_counter() - t


def reference_kind():
    from oracle_lib import reference
    return "reference" if reference() is not None else "port"


def piece_offsets(total_bytes, piece, count, rotate=0):
    """`count` piece-aligned offsets spread evenly over the workload; `rotate` shifts them so that successive steps
    sample different stretches."""
    slots = max(total_bytes // piece, 1)
    return [(((2 * i + 1) * slots) // (2 * count) + rotate * 7) % slots * piece for i in range(count)]


def _piece_weight(args):
    """Cheap estimate of what a piece costs the reference (it is quadratic in runs of one byte): share of repeated bytes."""
    offset, nbytes = args
    from smallz4_b200 import corpus
    d = corpus.make(KIND, nbytes, SEED, offset=offset)
    return float(np.count_nonzero(d[1:] == d[:-1])) / max(nbytes, 1)


def run_reference(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    piece = args.ref_piece_kb * 1024
    kind = reference_kind()
    total_mb = args.size_mb
    total = total_mb * MB
    per_step = cores * args.ref_pieces_per_core
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        # warm-up steps page the library and the corpus generator in; they use 64 KiB pieces so that the whole
        # run stays within minutes (a 1 MiB piece takes the reference 5-30 s at -9)
        for w in range(args.warmup):
            pool.map(_ref_worker, [(o, 65536, 9, kind) for o in piece_offsets(total, 65536, cores, w)], chunksize=1)
        # the pieces of every step, dearest first and handed out one by one: no core idles behind a slow piece for long
        plans = []
        for s in range(args.steps):
            offs = piece_offsets(total, piece, per_step, s)
            weights = pool.map(_piece_weight, [(o, piece) for o in offs], chunksize=1)
            plans.append([o for _, o in sorted(zip(weights, offs), reverse=True)])
        t0 = time.perf_counter()
        busy = 0.0
        for s in range(args.steps):
            busy += sum(pool.imap_unordered(_ref_worker, [(o, piece, 9, kind) for o in plans[s]], chunksize=1))
        dt = time.perf_counter() - t0
    gbs = per_step * piece * args.steps / dt / 1e9
    sample = (f"{per_step} contiguous pieces of {args.ref_piece_kb} KiB per step ({args.ref_pieces_per_core} per core, {cores} processes, "
              f"handed out one by one, dearest first), spread evenly over the workload and moved every step; each piece is compressed as "
              f"its own stream; cores busy {100 * busy / (dt * cores):.0f} % of the time")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(total_mb), "level": 9, "block_bytes": BLOCK, "sampled": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) > 2 + k and r[2 + k].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------- our arm (GPU)
# Algorithmic bytes per input position of each phase (DESIGN.md "Kernels"): what the phase has to read and write once if
# nothing were re-read.  sort: the data once for the histogram; pass 1 reads the data and writes key + position (12 B);
# passes 2-4 move 12 B in and out, pass 5 12 in / 20 out (the carried tables start), passes 6-8 20 in and out.
PHASE_BYTES = {"sort": 1 + (1 + 12) + 3 * 24 + (12 + 20) + 3 * 40,
               "chain": (20 + 8 + 2 + 4) + (1 + 4 + 2) + 2 + (8 + 2 + 1 + 4 + 4 + 2 + 2),      # extract, run helpers, tile cost, k_start
               "search": 1 + 2 + 2 + 6, "fixup": 6, "dp": 4 + 2 + 4 + 16, "path": 6, "emit": 2}
PHASE_KERNEL = {"sort": "k_lsd_pass2 x8 (+ k_lsd_hist, k_lsd_bases)",
                "chain": "k_lsd_extract, k_run_*, k_tile_cost/order, k_start",
                "search": "k_search (+ k_long)", "fixup": "k_seed_detect/k_seed_fix", "dp": "k_dp_spec (+ plan, verify)",
                "path": "k_path_*", "emit": "k_emit"}
# what ncu says limits the phase's main kernel (profiles/r2_summary.md has the counters)
PHASE_LIMIT = {"sort": "HBM copies with a permutation; barrier and shared-memory stalls between the phases of a tile keep them at about "
                       "half of the copy bandwidth",
               "chain": "L2 latency (gathers inside the 64 KiB window) and scatters that meet in L2",
               "search": "issue slots and shared-memory wavefronts of the chain walk (k_search); L2 latency (k_long): not HBM",
               "fixup": "latency", "dp": "issue slots / dependent-issue latency of the cost recurrence: not HBM", "path": "latency",
               "emit": "HBM"}
STEP_BYTES = sum(PHASE_BYTES.values())


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_counters():
    """Counters of the committed ncu --set full capture (profiles/r2_counters.json, written by tools/make_profile_summary.py)."""
    if os.path.exists(COUNTERS):
        with open(COUNTERS) as f:
            return json.load(f)
    return {}


def cpu_baseline(size_mb, pieces, piece_kb):
    """Single host core, unmodified reference when available, on contiguous pieces of the workload."""
    from oracle_lib import oracle_compress, reference_compress
    from smallz4_b200 import corpus
    kind = reference_kind()
    piece = piece_kb * 1024
    t = 0.0
    for off in piece_offsets(size_mb * MB, piece, pieces):
        data = corpus.make(KIND, piece, SEED, offset=off)
        t0 = time.perf_counter()
        if kind == "reference":
            reference_compress(data, 9)
        else:
            oracle_compress(data, 9)
        t += time.perf_counter() - t0
    return {"value": pieces * piece / t / 1e9, "unit": "GB/s", "cores": 1, "kind": kind,
            "sample": f"{pieces} contiguous pieces of {piece_kb} KiB spread evenly over the workload, each its own stream "
                      f"({t:.1f} s of CPU)"}


def split_records(body):
    """[bytes]: the [size][payload] block records in a rank's output (smallz4.h:769-780)."""
    out, at, n = [], 0, len(body)
    while at < n:
        word = int.from_bytes(body[at:at + 4], "little")
        size = word & 0x7FFFFFFF
        out.append(body[at:at + 4 + size])
        at += 4 + size
    assert at == n, "block records do not tile the segment"
    return out


def check_parity(digests, total_bytes):
    """digests: sha256 hex of every block record of the whole job, in stream order."""
    res = {"checked": False, "blocks": 0, "mismatches": 0, "against": None}
    if os.path.exists(GOLDEN_BLOCKS):
        with open(GOLDEN_BLOCKS) as f:
            gb = json.load(f)
        gold = gb["levels"].get("9")
        if gold and gb["kind"] == KIND and gb["seed"] == SEED:
            n = min(len(gold), len(digests), total_bytes // BLOCK)
            res.update(checked=n > 0, blocks=n, against="tests/golden/golden_blocks.json (unmodified reference, level 9)",
                       mismatches=sum(1 for k in range(n) if gold[k]["sha256"] != digests[k]))
    if len(digests) > res["blocks"] and os.path.exists(SELF_DIGESTS):
        with open(SELF_DIGESTS) as f:
            sd = json.load(f)
        if sd.get("total_bytes") == total_bytes and len(sd["sha256_16"]) == len(digests):
            res["self_blocks"] = len(digests)
            res["self_mismatches"] = sum(1 for a, b in zip(sd["sha256_16"], digests) if a != b[:16])
            res["self_against"] = "tests/golden/blocks_8gb_selfcheck.json (this library at N=1, committed)"
    return res


def run_ours(args, rank, world, local_rank, sharded):
    import torch
    from smallz4_b200 import corpus, shard
    from smallz4_b200.api import Compressor

    dist = None
    torch.cuda.set_device(local_rank)
    if "RANK" in os.environ and "MASTER_ADDR" in os.environ:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    total_mb = args.total_gb * 1024 if sharded else args.size_mb
    total = total_mb * MB
    begin, end = shard.plan(total, world)[rank]
    mine = end - begin
    halo = shard.halo_for(begin)
    host_in = torch.empty(halo + mine, dtype=torch.uint8).pin_memory()
    fill_parallel(host_in.numpy(), KIND, SEED, begin - halo)
    d_in = host_in.to(dev, non_blocking=False)
    cap = mine + 4 * (mine // BLOCK + 2) + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    host_out = torch.empty(cap + 64, dtype=torch.uint8).pin_memory()

    c = Compressor(device=local_rank, profile=1, batch_blocks=args.batch_blocks)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        if mine == 0:
            return 0
        return c.compress_device(d_in.data_ptr(), halo, mine, d_out.data_ptr(), cap, level=9, first=(begin == 0),
                                 last=(end == total))

    def host_step():
        # the public host entry points: the whole stream at N=1, the rank's own range (history + whole blocks) when sharded
        if mine == 0:
            return 0
        if not sharded:
            return c.compress_into(host_in.data_ptr(), mine, host_out.data_ptr(), cap + 64, level=9)
        return c.compress_range_into(host_in.data_ptr(), halo, mine, host_out.data_ptr(), cap + 64, level=9,
                                     first=(begin == 0), last=(end == total))

    # ---- device-resident timing
    seg_len = 0
    for _ in range(args.warmup):
        seg_len = device_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    kernel_ms, launches = 0.0, 0
    phases = {}
    for _ in range(args.steps):
        seg_len = device_step()
        ms, ln = c.last_stats()
        kernel_ms += ms
        launches += ln
        for k, v in c.last_phase_ms().items():
            phases[k] = phases.get(k, 0.0) + v
    barrier()
    dt = time.perf_counter() - t0
    # ---- end-to-end timing (host buffers, copies inside)
    host_step()
    barrier()
    t1 = time.perf_counter()
    frame_len = 0
    for _ in range(args.steps):
        frame_len = host_step()
    barrier()
    dt_e2e = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join()

    # ---- parity: hash what the timed device steps produced
    body = d_out[:seg_len].cpu().numpy().tobytes() if seg_len else b""
    my_digests = [hashlib.sha256(r).hexdigest() for r in split_records(body)]
    per_rank = {"rank": rank, "bytes": mine, "kernel_ms_per_step": kernel_ms / args.steps, "blocks": len(my_digests),
                "records_sha256": hashlib.sha256("".join(my_digests).encode()).hexdigest()[:16]}
    if dist is not None:
        t = torch.tensor([dt, dt_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = t.tolist()
        gathered = [None] * world
        dist.all_gather_object(gathered, (per_rank, my_digests, int(frame_len), int(seg_len)))
        ranks = [g[0] for g in gathered]
        digests = [d for g in gathered for d in g[1]]
        d2h = sum(g[2] for g in gathered)
        packed = sum(g[3] for g in gathered)
    else:
        ranks, digests, d2h, packed = [per_rank], my_digests, int(frame_len), int(seg_len)

    if rank == 0:
        peak, peak_kind = measured_peak()
        top = max(phases, key=phases.get)
        top_ms = phases[top] / args.steps
        achieved = PHASE_BYTES[top] * mine / (top_ms * 1e-3) / 1e9
        counters = ncu_counters().get(top, {})
        step_ms = dt / args.steps * 1e3
        parity = check_parity(digests, total)
        if args.write_digests:
            os.makedirs(os.path.dirname(args.write_digests), exist_ok=True)
            with open(args.write_digests, "w") as f:
                json.dump({"generator": "bench.py --write-digests (this library, not the reference)", "kind": KIND, "seed": SEED,
                           "total_bytes": total, "n_gpus": world, "sha256_16": [d[:16] for d in digests]}, f)
        line = {
            "metric": METRIC, "value": total * args.steps / dt / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(total_mb), "level": 9, "block_bytes": BLOCK,
                       "sharding": (f"{world} contiguous ranges of whole blocks, one per GPU, each with a 128 KiB halo in front; no collective "
                                    f"on the data path" if world > 1 else "one GPU, whole stream"),
                       "l2": f"every batch ({min(args.batch_blocks * 4, total_mb)} MB of input, ~75 B of arrays per byte) is larger than "
                             f"L2 (126 MB); no flush needed",
                       "batch_blocks": args.batch_blocks},
            "e2e": {"value": total * args.steps / dt_e2e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": total + (world - 1) * HALO,
                    "d2h_bytes_per_step": d2h + 64 * (total // (args.batch_blocks * BLOCK) + world)},
            "gpu_launches": int(launches),
            "parity": parity,
            "kernel_ms_per_step": kernel_ms / args.steps,
            "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "ranks": ranks,
            "compression_ratio": total / max(packed, 1),
            "roofline": {"bound": "hbm", "limited_by": PHASE_LIMIT[top], "kernel": PHASE_KERNEL[top],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (int(counters["dram_bytes_per_input_byte"] * mine) if "dram_bytes_per_input_byte" in counters else None),
                         "traffic_source": counters.get("capture"),
                         "peak_source": peak_kind, "algorithmic_bytes_per_input_byte": PHASE_BYTES[top], "ms_per_launch": top_ms,
                         "issue_frac": counters.get("issue_frac"), "smem_wavefront_frac": counters.get("smem_wavefront_frac"),
                         "active_lanes": counters.get("active_lanes"), "dram_frac": counters.get("dram_frac"),
                         "phases": {k: {"ms": phases[k] / args.steps, "algorithmic_bytes_per_input_byte": PHASE_BYTES[k],
                                        "achieved": PHASE_BYTES[k] * mine / max(phases[k] / args.steps * 1e-3, 1e-9) / 1e9,
                                        "frac": PHASE_BYTES[k] * mine / max(phases[k] / args.steps * 1e-3, 1e-9) / 1e9 / peak,
                                        "kernels": PHASE_KERNEL[k], "limited_by": PHASE_LIMIT[k],
                                        **{c: v for c, v in ncu_counters().get(k, {}).items() if c != "capture"}}
                                    for k in phases},
                         "step": {"algorithmic_bytes_per_input_byte": STEP_BYTES,
                                  "achieved": STEP_BYTES * mine / (kernel_ms / args.steps * 1e-3) / 1e9,
                                  "frac": STEP_BYTES * mine / (kernel_ms / args.steps * 1e-3) / 1e9 / peak}},
            "clocks": sampler.summary(),
        }
        if world == 1 and not sharded and not args.no_extra:
            line["extra"] = extra_runs(c, torch, dev, args)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(total_mb, args.cpu_pieces, args.ref_piece_kb)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def extra_runs(c, torch, dev, args):
    """Device-resident GB/s (one warm-up + two timed steps each, outside the headline timing): the four corpora of
    BASELINE.json's north_star plus zeros at -9, and levels 1..8 on the mixed corpus."""
    from smallz4_b200 import corpus
    n = args.size_mb * MB
    cap = n + 4 * (n // BLOCK + 2) + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    out = {"size_mb": args.size_mb, "unit": "GB/s (device-resident)", "corpora_level9": {}, "levels_mixed": {}}

    def timed(d_in, level):
        c.compress_device(d_in.data_ptr(), 0, n, d_out.data_ptr(), cap, level=level)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(2):
            c.compress_device(d_in.data_ptr(), 0, n, d_out.data_ptr(), cap, level=level)
        torch.cuda.synchronize()
        return round(2 * n / (time.perf_counter() - t) / 1e9, 3)

    host = np.empty(n, dtype=np.uint8)
    for kind in ["text", "binary", "runs", "random", "zeros"]:
        fill_parallel(host, kind, SEED, 0)
        d_in = torch.from_numpy(host).to(dev)
        out["corpora_level9"][kind] = timed(d_in, 9)
    fill_parallel(host, KIND, SEED, 0)
    d_in = torch.from_numpy(host).to(dev)
    for level in range(1, 9):
        out["levels_mixed"][str(level)] = timed(d_in, level)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mb", type=int, default=256, help="N=1 without torchrun: BASELINE configs[1]")
    ap.add_argument("--total-gb", type=int, default=8, help="under torchrun: BASELINE configs[3], the whole job's input")
    ap.add_argument("--sharded", action="store_true", help="use the 8 GB sharded configuration even without torchrun")
    ap.add_argument("--batch-blocks", type=int, default=64)
    ap.add_argument("--cpu-pieces", type=int, default=3)
    ap.add_argument("--ref-piece-kb", type=int, default=1024)
    ap.add_argument("--ref-pieces-per-core", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--write-digests", default=None, help="write the per-block digests of this run to a JSON file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sharded = args.sharded or world > 1
    if args.impl == "reference":
        if sharded:
            args.size_mb = args.total_gb * 1024
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank, sharded)


if __name__ == "__main__":
    main()
