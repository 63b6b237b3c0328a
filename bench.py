#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: input GB/s at level -9 on the synthetic mixed corpus.

A "step" is one pass of the hot path (smallz4 -9, 4 MiB blocks) over one rank's shard.  At N=1 the
workload is BASELINE configs[1]: 256 MB of the mixed corpus.  At N>1 every rank owns its own 256 MB
shard of an N x 256 MB corpus plus the 64 KiB halo in front of it (weak scaling, no collective on
the data path -- blocks only depend on their halo).

  value   whole-job input GB/s with the shard already resident in HBM (sz4_compress_device)
  e2e     the same through the host API (sz4_compress_host): pinned host input -> H2D -> kernels ->
          D2H of the frame into pinned host memory, all inside the timed region
  roofline   the dominant kernel (longest phase, measured live with CUDA events on the library's
          stream) against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference (oracle/_ref) on one host core, on a bounded sample

`--impl reference` times the reference's own CPU code on all host cores (one process per core over
file shards, as BASELINE.json's north_star describes) on bounded samples of the same workload.
"""
import argparse
import ctypes
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
HALO = 131072
METRIC = "input GB/s at -9 (smallz4 optimal parse, 4 MiB blocks), byte-identical to the reference"
SEED = 1


def workload_name(mb):
    return f"{mb} MB synthetic mixed corpus (text/binary/runs/random/zeros) at -9, maxChainLength 65535, 4 MiB blocks"


# --------------------------------------------------------------------------- reference arm (CPU)
def _ref_worker(args):
    """One process per core: compress one shard with the unmodified reference (or the oracle port)."""
    offset, nbytes, level, kind = args
    from oracle_lib import oracle_compress, reference, reference_compress
    from smallz4_b200 import corpus
    data = corpus.make("mixed", nbytes, SEED, offset=offset)
    t = time.perf_counter()
    if kind == "reference" and reference() is not None:
        reference_compress(data, level)
    else:
        oracle_compress(data, level)
    return time.perf_counter() - t


def reference_kind():
    from oracle_lib import reference
    return "reference" if reference() is not None else "port"


def run_reference(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    shard = args.ref_shard_kb * 1024
    kind = reference_kind()
    total_mb = args.size_mb
    # many small shards spread evenly over the whole workload (every corpus component is sampled in
    # proportion) and handed out dynamically, so that no core idles behind a slow shard
    pieces = cores * args.ref_shards_per_core
    stride = max(shard, (total_mb * MB // pieces) // shard * shard)
    jobs = [(i * stride, shard, 9, kind) for i in range(pieces)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    bytes_per_step = pieces * shard
    gbs = bytes_per_step * args.steps / dt / 1e9
    sample = f"{pieces} shards of {args.ref_shard_kb} KiB spread evenly over the workload, {cores} processes (one per core), per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(total_mb), "sampled": sample},
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
            time.sleep(0.2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) > 2 + k and r[2 + k].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------- our arm (GPU)
# algorithmic bytes per input position of each phase's dominant kernel (DESIGN.md "Kernels")
PHASE_BYTES = {"sort": 3 * (8 + 8 + 8) + 1, "chain": 8 + 2 + 1 + 4 + 2 + 2, "search": 1 + 2 + 4 + 2, "fixup": 6,
               "dp": 4 + 2 + 4 + 16, "path": 6, "emit": 2}
# DRAM bytes per input byte of the dominant kernels, from ncu --set full (profiles/r1_final_summary.md)
DRAM_BYTES_PER_INPUT_BYTE = {"search": 16.1}
PHASE_KERNEL = {"sort": "k_sort_scatter (+hist, scan)", "chain": "k_chain (+ run helpers k_flag_*)", "search": "k_search",
                "fixup": "k_seed_detect/k_seed_fix", "dp": "k_dp", "path": "k_path", "emit": "k_emit"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def cpu_baseline(size_mb, sample_kb):
    """Single host core, unmodified reference when available, on a bounded sample of the workload."""
    from oracle_lib import oracle_compress, reference, reference_compress
    from smallz4_b200 import corpus
    kind = reference_kind()
    pieces = 32           # many small pieces: the corpus components differ by 100x in reference speed
    piece = sample_kb * 1024 // pieces
    stride = (size_mb * MB // pieces) // piece * piece
    t = 0.0
    for i in range(pieces):
        data = corpus.make("mixed", piece, SEED, offset=i * stride)
        t0 = time.perf_counter()
        if kind == "reference":
            reference_compress(data, 9)
        else:
            oracle_compress(data, 9)
        t += time.perf_counter() - t0
    return {"value": pieces * piece / t / 1e9, "unit": "GB/s", "cores": 1, "kind": kind,
            "sample": f"{pieces} pieces of {piece // 1024} KiB spread evenly over the workload ({t:.1f} s of CPU)"}


def run_ours(args, rank, world, local_rank):
    import torch
    from smallz4_b200 import corpus
    from smallz4_b200.api import Compressor

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    shard = args.size_mb * MB
    halo = HALO if rank > 0 else 0
    host_in = torch.empty(halo + shard, dtype=torch.uint8).pin_memory()
    corpus.fill(host_in.numpy(), "mixed", SEED, offset=rank * shard - halo)
    d_in = host_in.to(dev, non_blocking=False)
    cap = shard + shard // 255 + 4 * (shard // (4 * MB) + 2) + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    host_out = torch.empty(cap + 64, dtype=torch.uint8).pin_memory()

    c = Compressor(device=local_rank, profile=1, batch_blocks=args.batch_blocks)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return c.compress_device(d_in.data_ptr(), halo, shard, d_out.data_ptr(), cap, level=9, first=(rank == 0),
                                 last=(rank == world - 1))

    def host_step():
        return c.compress_into(host_in.data_ptr() + halo, shard, host_out.data_ptr(), cap + 64, level=9)

    # ---- device-resident timing
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
    for _ in range(args.steps):
        frame_len = host_step()
    barrier()
    dt_e2e = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join()

    if dist is not None:
        t = torch.tensor([dt, dt_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = t.tolist()
    if rank == 0:
        total = shard * world
        peak, peak_kind = measured_peak()
        top = max(phases, key=phases.get)
        top_ms = phases[top] / args.steps
        achieved = PHASE_BYTES[top] * shard / (top_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total * args.steps / dt / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args.size_mb) + (f" per GPU, {world} shards with 128 KiB halos" if world > 1 else ""),
                       "level": 9, "block_bytes": 4 * MB, "l2": (f"input ({args.size_mb} MB) larger than L2 (126 MB); no flush needed" if args.size_mb > 126
                              else f"input ({args.size_mb} MB) fits L2: not a valid bench size"),
                       "batch_blocks": args.batch_blocks},
            "e2e": {"value": total * args.steps / dt_e2e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": shard,
                    "d2h_bytes_per_step": int(frame_len) + 8 * ((shard // (args.batch_blocks * 4 * MB)) + 1)},
            "gpu_launches": int(launches),
            "kernel_ms_per_step": kernel_ms / args.steps,
            "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "compression_ratio": shard / max(int(seg_len), 1),
            "roofline": {"bound": "hbm", "kernel": PHASE_KERNEL[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (int(DRAM_BYTES_PER_INPUT_BYTE[top] * shard) if top in DRAM_BYTES_PER_INPUT_BYTE else None),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture at 64 MB "
                                           "(profiles/r1_final_summary.md), scaled to this launch's input bytes",
                         "peak_source": peak_kind,
                         "algorithmic_bytes_per_input_byte": PHASE_BYTES[top], "ms_per_launch": top_ms},
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.size_mb, args.cpu_sample_kb)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mb", type=int, default=256)
    ap.add_argument("--batch-blocks", type=int, default=64)
    ap.add_argument("--cpu-sample-kb", type=int, default=4096)
    ap.add_argument("--ref-shard-kb", type=int, default=64)
    ap.add_argument("--ref-shards-per-core", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
