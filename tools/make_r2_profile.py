#!/usr/bin/env python
"""Builds profiles/r2_summary.md and profiles/r2_counters.json from what the final gpurun call of round 2 left in
gpurun_out/ (bench lines, ncu launch list, ncu --set full report).  Needs `ncu` on PATH to read the report (no GPU).

    python tools/make_r2_profile.py [--tag r2_final]

r2_counters.json is what bench.py quotes next to its live timings (`roofline.issue_frac` etc.): ncu counters of the main
kernel of each phase as fractions of their peaks, with the capture they come from."""
import argparse
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
SMS = 148
INPUT_BYTES = 256 << 20
PHASE_OF = {"k_lsd_pass2": "sort", "k_lsd_hist": "sort", "k_lsd_extract": "chain", "k_run_apply": "chain", "k_start": "chain",
            "k_search": "search", "k_long": "search", "k_dp_spec": "dp"}
MAIN = {"sort": "k_lsd_pass2<3", "chain": "k_start", "search": "k_search", "dp": "k_dp_spec"}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def short(name):
    n = name.split("(")[0].replace("void ", "").strip()
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2_final")
    a = ap.parse_args()
    tag = a.tag
    os.makedirs(P, exist_ok=True)
    for f in (f"{tag}_bench.json", f"{tag}_bench_reference.json", f"{tag}_launches.csv"):
        if os.path.exists(os.path.join(G, f)):
            shutil.copy(os.path.join(G, f), os.path.join(P, f))
    # ---- launch list
    rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 10]
    h = rows[0]
    t = collections.defaultdict(float)
    c = collections.Counter()
    for r in rows[1:]:
        d = dict(zip(h, r))
        n = short(d["Kernel Name"])
        t[n] += num(d["Metric Value"]) / 1e6
        c[n] += 1
    tot = sum(t.values())
    table = "\n".join(f"| `{k}` | {c[k]} | {v / c[k]:.3f} | {100 * v / tot:.1f} % |" for k, v in sorted(t.items(), key=lambda x: -x[1])[:24])
    # ---- full counters
    raw = subprocess.run(["ncu", "-i", os.path.join(G, f"{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    per_kernel = collections.OrderedDict()
    for vals in rr[2:]:
        d = dict(zip(hdr, vals))
        name = short(d["Kernel Name"])
        cyc = num(d["sm__cycles_elapsed.avg"])
        wf = num(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"])
        bc = num(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"])
        st = sorted(((k.split("issue_stalled_")[1].split("_per")[0], num(d[k])) for k in hdr
                     if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), key=lambda x: -x[1])[:5]
        e = {"ms": num(d["gpu__time_duration.sum"]) / (1e6 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1),
             "registers": num(d["launch__registers_per_thread"]), "warps_active_frac": num(d["sm__warps_active.avg.pct_of_peak_sustained_active"]) / 100,
             "warp_inst": num(d["smsp__inst_executed.sum"]), "active_lanes": num(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
             "issue_frac": num(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / 100,
             "smem_wavefront_frac": wf / (SMS * cyc) if cyc else 0.0, "smem_bank_conflict_share": bc / wf if wf else 0.0,
             "dram_read_gb": num(d["dram__bytes_read.sum"]), "dram_write_gb": num(d["dram__bytes_write.sum"]),
             "dram_frac": num(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]) / 100,
             "l2_frac": num(d["lts__throughput.avg.pct_of_peak_sustained_elapsed"]) / 100,
             "stalls": ", ".join(f"{s} {v:.2f}" for s, v in st)}
        for k in ("dram_read_gb", "dram_write_gb"):
            u = units[hdr.index("dram__bytes_read.sum" if k == "dram_read_gb" else "dram__bytes_write.sum")]
            e[k] *= {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1.0)
        per_kernel.setdefault(name, []).append(e)
    counters = {}
    lines = []
    for name, es in per_kernel.items():
        e = max(es, key=lambda x: x["ms"])
        lines.append(f"### `{name}` ({len(es)} launch{'es' if len(es) > 1 else ''} captured; the longest)\n"
                     f"- {e['ms']:.3f} ms, {e['registers']:.0f} registers/thread, warps active {100 * e['warps_active_frac']:.0f} % of peak\n"
                     f"- {e['warp_inst'] / 1e9:.2f} G warp instructions, {e['active_lanes']:.1f} of 32 lanes per instruction, "
                     f"**issue slots {100 * e['issue_frac']:.0f} % of peak**\n"
                     f"- **shared memory {e['smem_wavefront_frac']:.3f} wavefronts per SM-cycle (peak 1.0)**, "
                     f"{100 * e['smem_bank_conflict_share']:.0f} % of them bank-conflict replays\n"
                     f"- DRAM {e['dram_read_gb']:.2f} GB read + {e['dram_write_gb']:.2f} GB written = **{100 * e['dram_frac']:.0f} % of peak**; "
                     f"L2 {100 * e['l2_frac']:.0f} % of peak\n"
                     f"- stalls per issue: {e['stalls']}\n")
    for phase, kern in MAIN.items():
        match = [k for k in per_kernel if k.startswith(kern)]
        if not match:
            continue
        es = [x for k in match for x in per_kernel[k]]
        e = max(es, key=lambda x: x["ms"])
        # DRAM bytes of every captured launch that belongs to the phase, per input byte
        phase_bytes = sum((x["dram_read_gb"] + x["dram_write_gb"]) * 1e9 for k, v in per_kernel.items()
                          for x in v if PHASE_OF.get(k.split("<")[0], None) == phase)
        counters[phase] = {"kernel": match[0], "issue_frac": round(e["issue_frac"], 4), "smem_wavefront_frac": round(e["smem_wavefront_frac"], 4),
                           "active_lanes": round(e["active_lanes"], 2), "dram_frac": round(e["dram_frac"], 4),
                           "dram_bytes_per_input_byte": round(phase_bytes / INPUT_BYTES, 2),
                           "capture": f"profiles/{tag}.ncu-rep summary in profiles/r2_summary.md: ncu --set full --clock-control none of "
                                      f"`python bench.py --steps 1 --warmup 1` (256 MB mixed corpus, first batch), dram__bytes_read.sum + "
                                      f"dram__bytes_write.sum of the phase's captured launches"}
    with open(os.path.join(P, "r2_counters.json"), "w") as f:
        json.dump(counters, f, indent=1)
    b = json.load(open(os.path.join(P, f"{tag}_bench.json")))
    ref = json.load(open(os.path.join(P, f"{tag}_bench_reference.json"))) if os.path.exists(os.path.join(P, f"{tag}_bench_reference.json")) else None
    ph = {k: round(v, 1) for k, v in b["phase_ms_per_step"].items()}
    notes = open(os.path.join(P, "r2_reading.md")).read() if os.path.exists(os.path.join(P, "r2_reading.md")) else ""
    md = f"""# Round 2 — final measurements (B200, driver 580, CUDA 12.9)

(generated by `tools/make_r2_profile.py` from the files a `gpurun` call left behind; the reading at the end is `profiles/r2_reading.md`)

## bench.py (default command: `python bench.py`, 3 warm-up + 3 timed steps, 256 MB mixed corpus, -9)

* value (input resident in HBM): **{b['value']:.3f} GB/s**, {b['ms_per_step']:.1f} ms per step, {b['gpu_launches']} kernel launches per {b['steps']} steps
* e2e (pinned host -> H2D -> kernels -> D2H of the frame): **{b['e2e']['value']:.3f} GB/s**
* parity: {json.dumps(b['parity'])}
* phases per step (CUDA events on the library's stream, ms): {json.dumps(ph)}
* compression ratio {b['compression_ratio']:.3f}; clocks {b['clocks']}
* extra (device-resident GB/s, 256 MB each): {json.dumps(b.get('extra', {}))}
* cpu_baseline: {json.dumps(b.get('cpu_baseline', {}))}
* reference arm (`bench.py --impl reference`): {json.dumps({k: ref[k] for k in ('value', 'unit', 'ms_per_step', 'cpu_baseline')}) if ref else 'not run'}
* roofline of the dominant phase: {b['roofline']['kernel']}: {b['roofline']['achieved']:.0f} GB/s of algorithmic bytes
  ({b['roofline']['algorithmic_bytes_per_input_byte']} B per position) against {b['roofline']['peak']} GB/s measured HBM copy bandwidth
  = {b['roofline']['frac']:.3f}; limited by: {b['roofline']['limited_by']}
* whole step: {b['roofline']['step']['algorithmic_bytes_per_input_byte']} B per position -> {b['roofline']['step']['achieved']:.0f} GB/s = {b['roofline']['step']['frac']:.3f} of the measured peak
* full lines: `profiles/{tag}_bench.json`, `profiles/{tag}_bench_reference.json`

## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, `profiles/{tag}_launches.csv`)

Per-launch times are cold-cache and serialised; compare shares with the live CUDA-event phase times above.

| kernel | launches | avg ms | share |
|---|---|---|---|
{table}

## ncu --set full at 256 MB (`python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline`, first batch)

Every counter as a fraction of its peak; shared memory as `l1tex__data_pipe_lsu_wavefronts_mem_shared.sum` / (148 SMs x
`sm__cycles_elapsed.avg`), i.e. wavefronts per SM-cycle of a pipe that retires one per cycle.

""" + "\n".join(lines) + "\n" + notes
    open(os.path.join(P, "r2_summary.md"), "w").write(md)
    print(md[:3000])


if __name__ == "__main__":
    main()
