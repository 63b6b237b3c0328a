#!/usr/bin/env python
"""Key counters of every kernel in an ncu report (read here, no GPU):  python tools/ncu_metrics.py gpurun_out/x.ncu-rep
Each counter is also given as a fraction of its peak where ncu reports one; shared memory as wavefronts per SM-cycle."""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    for vals in rr[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print("###", d["Kernel Name"][:100])
        for w in WANT:
            if w in d:
                print(f"- {w} = {d[w]} {u[w]}")
        try:
            sms = 148
            cyc = float(d["sm__cycles_elapsed.avg"].replace(",", ""))
            wf = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", ""))
            print(f"- shared-memory wavefronts per SM-cycle = {wf / (sms * cyc):.3f} (peak 1.0)")
        except Exception:
            pass
        st = sorted(((k.split("issue_stalled_")[1].split("_per")[0], float(d[k].replace(",", ""))) for k in hdr
                     if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), key=lambda x: -x[1])[:7]
        print("- stalls per issue: " + ", ".join(f"{a} {b:.2f}" for a, b in st))


if __name__ == "__main__":
    main()
