#!/usr/bin/env python
"""Source lines of one kernel ranked by executed warp instructions (and stall samples), from an ncu report captured with
--import-source on:   python tools/ncu_hot_lines.py report.ncu-rep kernel_regex [top] [launch_skip]"""
import csv, subprocess, sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur_file, hdr = "", None
    lines = []
    seen_kernel = 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Function Name":
            seen_kernel += 1
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            d = dict(zip(hdr, r))
            i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples"); i_t = hdr.index("Thread Instructions Executed")
            def num(x):
                try:
                    return float(x)
                except ValueError:
                    return 0.0
            k = len(hdr)                         # source text with quotes may split into extra columns: count from the right
            lines.append((num(r[i_inst - k]), num(r[i_s - k]), num(r[i_t - k]), cur_file, r[0], r[1][:120]))
    tot = sum(x[0] for x in lines) or 1
    tots = sum(x[1] for x in lines) or 1
    print(f"total warp instructions {tot:.3e}, stall samples {tots:.0f}")
    for v, s, t, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * v / tot:5.1f}% inst {100 * s / tots:5.1f}% smp  lanes {t / v if v else 0:4.1f}  {f}:{ln:>4} {src.strip()}")


if __name__ == "__main__":
    main()
