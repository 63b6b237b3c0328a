#!/usr/bin/env python
"""profiles/r2_sass_excerpts.txt: the UBLKCP / mbarrier staging and the fast loop of k_search, and the tile input and
ranking of k_lsd_pass2, cut out of `cuobjdump -sass smallz4_b200/libsmallz4_b200.so` (no GPU needed)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "smallz4_b200", "libsmallz4_b200.so")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    return [l.split("Function :")[1].strip() for l in out.splitlines() if "Function :" in l]


def sass(mangled):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, SO], capture_output=True, text=True).stdout.splitlines()
    out = [l for l in out if not l.strip().startswith("/* 0x")]            # the second encoding line of every instruction
    return [l.split("/* 0x")[0].rstrip() for l in out]


def around(lines, words, before=1, after=1):
    idx = [i for i, l in enumerate(lines) if any(w in l for w in words)]
    seen, out = set(), []
    for i in idx:
        for j in range(max(0, i - before), min(len(lines), i + after + 1)):
            if j not in seen:
                seen.add(j)
                out.append(lines[j])
    return out


def main():
    fs = functions()
    ks = sass(next(f for f in fs if "8k_search" in f))
    lds = [i for i, l in enumerate(ks) if "LDS.U16" in l]
    best = max((sum(1 for x in lds if a <= x < a + 150), a) for a in lds)
    p2 = sass(next(f for f in fs if "k_lsd_pass2ILj3ELb0" in f))
    with open(os.path.join(ROOT, "profiles", "r2_sass_excerpts.txt"), "w") as f:
        f.write("SASS excerpts of libsmallz4_b200.so (sm_100a), cuobjdump -sass, round 2 final build (tools/make_sass_excerpts.py)\n")
        f.write("\n==== k_search: staging of the data / pe8 window (cp.async.bulk -> UBLKCP, mbarrier -> SYNCS) ====\n")
        f.write("\n".join(around(ks, ["UBLKCP", "SYNCS"], 2, 2)))
        f.write(f"\n\n==== k_search: the fast loop, unrolled by eight ({best[0]} chain-entry loads LDS.U16 in these 150 instructions; per candidate:\n"
                "==== one add, LDS.U16 chain entry, two LDS words + SHF funnel shift, compares and predicated state selects) ====\n")
        f.write("\n".join(ks[best[1] - 4: best[1] + 146]))
        f.write("\n\n\n==== k_lsd_pass2<3,false>: tile input by UBLKCP on an mbarrier, ranking by MATCH.ANY ====\n")
        f.write("\n".join(around(p2, ["UBLKCP", "SYNCS", "MATCH"])))
        f.write("\n")


if __name__ == "__main__":
    main()
