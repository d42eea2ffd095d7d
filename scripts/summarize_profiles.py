#!/usr/bin/env python
"""Turn the raw ncu outputs a gpurun call left in gpurun_out/ into the small text summaries that
are committed under profiles/ (gpurun_out/ itself is scratch and git-ignored).

  python scripts/summarize_profiles.py r01a      -> profiles/r01a_*.txt
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def launches(tag, name="launches.csv"):
    path = os.path.join(OUT, name)
    if not os.path.isfile(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h = rows[hdr]
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for r in rows[hdr + 1:]:
        d = dict(zip(h, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        k = d["Kernel Name"].split("(")[0].replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += v
        seq.append((d["ID"], k, d["Grid Size"], d["Block Size"], v))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, tag + "_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch, cold-cache, serialised)\n")
        f.write("# source: gpurun_out/%s, %d launches, %.3f ms total\n" % (name, len(seq), tot / 1e6))
        f.write("%-58s %6s %12s %7s\n" % ("kernel", "count", "total_ms", "share"))
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-58s %6d %12.3f %6.1f%%\n" % (k[:58], v[0], v[1] / 1e6, 100 * v[1] / tot))
        f.write("\n# launch list (id, kernel, grid, block, ns)\n")
        for s in seq:
            f.write("%s %s %s %s %.0f\n" % s)


def details(tag, rep, kernel_regex=None):
    path = os.path.join(OUT, rep)
    if not os.path.isfile(path):
        return
    base = os.path.splitext(rep)[0]
    txt = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
    # keep the first profiled launch only
    parts = txt.split("\n  void ")
    first = parts[0] + ("\n  void " + parts[1] if len(parts) > 1 else "")
    with open(os.path.join(PROF, "%s_%s_details.txt" % (tag, base)), "w") as f:
        f.write(first)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) >= 3:
        names, units, vals = rows[0], rows[1], rows[2]
        keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
                "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
                "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
                "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
                "smsp__issue_active.avg.pct_of_peak_sustained_active")
        with open(os.path.join(PROF, "%s_%s_raw.txt" % (tag, base)), "w") as f:
            f.write("# selected raw metrics of the first profiled launch (ncu --set full --clock-control none)\n")
            for n, u, v in zip(names, units, vals):
                if n in keep or n.startswith("sm__inst_executed_pipe_") and n.endswith(".sum") or "tensor" in n and n.endswith("pct_of_peak_sustained_active"):
                    f.write("%-90s %-14s %s\n" % (n, u, v))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    for rep in sorted(os.listdir(OUT)):
        if rep.endswith(".ncu-rep"):
            details(tag, rep)
    print("\n".join(sorted(os.listdir(PROF))))
