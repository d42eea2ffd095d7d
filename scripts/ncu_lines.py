#!/usr/bin/env python
"""Per-source-line instruction counts, stall samples and the dominant stall reasons of an ncu report:
scripts/ncu_lines.py gpurun_out/x.ncu-rep [N] [file-filter]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; data = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        try:
            st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
            data.append((cur, int(r[0]), r[1].strip()[:95], int(d["# Samples"]), int(d["Instructions Executed"]), st))
        except (ValueError, KeyError): pass
ts = sum(d[3] for d in data); ti = sum(d[4] for d in data)
print("samples %d  warp-instructions %d" % (ts, ti))
tot = {}
for d in data:
    for k, v in d[5].items(): tot[k] = tot.get(k, 0) + v
print("stall totals: " + "  ".join("%s %.1f%%" % (k, 100 * v / ts) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:12]))
for d in sorted(data, key=lambda x: -x[3])[:top]:
    s = " ".join("%s:%d" % (k, v) for k, v in sorted(d[5].items(), key=lambda x: -x[1])[:3])
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s   [%s]" % (100 * d[3] / max(ts, 1), 100 * d[4] / max(ti, 1), d[0], d[1], d[2], s))
