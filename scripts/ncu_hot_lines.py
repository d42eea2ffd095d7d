#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report: scripts/ncu_hot_lines.py gpurun_out/x.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; data = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try: data.append((cur, int(r[0]), r[1].strip()[:100], int(r[6]), int(r[7])))
        except ValueError: pass
ts = sum(d[3] for d in data); ti = sum(d[4] for d in data)
print("samples %d  warp-instructions %d" % (ts, ti))
for d in sorted(data, key=lambda x: -x[3])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * d[3] / max(ts, 1), 100 * d[4] / max(ti, 1), d[0], d[1], d[2]))
