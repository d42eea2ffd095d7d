#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which hardware paths a kernel uses (B200_PROFILING.md):
scripts/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lime_cikm25_b200", "liblime_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "LDGSTS", "LDSM", "HMMA", "FFMA2", "FMUL2", "FADD2",
        "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "SYNCS", "BAR", "ATOM", "RED", "LDL", "STL", "MATCH"]
cur, cnt, tot = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = re.sub(r"\(anonymous namespace\)::|lime::|void ", "", cur)
        if not cur or cur.startswith("_Z"):          # internal-linkage names c++filt leaves mangled
            name = None
            for mm in re.finditer(r"([a-z_0-9]+_kernel)E", m.group(1)):       # <length><name>E: find the split whose length matches
                t = mm.group(1)
                for i in range(len(t)):
                    if i >= 2 and t[i - 2:i].isdigit() and t[i].isalpha() and int(t[i - 2:i]) == len(t) - i:
                        name = t[i:]
            cur = name or m.group(1)[:46]
        while cur in cnt:
            cur += "'"
        cnt[cur], tot[cur] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        tot[cur] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or (k in ("LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "LDL", "STL") and op.startswith(k) and not op.startswith("LDGSTS") and not op.startswith("LDSM")):
                cnt[cur][k] += 1
                break
print("# cuobjdump -sass %s: static instruction counts per kernel (columns with a non-zero entry only)" % os.path.basename(lib))
print("# UTMALDG/UTMASTG/UBLKCP = TMA, UTCHMMA = tcgen05.mma (kind::f16), LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,")
print("# LDGSTS = cp.async, LDSM/HMMA = ldmatrix / mma.sync, FFMA2 = packed fp32x2, SYNCS = mbarrier ops, LDL/STL = spills")
used = [k for k in KEYS if any(c[k] for c in cnt.values())]
print("%-46s %6s " % ("kernel", "instrs") + " ".join("%7s" % k for k in used))
for f, c in cnt.items():
    print("%-46s %6d " % (f[:46], tot[f]) + " ".join("%7d" % c[k] for k in used))
