#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo).
usage: python scripts/ncu_source.py rep [kernel-id like :::1] [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
rows = list(csv.reader(lines[start:]))
hdr = rows[0]
iline, isrc = 0, 1
iinst = hdr.index("Instructions Executed")
isamp = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.OrderedDict()
cur = None
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    if r[iline]:
        cur = (r[iline], r[isrc].strip())
    if cur is None:
        continue
    a = agg.setdefault(cur, {"inst": 0, "samp": 0, "stalls": collections.Counter()})
    try:
        a["inst"] += int(r[iinst] or 0)
        a["samp"] += int(r[isamp] or 0)
        for i in stall_cols:
            if r[i]:
                a["stalls"][hdr[i]] += int(r[i])
    except ValueError:
        pass
tot_i = sum(a["inst"] for a in agg.values()) or 1
tot_s = sum(a["samp"] for a in agg.values()) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("--- by instructions ---")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
    print(f"{ln:>5} {100*a['inst']/tot_i:5.1f}% inst {100*a['samp']/tot_s:5.1f}% samp  {src[:90]}")
print("--- by stall samples ---")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in a["stalls"].most_common(3))
    print(f"{ln:>5} {100*a['samp']/tot_s:5.1f}% samp  [{st}]  {src[:70]}")
