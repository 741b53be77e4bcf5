"""Aggregate an ncu source page by CUDA source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if r and ("Source" in r) and ("Instructions Executed" in r):
        hdr = r; start = i + 1; break
print(hdr[:6])
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)")
iL = hdr.index("Line") if "Line" in hdr else None
agg = []
for r in rows[start:]:
    if len(r) != len(hdr): continue
    try: n = int(r[iI] or 0); w = int(r[iW] or 0)
    except ValueError: continue
    agg.append((n, w, r[0][:12], r[iS][:150]))
tot = sum(a[0] for a in agg); tw = sum(a[1] for a in agg)
print("total", tot, tw)
for n, w, l, src in sorted(agg, key=lambda a: -a[1])[:45]:
    print(f"{100*n/tot:5.1f}% instr {100*w/max(tw,1):5.1f}% stall | {l} | {src}")
