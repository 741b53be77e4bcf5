"""Aggregate an ncu source page (SASS view) by opcode: instructions executed and stall samples."""
import csv, subprocess, sys, collections, re
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# may contain several kernels: take the first block
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": 
        if hdr is not None: break
        hdr = r; continue
    if hdr is not None and len(r) == len(hdr): data.append(r)
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)")
byop = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r in data:
    op = r[iS].split()[0] if not r[iS].strip().startswith("@") else r[iS].split()[1]
    op = op.split(".")[0]
    n = int(r[iI] or 0); w = int(r[iW] or 0)
    byop[op][0] += n; byop[op][1] += w; byop[op][2] += 1
    tot_i += n; tot_s += w
print(f"total warp-instr {tot_i}, stall samples {tot_s}, static instrs {len(data)}")
for op, (n, w, c) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:12s} exec {n:12d} ({100*n/tot_i:5.1f}%)  samples {w:8d} ({100*w/max(tot_s,1):5.1f}%)  static {c}")
