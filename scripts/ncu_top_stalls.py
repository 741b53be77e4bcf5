"""Top stall-sample instructions of one kernel in an ncu report (SASS view).
usage: ncu_top_stalls.py rep name-substring [N]   (substring is matched against the demangled kernel name)"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur["hdr"] is not None and len(r) == len(cur["hdr"]): cur["data"].append(r)
sel = [b for b in blocks if kern in b["name"].replace("(int)", "").replace("(bool)", "")]
if not sel:
    print("no kernel matches; have:"); [print("  ", b["name"]) for b in blocks]; sys.exit(1)
b = sel[0]; hdr = b["hdr"]; data = b["data"]
print(b["name"])
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[iW] or 0) for r in data); toti = sum(int(r[iI] or 0) for r in data)
print("total samples", tot, "static instrs", len(data), "warp-instr executed", toti)
top = sorted(range(len(data)), key=lambda i: -int(data[i][iW] or 0))[:N]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} exec {r[iI]:>9s} samples {r[iW]:>7s} {100*int(r[iW] or 0)/max(tot,1):5.1f}%  {r[iS].strip()[:120]}")
