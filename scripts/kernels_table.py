"""Print the per-kernel table of a bench.py JSON line (file argument or stdin)."""
import json, sys
d = json.loads((open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin).read().strip().splitlines()[-1])
print(f"value {d['value']:.2f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.2f}  launches {d.get('gpu_launches')}")
for k in d.get("kernels", []):
    print(f"  {k['kernel']:18s} {k['ms_per_step']:8.3f} ms  x{k['launches_per_step']:3d}  {k['achieved']:9.1f} {k['unit']:8s} frac {k['frac']:.3f} ({k['bound']})")
if "lewin_block_us" in d:
    b = d["lewin_block_us"]
    print("  block us:", " ".join(f"{l['level']}:{l['fwd_us']:.0f}/{l['fwd_bwd_us']:.0f}" for l in b["per_level"]))
if d.get("train_step"):
    print("  train_step:", d["train_step"])
