"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares."""
import csv, re, sys, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    name = re.sub(r"<.*", lambda m: m.group(0)[:40], name)
    ns = float(row["Metric Value"].replace(",", ""))
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += ns
total = sum(v[1] for v in tot.values())
print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total/1e6:.2f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {v[0]:8d} {v[1]/1e6:10.3f} {100*v[1]/total:6.1f}%")
