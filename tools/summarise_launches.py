"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (profiles/)."""
import csv, sys, collections
src, title = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((r["Kernel Name"], ns))
agg = collections.OrderedDict()
for k, ns in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
qs_tot = sum(a[1] for k, a in agg.items() if "qs::" in k or k.startswith("k_"))
print(title)
print()
print("| kernel | launches | total ns | share of all | share of qs:: kernels |")
print("|---|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    mine = "qs::" in k or k.startswith("k_")
    print("| %s | %d | %d | %.2f%% | %s |" % (k[:90], a[0], a[1], 100 * a[1] / tot, ("%.2f%%" % (100 * a[1] / qs_tot)) if mine else "-"))
