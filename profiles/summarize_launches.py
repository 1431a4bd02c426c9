"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table on stdout)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1000.0 if u == "ns" else v * 1000.0 if u == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {v[0]} | {v[1]/1000:.2f} | {v[1]/v[0]:.1f} | {v[1]/tot*100:.1f}% |")
