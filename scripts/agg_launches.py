"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid)."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("femb::", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v * 1e-3 if u == "ns" else v if u == "us" else v * 1e3
    agg[(name[:60], row["Grid Size"])][0] += 1
    agg[(name[:60], row["Grid Size"])][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{k[0]:42s} grid={k[1]:>15s} n={v[0]:4d} total={v[1]/1e3:8.3f} ms ({v[1]/tot*100:5.1f}%) avg={v[1]/v[0]:8.1f} us")
