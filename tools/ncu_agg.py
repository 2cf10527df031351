"""Developer tool: sums an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import collections, csv, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    tot += ms
print(f"total {tot:.2f} ms")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% {n:6d} x {ms / n:8.4f}  {k[:100]}")
