"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    name = row["Kernel Name"][:78]
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:24]:
    print(f"{v / 1e6 / steps:9.3f} ms/step {100 * v / T:5.1f}% n={cnt[k]:4d}  {k}")
print(f"total {T / 1e6 / steps:.3f} ms/step over {steps:g} steps")
