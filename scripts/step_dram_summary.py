"""Per-kernel time and DRAM bytes of ONE step from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list of host-driven steps (steps are delimited by the mask kernel)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
ids = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = ids.setdefault(r["ID"], {"k": r["Kernel Name"][:60]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
seq = list(ids.values())
starts = [i for i, d in enumerate(seq) if "mask_bernoulli" in d["k"]]
a, b = starts[1], starts[2]
tot = collections.defaultdict(lambda: [0, 0, 0, 0])
for d in seq[a:b]:
    t = tot[d["k"]]
    t[0] += 1; t[1] += d.get("gpu__time_duration.sum", 0); t[2] += d.get("dram__bytes_read.sum", 0); t[3] += d.get("dram__bytes_write.sum", 0)
T = sum(v[1] for v in tot.values()); R = sum(v[2] for v in tot.values()); W = sum(v[3] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e3:8.1f} us n={v[0]:2d} rd {v[2] / 1e6:8.1f} MB wr {v[3] / 1e6:8.1f} MB {(v[2] + v[3]) / max(v[1], 1):6.0f} GB/s  {k}")
print(f"step: {T / 1e6:.3f} ms under ncu (serialised), DRAM read {R / 1e9:.3f} GB + written {W / 1e9:.3f} GB = {(R + W) / 1e9:.3f} GB")
