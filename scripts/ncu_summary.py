"""Summarise ncu outputs: launch-list shares by kernel, or key metrics of a full capture."""
import collections, csv, re, subprocess, sys


def launches(path, top=18):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        grid = row.get("Grid Size", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] / tot * 100:6.2f}%  n={v[0]:4d}  {v[1] / 1e3:10.1f} us  avg {v[1] / v[0] / 1e3:8.1f} us  {k[:80]}")
    print(f"total {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")


WANT = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_bytes.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active"]


def full(path, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    names = [w for w in list(WANT) + list(extra) if w in hdr]
    for r in rows[2:]:
        print("----")
        for w in names:
            i = hdr.index(w)
            print(f"  {w:75s} {r[i][:70]:>20s} {units[i]}")
    if extra == ("LIST",):
        print([h for h in hdr if "tensor" in h or "pipe" in h][:80])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], tuple(sys.argv[3:]))
