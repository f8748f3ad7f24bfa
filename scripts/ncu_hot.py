"""Top stall-sample SASS instructions of one kernel of an .ncu-rep (source page)."""
import csv, subprocess, sys
rep, kid = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "1")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:160])
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
iS, iN, iA = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Address")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[1:]):
    if len(r) <= iN: continue
    try: n = int(r[iN])
    except ValueError: continue
    data.append((n, idx, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, idx, r in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n/tot*100:5.1f}% #{idx:5d} {r[iS].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
