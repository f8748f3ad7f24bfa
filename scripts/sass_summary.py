"""Per-kernel counts of the Blackwell-native SASS mnemonics in libpmvae.so (cuobjdump -sass): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores, SYNCS = mbarrier ops.
    python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "posterior_matching_b200", "libpmvae.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)")
keys = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = pat.match(line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        counts[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(so, ROOT)} (sm_100a); kernels without tensor / TMA instructions are")
print("# listed with their FFMA count only.  " + " ".join(keys))
tot = collections.Counter()
for name, c in counts.items():
    tot.update(c)
    if any(c[k] for k in keys[:7]):
        print(f"{name[:110]:110s} " + " ".join(f"{k}={c[k]}" for k in keys if c[k]))
print()
print("TOTAL " + " ".join(f"{k}={tot[k]}" for k in keys))
print(f"kernels: {len(counts)}, of which with UTCHMMA: {sum(1 for c in counts.values() if c['UTCHMMA'])}")
