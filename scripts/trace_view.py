"""Pretty-prints one steady-state tile of a PMVAE_FUSED_TRACE log."""
import sys
lines = open(sys.argv[1]).read().splitlines()
which = int(sys.argv[2]) if len(sys.argv) > 2 else 2
roles = {}; cur = None
for l in lines:
    if l.startswith('trace role'):
        cur = int(l.split()[2]); roles[cur] = []
    elif cur is not None and '@' in l:
        a, b = l.split('@'); roles[cur].append((int(a), int(b)))
mma = roles[0]; epi = roles[1]
idx = [i for i,(t,c) in enumerate(mma) if t == 100]
print("tiles:", [mma[i][1] for i in idx][:8])
ev = [("MMA", t, c) for t, c in mma[idx[which]:idx[which+1]]]
idx = [i for i,(t,c) in enumerate(epi) if t == 1000]
ev += [("EPI", t, c) for t, c in epi[idx[which]:idx[which+1]]]
ev.sort(key=lambda e: e[2])
t0 = ev[0][2]; last = {"MMA": t0, "EPI": t0}
for r, t, c in ev:
    pad = "" if r == "MMA" else " " * 40
    print(f"{pad}{r} {t:5d} @ {c-t0:7d}  +{c-last[r]}")
    last[r] = c
