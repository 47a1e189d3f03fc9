"""Summarise CROSSCLR_PAIR_TRACE tables: mean cycles per S tile / P tile for each role of cluster 0."""
import sys
import numpy as np
for path in sys.argv[1:]:
    rows = [l.split() for l in open(path)]
    T = {(int(r[0]), int(r[1])): [int(x) for x in r[2:]] for r in rows}
    def ok(role, t, n=3): return (role, t) in T and all(x > 0 for x in T[(role, t)][:n])
    per = [T[(0, t + 1)][0] - T[(0, t)][0] for t in range(2, 26) if ok(0, t) and ok(0, t + 1)]
    swait = [T[(0, t)][1] - T[(0, t)][0] for t in range(2, 26) if ok(0, t)]
    iss = [T[(0, t)][2] - T[(0, t)][1] for t in range(2, 26) if ok(0, t)]
    out = [f"{path}: S period {np.mean(per):.0f} (sempty wait {np.mean(swait):.0f}, issue {np.mean(iss):.0f})" if per else f"{path}: no S"]
    for role in (1, 2, 3):
        ep = [T[(role, t)][3] - T[(role, t)][1] for t in range(1, 13) if ok(role, t, 4)]
        ew = [T[(role, t)][1] - T[(role, t)][0] for t in range(1, 13) if ok(role, t, 4)]
        eh = [T[(role, t)][2] - T[(role, t)][1] for t in range(1, 13) if ok(role, t, 4)]
        if ep: out.append(f"epi{role} active {np.mean(ep):.0f} (TMEM held {np.mean(eh):.0f}) wait {np.mean(ew):.0f}")
    gi = [T[(5, t)][2] - T[(5, t)][1] for t in range(2, 50) if ok(5, t)]
    gw = [T[(5, t)][1] - T[(5, t)][0] for t in range(2, 50) if ok(5, t)]
    gp = [T[(5, t + 1)][0] - T[(5, t)][0] for t in range(2, 50) if ok(5, t) and ok(5, t + 1)]
    if gi: out.append(f"G period {np.mean(gp):.0f} (P wait {np.mean(gw):.0f}, issue {np.mean(gi):.0f})")
    last = max(x for v in T.values() for x in v); first = min(x for v in T.values() for x in v if x > 0)
    out.append(f"span {last - first}")
    print("; ".join(out))
    w = [(t, T[(4, t)]) for t in range(16) if (4, t) in T and T[(4, t)][3] > 0]
    if w:
        print("  per epilogue warp of CTA 0 (cycles per tile): " + "  ".join(
            f"w{t + 4}[sfull {v[0] / v[3]:.0f} read {v[1] / v[3]:.0f} math {v[2] / v[3]:.0f}]" for t, v in w))
