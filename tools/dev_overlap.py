"""Dev: per-kernel event intervals of stream group 0 when it runs alone vs together with G-1 other groups."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from vil_fusion_b200 import cabi, synth
S, G, F = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 4, 40
seqs = [synth.Sequence("hdl64", F, seed=s) for s in range(S)]
cap = 115200
pin = cabi.host_alloc(S * F * cap * 16).view(np.float32).reshape(S, F, cap, 4)
scans = []
for s in range(S):
    row = []
    for f in range(F):
        x = seqs[s][f][0]
        pin[s, f, : x.shape[0]] = x
        row.append(pin[s, f, : x.shape[0]])
    scans.append(row)
cfg = cabi.default_config(max_scan_points=116000, max_map_points=1 << 18, max_ring_points=1864)
per = S // G
def run(ng):
    bs = [cabi.Batch(cfg, per) for _ in range(ng)]
    for f in range(8):
        ts = [b.submit([scans[g * per + i][f] for i in range(per)]) for g, b in enumerate(bs)]
        for b, t in zip(bs, ts): b.wait(t)
    bs[0].seqs[0].profile(True)
    t0 = time.perf_counter()
    pend = []
    for f in range(8, F):
        pend.append([b.submit([scans[g * per + i][f] for i in range(per)]) for g, b in enumerate(bs)])
        if len(pend) >= 3:
            for b, t in zip(bs, pend.pop(0)): b.wait(t)
    for ts in pend:
        for b, t in zip(bs, ts): b.wait(t)
    dt = time.perf_counter() - t0
    kt = bs[0].seqs[0].profile_kernels()
    for b in bs: b.close()
    return dt, kt
d1, k1 = run(1)
dg, kg = run(G)
print("alone: %.1f us/frame (group of %d)   with %d groups: %.1f us/step -> %.0f scans/s" % (d1 / (F - 8) * 1e6, per, G, dg / (F - 8) * 1e6, S * (F - 8) / dg))
tot1 = sum(v[0] for v in k1.values()); totg = sum(v[0] for v in kg.values())
print("sum of kernel intervals per frame: alone %.1f us, concurrent %.1f us" % (tot1 / (F - 8) * 1e3, totg / (F - 8) * 1e3))
for k in sorted(k1, key=lambda k: -k1[k][0]):
    a, b = k1[k], kg.get(k, (0, 1))
    print("  %-36s alone %7.1f us   concurrent %7.1f us   x%.2f" % ("/".join(k), 1e3 * a[0] / a[1], 1e3 * b[0] / max(b[1], 1), (b[0] / max(b[1], 1)) / (a[0] / a[1])))
