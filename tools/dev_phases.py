"""Dev: per-phase times of the cluster voxel kernel (run on the GPU box)."""
import sys, numpy as np
sys.path.insert(0, ".")
from vil_fusion_b200 import cabi, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = 12
seqs = [synth.Sequence("hdl64", F, seed=s) for s in range(S)]
b = cabi.Batch(cabi.default_config(max_scan_points=116000, max_map_points=1 << 18, max_ring_points=1864), S)
for f in range(F):
    scans = [np.ascontiguousarray(seqs[s][f][0]) for s in range(S)]
    b.wait(b.submit(scans))
names = ["append+bbox", "sort", "heads", "centroids", "grid", "-"]
for lane in (0, S - 1):
    c = b.seqs[lane].counts()
    print("lane", lane, c)
    for job, nm in enumerate(["scan edge", "scan surf", "map edge", "map surf"]):
        t = b.seqs[lane].voxel_phases(job)
        d = np.diff(t[:6]) / 1e3
        if job < 2: d[3] = (t[5] - t[3]) / 1e3; d[4] = 0
        print("  %-10s total %7.1f us  " % (nm, (t[5] - t[0]) / 1e3) + "  ".join("%s %.1f" % (n, x) for n, x in zip(names, d)))
t0 = min(b.seqs[l].voxel_phases(j)[0] for l in range(S) for j in (0, 1)); t1 = max(b.seqs[l].voxel_phases(j)[5] for l in range(S) for j in (0, 1))
print("scan kernel span %.1f us" % ((t1 - t0) / 1e3))
t0 = min(b.seqs[l].voxel_phases(j)[0] for l in range(S) for j in (2, 3)); t1 = max(b.seqs[l].voxel_phases(j)[5] for l in range(S) for j in (2, 3))
print("map kernel span %.1f us" % ((t1 - t0) / 1e3))
