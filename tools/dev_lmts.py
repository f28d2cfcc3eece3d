"""Dev: phase timestamps of k_solve (library built with -DVILF_LM_TIMING prints them at frame 20)."""
import sys, numpy as np
sys.path.insert(0, ".")
from vil_fusion_b200 import cabi, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = 23
seqs = [synth.Sequence("hdl64", F, seed=s) for s in range(S)]
b = cabi.Batch(cabi.default_config(max_scan_points=116000, max_map_points=1 << 18, max_ring_points=1864), S)
for f in range(F):
    scans = [np.ascontiguousarray(seqs[s][f][0]) for s in range(S)]
    b.wait(b.submit(scans))
print(b.seqs[0].counts())
