import sys, numpy as np
sys.path.insert(0,'.')
from vil_fusion_b200 import cabi, synth
flags=int(sys.argv[1])
seq=synth.Sequence("hdl64", 6, seed=13)
g=cabi.Odometry(cabi.default_config(flags=flags, max_scan_points=116000, max_map_points=1<<18))
for i in range(6):
    print(i, g.process_scan(seq[i][0])[4:], flush=True)
print("ok", flags)
