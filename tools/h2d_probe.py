"""Host-to-device copy ceiling with N ranks copying at once: what bounds `e2e` on a multi-GPU box.

Every rank binds to its GPU's NUMA node (as bench.py does), allocates the staging buffers of one bench step (32 scans x 1.8 MB x 8
frames) through the library (plain pinned, or write-combined with VILF_HOST_WC=1), fills them, and after a barrier all ranks copy
them to their GPU at the same time — one cudaMemcpyAsync per scan (the bench's pattern) and one per frame (32 scans contiguous).
Rank 0 prints one JSON line with the per-rank rates.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
    VILF_HOST_WC=1 python -m torch.distributed.run ... tools/h2d_probe.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (bind_to_gpu_numa_node)
from vil_fusion_b200 import cabi  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
numa = bench.bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, F, PTS = 32, 8, 113000
host = cabi.host_alloc(F * S * PTS * 16).view(np.float32).reshape(F, S, PTS * 4)
host[:] = 1.0  # touch every page (first touch = this rank's NUMA node); write-combined memory is only ever written
tgt = torch.empty((S, PTS * 4), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()


def run(per_scan: bool, reps: int = 6):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nbytes = 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with torch.cuda.stream(st):
        for rep in range(reps + 1):
            if rep == 1:
                ev0.record(st)
            for f in range(F):
                if per_scan:
                    for s in range(S):
                        cabi.memcpy_h2d_async(tgt[s].data_ptr(), host[f, s].ctypes.data, PTS * 16, st.cuda_stream)
                else:
                    cabi.memcpy_h2d_async(tgt.data_ptr(), host[f].ctypes.data, S * PTS * 16, st.cuda_stream)
                if rep >= 1:
                    nbytes += S * PTS * 16
        ev1.record(st)
    st.synchronize()
    return nbytes / (ev0.elapsed_time(ev1) * 1e-3) / 1e9


res = torch.tensor([run(True), run(False)], dtype=torch.float64, device="cuda")
if world > 1:
    allr = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    a = np.array([t.cpu().numpy() for t in allr])
    print(json.dumps(dict(n_gpus=world, staging="write-combined pinned" if os.environ.get("VILF_HOST_WC") else "pinned", host_placement_rank0=numa,
                          bytes_per_copy=dict(per_scan=PTS * 16, per_frame=S * PTS * 16),
                          gbs_per_gpu_one_copy_per_scan=[round(float(v), 2) for v in a[:, 0]], gbs_per_gpu_one_copy_per_frame=[round(float(v), 2) for v in a[:, 1]],
                          aggregate_gbs=dict(per_scan=round(float(a[:, 0].sum()), 1), per_frame=round(float(a[:, 1].sum()), 1)))))
if world > 1:
    dist.destroy_process_group()
