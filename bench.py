#!/usr/bin/env python
"""bench.py — scans/s of the scan-to-map lidar odometry hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--seqs S] [--workload hdl64|vlp32|beams128]

A "step" is one lock-step frame of every sequence in the job: S independent synthetic sequences per GPU
(independent worlds, trajectories and noise), N GPUs, so one step = S*N scans.  Frame 0 of every sequence
(localMapInited) and W-1 further frames are warm-up; exactly K steps are timed.

  value   whole-job scans/s with the scans already resident in HBM (device-to-device copy into the scan slot)
  e2e     the same through the public C ABI with HOST (pinned) scan buffers: per step H2D of S scans, D2H of S poses
  roofline  the kernel with the largest share of the step, timed with CUDA events inside the library in a
            second pass over the same frames; algorithmic bytes per SURVEY.md §8(d) / DESIGN.md §6
  cpu_baseline  the CPU oracle (restated reference path) on all host cores, one sequence per core, bounded sample

Multi-GPU: replicas only — every rank runs its own sequences, no collective on the data path;
torch.distributed (NCCL) is used for the start/stop barrier and the max-over-ranks of the device time.
--impl reference times the reference's CPU path (the oracle restatement; the ROS node cannot be built
offline) on the host cores and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vil_fusion_b200 import replicas, synth  # noqa: E402

WORKLOADS = {
    "hdl64": dict(sensor="hdl64", n_scan=64, n_rings=64, desc="synthetic HDL-64E 64x1800 sequence(s), leaf 0.4/0.8 (configs[3] per-GPU unit)"),
    "vlp32": dict(sensor="vlp32", n_scan=32, n_rings=32, desc="synthetic 32x1800 sequence(s), leaf 0.4/0.8 (configs[1])"),
    "beams128": dict(sensor="beams128", n_scan=0, n_rings=128, desc="synthetic 128x2048 sequence(s), explicit ring ids (configs[4] shape)"),
    # BASELINE.json configs[2]: dense world, 0.09 m voxels, ~1e6 points in the +-100 m crop once the maps have filled (vil_fusion_b200/synth.py DENSE)
    "hdl64_dense": dict(sensor="hdl64", n_scan=64, n_rings=64, seq_kw=dict(density=synth.DENSE["density"], speed=synth.DENSE["speed"], world_length=700.0),
                        cfg_kw=dict(edge_leaf=synth.DENSE["edge_leaf"], surf_leaf=synth.DENSE["surf_leaf"]), map_cap=synth.DENSE["max_map_points"], seqs=8,
                        preroll=320, cpu_frames=20,
                        desc="synthetic HDL-64E 64x1800 sequence(s) in a 3x denser world at 0.4 m/frame, leaf 0.09/0.09: >= 1e6 live map points per sequence (configs[2])"),
}


def make_sequence(w: dict, frames: int, seed: int):
    return synth.Sequence(w["sensor"], frames, seed=seed, **w.get("seq_kw", {}))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hdl64", choices=sorted(WORKLOADS))
    ap.add_argument("--seqs", type=int, default=None, help="independent sequences per GPU (default 32; 4 for hdl64_dense)")
    ap.add_argument("--preroll", type=int, default=None,
                    help="untimed frames every sequence runs before the timed steps, so that they see a saturated local map (default 300; 0 = time the young map like round 1)")
    ap.add_argument("--young-steps", type=int, default=60, help="timed steps of the extra young-map pass (frames right after map init); 0 = skip")
    ap.add_argument("--cpu-preroll", type=int, default=None, help="untimed frames per core before the CPU sample (default 100; the workload's preroll for hdl64_dense)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-sequence latency block")
    ap.add_argument("--depth", type=int, default=3, help="frames in flight (submit ahead of wait)")
    ap.add_argument("--groups", type=int, default=4, help="split the sequences of a GPU into this many lock-step batches, each on its own CUDA stream")
    ap.add_argument("--cpu-frames", type=int, default=300, help="frames per core of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the large-map kNN / map-update stage timings")
    ap.add_argument("--flags", type=int, default=0, help="vilf_config.flags (4 = VILF_FLAG_LEGACY_MAP: the round-1 radix / hashed-grid map path, for A/B runs)")
    ap.add_argument("--map-cap", type=int, default=None, help="capacity of each local map (points); overflow is an error, not a truncation")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle, one sequence per core
# ---------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    idx, scans, rings, n_scan, n_rings, cfg_kw, preroll, barrier = args
    from oracle import orc
    cfg = orc.config(n_scan=n_scan, n_rings=n_rings, **cfg_kw)
    od = orc.Odometry(cfg)
    for x, r in zip(scans[:preroll], rings[:preroll]):  # untimed: fills the local map
        od.process_scan(x, r if n_scan == 0 else None)
    t_pre = od.timing()
    barrier.wait()
    t0 = time.perf_counter()
    for x, r in zip(scans[preroll:], rings[preroll:]):
        od.process_scan(x, r if n_scan == 0 else None)
    dt = time.perf_counter() - t0
    tm = od.timing()
    tm = {k: tm[k] - t_pre[k] for k in tm}  # cumulative seconds per stage: keep the timed part
    return dt, tm


def cpu_baseline(workload: str, frames: int, cores: int | None = None, preroll: int = 0):
    """scans/s of the CPU oracle with `cores` independent sequences on `cores` processes, `frames` timed frames each after
    `preroll` untimed ones (frame 0 = map init)."""
    from oracle import orc
    orc.build()
    w = WORKLOADS[workload]
    cores = cores or (os.cpu_count() or 1)
    data = []
    for c in range(cores):
        seq = make_sequence(w, preroll + frames, 100 + c)
        sc = [seq[i] for i in range(preroll + frames)]
        data.append(([s[0] for s in sc], [s[1] for s in sc]))
    ctx = mp.get_context("fork")
    mgr_barrier = ctx.Barrier(cores)
    res_q = ctx.Queue()

    def run(idx):
        res_q.put(_cpu_worker((idx, data[idx][0], data[idx][1], w["n_scan"], w["n_rings"], w.get("cfg_kw", {}), preroll, mgr_barrier)))

    procs = [ctx.Process(target=run, args=(i,)) for i in range(cores)]
    for p in procs:
        p.start()
    res = [res_q.get() for _ in procs]
    for p in procs:
        p.join()
    wall = max(r[0] for r in res)
    per_stage = {k: float(np.mean([r[1][k] for r in res])) for k in ("extract", "scan_ds", "kd_build", "assoc", "solve", "map_update")}
    return dict(value=cores * frames / wall, unit="scans/s", cores=cores, kind="port",
                sample=f"{cores} independent {workload} sequences x {frames} timed frames after {preroll} untimed ones (frame 0 = map init), one per core, oracle -O3 single-thread each",
                seconds=wall, per_stage_s_per_seq=per_stage)


# ---------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(mx)) if mx else None, reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def algorithmic_bytes(phase: str, kernel: str, c: dict) -> float | None:
    """Algorithmic bytes of ONE launch (all S sequences of the rank), SURVEY.md §8(d) figures; c = mean counts per sequence."""
    N, F, Q, M, S = c["n_scan"], c["n_edge"] + c["n_surf"], c["n_ds"], c["n_map"], c["seqs"]
    per = {
        ("extract", "k_sort_keyhist<KeyGenRing>"): 16 * N + 8 * N,
        ("extract", "k_sort_scatter"): 16 * N,
        ("extract", "k_sector_select"): 4 * N + 16 * N + 20 * F,
        ("extract", "k_compact_features"): 40 * F,
        ("assoc_solve", "k_knn_assoc"): 136 * Q,            # 16 B query + 5 x 16 B neighbours + 40 B result
        ("assoc_solve", "k_fit"): 40 * Q + 80 * Q + 64 * Q,  # result + gathered neighbours + factor record
        ("assoc_solve", "k_solve"): 5 * 64 * Q,             # <= 5 evaluations of every factor record
        ("scan_ds", "k_voxel_cluster"): 16 * F + 16 * Q,    # every feature read once, every voxel written once
        ("map_update", "k_voxel_cluster"): 16 * (M + Q) + 16 * M + 32 * M,  # append + filter + hash build
        ("assoc_solve", "k_knn_cell_assoc"): 136 * Q,
        ("map_update", "k_merge<count>"): 16 * (M + Q),                    # every old and new point read once
        ("map_update", "k_merge<emit>"): 16 * (M + Q) + 16 * M,            # ... read again (L2) and every kept voxel written
        ("map_update", "k_merge_partition"): 24 * Q,
        ("map_update", "k_new_xform"): 32 * Q,
        ("grid_build", None): 32 * M,
        ("scan_ds", None): 16 * F + 16 * Q,
        ("map_update", None): 16 * (M + Q) + 16 * M,
    }
    v = per.get((phase, kernel), per.get((phase, None)))
    return None if v is None else float(v) * S


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and so the first-touch placement of its pinned staging buffers) to the NUMA node the GPU hangs off:
    with 8 ranks copying at once, host memory on the wrong socket halves the H2D rate.  Returns a description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa: single node"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception as e:  # not fatal: the run is just not NUMA-local
        return f"numa: not bound ({type(e).__name__})"


def issue_roofline(nq: int, ms: float, issue_peak: float):
    """Instruction-issue roofline of the 5-NN kernel (it is issue-, not bandwidth-bound): warp instructions per query from the
    committed ncu capture (profiles/r2_knn_issue.json, smsp__inst_executed.sum / queries on this very case) x the query rate
    measured in this run, against 148 SMs x 4 schedulers x 1 instruction per clock."""
    try:
        ent = json.load(open(os.path.join(ROOT, "profiles", "r2_knn_issue.json")))
        ipq = float(ent["warp_instructions_per_query"])
    except Exception:
        return None
    rate = ipq * nq / (ms * 1e-3)
    return dict(warp_instructions_per_query=ipq, source="profiles/r2_knn_issue.json (committed ncu capture of this case; not re-measured in this run)",
                achieved_ginst_s=rate / 1e9, peak_ginst_s=issue_peak / 1e9, frac=rate / issue_peak)


def run_gpu(args, rank: int, world: int, local_rank: int):
    import torch
    from vil_fusion_b200 import cabi

    numa = bind_to_gpu_numa_node(local_rank)

    torch.cuda.set_device(local_rank)
    w = WORKLOADS[args.workload]
    S = args.seqs if args.seqs else w.get("seqs", 32)
    K, W, D = args.steps, max(args.warmup, 3), max(1, min(args.depth, 6))
    P = args.preroll if args.preroll is not None else w.get("preroll", 300)
    KY = min(args.young_steps, K) if P > 0 else 0       # young-map pass (frames right after map init), reported beside the headline
    KP = 0 if args.no_roofline else 24                  # profiled frames (one stream group alone, an event after every kernel)
    map_cap = args.map_cap if args.map_cap else w.get("map_cap", 1 << 18)
    sensor = synth.SENSORS[w["sensor"]]()
    cap = sensor.rings * sensor.n_az
    use_ring = w["n_scan"] == 0
    # frame layout of every sequence: [0, Y) young pass, replayed as the head of the preroll | [Y, P) rest of the preroll, generated on
    # the fly | [P, P+W+K) device-resident pass | [.., +W+K) host (end-to-end) pass | [.., +KP) profiled frames
    Y = (W + KY) if KY else 0
    P = max(P, Y)
    f_dev, f_host, f_prof = P, P + W + K, P + 2 * (W + K)
    F = f_prof + KP
    pinned_budget = 20e9
    if S * (W + K) * cap * 18.0 > pinned_budget:  # the end-to-end pass keeps its scans in pinned host memory (~1.85 MB per scan)
        S_fit = max(4, int(pinned_budget // ((W + K) * cap * 18.0)))
        if rank == 0:
            print(f"bench.py: {S} sequences x {W + K} frames exceed the {pinned_budget / 1e9:.0f} GB pinned staging budget; using {S_fit} sequences per GPU", file=sys.stderr)
        S = S_fit
    t_gen = time.time()
    seqs = [make_sequence(w, F, sd) for sd in replicas.sequence_seeds(rank, world, S)]
    counts = np.zeros((S, F), np.int32)

    def gen_into(f, dst, dst_ring):
        """scan f of every sequence -> dst [S, cap, 4] (host), dst_ring [S, cap]"""
        for s_ in range(S):
            x, r = seqs[s_][f]
            counts[s_, f] = x.shape[0]
            dst[s_, : x.shape[0]] = x
            dst_ring[s_, : x.shape[0]] = r

    chunk = cabi.host_alloc(S * cap * 16).view(np.float32).reshape(S, cap, 4)
    chunk_ring = cabi.host_alloc(S * cap * 2).view(np.uint16).reshape(S, cap)

    def stage_device(f0, f1):
        """frames [f0, f1) of every sequence -> HBM"""
        n = max(f1 - f0, 1)
        dev = torch.empty((n, S, cap, 4), dtype=torch.float32, device="cuda")
        dev_ring = torch.empty((n, S, cap), dtype=torch.int16, device="cuda")
        for f in range(f0, f1):
            gen_into(f, chunk, chunk_ring)
            dev[f - f0].copy_(torch.from_numpy(chunk), non_blocking=False)
            dev_ring[f - f0].copy_(torch.from_numpy(chunk_ring.view(np.int16)), non_blocking=False)
        return dev, dev_ring

    cfg = cabi.default_config(n_scan=w["n_scan"], n_rings=w["n_rings"], max_scan_points=max(cap, 1024), max_map_points=map_cap,
                              max_ring_points=sensor.n_az + 64, flags=args.flags, **w.get("cfg_kw", {}))
    barrier = replicas.barrier
    G = max(1, min(args.groups, S))
    bounds = [round(g * S / G) for g in range(G + 1)]  # group g owns sequences [bounds[g], bounds[g+1])

    def make_batches():
        return [cabi.Batch(cfg, bounds[g + 1] - bounds[g], device=local_rank) for g in range(G)]

    def submit_dev(bs, dev, dev_ring, f0, f, groups=None):
        ts = []
        for g, b in enumerate(bs):
            if groups is not None and g not in groups:
                continue
            rng = range(bounds[g], bounds[g + 1])
            ptrs = [dev[f - f0, s_].data_ptr() for s_ in rng]
            rp = [dev_ring[f - f0, s_].data_ptr() for s_ in rng] if use_ring else None
            ts.append((b, b.submit_dev(ptrs, [int(counts[s_, f]) for s_ in rng], rp)))
        return ts

    def submit_host(bs, host, host_ring, f0, f):
        ts = []
        for g, b in enumerate(bs):
            rng = range(bounds[g], bounds[g + 1])
            ts.append((b, b.submit([host[f - f0, s_, : counts[s_, f]] for s_ in rng], [host_ring[f - f0, s_, : counts[s_, f]] for s_ in rng] if use_ring else None)))
        return ts

    def wait(ts):
        return np.concatenate([b.wait(t) for b, t in ts], axis=0)

    def timed(bs, submit, f_first, n_warm, n_timed):
        """n_warm untimed + n_timed timed lock-step frames starting at frame f_first; D frames in flight.  (device ms, wall ms, launches, poses)"""
        stream = torch.cuda.ExternalStream(bs[0].seqs[0].stream())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        poses = None
        for f in range(f_first, f_first + n_warm):
            poses = wait(submit(f))
        torch.cuda.synchronize()
        barrier()
        l0 = sum(b.seqs[0].launch_count() for b in bs)
        ev0.record(stream)
        t0 = time.perf_counter()
        tickets = []
        for f in range(f_first + n_warm, f_first + n_warm + n_timed):
            tickets.append(submit(f))
            if len(tickets) >= D:
                poses = wait(tickets.pop(0))
        while tickets:
            poses = wait(tickets.pop(0))
        ev1.record(stream)  # every ticket of every group has been waited for: this closes the device interval of the whole pass
        torch.cuda.synchronize()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        return ev0.elapsed_time(ev1), wall_ms, sum(b.seqs[0].launch_count() for b in bs) - l0, poses

    def close_all(bs):
        for b in bs:
            b.close()

    # ---- young-map pass: the frames right after localMapInited (what round 1 timed) ----
    young = None
    dev_y = dev_y_ring = None
    if KY:
        dev_y, dev_y_ring = stage_device(0, Y)
        bs = make_batches()
        ms_y, _, _, _ = timed(bs, lambda f: submit_dev(bs, dev_y, dev_y_ring, 0, f), 0, W, KY)
        cy = [sq.counts() for b in bs for sq in b.seqs]
        close_all(bs)
        ms_y = replicas.max_over_ranks([ms_y], device="cuda")[0]
        young = dict(value=replicas.job_scans(world, S, KY) / (ms_y * 1e-3), unit="scans/s", steps=KY, ms_per_step=ms_y / KY, frames=f"{W}..{W + KY - 1}",
                     mean_map_points=float(np.mean([c["n_map_edge"] + c["n_map_surf"] for c in cy])),
                     note="device-resident scans/s on the frames right after map init, where the local maps are still small (the window round 1 timed)")

    # ---- steady state: untimed preroll, then the timed passes on maps the +-100 m crop has filled ----
    bs = make_batches()
    for f in range(0, Y):                      # head of the preroll: the frames already resident
        wait(submit_dev(bs, dev_y, dev_y_ring, 0, f))
    del dev_y, dev_y_ring
    if P > Y:                                  # rest of the preroll: generated on the fly, submitted from pinned memory, untimed
        ring_sz = 2
        pre = cabi.host_alloc(ring_sz * S * cap * 16).view(np.float32).reshape(ring_sz, S, cap, 4)
        pre_ring = cabi.host_alloc(ring_sz * S * cap * 2).view(np.uint16).reshape(ring_sz, S, cap)
        pend = None
        for f in range(Y, P):
            slot = f % ring_sz
            gen_into(f, pre[slot], pre_ring[slot])
            ts = []
            for g, b in enumerate(bs):
                rng = range(bounds[g], bounds[g + 1])
                ts.append((b, b.submit([pre[slot, s_, : counts[s_, f]] for s_ in rng], [pre_ring[slot, s_, : counts[s_, f]] for s_ in rng] if use_ring else None)))
            if pend is not None:
                wait(pend)
            pend = ts
        if pend is not None:
            wait(pend)
        cabi.host_free(pre.reshape(-1).view(np.uint8)); cabi.host_free(pre_ring.reshape(-1).view(np.uint8))
    dev, dev_ring = stage_device(f_dev, f_dev + W + K)
    host = cabi.host_alloc((W + K) * S * cap * 16).view(np.float32).reshape(W + K, S, cap, 4)
    host_ring = cabi.host_alloc((W + K) * S * cap * 2).view(np.uint16).reshape(W + K, S, cap)
    for f in range(f_host, f_host + W + K):
        gen_into(f, host[f - f_host], host_ring[f - f_host])
    dev_p = dev_p_ring = None
    if KP:
        dev_p, dev_p_ring = stage_device(f_prof, f_prof + KP)
    t_gen = time.time() - t_gen
    torch.cuda.synchronize()

    clocks = Clocks(local_rank)
    clocks.start()
    ms_dev, wall_dev, launches, poses_dev = timed(bs, lambda f: submit_dev(bs, dev, dev_ring, f_dev, f), f_dev, W, K)
    clk = clocks.stop()
    cnt = [sq.counts() for b in bs for sq in b.seqs]
    ms_host, wall_host, _, poses_host = timed(bs, lambda f: submit_host(bs, host, host_ring, f_host, f), f_host, W, K)

    # ---- what the host link can do: the same pinned scans copied with nothing else running (the ceiling of e2e) ----
    def h2d_copy_only(batched: bool):
        """The timed frames' scans copied with the GPU otherwise idle.  batched = False: one cudaMemcpyAsync per scan; True: one copy per
        stream group and frame, rows read up to the group's longest scan (what vilf_submit_scan_batch issues for scans that share a slab).
        Useful bytes (n x 16 per scan) over device time."""
        frames = list(range(f_host + W, min(f_host + W + K, f_host + W + 20)))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tgt = torch.empty((S, cap, 4), dtype=torch.float32, device="cuda")
        st = torch.cuda.Stream()
        nbytes = 0
        with torch.cuda.stream(st):
            for rep in range(2):  # first repetition warms up
                if rep == 1:
                    ev0.record(st)
                for f in frames:
                    if batched:
                        for g in range(G):
                            lo, hi = bounds[g], bounds[g + 1]
                            span = (hi - 1 - lo) * cap + int(counts[lo:hi, f].max())
                            cabi.memcpy_h2d_async(tgt[lo].data_ptr(), host[f - f_host, lo].ctypes.data, span * 16, st.cuda_stream)
                    else:
                        for s_ in range(S):
                            cabi.memcpy_h2d_async(tgt[s_].data_ptr(), host[f - f_host, s_].ctypes.data, int(counts[s_, f]) * 16, st.cuda_stream)
                    if rep == 1:
                        nbytes += int(counts[:, f].sum()) * 16
            ev1.record(st)
        st.synchronize()
        return nbytes / (ev0.elapsed_time(ev1) * 1e-3) / 1e9

    h2d_peak_scan = h2d_copy_only(False)
    h2d_peak_batch = h2d_copy_only(True)
    h2d_peak = max(h2d_peak_scan, h2d_peak_batch)
    timed_counts = counts[:, f_dev + W: f_dev + W + K]
    host_counts = counts[:, f_host + W: f_host + W + K]

    roof = None
    kern_table = None
    if KP:
        # stream group 0 alone runs KP more frames with an event after every kernel, so kernel durations are not mixed with other streams' work
        g0 = bs[0].seqs[0]
        for f in range(f_prof, f_prof + 4):
            wait(submit_dev(bs, dev_p, dev_p_ring, f_prof, f, groups={0}))
        g0.profile(True)
        for f in range(f_prof + 4, f_prof + KP):
            wait(submit_dev(bs, dev_p, dev_p_ring, f_prof, f, groups={0}))
        kt = g0.profile_kernels()
        g0.profile(False)
        mean_counts = dict(
            n_scan=float(timed_counts.mean()), n_edge=float(np.mean([c["n_edge"] for c in cnt])), n_surf=float(np.mean([c["n_surf"] for c in cnt])),
            n_ds=float(np.mean([c["n_ds_edge"] + c["n_ds_surf"] for c in cnt])), n_map=float(np.mean([c["n_map_edge"] + c["n_map_surf"] for c in cnt])), seqs=bounds[1] - bounds[0])
        tot = sum(v[0] for v in kt.values()) or 1.0
        kern_table = sorted(([f"{p}/{k}", v[0], v[1]] for (p, k), v in kt.items()), key=lambda r: -r[1])
        (dp, dk), (dms, dn) = max(kt.items(), key=lambda kv: kv[1][0])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ab = algorithmic_bytes(dp, dk, mean_counts)
        dur_s = dms / max(dn, 1) * 1e-3
        ach = (ab / dur_s / 1e9) if ab else None
        traffic = None
        traffic_src = None
        issue = None
        try:  # DRAM bytes / warp instructions of the same kernel from a committed ncu --set full capture, scaled to this launch's sequence count
            tfile = "r2_traffic_dense.json" if args.workload == "hdl64_dense" else "r2_traffic.json"
            tt = json.load(open(os.path.join(ROOT, "profiles", tfile)))["kernels"]
            ent = tt.get({"k_sector_select": "k_sector_warp<16>", "k_new_xform": "k_new_cluster", "k_merge<count>": "k_merge<0>", "k_merge<emit>": "k_merge<1>"}.get(dk, dk))
            if ent:
                per = ent["dram_bytes_per_launch"]
                # k_voxel_cluster appears twice per frame on the radix path: the scan job (~100 k points per sequence) is the one with the
                # larger traffic, the map job the other; everything else: first launch
                li = (int(np.argmin(per)) if dp == "map_update" else int(np.argmax(per))) if dk == "k_voxel_cluster" else 0
                scale = (bounds[1] - bounds[0]) / ent["sequences"]
                pick = (lambda v: float(np.mean(v))) if dk == "k_knn_cell_assoc" else (lambda v: v[li])  # the search runs twice per frame (unseeded, seeded): the timing averages both
                traffic = pick(per) * scale
                traffic_src = f"committed ncu --set full capture profiles/{tfile} (dram__bytes_read.sum + dram__bytes_write.sum per launch), scaled to this launch's sequence count; NOT measured in this run"
                winst = pick(ent["warp_instructions_per_launch"]) * scale
                sm_hz = float(clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0) * 1e6
                issue_peak = 148 * 4 * sm_hz  # warp instructions per second: 148 SMs x 4 schedulers x 1 instruction per clock
                issue = dict(warp_instructions_per_launch=winst, achieved_ginst_s=winst / dur_s / 1e9, peak_ginst_s=issue_peak / 1e9, frac=winst / dur_s / issue_peak,
                             issue_active_pct_under_ncu=pick(ent["issue_active_pct"]), threads_per_warp_instruction=pick(ent["threads_per_warp_instruction"]),
                             source=f"instruction count from profiles/{tfile} (same launch geometry), duration measured in this run",
                             note="the instruction-issue roofline next to the HBM one: this kernel works on L2-resident data, so what bounds it is how many warp instructions the SMs it occupies can issue (and the latency between them), not DRAM bandwidth")
        except Exception:
            traffic = None
        roof = dict(bound="hbm", kernel=f"{dp}/{dk}", achieved=ach, peak=peak, unit="GB/s", frac=(ach / peak) if ach else None, traffic=traffic,
                    traffic_source=traffic_src, issue=issue,
                    peak_source="MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    share_of_step=dms / tot, launches_timed=dn, avg_launch_us=dur_s * 1e6, algorithmic_bytes_per_launch=ab,
                    note="event-to-event interval (includes the launch gap) on steady-state frames; per-frame working set of this workload is L2-resident, so the path is latency- / issue-bound, not HBM-bound (DESIGN.md §6)")
    close_all(bs)
    same = None  # (the two passes run different frames of the same sequences; bit-equality of the two entry points is a test, tests/test_gpu_sequence.py)

    # ---- single-sequence latency: what the ROS drop-in sees (one scan in, one pose out, blocking), GPU vs the CPU path on one core ----
    latency = None
    if rank == 0 and world == 1 and not args.no_latency:
        try:
            from oracle import orc
            L0, L1 = 100, 200  # frames [L0, L1) are timed, after L0 untimed frames that fill the map
            sq = make_sequence(w, L1, 9000)
            xs = []
            pin = cabi.host_alloc(L1 * cap * 16).view(np.float32).reshape(L1, cap, 4)
            for f in range(L1):
                x, r = sq[f]
                pin[f, : x.shape[0]] = x
                xs.append((pin[f, : x.shape[0]], r if use_ring else None))
            g1 = cabi.Odometry(cfg, device=local_rank)
            for f in range(L0):
                g1.process_scan(*xs[f])
            t0 = time.perf_counter()
            for f in range(L0, L1):
                g1.process_scan(*xs[f])
            gpu_ms = (time.perf_counter() - t0) / (L1 - L0) * 1e3
            g1.close()
            o1 = orc.Odometry(orc.config(n_scan=w["n_scan"], n_rings=w["n_rings"], **w.get("cfg_kw", {})))
            for f in range(L0):
                o1.process_scan(*xs[f])
            t0 = time.perf_counter()
            for f in range(L0, L1):
                o1.process_scan(*xs[f])
            cpu_ms = (time.perf_counter() - t0) / (L1 - L0) * 1e3
            latency = dict(gpu_ms_per_frame=gpu_ms, cpu_one_core_ms_per_frame=cpu_ms, ratio=cpu_ms / gpu_ms, frames=f"{L0}..{L1 - 1}",
                           note="ONE sequence, blocking vilf_process_scan per frame from pinned host memory (H2D + all stages + D2H pose, wall clock) against the CPU "
                                "oracle on one core on the same frames (BASELINE.md 3.2a); this is the latency the ROS node sees, the headline is throughput over many sequences")
            cabi.host_free(pin.reshape(-1).view(np.uint8))
        except Exception as e:
            latency = dict(error=repr(e))

    # ---- the stages the north star names, at a size where HBM matters (BASELINE configs[2]: ~1e6 map points) ----
    large = None
    if rank == 0 and not args.no_sweep:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import roofline_sweep as rs
            rng = np.random.default_rng(7)
            g = cabi.Odometry(cabi.default_config(max_scan_points=300000, max_map_points=(1 << 20) + 1024, flags=args.flags), device=local_rank)
            mp = rs.make_map(1_000_000, rng)
            nq = 260_000
            q = mp[rng.integers(0, mp.shape[0], nq)].copy()
            q[:, :3] += rng.normal(0, 0.03, (nq, 3)).astype(np.float32)
            k = g.bench_stage(0, mp, q, leaf=0.2, iters=5)  # 0.24 m point spacing: the search structure a 0.2 m voxel-filtered map gets
            u = g.bench_stage(1, mp, leaf=0.4, iters=5)
            # the per-frame map maintenance at this size: a 0.2 m-filtered map of ~1e6 points + the 20 k new points of one scan
            nn = 20_000
            newp = mp[rng.integers(0, mp.shape[0], nn)].copy()
            newp[:, :3] += rng.normal(0, 0.15, (nn, 3)).astype(np.float32)
            u2 = g.bench_stage(2, mp, newp, leaf=0.2, iters=5)
            # the caller-side stage next to the path (DESIGN.md §0 row f): lidar depth for 150 visual features on an HDL-64 scan
            from oracle import orc
            x0 = np.ascontiguousarray(host[W, 0, : counts[0, f_host + W]])
            T = np.eye(4); T[:3, :3] = [[0, -1, 0], [0, 0, -1], [1, 0, 0]]
            fts = np.stack([rng.uniform(-1.2, 1.2, 150), rng.uniform(-0.4, 0.12, 150), np.ones(150)], 1).astype(np.float32)
            g.feature_extract(x0)
            g.feature_depth(fts, T_lidar_cam=T)
            t0 = time.perf_counter()
            for _ in range(50):
                dg, _, ncl = g.feature_depth(fts, T_lidar_cam=T)
            gpu_ms = (time.perf_counter() - t0) / 50 * 1e3
            t0 = time.perf_counter()
            do, _ = orc.feature_depth(orc.camera_cloud(x0, T), fts)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            depth_assoc = dict(features=150, cloud_points=int(ncl), gpu_ms_per_call=gpu_ms, cpu_port_ms=cpu_ms, identical_to_cpu_port=bool(np.array_equal(dg, do)),
                               note="vilf_feature_depth on the resident scan: wall time of the blocking call (1.9 KB up, 0.6 KB down); cpu = oracle restatement of NODE:54-140, :348-361, one core")
            g.close()
            pk = rs.peak
            issue_peak = 148 * 4 * float(clk.get("sm_max_mhz") or 1965.0) * 1e6  # warp instructions per second: 4 schedulers per SM, one per clock
            large = dict(map_points=int(mp.shape[0]), queries=nq, peak_gbs=pk,
                         knn_build=dict(ms=k[0], gbs=32.0 * mp.shape[0] / (k[0] * 1e-3) / 1e9, frac=32.0 * mp.shape[0] / (k[0] * 1e-3) / 1e9 / pk,
                                        note="search structure of an arbitrary (unsorted) cloud: first frame / state import only on the cell-ordered path"),
                         knn_query=dict(ms=k[1], queries_per_s=nq / (k[1] * 1e-3), gbs=136.0 * nq / (k[1] * 1e-3) / 1e9, frac=136.0 * nq / (k[1] * 1e-3) / 1e9 / pk,
                                        shells=int(k[2]), cell_m=float(k[3]),
                                        issue_roofline=issue_roofline(nq, k[1], issue_peak)),
                         map_update=dict(ms=u[0], voxels_out=int(u[2]), points_per_s=mp.shape[0] / (u[0] * 1e-3),
                                         gbs=(16.0 * mp.shape[0] + 16.0 * u[2]) / (u[0] * 1e-3) / 1e9, frac=(16.0 * mp.shape[0] + 16.0 * u[2]) / (u[0] * 1e-3) / 1e9 / pk,
                                         note="crop box + voxel filter of an UNSORTED 1e6-point cloud (radix path): not what a frame does on the cell-ordered path"),
                         map_update_per_frame=dict(ms=u2[0], map_points_in=int(u2[2]), new_points=nn, map_points_out=int(u2[3]), build_ms_untimed=u2[1],
                                                   gbs=(16.0 * (u2[2] + nn) + 16.0 * u2[3]) / (u2[0] * 1e-3) / 1e9,
                                                   frac=(16.0 * (u2[2] + nn) + 16.0 * u2[3]) / (u2[0] * 1e-3) / 1e9 / pk,
                                                   note="createSubMap (EM:298-352) on a voxel-filtered map: append + crop box + voxel filter; the cell-ordered path does it as one merge update that also writes the search table (no separate knn_build per frame); with --flags 4 it is the radix filter over the concatenation and knn_build comes on top"),
                         depth_association=depth_assoc,
                         note="device-resident stage timings (vilf_bench_stage) on a synthetic 1e6-point map, algorithmic bytes: search build 32 B/point, query 136 B, map update 16 B/point in + 16 B/point out")
        except Exception as e:  # the headline numbers stand on their own
            large = dict(error=repr(e))

    # ---- max over ranks ----
    ms_dev_max, ms_host_max = replicas.max_over_ranks([ms_dev, ms_host], device="cuda")
    scans = replicas.job_scans(world, S, K)
    e2e_value = scans / (ms_host_max * 1e-3)
    h2d_gbs = float(host_counts.sum()) * 16 / (ms_host * 1e-3) / 1e9
    out = dict(
        metric="scans/s scan-to-map (HDL-64 synthetic)" if args.workload == "hdl64" else f"scans/s scan-to-map ({args.workload} synthetic)",
        value=scans / (ms_dev_max * 1e-3), unit="scans/s", n_gpus=world, steps=K, warmup=W, ms_per_step=ms_dev_max / K,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32 geometry / f64 solve", data="synthetic",
        config=dict(workload=w["desc"], sequences_per_gpu=S, scans_per_step=S * world, points_per_scan=int(timed_counts.mean()),
                    preroll_frames=P, timed_frames=f"{f_dev + W}..{f_dev + W + K - 1} (value), {f_host + W}..{f_host + W + K - 1} (e2e)",
                    mean_map_points=float(np.mean([c["n_map_edge"] + c["n_map_surf"] for c in cnt])), map_capacity=map_cap,
                    map_path="cell-ordered (merge update)" if (args.flags & 8 or (not args.flags & 4 and map_cap + max(cap, 1024) > (1 << 19))) else "radix-sorted (one cluster per cloud)",
                    frames_in_flight=D, l2="inputs of one step are distinct frames (S x 1.8 MB) and all K steps use fresh scans; no cache flush needed",
                    parallelism=f"{world} replica(s) x {S} independent sequences in {G} lock-step batch(es) on {G} stream(s), no collective",
                    value_is="scans resident in HBM when the timed region starts (device-to-device copy into the scan slot + all stages); the SURVEY 8d(i) metric "
                             "(H2D scan + stages + D2H pose through the C ABI) is e2e, which is the headline against the CPU arm"),
        e2e=dict(value=e2e_value, unit="scans/s", h2d_bytes_per_step=int(host_counts.mean() * 16 * S), d2h_bytes_per_step=int(S * 7 * 8),
                 ms_per_step=ms_host_max / K, h2d_gbs=h2d_gbs, h2d_copy_only_gbs=h2d_peak, h2d_copy_only_per_scan_gbs=h2d_peak_scan, h2d_copy_only_per_group_gbs=h2d_peak_batch,
                 e2e_efficiency=h2d_gbs / h2d_peak if h2d_peak else None,
                 staging="write-combined pinned host memory (VILF_HOST_WC)" if os.environ.get("VILF_HOST_WC") else "pinned host memory",
                 host_placement=numa,
                 note="the step moves S packed scans (16 B per point, the payload of pcl::PointXYZI) over PCIe, one strided copy per stream group (the scans share a "
                      "pinned slab); h2d_copy_only_*: the same buffers copied with the GPU idle, one copy per scan / one contiguous copy per group (useful bytes per "
                      "second); e2e_efficiency = h2d_gbs / the better of the two (1.0 = the run moves scans as fast as the link alone can)"),
        gpu_launches=int(launches), clocks=clk,
        wall_ms=dict(dev=wall_dev, host=wall_host), gen_s=t_gen,
    )
    if young:
        out["young_map"] = young
    if latency:
        out["latency"] = latency
    if large:
        out["large_map"] = large
    if roof:
        out["roofline"] = roof
        out["kernels_ms"] = kern_table[:14]
        n_frames = max(1, max((r[2] for r in kern_table if r[0].endswith("k_frame_reset")), default=1))
        per_stage = {}
        for name, ms_k, _n in kern_table:  # event-to-event kernel intervals of one stream group, summed per stage, per frame
            per_stage[name.split("/")[0]] = per_stage.get(name.split("/")[0], 0.0) + ms_k / n_frames
        out["stage_ms"] = per_stage
        try:  # 5-NN queries per second inside the frame = queries of the timed launches / their summed kernel intervals
            for name, ms_k, n_l in kern_table:
                if name.split("/")[-1] in ("k_knn_assoc", "k_knn_cell_assoc") and ms_k > 0:
                    q_per_launch = mean_counts["n_ds"] * mean_counts["seqs"]
                    out["knn_in_frame"] = dict(kernel=name, queries_per_launch=q_per_launch, launches=int(n_l), queries_per_s=q_per_launch * n_l / (ms_k * 1e-3),
                                               map_points_per_sequence=mean_counts["n_map"],
                                               note="voxel-filtered scan features searched against the live local maps, one stream group alone; the search structure is part of the map update (no separate build per frame)")
        except Exception:
            pass
    return out


def main():
    args = parse()
    # torchrun pins OMP_NUM_THREADS to 1; the synthetic ray caster (OpenMP, host side, untimed) may use this rank's share of the cores
    lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(lws, 1)))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        wl = WORKLOADS[args.workload]
        cpre = args.cpu_preroll if args.cpu_preroll is not None else (wl.get("preroll", 100) if "cpu_frames" in wl else 100)
        cb = cpu_baseline(args.workload, max(10, min(wl.get("cpu_frames", args.cpu_frames), args.steps + args.warmup)), preroll=cpre)
        w = WORKLOADS[args.workload]
        print(json.dumps(dict(
            impl="reference", metric="scans/s scan-to-map (HDL-64 synthetic)" if args.workload == "hdl64" else f"scans/s scan-to-map ({args.workload} synthetic)",
            value=cb["value"], unit="scans/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * cb["cores"] / cb["value"],
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32 geometry / f64 solve", data="synthetic",
            config=dict(workload=w["desc"], parallelism=f"{cb['cores']} host processes, one sequence each"),
            cpu_baseline=cb, e2e=dict(value=cb["value"], unit="scans/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
            note="reference ROS node cannot be built offline (no ROS/PCL/Ceres); this is the CPU restatement under oracle/ (DESIGN.md §2)")))
        return

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:  # before CUDA is initialised in this process (the workers are forked)
        try:
            wl = WORKLOADS[args.workload]
            cpre = args.cpu_preroll if args.cpu_preroll is not None else (wl.get("preroll", 100) if "cpu_frames" in wl else 100)
            cb = cpu_baseline(args.workload, wl.get("cpu_frames", args.cpu_frames), preroll=cpre)
        except Exception as e:  # the GPU numbers stand on their own
            cb = dict(error=repr(e))
    import torch
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_gpu(args, rank, world, local_rank)
    if rank == 0:
        if cb is not None:
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
