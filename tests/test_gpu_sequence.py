"""End-to-end parity of the per-frame path (C ABI) against the CPU oracle: free-running and teacher-forced
sequences, every sensor layout, the lock-step batch, the asynchronous API, state export/import."""
import numpy as np
import pytest

from conftest import check_knn, pose_err

pytestmark = pytest.mark.gpu

TOL_ROT, TOL_TRANS, TOL_MAP = 1e-4, 1e-3, 1e-5  # BASELINE.json north_star tolerances


def run_pair(cabi, orc, synth, sensor, n_scan, n_rings, frames, seed, cap_scan, use_ring=False):
    seq = synth.Sequence(sensor, frames, seed=seed)
    o = orc.Odometry(orc.config(n_scan=n_scan, n_rings=n_rings))
    g = cabi.Odometry(cabi.default_config(n_scan=n_scan, n_rings=n_rings, max_scan_points=cap_scan, max_map_points=1 << 19))
    worst = (0.0, 0.0)
    for i in range(frames):
        x, r = seq[i]
        po, _, _ = o.process_scan(x, r if use_ring else None)
        pg = g.process_scan(x, r if use_ring else None)
        e = pose_err(po, pg)
        worst = (max(worst[0], e[0]), max(worst[1], e[1]))
        assert e[0] <= TOL_ROT and e[1] <= TOL_TRANS, (i, e)
    return o, g, worst


def maps_close(orc, cabi, o, g):
    for which in (0, 1):
        mo, mg = o.cloud(which), g.cloud(which)
        assert mo.shape == mg.shape
        assert np.abs(mo - mg).max() <= TOL_MAP


def test_hdl64_100_frames_free_running(cabi, orc, synth):
    """configs[0]: 100-frame HDL-64E sequence, GPU free-running against the CPU oracle."""
    o, g, worst = run_pair(cabi, orc, synth, "hdl64", 64, 64, 100, 0, 116000)
    maps_close(orc, cabi, o, g)
    so, sg = o.solves(), g.solves()
    assert np.array_equal(so[:, :4], sg[:, :4])  # same factor counts, termination and iteration counts
    c = g.counts()
    assert c["frames"] == 99 and c["status"] == 0
    assert np.array_equal(o.cloud(orc.NO_REGISTERED), g.cloud(cabi.NO_REGISTERED))
    assert np.abs(o.cloud(orc.REGISTERED) - g.cloud(cabi.REGISTERED)).max() <= TOL_MAP
    g.close()


def test_vlp32_and_16_and_128(cabi, orc, synth):
    for sensor, n_scan, n_rings, cap, ring in (("vlp32", 32, 32, 58000, False), ("vlp16", 16, 16, 10000, False), ("beams128", 0, 128, 263000, True)):
        o, g, _ = run_pair(cabi, orc, synth, sensor, n_scan, n_rings, 8, 5, cap, use_ring=ring)
        maps_close(orc, cabi, o, g)
        g.close()


def test_golden_sequence(cabi, golden):
    g = cabi.Odometry(cabi.default_config(n_scan=16, n_rings=16, max_scan_points=10000, max_map_points=1 << 16))
    for i in range(golden["poses"].shape[0]):
        p = g.process_scan(golden[f"scan{i}"])
        e = pose_err(p, golden["poses"][i])
        assert e[0] <= TOL_ROT and e[1] <= TOL_TRANS, (i, e)
    assert np.abs(g.cloud(0) - golden["final_map_edge"]).max() <= TOL_MAP
    assert np.abs(g.cloud(1) - golden["final_map_surf"]).max() <= TOL_MAP
    g.close()


def test_teacher_forced(cabi, orc, synth):
    """Every frame starts from the oracle's state (pose, previous pose, both maps): per-frame parity without feedback
    (the teacher-forced gate on every frame of the configs[0] sequence)."""
    frames = 100
    seq = synth.Sequence("hdl64", frames, seed=3)
    o = orc.Odometry(orc.config())
    g = cabi.Odometry(cabi.default_config(max_scan_points=116000, max_map_points=1 << 19))
    x, _ = seq[0]
    o.process_scan(x)
    for i in range(1, frames):
        g.set_state(o.state(), o.cloud(orc.MAP_EDGE), o.cloud(orc.MAP_SURF))
        x, _ = seq[i]
        po, _, _ = o.process_scan(x)
        g.feature_extract(x)
        pg = g.update()
        e = pose_err(po, pg)
        assert e[0] <= TOL_ROT and e[1] <= TOL_TRANS, (i, e)
        assert np.abs(o.cloud(orc.MAP_SURF) - g.cloud(1)).max() <= TOL_MAP
        assert np.abs(g.state() - o.state()).max() < 1e-6
    g.close()


def test_method_surface_equals_fused_path(cabi, synth):
    """extractFeature -> localMapInited / optimation_processing call by call == vilf_process_scan."""
    seq = synth.Sequence("vlp32", 6, seed=8)
    cfg = cabi.default_config(n_scan=32, n_rings=32, max_scan_points=58000, max_map_points=1 << 18)
    a, b, c = cabi.Odometry(cfg), cabi.Odometry(cfg), cabi.Odometry(cfg)
    for i in range(6):
        x, _ = seq[i]
        pa = a.process_scan(x)
        b.feature_extract(x)
        if i == 0:
            b.map_init()
            pb = b.pose()[0]
        else:
            pb = b.update()
        e, _ = b.features(0)
        s, _ = b.features(1)
        if i == 0:
            c.map_init(e, s)
            pc = c.pose()[0]
        else:
            pc = c.update(e, s)  # host clouds, like the reference's optimation_processing(edge, surf)
        assert np.array_equal(pa, pb) and np.array_equal(pa, pc), i
    for which in range(4):
        assert np.array_equal(a.cloud(which), b.cloud(which)) and np.array_equal(a.cloud(which), c.cloud(which))
    pose, rt = a.pose()
    R = rt[:9].reshape(3, 3)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.array_equal(rt[9:], pose[4:])
    for o in (a, b, c):
        o.close()


def test_batch_lockstep_equals_single(cabi, synth):
    S, frames = 3, 7
    seqs = [synth.Sequence("vlp32", frames, seed=20 + s) for s in range(S)]
    cfg = cabi.default_config(n_scan=32, n_rings=32, max_scan_points=58000, max_map_points=1 << 18)
    singles = [cabi.Odometry(cfg) for _ in range(S)]
    batch = cabi.Batch(cfg, S)
    for f in range(frames):
        scans = [np.ascontiguousarray(seqs[s][f][0]) for s in range(S)]
        poses = batch.wait(batch.submit(scans))
        for s in range(S):
            assert np.array_equal(poses[s], singles[s].process_scan(scans[s])), (f, s)
    for s in range(S):
        for which in (0, 1):
            assert np.array_equal(batch.seqs[s].cloud(which), singles[s].cloud(which))
        singles[s].close()
    batch.close()


def test_batch_strided_host_array_equals_separate_scans(cabi, synth):
    """Scans of a batch that lie in ONE host array with a row pitch of max_scan_points points are moved with one strided copy
    (vilf_api.cu: submit_common); rows are ragged (every sequence has its own point count, one is empty in one frame).  Same poses
    and maps as the same scans submitted from separate arrays; explicit ring ids take the same route."""
    S, frames, cap = 5, 6, 58000
    seqs = [synth.Sequence("vlp32", frames, seed=40 + s) for s in range(S)]
    cfg = cabi.default_config(n_scan=32, n_rings=32, max_scan_points=cap, max_map_points=1 << 18)
    a, b = cabi.Batch(cfg, S), cabi.Batch(cfg, S)
    slab = cabi.host_alloc(frames * S * cap * 16).view(np.float32).reshape(frames, S, cap, 4)
    slab[:] = np.nan  # what lies behind a scan's last point must never matter
    for f in range(frames):
        scans = [np.ascontiguousarray(seqs[s][f][0][: (None if s != 2 else 40000 + 1000 * f)]) for s in range(S)]
        if f == 3:
            scans[4] = scans[4][:0]
        for s in range(S):
            slab[f, s, : len(scans[s])] = scans[s]
        pa = a.wait(a.submit([slab[f, s, : len(scans[s])] for s in range(S)]))
        pb = b.wait(b.submit([scans[s].copy() for s in range(S)]))
        assert np.array_equal(pa, pb), f
    for s in range(S):
        for which in (0, 1, 5):
            assert np.array_equal(a.seqs[s].cloud(which), b.seqs[s].cloud(which))
    a.close(); b.close()
    # explicit rings (n_scan == 0): the ring ids travel as a second strided copy
    seq = synth.Sequence("beams128", 3, seed=1)
    cap = 263000
    cfg = cabi.default_config(n_scan=0, n_rings=128, max_scan_points=cap, max_map_points=1 << 18, max_ring_points=2048 + 64)
    a, b = cabi.Batch(cfg, 3), cabi.Batch(cfg, 3)
    slab = cabi.host_alloc(3 * cap * 16).view(np.float32).reshape(3, cap, 4)
    rslab = cabi.host_alloc(3 * cap * 2).view(np.uint16).reshape(3, cap)
    for f in range(3):
        x, r = seq[f]
        ns = [len(x), len(x) - 5000, len(x) - 123]
        for s in range(3):
            slab[s, : ns[s]] = x[: ns[s]]; rslab[s, : ns[s]] = r[: ns[s]]
        pa = a.wait(a.submit([slab[s, : ns[s]] for s in range(3)], [rslab[s, : ns[s]] for s in range(3)]))
        pb = b.wait(b.submit([np.ascontiguousarray(x[: ns[s]]) for s in range(3)], [np.ascontiguousarray(r[: ns[s]]) for s in range(3)]))
        assert np.array_equal(pa, pb), f
    a.close(); b.close()
    cabi.host_free(slab); cabi.host_free(rslab)


def test_async_pipeline_equals_blocking(cabi, synth):
    frames = 12
    seq = synth.Sequence("vlp32", frames, seed=31)
    scans = [np.ascontiguousarray(seq[i][0]) for i in range(frames)]
    cfg = cabi.default_config(n_scan=32, n_rings=32, max_scan_points=58000, max_map_points=1 << 18)
    a, b = cabi.Odometry(cfg), cabi.Odometry(cfg)
    ref = [a.process_scan(x) for x in scans]
    tickets, got = [], []
    for x in scans:
        tickets.append(b.submit_scan(x))
        if len(tickets) == 4:
            got.append(b.wait(tickets.pop(0)))
    while tickets:
        got.append(b.wait(tickets.pop(0)))
    assert all(np.array_equal(r, p) for r, p in zip(ref, got))
    # more than 8 frames in flight is refused, not silently dropped
    c = cabi.Odometry(cfg)
    ts = [c.submit_scan(x) for x in scans[:8]]
    with pytest.raises(cabi.VilfError) as e:
        c.submit_scan(scans[8])
    assert e.value.code == 5
    for t in ts:
        c.wait(t)
    for o in (a, b, c):
        o.close()


def test_state_roundtrip_and_sequence_errors(cabi, synth):
    seq = synth.Sequence("vlp32", 5, seed=40)
    cfg = cabi.default_config(n_scan=32, n_rings=32, max_scan_points=58000, max_map_points=1 << 18)
    a = cabi.Odometry(cfg)
    with pytest.raises(cabi.VilfError) as e:
        a.update()
    assert e.value.code == 5
    for i in range(3):
        a.process_scan(seq[i][0])
    b = cabi.Odometry(cfg)
    b.set_state(a.state(), a.cloud(0), a.cloud(1))
    for i in range(3, 5):
        assert np.array_equal(a.process_scan(seq[i][0]), b.process_scan(seq[i][0]))
    a.close(); b.close()


def test_map_too_small_skips_optimisation(cabi, orc):
    """EM:254 guard: the pose stays at the prediction, the map is still maintained."""
    rng = np.random.default_rng(0)
    e = np.zeros((5, 4), np.float32); e[:, :3] = rng.normal(0, 5, (5, 3))
    s = np.zeros((30, 4), np.float32); s[:, :3] = rng.normal(0, 5, (30, 3))
    g = cabi.Odometry(cabi.default_config(max_scan_points=2048, max_map_points=4096))
    o = orc.Odometry(orc.config())
    g.map_init(e, s); o.init_map(e, s)
    pg, po = g.update(e, s), o.update(e, s)
    assert np.array_equal(pg, po) and len(g.solves()) == 0
    assert np.array_equal(g.cloud(0), o.cloud(orc.MAP_EDGE)) and np.array_equal(g.cloud(1), o.cloud(orc.MAP_SURF))
    g.close()


def test_map_capacity_overflow_is_reported(cabi, synth):
    seq = synth.Sequence("vlp32", 3, seed=41)
    g = cabi.Odometry(cabi.default_config(n_scan=32, n_rings=32, max_scan_points=58000, max_map_points=1024))
    g.process_scan(seq[0][0])
    with pytest.raises(cabi.VilfError) as e:
        g.process_scan(seq[1][0])
    assert e.value.code == 3
    g.close()


def test_cluster_and_grid_wide_paths_are_bit_identical(cabi, synth):
    """One-cluster-per-cloud kernels (k_cluster.cu) vs the grid-wide multi-launch path: same poses, maps and stage outputs, bit for bit."""
    frames = 10
    seq = synth.Sequence("hdl64", frames, seed=11)
    kw = dict(max_scan_points=116000, max_map_points=1 << 18)
    a = cabi.Odometry(cabi.default_config(**kw))
    b = cabi.Odometry(cabi.default_config(flags=cabi.FLAG_NO_CLUSTER, **kw))
    for i in range(frames):
        x, _ = seq[i]
        pa, pb = a.process_scan(x), b.process_scan(x)
        assert np.array_equal(pa, pb), i
        for which in (cabi.DS_EDGE, cabi.DS_SURF, cabi.MAP_EDGE, cabi.MAP_SURF, cabi.REGISTERED):
            assert np.array_equal(a.cloud(which), b.cloud(which)), (i, which)
    # stage-level entry points take the cluster path for small clouds as well
    rng = np.random.default_rng(5)
    pts = (rng.random((50000, 4), dtype=np.float32) - 0.5) * np.float32(60.0)
    for leaf in (0.4, 0.8, 0.05):
        va, ga = a.voxel_downsample(pts, leaf)
        vb, gb = b.voxel_downsample(pts, leaf)
        assert ga == gb and np.array_equal(va, vb)
    ca = a.crop_voxel_downsample(pts, [1.0, -2.0, 0.5], 20.0, 0.4)
    cb = b.crop_voxel_downsample(pts, [1.0, -2.0, 0.5], 20.0, 0.4)
    assert np.array_equal(ca, cb)
    q = pts[:2000]
    ia, da = a.knn5(pts, q)
    ib, db = b.knn5(pts, q)
    assert np.array_equal(ia, ib) and np.array_equal(da, db)
    a.close(); b.close()


@pytest.mark.parametrize("sensor,n_scan,cap", [("vlp32", 32, 58000), ("hdl64", 64, 116000)])
def test_1000_frames_free_running(cabi, orc, synth, sensor, n_scan, cap):
    """BASELINE.json configs[1] / configs[2] and the north-star's parity clause: 1000-frame synthetic sequences, the GPU
    free-running (never reset to the oracle's state), every frame within 1e-4 rad / 1e-3 m of the CPU oracle, final maps
    within 1e-5 m.  Frames are submitted ahead of the waits (3 in flight) like bench.py does."""
    frames = 1000
    seq = synth.Sequence(sensor, frames, seed=21)
    o = orc.Odometry(orc.config(n_scan=n_scan, n_rings=n_scan))
    g = cabi.Odometry(cabi.default_config(n_scan=n_scan, n_rings=n_scan, max_scan_points=cap, max_map_points=1 << 18, max_ring_points=1864))
    pend = []
    worst = [0.0, 0.0]

    def check(i, x, pg):
        po, _, _ = o.process_scan(x)
        e = pose_err(po, pg)
        worst[0], worst[1] = max(worst[0], e[0]), max(worst[1], e[1])
        assert e[0] <= TOL_ROT and e[1] <= TOL_TRANS, (i, e)

    for i in range(frames):
        x = np.ascontiguousarray(seq[i][0])
        pend.append((i, x, g.submit_scan(x)))
        if len(pend) >= 3:
            j, xj, t = pend.pop(0)
            check(j, xj, g.wait(t))
    for j, xj, t in pend:
        check(j, xj, g.wait(t))
    maps_close(orc, cabi, o, g)
    c = g.counts()
    assert c["frames"] == frames - 1 and c["status"] == 0
    # the trajectory really moved: ~1 m per frame
    assert np.linalg.norm(g.pose()[0][4:]) > 500.0
    g.close()


def test_graph_replay_equals_individual_launches(cabi, synth):
    """Steady-state frames are replayed as captured CUDA graphs (one per scan-buffer / map-buffer combination); the same
    kernels launched one by one (VILF_FLAG_NO_GRAPH) must give identical poses, maps and launch counts."""
    frames = 9
    seq = synth.Sequence("hdl64", frames, seed=17)
    kw = dict(max_scan_points=116000, max_map_points=1 << 18)
    a = cabi.Odometry(cabi.default_config(**kw))
    b = cabi.Odometry(cabi.default_config(flags=cabi.FLAG_NO_GRAPH, **kw))
    for i in range(frames):
        x, _ = seq[i]
        assert np.array_equal(a.process_scan(x), b.process_scan(x)), i
    for which in (cabi.MAP_EDGE, cabi.MAP_SURF, cabi.DS_SURF):
        assert np.array_equal(a.cloud(which), b.cloud(which))
    assert a.launch_count() == b.launch_count()
    a.close(); b.close()


def test_config5_128_beams_fine_voxels(cabi, orc, synth):
    """BASELINE.json configs[4]: dense 128-beam scans (explicit ring ids) with a 0.2 m edge / 0.4 m surf map voxel.  The edge
    map then gets 0.5 m search cells and the two-shell 5-NN walk with early exit, the surf map keeps 1 m cells — both kinds
    of grid are queried by the same warps of k_knn_assoc.  Free-running against the oracle, maps compared at the end."""
    frames = 8
    seq = synth.Sequence("beams128", frames, seed=31)
    o = orc.Odometry(orc.config(n_scan=0, n_rings=128, edge_leaf=0.2, surf_leaf=0.4))
    g = cabi.Odometry(cabi.default_config(n_scan=0, n_rings=128, edge_leaf=0.2, surf_leaf=0.4, max_scan_points=263000, max_map_points=1 << 18,
                                          max_ring_points=2048 + 64))
    for i in range(frames):
        x, r = seq[i]
        po, _, _ = o.process_scan(x, r)
        pg = g.process_scan(x, r)
        e = pose_err(po, pg)
        assert e[0] <= TOL_ROT and e[1] <= TOL_TRANS, (i, e)
    maps_close(orc, cabi, o, g)
    assert np.array_equal(o.solves()[:, :4], g.solves()[:, :4])  # same factor counts, terminations and iteration counts
    c = g.counts()
    assert c["status"] == 0 and c["n_map_edge"] > 5000
    g.close()


def test_config3_dense_hdl64_million_point_maps(cabi, orc, synth):
    """BASELINE.json configs[2]: a dense-world HDL-64E sequence whose live local maps (EM:327-350 crop + voxel filter, EM:256-257
    rebuild of the search structure) hold ~1e6 points; the large-map path (maps beyond one cluster's 2^19 points) is the one that
    runs.  Free-running against the oracle on every frame; then the three large-map stages on the oracle's own ~1e6 map points.

    Tie class T2 at this scale: 4.4e4 queries per outer iteration meet an exact fp32 distance tie at the 5th / 6th neighbour
    about every 50 frames (first at frame 46 of this sequence: indices 197469 vs 203676, d^2 = 0.015903158 both).  FLANN keeps
    the first visited, the CUDA path the lower map index; the poses then split by ~2e-8 m and, a few frames later, single
    points fall on the other side of a voxel face (seen with 0.1 m voxels; the workload now uses 0.09 m, vil_fusion_b200/synth.py DENSE).  So the strict comparison (every frame, maps to 1e-5 m, identical
    solver summaries) runs against the oracle with the CANONICAL tie rule (oracle Config.knn_ties = 1, checked against brute
    force in tests/test_oracle.py); against the FLANN-order oracle the poses are held to the north-star tolerance."""
    D = synth.DENSE
    frames, flann_frames = 300, 80
    seq = synth.Sequence(D["sensor"], frames, seed=7, density=D["density"], speed=D["speed"])
    o = orc.Odometry(orc.config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"], knn_ties=1))
    of = orc.Odometry(orc.config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"]))
    g = cabi.Odometry(cabi.default_config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"], max_scan_points=116000, max_map_points=D["max_map_points"],
                                          max_ring_points=1864))
    worst = [0.0, 0.0]
    for i in range(frames):
        x = np.ascontiguousarray(seq[i][0])
        po, _, _ = o.process_scan(x)
        pg = g.process_scan(x)
        e = pose_err(po, pg)
        worst = [max(worst[0], e[0]), max(worst[1], e[1])]
        assert e[0] <= TOL_ROT and e[1] <= TOL_TRANS, (i, e)
        if i < flann_frames:
            pf, _, _ = of.process_scan(x)
            ef = pose_err(pf, pg)
            assert ef[0] <= TOL_ROT and ef[1] <= TOL_TRANS, (i, ef)
        if i % 50 == 49:  # maps along the way, not only at the end
            maps_close(orc, cabi, o, g)
            assert np.array_equal(o.solves()[:, :4], g.solves()[:, :4]), i
    assert worst[0] < 1e-9 and worst[1] < 1e-9, worst  # with one tie rule the two free-running trajectories do not separate
    maps_close(orc, cabi, o, g)
    assert np.array_equal(o.solves()[:, :4], g.solves()[:, :4])
    c = g.counts()
    assert c["status"] == 0 and c["frames"] == frames - 1
    assert c["n_map_edge"] + c["n_map_surf"] >= 1_000_000, c
    assert max(c["n_map_edge"], c["n_map_surf"]) > (1 << 19), c
    # ---- stage level, on the ~1e6 real map points (all of them, and all queries) ----
    me, ms = o.cloud(orc.MAP_EDGE), o.cloud(orc.MAP_SURF)
    big = np.ascontiguousarray(np.concatenate([me, ms]))
    assert big.shape[0] >= 1_000_000
    for leaf in (0.2, 0.4):
        vo, _ = orc.voxel_grid(big, leaf)
        vg, guard = g.voxel_downsample(big, leaf)
        assert guard == 0 and np.array_equal(vg, vo), leaf
    ctr = [float(v) for v in g.pose()[0][4:]]
    half = 60.0
    co, _ = orc.voxel_grid(orc.crop_box(big, [ctr[a] - half for a in range(3)], [ctr[a] + half for a in range(3)]), 0.2)
    cg = g.crop_voxel_downsample(big, ctr, half, 0.2)
    assert np.array_equal(cg, co)
    rng = np.random.default_rng(3)
    q = ms[rng.integers(0, ms.shape[0], 200_000)].copy()
    q[:, :3] += rng.normal(0, 0.05, (q.shape[0], 3)).astype(np.float32)
    io, do = orc.knn(ms, q, 5, canonical=True)
    ig, dg = g.knn5(ms, q)
    inside = do < np.float32(1.0)  # exact wherever the reference uses the result (EM:129 / :189 gate)
    assert inside[:, 4].mean() > 0.9
    assert np.array_equal(dg[inside], do[inside]) and np.array_equal(ig[inside], io[inside])
    i2, d2 = orc.knn(ms, q, 5)  # FLANN visiting order: same distances, indices equal up to exact ties
    check_knn(ig, dg, i2, d2)
    g.close()
