"""The CPU oracle against (a) the committed golden vectors, (b) the reference's own vendored kd-tree,
(c) independent numpy / scipy statements of the same mathematics.  No GPU."""
import numpy as np
import pytest


def test_kdtree_matches_reference_nanoflann_goldens(orc, golden):
    """Index-for-index equal to the reference's vendored nanoflann 1.3.2 (vectors written by oracle/_ref)."""
    idx, d2 = orc.knn(golden["knn_map"], golden["knn_q"])
    assert np.array_equal(idx, golden["ref_knn_idx"])
    assert np.array_equal(d2, golden["ref_knn_d2"])


def test_kdtree_matches_reference_nanoflann_live(orc, synth):
    if orc.ref_lib() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(1)
    for m in (1, 4, 17, 5000):
        mp = np.zeros((m, 4), np.float32)
        mp[:, :3] = rng.uniform(-20, 20, (m, 3))
        q = np.zeros((300, 4), np.float32)
        q[:, :3] = rng.uniform(-22, 22, (300, 3))
        a, b = orc.knn(mp, q), orc.ref_knn(mp, q)
        k = min(m, 5)  # nanoflann leaves the slots beyond the map size uninitialised
        assert np.array_equal(a[0][:, :k], b[0][:, :k]) and np.array_equal(a[1][:, :k], b[1][:, :k])


def test_kdtree_matches_flann_single_index_of_opencv(orc, synth):
    """PCL's KdTreeFLANN is FLANN's KDTreeSingleIndex (leaf 15, exact, sorted).  OpenCV vendors the FLANN sources
    (cv2.flann_Index, algorithm 4 = FLANN_INDEX_KDTREE_SINGLE): the oracle's tree against that library on a real local map —
    same squared fp32 distances bit for bit, same indices except on exact distance ties."""
    cv2 = pytest.importorskip("cv2")
    from conftest import check_knn
    o = orc.Odometry(orc.config())
    seq = synth.Sequence("hdl64", 4, seed=6)
    for i in range(4):
        o.process_scan(seq[i][0])
    mp = o.cloud(orc.MAP_SURF)
    q = o.cloud(orc.DS_SURF).copy()
    q[:, :3] += np.float32(0.05)
    index = cv2.flann_Index(np.ascontiguousarray(mp[:, :3]), dict(algorithm=4, leaf_max_size=15))
    fi, fd = index.knnSearch(np.ascontiguousarray(q[:, :3]), 5, params=dict(checks=-1, eps=0.0, sorted=True))
    oi, od = orc.knn(mp, q)
    assert len(mp) > 5000 and len(q) > 2000
    check_knn(fi.astype(np.int32), fd.astype(np.float32), oi, od, gate=np.float32(1e30))
    assert (fi == oi).mean() > 0.9999


def test_kdtree_is_exact_vs_bruteforce_and_scipy(orc):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(2)
    mp = np.zeros((3000, 4), np.float32)
    mp[:, :3] = rng.normal(0, 8, (3000, 3))
    q = np.zeros((200, 4), np.float32)
    q[:, :3] = rng.normal(0, 8, (200, 3))
    idx, d2 = orc.knn(mp, q)
    d = ((q[:, None, :3] - mp[None, :, :3]) ** 2)
    bf = (d[..., 0] + d[..., 1]) + d[..., 2]  # fp32, same association as L2_Simple
    order = np.argsort(bf, axis=1, kind="stable")[:, :5]
    assert np.array_equal(np.take_along_axis(bf, order, 1), d2)
    _, si = cKDTree(mp[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64), k=5)
    assert (np.sort(si, 1) == np.sort(idx, 1)).mean() > 0.999  # fp64 vs fp32 distance ties aside


def test_kdtree_canonical_tie_rule_vs_bruteforce(orc):
    """Config.knn_ties = 1 / orc.knn(canonical=True): equal fp32 distances (tie class T2) ordered by map index — the rule of
    the CUDA path — instead of FLANN's first-visited-wins.  Same distances as the FLANN-order search, indices == brute force
    sorted by (d^2, index) on a lattice full of exact ties."""
    rng = np.random.default_rng(0)
    mp = np.zeros((20000, 4), np.float32); mp[:, :3] = np.round(rng.uniform(-8, 8, (20000, 3)) * 4) / 4
    q = np.zeros((1500, 4), np.float32); q[:, :3] = np.round(rng.uniform(-8, 8, (1500, 3)) * 2) / 2
    i0, d0 = orc.knn(mp, q)
    i1, d1 = orc.knn(mp, q, canonical=True)
    assert np.array_equal(d0, d1) and (i0 != i1).any()
    for r in range(0, 1500, 5):
        d = q[r, :3][None] - mp[:, :3]
        bf = ((d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        o = np.lexsort((np.arange(len(bf)), bf))[:5]
        assert np.array_equal(o, i1[r]) and np.array_equal(bf[o], d1[r])


def test_extract_goldens(orc, golden):
    cfg = orc.config(n_scan=16, n_rings=16)
    e, es, s, ss = orc.extract(cfg, golden["scan0"])
    assert np.array_equal(es, golden["edge_src"]) and np.array_equal(ss, golden["surf_src"])
    assert np.array_equal(e, golden["edge"]) and np.array_equal(s, golden["surf"])


def numpy_extract(xyzi, n_scan=16, lidar_min=3.0, lidar_max=90.0, thr=0.1):
    """Independent (slow, pure numpy/python) statement of FE:54-220 for the 16-beam branch."""
    rings = [[] for _ in range(n_scan)]
    for i, p in enumerate(xyzi):
        d = float(np.sqrt(np.float32(p[0] * p[0] + p[1] * p[1])))
        if d < lidar_min or d > lidar_max:
            continue
        ang = np.arctan(float(p[2]) / d) * 180 / np.pi
        rid = int((ang + 15) / 2 + 0.5)
        if rid > 15 or rid < 0:
            continue
        rings[rid].append(i)
    edge, surf = [], []
    for idxs in rings:
        n = len(idxs)
        if n < 131:
            continue
        P = xyzi[idxs, :3].astype(np.float32)
        curv = []
        for j in range(5, n - 5):
            acc = P[j - 5].copy()
            for k in (-4, -3, -2, -1):
                acc = acc + P[j + k]
            acc = acc - np.float32(10) * P[j]
            for k in (1, 2, 3, 4, 5):
                acc = acc + P[j + k]
            a = acc.astype(np.float64)
            curv.append((a[0] * a[0] + a[1] * a[1] + a[2] * a[2], j))
        cs = n - 10
        ln = cs // 6
        for s in range(6):
            st, en = ln * s, ln * (s + 1) - 1
            if s == 5:
                en = cs - 1
            sub = sorted(curv[st:en])
            picked, cnt = set(), 0
            for v, ind in reversed(sub):
                if ind in picked:
                    continue
                if v <= thr:
                    break
                cnt += 1
                picked.add(ind)
                if cnt <= 20:
                    edge.append(idxs[ind])
                else:
                    break
                for k in range(1, 6):
                    g = (P[ind + k] - P[ind + k - 1]).astype(np.float64)
                    if g @ g > 0.05:
                        break
                    picked.add(ind + k)
                for k in range(1, 6):
                    g = (P[ind - k] - P[ind - k + 1]).astype(np.float64)
                    if g @ g > 0.05:
                        break
                    picked.add(ind - k)
            surf += [idxs[ind] for v, ind in sub if ind not in picked]
    return np.asarray(edge, np.int32), np.asarray(surf, np.int32)


def test_extract_drops_non_finite_returns(orc, golden):
    """FE:56-57 leaves NaN rows in the cloud; the x86 reference drops them at the scanID range test (int(NaN) == INT_MIN).
    Extraction of a scan with NaN / Inf rows == extraction of the scan with those rows deleted, indices remapped."""
    cfg = orc.config(n_scan=16, n_rings=16)
    x = golden["scan0"].copy()
    bad = np.arange(7, x.shape[0], 53)
    x[bad[0::3], 0] = np.nan
    x[bad[1::3], 2] = np.nan
    x[bad[2::3], 2] = np.inf
    e, es, s, ss = orc.extract(cfg, x)
    keep = np.setdiff1d(np.arange(x.shape[0]), bad)
    ce, ces, cs, css = orc.extract(cfg, x[keep])
    assert np.array_equal(keep[ces], es) and np.array_equal(keep[css], ss)
    assert np.isfinite(e).all() and np.isfinite(s).all()


def test_extract_vs_independent_numpy(orc, golden):
    cfg = orc.config(n_scan=16, n_rings=16)
    _, es, _, ss = orc.extract(cfg, golden["scan1"])
    ne, ns = numpy_extract(golden["scan1"])
    assert np.array_equal(es, ne) and np.array_equal(ss, ns)


def numpy_voxel(pts, leaf):
    """Independent statement of pcl::VoxelGrid::applyFilter (stable order inside a voxel)."""
    inv = np.float32(1.0) / np.float32(leaf)
    mn = pts[:, :3].min(0)
    mx = pts[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(pts[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(key, kind="stable")
    out = []
    i = 0
    ks = key[order]
    while i < len(order):
        j = i
        acc = np.zeros(4, np.float32)
        while j < len(order) and ks[j] == ks[i]:
            acc = acc + pts[order[j]]
            j += 1
        out.append(acc / np.float32(j - i))
        i = j
    return np.asarray(out, np.float32)


def test_voxel_grid_goldens_and_numpy(orc, golden):
    for leaf in (0.4, 0.8):
        v, ok = orc.voxel_grid(golden["surf"], leaf)
        assert ok and np.array_equal(v, golden[f"vox_surf_{leaf}"])
    v, _ = orc.voxel_grid(golden["edge"], 0.4)
    assert np.array_equal(v, golden["vox_edge_0.4"])
    assert np.array_equal(v, numpy_voxel(golden["edge"], 0.4))
    sub = golden["surf"][:1500]
    assert np.array_equal(orc.voxel_grid(sub, 0.8)[0], numpy_voxel(sub, 0.8))


def test_voxel_grid_properties(orc, golden):
    pts = golden["surf"]
    v, _ = orc.voxel_grid(pts, 0.8)
    inv = np.float32(1 / np.float32(0.8))
    cells = np.floor(v[:, :3] * inv)
    assert len(np.unique(cells, axis=0)) == len(v)          # at most one output per voxel
    assert np.array_equal(orc.voxel_grid(v, 0.8)[0], v)     # idempotent
    perm = np.random.default_rng(0).permutation(len(pts))
    v2, _ = orc.voxel_grid(pts[perm], 0.8)
    assert v2.shape == v.shape and np.abs(v2 - v).max() < 1e-5  # summation order only (tie class T3)
    assert orc.voxel_grid(np.zeros((0, 4), np.float32), 0.8)[0].shape == (0, 4)
    big = np.array([[0, 0, 0, 0], [3e3, 3e3, 3e3, 0]], np.float32)
    out, ok = orc.voxel_grid(big, 0.01)                     # PCL's int32 guard: output = input
    assert not ok and np.array_equal(out, big)


def test_crop_box_golden(orc, golden):
    c = golden["crop_center"]
    out = orc.crop_box(golden["surf"], c - 15.0, c + 15.0)
    assert np.array_equal(out, golden["crop_surf"])
    p = golden["surf"]
    lo, hi = (c - 15.0).astype(np.float32), (c + 15.0).astype(np.float32)
    keep = np.all((p[:, :3] >= lo) & (p[:, :3] <= hi), axis=1)
    assert np.array_equal(out, p[keep])


def test_eig3_and_lstsq_vs_numpy(orc):
    rng = np.random.default_rng(3)
    for _ in range(50):
        A = rng.normal(size=(5, 3)) * rng.uniform(0.01, 3)
        Cm = A.T @ A
        w, V = orc.eig3(Cm)
        wn = np.linalg.eigvalsh(Cm)
        assert np.allclose(w, wn, rtol=1e-12, atol=1e-14 * np.abs(wn).max())
        assert np.allclose(Cm @ V, V * w, atol=1e-12 * max(1.0, np.abs(wn).max()))
        Apts = rng.normal(size=(5, 3)) + np.array([3.0, -2.0, 1.0])
        n = orc.lstsq5x3(Apts, -np.ones(5))
        nn = np.linalg.lstsq(Apts, -np.ones(5), rcond=None)[0]
        assert np.allclose(n, nn, rtol=1e-9, atol=1e-12)


def test_sensitivity_alternates_agree_with_the_oracle_proper(orc, golden):
    """The switchable alternates of the third-party restatements (tools/sensitivity.py): Eigen 3.3.7's tridiagonal QR iteration
    against the cyclic Jacobi solver and numpy, the unpivoted Householder plane fit against the column-pivoted one, and the
    reciprocal / unstable-order voxel centroids against the canonical ones (equal to the last bit or two, same voxels)."""
    rng = np.random.default_rng(8)
    for t in range(300):
        P = rng.normal(size=(5, 3)) * rng.uniform(0.01, 2.0, 3)
        if t % 5 == 0:
            P[:, 2] = 0.3 * P[:, 0]  # rank-deficient neighbourhoods (collinear / coplanar points are the common case)
        Z = P - P.mean(0)
        Cm = Z.T @ Z
        w0, V0 = orc.eig3(Cm)
        w1, V1 = orc.eig3(Cm, alt=True)
        wn = np.linalg.eigvalsh(Cm)
        sc = max(np.abs(wn).max(), 1e-300)
        assert np.abs(w0 - w1).max() <= 1e-14 * sc and np.abs(w1 - wn).max() <= 1e-14 * sc
        assert np.allclose(Cm @ V1, V1 * w1, atol=1e-13 * sc) and np.allclose(V1.T @ V1, np.eye(3), atol=1e-14)
        if wn[2] > 1.5 * wn[1]:
            assert 1 - abs(V0[:, 2] @ V1[:, 2]) < 1e-13
        A = rng.normal(size=(5, 3)) * 0.2 + np.array([30.0, -12.0, 1.5])
        assert np.allclose(orc.lstsq5x3(A, -np.ones(5)), orc.lstsq5x3(A, -np.ones(5), alt=True), rtol=1e-9, atol=1e-13)
    pts = golden["voxel_in"] if "voxel_in" in golden.files else None
    if pts is None:
        pts = (rng.uniform(-20, 20, (20000, 4))).astype(np.float32)
    a, _ = orc.voxel_grid(pts, 0.4, 0)
    for mode in (1, 2, 3):  # unstable order, reciprocal centroid, both
        b, _ = orc.voxel_grid(pts, 0.4, mode)
        assert a.shape == b.shape and np.abs(a - b).max() <= 4e-6  # same voxels, centroids equal to a few ulp of ~20 m


def test_se3_plus_and_jacobians(orc):
    rng = np.random.default_rng(4)
    x = np.array([0.1, -0.05, 0.02, 0.0, 1.0, 2.0, -0.5])
    x[3] = np.sqrt(1 - (x[:3] ** 2).sum())
    assert np.allclose(orc.se3_plus(x, np.zeros(6)), x, atol=1e-15)
    y = orc.se3_plus(x, np.array([0.01, -0.02, 0.03, 0.1, 0.2, -0.1]))
    assert abs(np.linalg.norm(y[:4]) - 1) < 1e-12
    pab = np.concatenate([rng.normal(size=3) * 5, [1.0, 2.0, 0.5], [1.05, 2.15, 0.6]])
    pnd = np.concatenate([rng.normal(size=3) * 5, [0.6, 0.0, 0.8], [0.7]])
    r0, J = orc.edge_eval(x, pab)
    s0, Js = orc.surf_eval(x, pnd)
    eps = 1e-6
    for k in range(6):  # left perturbation x (+) delta, the parameterisation of EM:34-49
        d = np.zeros(6); d[k] = eps
        xp, xm = orc.se3_plus(x, d), orc.se3_plus(x, -d)
        assert np.allclose((orc.edge_eval(xp, pab)[0] - orc.edge_eval(xm, pab)[0]) / (2 * eps), J[:, k], atol=1e-6)
        assert np.allclose((orc.surf_eval(xp, pnd)[0] - orc.surf_eval(xm, pnd)[0]) / (2 * eps), Js[:, k], atol=1e-6)


def test_factors_normal_equations_and_solve_goldens(orc, golden):
    cfg = orc.config(n_scan=16, n_rings=16)
    f = orc.factors(cfg, golden["fac_pose"], golden["f3_ds_edge"], golden["f3_ds_surf"], golden["f3_map_edge"], golden["f3_map_surf"])
    for k, v in f.items():
        assert np.array_equal(v, golden[f"fac_{k}"]), k
    pab, pnd = orc.pack_factors(golden["f3_ds_edge"], golden["f3_ds_surf"], f)
    H, g, cost = orc.normal_eq(cfg.huber, golden["fac_pose"], pab, pnd)
    assert np.array_equal(H, golden["ne_H"]) and np.array_equal(g, golden["ne_g"]) and cost == golden["ne_cost"][0]
    p, tr, term = orc.solve(cfg.huber, 4, golden["fac_pose"], pab, pnd)
    assert np.array_equal(p, golden["solve_pose"]) and np.array_equal(tr, golden["solve_trace"]) and term == golden["solve_term"][0]
    # the LM restatement must behave like a trust-region solver: monotone cost, gradient of the final point small vs the first
    costs = tr[tr[:, 2] == 1, 3]
    assert np.all(np.diff(costs) <= 0)


def test_solve_reaches_the_scipy_optimum(orc, golden):
    """Independent check of the restated Ceres loop: more iterations converge to scipy's robust least-squares optimum."""
    from scipy.optimize import least_squares
    cfg = orc.config(n_scan=16, n_rings=16)
    f = {k[4:]: golden[k] for k in golden.files if k.startswith("fac_") and k != "fac_pose"}
    pab, pnd = orc.pack_factors(golden["f3_ds_edge"], golden["f3_ds_surf"], f)
    x0 = golden["fac_pose"]
    p, _, _ = orc.solve(cfg.huber, 50, x0, pab, pnd)

    def resid(d):
        x = orc.se3_plus(x0, d)
        out = []
        for row in pab:
            r = orc.edge_eval(x, row)[0]
            s = float(r @ r)
            out.append(np.sqrt(s if s <= 0.01 else 2 * 0.1 * np.sqrt(s) - 0.01))
        for row in pnd:
            r = orc.surf_eval(x, row)[0]
            s = float(r @ r)
            out.append(np.sqrt(s if s <= 0.01 else 2 * 0.1 * np.sqrt(s) - 0.01))
        return np.asarray(out)

    sol = least_squares(resid, np.zeros(6), xtol=1e-12, ftol=1e-12, gtol=1e-12)
    xs = orc.se3_plus(x0, sol.x)
    c_orc = 0.5 * (resid_at(orc, p, x0, pab, pnd) ** 2).sum()
    c_sp = 0.5 * (resid(sol.x) ** 2).sum()
    assert c_orc <= c_sp * (1 + 1e-6)
    assert np.abs(xs - p).max() < 5e-4


def resid_at(orc, x, x0, pab, pnd):
    out = []
    for row in pab:
        r = orc.edge_eval(x, row)[0]
        s = float(r @ r)
        out.append(np.sqrt(s if s <= 0.01 else 2 * 0.1 * np.sqrt(s) - 0.01))
    for row in pnd:
        r = orc.surf_eval(x, row)[0]
        s = float(r @ r)
        out.append(np.sqrt(s if s <= 0.01 else 2 * 0.1 * np.sqrt(s) - 0.01))
    return np.asarray(out)


def test_sequence_goldens(orc, golden):
    cfg = orc.config(n_scan=16, n_rings=16)
    od = orc.Odometry(cfg)
    for i in range(golden["poses"].shape[0]):
        p, _, _ = od.process_scan(golden[f"scan{i}"])
        assert np.array_equal(p, golden["poses"][i]), i
    assert np.array_equal(od.cloud(orc.MAP_EDGE), golden["final_map_edge"])
    assert np.array_equal(od.cloud(orc.MAP_SURF), golden["final_map_surf"])


def test_map_too_small_skips_optimisation(orc):
    """EM:254: with <= 10 edge or <= 50 surf map points the solve is skipped but the map is still updated."""
    cfg = orc.config()
    od = orc.Odometry(cfg)
    rng = np.random.default_rng(0)
    e = np.zeros((5, 4), np.float32); e[:, :3] = rng.normal(0, 5, (5, 3))
    s = np.zeros((30, 4), np.float32); s[:, :3] = rng.normal(0, 5, (30, 3))
    od.init_map(e, s)
    p = od.update(e, s)
    assert np.array_equal(p, [0, 0, 0, 1, 0, 0, 0]) and len(od.solves()) == 0
    assert od.cloud(orc.MAP_EDGE).shape[0] > 0
