"""Sensitivity envelope of the "parity unpinned" third-party restatements (What's weak #1).

The reference's PCL / Eigen / Ceres cannot be built here, so the oracle restates them.  This tool bounds what a misreading of
those libraries' ARITHMETIC could cost: every third-party piece gets a switchable alternate inside the oracle
(oracle/orc_pipeline.hpp Config, oracle/orc_math.hpp), and the 1000-frame HDL-64 sequence is run

  teacher-forced   before every frame the alternate is reset to the baseline's state (pose, globalOdom, both maps), both process
                   the same features: counts the decisions that flip within ONE frame (lambda_2 > 3 lambda_1 EM:153, planeValid
                   EM:202-213, LM termination / iteration schedule EM:283) and the pose / map difference one frame can create;
  free-running     the alternate runs alone from frame 0: pose difference to the baseline on every frame and the final maps.

Alternates (reference call site):
  eig_qr        EM:150   Eigen 3.3.7 SelfAdjointEigenSolver::compute (tridiagonal QR iteration) instead of cyclic Jacobi
  plane_nopivot EM:198   HouseholderQR (no column pivoting) instead of colPivHouseholderQr
  centroid_rcp  EM:248-251, :347-350   voxel centroid `sum * (1 / n)` (Eigen 3.2 operator/=) instead of `sum / n`
  voxel_unstable  same   std::sort (unstable, what PCL calls) instead of the stable order inside a voxel
  lm_cholesky   EM:283   normal equations + Cholesky instead of Householder QR of [J; D] (the CUDA path's choice)
  knn_canonical EM:128, :185   (d^2, index) tie order instead of FLANN's visiting order (the CUDA path's choice)
  all           every switch at once

    python tools/sensitivity.py [frames] [hdl64|vlp32] > profiles/r2_sensitivity.json       (CPU only, ~7 min on 8 cores)
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ALTS = {
    "eig_qr": dict(eig_alg=1),
    "plane_nopivot": dict(plane_alg=1),
    "centroid_rcp": dict(centroid_div=1),
    "voxel_unstable": dict(voxel_order=1),
    "lm_cholesky": dict(lm_solver=1),
    "knn_canonical": dict(knn_ties=1),
    "all": dict(eig_alg=1, plane_alg=1, centroid_div=1, voxel_order=1, lm_solver=1, knn_ties=1),
}
TOL = dict(rot_rad=1e-4, trans_m=1e-3, map_m=1e-5)
SENSOR, SEED = "hdl64", 21
N_SCAN = {"hdl64": 64, "vlp32": 32}


def pose_err(a, b):
    qa, qb = a[:4] / np.linalg.norm(a[:4]), b[:4] / np.linalg.norm(b[:4])
    chord = min(np.linalg.norm(qa - qb), np.linalg.norm(qa + qb))
    return 4.0 * float(np.arcsin(min(1.0, chord / 2.0))), float(np.linalg.norm(a[4:] - b[4:]))


def map_delta(a, b):
    """(points that have no bit-identical partner, largest distance from a point of a to its nearest point of b)"""
    if a.shape == b.shape and np.array_equal(a, b):
        return 0, 0.0
    from scipy.spatial import cKDTree
    sa = set(map(bytes, np.ascontiguousarray(a[:, :3])))
    sb = set(map(bytes, np.ascontiguousarray(b[:, :3])))
    only = len(sa ^ sb)
    d = max(float(cKDTree(b[:, :3]).query(a[:, :3])[0].max()), float(cKDTree(a[:, :3]).query(b[:, :3])[0].max())) if len(a) and len(b) else float("inf")
    return only, d


def features(frames):
    from oracle import orc
    from vil_fusion_b200 import synth
    seq = synth.Sequence(SENSOR, frames, seed=SEED)
    cfg = orc.config(n_scan=N_SCAN[SENSOR], n_rings=N_SCAN[SENSOR])
    out, gt = [], []
    for i in range(frames):
        e, _, s, _ = orc.extract(cfg, np.ascontiguousarray(seq[i][0]))
        out.append((e, s))
        gt.append(seq.gt_pose(i))
    return out, gt


def run_free(args):
    name, frames = args
    from oracle import orc
    feats = FEATS
    o = orc.Odometry(orc.config(n_scan=N_SCAN[SENSOR], n_rings=N_SCAN[SENSOR], **ALTS.get(name, {})))
    poses = np.zeros((frames, 7))
    solves = []
    for i, (e, s) in enumerate(feats[:frames]):
        if i == 0:
            o.init_map(e, s); poses[0] = [0, 0, 0, 1, 0, 0, 0]
        else:
            poses[i] = o.update(e, s)
        solves.append(o.solves()[:, :4].copy())
    return name, poses, solves, o.cloud(0), o.cloud(1)


def run_forced(args):
    """baseline B and alternate A, A re-seeded from B before every frame"""
    name, frames = args
    from oracle import orc
    feats = FEATS
    cb, ca = orc.config(n_scan=N_SCAN[SENSOR], n_rings=N_SCAN[SENSOR]), orc.config(n_scan=N_SCAN[SENSOR], n_rings=N_SCAN[SENSOR], **ALTS[name])
    B, A = orc.Odometry(cb), orc.Odometry(ca)
    r = dict(frames=0, edge_decisions=0, surf_decisions=0, edge_flips=0, surf_flips=0, knn_sets_differ=0, solve_summaries_differ=0, frames_with_any_flip=0,
             max_rot_rad=0.0, max_trans_m=0.0, map_points_not_identical=0, max_map_dist_m=0.0, frames_map_size_differs=0)
    for i, (e, s) in enumerate(feats[:frames]):
        if i == 0:
            B.init_map(e, s)
            continue
        st, me, ms = B.state(), B.cloud(0), B.cloud(1)
        A.set_state(st); A.set_cloud(0, me); A.set_cloud(1, ms)
        pb = B.update(e, s)
        pa = A.update(e, s)
        er = pose_err(pb, pa)
        r["max_rot_rad"] = max(r["max_rot_rad"], er[0]); r["max_trans_m"] = max(r["max_trans_m"], er[1])
        sb, sa = B.solves()[:, :4], A.solves()[:, :4]
        if sb.shape != sa.shape or not np.array_equal(sb, sa):
            r["solve_summaries_differ"] += 1
        # the per-feature decisions of one association pass, same inputs for both: the frame's features at the baseline's final
        # pose against the maps the frame started from
        de, ds = B.cloud(orc.DS_EDGE), B.cloud(orc.DS_SURF)
        fb = orc.factors(cb, pb, de, ds, me, ms)
        fa = orc.factors(ca, pb, de, ds, me, ms)
        ef = int(np.count_nonzero(fb["edge_valid"] != fa["edge_valid"])); sf = int(np.count_nonzero(fb["surf_valid"] != fa["surf_valid"]))
        r["edge_decisions"] += len(de); r["surf_decisions"] += len(ds)
        r["edge_flips"] += ef; r["surf_flips"] += sf
        kd = int(np.count_nonzero((fb["edge_nn"] != fa["edge_nn"]).any(axis=1))) + int(np.count_nonzero((fb["surf_nn"] != fa["surf_nn"]).any(axis=1)))
        r["knn_sets_differ"] += kd
        r["frames_with_any_flip"] += int(ef + sf > 0)
        for w in (0, 1):
            mb, ma = B.cloud(w), A.cloud(w)
            if mb.shape != ma.shape:
                r["frames_map_size_differs"] += 1
            only, d = map_delta(ma, mb)
            r["map_points_not_identical"] += only
            r["max_map_dist_m"] = max(r["max_map_dist_m"], d)
        r["frames"] += 1
    return name, r


FEATS = None


def main():
    global FEATS, SENSOR
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    if len(sys.argv) > 2:
        SENSOR = sys.argv[2]
    t0 = time.time()
    FEATS, gt = features(frames)  # forked workers inherit FEATS
    ctx = mp.get_context("fork")
    with ctx.Pool(min(8, os.cpu_count() or 1)) as pool:
        free = {n: (p, s, me, ms) for n, p, s, me, ms in pool.map(run_free, [(n, frames) for n in ["baseline"] + list(ALTS)])}
        forced = dict(pool.map(run_forced, [(n, frames) for n in ALTS]))
    bp, bs, bme, bms = free["baseline"]
    R0, t0g = np.asarray(gt[0][0]), np.asarray(gt[0][1])
    gt_rel = np.array([R0.T @ (np.asarray(gt[i][1]) - t0g) for i in range(frames)])  # ground-truth position in the frame of scan 0

    def gt_err(p):
        e = np.linalg.norm(p[:, 4:] - gt_rel, axis=1)
        return dict(max_m=float(e.max()), final_m=float(e[-1]))

    out = dict(sequence=dict(sensor=SENSOR, seed=SEED, frames=frames, travelled_m=float(np.linalg.norm(bp[-1][4:]))), tolerance=TOL,
               baseline_error_vs_ground_truth=gt_err(bp), alternates={})
    for n in ALTS:
        p, s, me, ms = free[n]
        errs = np.array([pose_err(bp[i], p[i]) for i in range(frames)])
        first_split = next((i for i in range(frames) if not np.array_equal(bp[i], p[i])), None)
        sched = sum(1 for i in range(frames) if bs[i].shape != s[i].shape or not np.array_equal(bs[i], s[i]))
        oe, de = map_delta(me, bme)
        os_, ds_ = map_delta(ms, bms)
        fr = dict(error_vs_ground_truth=gt_err(p), max_rot_rad=float(errs[:, 0].max()), max_trans_m=float(errs[:, 1].max()), first_frame_with_different_pose=first_split,
                  frames_with_different_solve_summary=sched,
                  per_100_frames=[dict(frames=[k, min(k + 99, frames - 1)], max_rot_rad=float(errs[k:k + 100, 0].max()), max_trans_m=float(errs[k:k + 100, 1].max()))
                                  for k in range(0, frames, 100)],
                  final_maps=dict(edge=dict(points=int(len(me)), baseline_points=int(len(bme)), not_identical=oe, max_dist_m=de),
                                  surf=dict(points=int(len(ms)), baseline_points=int(len(bms)), not_identical=os_, max_dist_m=ds_)))
        inside = fr["max_rot_rad"] <= TOL["rot_rad"] and fr["max_trans_m"] <= TOL["trans_m"]
        maps_inside = max(de, ds_) <= TOL["map_m"] and len(me) == len(bme) and len(ms) == len(bms)
        out["alternates"][n] = dict(switches=ALTS[n], teacher_forced=forced[n], free_running=fr, pose_inside_tolerance=bool(inside), final_maps_inside_tolerance=bool(maps_inside))
    out["seconds"] = time.time() - t0
    out["note"] = ("teacher_forced: flips / differences created within single frames from identical inputs; free_running: accumulated over the whole sequence. "
                   "A map point 'not identical' is one whose xyz bits have no partner in the other map; a flipped voxel membership shows up as a point a leaf away, "
                   "so max_dist_m above 1e-5 with a handful of not-identical points is tie class T5 (a razor-edge decision), not arithmetic drift. "
                   "error_vs_ground_truth: distance of the estimated position to the simulator's true position, for scale: the spread between alternates has to be read against it.")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
