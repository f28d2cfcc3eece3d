"""Lidar depth for visual features (the caller-side stage next to the odometry path): the CUDA entry point
vilf_feature_depth against the oracle's restatement of getFeatureDepth (feature_tracker_node.cpp:54-140, :348-361)."""
import numpy as np
import pytest

T_LC = np.eye(4)
T_LC[:3, :3] = [[0, -1, 0], [0, 0, -1], [1, 0, 0]]  # camera z = lidar x (the comment at NODE:150)
T_LC[:3, 3] = [0.05, -0.1, 0.02]


def features(rng, m):
    return np.stack([rng.uniform(-1.2, 1.2, m), rng.uniform(-0.4, 0.12, m), np.ones(m)], 1).astype(np.float32)


def test_oracle_depth_of_a_wall(orc):
    """A wall 10 m ahead of the camera, sampled like a lidar: every feature that looks at it gets depth ~10 m (camera z)."""
    rng = np.random.default_rng(0)
    az = np.deg2rad(np.arange(-60, 60, 0.2))
    el = np.deg2rad(np.arange(-20, 10, 0.4))
    A, E = np.meshgrid(az, el)
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    pts = d * (10.0 / d[:, :1])  # lidar frame: wall at x = 10
    scan = np.concatenate([pts, np.zeros((pts.shape[0], 1))], 1).astype(np.float32)
    T = np.eye(4); T[:3, :3] = T_LC[:3, :3]
    cloud = orc.camera_cloud(scan, T)
    assert cloud.shape[0] == scan.shape[0] and np.allclose(cloud[:, 2], 10.0, atol=1e-4)
    f = np.stack([rng.uniform(-0.8, 0.8, 200), rng.uniform(-0.12, 0.3, 200), np.ones(200)], 1).astype(np.float32)
    depth, nn = orc.feature_depth(cloud, f)
    assert (depth > 0).mean() > 0.9
    assert np.abs(depth[depth > 0] - 10.0).max() < 0.05
    # too few points: nothing is returned (NODE:91-95)
    depth, _ = orc.feature_depth(cloud[:9], f)
    assert (depth == -1).all()


def test_oracle_camera_cloud_filter(orc):
    scan = np.array([[1, 0, 0, 0.5], [-1, 0, 0, 0.5], [1, 10.0, 0, 0.1], [1, 10.5, 0, 0.1], [1, 0, -10.5, 0.1], [0, 1, 1, 0.2]], np.float32)
    out = orc.camera_cloud(scan, np.eye(4))
    assert np.array_equal(out, scan[[0, 2]])  # x > 0, |y/x| <= 10, |z/x| <= 10 (NODE:353)


@pytest.mark.gpu
def test_feature_depth_matches_oracle(cabi, orc, synth):
    rng = np.random.default_rng(3)
    seq = synth.Sequence("hdl64", 3, seed=8)
    g = cabi.Odometry(cabi.default_config(max_scan_points=116000, max_map_points=1 << 17))
    for i in range(3):
        x, _ = seq[i]
        f = features(rng, 180)
        cloud = orc.camera_cloud(x, T_LC)
        od, onn = orc.feature_depth(cloud, f)
        assert (od > 0).sum() > 30
        # (a) explicit camera-frame cloud
        gd, gnn, cnt = g.feature_depth(f, cloud_cam=cloud)
        assert cnt == cloud.shape[0]
        assert np.array_equal(gd, od)
        assert np.array_equal(gnn, onn)
        # (b) the scan resident from the odometry step, filter + extrinsic on the device
        g.process_scan(x)
        gd2, gnn2, cnt2 = g.feature_depth(f, T_lidar_cam=T_LC)
        assert cnt2 == cloud.shape[0]
        assert np.array_equal(gd2, od)
        # neighbours are reported as scan indices: they must be the same points
        keep = (x[:, 0] > 0) & (np.abs(x[:, 1] / x[:, 0]) <= 10) & (np.abs(x[:, 2] / x[:, 0]) <= 10)
        scan_idx = np.nonzero(keep)[0]
        assert np.array_equal(gnn2, scan_idx[onn])
    # edge cases: no features, fewer than 10 points, a feature far from every return
    gd, _, _ = g.feature_depth(np.zeros((0, 3), np.float32), cloud_cam=cloud)
    assert gd.shape == (0,)
    gd, _, _ = g.feature_depth(f, cloud_cam=cloud[:9])
    assert (gd == -1).all()
    up = np.array([[0.0, -50.0, 1.0]], np.float32)  # looks straight up: no lidar return within the angular gate
    assert g.feature_depth(up, cloud_cam=cloud)[0][0] == -1 and orc.feature_depth(cloud, up)[0][0] == -1
    g.close()
