"""The host-side mirror of the reference's classes (include/vilf/*.hpp): it compiles with a plain C++14 compiler against
the C ABI, keeps the reference's method surface, refuses to run without a GPU, and — on a GPU — reproduces the oracle's
poses when driven in the order the reference's ROS node uses (feature_tracker_node.cpp:339-389)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import pose_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
INC = os.path.join(ROOT, "include", "vilf")


def build_shim():
    subprocess.run(["make", "-C", CPP], check=True, capture_output=True)
    return os.path.join(CPP, "shim_test")


def write_scans(path, scans):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(scans)))
        for x in scans:
            x = np.ascontiguousarray(x, dtype=np.float32)
            f.write(struct.pack("<i", x.shape[0]))
            f.write(x.tobytes())


def test_mirror_keeps_the_reference_method_surface():
    """Every public method / member name of the reference's two classes and the F-LOAM aliases."""
    fe = open(os.path.join(INC, "featureExtraction.hpp")).read()
    em = open(os.path.join(INC, "EstimationMapping.hpp")).read()
    for name in ("class featureExtraction", "void initParam(", "void extractFeature(", "class LaserProcessingClass", "void featureExtraction("):
        assert name in fe, name
    for name in ("class EstimationMapping", "void initParameter(", "void allocateMemory(", "void localMapInited(", "void optimation_processing(",
                 "int EdgeCostFactor(", "int SurfCostFactor(", "void createSubMap(", "void pointAssociaToMap(", "void getMapCloud(",
                 "double parameter_opti[7]", "Isometry3d globalOdom;", "Isometry3d globalOdom_last;", "CloudPtr localMapEdge;", "CloudPtr localMapSurf;",
                 "CloudPtr cloudRegistered;", "CloudPtr cloudNoRegistered;", "double edgeMapLeafSize;", "double surfMapLeafSize;",
                 "class OdomEstimationClass", "void initMapWithPoints(", "void updatePointsToMap(", "int addEdgeCostFactor(", "int addSurfCostFactor(",
                 "void addPointsToMap(", "void pointAssociateToMap(", "void downSamplingToMap(", "void getMap("):
        assert name in em, name
    sc = open(os.path.join(INC, "Scancontext.hpp")).read()
    for name in ("class SCManager", "void makeAndSaveScancontextAndKeys(", "std::pair<int, float> detectLoopClosureID(", "void setSCdistThres(", "void setMaximumRadius(",
                 "distanceBtnScanContext(", "double LIDAR_HEIGHT", "int PC_NUM_RING", "int PC_NUM_SECTOR", "double PC_MAX_RADIUS", "int NUM_EXCLUDE_RECENT",
                 "int NUM_CANDIDATES_FROM_TREE", "double SEARCH_RATIO", "double SC_DIST_THRES", "int TREE_MAKING_PERIOD_"):
        assert name in sc, name
    fx = open(os.path.join(INC, "featureExtract.hpp")).read()
    for name in ("class featureExtract", "void initParam(", "bool extractFeature(", '"Horizon_SCAN"', '"N_SCAN"', '"downsampleRate"', '"lidarMinRange"', '"lidarMaxRange"',
                 '"edgeThreshold"', '"surfThreshold"', '"SurfLeafSize"', "struct VelodynePointXYZIRT"):
        assert name in fx, name
    # the mirror goes through the C ABI only: no CUDA, torch or oracle in the headers
    for txt in (fe, em, sc, fx, open(os.path.join(INC, "session.hpp")).read(), open(os.path.join(INC, "cloud.hpp")).read()):
        assert not re.search(r"cuda_runtime|torch|oracle", txt)


def test_mirror_compiles_and_fails_loudly_without_gpu(tmp_path, synth):
    exe = build_shim()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_mirror_matches_oracle_in_node_order")
    seq = synth.Sequence("vlp16", 2, seed=1)
    write_scans(tmp_path / "scans.bin", [seq[i][0] for i in range(2)])
    r = subprocess.run([exe, str(tmp_path / "scans.bin"), "16", str(tmp_path / "poses.bin")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_mirror_matches_oracle_in_node_order(tmp_path, synth, orc):
    exe = build_shim()
    frames = 6
    seq = synth.Sequence("vlp32", frames, seed=4)
    scans = [seq[i][0] for i in range(frames)]
    write_scans(tmp_path / "scans.bin", scans)
    r = subprocess.run([exe, str(tmp_path / "scans.bin"), "32", str(tmp_path / "poses.bin"), "32"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "featureExtract + EstimationMapping" in r.stdout  # the ring-field extractor class in front of the same estimator
    poses = np.fromfile(tmp_path / "poses.bin", dtype=np.float64).reshape(frames, 7)
    o = orc.Odometry(orc.config(n_scan=32, n_rings=32))
    for i in range(frames):
        po, _, _ = o.process_scan(scans[i])
        e = pose_err(po, poses[i])
        assert e[0] <= 1e-4 and e[1] <= 1e-3, (i, e)
