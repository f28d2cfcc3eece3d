import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure), compiled on first use."""
    from oracle import orc as o
    o.build()
    return o


@pytest.fixture(scope="session")
def synth():
    from vil_fusion_b200 import synth as s
    s.build_native()
    return s


@pytest.fixture(scope="session")
def cabi():
    from vil_fusion_b200 import cabi as c
    return c


@pytest.fixture(scope="session")
def golden():
    """Committed golden vectors (tests/golden/make_golden.py wrote them in the build container)."""
    return np.load(os.path.join(GOLDEN, "small16.npz"))


@pytest.fixture(scope="session")
def golden2():
    """Committed golden vectors of the round-2 stages (tests/golden/make_golden_r2.py)."""
    return np.load(os.path.join(GOLDEN, "round2.npz"))


@pytest.fixture(scope="session")
def hdl64_frames(synth):
    seq = synth.Sequence("hdl64", 6, seed=0)
    return [seq[i] for i in range(6)]


def pose_err(a, b):
    """(rotation angle between the two orientations [rad], translation difference [m]); pose = qx qy qz qw tx ty tz."""
    qa, qb = np.asarray(a[:4], float), np.asarray(b[:4], float)
    qa, qb = qa / np.linalg.norm(qa), qb / np.linalg.norm(qb)
    chord = min(np.linalg.norm(qa - qb), np.linalg.norm(qa + qb))  # = 2 sin(angle / 4); stable near zero
    ang = 4.0 * np.arcsin(min(1.0, chord / 2.0))
    return float(ang), float(np.linalg.norm(np.asarray(a[4:], float) - np.asarray(b[4:], float)))


def check_knn(gi, gd, oi, od, gate=1.0):
    inside = od < gate
    assert np.array_equal(gd[inside], od[inside])
    neq = inside & (gi != oi)
    if neq.any():  # only documented distance ties (T2) may pick a different index
        rows = np.nonzero(neq.any(axis=1))[0]
        for r in rows:
            for k in np.nonzero(neq[r])[0]:
                assert (od[r] == od[r, k]).sum() >= 2 or gd[r, k] == od[r, k], (r, k)
