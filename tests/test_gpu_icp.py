"""GPU: loop verification after the path (mapOptmization.cpp:1053-1078): scgpu_verify_loop against the
restated pcl::IterativeClosestPoint (oracle/icp_oracle.cpp).  PARITY UNPINNED -- PCL is absent from the reference tree and
from this image, both sides restate its published algorithm; tolerances cover the two independent eigen-solvers."""
import numpy as np
import pytest

from test_icp_oracle import clouds

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,noise", [(3, 0.0), (4, 0.02), (8, 0.02)])
def test_verify_loop_equals_icp_oracle(k, noise):
    from oracle import oracle as orc
    from sc_lego_loam_b200.scgpu import SCManager
    src3, tgt3, T = clouds(n_src=1500, n_tgt=20000, noise=noise)
    src, tgt = np.zeros((len(src3), k), np.float32), np.zeros((len(tgt3), k), np.float32)
    src[:, :3], tgt[:, :3] = src3, tgt3
    m = SCManager()
    got = m.verify_loop(src, tgt)
    want = orc.Icp().align(src, tgt)
    assert got["converged"] == want["converged"] and got["accepted"]
    assert abs(got["iterations"] - want["iterations"]) <= 1
    assert np.allclose(got["T"], want["T"], atol=1e-4), got["T"] - want["T"]
    assert abs(got["fitness"] - want["fitness"]) <= 1e-5 + 1e-3 * want["fitness"]
    assert np.allclose(got["T"], T, atol=5e-3 + 2 * noise)


def test_verify_loop_rejects_and_seeds():
    from sc_lego_loam_b200.scgpu import SCManager
    m = SCManager()
    src, tgt, T = clouds(n_src=1200, n_tgt=15000, yaw=0.9)          # a revisit rotated by ~52 degrees
    cold = m.verify_loop(src, tgt)                                  # identity initial guess (what the reference runs)
    warm = m.verify_loop(src, tgt, seed_axis=2, seed_angle=0.9)      # seeded with the yaw Scan Context reports
    assert warm["accepted"] and warm["fitness"] < 1e-3 and np.allclose(warm["T"], T, atol=5e-3)
    assert warm["fitness"] <= cold["fitness"] + 1e-9 and warm["iterations"] <= cold["iterations"]
    rng = np.random.default_rng(0)
    junk = rng.uniform(-30, 30, (1000, 3)).astype(np.float32) + np.float32([0, 0, 40])
    bad = m.verify_loop(junk, tgt)
    assert not bad["accepted"] and bad["fitness"] > 1.5               # historyKeyframeFitnessScore
    none = m.verify_loop(junk + np.float32(1000), tgt, max_correspondence_distance=1.0)
    assert not none["converged"] and not none["accepted"]
    # a device-list handle routes the call to its first shard
    g = SCManager(devices=[0, 0], capacity_hint=1024)
    assert g.verify_loop(src, tgt, seed_axis=2, seed_angle=0.9)["accepted"]
