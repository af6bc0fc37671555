"""CPU: the ICP oracle (oracle/icp_oracle.cpp, a restatement of pcl::IterativeClosestPoint -- parity unpinned, PCL is absent)
recovers a known rigid motion and reports PCL's fitness (mean squared nearest-neighbour distance)."""
import numpy as np


def clouds(n_src=600, n_tgt=4000, yaw=0.06, t=(0.3, -0.2, 0.05), seed=3, noise=0.0):
    rng = np.random.default_rng(seed)
    # a structured "submap": points on a few planes and poles, so that point-to-point ICP has something to lock on
    tgt = np.concatenate([
        np.c_[rng.uniform(-20, 20, n_tgt // 2), rng.uniform(-20, 20, n_tgt // 2), np.zeros(n_tgt // 2)],
        np.c_[np.full(n_tgt // 4, 8.0), rng.uniform(-20, 20, n_tgt // 4), rng.uniform(0, 4, n_tgt // 4)],
        np.c_[rng.uniform(-20, 20, n_tgt - n_tgt // 2 - n_tgt // 4), np.full(n_tgt - n_tgt // 2 - n_tgt // 4, -6.0), rng.uniform(0, 4, n_tgt - n_tgt // 2 - n_tgt // 4)],
    ]).astype(np.float32)
    pick = rng.choice(n_tgt, n_src, replace=False)
    c, s = np.cos(yaw), np.sin(yaw)
    Rm = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    src = ((tgt[pick].astype(np.float64) - np.array(t)) @ Rm).astype(np.float32)     # src = R^T (tgt - t)  ->  T(src) = tgt
    src += rng.normal(0, noise, src.shape).astype(np.float32)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = Rm, t
    return src, tgt, T


def test_icp_oracle_recovers_motion():
    from oracle import oracle as orc
    src, tgt, T = clouds()
    r = orc.Icp().align(src, tgt)
    assert r["converged"] and 1 <= r["iterations"] <= 100
    assert np.allclose(r["T"], T, atol=2e-3), r["T"] - T
    assert r["fitness"] < 1e-4
    far = orc.Icp().align(src + np.float32(500.0), tgt, max_corr=1.0)        # no correspondences within the gate
    assert not far["converged"] and far["iterations"] == 0
