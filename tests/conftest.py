import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_builds():
    """Host-side helpers every test may need: the oracle port (+ the verbatim reference where
    /root/reference exists) and the synthetic scan generator.  libscgpu.so is built by __graft_entry__.build()."""
    from oracle import oracle as orc
    from sc_lego_loam_b200 import build
    orc.build(ref=True)
    build.build_scangen()
    yield


def golden(variant):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", f"golden_{variant}.npz"), allow_pickle=False)


def params_from_golden(g):
    from oracle.oracle import Params
    R, S, K, excl, period = (int(v) for v in g["params"])
    lh, mr, ratio, thres = (float(v) for v in g["params_f"])
    return Params(R=R, S=S, num_candidates=K, exclude_recent=excl, tree_period=period, lidar_height=lh,
                  max_radius=mr, search_ratio=ratio, dist_thres=thres)


def seq_scans(g):
    """Regenerates the sequential-run scans of a golden file; returns None when the bytes differ from the
    ones the fixture was made from (different libm) -- callers then skip."""
    import hashlib
    from sc_lego_loam_b200.synth import ScanGen
    n, places = int(g["seq_n"]), int(g["seq_places"])
    R = int(g["params"][0])
    gen = ScanGen("hdl64", seed=777, n_places=places, n_azim=int(g["seq_azim"]),
                  max_range=45.0 if R == 7 else 110.0)
    scans = [gen.scan(i, 3) for i in range(n)]
    h = hashlib.sha256()
    for s in scans:
        h.update(s.tobytes())
    if h.hexdigest() != str(g["seq_sha256"]):
        return None
    return scans


VARIANTS = ["default", "k50", "40x120", "full"]
