"""CPU: the submap-assembly oracle (mapOptmization.cpp:598-627, 928-949).  The restatement
(oracle/submap_shim.cpp, "port") is pinned against the reference's OWN transformPointCloud -- its text cut out of
mapOptmization.cpp at build time and compiled against stub types (oracle/_ref/libsubmapref.so) -- bit for bit."""
import numpy as np
import pytest

from oracle import oracle as orc


def icol(floats):
    """float index of the intensity: pcl::PointXYZI (32-byte records) keeps it at byte 16, packed float4 records at byte 12"""
    return 4 if floats == 8 else 3


def keyframes(seed=0, n_clouds=7, n=400, floats=8):
    rng = np.random.default_rng(seed)
    clouds, poses = [], []
    for i in range(n_clouds):
        c = np.zeros((n + 13 * i, floats), np.float32)
        c[:, :3] = rng.uniform(-60, 60, (len(c), 3)).astype(np.float32)
        c[:, 2] *= np.float32(0.1)
        if floats > 3:
            c[:, icol(floats)] = rng.uniform(-3, 200, len(c)).astype(np.float32)      # LeGO-LOAM stores row + col/10000 here; negatives are dropped
        clouds.append(c)
        poses.append(np.float32([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(-5, 5),
                                 rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), rng.uniform(-np.pi, np.pi)]))
    return clouds, np.stack(poses)


@pytest.mark.skipif(not orc.Submap.available("reference"), reason="oracle/_ref/libsubmapref.so not built (needs /root/reference)")
@pytest.mark.parametrize("floats", [3, 4, 8])
def test_port_equals_the_reference_function(floats):
    ref, port = orc.Submap("reference"), orc.Submap("port")
    clouds, poses = keyframes(seed=3, floats=floats)
    poses[0] = 0                                     # identity
    poses[1, 3:] = [np.pi / 2, -np.pi / 2, np.pi]    # axis-aligned rotations
    poses[2, :3] = [1e6, -1e6, 1e-6]                 # large / tiny translations
    clouds[3][:5, :3] = [[0, 0, 0], [np.inf, 1, 1], [np.nan, 0, 0], [1e30, 1e30, 1e30], [-0.0, 0.0, -0.0]]
    for c, p in zip(clouds, poses):
        a, b = ref.transform(c, p, icol(floats)), port.transform(c, p, icol(floats))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    v = np.float32([-2, -1, -0.999, -0.0, 0, 0.5, 3, 2.2e9, -2.2e9, np.nan, np.inf])
    assert np.array_equal(ref.keeps(v), port.keeps(v))
    assert list(port.keeps(np.float32([-1, -0.999, 0, 7]))) == [False, True, True, True]


def test_assemble_concatenates_in_order_and_filters():
    port = orc.Submap("port")
    clouds, poses = keyframes(seed=5, n_clouds=3, n=50, floats=4)
    out = port.assemble(clouds, poses)
    assert out.shape == (sum(len(c) for c in clouds), 4)
    assert np.array_equal(out[:50], port.transform(clouds[0], poses[0]))
    kept = port.assemble(clouds, poses, drop_negative_intensity=True)
    assert len(kept) == int(sum((c[:, 3] > -1).sum() for c in clouds)) and (kept[:, 3] > -1).all()
