"""CPU: oracle/voxel_oracle.cpp (the restatement of pcl::VoxelGrid; PARITY UNPINNED -- PCL is not in the image) against an
independent numpy statement of the same published algorithm: grid origin / extent, leaf indices and their order, points
per voxel bit-exact; centroids (FP32 running sums in std::sort order in the oracle) within n * ulp of the float64 mean."""
import numpy as np
import pytest


def _numpy_voxel_grid(pts, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    xyz = pts[:, :3]
    fin = np.isfinite(xyz).all(axis=1)
    xyz = xyz[fin]
    mn, mx = xyz.min(axis=0), xyz.max(axis=0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div_b = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int32)           # FP32 throughout, as PCL
    idx = (ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]).astype(np.uint32)
    uniq, inverse, count = np.unique(idx, return_inverse=True, return_counts=True)
    mean = np.zeros((len(uniq), 3), np.float64)
    np.add.at(mean, inverse, xyz.astype(np.float64))
    mean /= count[:, None]
    return min_b, div_b, uniq, count.astype(np.uint32), mean


@pytest.mark.parametrize("leaf", [0.5, 0.3, 1.0])
def test_voxel_oracle_matches_numpy_statement(leaf):
    from oracle import oracle as orc
    from sc_lego_loam_b200.synth import ScanGen
    vox = orc.Voxel()
    for i in (0, 3):
        s = ScanGen("hdl64", seed=11, n_places=40, n_azim=600).scan(i, 4)
        s[::211, 1] = np.nan
        got = vox.downsample(s, leaf)
        min_b, div_b, idx, count, mean = _numpy_voxel_grid(s, leaf)
        assert np.array_equal(got["min_b"], min_b) and np.array_equal(got["div_b"], div_b)
        assert np.array_equal(got["idx"], idx) and np.array_equal(got["count"], count)
        ulp = np.spacing(np.maximum(np.abs(mean), 1e-3).astype(np.float32)).astype(np.float64)
        assert (np.abs(got["points"][:, :3] - mean) <= count[:, None] * ulp + 1e-7).all()


def test_voxel_oracle_edges():
    from oracle import oracle as orc
    vox = orc.Voxel()
    assert len(vox.downsample(np.zeros((0, 4), np.float32), 0.5)["idx"]) == 0
    assert len(vox.downsample(np.full((5, 4), np.nan, np.float32), 0.5)["idx"]) == 0
    one = vox.downsample(np.array([[1.25, -2.5, 0.75, 9.0]], np.float32), 0.5)
    assert np.array_equal(one["points"], [[1.25, -2.5, 0.75, 9.0]]) and one["count"][0] == 1
    # intensity is averaged like the coordinates (AccumulatorIntensity)
    two = vox.downsample(np.array([[0.1, 0.1, 0.1, 2.0], [0.2, 0.2, 0.2, 4.0]], np.float32), 0.5)
    assert len(two["idx"]) == 1 and two["points"][0, 3] == 3.0
    far = np.zeros((2, 4), np.float32)
    far[0, :3], far[1, :3] = 1500.0, -1500.0
    with pytest.raises(ValueError):            # PCL: "Leaf size is too small for the input dataset"
        vox.downsample(far, 0.1)
