"""GPU: voxel-grid downsample in front of the path (mapOptmization.cpp:264, 1235-1237, 1628-1630)
against oracle/voxel_oracle.cpp, the restatement of pcl::VoxelGrid (PARITY UNPINNED: PCL is not in the image).

Bit-exact: grid origin / extent, the set of occupied leaf indices, points per voxel.  Centroids: the device returns the
correctly rounded mean, PCL an FP32 running sum in std::sort order -- tolerance n_points_in_voxel * ulp(|coordinate|)
(written below).  Fused path: the descriptor built from the raw scan with the leaf set equals the descriptor the plain
path builds from the device's own downsampled cloud bit for bit (and the oracle's descriptor of that cloud); against the
full oracle chain (oracle voxel grid -> oracle descriptor) cell heights agree within 1e-5 relative except for at most a
handful of cells (a centroid that moved by an ulp across a bin boundary)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mgr():
    from sc_lego_loam_b200.scgpu import SCManager
    m = SCManager()
    yield m
    m.close()


def _scan(seed, i, floats=4, kind="hdl64"):
    from sc_lego_loam_b200.synth import ScanGen
    return ScanGen(kind, seed=seed, n_places=40).scan(i, floats)


def _compare(got, want, pts_abs_max):
    assert np.array_equal(got["min_b"], want["min_b"]) and np.array_equal(got["div_b"], want["div_b"])
    assert np.array_equal(got["idx"], want["idx"])                                 # same voxels, PCL's order after the sort
    assert np.array_equal(got["points"][:, 3].astype(np.uint32), want["count"])    # same number of points in each
    n = want["count"].astype(np.float64)[:, None]
    ulp = np.spacing(np.maximum(np.abs(want["points"][:, :3]), 1e-3).astype(np.float32)).astype(np.float64)
    err = np.abs(got["points"][:, :3].astype(np.float64) - want["points"][:, :3].astype(np.float64))
    tol = n * ulp + 1e-7                                                           # FP32 running sum of n terms vs exact mean
    assert (err <= tol).all(), float((err / tol).max())
    return float(err.max())


@pytest.mark.parametrize("leaf", [0.5, 0.3, 1.0])
@pytest.mark.parametrize("floats", [3, 4, 8])
def test_voxel_grid_matches_oracle(mgr, leaf, floats):
    from oracle import oracle as orc
    vox = orc.Voxel()
    for i in (0, 7):
        s = _scan(11, i, floats)
        want = vox.downsample(s, leaf)
        got = mgr.voxel_downsample(s, leaf)
        assert not got["refused"]
        _compare(got, want, np.abs(s[:, :3]).max())


def test_voxel_grid_edges(mgr):
    from oracle import oracle as orc
    vox = orc.Voxel()
    rng = np.random.default_rng(5)
    # non-finite points are skipped; negative coordinates; points exactly on leaf boundaries; duplicates
    s = rng.uniform(-30, 30, (5000, 4)).astype(np.float32)
    s[::97, 0] = np.nan
    s[5::101, 1] = np.inf
    s[10:200, :3] = np.round(s[10:200, :3] * 2) / 2          # multiples of the leaf size
    s[300:400] = s[300]                                       # 100 identical points
    want = vox.downsample(s, 0.5)
    got = mgr.voxel_downsample(s, 0.5)
    _compare(got, want, 30.0)
    # a single point, an empty cloud, a cloud of NaNs
    one = np.array([[1.25, -2.5, 0.75, 9.0]], np.float32)
    got = mgr.voxel_downsample(one, 0.5)
    assert len(got["idx"]) == 1 and np.array_equal(got["points"][0], [1.25, -2.5, 0.75, 1.0])
    assert len(mgr.voxel_downsample(np.zeros((0, 4), np.float32), 0.5)["idx"]) == 0
    assert len(mgr.voxel_downsample(np.full((64, 4), np.nan, np.float32), 0.5)["idx"]) == 0
    # PCL's refusal (leaf index would overflow int32): the input is passed through
    far = rng.uniform(-1, 1, (1000, 4)).astype(np.float32)
    far[0, :3] = [1500.0, 1500.0, 1500.0]      # 30,000 leaves per axis: 2.7e13 > INT32_MAX (and no int64 overflow in dx*dy*dz)
    far[1, :3] = [-1500.0, -1500.0, -1500.0]
    with pytest.raises(ValueError):
        vox.downsample(far, 0.1)
    got = mgr.voxel_downsample(far, 0.1)
    assert got["refused"] and len(got["idx"]) == 1000
    assert np.array_equal(got["points"][:, :3], far[:, :3])


def test_table_overflow_takes_more_passes(mgr):
    """More occupied voxels than the cluster's table holds (8 x 3,072 slots): the scan is redone in key partitions."""
    from oracle import oracle as orc
    rng = np.random.default_rng(9)
    s = rng.uniform(-100, 100, (120000, 4)).astype(np.float32)   # ~every point its own voxel
    s[:, 2] = rng.uniform(-5, 5, 120000)
    want = orc.Voxel().downsample(s, 0.5)
    got = mgr.voxel_downsample(s, 0.5)
    assert len(want["idx"]) > 8 * 3072 and got["passes"] > 1
    _compare(got, want, 100.0)


def test_fused_descriptor_equals_plain_path_on_downsampled_cloud():
    from oracle import oracle as orc
    from sc_lego_loam_b200.scgpu import SCManager
    port, vox = orc.Port(), orc.Voxel()
    raw, fused = SCManager(), SCManager()
    fused.set_downsample_leaf(0.5)
    cells_off = 0
    for i in range(6):
        s = _scan(23, i, 4)
        ds = raw.voxel_downsample(s, 0.5)["points"].copy()
        ds[:, 3] = 0
        a = fused.makeScancontext(s)                   # raw scan in, voxel grid + binning on the device
        b = raw.makeScancontext(ds)                    # plain path on the device's own centroids
        assert np.array_equal(a, b)
        assert np.array_equal(b, port.make_sc(ds))     # ... which is the oracle's descriptor of that cloud
        c = port.make_sc(vox.downsample(s, 0.5)["points"])   # the full oracle chain (PCL restated -> SC restated)
        cells_off += int((np.abs(a - c) > 1e-5 * np.maximum(1.0, np.abs(c))).sum())   # heights are centroid z: ulp-level differences expected
    assert cells_off <= 6, cells_off                   # a centroid that moved by an ulp across a bin boundary is rare
    # the fused path feeds the database: append + detect work as with plain scans
    scans = np.stack([_scan(23, i % 4, 4) for i in range(60)])
    fused.append_scans(scans)
    assert fused.size() == 60
    e = fused.get_entry(5)
    assert np.array_equal(np.asarray(e[0], np.float64).ravel(), fused.makeScancontext(scans[5]).ravel())
    raw.close()
    fused.close()
