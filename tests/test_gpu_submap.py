"""GPU: submap assembly in front of the ICP (mapOptmization.cpp:598-627, 924-949) through the C ABI:
scgpu_assemble_submap against the reference's own transformPointCloud (oracle/_ref/libsubmapref.so where it travelled, else
the pinned restatement) -- transformed coordinates BIT-EXACT, order and intensity filter as the reference -- and
scgpu_verify_loop_keyframes against the composed oracle (transform -> concatenate -> restated VoxelGrid -> restated ICP; the
last two PARITY UNPINNED: PCL is absent)."""
import numpy as np
import pytest

from test_icp_oracle import clouds as icp_clouds
from test_submap_oracle import icol, keyframes

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    """bit-identical floats; a NaN matches a NaN whatever its sign / payload (x86 and the GPU differ in the NaN they generate)"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def _oracle():
    from oracle import oracle as orc
    return orc.Submap("reference" if orc.Submap.available("reference") else "port")


@pytest.mark.parametrize("floats", [3, 4, 8])
def test_assembled_cloud_is_bit_identical(floats):
    from sc_lego_loam_b200.scgpu import SCManager
    o, m = _oracle(), SCManager()
    clouds, poses = keyframes(seed=11, n_clouds=9, n=3000, floats=floats)
    clouds[4] = clouds[4][:0]                                         # an empty keyframe cloud
    clouds[2][:4, :3] = [[0, 0, 0], [np.inf, 1, 1], [np.nan, 0, 0], [1e30, 1e30, 1e30]]
    ic = icol(floats)
    got = m.assemble_submap(clouds, poses, intensity_column=ic)
    want = o.assemble(clouds, poses, intensity_column=ic)
    assert same_bits(got, want) and np.isnan(got).any()
    if floats > 3:
        clouds[1][::7, ic] = -1.0
        clouds[1][1::7, ic] = -0.5
        clouds[3][2::5, ic] = np.nan
        clouds[5][3::5, ic] = 2.5e9
        got = m.assemble_submap(clouds, poses, drop_negative_intensity=True, intensity_column=ic)
        want = o.assemble(clouds, poses, drop_negative_intensity=True, intensity_column=ic)
        assert len(got) < sum(len(c) for c in clouds)
        assert same_bits(got, want)
    assert len(m.assemble_submap([], np.zeros((0, 6), np.float32))) == 0


def test_history_submap_goes_through_the_voxel_grid():
    from oracle import oracle as orc
    from scipy.spatial import cKDTree
    from sc_lego_loam_b200.scgpu import SCManager
    o, m = _oracle(), SCManager()
    clouds, poses = keyframes(seed=13, n_clouds=11, n=4000, floats=4)
    poses[:, :3] *= np.float32(0.05)                                  # overlapping keyframes, as around a loop
    got = m.assemble_submap(clouds, poses, leaf=0.3)
    want = orc.Voxel().downsample(o.assemble(clouds, poses), 0.3)
    assert len(got) == len(want["idx"]) and int(got[:, 3].sum()) == sum(len(c) for c in clouds)
    d, j = cKDTree(want["points"][:, :3].astype(np.float64)).query(got[:, :3].astype(np.float64))
    assert len(set(j.tolist())) == len(got) and d.max() < 1e-4        # the same voxels: centroids within the FP32 summation error
    assert np.array_equal(got[:, 3].astype(np.uint32), want["count"][j])


def test_verify_loop_from_keyframes_equals_composed_oracle():
    from oracle import oracle as orc
    from sc_lego_loam_b200.scgpu import SCManager
    o, m = _oracle(), SCManager()
    rng = np.random.default_rng(17)
    src3, tgt3, T = icp_clouds(n_src=1500, n_tgt=24000, noise=0.01)
    # the history submap as 12 keyframe clouds in their own sensor frames, each with its key pose; the query as corner + surface cloud
    poses = np.stack([np.float32([rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-1, 1), rng.uniform(-0.05, 0.05),
                                  rng.uniform(-0.05, 0.05), rng.uniform(-np.pi, np.pi)]) for _ in range(12)])

    def to_local(world, p):      # inverse of the key pose (FP64; the result is just some cloud that the pose maps near `world`)
        x, y, z, r, pt, yw = [float(v) for v in p]
        Rz = np.array([[np.cos(yw), -np.sin(yw), 0], [np.sin(yw), np.cos(yw), 0], [0, 0, 1]])
        Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
        Ry = np.array([[np.cos(pt), 0, np.sin(pt)], [0, 1, 0], [-np.sin(pt), 0, np.cos(pt)]])
        return ((world.astype(np.float64) - [x, y, z]) @ (Ry @ Rx @ Rz)).astype(np.float32)

    parts = np.array_split(tgt3, 12)
    tgt_clouds = []
    for part, p in zip(parts, poses):
        c = np.zeros((len(part), 4), np.float32)
        c[:, :3] = to_local(part, p)
        c[:, 3] = rng.uniform(0, 64, len(part))
        tgt_clouds.append(c)
    spose = poses[5]
    src = np.zeros((len(src3), 4), np.float32)
    src[:, :3] = to_local(src3, spose)
    src[:, 3] = rng.uniform(0, 64, len(src))
    src[::50, 3] = -1.0                                               # points the reference drops (mapOptmization.cpp:932-939)
    src_clouds = [src[:600], src[600:]]
    got = m.verify_loop_keyframes(src_clouds, spose, tgt_clouds, poses, leaf=0.3)
    o_src = o.assemble(src_clouds, [spose, spose], drop_negative_intensity=True)
    o_tgt = orc.Voxel().downsample(o.assemble(tgt_clouds, poses), 0.3)["points"]
    assert got["n_source"] == len(o_src) and got["n_target"] == len(o_tgt)
    want = orc.Icp().align(o_src, o_tgt)
    assert got["converged"] == want["converged"] and got["accepted"]
    assert abs(got["iterations"] - want["iterations"]) <= 1
    assert np.allclose(got["T"], want["T"], atol=2e-4), got["T"] - want["T"]
    assert abs(got["fitness"] - want["fitness"]) <= 1e-5 + 2e-3 * want["fitness"]
    assert np.allclose(got["T"], T, atol=2e-2)                        # and it recovers the planted loop transform
