"""GPU: the peer-sharded database (include/scgpu.h "peer-sharded database").

(1) A device-list handle -- ONE scgpu handle over G shards (scgpu_config.n_devices) -- must reproduce the single-device
    results bit for bit through every public call: sequential append+detect, batched replay (sync and async, several steps
    in flight), query_batched, exhaustive search, save/load, candidate dumps.  On a 1-GPU box the device list names the same
    GPU several times: the shards are separate allocations read through the same peer tables, the kernels are the
    multi-GPU ones (synchronised by events, never by spinning -- see B200_PROFILING.md).
(2) With >= 2 GPUs: one process per GPU, cudaIpc-mapped shards, in-kernel flag barriers (scgpu_peer_replay_async) against a
    single-GPU replay of the same scans; and the device-list handle over distinct GPUs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

KEYS = ("loop_id", "yaw", "min_dist", "nn_idx", "nn_shift")


def _gen(azim=150):
    from sc_lego_loam_b200.synth import ScanGen
    return ScanGen("hdl64", seed=99, n_places=90, n_azim=azim)


def _same(a, b):
    for k in KEYS:
        assert np.array_equal(np.asarray(a[k]).view(np.uint8), np.asarray(b[k]).view(np.uint8)), k


@pytest.mark.parametrize("G,K", [(2, 10), (3, 10), (4, 50)])
def test_device_list_replay_equals_single_device(G, K):
    from sc_lego_loam_b200.scgpu import SCManager
    gen = _gen()
    scans = gen.scans(0, 24 * G + 5, 4)
    pre = gen.descs(1000, 61, 20, 60)
    single = SCManager(num_candidates=K)
    single.append_descs(pre)
    grp = SCManager(num_candidates=K, devices=[0] * G, capacity_hint=4096)
    grp.append_descs(pre)
    assert grp.size() == 61
    cut = 17                                           # not a multiple of G: the second batch starts on another shard
    for lo, hi in ((0, cut), (cut, len(scans))):
        want = single.replay(scans[lo:hi])
        got = grp.replay(scans[lo:hi])
        _same(got, want)
    assert (want["loop_id"] >= 0).any()
    assert grp.size() == single.size()
    for i in (0, 1, 60, 61, 62, 63, grp.size() - 1):     # entries live on shard i % G; every ring key is in every replica
        a, b = grp.get_entry(i), single.get_entry(i)
        assert all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))
    # batched queries and exhaustive search over the shards
    n = grp.size()
    _same(grp.query_batched(n - 23, 23), single.query_batched(n - 23, 23))
    for q, ns in ((n - 1, n - 51), (70, 30), (n - 5, 7)):
        assert grp.exhaustive(q, ns) == single.exhaustive(q, ns)
    qs = [n - 1 - 3 * i for i in range(9)]
    for a, b in zip(grp.exhaustive_batched(qs, n - 50), single.exhaustive_batched(qs, n - 50)):
        assert np.array_equal(a, b)


def test_device_list_sequential_calls_and_candidates():
    """The reference's own call pattern (one append, one detect) on a device-list handle: the scan is binned on the shard
    that owns it, the detect runs on that shard with candidates fetched from the others."""
    from sc_lego_loam_b200.scgpu import SCManager
    gen = _gen()
    single, grp = SCManager(), SCManager(devices=[0, 0, 0], capacity_hint=1024)
    n_loop = 0
    for i in range(130):
        s = gen.scan(i, 4)
        single.makeAndSaveScancontextAndKeys(s)
        grp.makeAndSaveScancontextAndKeys(s)
        a, b = single.detectLoopClosureID(details=True), grp.detectLoopClosureID(details=True)
        assert a["loop_id"] == b["loop_id"] and np.float32(a["yaw"]).tobytes() == np.float32(b["yaw"]).tobytes(), (i, a, b)
        assert a["min_dist"] == b["min_dist"] and a["nn_idx"] == b["nn_idx"] and a["nn_shift"] == b["nn_shift"]
        if i >= 50 and i % 7 == 0:
            ca, cb = single.candidates(), grp.candidates()
            assert all(np.array_equal(ca[k], cb[k]) for k in ("cand_idx", "cand_shift")) and ca["n_tree"] == cb["n_tree"]
            assert np.array_equal(ca["cand_d2"].view(np.uint32), cb["cand_d2"].view(np.uint32))
            assert np.array_equal(ca["cand_dist"], cb["cand_dist"], equal_nan=True)
        n_loop += a["loop_id"] >= 0
    assert n_loop > 0
    sc = gen.scan(3, 4)
    assert np.array_equal(single.makeScancontext(sc), grp.makeScancontext(sc))


def test_async_replays_in_flight_equal_sync():
    """Several replay steps enqueued back to back (the build of step n+1 overlaps the query stage of step n; its append
    waits for those queries) give the results of the synchronous call; chunked steps (SCGPU_REPLAY_CHUNKS) too."""
    import torch
    from sc_lego_loam_b200.scgpu import SCManager
    gen = _gen()
    scans = gen.scans(0, 150, 4)
    d = torch.from_numpy(scans).cuda()
    ref = SCManager()
    want = ref.replay(scans)
    for devices in (None, [0, 0]):
        m = SCManager(capacity_hint=1024) if devices is None else SCManager(devices=devices, capacity_hint=1024)
        for _ in range(4):
            m.truncate(0)
            m.replay_async((d.data_ptr(), 150, scans.shape[1], 16, 1))
        _same(m.replay_results(150), want)
        # host scans, pageable (packed to 12 bytes per point by the host threads) and pinned (copied as they are)
        m.truncate(0)
        _same(m.replay(scans), want)
        pinned = torch.from_numpy(scans).pin_memory()
        m.truncate(0)
        _same(m.replay((pinned.data_ptr(), 150, scans.shape[1], 16, 0)), want)
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np\nfrom sc_lego_loam_b200.scgpu import SCManager\nfrom sc_lego_loam_b200.synth import ScanGen\n"
            "gen = ScanGen('hdl64', seed=99, n_places=90, n_azim=150)\nscans = gen.scans(0, 150, 4)\n"
            "a = SCManager().replay(scans)\nnp.save(sys.argv[1], np.stack([a['loop_id'].astype(np.float64), a['yaw'].astype(np.float64), a['min_dist']]))\n") % ROOT
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "r.npy")
        r = subprocess.run([sys.executable, "-c", code, out], env=dict(os.environ, SCGPU_REPLAY_CHUNKS="4"), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        got = np.load(out)
    assert np.array_equal(got[0], want["loop_id"].astype(np.float64)) and np.array_equal(got[2], want["min_dist"])
    assert np.array_equal(got[1].astype(np.float32).view(np.uint32), want["yaw"].view(np.uint32))


@pytest.mark.parametrize("k,pinned", [(3, False), (4, False), (8, False), (5, False), (3, True), (8, True)])
def test_host_strides_and_packing(k, pinned):
    """Host scans of stride 12 / 16 / 32 / 20 bytes, pageable (packed to 12 B by the thread pool; k_build_tma<12> when the
    point count is a multiple of 4, the register-staged kernel otherwise) and pinned (copied as given)."""
    import torch
    from sc_lego_loam_b200.scgpu import SCManager
    gen = _gen(azim=150)
    base = gen.scans(0, 40, 4)[:, :, :3]
    for npts in (base.shape[1], base.shape[1] - 3):
        scans = np.zeros((40, npts, k), np.float32)
        scans[:, :, :3] = base[:, :npts]
        scans[:, :, 3:] = 7.5                                  # payload that must be ignored (intensity, padding)
        want = SCManager().replay(np.ascontiguousarray(base[:, :npts]))
        m = SCManager()
        if pinned:
            t = torch.from_numpy(scans).pin_memory()
            got = m.replay((t.data_ptr(), 40, npts, 4 * k, 0))
        else:
            got = m.replay(scans)
        _same(got, want)
        for i in (0, 39):
            assert np.array_equal(m.get_entry(i)[0], SCManager().makeScancontext(scans[i]).astype(np.float32))


def test_save_load_across_device_lists(tmp_path):
    from sc_lego_loam_b200.scgpu import SCManager, ScgpuError
    gen = _gen()
    scans = gen.scans(0, 90, 4)
    a = SCManager(devices=[0, 0, 0], capacity_hint=1024)
    ra = a.replay(scans[:70])
    path = str(tmp_path / "db.scgpu")
    a.save(path)
    b = SCManager()
    b.load(path)
    c = SCManager(devices=[0, 0], capacity_hint=1024)
    c.load(path)
    assert b.size() == c.size() == 70
    _same(b.replay(scans[70:]), a.replay(scans[70:]))        # the snapshot state travels with the file
    c.replay(scans[70:])
    for i in (0, 33, 69, 89):
        assert np.array_equal(b.get_entry(i)[0], c.get_entry(i)[0])
    with pytest.raises(ScgpuError):                            # load replaces: only into an empty handle
        b.load(path)
    with pytest.raises(ScgpuError):                            # other binning constants: refuse
        SCManager(lidar_height=1.0).load(path)
    bad = str(tmp_path / "bad.scgpu")
    raw = bytearray(open(path, "rb").read())
    raw[24:32] = (10 ** 12).to_bytes(8, "little")              # entry count from an untrusted header
    open(bad, "wb").write(raw)
    with pytest.raises(ScgpuError):
        SCManager().load(bad)
    assert ra["loop_id"].shape == (70,)


def test_staged_exhaustive_overflow_falls_back_to_exact():
    """A database of duplicates overflows the rescoring list (65,536 entries): the staged call reports it in n_rescored and
    ShardedSearch.exhaustive_records redoes those queries exactly on the shard."""
    import torch
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.sharded import EXH_LIST_CAP, GpuStages, ShardedSearch
    gen = _gen()
    d = gen.descs(0, 4, 20, 60)
    n = 70000
    descs = np.repeat(d[1:2], n, axis=0)
    descs[0] = d[0]
    descs[n - 1] = d[2]
    m = SCManager(capacity_hint=n + 8)
    m.append_descs(descs)
    st = GpuStages(m, "cuda:0")
    search = ShardedSearch(st, rank=0, world=1)
    rec = torch.stack([st.gather(n - 1), st.gather(5), st.gather(0)])
    raw = st.exhaustive(rec, n - 2)
    assert int(raw[1, 1] & 0xffffffff) > EXH_LIST_CAP          # the duplicate query overflowed the shared list
    dist_, shift, idx = search.exhaustive_records(rec, n - 2)
    want = [m.exhaustive(q, n - 2) for q in (n - 1, 5, 0)]
    for i, w in enumerate(want):
        assert (float(dist_[i]), int(shift[i]), int(idx[i])) == w[:3], (i, w)
    assert int(idx[1]) == 1 and float(dist_[1]) == 0.0           # smallest index among 69,997 exact ties


_PEER_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SCGPU_ROOT"])
from sc_lego_loam_b200.scgpu import SCManager, FLAG_PEER
from sc_lego_loam_b200.sharded import PeerShardedSearch
from sc_lego_loam_b200.synth import ScanGen
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gen = ScanGen("hdl64", seed=99, n_places=90, n_azim=150)
B, steps = 16, 3
m = SCManager(device=local, shard_rank=rank, shard_count=world, flags=FLAG_PEER, capacity_hint=4096)
search = PeerShardedSearch(m, rank, world)
N0 = 61                                    # not a multiple of the ranks: batches start on a shard other than 0
pre = gen.descs(1000, N0, 20, 60)
search.prefill_descs(pre)
outs = []
size = N0
def batch(size, n_total):
    mine = [i for i in range(n_total) if (size + i) % world == rank]
    return np.stack([gen.scan(size - N0 + i, 4) for i in mine])
for s in range(steps):
    n_total = B * world + (1 if s == 1 else 0)        # a ragged batch in the middle
    scans = batch(size, n_total)
    d = torch.from_numpy(scans).cuda()
    if s == 2:                                        # host scans on the last step (H2D inside)
        search.step_async(scans, n_total)
    else:
        search.step_async((d.data_ptr(), scans.shape[0], scans.shape[1], 16, 1), n_total)
    outs.append(search.results(n_total))
    size += n_total
# several steps in flight, each replayed from the prefilled size: the first batch again
scans = batch(N0, B * world)
d = torch.from_numpy(scans).cuda()
for _ in range(3):
    search.truncate(N0)
    search.step_async((d.data_ptr(), scans.shape[0], scans.shape[1], 16, 1), B * world)
again = search.results(B * world)
for k in again:
    assert np.array_equal(again[k].view(np.uint8), outs[0][k].view(np.uint8)), k
if rank == 0:
    single = SCManager(device=0)
    single.append_descs(pre)
    want = single.replay(gen.scans(0, size - N0, 4))
    for k in want:
        got = np.concatenate([o[k] for o in outs])
        assert np.array_equal(got.view(np.uint8), want[k].view(np.uint8)), k
    print("PEER_OK", world, int((want["loop_id"] >= 0).sum()))
dist.barrier()
dist.destroy_process_group()
'''


def test_peer_ranks_equal_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    n = 2 if n < 4 else 4
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_WORKER)
    env = dict(os.environ, SCGPU_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29400 + os.getpid() % 500), str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "PEER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_device_list_over_distinct_gpus():
    import torch
    from sc_lego_loam_b200.scgpu import SCManager
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    gen = _gen()
    scans = gen.scans(0, 120, 4)
    single = SCManager()
    want = single.replay(scans)
    grp = SCManager(devices=list(range(min(n, 4))), capacity_hint=1024)
    _same(grp.replay(scans), want)
    q = grp.size() - 1
    assert grp.exhaustive(q, q - 50) == single.exhaustive(q, q - 50)
