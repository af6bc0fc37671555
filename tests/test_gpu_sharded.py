"""GPU: the sharded path through the staged C ABI (scgpu_stage_*).

(1) On ONE device, G shard handles emulate G ranks (the three all_gathers become torch.stack of the shards'
    outputs): ownership, candidate merge and the (distance, position) reduction must reproduce the single-shard
    result bit for bit.  This is how the multi-rank logic is tested where fewer GPUs than ranks are available.
(2) With >= 2 GPUs visible, the real thing: torchrun, one process per GPU, NCCL all_gathers
    (sc_lego_loam_b200/sharded.py), compared with a single-GPU replay of the same scans."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _scans(n, first=0, azim=150):
    from sc_lego_loam_b200.synth import ScanGen
    gen = ScanGen("hdl64", seed=99, n_places=90, n_azim=azim)
    return gen, gen.scans(first, n, 4)


@pytest.mark.parametrize("G,K", [(3, 10), (4, 50)])
def test_emulated_shards_equal_single_device(G, K):
    import torch
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.sharded import GpuStages
    gen, scans = _scans(24 * G)
    pre = gen.descs(1000, 61, 20, 60)
    single = SCManager(num_candidates=K)
    single.append_descs(pre)
    want = single.replay(scans)

    mgrs = [SCManager(num_candidates=K, shard_rank=r, shard_count=G) for r in range(G)]
    st = [GpuStages(m, "cuda:0") for m in mgrs]
    for s in st:
        s.prefill(pre)
    size = 61
    got = {k: [] for k in want}
    B = 8
    for b0 in range(0, 24, B):                       # three steps of G*B scans; scan j of rank r = entry size + j*G + r
        chunk = scans[b0 * G:(b0 + B) * G]
        local = [torch.from_numpy(np.ascontiguousarray(chunk[r::G])).cuda() for r in range(G)]
        gathered = torch.stack([st[r].build(local[r]) for r in range(G)])                 # exchange 1
        rec_global = gathered.transpose(0, 1).reshape(G * B, -1).contiguous()
        ns = []
        for r in range(G):
            st[r].append(rec_global, size, 1, G * B)
            st[r].set_size(size + G * B)
            ns.append(st[r].plan_n_search(size + 1, G * B))
        size += G * B
        assert all(torch.equal(ns[0], x) for x in ns)
        keys_parts = torch.stack([st[r].topk(rec_global, ns[r]) for r in range(G)])       # exchange 2
        keys = [st[r].merge(keys_parts) for r in range(G)]
        assert all(torch.equal(keys[0], k) for k in keys)
        best = torch.stack([st[r].score(rec_global, keys[r], ns[r]) for r in range(G)])   # exchange 3
        for r in range(G):
            out = st[r].finalize(best, ns[r])
            if r == 0:
                first_out = out
            else:
                assert all(torch.equal(out[k], first_out[k]) for k in out)
        for k in got:
            got[k].append(first_out[k].cpu().numpy())
    for k in want:
        g = np.concatenate(got[k])
        assert np.array_equal(g.view(np.uint8), want[k].view(np.uint8)), k
    assert (want["loop_id"] >= 0).any()
    assert sum(m.size() for m in mgrs) == G * (61 + 24 * G)   # every shard tracks the global size
    # each shard holds only its own entries
    for r, m in enumerate(mgrs):
        for i in range(3 * G):
            if i % G == r:
                m.get_entry(i)
            else:
                with pytest.raises(Exception):
                    m.get_entry(i)


def test_emulated_sharded_exhaustive_equals_single_device():
    """BASELINE config 4 sharded: each shard screens + rescans its own entries (scgpu_stage_exhaustive), the global
    winner is the minimum by (distance, index) -- identical to the single-device exhaustive search."""
    import torch
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.sharded import GpuStages, reduce_exhaustive
    from sc_lego_loam_b200.synth import ScanGen
    G, n = 4, 6000
    descs = ScanGen("hdl64", seed=5, n_places=2500).descs(0, n)
    descs[5000] = np.roll(descs[77].reshape(60, 20), 9, axis=0).ravel()
    descs[123] = descs[77]                    # duplicate on another shard: the lower index must win
    single = SCManager(capacity_hint=n)
    single.append_descs(descs)
    st = [GpuStages(SCManager(shard_rank=r, shard_count=G, capacity_hint=n), "cuda:0") for r in range(G)]
    for s_ in st:
        s_.prefill(descs)
    for q, ns in ((5000, 4000), (5999, 5949), (300, 250), (5000, 100)):
        rec = st[q % G].gather(q)
        parts = torch.stack([s_.exhaustive(rec, ns) for s_ in st])
        got = reduce_exhaustive(parts)
        want = single.exhaustive(q, ns)
        assert got == want[:3], (q, ns, got, want)
    assert reduce_exhaustive(torch.stack([s_.exhaustive(st[0].gather(5000 - 5000 % G), 4000) for s_ in st]))[2] >= 0
    assert single.exhaustive(5000, 4000)[2] == 77


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SCGPU_ROOT"])
from sc_lego_loam_b200.scgpu import SCManager
from sc_lego_loam_b200.sharded import GpuStages, ShardedSearch
from sc_lego_loam_b200.synth import ScanGen
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gen = ScanGen("hdl64", seed=99, n_places=90, n_azim=150)
B, steps = 16, 3
m = SCManager(device=local, shard_rank=rank, shard_count=world)
search = ShardedSearch(GpuStages(m, f"cuda:{local}"), rank, world)
pre = gen.descs(1000, 61, 20, 60)
search.prefill_descs(pre)
outs = []
for s in range(steps):
    first = s * B * world
    scans = torch.from_numpy(np.stack([gen.scan(first + j * world + rank, 4) for j in range(B)])).cuda()
    r = search.step(scans)
    outs.append({k: v.cpu().numpy() for k, v in r.items()})
ex = search.exhaustive(61 + 5, 61)          # sharded exhaustive search, query = a stored keyframe
if rank == 0:
    single = SCManager(device=0)
    single.append_descs(pre)
    want = single.replay(gen.scans(0, steps * B * world, 4))
    assert ex == single.exhaustive(61 + 5, 61)[:3], ex
    for k in want:
        got = np.concatenate([o[k] for o in outs])
        assert np.array_equal(got.view(np.uint8), want[k].view(np.uint8)), k
    print("SHARDED_OK", world, int((want["loop_id"] >= 0).sum()))
dist.barrier()
dist.destroy_process_group()
'''


def test_nccl_ranks_equal_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    n = 2 if n < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, SCGPU_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29400 + os.getpid() % 500), str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
