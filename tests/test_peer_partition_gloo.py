"""CPU, world_size 2 over gloo: the host logic of the PEER-SHARDED step (include/scgpu.h "peer-sharded database") -- which rank
bins / stores / searches for which scan of a batch (scgpu_peer_partition, the arithmetic scgpu_peer_replay_async applies), the
snapshot plan every rank derives identically, and the placement of each rank's results in the global order.

The device work of a rank (k_build, k_append with ring keys pushed to every replica, k_topk over the replica, peer fetch of the
candidates, k_best_finalize pushing results) is replaced by the ORACLE here: every rank keeps a full oracle database that the
all_gather of descriptors stands in for -- what the NVLink peer mappings give the kernels.  The product path runs on GPUs in
tests/test_gpu_peer.py; this test pins the partition and ordering that both share."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_partition_arithmetic_without_a_device():
    from sc_lego_loam_b200.scgpu import peer_partition
    for G in (1, 2, 3, 4, 8):
        for first in (0, 1, 5, 61, 4541):
            for n in (0, 1, 2, 7, 16, 33):
                seen = []
                for r in range(G):
                    off, cnt = peer_partition(first, n, G, r)
                    mine = [off + j * G for j in range(cnt)]
                    assert all((first + i) % G == r and i < n for i in mine)
                    seen += mine
                assert sorted(seen) == list(range(n))          # every scan of the batch has exactly one owner


def _worker(rank, world, port_no, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from sc_lego_loam_b200.scgpu import peer_partition
    from sc_lego_loam_b200.synth import ScanGen
    gen = ScanGen("hdl64", seed=99, n_places=90, n_azim=120)
    P = orc.Params()
    port = orc.Port(P)                    # this rank's view of the database (own shard + what the peer mappings would show)
    n0 = 61
    for d in gen.descs(1000, n0, 20, 60):
        port.append_desc(d.astype(np.float64))
    size, results = n0, {}
    for n_total in (24, 17, 30):          # ragged batches: the second and third start on the other shard
        off, cnt = peer_partition(size, n_total, world, rank)
        mine = [off + j * world for j in range(cnt)]
        # stage 1+2 on this rank's own scans only
        local = np.stack([port.make_sc(gen.scan(size - n0 + i, 4)) for i in mine]) if mine else np.zeros((0, 1200))
        # "peer memory": every rank can read every descriptor of the batch (here: one all_gather of padded blocks)
        B = (n_total + world - 1) // world
        pad = np.zeros((B, 1200))
        pad[:len(local)] = local
        got = [torch.zeros((B, 1200), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(got, torch.from_numpy(pad))
        batch = np.zeros((n_total, 1200))
        for r in range(world):
            o, c = peer_partition(size, n_total, world, r)
            for j in range(c):
                batch[o + j * world] = got[r][j].numpy()
        # every rank advances the same database and snapshot state, but SEARCHES only for its own scans
        for i in range(n_total):
            port.append_desc(batch[i])
            d = port.detect()             # (advances the tree-snapshot counter on every rank, as plan() does)
            if i in mine:
                results[size + i] = (d["loop_id"], np.float32(d["yaw"]).tobytes())
        size += n_total
    out = [None] * world
    dist.all_gather_object(out, results)
    if rank == 0:
        merged = {}
        for o in out:
            assert not set(merged) & set(o)                    # no query was searched twice
            merged.update(o)
        ref = orc.Port(P)
        for d in gen.descs(1000, n0, 20, 60):
            ref.append_desc(d.astype(np.float64))
        ok = sorted(merged) == list(range(n0, size))
        loops = 0
        for i in range(size - n0):
            ref.append_scan(gen.scan(i, 4))
            d = ref.detect()
            ok &= merged[n0 + i] == (d["loop_id"], np.float32(d["yaw"]).tobytes())
            loops += d["loop_id"] >= 0
        q.put((ok, loops))
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_search_over_gloo_equals_sequential():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, loops = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert ok and loops > 0
