"""CPU, world_size 2 over gloo: the sharding orchestration of sc_lego_loam_b200/sharded.py (entry i on rank i % G,
three all_gathers, deterministic merges) reproduces the single-process result exactly.

The stage backend here is a stand-in built on the ORACLE (numpy) with the same stage contract as GpuStages; the
product backend (the staged C ABI) is exercised on GPUs by tests/test_gpu_sharded.py.  What this test pins is
the host logic that is identical for both: ordering of queries (scan j of rank r = entry first + j*G + r),
ownership, the snapshot plan shared by all ranks, key merge order and the (distance, position) final reduction."""
import os
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

R, S, K, EXCL, PERIOD = 20, 60, 10, 50, 10
NONE = -1  # UINT64_MAX as int64


def _f2u(f):
    return struct.unpack("<I", struct.pack("<f", float(f)))[0]


class OracleStages:
    def __init__(self, rank, world):
        from oracle import oracle as orc
        self.port = orc.Port(orc.Params(R=R, S=S, num_candidates=K, exclude_recent=EXCL, tree_period=PERIOD))
        self.rank, self.G = rank, world
        self.db = {}          # global idx -> (sc float64, ring float32)
        self.n = 0
        self.counter, self.n_tree = 0, 0

    def _rec(self, sc):
        ring = self.port.ringkey(sc).astype(np.float32)
        return np.concatenate([sc.astype(np.float32), ring])

    def build(self, scans):
        return torch.from_numpy(np.stack([self._rec(self.port.make_sc(s.numpy())) for s in scans]))

    def prefill(self, descs):
        for i, d in enumerate(descs):
            if i % self.G == self.rank:
                r = self._rec(d.astype(np.float64))
                self.db[i] = (r[:R * S].astype(np.float64), r[R * S:])
        self.n = len(descs)

    def size(self):
        return self.n

    def append(self, rec, first, step, n):
        rec = rec.numpy()
        for i in range(n):
            g = first + i * step
            if g % self.G == self.rank:
                self.db[g] = (rec[i, :R * S].astype(np.float64), rec[i, R * S:].copy())
        self.n = max(self.n, first + (n - 1) * step + 1)

    def set_size(self, n):
        self.n = n

    def plan_n_search(self, first_size, n):   # Scancontext.cpp:257-276
        out = np.zeros(n, np.int64)
        for i in range(n):
            size = first_size + i
            if size < EXCL + 1:
                continue
            if self.counter % PERIOD == 0:
                self.n_tree = size - EXCL
            self.counter += 1
            out[i] = self.n_tree
        return torch.from_numpy(out)

    def topk(self, qrec, n_search):
        q = qrec.numpy()
        out = np.full((len(q), K), NONE, np.int64)
        for i in range(len(q)):
            cand = sorted((( _f2u(self.port.L.sco_key_dist2(np.ascontiguousarray(q[i, R * S:]), np.ascontiguousarray(ring), R)) << 32) | g)
                          for g, (_, ring) in self.db.items() if g < int(n_search[i]))[:K]
            out[i, :len(cand)] = np.array(cand, np.uint64).view(np.int64)
        return torch.from_numpy(out)

    def merge(self, parts):
        p = parts.numpy().view(np.uint64)           # [G, NQ, K]
        allk = np.concatenate(list(p), axis=1)
        allk.sort(axis=1)
        return torch.from_numpy(allk[:, :K].copy().view(np.int64))

    def score(self, qrec, keys, n_search):
        q, keys = qrec.numpy(), keys.numpy().view(np.uint64)
        out = np.zeros((len(q), 3), np.int64)
        for i in range(len(q)):
            best = (1e7, K, 0, 0)
            if int(n_search[i]) != 0:
                for k in range(K):
                    g = 0 if keys[i, k] == np.uint64(0xFFFFFFFFFFFFFFFF) else int(keys[i, k] & np.uint64(0xFFFFFFFF))
                    if g % self.G != self.rank:
                        continue
                    d, s = self.port.distance(q[i, :R * S].astype(np.float64), self.db[g][0])
                    if d < best[0]:
                        best = (d, k, s, g)
            out[i] = [np.float64(best[0]).view(np.int64), best[1] | (best[2] << 32), best[3]]
        return torch.from_numpy(out)

    def finalize(self, parts, n_search):
        p = parts.numpy()                              # [G, NQ, 3]
        nq = p.shape[1]
        res = dict(loop_id=np.full(nq, -1, np.int32), yaw=np.zeros(nq, np.float32), min_dist=np.full(nq, 1e7),
                   nn_idx=np.zeros(nq, np.int32), nn_shift=np.zeros(nq, np.int32))
        for i in range(nq):
            if int(n_search[i]) == 0:
                continue
            best = (1e7, K, 0, 0)
            for g in range(p.shape[0]):
                d = float(p[g, i, 0:1].view(np.float64)[0])
                rank, shift, idx = int(p[g, i, 1] & 0xFFFFFFFF), int(p[g, i, 1] >> 32), int(p[g, i, 2])
                if rank < K and (d < best[0] or (d == best[0] and rank < best[1])):
                    best = (d, rank, shift, idx)
            res["min_dist"][i], res["nn_shift"][i], res["nn_idx"][i] = best[0], best[2], best[3]
            if best[0] < 0.5:
                res["loop_id"][i] = best[3]
            deg = np.float32(best[2] * (360.0 / S))
            res["yaw"][i] = np.float32(np.float64(deg) * np.pi / 180.0)
        return {k: torch.from_numpy(v) for k, v in res.items()}


def _worker(rank, world, port_no, ret):
    try:
        _worker_body(rank, world, port_no, ret)
    except Exception as e:  # surface the failure instead of letting the parent wait for its timeout
        import traceback
        ret.put(("error", rank, traceback.format_exc()))
        raise


def _worker_body(rank, world, port_no, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sc_lego_loam_b200.sharded import ShardedSearch
    from sc_lego_loam_b200.synth import ScanGen
    gen = ScanGen("hdl64", seed=4242, n_places=60, n_azim=60)
    search = ShardedSearch(OracleStages(rank, world))
    search.prefill_descs(gen.descs(0, 44, R, S))
    out = []
    B = 5
    for step in range(3):
        first = 44 + step * B * world
        scans = torch.from_numpy(np.stack([gen.scan(first + j * world + rank, 3) for j in range(B)]))
        r = search.step(scans)
        out.append({k: v.numpy().copy() for k, v in r.items()})
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_ranks_equal_single_process():
    from oracle import oracle as orc
    from sc_lego_loam_b200.synth import ScanGen
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port_no = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=240)
    assert not (isinstance(got, tuple) and got[0] == "error"), got
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process oracle: the same 44 descriptors then the same scans in global order
    gen = ScanGen("hdl64", seed=4242, n_places=60, n_azim=60)
    port = orc.Port(orc.Params(R=R, S=S, num_candidates=K, exclude_recent=EXCL, tree_period=PERIOD))
    for d in gen.descs(0, 44, R, S):
        port.append_desc(d.astype(np.float64))
    want = []
    for i in range(44, 44 + 3 * 10):
        port.append_scan(gen.scan(i, 3))
        want.append(port.detect())
    flat = {k: np.concatenate([g[k] for g in got]) for k in got[0]}
    assert np.array_equal(flat["loop_id"], [w["loop_id"] for w in want])
    assert np.array_equal(flat["yaw"].view(np.uint32), np.array([w["yaw"] for w in want], np.float32).view(np.uint32))
    assert np.array_equal(flat["min_dist"], [w["min_dist"] for w in want])
    assert (flat["loop_id"] >= 0).any() or len(want) > 0
