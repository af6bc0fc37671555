"""Keyframe database sharded over the ranks of one box.

Entry i lives on rank i % G.  One process per GPU; torch.distributed supplies the communicator (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The path has exactly three exchange steps per batch of queries
and nothing else crosses ranks:

    1. all_gather of the freshly built descriptor records (each rank bins its own scans)       ~5.8 KB / scan
    2. all_gather of the per-shard top-K ring-key candidates  (K x 8 B per query per rank)
    3. all_gather of the per-shard best (distance, position, shift, index) (24 B per query per rank)

followed by a deterministic merge that reproduces the single-device result bit for bit: candidate lists merge in
(squared distance, index) order, and the final argmin is the strict-min in candidate order (Scancontext.cpp:296-311).

``ShardedSearch`` is backend-agnostic: ``stages`` provides build / append / topk / merge / score / finalize on
torch tensors.  ``GpuStages`` (below) is the product backend -- the staged C ABI of include/scgpu.h on the
current CUDA stream.  The CPU tests drive the same orchestration over gloo with a stand-in backend built on
the oracle (tests/test_sharded_gloo.py).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class ShardedSearch:
    def __init__(self, stages, rank=None, world=None, group=None):
        self.st = stages
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.size = 0  # global number of entries

    def _all_gather(self, t):
        """[...]-> [G, ...] (same shape on every rank)."""
        if self.world == 1:
            return t.unsqueeze(0)
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)     # rank-major concatenation along dim 0
        return out.view((self.world,) + tuple(t.shape))

    def prefill_descs(self, descs):
        """Every rank passes the SAME (n, R*S) float32 descriptors; each keeps the entries it owns."""
        self.st.prefill(descs)          # ownership (i % G == rank) is applied by the backend
        self.size = self.st.size()

    def exhaustive(self, global_idx, n_search):
        """Every entry [0, n_search) of every shard scored against stored entry global_idx (BASELINE config 4):
        the owner broadcasts the query record, each shard searches its part, one all_gather of 24 bytes per rank."""
        owner = global_idx % self.world
        rec = self.st.gather(global_idx) if self.rank == owner else torch.empty(self.st.rec_bytes, dtype=torch.uint8, device=self.st.device)
        if self.world > 1:
            dist.broadcast(rec, src=owner, group=self.group)
        d, s, i = self.exhaustive_records(rec.unsqueeze(0), n_search)
        return float(d[0].item()), int(s[0].item()), int(i[0].item())

    def exhaustive_stored(self, first_idx, nq, n_search):
        """Exhaustive search for the stored entries first_idx .. first_idx+nq-1 as queries (first_idx and nq multiples
        of the world size): every rank gathers the records of the queries it owns, one all_gather makes them global."""
        G = self.world
        assert first_idx % G == 0 and nq % G == 0
        mine = torch.stack([self.st.gather(first_idx + j * G + self.rank) for j in range(nq // G)])     # [nq/G, rec]
        recs = self._all_gather(mine).transpose(0, 1).reshape(nq, -1).contiguous()                      # query i = j*G + r
        return self.exhaustive_records(recs, n_search)

    def exhaustive_records(self, qrecs, n_search, flipped=False):
        """The same for a batch of query records present on every rank ([nq, rec]); returns device tensors
        (dist, shift, idx) per query.  One all_gather of 24 bytes per query per rank.  A shard whose rescoring list
        overflowed (n_rescored > EXH_LIST_CAP: a database of near-duplicates) reports an unreliable winner; those
        queries are redone exactly on that shard (scgpu_stage_exhaustive_exact) and gathered again."""
        local = self.st.exhaustive(qrecs, n_search, flipped)
        parts = self._all_gather(local)
        over = (parts[..., 1] & 0xffffffff) > EXH_LIST_CAP                        # [G, nq]
        if bool(over.any().item()):
            mine = over[self.rank].nonzero().flatten().tolist()
            ns = np.broadcast_to(np.asarray(n_search, np.uint64), (qrecs.shape[0],))
            for q in mine:
                local[q] = self.st.exhaustive_exact(qrecs[q], int(ns[q]))
            parts = self._all_gather(local)
        return reduce_exhaustive_batch(parts)

    def step(self, scans_local):
        """scans_local: this rank's B scans ([B, P, k] float32 on the stage device).  Scan j of rank r becomes
        global entry size + j*G + r.  Returns the detect results of all G*B new entries in global order
        (identical on every rank): dict(loop_id, yaw, min_dist, nn_idx, nn_shift)."""
        G, B = self.world, scans_local.shape[0]
        ev = getattr(self, "build_events", None)      # bench.py: CUDA events around the build launch (roofline.achieved)
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rec_local = self.st.build(scans_local)                                   # [B, rec]
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        gathered = self._all_gather(rec_local)                                   # [G, B, rec]   exchange 1
        rec_global = gathered.transpose(0, 1).reshape(G * B, -1).contiguous()    # query q = j*G + r
        first = self.size
        self.st.append(rec_global, first, 1, G * B)
        self.size = first + G * B
        self.st.set_size(self.size)
        n_search = self.st.plan_n_search(first + 1, G * B)                       # same on every rank
        keys_local = self.st.topk(rec_global, n_search)                          # [GB, K]
        keys = self.st.merge(self._all_gather(keys_local))                       # exchange 2
        best_local = self.st.score(rec_global, keys, n_search)                   # [GB, 3] int64
        return self.st.finalize(self._all_gather(best_local), n_search)          # exchange 3


EXH_LIST_CAP = 65536      # SCGPU_EXH_LIST_CAP (include/scgpu.h)


def reduce_exhaustive_batch(parts):
    """parts: [G, nq, 3] int64 -> (dist [nq] f64, shift [nq] i32, idx [nq] i64) of the global winners, on the device:
    per query the minimum by (dist, idx) over the shards with dist < 1e7; (1e7, 0, 0) when none."""
    d = parts[..., 0].contiguous().view(torch.float64)                      # [G, nq]
    idx = parts[..., 2]
    shift = parts[..., 1] >> 32
    big = torch.iinfo(torch.int64).max
    dmin = d.min(dim=0).values                                              # [nq]
    cand = (d == dmin) & (d < 1e7)
    idx_masked = torch.where(cand, idx, torch.full_like(idx, big))
    g = idx_masked.argmin(dim=0)                                            # shard of the winner
    none = ~cand.any(dim=0)
    take = lambda t: t.gather(0, g.unsqueeze(0)).squeeze(0)                 # noqa: E731
    out_d = torch.where(none, torch.full_like(dmin, 1e7), take(d))
    out_s = torch.where(none, torch.zeros_like(g), take(shift) & 0x3fffffff).to(torch.int32)       # bit 30: flipped winner
    out_i = torch.where(none, torch.zeros_like(g), take(idx))
    return out_d, out_s, out_i


def reduce_exhaustive(parts):
    """parts: [G, 3] int64 views of {f64 dist, (i32 n_rescored, i32 shift), i64 global idx} -> (dist, shift, idx) of the
    global winner: minimum by (dist, idx) over shards with dist < 1e7 (the strict-min in index order of
    Scancontext.cpp:296-311 over every entry); (1e7, 0, 0) when no shard has a finite distance."""
    p = parts.cpu().numpy()
    best = (1e7, 0, 0)
    found = False
    for g in range(p.shape[0]):
        d = float(p[g, 0:1].view(np.float64)[0])
        shift, idx = int(p[g, 1] >> 32), int(p[g, 2])
        if d < 1e7 and (not found or d < best[0] or (d == best[0] and idx < best[2])):
            best, found = (d, shift, idx), True
    return best


class GpuStages:
    """The staged C ABI (scgpu_stage_*) on torch CUDA tensors, launched on torch's current stream so that the
    kernels and the NCCL collectives are ordered without host synchronisation."""

    def __init__(self, manager, device):
        self.m = manager
        self.lib = manager.lib
        self.h = manager.h
        self.device = torch.device(device)
        self.rec_bytes = manager.record_bytes()
        self.K = manager.K

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        from .scgpu import _check
        _check(rc)

    def build(self, scans):
        assert scans.is_cuda and scans.dtype == torch.float32 and scans.is_contiguous()
        B, P, k = scans.shape
        rec = torch.empty((B, self.rec_bytes), dtype=torch.uint8, device=self.device)
        self._check(self.lib.scgpu_stage_build(self.h, scans.data_ptr(), B, P, k * 4, rec.data_ptr(), self._stream()))
        return rec

    def prefill(self, descs):
        self.m.append_descs(descs)      # scgpu_append_descs keeps the entries this shard owns

    def size(self):
        return self.m.size()

    def append(self, rec, first, step, n):
        self._check(self.lib.scgpu_stage_append(self.h, rec.data_ptr(), first, step, n, self._stream()))

    def set_size(self, n):
        self._check(self.lib.scgpu_stage_set_size(self.h, n))

    def plan_n_search(self, first_size, n):
        ns = self.m.plan_n_search(first_size, n)
        return torch.from_numpy(ns.astype(np.int64)).to(self.device, non_blocking=False)

    def topk(self, qrec, n_search):
        nq = qrec.shape[0]
        keys = torch.empty((nq, self.K), dtype=torch.int64, device=self.device)
        self._check(self.lib.scgpu_stage_topk(self.h, qrec.data_ptr(), nq, n_search.data_ptr(), keys.data_ptr(), self._stream()))
        return keys

    def merge(self, parts):
        G, nq, K = parts.shape
        if G == 1:
            return parts[0]
        out = torch.empty((nq, K), dtype=torch.int64, device=self.device)
        self._check(self.lib.scgpu_stage_merge(self.h, parts.data_ptr(), G, nq, out.data_ptr(), self._stream()))
        return out

    def score(self, qrec, keys, n_search):
        nq = qrec.shape[0]
        best = torch.empty((nq, 3), dtype=torch.int64, device=self.device)   # {f64 dist, i32 rank, i32 shift, i64 idx}
        self._check(self.lib.scgpu_stage_score(self.h, qrec.data_ptr(), nq, keys.data_ptr(), n_search.data_ptr(),
                                               best.data_ptr(), self._stream()))
        return best

    def gather(self, global_idx):
        rec = torch.empty(self.rec_bytes, dtype=torch.uint8, device=self.device)
        self._check(self.lib.scgpu_stage_gather(self.h, global_idx, rec.data_ptr(), self._stream()))
        return rec

    def exhaustive(self, qrec, n_search, flipped=False):
        """This shard's exhaustive winners for the query record(s) qrec ([rec] or [nq, rec]): [3] / [nq, 3] int64.
        flipped: also score every entry with its columns reversed (a flipped winner has bit 30 of the shift set)."""
        single = qrec.dim() == 1
        nq = 1 if single else qrec.shape[0]
        ns = np.ascontiguousarray(np.broadcast_to(np.asarray(n_search, np.uint64), (nq,)))
        best = torch.empty((nq, 3), dtype=torch.int64, device=self.device)
        self._check(self.lib.scgpu_stage_exhaustive2(self.h, qrec.data_ptr(), nq, ns.ctypes.data, int(bool(flipped)), best.data_ptr(),
                                                     self._stream()))
        return best[0] if single else best

    def exhaustive_exact(self, qrec, n_search):
        """Every local entry scored exactly for one query record (overflow fallback): [3] int64."""
        best = torch.empty(3, dtype=torch.int64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.lib.scgpu_stage_exhaustive_exact(self.h, qrec.data_ptr(), int(n_search), best.data_ptr()))
        return best

    def finalize(self, parts, n_search):
        G, nq, _ = parts.shape
        dev = self.device
        out = dict(loop_id=torch.empty(nq, dtype=torch.int32, device=dev), yaw=torch.empty(nq, dtype=torch.float32, device=dev),
                   min_dist=torch.empty(nq, dtype=torch.float64, device=dev), nn_idx=torch.empty(nq, dtype=torch.int32, device=dev),
                   nn_shift=torch.empty(nq, dtype=torch.int32, device=dev))
        self._check(self.lib.scgpu_stage_finalize(self.h, parts.contiguous().data_ptr(), G, nq, n_search.data_ptr(),
                                                  out["loop_id"].data_ptr(), out["yaw"].data_ptr(), out["min_dist"].data_ptr(),
                                                  out["nn_idx"].data_ptr(), out["nn_shift"].data_ptr(), self._stream()))
        return out


class PeerShardedSearch:
    """The database sharded over the ranks with the shards reading and writing each other's HBM over NVLink
    (include/scgpu.h "peer-sharded database").  torch.distributed is used ONCE, to exchange the shards' IPC handles;
    after that a step is one collective C call per rank -- ring keys are pushed into every shard's replica by k_append,
    queries are partitioned (the rank that binned a scan searches for it), candidate rows are fetched from their owners
    inside k_cand_screen / k_score_pairs, results are pushed to every rank by k_finalize, and the two synchronisation
    points of a step are in-kernel flag barriers over peer memory.  No NCCL call on the data path."""

    def __init__(self, manager, rank=None, world=None, group=None):
        self.m = manager
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        dev = torch.device("cuda", manager.cfg.device)
        mine = torch.from_numpy(manager.peer_export()).to(dev)
        blobs = torch.empty((self.world, mine.numel()), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(blobs, mine, group=group)
        manager.peer_attach(blobs.cpu().numpy())
        dist.barrier(group=group)          # every rank has mapped every shard before anyone stores into a peer

    def prefill_descs(self, descs):
        """Every rank passes the SAME descriptors; each keeps the entries it owns and every ring key."""
        self.m.append_descs(descs)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def truncate(self, n):
        self.m.truncate(n)

    def step_async(self, scans_local, n_total):
        """scans_local: this rank's scans of the batch ((ptr, n, pts, stride, location) tuple or numpy array)."""
        self.m.peer_replay_async(scans_local, n_total)

    def results(self, n_total, out=None):
        return self.m.replay_results(n_total, out)

    def step(self, scans_local, n_total):
        self.step_async(scans_local, n_total)
        return self.results(n_total)
