#!/usr/bin/env python
"""BASELINE config 4: exhaustive exact search -- every entry of an N-keyframe database (20x60) scored against the
query.  Prints one JSON object: queries/s, ms per database pass, the screening kernel's HBM roofline fraction
(algorithmic bytes = 4*R*S*N per pass) and, with --cpu N, the reference timed on N entries."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main_sharded(a):
    """torchrun: the N-entry database sharded i % G over the ranks; batches of 64 stored entries as queries."""
    import torch
    import torch.distributed as dist
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.sharded import GpuStages, ShardedSearch
    from sc_lego_loam_b200.synth import ScanGen
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("NCCL_DEBUG_FILE", f"/tmp/scgpu_nccl_{os.getpid()}_%h_%p.log")  # keep stdout to the JSON line
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R, S = 20, 60
    descs = ScanGen("hdl64", seed=20181004, n_places=70000).descs(0, a.n, R, S, threads=max(2, (os.cpu_count() or 8) // world))
    m = SCManager(device=local, shard_rank=rank, shard_count=world, capacity_hint=a.n + 8)
    search = ShardedSearch(GpuStages(m, f"cuda:{local}"), rank, world)
    search.prefill_descs(descs)
    nq = 64
    first = (a.n - 64 - 8 * world) // world * world
    ns = first - 50
    for _ in range(3):
        out = search.exhaustive_stored(first, nq, ns)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, a.queries // nq)
    e0.record()
    for _ in range(reps):
        out = search.exhaustive_stored(first, nq, ns)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak = 6523.3
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        per_q = ms.item() / (reps * nq)
        print(json.dumps({"config": f"exhaustive_{a.n}_20x60_sharded", "n_gpus": world, "queries": reps * nq, "ms_per_query": per_q,
                          "queries_per_sec": 1e3 / per_q, "aggregate_db_stream_gbs": 4 * R * S * ns / (per_q * 1e-3) / 1e9,
                          "frac_of_aggregate_hbm_roofline": 4 * R * S * ns / (per_q * 1e-3) / 1e9 / (peak * world),
                          "winner_idx_sample": [int(x) for x in out[2][:4].cpu()]}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entries", dest="n", type=int, default=100000)
    ap.add_argument("--queries", type=int, default=32)
    ap.add_argument("--cpu", type=int, default=0, help="time the CPU reference on this many entries (0 = skip)")
    a = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_sharded(a)
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.synth import ScanGen
    R, S = 20, 60
    gen = ScanGen("hdl64", seed=20181004, n_places=70000)
    descs = gen.descs(0, a.n, R, S)
    peak = 6523.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    m = SCManager(capacity_hint=a.n + 8)
    m.append_descs(descs)
    qs = [a.n - 1 - 37 * i for i in range(a.queries)]
    for q in qs[:3]:
        m.exhaustive(q, a.n - 50)
    t_tot, t_scr, resc, res = [], [], [], []
    t0 = time.perf_counter()
    for q in qs:
        res.append(m.exhaustive(q, a.n - 50))
        tot, scr, _ = m.timing()
        t_tot.append(tot)
        t_scr.append(scr)
        resc.append(m.exhaustive_rescored())
    wall = time.perf_counter() - t0
    # throughput: all queries enqueued back to back, one synchronisation
    m.exhaustive_batched(qs, a.n - 50)
    t0 = time.perf_counter()
    bd, bs, bi = m.exhaustive_batched(qs, a.n - 50)
    wall_b = time.perf_counter() - t0
    ms_batched = m.timing()[0] / len(qs)
    same = all(res[i][1] == bs[i] and res[i][2] == bi[i] and res[i][0] == bd[i] for i in range(len(qs)))
    ms_scr, ms_tot = float(np.median(t_scr)), float(np.median(t_tot))
    algo = 4 * R * S * (a.n - 50)
    out = {"config": f"exhaustive_{a.n}_20x60", "queries": len(qs), "queries_per_sec_device": 1e3 / ms_tot,
           "queries_per_sec_wall": len(qs) / wall,
           "batched": {"queries_per_sec_device": 1e3 / ms_batched, "queries_per_sec_wall": len(qs) / wall_b, "ms_per_query_device": ms_batched,
                       "frac_of_hbm_roofline_whole_query": 4 * R * S * (a.n - 50) / (ms_batched * 1e-3) / 1e9 / peak, "equal_to_single": bool(same)}, "ms_per_query_device": ms_tot, "ms_screen_kernel": ms_scr,
           "roofline": {"kernel": "k_exh_screen", "bound": "hbm", "achieved": algo / (ms_scr * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": algo / (ms_scr * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": algo,
                        "frac_whole_query": algo / (ms_tot * 1e-3) / 1e9 / peak},
           "rescored_per_query_median": float(np.median(resc)), "rescored_max": int(max(resc))}
    if a.cpu:
        from oracle import oracle as orc
        kind = "reference" if orc.ref_available("default") else "port"
        obj = orc.Ref("default") if kind == "reference" else orc.Port()
        for d in descs[:a.cpu]:
            obj.append_desc(d.astype(np.float64))
        q = qs[0]
        if kind == "reference":
            sec, d, s, i = obj.time_exhaustive(descs[q].astype(np.float64), a.cpu)
        else:
            t = time.perf_counter()
            d, s, i, _ = obj.exhaustive(descs[q].astype(np.float64), a.cpu)
            sec = time.perf_counter() - t
        g = m.exhaustive(q, a.cpu)
        out["cpu_baseline"] = {"kind": kind, "cores": 1, "entries": a.cpu, "seconds": sec, "us_per_pair": 1e6 * sec / a.cpu,
                               "extrapolated_queries_per_sec_at_n": 1.0 / (sec / a.cpu * (a.n - 50)),
                               "winner_equals_gpu": bool(g[1] == s and g[2] == i and abs(g[0] - d) <= 1e-5 * abs(d) + 1e-9)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
