#!/usr/bin/env python
"""Round 1's multi-GPU step, kept for A/B (`bench.py --gpus N --mode allgather` under torchrun): the database sharded over
the ranks with three NCCL all_gathers per step (built records, per-shard top-K keys, per-shard best) and every rank running
the per-query kernels for ALL queries.  The default mode of bench.py is the peer-sharded database (include/scgpu.h)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ALGO_BYTES_PER_SCAN, DB_SIZE, PTS, R, S, WORKLOAD, ClockSampler, gen_scans, measured_peak, ncu_traffic  # noqa: E402


def run_multi_gpu_allgather(args):
    """STRONG scaling of the same job: the 4,541-keyframe run (rounded up to a multiple of G) replayed from an
    empty database, keyframe i built by and stored on rank i % G; every query searches all shards."""
    import torch
    import torch.distributed as dist
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.sharded import GpuStages, ShardedSearch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G = world
    total = min(args.batch, DB_SIZE)
    B = (total + G - 1) // G                     # keyframes per rank per step
    n0 = (DB_SIZE - total) // G * G              # pre-filled part when --batch < the whole run
    h_scans = torch.empty((B, PTS, 4), dtype=torch.float32, pin_memory=True)
    gen = gen_scans(n0 + np.arange(B) * G + rank, h_scans.numpy())
    m = SCManager(device=local, shard_rank=rank, shard_count=G, capacity_hint=n0 + G * B + 64)
    search = ShardedSearch(GpuStages(m, f"cuda:{local}"), rank, world)
    if n0:
        search.prefill_descs(gen.descs(0, n0, R, S))
    d_scans = h_scans.cuda()
    d_stage = torch.empty_like(d_scans)

    ns_cache = {}
    plan_uncached = search.st.plan_n_search

    def plan_cached(first_size, n):
        """Every step restarts from the same database size (truncate resets the snapshot state), so the n_search
        plan is the same device tensor each time: computing it once keeps the step free of host synchronisation."""
        if (first_size, n) not in ns_cache:
            ns_cache[(first_size, n)] = plan_uncached(first_size, n)
        return ns_cache[(first_size, n)]

    search.st.plan_n_search = plan_cached

    def step(e2e):
        m.truncate(n0)
        search.size = n0
        src = d_scans
        if e2e:
            d_stage.copy_(h_scans, non_blocking=True)
            src = d_stage
        r = search.step(src)
        if e2e:
            return {k: v.cpu() for k, v in r.items()}
        return r

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    def timed(e2e):
        for _ in range(args.warmup):
            step(e2e)
        torch.cuda.synchronize()
        dist.barrier()
        l0 = m.launch_count()
        clocks.active = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            r = step(e2e)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        clocks.active = False
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() * 1e-3, m.launch_count() - l0, r

    dt_dev, launches, r = timed(False)
    # roofline of the dominant kernel (k_build_tma: every rank bins its own B scans): CUDA events around the build
    # launch of each step of one more device-leg pass, on the stream it is launched on; max over ranks
    search.build_events = []
    timed(False)
    ms_build = torch.tensor([sum(a.elapsed_time(b) for a, b in search.build_events[-args.steps:]) / args.steps], device="cuda")
    search.build_events = None
    dist.all_reduce(ms_build, op=dist.ReduceOp.MAX)
    ms_build = ms_build.item()
    dt_e2e, _, r2 = timed(True)
    same = all(torch.equal(r[k].cpu(), r2[k]) for k in ("loop_id", "nn_idx", "nn_shift"))
    if rank == 0:
        clocks.stop()
        nq = G * B
        line = {
            "metric": "sc_loop_queries_per_sec", "value": nq * args.steps / dt_dev, "unit": "queries/s", "n_gpus": G,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 binning + f64 keys/distance", "data": "synthetic",
            "config": {"workload": WORKLOAD, "db_keyframes": n0 + nq, "keyframes_per_step": nq, "keyframes_per_gpu_per_step": B,
                       "points_per_scan": PTS, "point_stride_bytes": 16, "l2": "inputs larger than L2",
                       "parallelism": f"database sharded i%{G} over {G} gpus; each rank bins its own scans; 3 NCCL all_gathers per step"},
            "e2e": {"value": nq * args.steps / dt_e2e, "unit": "queries/s", "h2d_bytes_per_step": nq * PTS * 16 + nq * 8,
                    "d2h_bytes_per_step": nq * 24 * G, "ms_per_step": 1e3 * dt_e2e / args.steps, "results_equal_device_leg": bool(same)},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "roofline": {"kernel": "k_build_tma", "bound": "hbm", "achieved": ALGO_BYTES_PER_SCAN * B / (ms_build * 1e-3) / 1e9,
                         "peak": measured_peak()[0], "unit": "GB/s", "frac": ALGO_BYTES_PER_SCAN * B / (ms_build * 1e-3) / 1e9 / measured_peak()[0],
                         "traffic": (ncu_traffic()[0] * B if ncu_traffic() else None), "peak_source": measured_peak()[1], "per": "gpu (max over ranks of the build launch time)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SCAN * B, "ms_per_launch": ms_build},
            "stages": {"loops_found": int((r["loop_id"] >= 0).sum().item()), "build_ms_per_step": ms_build},
        }
        print(json.dumps(line))
    dist.destroy_process_group()


