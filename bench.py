#!/usr/bin/env python
"""bench.py -- Scan Context loop-closure hot path on B200: loop queries/s (each = one keyframe through
makeAndSaveScancontextAndKeys + detectLoopClosureID), next to the reference CPU SCManager.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl scgpu|reference] [--sweep]

Workload (BASELINE.json configs[1]): KITTI-00-shaped synthetic run -- a database of 4,541 keyframes, HDL-64
scans (64 x 1875 = 120,000 points, float4 records), 20x60 descriptor, 10 candidates, exclude-recent 50.  One
"step" replays the LAST `batch` keyframes of that run: for each of them descriptor build (k_build) + keys +
append + exact ring-key top-10 (k_topk) + column-shifted cosine distance of the 10 candidates (k_score) + argmin,
threshold and yaw, with the reference's periodic tree-snapshot semantics.  The database is truncated back to
4,541 - batch before every step, so every step does identical work.

  value   : keyframes(queries)/s with the scans already resident in HBM (device leg)
  e2e     : the same through the C ABI with HOST (pinned) scan buffers: H2D of the scans and D2H of the results
            inside the timed region
  roofline: k_build, the dominant kernel: algorithmic bytes (16 B/point + descriptor record) / CUDA-event time
            of the k_build launches, against the measured copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the reference's own Scancontext.cpp compiled verbatim (oracle/_ref), or the
            oracle port when that library is absent, on the host cores of this box.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DB_SIZE = 4541
PTS = 120000
R, S, K_CAND = 20, 60, 10
ALGO_BYTES_PER_SCAN = 16 * PTS + 4 * R * S + 4 * R + 8 * S   # 1,925,360 B
SEED = 20181002
WORKLOAD = "kitti00_shaped_4541kf_hdl64_120kpts_sc20x60_k10_excl50"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="scgpu", choices=["scgpu", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="keyframes replayed per step (per GPU)")
    ap.add_argument("--sweep", action="store_true", help="also time query-only throughput vs database size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self._stop, self._t, self.active = index, [], threading.Event(), None, False

    def _run(self):
        while not self._stop.is_set():
            if self.active:
                try:
                    o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                       capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append([x.strip() for x in o])
                except Exception:
                    pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_inputs(batch, first_scan, n_db_prefill, stride_floats=4, scan_step=1, scan_offset=0):
    from sc_lego_loam_b200.synth import ScanGen
    gen = ScanGen("hdl64", seed=SEED, n_places=3500)
    descs = gen.descs(0, n_db_prefill, R, S)
    scans = np.empty((batch, PTS, stride_floats), np.float32)
    for j in range(batch):
        gen.scan(first_scan + j * scan_step + scan_offset, stride_floats, scans[j])
    return descs, scans


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------------
_W = {}


def _ref_worker_init(n_prefill, n_scans, worker_seed_base):
    from oracle import oracle as orc
    from sc_lego_loam_b200.synth import ScanGen
    kind = "reference" if orc.ref_available("default") else "port"
    gen = ScanGen("hdl64", seed=SEED, n_places=3500)
    obj = orc.Ref("default") if kind == "reference" else orc.Port()
    for d in gen.descs(0, n_prefill, R, S):
        obj.append_desc(d.astype(np.float64))
    wid = os.getpid()
    scans = gen.scans(n_prefill + (wid % 64) * n_scans, n_scans, 4)
    _W.update(obj=obj, scans=scans, kind=kind)


def _ref_worker_step(_):
    obj, scans = _W["obj"], _W["scans"]
    t0 = time.perf_counter()
    if _W["kind"] == "reference":
        tb, td = obj.time_run(scans)
    else:
        tb = td = 0.0
        for s in scans:
            a = time.perf_counter()
            obj.append_scan(s)
            b = time.perf_counter()
            obj.detect()
            tb += b - a
            td += time.perf_counter() - b
    return time.perf_counter() - t0, tb, td, len(scans)


def run_reference_arm(args):
    """The reference's own CPU SCManager, one independent replica per host core (the reference is single-threaded;
    SCManager has no internal parallelism), each with the same 4,541-keyframe database."""
    import multiprocessing as mp
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    per = 6   # keyframes per replica per step: ~55 ms of CPU work each
    kind = "reference" if orc.ref_available("default") else "port"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_worker_init, initargs=(DB_SIZE - per, per, 0)) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker_step, range(cores), chunksize=1)
        t0 = time.perf_counter()
        tb = td = 0.0
        for _ in range(args.steps):
            res = pool.map(_ref_worker_step, range(cores), chunksize=1)
            tb += sum(r[1] for r in res)
            td += sum(r[2] for r in res)
        dt = time.perf_counter() - t0
    n = cores * per * args.steps
    val = n / dt
    line = {"metric": "sc_loop_queries_per_sec", "value": val, "unit": "queries/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "db_keyframes": DB_SIZE, "replicas": cores, "keyframes_per_replica_per_step": per},
            "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": kind,
                             "sample": f"{cores} independent SCManager replicas x {per} keyframes/step x {args.steps} steps; "
                                       f"each keyframe = makeAndSaveScancontextAndKeys(120k pts) + detectLoopClosureID over a {DB_SIZE}-keyframe DB; "
                                       f"per-keyframe CPU time build {1e3 * tb / n:.2f} ms + detect {1e3 * td / n:.2f} ms"},
            "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_single(n_prefill, scans, passes):
    """One thread of the reference (kind 'reference') or the oracle port on a bounded sample of the workload."""
    from oracle import oracle as orc
    from sc_lego_loam_b200.synth import ScanGen
    kind = "reference" if orc.ref_available("default") else "port"
    obj = orc.Ref("default") if kind == "reference" else orc.Port()
    for d in ScanGen("hdl64", seed=SEED, n_places=3500).descs(0, n_prefill, R, S):
        obj.append_desc(d.astype(np.float64))
    tb = td = 0.0
    n = 0
    for _ in range(passes):
        if kind == "reference":
            a, b = obj.time_run(scans)
        else:
            a = b = 0.0
            for s in scans:
                t0 = time.perf_counter()
                obj.append_scan(s)
                t1 = time.perf_counter()
                obj.detect()
                a += t1 - t0
                b += time.perf_counter() - t1
        tb, td, n = tb + a, td + b, n + len(scans)
    return {"value": n / (tb + td), "unit": "queries/s", "cores": 1, "kind": kind,
            "sample": f"{n} keyframes (the step's first {len(scans)} scans x {passes} passes), DB {n_prefill}->{n_prefill + n}; "
                      f"build {1e3 * tb / n:.2f} ms + detect {1e3 * td / n:.2f} ms per keyframe; host has {os.cpu_count()} cores"}


# ------------------------------------------------------------------------------------------------------
def run_single_gpu(args):
    import torch
    from sc_lego_loam_b200.scgpu import SCManager
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scgpu path has no CPU fallback")
    B = args.batch
    n0 = DB_SIZE - B
    descs, scans = make_inputs(B, n0, n0)
    m = SCManager(device=0, capacity_hint=DB_SIZE + 64)
    m.append_descs(descs)
    torch.cuda.set_device(0)
    d_scans = torch.from_numpy(scans).cuda()
    h_scans = torch.from_numpy(scans).pin_memory()
    out = dict(loop_id=np.empty(B, np.int32), yaw=np.empty(B, np.float32), min_dist=np.empty(B, np.float64),
               nn_idx=np.empty(B, np.int32), nn_shift=np.empty(B, np.int32))

    def step(ptr, loc):
        m.truncate(n0)
        m.replay((ptr, B, PTS, 16, loc), out)

    clocks = ClockSampler(0)
    clocks.start()

    def timed(ptr, loc):
        for _ in range(args.warmup):
            step(ptr, loc)
        torch.cuda.synchronize()
        l0 = m.launch_count()
        tb = tq = 0.0
        clocks.active = True
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(ptr, loc)          # returns after the results are on the host (stream synchronised inside)
            _, b, q = m.timing()
            tb, tq = tb + b, tq + q
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clocks.active = False
        return dt, tb / args.steps, tq / args.steps, m.launch_count() - l0

    dt_dev, ms_build, ms_query, launches = timed(d_scans.data_ptr(), 1)
    res_dev = {k: v.copy() for k, v in out.items()}
    dt_e2e, ms_build_e2e, _, _ = timed(h_scans.data_ptr(), 0)
    same = all(np.array_equal(res_dev[k], out[k], equal_nan=True) for k in out)
    clocks.stop()

    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_SCAN * B / (ms_build * 1e-3) / 1e9
    line = {
        "metric": "sc_loop_queries_per_sec", "value": B * args.steps / dt_dev, "unit": "queries/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 binning + f64 keys/distance", "data": "synthetic",
        "config": {"workload": WORKLOAD, "db_keyframes": DB_SIZE, "keyframes_per_step": B, "points_per_scan": PTS,
                   "point_stride_bytes": 16, "l2": f"inputs larger than L2 ({B * PTS * 16 >> 20} MiB of points per step)",
                   "parallelism": "1 gpu"},
        "e2e": {"value": B * args.steps / dt_e2e, "unit": "queries/s", "h2d_bytes_per_step": B * PTS * 16 + B * 8,
                "d2h_bytes_per_step": B * 24, "ms_per_step": 1e3 * dt_e2e / args.steps, "results_equal_device_leg": bool(same)},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_build", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SCAN * B,
                     "ms_per_launch": ms_build},
        "stages": {"build_ms_per_step": ms_build, "query_ms_per_step": ms_query, "builds_per_sec": B / (ms_build * 1e-3),
                   "queries_only_per_sec": B / (ms_query * 1e-3), "loops_found": int((res_dev["loop_id"] >= 0).sum())},
        "clocks": clocks.summary(),
    }
    if args.sweep:
        line["sweep"] = db_size_sweep(m, torch)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single(n0, scans[:64], 3)
    print(json.dumps(line))


def db_size_sweep(m, torch):
    """Query-only throughput (k_topk + k_score + argmin for 256 stored keyframes) vs database size."""
    from sc_lego_loam_b200.synth import ScanGen
    gen = ScanGen("hdl64", seed=SEED + 1, n_places=80000)
    out = []
    m.truncate(0)
    have = 0
    for n in (1000, 2000, 5000, 10000, 20000, 50000, 100000):
        m.append_descs(gen.descs(have, n - have, R, S))
        have = n
        nq = 256
        for _ in range(3):
            m.query_batched(n - nq, nq)
        ts = []
        for _ in range(10):
            m.query_batched(n - nq, nq)
            ts.append(m.timing()[2])
        ms = float(np.median(ts))
        out.append({"db": n, "queries_per_sec": nq / (ms * 1e-3), "ms_per_256_queries": ms,
                    "ringkey_stream_gbs": nq * n * 4 * R / (ms * 1e-3) / 1e9})
    return out


def run_multi_gpu(args):
    import torch
    import torch.distributed as dist
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.sharded import GpuStages, ShardedSearch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G, B = world, args.batch
    n0 = (DB_SIZE - B) // G * G       # database before the step (a multiple of G keeps shards equal)
    descs, scans = make_inputs(B, n0, n0, scan_step=G, scan_offset=rank)
    m = SCManager(device=local, shard_rank=rank, shard_count=G, capacity_hint=DB_SIZE + G * B + 64)
    search = ShardedSearch(GpuStages(m, f"cuda:{local}"), rank, world)
    search.prefill_descs(descs)
    d_scans = torch.from_numpy(scans).cuda()
    h_scans = torch.from_numpy(scans).pin_memory()
    d_stage = torch.empty_like(d_scans)

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
    dt_e2e, _, r2 = timed(True)
    if rank == 0:
        clocks.stop()
        nq = G * B
        line = {
            "metric": "sc_loop_queries_per_sec", "value": nq * args.steps / dt_dev, "unit": "queries/s", "n_gpus": G,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 binning + f64 keys/distance", "data": "synthetic",
            "config": {"workload": WORKLOAD, "db_keyframes": n0 + nq, "keyframes_per_step": nq, "keyframes_per_gpu_per_step": B,
                       "points_per_scan": PTS, "point_stride_bytes": 16, "l2": "inputs larger than L2",
                       "parallelism": f"database sharded i%{G} over {G} gpus; 3 NCCL all_gathers per step"},
            "e2e": {"value": nq * args.steps / dt_e2e, "unit": "queries/s", "h2d_bytes_per_step": nq * PTS * 16 + nq * 8,
                    "d2h_bytes_per_step": nq * 24 * G, "ms_per_step": 1e3 * dt_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
            "stages": {"loops_found": int((r["loop_id"] >= 0).sum().item())},
        }
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference_arm(args)
        return
    if world > 1:
        run_multi_gpu(args)
    else:
        run_single_gpu(args)


if __name__ == "__main__":
    main()
