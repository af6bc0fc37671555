#!/usr/bin/env python
"""BASELINE configs 3 and 5 on one B200 (the sharded versions are exercised by tests/test_gpu_sharded.py and
bench.py --gpus N).  Prints one JSON object per config, each with a parity check against the reference compiled
with that config's constants (oracle/_ref/libscref_k50.so, libscref_40x120.so) or the oracle port.

  config 3: MulRan-scale OS1-64 -- 40,000-keyframe database, 20x60, K = 50 candidates; a batch of 512 full-size
            scans (65,536 points) replayed on top (append + detect each).
  config 5: 40x120 descriptors, 20,000 keyframes, exhaustive search forward + column-reversed ("flipped").
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from sc_lego_loam_b200.scgpu import SCManager  # noqa: E402
from sc_lego_loam_b200.synth import ScanGen  # noqa: E402


def config1(n=2000):
    """BASELINE config 1: the reference's own CPU-runnable case -- 2,000 synthetic HDL-64 scans (120,000 points,
    32-byte pcl::PointXYZI records), 20x60, 10 candidates, exclude-recent 50 -- run through the reference (one thread)
    and through the GPU path (sequential API semantics via one batched replay), EVERYTHING compared: loop id, yaw bits,
    nearest distance, and for a sample of queries the candidate lists with their squared distances, shifts, distances."""
    gen = ScanGen("hdl64", seed=20181001, n_places=(n * 10) // 13)
    kind = "reference" if orc.ref_available("default") else "port"
    ref = orc.Ref("default") if kind == "reference" else orc.Port()
    m = SCManager(capacity_hint=n + 8)
    chunk = 250
    ids, yaws, dists, cands = [], [], [], {}
    t_ref = t_gpu = 0.0
    sample = set(range(60, n, 97))
    for c0 in range(0, n, chunk):
        scans = gen.scans(c0, chunk, 8)
        t0 = time.perf_counter()
        out = m.replay(scans)
        t_gpu += time.perf_counter() - t0
        for q in range(chunk):
            if c0 + q in sample:
                cands[c0 + q] = m.candidates(q)
        t0 = time.perf_counter()
        for j, s in enumerate(scans):
            ref.append_scan(s)
            d = ref.detect(details=(c0 + j) in sample) if kind == "reference" else ref.detect()
            ids.append(d["loop_id"]); yaws.append(d["yaw"]); dists.append(d.get("min_dist", np.nan))
            if (c0 + j) in sample and d["k"]:
                g = cands[c0 + j]
                assert np.array_equal(g["cand_idx"], d["cand_idx"]) and np.array_equal(g["cand_shift"], d["cand_shift"]), c0 + j
                assert np.array_equal(g["cand_d2"].view(np.uint32), d["cand_d2"].view(np.uint32)), c0 + j
                assert np.allclose(g["cand_dist"], d["cand_dist"], rtol=1e-5, atol=1e-9, equal_nan=True), c0 + j
                assert g["n_tree"] == d["n_tree"]
        t_ref += time.perf_counter() - t0
        if c0 == 0:
            got = {k: [v.copy()] for k, v in out.items()}
        else:
            for k, v in out.items():
                got[k].append(v.copy())
    got = {k: np.concatenate(v) for k, v in got.items()}
    ids, yaws = np.array(ids), np.array(yaws, np.float32)
    res = {"config": "1: 2,000 HDL-64 scans (120k pts, PointXYZI stride), 20x60, K=10, exclude 50: reference vs GPU", "keyframes": n,
           "loop_ids_equal": bool(np.array_equal(ids, got["loop_id"])), "yaw_bits_equal": bool(np.array_equal(yaws.view(np.uint32), got["yaw"].view(np.uint32))),
           "loops_found": int((ids >= 0).sum()), "candidate_lists_checked": len(cands), "reference_kind": kind,
           "reference_seconds_1_core": t_ref, "reference_queries_per_sec": n / t_ref, "gpu_seconds_incl_pageable_h2d": t_gpu,
           "gpu_queries_per_sec_from_pageable_host": n / t_gpu}
    print(json.dumps(res))


def config3(n_db=40000, batch=512, check=48):
    gen = ScanGen("os1", seed=20181003, n_places=30000)
    descs = gen.descs(0, n_db, 20, 60, threads=16)
    scans = gen.scans(n_db, batch, 4)
    import torch
    m = SCManager(num_candidates=50, capacity_hint=n_db + batch + 8)
    m.append_descs(descs)
    d_scans = torch.from_numpy(scans).cuda()          # device-resident scans: the step is pure device time
    dev = (d_scans.data_ptr(), batch, scans.shape[1], 16, 1)
    for _ in range(2):
        m.truncate(n_db)
        m.replay(dev)
    ts = []
    for _ in range(5):
        m.truncate(n_db)
        out = m.replay(dev)
        ts.append(m.timing())
    tot, b, q = (float(np.median([t[i] for t in ts])) for i in range(3))
    res = {"config": "3: 40k OS1-64 keyframes, 20x60, K=50, 1 B200", "keyframes_per_step": batch, "ms_per_step": tot,
           "queries_per_sec": batch / (tot * 1e-3), "build_ms": b, "query_ms": q, "loops_found": int((out["loop_id"] >= 0).sum())}
    kind = "reference" if orc.ref_available("k50") else "port"
    ref = orc.Ref("k50") if kind == "reference" else orc.Port(orc.Params(num_candidates=50))
    for d in descs:
        ref.append_desc(d.astype(np.float64))
    t0 = time.perf_counter()
    ids, yaws = [], []
    for s in scans[:check]:
        ref.append_scan(s)
        d = ref.detect(details=False) if kind == "reference" else ref.detect()
        ids.append(d["loop_id"])
        yaws.append(d["yaw"])
    sec = time.perf_counter() - t0
    res["cpu"] = {"kind": kind, "cores": 1, "keyframes": check, "queries_per_sec": check / sec}
    res["first_%d_equal_reference" % check] = bool(np.array_equal(ids, out["loop_id"][:check]) and
                                                   np.array_equal(np.array(yaws, np.float32).view(np.uint32), out["yaw"][:check].view(np.uint32)))
    print(json.dumps(res))


def config5(n_db=20000, n_check=1500):
    R, S = 40, 120
    gen = ScanGen("hdl64", seed=20181005, n_places=15000)
    descs = gen.descs(0, n_db, R, S, threads=16)
    descs[n_db - 1] = descs[4242].reshape(S, R)[::-1].ravel()      # the query is a column-reversed revisit of entry 4242
    m = SCManager(num_ring=R, num_sector=S, capacity_hint=n_db + 8)
    m.append_descs(descs)
    q = n_db - 1
    m.exhaustive(q, n_db - 50, True)
    m.exhaustive(q, n_db - 50, False)       # builds the screening copy (lazy) -- not part of a query
    t0 = time.perf_counter()
    fwd = m.exhaustive(q, n_db - 50, False)
    t1 = time.perf_counter()
    ms_fwd_dev, ms_fwd_screen, _ = m.timing()
    fwd_rescored = m.exhaustive_rescored()
    both = m.exhaustive(q, n_db - 50, True)
    t2 = time.perf_counter()
    res = {"config": "5: 40x120, 20k keyframes, exhaustive search on 1 B200: FP32 screening (k_exh_screen<40,120,6>; the column-reversed "
                     "pass reads the candidate rows back to front) + exact FP64 rescoring of the survivors",
           "ms_per_query_forward": 1e3 * (t1 - t0), "ms_per_query_forward_device": ms_fwd_dev, "ms_forward_screen_kernel": ms_fwd_screen,
           "forward_rescored": fwd_rescored, "ms_per_query_forward_plus_flipped": 1e3 * (t2 - t1),
           "winner_forward": fwd, "winner_flipped_search": both, "finds_reversed_revisit": bool(both[2] == 4242 and both[3] == 1)}
    port = orc.Port(orc.Params(R=R, S=S))
    for d in descs[:n_check]:
        port.append_desc(d.astype(np.float64))
    t0 = time.perf_counter()
    want = port.exhaustive(descs[q].astype(np.float64), n_check, True)
    sec = time.perf_counter() - t0
    got = m.exhaustive(q, n_check, True)
    res["cpu"] = {"kind": "port (composed flipped oracle)", "cores": 1, "entries": n_check, "seconds": sec,
                  "extrapolated_seconds_per_query_at_20k": sec / n_check * (n_db - 50)}
    res["winner_equals_oracle_on_%d" % n_check] = bool(got[1:] == want[1:] and abs(got[0] - want[0]) <= 1e-5 * abs(want[0]) + 1e-9)
    print(json.dumps(res))


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "3", "5"]
    if "1" in which:
        config1()
    if "3" in which:
        config3()
    if "5" in which:
        config5()
