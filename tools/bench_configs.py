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
    t0 = time.perf_counter()
    fwd = m.exhaustive(q, n_db - 50, False)
    t1 = time.perf_counter()
    both = m.exhaustive(q, n_db - 50, True)
    t2 = time.perf_counter()
    res = {"config": "5: 40x120, 20k keyframes, exhaustive forward + flipped, 1 B200 (exact FP64 pair kernel for every entry)",
           "ms_per_query_forward": 1e3 * (t1 - t0), "ms_per_query_forward_plus_flipped": 1e3 * (t2 - t1),
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
    which = sys.argv[1:] or ["3", "5"]
    if "3" in which:
        config3()
    if "5" in which:
        config5()
