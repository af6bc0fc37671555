#!/usr/bin/env python
"""Full-shift exhaustive search (SEARCH_RATIO = 1, 20x60) over N keyframes: tensor-core screening (k_tc_fullshift, tcgen05
3xTF32) vs its FFMA2 SIMT counterpart (k_fullshift_simt; run with SCGPU_FULLSHIFT_SIMT=1).  Prints one JSON object.

    python tools/bench_fullshift.py [N] [Q]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_lego_loam_b200.scgpu import SCManager  # noqa: E402
from sc_lego_loam_b200.synth import ScanGen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 64
R, S = 20, 60
simt = os.environ.get("SCGPU_FULLSHIFT_SIMT", "0") not in ("", "0")
gen = ScanGen("hdl64", seed=20181003, n_places=int(n * 0.8))
descs = gen.descs(0, n, R, S, threads=min(16, os.cpu_count() or 1))
m = SCManager(search_ratio=1.0, capacity_hint=n + 8)
m.append_descs(descs)
qs = [n - 1 - 37 * i for i in range(Q)]
for _ in range(2):
    m.exhaustive_batched(qs, n - 50)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    d, s, i = m.exhaustive_batched(qs, n - 50)
    ts.append((time.perf_counter() - t0, m.timing()[0]))
wall, dev = (float(np.median([t[k] for t in ts])) for k in (0, 1))
single = []
for q in qs[:4]:
    m.exhaustive(q, n - 50)
    t = m.timing()
    single.append((t[0], t[1]))
pairs = Q * (n - 50)
flop = 2.0 * R * S * S * pairs                      # the contraction itself (one pass; 3xTF32 issues three)
res = {"kernel": "k_fullshift_simt (FFMA2)" if simt else "k_tc_fullshift (tcgen05 kind::tf32 x3, TMEM, TMA)",
       "db_keyframes": n, "queries_per_batch": Q, "ms_per_batch_device": dev, "ms_per_batch_wall": 1e3 * wall,
       "queries_per_sec": Q / (dev * 1e-3), "pairs_per_sec": pairs / (dev * 1e-3),
       "contraction_tflops_useful": flop / (dev * 1e-3) / 1e12,
       "tensor_tflops_issued": None if simt else 3 * flop / (dev * 1e-3) / 1e12,
       "ms_single_query": float(np.median([x[0] for x in single])), "ms_single_query_screen_kernel": float(np.median([x[1] for x in single])),
       "rescored_last": m.exhaustive_rescored(), "winner_sample": [float(d[0]), int(s[0]), int(i[0])]}
print(json.dumps(res))
