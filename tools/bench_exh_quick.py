#!/usr/bin/env python
"""Config 4 quick look: windowed exhaustive search over N keyframes (20x60), Q = 1 and Q = 64, one GPU."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sc_lego_loam_b200.scgpu import SCManager
from sc_lego_loam_b200.synth import ScanGen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
R, S = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20, 60)
gen = ScanGen("hdl64", seed=bench.SEED + 1, n_places=int(n * 0.8))
descs = gen.descs(0, n, R, S, threads=min(16, os.cpu_count() or 1))
m = SCManager(num_ring=R, num_sector=S, capacity_hint=n + 8)
m.append_descs(descs)
peak, _ = bench.measured_peak()
qs = [n - 1 - 37 * i for i in range(64)]
bench.R, bench.S = R, S
print(json.dumps(bench.exhaustive_extra(m, n, peak, qs)))
