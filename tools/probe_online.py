#!/usr/bin/env python
"""Where the time of one online keyframe goes (scgpu_append_scan of a pageable host scan + scgpu_detect)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sc_lego_loam_b200.scgpu import SCManager
from sc_lego_loam_b200.synth import ScanGen

gen = ScanGen("hdl64", seed=3, n_places=3000)
m = SCManager(capacity_hint=8192)
m.append_descs(gen.descs(0, 4600, 20, 60, threads=8))
base = gen.scans(0, 8, 4)
for name, k, pin in (("pageable32", 8, False), ("pageable16", 4, False), ("pinned16", 4, True), ("pinned32", 8, True)):
    src = np.zeros((8, base.shape[1], k), np.float32)
    src[:, :, :3] = base[:, :, :3]
    if pin:
        t = torch.from_numpy(src).pin_memory()
        src = t.numpy()
    ta, td = [], []
    for i in range(60):
        t0 = time.perf_counter()
        m.makeAndSaveScancontextAndKeys(src[i % 8])
        t1 = time.perf_counter()
        m.detectLoopClosureID()
        t2 = time.perf_counter()
        ta.append(t1 - t0); td.append(t2 - t1)
    print(f"{name:12s} threads={os.environ.get('SCGPU_HOST_THREADS','auto'):>4s} append {1e6*np.median(ta[10:]):7.1f} us  detect {1e6*np.median(td[10:]):7.1f} us")
