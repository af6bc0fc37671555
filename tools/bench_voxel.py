#!/usr/bin/env python
"""voxel-grid downsample (pcl::VoxelGrid, leaf 0.5 m, mapOptmization.cpp:264,1235-1237) fused in
front of the descriptor build, on N synthetic HDL-64 scans resident in HBM.  Prints one JSON object: scans/s of
k_build_voxel next to the plain k_build_tma on the same scans, the fraction of the HBM roofline (algorithmic bytes = one
read of the points, 16 B each: the min/max pass and the accumulate pass share the read through L2) and the CPU oracle
(restated PCL filter + reference descriptor) on a few scans."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=1184)
    ap.add_argument("--leaf", type=float, default=0.5)
    ap.add_argument("--cpu", type=int, default=8)
    a = ap.parse_args()
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from sc_lego_loam_b200.scgpu import SCManager
    from sc_lego_loam_b200.synth import ScanGen
    PTS = 120000
    gen = ScanGen("hdl64", seed=20181002, n_places=3500)
    host = np.empty((a.scans, PTS, 4), np.float32)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda j: gen.scan(j, 4, host[j]), range(a.scans)))
    dev = torch.from_numpy(host).cuda()
    peak = 6523.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"config": f"voxel_grid_leaf{a.leaf}_hdl64_{a.scans}_scans", "scans": a.scans}
    for name, leaf in (("plain", 0.0), ("voxel", a.leaf)):
        m = SCManager(capacity_hint=a.scans * 6 + 8)
        m.set_downsample_leaf(leaf)
        ms = []
        for _ in range(5):
            m.truncate(0)
            m.append_scans((dev.data_ptr(), a.scans, PTS, 16, 1))
            ms.append(m.timing()[1])
        t = float(np.median(ms[2:]))
        out[name] = {"ms": t, "scans_per_sec": a.scans / (t * 1e-3), "read_gbs": a.scans * PTS * 16 / (t * 1e-3) / 1e9,
                     "frac_of_hbm_roofline": a.scans * PTS * 16 / (t * 1e-3) / 1e9 / peak}
        if leaf:
            v = m.voxel_downsample(host[0], leaf)
            out[name]["voxels_scan0"] = int(len(v["idx"]))
            out[name]["passes_scan0"] = int(v["passes"])
        m.close()
    if a.cpu:
        from oracle import oracle as orc
        vox, port = orc.Voxel(), orc.Port()
        t0 = time.perf_counter()
        for j in range(a.cpu):
            ds = vox.downsample(host[j], a.leaf)["points"]
        t1 = time.perf_counter()
        for j in range(a.cpu):
            port.make_sc(ds)
        t2 = time.perf_counter()
        out["cpu_baseline"] = {"kind": "port (restated pcl::VoxelGrid + descriptor)", "cores": 1, "scans": a.cpu,
                               "voxel_ms_per_scan": 1e3 * (t1 - t0) / a.cpu, "descriptor_ms_per_downsampled_scan": 1e3 * (t2 - t1) / a.cpu,
                               "scans_per_sec": a.cpu / (t2 - t0)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
