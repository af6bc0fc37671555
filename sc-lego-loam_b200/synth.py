"""Deterministic synthetic scans / descriptors (wrapper of csrc/scangen.c; host-only data generator).

Shapes per SURVEY.md 8(d): HDL-64 = 64 x 1875 = 120,000 points, OS1-64 = 64 x 1024 = 65,536 points.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build


class _Cfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_beams", C.c_int), ("n_azim", C.c_int), ("n_places", C.c_int),
                ("sensor_h", C.c_float), ("max_range", C.c_float), ("jitter", C.c_float), ("range_sigma", C.c_float)]


_lib = None


def _L():
    global _lib
    if _lib is None:
        path = _build.LIB_SCANGEN
        if not os.path.exists(path):
            _build.build_scangen()
        _lib = C.CDLL(path)
        _lib.scangen_default_cfg.argtypes = [C.POINTER(_Cfg), C.c_int]
        _lib.scangen_pose.argtypes = [C.POINTER(_Cfg), C.c_uint64, C.POINTER(C.c_int)] + [C.POINTER(C.c_float)] * 3
        _lib.scangen_scan.argtypes = [C.POINTER(_Cfg), C.c_uint64, C.c_void_p, C.c_size_t]
        _lib.scangen_desc.argtypes = [C.POINTER(_Cfg), C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    return _lib


class ScanGen:
    """sensor: 'hdl64' (120,000 pts) | 'os1' (65,536 pts); n_azim overrides the azimuth count (small tests)."""

    def __init__(self, sensor="hdl64", seed=20181001, n_places=1000, n_azim=None, n_beams=None, **kw):
        self.cfg = _Cfg()
        _L().scangen_default_cfg(C.byref(self.cfg), 1 if sensor == "hdl64" else 0)
        self.cfg.seed = seed
        self.cfg.n_places = n_places
        if n_azim:
            self.cfg.n_azim = n_azim
        if n_beams:
            self.cfg.n_beams = n_beams
        for k, v in kw.items():
            setattr(self.cfg, k, v)

    @property
    def points_per_scan(self):
        return self.cfg.n_beams * self.cfg.n_azim

    def pose(self, i):
        p, y, dx, dy = C.c_int(), C.c_float(), C.c_float(), C.c_float()
        _L().scangen_pose(C.byref(self.cfg), i, C.byref(p), C.byref(y), C.byref(dx), C.byref(dy))
        return p.value, y.value, dx.value, dy.value

    def scan(self, i, floats_per_point=4, out=None):
        """(points, floats_per_point) float32; x,y,z first, remaining lanes zero (4 = float4, 8 = pcl::PointXYZI)."""
        if out is None:
            out = np.empty((self.points_per_scan, floats_per_point), np.float32)
        assert out.flags.c_contiguous and out.dtype == np.float32
        _L().scangen_scan(C.byref(self.cfg), i, out.ctypes.data, floats_per_point * 4)
        return out

    def scans(self, start, count, floats_per_point=4, out=None):
        if out is None:
            out = np.empty((count, self.points_per_scan, floats_per_point), np.float32)
        for j in range(count):
            self.scan(start + j, floats_per_point, out[j])
        return out

    def desc(self, i, R=20, S=60):
        """Descriptor generated in descriptor space: R*S float32, column-major (element (r,c) at c*R+r)."""
        out = np.empty(R * S, np.float32)
        _L().scangen_desc(C.byref(self.cfg), i, R, S, out.ctypes.data)
        return out

    def descs(self, start, count, R=20, S=60, threads=0):
        """count descriptors; threads > 0 generates in parallel (the C generator releases the GIL)."""
        out = np.empty((count, R * S), np.float32)
        lib, cfg = _L(), C.byref(self.cfg)

        def run(lo, hi):
            for j in range(lo, hi):
                lib.scangen_desc(cfg, start + j, R, S, out[j].ctypes.data)

        if threads and count >= 4 * threads:
            from concurrent.futures import ThreadPoolExecutor
            edges = np.linspace(0, count, threads + 1).astype(int)
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(lambda k: run(int(edges[k]), int(edges[k + 1])), range(threads)))
        else:
            run(0, count)
        return out
