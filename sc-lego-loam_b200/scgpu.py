"""ctypes binding of the C ABI (include/scgpu.h) and a Python mirror of the reference's SCManager surface.

The class below keeps the reference's method names and argument meaning (Scancontext.h:63-73) so that parity
tests read like calls on the reference object.  Every method goes through libscgpu.so; there is no Python or
CPU implementation of any stage -- if the library or a CUDA device is missing, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

OK = 0
FLAG_FRESH_TREE = 1
FLAG_EXACT_BINNING = 2
FLAG_NO_SCREENING = 4
FLAG_NO_TMA_BUILD = 8
FLAG_PEER = 16
FLAG_INTENSITY = 32


class ScgpuError(RuntimeError):
    pass


class Config(C.Structure):
    """scgpu_config: the reference's constants (Scancontext.h:77-96) + placement."""
    _fields_ = [("num_ring", C.c_int32), ("num_sector", C.c_int32), ("lidar_height", C.c_double),
                ("max_radius", C.c_double), ("exclude_recent", C.c_int32), ("num_candidates", C.c_int32),
                ("search_ratio", C.c_double), ("dist_thres", C.c_double), ("tree_period", C.c_int32),
                ("device", C.c_int32), ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
                ("capacity_hint", C.c_uint64), ("flags", C.c_uint32), ("n_devices", C.c_int32),
                ("devices", C.c_int32 * 8)]


class IcpParams(C.Structure):
    """scgpu_icp_params: pcl::IterativeClosestPoint settings of mapOptmization.cpp:1053-1058 + the acceptance threshold."""
    _fields_ = [("max_iterations", C.c_int32), ("seed_axis", C.c_int32), ("max_correspondence_distance", C.c_double),
                ("transformation_epsilon", C.c_double), ("euclidean_fitness_epsilon", C.c_double),
                ("fitness_threshold", C.c_double), ("seed_angle", C.c_float), ("reserved", C.c_float)]


_vp, _sz, _i, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
_pi, _pf, _pd = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)

# name -> (argtypes); every function returns int status unless noted
_SIGNATURES = {
    "scgpu_default_config": [C.POINTER(Config)],
    "scgpu_create": [C.POINTER(Config), C.POINTER(_vp)],
    "scgpu_destroy": [_vp],
    "scgpu_make_sc": [_vp, _vp, _sz, _sz, _vp],
    "scgpu_ringkey": [_vp, _vp, _vp],
    "scgpu_sectorkey": [_vp, _vp, _vp],
    "scgpu_fast_align": [_vp, _vp, _vp, _pi],
    "scgpu_dist_direct": [_vp, _vp, _vp, _pd],
    "scgpu_distance": [_vp, _vp, _vp, _pd, _pi],
    "scgpu_append_scan": [_vp, _vp, _sz, _sz],
    "scgpu_detect": [_vp, _pi, _pf, _pd, _pi, _pi],
    "scgpu_size": [_vp, C.POINTER(_u64)],
    "scgpu_append_scans_batched": [_vp, _vp, _sz, _sz, _sz, _i],
    "scgpu_append_descs": [_vp, _vp, _sz],
    "scgpu_replay_batched": [_vp, _vp, _sz, _sz, _sz, _i, _vp, _vp, _vp, _vp, _vp],
    "scgpu_query_batched": [_vp, _u64, _sz, _vp, _vp, _vp, _vp, _vp],
    "scgpu_replay_async": [_vp, _vp, _sz, _sz, _sz, _i],
    "scgpu_replay_results": [_vp, _sz, _vp, _vp, _vp, _vp, _vp],
    "scgpu_default_icp_params": [C.POINTER(IcpParams)],
    "scgpu_verify_loop": [_vp, _vp, _sz, _vp, _sz, _sz, C.POINTER(IcpParams), _vp, _pd, _pi, _pi, _pi],
    "scgpu_host_info": [_pi, _pi],
    "scgpu_assemble_submap": [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _i, C.c_float, _vp, _sz, C.POINTER(_sz)],
    "scgpu_verify_loop_keyframes": [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _sz, _sz, _sz, C.c_float, C.POINTER(IcpParams), _vp, _pd, _pi, _pi, _pi,
                                    C.POINTER(_sz), C.POINTER(_sz)],
    "scgpu_growth_stats": [_vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(_u64)],
    "scgpu_timer_start": [_vp],
    "scgpu_timer_stop": [_vp, _pd],
    "scgpu_peer_partition": [_u64, _sz, _i, _i, C.POINTER(_sz), C.POINTER(_sz)],
    "scgpu_peer_export": [_vp, _vp, _sz],
    "scgpu_peer_attach": [_vp, _vp, _i],
    "scgpu_peer_replay_async": [_vp, _vp, _sz, _sz, _sz, _i],
    "scgpu_peer_barrier": [_vp, _vp],
    "scgpu_stage_exhaustive_exact": [_vp, _vp, _u64, _vp],
    "scgpu_get_candidates": [_vp, _vp, _vp, _vp, _vp, C.POINTER(_u64)],
    "scgpu_get_batch_candidates": [_vp, _sz, _vp, _vp, _vp, _vp, C.POINTER(_u64)],
    "scgpu_get_entry": [_vp, _u64, _vp, _vp, _vp],
    "scgpu_truncate": [_vp, _u64],
    "scgpu_exhaustive": [_vp, _u64, _u64, _i, _pd, _pi, C.POINTER(C.c_int64), _pi],
    "scgpu_exhaustive_batched": [_vp, _vp, _vp, _sz, _vp, _vp, _vp],
    "scgpu_exhaustive_stats": [_vp, C.POINTER(_u64)],
    "scgpu_probe_screen": [_vp, _u64, _u64, _vp, _vp],
    "scgpu_stage_exhaustive": [_vp, _vp, _sz, _vp, _vp, _vp],
    "scgpu_stage_exhaustive2": [_vp, _vp, _sz, _vp, _i, _vp, _vp],
    "scgpu_stage_gather": [_vp, _u64, _vp, _vp],
    "scgpu_set_downsample_leaf": [_vp, C.c_float],
    "scgpu_voxel_downsample": [_vp, _vp, _sz, _sz, C.c_float, _vp, _vp, _sz, C.POINTER(_sz), _vp, _vp, _pi],
    "scgpu_save": [_vp, C.c_char_p],
    "scgpu_load": [_vp, C.c_char_p],
    "scgpu_record_bytes": [_vp, C.POINTER(_sz)],
    "scgpu_stage_build": [_vp, _vp, _sz, _sz, _sz, _vp, _vp],
    "scgpu_stage_append": [_vp, _vp, _u64, _u64, _sz, _vp],
    "scgpu_stage_set_size": [_vp, _u64],
    "scgpu_get_timing": [_vp, _pd, _pd, _pd],
    "scgpu_stage_topk": [_vp, _vp, _sz, _vp, _vp, _vp],
    "scgpu_stage_merge": [_vp, _vp, _i, _sz, _vp, _vp],
    "scgpu_stage_score": [_vp, _vp, _sz, _vp, _vp, _vp, _vp],
    "scgpu_stage_finalize": [_vp, _vp, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "scgpu_plan_n_search": [_vp, _u64, _sz, _vp],
    "scgpu_launch_count": [_vp, C.POINTER(_u64)],
    "scgpu_xy2theta": [C.c_float, C.c_float, _pf],
    "scgpu_probe_atanf": [_vp, _sz, _vp],
    "scgpu_probe_bins": [_vp, _vp, _sz, _vp, _vp, _vp],
    "scgpu_probe_selfcheck": [_vp, _u64, _u64, _i, C.POINTER(_u64), C.POINTER(_u64), _vp],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["scgpu_last_error", "scgpu_version"])

_lib = None


def load_library(build_if_missing=False):
    """dlopen libscgpu.so (built in-tree by build.py).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_SCGPU
    if not os.path.exists(path):
        if not build_if_missing:
            raise ScgpuError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        _build.build_scgpu()
    lib = C.CDLL(path)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    lib.scgpu_last_error.restype = C.c_char_p
    lib.scgpu_version.restype = C.c_char_p
    _lib = lib
    return lib


def _check(rc):
    if rc != OK:
        raise ScgpuError(f"scgpu error {rc}: {load_library().scgpu_last_error().decode()}")


def _pts(pts):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise ValueError("points must be (n, k>=3) float32")
    return pts, pts.ctypes.data, pts.shape[0], pts.shape[1] * 4


def _d(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


class SCManager:
    """Mirror of the reference's ``class SCManager`` (Scancontext.h:58-108) on one B200.

    Matrices are numpy float64 arrays in the reference's memory layout (column-major R x S, i.e. a flat array
    with element (ring r, sector c) at c*R + r); ``as_matrix`` reshapes to (R, S).
    """

    def __init__(self, **cfg):
        self.lib = load_library()
        self.cfg = Config()
        _check(self.lib.scgpu_default_config(C.byref(self.cfg)))
        devices = cfg.pop("devices", None)
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise TypeError(f"unknown config field {k}")
            setattr(self.cfg, k, v)
        if devices is not None:      # device list: ONE handle drives a database sharded over these GPUs (entry i on devices[i % n])
            devices = [int(d) for d in devices]
            if not 1 <= len(devices) <= 8:
                raise ValueError("1 to 8 devices")
            self.cfg.n_devices = len(devices)
            for i, d in enumerate(devices):
                self.cfg.devices[i] = d
        self.h = _vp()
        _check(self.lib.scgpu_create(C.byref(self.cfg), C.byref(self.h)))
        self.R, self.S, self.K = self.cfg.num_ring, self.cfg.num_sector, self.cfg.num_candidates
        # the reference's public constants, same names
        self.LIDAR_HEIGHT = self.cfg.lidar_height
        self.PC_NUM_RING, self.PC_NUM_SECTOR = self.R, self.S
        self.PC_MAX_RADIUS = self.cfg.max_radius
        self.PC_UNIT_SECTORANGLE = 360.0 / float(self.S)
        self.PC_UNIT_RINGGAP = self.cfg.max_radius / float(self.R)
        self.NUM_EXCLUDE_RECENT = self.cfg.exclude_recent
        self.NUM_CANDIDATES_FROM_TREE = self.K
        self.SEARCH_RATIO = self.cfg.search_ratio
        self.SC_DIST_THRES = self.cfg.dist_thres
        self.TREE_MAKING_PERIOD_ = self.cfg.tree_period

    def close(self):
        if getattr(self, "h", None):
            self.lib.scgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def as_matrix(self, flat):
        return np.asarray(flat).reshape(self.S, self.R).T

    # ---- the reference's eight methods ---------------------------------------------------------
    def makeScancontext(self, scan):
        pts, ptr, n, stride = _pts(scan)
        out = np.empty(self.R * self.S, np.float64)
        _check(self.lib.scgpu_make_sc(self.h, ptr, n, stride, out.ctypes.data))
        return out

    def makeRingkeyFromScancontext(self, desc):
        desc = _d(desc, self.R * self.S)
        out = np.empty(self.R, np.float64)
        _check(self.lib.scgpu_ringkey(self.h, desc.ctypes.data, out.ctypes.data))
        return out

    def makeSectorkeyFromScancontext(self, desc):
        desc = _d(desc, self.R * self.S)
        out = np.empty(self.S, np.float64)
        _check(self.lib.scgpu_sectorkey(self.h, desc.ctypes.data, out.ctypes.data))
        return out

    def fastAlignUsingVkey(self, vkey1, vkey2):
        a, b = _d(vkey1, self.S), _d(vkey2, self.S)
        s = C.c_int()
        _check(self.lib.scgpu_fast_align(self.h, a.ctypes.data, b.ctypes.data, C.byref(s)))
        return s.value

    def distDirectSC(self, sc1, sc2):
        a, b = _d(sc1, self.R * self.S), _d(sc2, self.R * self.S)
        d = C.c_double()
        _check(self.lib.scgpu_dist_direct(self.h, a.ctypes.data, b.ctypes.data, C.byref(d)))
        return d.value

    def distanceBtnScanContext(self, sc1, sc2):
        a, b = _d(sc1, self.R * self.S), _d(sc2, self.R * self.S)
        d, s = C.c_double(), C.c_int()
        _check(self.lib.scgpu_distance(self.h, a.ctypes.data, b.ctypes.data, C.byref(d), C.byref(s)))
        return d.value, s.value

    def makeAndSaveScancontextAndKeys(self, scan):
        pts, ptr, n, stride = _pts(scan)
        _check(self.lib.scgpu_append_scan(self.h, ptr, n, stride))

    def detectLoopClosureID(self, details=False):
        lid, yaw, nd, ni, ns = C.c_int(), C.c_float(), C.c_double(), C.c_int(), C.c_int()
        _check(self.lib.scgpu_detect(self.h, C.byref(lid), C.byref(yaw), C.byref(nd), C.byref(ni), C.byref(ns)))
        if not details:
            return lid.value, np.float32(yaw.value)
        return dict(loop_id=lid.value, yaw=np.float32(yaw.value), min_dist=nd.value, nn_idx=ni.value,
                    nn_shift=ns.value)

    # ---- extras ------------------------------------------------------------------------------------
    def size(self):
        n = _u64()
        _check(self.lib.scgpu_size(self.h, C.byref(n)))
        return n.value

    def launch_count(self):
        n = _u64()
        _check(self.lib.scgpu_launch_count(self.h, C.byref(n)))
        return n.value

    @staticmethod
    def _scans(scans):
        """(n_scans, pts, k) float32 numpy array, or (ptr, n_scans, pts, stride, location) tuple."""
        if isinstance(scans, tuple):
            return None, scans
        scans = np.ascontiguousarray(scans, dtype=np.float32)
        if scans.ndim != 3 or scans.shape[2] < 3:
            raise ValueError("scans must be (n_scans, pts, k>=3) float32")
        return scans, (scans.ctypes.data, scans.shape[0], scans.shape[1], scans.shape[2] * 4, 0)

    def append_scans(self, scans):
        keep, (ptr, n, pts, stride, loc) = self._scans(scans)
        _check(self.lib.scgpu_append_scans_batched(self.h, ptr, n, pts, stride, loc))

    def append_descs(self, descs):
        descs = np.ascontiguousarray(descs, dtype=np.float32)
        n = descs.size // (self.R * self.S)
        _check(self.lib.scgpu_append_descs(self.h, descs.ctypes.data, n))

    def replay(self, scans, out=None):
        """The bench step: append each scan and detect after it; returns dict of arrays (n_scans,)."""
        keep, (ptr, n, pts, stride, loc) = self._scans(scans)
        if out is None:
            out = dict(loop_id=np.empty(n, np.int32), yaw=np.empty(n, np.float32), min_dist=np.empty(n, np.float64),
                       nn_idx=np.empty(n, np.int32), nn_shift=np.empty(n, np.int32))
        _check(self.lib.scgpu_replay_batched(self.h, ptr, n, pts, stride, loc, out["loop_id"].ctypes.data,
                                             out["yaw"].ctypes.data, out["min_dist"].ctypes.data,
                                             out["nn_idx"].ctypes.data, out["nn_shift"].ctypes.data))
        return out

    def replay_async(self, scans):
        """Enqueue the bench step and return; ``replay_results`` waits for the last enqueued step.  Device-resident
        scans (tuple form, location 1) must stay valid until then; host scans too."""
        keep, (ptr, n, pts, stride, loc) = self._scans(scans)
        self._keep = keep
        _check(self.lib.scgpu_replay_async(self.h, ptr, n, pts, stride, loc))
        return n

    def replay_results(self, n, out=None):
        if out is None:
            out = dict(loop_id=np.empty(n, np.int32), yaw=np.empty(n, np.float32), min_dist=np.empty(n, np.float64),
                       nn_idx=np.empty(n, np.int32), nn_shift=np.empty(n, np.int32))
        _check(self.lib.scgpu_replay_results(self.h, n, out["loop_id"].ctypes.data, out["yaw"].ctypes.data,
                                             out["min_dist"].ctypes.data, out["nn_idx"].ctypes.data, out["nn_shift"].ctypes.data))
        return out

    # ---- peer-sharded database, one process per GPU (include/scgpu.h "peer-sharded database") ----------------------
    def peer_export(self):
        blob = np.zeros(128, np.uint8)
        _check(self.lib.scgpu_peer_export(self.h, blob.ctypes.data, blob.size))
        return blob

    def peer_attach(self, blobs):
        blobs = np.ascontiguousarray(blobs, np.uint8).reshape(-1, 128)
        _check(self.lib.scgpu_peer_attach(self.h, blobs.ctypes.data, blobs.shape[0]))

    def peer_replay_async(self, scans_local, n_total):
        """Collective: this rank's scans of a batch of n_total (see scgpu_peer_replay_async)."""
        keep, (ptr, n, pts, stride, loc) = self._scans(scans_local)
        self._keep = keep
        _check(self.lib.scgpu_peer_replay_async(self.h, ptr, n_total, pts, stride, loc))

    def query_batched(self, first, n):
        out = dict(loop_id=np.empty(n, np.int32), yaw=np.empty(n, np.float32), min_dist=np.empty(n, np.float64),
                   nn_idx=np.empty(n, np.int32), nn_shift=np.empty(n, np.int32))
        _check(self.lib.scgpu_query_batched(self.h, first, n, out["loop_id"].ctypes.data, out["yaw"].ctypes.data,
                                            out["min_dist"].ctypes.data, out["nn_idx"].ctypes.data,
                                            out["nn_shift"].ctypes.data))
        return out

    def candidates(self, q=0):
        K = self.K
        ci, cd = np.zeros(K, np.uint64), np.zeros(K, np.float32)
        sd, ss = np.zeros(K, np.float64), np.zeros(K, np.int32)
        ns = _u64()
        _check(self.lib.scgpu_get_batch_candidates(self.h, q, ci.ctypes.data, cd.ctypes.data, sd.ctypes.data,
                                                   ss.ctypes.data, C.byref(ns)))
        return dict(cand_idx=ci, cand_d2=cd, cand_dist=sd, cand_shift=ss, n_tree=ns.value)

    def get_entry(self, i):
        sc, rk, sk = np.empty(self.R * self.S, np.float32), np.empty(self.R, np.float32), np.empty(self.S, np.float64)
        _check(self.lib.scgpu_get_entry(self.h, i, sc.ctypes.data, rk.ctypes.data, sk.ctypes.data))
        return sc, rk, sk

    def truncate(self, n):
        _check(self.lib.scgpu_truncate(self.h, n))

    def exhaustive(self, q, n_search, flipped=False):
        d, s, i, f = C.c_double(), C.c_int(), C.c_int64(), C.c_int()
        _check(self.lib.scgpu_exhaustive(self.h, q, n_search, int(flipped), C.byref(d), C.byref(s), C.byref(i),
                                         C.byref(f)))
        return d.value, s.value, i.value, f.value

    def exhaustive_batched(self, queries, n_search):
        q = np.ascontiguousarray(queries, np.uint64)
        ns = np.ascontiguousarray(np.broadcast_to(np.asarray(n_search, np.uint64), q.shape))
        d, s, i = np.empty(q.size, np.float64), np.empty(q.size, np.int32), np.empty(q.size, np.int64)
        _check(self.lib.scgpu_exhaustive_batched(self.h, q.ctypes.data, ns.ctypes.data, q.size, d.ctypes.data, s.ctypes.data,
                                                 i.ctypes.data))
        return d, s, i

    def probe_screen(self, q, n_search, shifts=False):
        """Screened distances (float32; -1 = rescore, inf = no valid column pair) of entry q vs entries [0, n_search)."""
        d = np.empty(n_search, np.float32)
        s = np.empty(n_search, np.uint32) if shifts else None
        _check(self.lib.scgpu_probe_screen(self.h, q, n_search, d.ctypes.data, s.ctypes.data if shifts else None))
        return (d, s) if shifts else d

    def exhaustive_rescored(self):
        n = _u64()
        _check(self.lib.scgpu_exhaustive_stats(self.h, C.byref(n)))
        return n.value

    def save(self, path):
        _check(self.lib.scgpu_save(self.h, os.fsencode(path)))

    def load(self, path):
        _check(self.lib.scgpu_load(self.h, os.fsencode(path)))

    def set_downsample_leaf(self, leaf):
        """downSizeFilterScancontext.setLeafSize(leaf, leaf, leaf) (mapOptmization.cpp:264); 0 = off."""
        _check(self.lib.scgpu_set_downsample_leaf(self.h, float(leaf)))

    def voxel_downsample(self, scan, leaf):
        """pcl::VoxelGrid::filter (mapOptmization.cpp:1235-1237) on the device -> dict(points (m,4) f32 = x, y, z, count;
        idx (m,) u32 leaf index; min_b, div_b (3,) i32; refused; passes), sorted by leaf index (PCL's output order)."""
        pts, ptr, n, stride = _pts(scan)
        out = np.zeros((max(n, 1), 4), np.float32)
        idx = np.zeros(max(n, 1), np.uint32)
        mb, db = np.zeros(3, np.int32), np.zeros(3, np.int32)
        m, status = _sz(), C.c_int()
        _check(self.lib.scgpu_voxel_downsample(self.h, ptr, n, stride, float(leaf), out.ctypes.data, idx.ctypes.data, out.shape[0],
                                               C.byref(m), mb.ctypes.data, db.ctypes.data, C.byref(status)))
        order = np.argsort(idx[:m.value], kind="stable")
        return {"points": out[:m.value][order], "idx": idx[:m.value][order], "min_b": mb, "div_b": db,
                "refused": bool(status.value & 1), "passes": status.value >> 8}

    def plan_n_search(self, first_size, n):
        out = np.empty(n, np.uint64)
        _check(self.lib.scgpu_plan_n_search(self.h, first_size, n, out.ctypes.data))
        return out

    def probe_bins(self, xyz):
        """Per point: 0-based bin (sector*R + ring, -1 = not binned), stored height, azimuth [deg] -- device side."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        n = xyz.shape[0]
        b, hh, th = np.empty(n, np.int32), np.empty(n, np.float32), np.empty(n, np.float32)
        _check(self.lib.scgpu_probe_bins(self.h, xyz.ctypes.data, n, b.ctypes.data, hh.ctypes.data, th.ctypes.data))
        return b, hh, th

    def timing(self):
        """(ms_total, ms_build, ms_query) of the last batched call, from CUDA events on the library's stream."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(self.lib.scgpu_get_timing(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def verify_loop(self, source, target, seed_axis=-1, seed_angle=0.0, **overrides):
        """Point-to-point ICP of `source` onto `target` with the reference's settings (mapOptmization.cpp:1053-1078):
        dict(T (4,4), fitness, converged, iterations, accepted).  seed_axis 0/1/2 seeds the rotation with the Scan Context yaw."""
        prm = IcpParams()
        _check(self.lib.scgpu_default_icp_params(C.byref(prm)))
        prm.seed_axis, prm.seed_angle = int(seed_axis), float(seed_angle)
        for k, v in overrides.items():
            setattr(prm, k, v)
        s, sp, sn, ss = _pts(source)
        t, tp, tn, ts = _pts(target)
        if ss != ts:
            raise ValueError("source and target must have the same point stride")
        T = np.empty(16, np.float64)
        fit, conv, its, acc = C.c_double(), C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.scgpu_verify_loop(self.h, sp, sn, tp, tn, ss, C.byref(prm), T.ctypes.data, C.byref(fit), C.byref(conv),
                                          C.byref(its), C.byref(acc)))
        return dict(T=T.reshape(4, 4), fitness=fit.value, converged=bool(conv.value), iterations=its.value, accepted=bool(acc.value))

    @staticmethod
    def _cloud_list(clouds):
        """list of (n, k) float32 arrays with one k -> (kept arrays, pointer array, size array, stride bytes)."""
        arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
        k = arrs[0].shape[1] if arrs else 4
        if any(a.ndim != 2 or a.shape[1] != k for a in arrs):
            raise ValueError("clouds must be (n, k) arrays with the same k")
        ptrs = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
        sizes = (_sz * max(1, len(arrs)))(*[a.shape[0] for a in arrs])
        return arrs, ptrs, sizes, 4 * k

    def assemble_submap(self, clouds, poses, leaf=0.0, drop_negative_intensity=False, intensity_column=3):
        """mapOptmization.cpp:928-949: clouds[i] moved by poses[i] = (x, y, z, roll, pitch, yaw), concatenated; optional
        intensity filter (932-939) and voxel grid (948-949).  Returns (n, 4) float32: x, y, z, intensity (or, with leaf > 0,
        centroid x, y, z and the number of points, in unspecified order)."""
        arrs, ptrs, sizes, stride = self._cloud_list(clouds)
        ps = np.ascontiguousarray(np.asarray(poses, np.float32).reshape(-1, 6))
        if len(ps) != len(arrs):
            raise ValueError("one pose per cloud")
        ioff = 4 * intensity_column if arrs and arrs[0].shape[1] > intensity_column else 0
        cap = max(1, sum(a.shape[0] for a in arrs))
        out = np.empty((cap, 4), np.float32)
        n = _sz()
        _check(self.lib.scgpu_assemble_submap(self.h, ptrs, sizes, ps.ctypes.data, len(arrs), stride, ioff, int(drop_negative_intensity),
                                              float(leaf), out.ctypes.data, cap, C.byref(n)))
        return out[:n.value]

    def verify_loop_keyframes(self, src_clouds, src_pose, tgt_clouds, tgt_poses, leaf=0.3, intensity_column=3, seed_axis=-1,
                              seed_angle=0.0, **overrides):
        """mapOptmization.cpp:924-949 + 1053-1078 in one call: submap assembly on the device, then the ICP (see verify_loop)."""
        prm = IcpParams()
        _check(self.lib.scgpu_default_icp_params(C.byref(prm)))
        prm.seed_axis, prm.seed_angle = int(seed_axis), float(seed_angle)
        for k, v in overrides.items():
            setattr(prm, k, v)
        sa, sp, ss, stride = self._cloud_list(src_clouds)
        ta, tp, ts, stride_t = self._cloud_list(tgt_clouds)
        if stride != stride_t:
            raise ValueError("query and history clouds must have the same point stride")
        spose = np.ascontiguousarray(np.asarray(src_pose, np.float32).reshape(6))
        tposes = np.ascontiguousarray(np.asarray(tgt_poses, np.float32).reshape(-1, 6))
        if len(tposes) != len(ta):
            raise ValueError("one pose per history cloud")
        ioff = 4 * intensity_column if stride > 4 * intensity_column else 0
        T = np.empty(16, np.float64)
        fit, conv, its, acc, ns, nt = C.c_double(), C.c_int(), C.c_int(), C.c_int(), _sz(), _sz()
        _check(self.lib.scgpu_verify_loop_keyframes(self.h, sp, ss, len(sa), spose.ctypes.data, tp, ts, tposes.ctypes.data, len(ta), stride, ioff,
                                                    float(leaf), C.byref(prm), T.ctypes.data, C.byref(fit), C.byref(conv), C.byref(its),
                                                    C.byref(acc), C.byref(ns), C.byref(nt)))
        return dict(T=T.reshape(4, 4), fitness=fit.value, converged=bool(conv.value), iterations=its.value, accepted=bool(acc.value),
                    n_source=ns.value, n_target=nt.value)

    def timer_start(self):
        _check(self.lib.scgpu_timer_start(self.h))

    def timer_stop(self):
        """Device milliseconds since timer_start (waits for everything enqueued)."""
        ms = C.c_double()
        _check(self.lib.scgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def selfcheck_binning(self, n, seed=1, mode=0):
        """(mismatches, fallbacks, first_bad) of the binning front end vs the exact path over n device-generated points."""
        mm, fb = _u64(), _u64()
        bad = np.zeros(5, np.float32)
        _check(self.lib.scgpu_probe_selfcheck(self.h, n, seed, mode, C.byref(mm), C.byref(fb), bad.ctypes.data))
        return mm.value, fb.value, bad

    def growth_stats(self):
        """(grown in place, grown by copy, local capacity): include/scgpu.h scgpu_growth_stats."""
        a, b, c = C.c_uint(), C.c_uint(), _u64()
        _check(self.lib.scgpu_growth_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def record_bytes(self):
        n = _sz()
        _check(self.lib.scgpu_record_bytes(self.h, C.byref(n)))
        return n.value


def host_info():
    """(pool_threads, packs_pinned) of the library's H2D packing path."""
    a, b = C.c_int(), C.c_int()
    _check(load_library().scgpu_host_info(C.byref(a), C.byref(b)))
    return a.value, bool(b.value)


def peer_partition(first, n_total, G, rank):
    """(first_index, count): the scans of a batch shard `rank` of G owns (include/scgpu.h scgpu_peer_partition).  No GPU needed."""
    a, b = _sz(), _sz()
    _check(load_library().scgpu_peer_partition(first, n_total, G, rank, C.byref(a), C.byref(b)))
    return a.value, b.value


def probe_atanf(x):
    """The device's atanf restatement (glibc/fdlibm algorithm in plain binary32) for an array of floats."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    out = np.empty_like(x)
    _check(lib.scgpu_probe_atanf(x.ctypes.data, x.size, out.ctypes.data))
    return out


def xy2theta(x, y):
    out = C.c_float()
    _check(load_library().scgpu_xy2theta(float(x), float(y), C.byref(out)))
    return np.float32(out.value)
