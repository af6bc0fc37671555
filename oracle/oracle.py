"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for (a) the plain-C restatement ``oracle/_build/libscoracle.so`` (class ``Port``) and
(b) the reference's own Scancontext.cpp compiled verbatim, ``oracle/_ref/libscref*.so`` (class ``Ref``).
Only tests/, ``__graft_entry__.smoke()``, bench.py's cpu_baseline / ``--impl reference`` legs and the cpu_baseline legs
of the per-config bench tools (tools/bench_configs.py, bench_exhaustive.py, bench_voxel.py) may import this module; it is
the checker and the CPU baseline, never the thing measured as the product.  The product package (sc-lego-loam_b200/,
include/) never imports, links or executes anything under oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/SC-LeGO-LOAM/LeGO-LOAM"

_vp, _sz, _i, _d, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_float
_pd = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_pf = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_pu64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_pi32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(ref=True, quiet=True):
    """Compile the port (always) and, when /root/reference exists, the verbatim reference variants."""
    targets = ["port"] + (["ref"] if ref and os.path.exists(REF_ROOT) else [])
    subprocess.run(["make", "-C", HERE] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class Params(C.Structure):
    """sco_params (SC.h:77-96)."""
    _fields_ = [("R", _i), ("S", _i), ("lidar_height", _d), ("max_radius", _d), ("exclude_recent", _i),
                ("num_candidates", _i), ("search_ratio", _d), ("dist_thres", _d), ("tree_period", _i)]

    def __init__(self, **kw):
        super().__init__(R=20, S=60, lidar_height=2.0, max_radius=80.0, exclude_recent=50,
                         num_candidates=10, search_ratio=0.1, dist_thres=0.5, tree_period=10)
        for k, v in kw.items():
            setattr(self, k, v)

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _pts_args(pts):
    """pts: (n, k) float32 array, k>=3, C-contiguous -> (pointer, n, stride_bytes)."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] >= 3
    return pts, pts.ctypes.data_as(_vp), pts.shape[0], pts.shape[1] * 4


class Port:
    """The plain-C restatement (oracle/sc_oracle.c)."""

    def __init__(self, params=None):
        path = os.path.join(HERE, "_build", "libscoracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        self.p = params or Params()
        pp = C.POINTER(Params)
        L.sco_atanf.restype = _f; L.sco_atanf.argtypes = [_f]
        L.sco_xy2theta.restype = _f; L.sco_xy2theta.argtypes = [_f, _f]
        L.sco_bin_point.restype = _i
        L.sco_bin_point.argtypes = [pp, _f, _f, _f, C.POINTER(_i), C.POINTER(_i), C.POINTER(_f)]
        L.sco_make_sc.argtypes = [pp, _vp, _sz, _sz, _pd]
        L.sco_ringkey.argtypes = [pp, _pd, _pd]
        L.sco_sectorkey.argtypes = [pp, _pd, _pd]
        L.sco_fast_align.restype = _i; L.sco_fast_align.argtypes = [pp, _pd, _pd]
        L.sco_dist_direct_shifted.restype = _d; L.sco_dist_direct_shifted.argtypes = [pp, _pd, _pd, _i]
        L.sco_distance.argtypes = [pp, _pd, _pd, C.POINTER(_d), C.POINTER(_i)]
        L.sco_key_dist2.restype = _f; L.sco_key_dist2.argtypes = [_pf, _pf, _i]
        L.sco_knn.restype = _i; L.sco_knn.argtypes = [pp, _pf, _sz, _pf, _pu64, _pf]
        L.sco_db_create.restype = _vp; L.sco_db_create.argtypes = [pp]
        L.sco_db_destroy.argtypes = [_vp]
        L.sco_db_size.restype = _sz; L.sco_db_size.argtypes = [_vp]
        L.sco_db_append_scan.argtypes = [_vp, _vp, _sz, _sz]
        L.sco_db_append_desc.argtypes = [_vp, _pd]
        L.sco_db_get_entry.argtypes = [_vp, _sz, _vp, _vp, _vp, _vp]
        L.sco_db_detect.restype = _i
        L.sco_db_detect.argtypes = [_vp, C.POINTER(_i), C.POINTER(_f), C.POINTER(_d), _pu64, _pf, _pd, _pi32,
                                    C.POINTER(C.c_uint64)]
        L.sco_db_exhaustive.argtypes = [_vp, _pd, _sz, _i, C.POINTER(_d), C.POINTER(_i), C.POINTER(C.c_int64),
                                        C.POINTER(_i)]
        L.sco_atanf_many.argtypes = [_pf, _sz, _vp, _vp]
        L.sco_bin_points.argtypes = [pp, _pf, _sz, _pi32, _pf, _pf]
        self.db = L.sco_db_create(C.byref(self.p))

    def __del__(self):
        try:
            self.L.sco_db_destroy(self.db)
        except Exception:
            pass

    # --- stateless functions -------------------------------------------------------------------
    def atanf(self, x):
        return self.L.sco_atanf(float(x))

    def xy2theta(self, x, y):
        return self.L.sco_xy2theta(float(x), float(y))

    def atanf_many(self, x):
        """(restatement, host libm) atanf of every element."""
        x = np.ascontiguousarray(x, np.float32).ravel()
        a, b = np.empty_like(x), np.empty_like(x)
        self.L.sco_atanf_many(x, x.size, a.ctypes.data, b.ctypes.data)
        return a, b

    def bin_points(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        n = xyz.shape[0]
        b, h, t = np.empty(n, np.int32), np.empty(n, np.float32), np.empty(n, np.float32)
        self.L.sco_bin_points(C.byref(self.p), xyz.ravel(), n, b, h, t)
        return b, h, t

    def bin_point(self, x, y, z):
        r, s, h = _i(), _i(), _f()
        ok = self.L.sco_bin_point(C.byref(self.p), x, y, z, C.byref(r), C.byref(s), C.byref(h))
        return (r.value, s.value, h.value) if ok else None

    def make_sc(self, pts):
        pts, ptr, n, stride = _pts_args(pts)
        out = np.empty(self.p.R * self.p.S, np.float64)
        self.L.sco_make_sc(C.byref(self.p), ptr, n, stride, out)
        return out  # column-major R*S

    def ringkey(self, sc):
        out = np.empty(self.p.R, np.float64)
        self.L.sco_ringkey(C.byref(self.p), np.ascontiguousarray(sc, np.float64), out)
        return out

    def sectorkey(self, sc):
        out = np.empty(self.p.S, np.float64)
        self.L.sco_sectorkey(C.byref(self.p), np.ascontiguousarray(sc, np.float64), out)
        return out

    def fast_align(self, v1, v2):
        return self.L.sco_fast_align(C.byref(self.p), np.ascontiguousarray(v1, np.float64),
                                     np.ascontiguousarray(v2, np.float64))

    def dist_direct_shifted(self, a, b, shift):
        return self.L.sco_dist_direct_shifted(C.byref(self.p), np.ascontiguousarray(a, np.float64),
                                              np.ascontiguousarray(b, np.float64), int(shift))

    def distance(self, a, b):
        d, s = _d(), _i()
        self.L.sco_distance(C.byref(self.p), np.ascontiguousarray(a, np.float64),
                            np.ascontiguousarray(b, np.float64), C.byref(d), C.byref(s))
        return d.value, s.value

    def knn(self, keys, query):
        keys = np.ascontiguousarray(keys, np.float32)
        K = self.p.num_candidates
        idx, d2 = np.zeros(K, np.uint64), np.zeros(K, np.float32)
        cnt = self.L.sco_knn(C.byref(self.p), keys, keys.shape[0], np.ascontiguousarray(query, np.float32), idx, d2)
        return cnt, idx, d2

    # --- database ------------------------------------------------------------------------------
    def size(self):
        return self.L.sco_db_size(self.db)

    def append_scan(self, pts):
        pts, ptr, n, stride = _pts_args(pts)
        self.L.sco_db_append_scan(self.db, ptr, n, stride)

    def append_desc(self, sc):
        self.L.sco_db_append_desc(self.db, np.ascontiguousarray(sc, np.float64))

    def get_entry(self, i):
        R, S = self.p.R, self.p.S
        sc, rk, sk, rkf = np.empty(R * S), np.empty(R), np.empty(S), np.empty(R, np.float32)
        self.L.sco_db_get_entry(self.db, i, sc.ctypes.data, rk.ctypes.data, sk.ctypes.data, rkf.ctypes.data)
        return sc, rk, sk, rkf

    def detect(self):
        K = self.p.num_candidates
        lid, yaw, md, nt = _i(), _f(), _d(), C.c_uint64()
        ci, cd = np.zeros(K, np.uint64), np.zeros(K, np.float32)
        sd, ss = np.zeros(K, np.float64), np.zeros(K, np.int32)
        k = self.L.sco_db_detect(self.db, C.byref(lid), C.byref(yaw), C.byref(md), ci, cd, sd, ss, C.byref(nt))
        return dict(loop_id=lid.value, yaw=np.float32(yaw.value), min_dist=md.value, k=k, cand_idx=ci, cand_d2=cd,
                    cand_dist=sd, cand_shift=ss, n_tree=nt.value)

    def exhaustive(self, query_sc, n, flipped=False):
        d, s, i, f = _d(), _i(), C.c_int64(), _i()
        self.L.sco_db_exhaustive(self.db, np.ascontiguousarray(query_sc, np.float64), n, int(flipped),
                                 C.byref(d), C.byref(s), C.byref(i), C.byref(f))
        return d.value, s.value, i.value, f.value


REF_VARIANTS = {
    "default": "libscref.so",          # 20x60, K=10, exclude 50, period 10, ratio 0.1
    "k50": "libscref_k50.so",          # K=50
    "40x120": "libscref_40x120.so",    # R=40, S=120
    "full": "libscref_full.so",        # 7x18, ratio 1.0, K=3, exclude 5, period 3, lidar_height 0, radius 33.3
    "full60": "libscref_full60.so",    # 20x60, ratio 1.0 (every shift searched)
    "intensity": "libscref_intensity.so",  # bin value = point intensity instead of z + LIDAR_HEIGHT (Scancontext.h:41)
}


class Voxel:
    """pcl::VoxelGrid restated (oracle/voxel_oracle.cpp; parity unpinned: PCL is not in this image)."""

    def __init__(self):
        path = os.path.join(HERE, "_build", "libvoxoracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.L = C.CDLL(path)
        self.L.vox_downsample.restype = C.c_long
        self.L.vox_downsample.argtypes = [_vp, _sz, _sz, _f, _vp, _vp, _vp, _sz, _vp, _vp]

    def downsample(self, pts, leaf):
        """-> dict(points (m,4) f32 in ascending leaf-index order, idx (m,) u32, count (m,) u32, min_b, div_b)."""
        pts, ptr, n, stride = _pts_args(pts)
        out = np.zeros((max(n, 1), 4), np.float32)
        idx = np.zeros(max(n, 1), np.uint32)
        cnt = np.zeros(max(n, 1), np.uint32)
        mb, db = np.zeros(3, np.int32), np.zeros(3, np.int32)
        m = self.L.vox_downsample(ptr, n, stride // 4, leaf, out.ctypes.data_as(_vp), idx.ctypes.data_as(_vp),
                                  cnt.ctypes.data_as(_vp), out.shape[0], mb.ctypes.data_as(_vp), db.ctypes.data_as(_vp))
        if m < 0:
            raise ValueError("leaf size too small for the input (PCL refuses)" if m == -1 else "capacity")
        return {"points": out[:m].copy(), "idx": idx[:m].copy(), "count": cnt[:m].copy(), "min_b": mb, "div_b": db}


class Icp:
    """pcl::IterativeClosestPoint restated (oracle/icp_oracle.cpp; parity unpinned: PCL is not in this image)."""

    def __init__(self):
        path = os.path.join(HERE, "_build", "libicporacle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.L = C.CDLL(path)
        self.L.icp_oracle.argtypes = [_vp, _sz, _vp, _sz, _sz, _i, _d, _d, _d, _vp, _vp, C.POINTER(_d), C.POINTER(_i), C.POINTER(_i)]

    def align(self, src, tgt, T0=None, max_iter=100, max_corr=100.0, trans_eps=1e-6, fit_eps=1e-6):
        src = np.ascontiguousarray(src, np.float32)
        tgt = np.ascontiguousarray(tgt, np.float32)
        assert src.shape[1] == tgt.shape[1] >= 3
        T0 = np.eye(4) if T0 is None else np.asarray(T0, np.float64)
        T0 = np.ascontiguousarray(T0.reshape(16))
        T = np.empty(16, np.float64)
        fit, conv, its = _d(), _i(), _i()
        self.L.icp_oracle(src.ctypes.data, src.shape[0], tgt.ctypes.data, tgt.shape[0], src.shape[1], max_iter, max_corr, trans_eps, fit_eps,
                          T0.ctypes.data, T.ctypes.data, C.byref(fit), C.byref(conv), C.byref(its))
        return dict(T=T.reshape(4, 4), fitness=fit.value, converged=bool(conv.value), iterations=its.value)


class Submap:
    """transformPointCloud(cloud, pose) of the submap assembly (mapOptmization.cpp:598-627).  kind="reference": the reference's
    own function text compiled from where it lies (oracle/_ref/libsubmapref.so; oracle/Makefile); kind="port": the restatement
    in oracle/submap_shim.cpp."""

    def __init__(self, kind="port"):
        path = os.path.join(HERE, "_ref", "libsubmapref.so") if kind == "reference" else os.path.join(HERE, "_build", "libsubmaporacle.so")
        if not os.path.exists(path):
            build(ref=kind == "reference")
        self.L = C.CDLL(path)
        self.L.submap_transform.argtypes = [_vp, _sz, _sz, _i, _vp, _vp]
        self.L.submap_keeps.argtypes = [_f]
        self.L.submap_keeps.restype = _i

    @staticmethod
    def available(kind):
        return os.path.exists(os.path.join(HERE, "_ref", "libsubmapref.so")) if kind == "reference" else True

    def transform(self, cloud, pose6, intensity_column=3):
        cloud = np.ascontiguousarray(cloud, np.float32)
        pose = np.ascontiguousarray(pose6, np.float32).reshape(6)
        out = np.empty((cloud.shape[0], 4), np.float32)
        self.L.submap_transform(cloud.ctypes.data, cloud.shape[0], cloud.shape[1], intensity_column if cloud.shape[1] > intensity_column else -1,
                                pose.ctypes.data, out.ctypes.data)
        return out

    def keeps(self, intensity):
        return np.array([bool(self.L.submap_keeps(float(v))) for v in np.asarray(intensity, np.float32)])

    def assemble(self, clouds, poses, drop_negative_intensity=False, intensity_column=3):
        """clouds moved by their poses, concatenated (operator+=), optionally filtered as mapOptmization.cpp:932-939."""
        parts = [self.transform(c, p, intensity_column) for c, p in zip(clouds, poses)]
        out = np.concatenate(parts) if parts else np.zeros((0, 4), np.float32)
        if drop_negative_intensity:
            out = out[self.keeps(out[:, 3])]
        return out


def ref_available(variant="default"):
    return os.path.exists(os.path.join(HERE, "_ref", REF_VARIANTS[variant]))


class Ref:
    """The reference's own SCManager (Scancontext.cpp compiled verbatim) behind oracle/ref_shim.cpp."""

    def __init__(self, variant="default"):
        path = os.path.join(HERE, "_ref", REF_VARIANTS[variant])
        if not os.path.exists(path):
            build(ref=True)
        L = self.L = C.CDLL(path)
        L.scref_create.restype = _vp
        L.scref_destroy.argtypes = [_vp]
        L.scref_params.argtypes = [_vp] + [C.POINTER(_i)] * 5 + [C.POINTER(_d)] * 4
        L.scref_xy2theta.restype = _f; L.scref_xy2theta.argtypes = [_f, _f]
        L.scref_make_sc.argtypes = [_vp, _vp, _sz, _sz, _pd]
        L.scref_ringkey.argtypes = [_vp, _pd, _pd]
        L.scref_sectorkey.argtypes = [_vp, _pd, _pd]
        L.scref_fast_align.restype = _i; L.scref_fast_align.argtypes = [_vp, _pd, _pd]
        L.scref_dist_direct.restype = _d; L.scref_dist_direct.argtypes = [_vp, _pd, _pd]
        L.scref_distance.argtypes = [_vp, _pd, _pd, C.POINTER(_d), C.POINTER(_i)]
        L.scref_append_scan.argtypes = [_vp, _vp, _sz, _sz]
        L.scref_append_desc.argtypes = [_vp, _pd]
        L.scref_size.restype = _sz; L.scref_size.argtypes = [_vp]
        L.scref_get_entry.argtypes = [_vp, _sz, _vp, _vp, _vp, _vp]
        L.scref_detect.argtypes = [_vp, C.POINTER(_i), C.POINTER(_f), C.c_char_p, _sz]
        L.scref_last_candidates.restype = _i
        L.scref_last_candidates.argtypes = [_vp, _pu64, _pf, _pd, _pi32, C.POINTER(C.c_uint64)]
        L.scref_knn.restype = _i; L.scref_knn.argtypes = [_vp, _pf, _sz, _pf, _pu64, _pf]
        L.scref_time_run.argtypes = [_vp, _vp, _sz, _sz, _sz, C.POINTER(_d), C.POINTER(_d), _vp, _vp]
        L.scref_time_exhaustive.restype = _d
        L.scref_time_exhaustive.argtypes = [_vp, _pd, _sz, C.POINTER(_d), C.POINTER(_i), C.POINTER(C.c_int64)]
        self.h = L.scref_create()
        v = [_i() for _ in range(5)] + [_d() for _ in range(4)]
        L.scref_params(self.h, *[C.byref(x) for x in v])
        self.p = Params(R=v[0].value, S=v[1].value, num_candidates=v[2].value, exclude_recent=v[3].value,
                        tree_period=v[4].value, lidar_height=v[5].value, max_radius=v[6].value,
                        search_ratio=v[7].value, dist_thres=v[8].value)

    def __del__(self):
        try:
            self.L.scref_destroy(self.h)
        except Exception:
            pass

    def xy2theta(self, x, y):
        return self.L.scref_xy2theta(float(x), float(y))

    def make_sc(self, pts):
        pts, ptr, n, stride = _pts_args(pts)
        out = np.empty(self.p.R * self.p.S, np.float64)
        self.L.scref_make_sc(self.h, ptr, n, stride, out)
        return out

    def ringkey(self, sc):
        out = np.empty(self.p.R, np.float64)
        self.L.scref_ringkey(self.h, np.ascontiguousarray(sc, np.float64), out)
        return out

    def sectorkey(self, sc):
        out = np.empty(self.p.S, np.float64)
        self.L.scref_sectorkey(self.h, np.ascontiguousarray(sc, np.float64), out)
        return out

    def fast_align(self, v1, v2):
        return self.L.scref_fast_align(self.h, np.ascontiguousarray(v1, np.float64),
                                       np.ascontiguousarray(v2, np.float64))

    def dist_direct(self, a, b):
        return self.L.scref_dist_direct(self.h, np.ascontiguousarray(a, np.float64),
                                        np.ascontiguousarray(b, np.float64))

    def distance(self, a, b):
        d, s = _d(), _i()
        self.L.scref_distance(self.h, np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64),
                              C.byref(d), C.byref(s))
        return d.value, s.value

    def knn(self, keys, query):
        keys = np.ascontiguousarray(keys, np.float32)
        K = self.p.num_candidates
        idx, d2 = np.zeros(K, np.uint64), np.zeros(K, np.float32)
        cnt = self.L.scref_knn(self.h, keys, keys.shape[0], np.ascontiguousarray(query, np.float32), idx, d2)
        return cnt, idx, d2

    def size(self):
        return self.L.scref_size(self.h)

    def append_scan(self, pts):
        pts, ptr, n, stride = _pts_args(pts)
        self.L.scref_append_scan(self.h, ptr, n, stride)

    def append_desc(self, sc):
        self.L.scref_append_desc(self.h, np.ascontiguousarray(sc, np.float64))

    def get_entry(self, i):
        R, S = self.p.R, self.p.S
        sc, rk, sk, rkf = np.empty(R * S), np.empty(R), np.empty(S), np.empty(R, np.float32)
        self.L.scref_get_entry(self.h, i, sc.ctypes.data, rk.ctypes.data, sk.ctypes.data, rkf.ctypes.data)
        return sc, rk, sk, rkf

    def detect(self, details=True):
        lid, yaw = _i(), _f()
        log = C.create_string_buffer(1024)
        self.L.scref_detect(self.h, C.byref(lid), C.byref(yaw), log, 1024)
        out = dict(loop_id=lid.value, yaw=np.float32(yaw.value), log=log.value.decode(), k=0)
        if details:
            K = self.p.num_candidates
            ci, cd = np.zeros(K, np.uint64), np.zeros(K, np.float32)
            sd, ss = np.zeros(K, np.float64), np.zeros(K, np.int32)
            nt = C.c_uint64()
            k = self.L.scref_last_candidates(self.h, ci, cd, sd, ss, C.byref(nt))
            out.update(k=k, cand_idx=ci, cand_d2=cd, cand_dist=sd, cand_shift=ss, n_tree=nt.value)
            if k:
                out["min_dist"] = float(np.min(sd)) if not np.all(np.isnan(sd)) else 10000000.0
        return out

    def time_run(self, scans, want_results=False):
        """scans: (n_scans, pts, k>=3) float32.  Returns (sec_build, sec_detect[, loop_ids, yaws])."""
        scans = np.ascontiguousarray(scans, np.float32)
        n, pts, k = scans.shape
        tb, td = _d(), _d()
        ids = np.zeros(n, np.int32)
        yaws = np.zeros(n, np.float32)
        self.L.scref_time_run(self.h, scans.ctypes.data, n, pts, k * 4, C.byref(tb), C.byref(td),
                              ids.ctypes.data, yaws.ctypes.data)
        return (tb.value, td.value, ids, yaws) if want_results else (tb.value, td.value)

    def time_exhaustive(self, query_sc, n):
        d, s, i = _d(), _i(), C.c_int64()
        sec = self.L.scref_time_exhaustive(self.h, np.ascontiguousarray(query_sc, np.float64), n, C.byref(d),
                                           C.byref(s), C.byref(i))
        return sec, d.value, s.value, i.value
