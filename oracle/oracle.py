"""ctypes binding of the CPU oracle (oracle/libfccf_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_DT = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.int64}


def build(force=False):
    so = os.path.join(_HERE, "libfccf_oracle.so")
    src = os.path.join(_HERE, "fccf_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfccf_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp, ip, lp, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_keep_blobs.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_register.argtypes = [C.c_void_p, fp, C.c_int64, fp, C.c_int64, C.c_float, fp]
        L.orc_time_pipeline.argtypes = [C.c_void_p]
        L.orc_time_pipeline.restype = C.c_double
        L.orc_time_total.argtypes = [C.c_void_p]
        L.orc_time_total.restype = C.c_double
        L.orc_voxelgrid.argtypes = [C.c_void_p, fp, C.c_int64, C.c_float, fp, lp, ip]
        L.orc_voxelgrid.restype = C.c_int64
        L.orc_octree.argtypes = [fp, C.c_int64, C.c_double, ip, ip, ip, dp, ip]
        L.orc_face_extract.argtypes = [C.c_void_p, fp, C.c_int64]
        L.orc_plane_fit.argtypes = [fp, C.c_int, fp]
        L.orc_normal_angle.argtypes = [C.c_float] * 6
        L.orc_normal_angle.restype = C.c_float
        L.orc_quat_from_matrix.argtypes = [fp, fp]
        L.orc_quat_to_matrix.argtypes = [fp, fp]
        L.orc_fine_verify.argtypes = [C.c_void_p, fp, fp, C.c_int64, fp, C.c_int64, ip, C.c_int, ip]
        L.orc_fine_verify.restype = C.c_float
        L.orc_bench_fine_verify.argtypes = [C.c_void_p, fp, C.c_int, fp, C.c_int64, fp, C.c_int64, C.c_int, fp]
        L.orc_bench_fine_verify.restype = C.c_double
        L.orc_quick_verify.argtypes = [C.c_void_p, fp, fp, C.c_int, fp, C.c_int, ip, ip, ip]
        L.orc_quick_verify.restype = C.c_float
        L.orc_hypotheses.argtypes = [C.c_void_p, fp, dp, C.c_int, fp, dp, C.c_int]
        L.orc_cluster.argtypes = [C.c_void_p, fp, ip]
        L.orc_fuse.argtypes = [C.c_void_p, fp, fp, fp, ip, C.c_int, fp]
        L.orc_blob_bytes.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_blob_bytes.restype = C.c_int64
        L.orc_blob_dtype.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_blob_copy.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.orc_blob_names.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.orc_format_output.argtypes = [C.c_float, fp, C.c_char_p, C.c_int]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    def __init__(self, **params):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        for k, v in params.items():
            self.set_param(k, v)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set_param(self, name, value):
        if self.L.orc_set_param(self.h, name.encode(), float(value)) != 0:
            raise KeyError(name)

    def keep_blobs(self, on):
        self.L.orc_keep_blobs(self.h, int(bool(on)))

    def register(self, src, tar, leaf):
        src = np.ascontiguousarray(src, np.float32)
        tar = np.ascontiguousarray(tar, np.float32)
        T = np.zeros(16, np.float32)
        self.L.orc_register(self.h, _f(src), len(src), _f(tar), len(tar), C.c_float(leaf), _f(T))
        return T.reshape(4, 4)

    @property
    def time_pipeline(self):
        return self.L.orc_time_pipeline(self.h)

    @property
    def time_total(self):
        return self.L.orc_time_total(self.h)

    def blob(self, name):
        nb = self.L.orc_blob_bytes(self.h, name.encode())
        if nb < 0:
            raise KeyError(name)
        dt = _DT[self.L.orc_blob_dtype(self.h, name.encode())]
        out = np.zeros(nb // np.dtype(dt).itemsize, dt)
        if nb:
            self.L.orc_blob_copy(self.h, name.encode(), out.ctypes.data_as(C.c_void_p))
        return out

    def blob_names(self):
        buf = C.create_string_buffer(1 << 16)
        n = self.L.orc_blob_names(self.h, buf, len(buf))
        return buf.value.decode().split("\n")[:-1] if n >= 0 else []

    def voxelgrid(self, xyz, leaf):
        xyz = np.ascontiguousarray(xyz, np.float32)
        n = len(xyz)
        out = np.zeros((max(n, 1), 3), np.float32)
        cell = np.zeros(max(n, 1), np.int64)
        cnt = np.zeros(max(n, 1), np.int32)
        m = self.L.orc_voxelgrid(self.h, _f(xyz), n, C.c_float(leaf), _f(out),
                                 cell.ctypes.data_as(C.POINTER(C.c_int64)), _i(cnt))
        return out[:m].copy(), cell[:m].copy(), cnt[:m].copy()

    def octree(self, xyz, res):
        xyz = np.ascontiguousarray(xyz, np.float32)
        n = len(xyz)
        keys = np.zeros((max(n, 1), 3), np.int32)
        start = np.zeros(n + 2, np.int32)
        pidx = np.zeros(max(n, 1), np.int32)
        mn = np.zeros(3, np.float64)
        depth = C.c_int(0)
        V = self.L.orc_octree(_f(xyz), n, C.c_double(res), _i(keys), _i(start), _i(pidx),
                              mn.ctypes.data_as(C.POINTER(C.c_double)), C.byref(depth))
        return keys[:V].copy(), start[:V + 1].copy(), pidx[:n].copy(), mn, depth.value

    def face_extract(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        return self.L.orc_face_extract(self.h, _f(xyz), len(xyz))

    def plane_fit(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        out = np.zeros(8, np.float32)
        self.L.orc_plane_fit(_f(xyz), len(xyz), _f(out))
        return out

    def normal_angle(self, a, b):
        return self.L.orc_normal_angle(*(C.c_float(float(v)) for v in (*a, *b)))

    def quat_from_matrix(self, R):
        R = np.ascontiguousarray(R, np.float32)
        q = np.zeros(4, np.float32)
        self.L.orc_quat_from_matrix(_f(R), _f(q))
        return q

    def quat_to_matrix(self, q):
        q = np.ascontiguousarray(q, np.float32)
        R = np.zeros(9, np.float32)
        self.L.orc_quat_to_matrix(_f(q), _f(R))
        return R.reshape(3, 3)

    def fine_verify(self, T, s1, s2, want_counts=True):
        T = np.ascontiguousarray(T, np.float32).reshape(16)
        s1 = np.ascontiguousarray(s1, np.float32)
        s2 = np.ascontiguousarray(s2, np.float32)
        cap = len(s1) + len(s2) + 1
        counts = np.zeros((cap, 5), np.int32)
        nrows = C.c_int(0)
        sc = self.L.orc_fine_verify(self.h, _f(T), _f(s1), len(s1), _f(s2), len(s2),
                                    _i(counts) if want_counts else None, cap, C.byref(nrows))
        return float(sc), counts[:nrows.value].copy()

    def bench_fine_verify(self, Ts, s1, s2, nrep):
        Ts = np.ascontiguousarray(Ts, np.float32).reshape(-1, 16)
        s1 = np.ascontiguousarray(s1, np.float32)
        s2 = np.ascontiguousarray(s2, np.float32)
        chk = C.c_float(0)
        return self.L.orc_bench_fine_verify(self.h, _f(Ts), len(Ts), _f(s1), len(s1), _f(s2), len(s2), nrep, C.byref(chk))

    def quick_verify(self, T, planes1, planes2):
        T = np.ascontiguousarray(T, np.float32).reshape(16).copy()
        p1 = np.ascontiguousarray(planes1, np.float32)
        p2 = np.ascontiguousarray(planes2, np.float32)
        pairs = np.zeros((64, 2), np.int32)
        npairs = C.c_int(0)
        iters = C.c_int(0)
        s = self.L.orc_quick_verify(self.h, _f(T), _f(p1), len(p1), _f(p2), len(p2), _i(pairs), C.byref(npairs), C.byref(iters))
        return float(s), T.reshape(4, 4), pairs[:npairs.value].copy(), iters.value


def _stage_methods():
    def hypotheses(self, planes1, theta1, planes2, theta2):
        p1 = np.ascontiguousarray(planes1, np.float32).reshape(-1, 7)
        p2 = np.ascontiguousarray(planes2, np.float32).reshape(-1, 7)
        t1 = np.ascontiguousarray(theta1, np.float64)
        t2 = np.ascontiguousarray(theta2, np.float64)
        dp = C.POINTER(C.c_double)
        self.L.orc_hypotheses(self.h, _f(p1), t1.ctypes.data_as(dp), len(p1), _f(p2), t2.ctypes.data_as(dp), len(p2))
        return self.blob("n_hyp")

    def cluster(self, qt7, n_hyp):
        q = np.ascontiguousarray(qt7, np.float32).reshape(-1, 7)
        nh = np.ascontiguousarray(n_hyp, np.int32)
        self.L.orc_cluster(self.h, _f(q), _i(nh))
        return self.blob("n_centres")

    def fuse(self, top_T, s1, s2, n_top):
        T = np.ascontiguousarray(top_T, np.float32)
        k = T.shape[1]
        a = np.ascontiguousarray(s1, np.float32).reshape(3, k)
        b = np.ascontiguousarray(s2, np.float32).reshape(3, k)
        nt = np.ascontiguousarray(n_top, np.int32)
        out = np.zeros(16, np.float32)
        self.L.orc_fuse(self.h, _f(T), _f(a), _f(b), _i(nt), k, _f(out))
        return out.reshape(4, 4)

    Oracle.hypotheses, Oracle.cluster, Oracle.fuse = hypotheses, cluster, fuse


_stage_methods()


def format_output(leaf, T):
    """stdout of the reference's main() for a leaf size and a 4x4 result."""
    T = np.ascontiguousarray(T, np.float32).reshape(16)
    buf = C.create_string_buffer(4096)
    n = lib().orc_format_output(C.c_float(leaf), _f(T), buf, len(buf))
    assert n >= 0
    return buf.value.decode()
