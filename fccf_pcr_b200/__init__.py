"""fccf_pcr_b200 — B200-native FCCF-PCR registration path behind the reference's drop-in surface.

The product is `libfccf.so` (C-ABI in include/fccf.h, hand-written sm_100a CUDA kernels) and the
`FCCF {src} {tar} {voxel}` program; this Python module is only the ctypes binding the tests and
bench.py use, mirroring the reference's operator `computer_transform_guess` (FCCF.cpp:1370) and
its stages.  There is no CPU fallback: without the CUDA library and a GPU every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfccf.so")
CLI_PATH = os.path.join(_HERE, "FCCF")
_DT = {0: np.float32, 1: np.float64, 2: np.int32, 3: np.int64}

PARAM_FIELDS = [
    "parameter_l1", "parameter_l2", "parameter_k1", "parameter_k2", "normal_vector_threshold1",
    "normal_vector_threshold2", "face_voxel_size", "voxel_point_threshold", "curvature_threshold",
    "select_plane_number", "quick_verify_angel_threshold", "quick_verify_distance_threshold",
    "required_optimize_plane", "fine_verify_voxel_size", "fine_verify_number",
    "included_angle_same_threshold", "included_angle_min_threshold", "included_angle_max_threshold",
    "third_plane_threshold", "third_plane_normal_threshold", "cluster_number_threshold",
    "cluster_angel_threshold", "cluster_distance_threshold", "seclct_cluster_number", "rough_threshold_gl",
]


class Params(C.Structure):
    _fields_ = [(n, C.c_float) for n in PARAM_FIELDS] + [("emulate_pcl_overflow", C.c_int), ("batch_lanes", C.c_int), ("reserved", C.c_int * 2)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("downsample_ms", C.c_float), ("pipeline_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("n_launches", C.c_int),
                ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong), ("stage_ms", C.c_float * 8)]


EXPORTS = [
    "fccf_default_params", "fccf_create", "fccf_destroy", "fccf_last_error", "fccf_set_params", "fccf_set_stage_timing", "fccf_register",
    "fccf_register_device", "fccf_register_batch", "fccf_voxelgrid", "fccf_extract_planes", "fccf_score_hypotheses",
    "fccf_score_hypotheses_bench", "fccf_score_counts", "fccf_quick_verify", "fccf_debug_blob", "fccf_launch_count",
    "fccf_stream_handle", "fccf_score_best", "fccf_register_batch_device", "fccf_register_batch_multi", "fccf_score_sharded",
    "fccf_device_count", "fccf_base_pairs", "fccf_hypotheses", "fccf_cluster", "fccf_fuse",
]


def build(force=False):
    """Compile libfccf.so and the FCCF CLI for sm_100a (nvcc cross-compiles without a GPU)."""
    srcdir = os.path.join(_HERE, "csrc")
    args = ["make", "-C", srcdir, "-j8"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libfccf.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.fccf_default_params.argtypes = [C.POINTER(Params)]
        L.fccf_create.argtypes = [C.c_int, C.POINTER(Params)]
        L.fccf_create.restype = vp
        L.fccf_destroy.argtypes = [vp]
        L.fccf_last_error.argtypes = [vp]
        L.fccf_last_error.restype = C.c_char_p
        L.fccf_set_params.argtypes = [vp, C.POINTER(Params)]
        L.fccf_set_stage_timing.argtypes = [vp, C.c_int]
        L.fccf_register.argtypes = [vp, fp, C.c_size_t, fp, C.c_size_t, C.c_float, fp, C.POINTER(Timing)]
        L.fccf_register_device.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_float, fp, C.POINTER(Timing)]
        L.fccf_register_batch.argtypes = [vp, C.c_int, C.POINTER(fp), C.POINTER(C.c_size_t), C.POINTER(fp), C.POINTER(C.c_size_t), C.c_float, fp, C.POINTER(Timing)]
        L.fccf_register_batch_device.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(C.c_size_t), C.c_float, fp, C.POINTER(Timing)]
        L.fccf_voxelgrid.argtypes = [vp, fp, C.c_size_t, C.c_float, fp, C.POINTER(C.c_int64), ip, C.POINTER(C.c_size_t)]
        L.fccf_extract_planes.argtypes = [vp, fp, C.c_size_t, ip]
        L.fccf_score_hypotheses.argtypes = [vp, fp, C.c_size_t, fp, C.c_size_t, fp, C.c_size_t, fp]
        L.fccf_score_hypotheses_bench.argtypes = [vp, fp, C.c_size_t, fp, C.c_size_t, fp, C.c_size_t, C.c_int, fp, fp]
        L.fccf_score_best.argtypes = [vp, C.c_size_t, C.POINTER(C.c_int64), vp]
        L.fccf_score_counts.argtypes = [vp, C.c_size_t, ip, C.c_size_t, C.POINTER(C.c_size_t)]
        L.fccf_quick_verify.argtypes = [vp, fp, C.c_size_t, fp, C.c_int, fp, C.c_int, fp, ip, ip, ip]
        L.fccf_debug_blob.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t), ip]
        L.fccf_launch_count.argtypes = [vp]
        L.fccf_launch_count.restype = C.c_uint64
        L.fccf_stream_handle.argtypes = [vp]
        L.fccf_stream_handle.restype = vp
        L.fccf_register_batch_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(fp), C.POINTER(C.c_size_t), C.POINTER(fp), C.POINTER(C.c_size_t), C.c_float, fp, C.POINTER(Timing)]
        L.fccf_score_sharded.argtypes = [C.POINTER(vp), C.c_int, fp, C.c_size_t, fp, C.c_size_t, fp, C.c_size_t, fp, fp, C.POINTER(C.c_int64), ip]
        L.fccf_device_count.restype = C.c_int
        dp = C.POINTER(C.c_double)
        L.fccf_base_pairs.argtypes = [vp, fp, dp, C.c_int, ip, fp, ip]
        L.fccf_hypotheses.argtypes = [vp, fp, dp, C.c_int, fp, dp, C.c_int, ip]
        L.fccf_cluster.argtypes = [vp, fp, ip, ip]
        L.fccf_fuse.argtypes = [vp, fp, fp, fp, ip, C.c_int, fp]
        _LIB = L
    return _LIB


class FccfError(RuntimeError):
    pass


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def default_params(**over):
    p = Params()
    lib().fccf_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


class HostBatch:
    """Host clouds of a batch with their C argument arrays (see Context.prepare_batch)."""

    def __init__(self, srcs, tars):
        self.n = n = len(srcs)
        self.srcs = [np.ascontiguousarray(a, np.float32) for a in srcs]      # keeps the buffers alive
        self.tars = [np.ascontiguousarray(a, np.float32) for a in tars]
        fp = C.POINTER(C.c_float)
        self.sp = (fp * n)(*[_f(a) for a in self.srcs])
        self.tp = (fp * n)(*[_f(a) for a in self.tars])
        self.ns = (C.c_size_t * n)(*[len(a) for a in self.srcs])
        self.nt = (C.c_size_t * n)(*[len(a) for a in self.tars])
        self.T = np.zeros((n, 16), np.float32)


def device_count():
    return int(lib().fccf_device_count())


def register_batch_multi(ctxs, hb, leaf):
    """fccf_register_batch_multi: a prepared host batch (HostBatch) over several contexts, one host thread per GPU."""
    L = lib()
    n = len(ctxs)
    hs = (C.c_void_p * n)(*[c.h for c in ctxs])
    tms = (Timing * n)()
    rc = L.fccf_register_batch_multi(hs, n, hb.n, hb.sp, hb.ns, hb.tp, hb.nt, C.c_float(leaf), _f(hb.T), tms)
    if rc not in (0, 3):
        raise FccfError("libfccf error %d: %s" % (rc, L.fccf_last_error(ctxs[0].h).decode()))
    return hb.T.reshape(hb.n, 4, 4), list(tms)


def score_sharded(ctxs, Ts, s1, s2):
    """fccf_score_sharded: (scores, best score, best global index, used_nccl)."""
    L = lib()
    n = len(ctxs)
    hs = (C.c_void_p * n)(*[c.h for c in ctxs])
    Ts = np.ascontiguousarray(Ts, np.float32).reshape(-1, 16)
    s1 = np.ascontiguousarray(s1, np.float32)
    s2 = np.ascontiguousarray(s2, np.float32)
    sc = np.zeros(max(len(Ts), 1), np.float32)
    best = C.c_float(0)
    idx = C.c_int64(0)
    used = C.c_int(0)
    rc = L.fccf_score_sharded(hs, n, _f(Ts), len(Ts), _f(s1), len(s1), _f(s2), len(s2), _f(sc), C.byref(best), C.byref(idx), C.byref(used))
    if rc not in (0, 3):
        raise FccfError("libfccf error %d: %s" % (rc, L.fccf_last_error(ctxs[0].h).decode()))
    return sc[:len(Ts)], float(best.value), int(idx.value), bool(used.value)


class Context:
    """One registration context on one GPU (fccf_ctx)."""

    def __init__(self, device=0, **params):
        self.L = lib()
        self.params = default_params(**params)
        h = self.L.fccf_create(int(device), C.byref(self.params))
        if not h:
            raise FccfError("fccf_create failed: no usable CUDA device (libfccf has no CPU fallback)")
        self.h = C.c_void_p(h)
        self.timing = Timing()

    def set_stage_timing(self, on):
        """Per-stage device times in timing.stage_ms[1..6] (six more event nodes per captured sequence)."""
        self._check(self.L.fccf_set_stage_timing(self.h, 1 if on else 0))

    def close(self):
        if getattr(self, "h", None):
            self.L.fccf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, soft=False):
        if rc != 0:
            msg = self.L.fccf_last_error(self.h).decode()
            if soft and rc == 3:
                self.warning = msg
                return
            raise FccfError("libfccf error %d: %s" % (rc, msg))

    def set_params(self, **over):
        for k, v in over.items():
            setattr(self.params, k, v)
        self._check(self.L.fccf_set_params(self.h, C.byref(self.params)))

    # computer_transform_guess + main()'s downsampling: src = argv[1], tar = argv[2]
    def register(self, src, tar, leaf):
        src = np.ascontiguousarray(src, np.float32)
        tar = np.ascontiguousarray(tar, np.float32)
        T = np.zeros(16, np.float32)
        self._check(self.L.fccf_register(self.h, _f(src), len(src), _f(tar), len(tar), C.c_float(leaf), _f(T), C.byref(self.timing)))
        return T.reshape(4, 4)

    def register_device(self, d_src_ptr, n_src, d_tar_ptr, n_tar, leaf):
        T = np.zeros(16, np.float32)
        self._check(self.L.fccf_register_device(self.h, C.c_void_p(d_src_ptr), n_src, C.c_void_p(d_tar_ptr), n_tar, C.c_float(leaf), _f(T), C.byref(self.timing)))
        return T.reshape(4, 4)

    def register_batch(self, srcs, tars, leaf):
        n = len(srcs)
        srcs = [np.ascontiguousarray(a, np.float32) for a in srcs]
        tars = [np.ascontiguousarray(a, np.float32) for a in tars]
        fp = C.POINTER(C.c_float)
        sp = (fp * n)(*[_f(a) for a in srcs])
        tp = (fp * n)(*[_f(a) for a in tars])
        ns = (C.c_size_t * n)(*[len(a) for a in srcs])
        nt = (C.c_size_t * n)(*[len(a) for a in tars])
        T = np.zeros((n, 16), np.float32)
        self._check(self.L.fccf_register_batch(self.h, n, sp, ns, tp, nt, C.c_float(leaf), _f(T), C.byref(self.timing)))
        return T.reshape(n, 4, 4)

    def prepare_batch(self, srcs, tars):
        """Argument block of fccf_register_batch for host clouds, built once: the pointer / size arrays a C
        caller would hold anyway (building them through ctypes costs milliseconds for hundreds of pairs)."""
        return HostBatch(srcs, tars)

    def register_batch_prepared(self, hb, leaf):
        self._check(self.L.fccf_register_batch(self.h, hb.n, hb.sp, hb.ns, hb.tp, hb.nt, C.c_float(leaf), _f(hb.T), C.byref(self.timing)))
        return hb.T.reshape(hb.n, 4, 4)

    def register_batch_device(self, d_src_ptrs, n_srcs, d_tar_ptrs, n_tars, leaf):
        n = len(d_src_ptrs)
        sp = (C.c_void_p * n)(*[int(p) for p in d_src_ptrs])
        tp = (C.c_void_p * n)(*[int(p) for p in d_tar_ptrs])
        ns = (C.c_size_t * n)(*[int(v) for v in n_srcs])
        nt = (C.c_size_t * n)(*[int(v) for v in n_tars])
        T = np.zeros((n, 16), np.float32)
        self._check(self.L.fccf_register_batch_device(self.h, n, sp, ns, tp, nt, C.c_float(leaf), _f(T), C.byref(self.timing)))
        return T.reshape(n, 4, 4)

    def voxelgrid(self, xyz, leaf):
        xyz = np.ascontiguousarray(xyz, np.float32)
        n = len(xyz)
        out = np.zeros((max(n, 1), 3), np.float32)
        cell = np.zeros(max(n, 1), np.int64)
        cnt = np.zeros(max(n, 1), np.int32)
        m = C.c_size_t(0)
        self._check(self.L.fccf_voxelgrid(self.h, _f(xyz), n, C.c_float(leaf), _f(out), cell.ctypes.data_as(C.POINTER(C.c_int64)), _i(cnt), C.byref(m)))
        return out[:m.value].copy(), cell[:m.value].copy(), cnt[:m.value].copy()

    def extract_planes(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        nf = C.c_int(0)
        self._check(self.L.fccf_extract_planes(self.h, _f(xyz), len(xyz), C.byref(nf)))
        return nf.value

    def score_hypotheses(self, Ts, s1, s2):
        Ts = np.ascontiguousarray(Ts, np.float32).reshape(-1, 16)
        s1 = np.ascontiguousarray(s1, np.float32)
        s2 = np.ascontiguousarray(s2, np.float32)
        sc = np.zeros(max(len(Ts), 1), np.float32)
        self._check(self.L.fccf_score_hypotheses(self.h, _f(Ts), len(Ts), _f(s1), len(s1), _f(s2), len(s2), _f(sc)))
        return sc[:len(Ts)]

    def score_hypotheses_bench(self, Ts, s1, s2, repeat):
        Ts = np.ascontiguousarray(Ts, np.float32).reshape(-1, 16)
        s1 = np.ascontiguousarray(s1, np.float32)
        s2 = np.ascontiguousarray(s2, np.float32)
        sc = np.zeros(max(len(Ts), 1), np.float32)
        ms = C.c_float(0)
        self._check(self.L.fccf_score_hypotheses_bench(self.h, _f(Ts), len(Ts), _f(s1), len(s1), _f(s2), len(s2), int(repeat), _f(sc), C.byref(ms)))
        return sc[:len(Ts)], ms.value

    def score_best(self, index_base=0, device_ptr=None):
        """Packed (score, index) maximum of the last scored list (see fccf_score_best / dist.unpack_score_index)."""
        out = C.c_int64(0)
        self._check(self.L.fccf_score_best(self.h, int(index_base), C.byref(out), C.c_void_p(device_ptr) if device_ptr else None))
        return np.int64(out.value)

    def score_counts(self, hyp, cap_rows=1 << 20):
        rows = np.zeros((cap_rows, 5), np.int32)
        n = C.c_size_t(0)
        self._check(self.L.fccf_score_counts(self.h, int(hyp), _i(rows), cap_rows, C.byref(n)))
        r = rows[:min(n.value, cap_rows)]
        order = np.lexsort((r[:, 2], r[:, 1], r[:, 0]))
        return r[order].copy()

    def quick_verify(self, Ts, planes1, planes2):
        Ts = np.ascontiguousarray(Ts, np.float32).reshape(-1, 16).copy()
        p1 = np.ascontiguousarray(planes1, np.float32).reshape(-1, 7)
        p2 = np.ascontiguousarray(planes2, np.float32).reshape(-1, 7)
        n = len(Ts)
        sc = np.zeros(n, np.float32)
        npair = np.zeros(n, np.int32)
        pairs = np.zeros((n, 16, 2), np.int32)
        iters = np.zeros(n, np.int32)
        self._check(self.L.fccf_quick_verify(self.h, _f(Ts), n, _f(p1), len(p1), _f(p2), len(p2), _f(sc), _i(npair), _i(pairs), _i(iters)))
        return sc, Ts.reshape(-1, 4, 4), npair, pairs, iters

    def base_pairs(self, planes, theta):
        """select_base on one plane table: (pairs [B,3] = i, j, type; angles [B])."""
        p = np.ascontiguousarray(planes, np.float32).reshape(-1, 7)
        th = np.ascontiguousarray(theta, np.float64)
        pairs = np.zeros((120, 3), np.int32)
        ang = np.zeros(120, np.float32)
        n = C.c_int(0)
        self._check(self.L.fccf_base_pairs(self.h, _f(p), th.ctypes.data_as(C.POINTER(C.c_double)), len(p), _i(pairs), _f(ang), C.byref(n)))
        return pairs[:n.value].copy(), ang[:n.value].copy()

    def hypotheses(self, planes1, theta1, planes2, theta2):
        """select_base x 2 + match loop + computer_transform: n_hyp[3]; pools via blob('hyp0') ..."""
        p1 = np.ascontiguousarray(planes1, np.float32).reshape(-1, 7)
        p2 = np.ascontiguousarray(planes2, np.float32).reshape(-1, 7)
        t1 = np.ascontiguousarray(theta1, np.float64)
        t2 = np.ascontiguousarray(theta2, np.float64)
        nh = np.zeros(3, np.int32)
        dp = C.POINTER(C.c_double)
        self._check(self.L.fccf_hypotheses(self.h, _f(p1), t1.ctypes.data_as(dp), len(p1), _f(p2), t2.ctypes.data_as(dp), len(p2), _i(nh)), soft=True)
        return nh

    def cluster(self, qt7, n_hyp):
        """cluster_num + transform_cluster of three concatenated pools of (qw qx qy qz tx ty tz) rows: n_centres[3]."""
        q = np.ascontiguousarray(qt7, np.float32).reshape(-1, 7)
        nh = np.ascontiguousarray(n_hyp, np.int32)
        nc = np.zeros(3, np.int32)
        self._check(self.L.fccf_cluster(self.h, _f(q), _i(nh), _i(nc)), soft=True)
        return nc

    def fuse(self, top_T, s1, s2, n_top):
        """per-type best + 0.8 gate + fuse_answer: top_T [3,k,4,4], s1 / s2 [3,k], n_top[3] -> 4x4."""
        T = np.ascontiguousarray(top_T, np.float32)
        k = T.shape[1]
        a = np.ascontiguousarray(s1, np.float32).reshape(3, k)
        b = np.ascontiguousarray(s2, np.float32).reshape(3, k)
        nt = np.ascontiguousarray(n_top, np.int32)
        out = np.zeros(16, np.float32)
        self._check(self.L.fccf_fuse(self.h, _f(T), _f(a), _f(b), _i(nt), k, _f(out)))
        return out.reshape(4, 4)

    def blob(self, name):
        nb = C.c_size_t(0)
        dt = C.c_int(0)
        self._check(self.L.fccf_debug_blob(self.h, name.encode(), None, 0, C.byref(nb), C.byref(dt)))
        out = np.zeros(nb.value // np.dtype(_DT[dt.value]).itemsize, _DT[dt.value])
        if nb.value:
            self._check(self.L.fccf_debug_blob(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), nb.value, C.byref(nb), C.byref(dt)))
        return out

    @property
    def stream_handle(self):
        """cudaStream_t of this context as an integer (e.g. for torch.cuda.ExternalStream)."""
        return int(self.L.fccf_stream_handle(self.h) or 0)

    @property
    def launch_count(self):
        return int(self.L.fccf_launch_count(self.h))
