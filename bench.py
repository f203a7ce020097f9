#!/usr/bin/env python
"""bench.py — ms/registration + hypotheses scored/s on the 200k-point synthetic indoor pair
(BASELINE.json configs[1]: 200k points per cloud, voxel 0.2 m), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the registration path over one batch of `--pairs` independent scan pairs per GPU
(BASELINE config 4 distributes such batches over 1/2/4/8 GPUs: pair b -> rank b mod N, no data-path
collective, weak scaling).  `value` = ms per registration with both raw clouds already resident in HBM
(fccf_register_device), timed with CUDA events on the library's own stream, L2 flushed between steps,
max over ranks.  `e2e` = the same through fccf_register_batch with pinned HOST buffers (H2D of both
clouds and D2H of the result inside the timed region).  The hypothesis-scoring leg times the
fine_verify-equivalent scoring kernel over `--score-hyps` hypotheses per GPU on the leftover clouds of
the same pair; its roofline is the `roofline` object (20 B per moving point per
hypothesis against measured HBM bandwidth, plus FP32/FP64 instruction-issue figures).
`cpu_baseline` / `--impl reference`: the CPU oracle (a port: the reference needs PCL/Eigen/Ceres/FLANN,
which are not available, see oracle/fccf_oracle.cpp) timed on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware work queue per in-flight registration (the driver default of 8 serialises lanes); must be
# set before the CUDA context exists, i.e. before torch touches the device
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "ms/registration + hypotheses scored/s (200k-pt pair)"
KIND, NPTS, LEAF = "indoor", 200_000, 0.2
SEED0 = 100    # BASELINE config 4: seeds 100..163


def workload_name(pairs):
    return "synthetic %s pair, %dk+%dk points, voxel %.1f m, %d pairs/step/GPU (seeds %d+)" % (KIND, NPTS // 1000, NPTS // 1000, LEAF, pairs, SEED0)


def make_pairs(rank, pairs, npts=NPTS):
    from fccf_pcr_b200 import scenes

    out = []
    for i in range(pairs):
        src, tar, Tgt = scenes.make_pair(KIND, npts, SEED0 + rank * pairs + i)
        out.append((src, tar, Tgt))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def perturbed_hypotheses(T, n, seed):
    """n rigid hypotheses around T: rotations up to +-8 degrees about random axes, shifts up to 0.5 m."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = np.radians(rng.uniform(-8, 8, n))
    K = np.zeros((n, 3, 3)); K[:, 0, 1] = -ax[:, 2]; K[:, 0, 2] = ax[:, 1]; K[:, 1, 0] = ax[:, 2]; K[:, 1, 2] = -ax[:, 0]; K[:, 2, 0] = -ax[:, 1]; K[:, 2, 1] = ax[:, 0]
    R = np.eye(3)[None] + np.sin(ang)[:, None, None] * K + (1 - np.cos(ang))[:, None, None] * (K @ K)
    out = np.tile(np.eye(4), (n, 1, 1))
    out[:, :3, :3] = R @ np.asarray(T, float)[:3, :3]
    out[:, :3, 3] = np.asarray(T, float)[:3, 3] + rng.uniform(-0.5, 0.5, (n, 3))
    out[0] = np.asarray(T, float)
    return out.astype(np.float32)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fccf_pcr_b200 as fccf

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    B, K, W = args.pairs, args.steps, args.warmup
    pairs = make_pairs(rank, B, args.points)
    ctx = fccf.Context(local, batch_lanes=args.lanes)   # raises without a CUDA device: no CPU path
    d_pairs = [(torch.from_numpy(s).to(dev), torch.from_numpy(t).to(dev)) for s, t, _ in pairs]
    h_src = [torch.from_numpy(s).pin_memory() for s, _, _ in pairs]
    h_tar = [torch.from_numpy(t).pin_memory() for _, t, _ in pairs]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    # ---- device-resident leg: `value` -------------------------------------------------------
    sp = [ds.data_ptr() for ds, _ in d_pairs]; tp = [dt.data_ptr() for _, dt in d_pairs]
    ns = [ds.shape[0] for ds, _ in d_pairs]; nt = [dt.shape[0] for _, dt in d_pairs]

    def step_device():
        return ctx.register_batch_device(sp, ns, tp, nt, args.leaf)     # B pairs, several in flight

    for _ in range(W):
        step_device()
    l2_flush()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ctx.launch_count
    stage = np.zeros(8)
    dev_ms = 0.0
    t_wall0 = time.perf_counter()
    for k in range(K):
        Ts = step_device()
        tm = ctx.timing
        dev_ms += tm.total_ms            # CUDA events spanning the whole batch (all lanes), recorded by the library
        stage += np.array(list(tm.stage_ms))
        l2_flush()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - l0
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / K
    ms_per_reg = dev_ms / (K * B * world)
    T_check = Ts
    # single-pair latency (one registration at a time, device-resident inputs)
    lat = []
    for k in range(max(5, min(K, 20))):
        ctx.register_device(sp[0], ns[0], tp[0], nt[0], args.leaf)
        lat.append(ctx.timing.total_ms)
        l2_flush()

    # ---- end-to-end leg: host buffers through the batch entry point --------------------------
    srcs = [t.numpy() for t in h_src]; tars = [t.numpy() for t in h_tar]
    for _ in range(max(1, W // 2)):
        ctx.register_batch(srcs, tars, args.leaf)
    l2_flush()
    barrier()
    e2e_s = 0.0
    for k in range(K):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Tb = ctx.register_batch(srcs, tars, args.leaf)       # pinned host -> device, pipeline, result -> host
        e2e_s += time.perf_counter() - t0
        h2d_step, d2h_step = int(ctx.timing.h2d_bytes), int(ctx.timing.d2h_bytes)
        h2d_ms_step = float(ctx.timing.h2d_ms)
        l2_flush()
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    e2e_ms_per_reg = 1e3 * e2e_s / (K * B * world)
    for a, b in zip(Tb, T_check):
        assert np.array_equal(a, b), "host and device entry points disagree"

    # ---- hypothesis-scoring leg -----------------------------------------------------------------
    src0, tar0, Tgt0 = pairs[0]
    T0 = ctx.register(src0, tar0, args.leaf)
    s1 = ctx.blob("sub1").reshape(-1, 3).copy(); s2 = ctx.blob("sub2").reshape(-1, 3).copy()
    H = args.score_hyps
    hyps = perturbed_hypotheses(T0, H, 1234 + rank)
    l1 = ctx.launch_count
    barrier()
    scores, kernel_ms = ctx.score_hypotheses_bench(hyps, s1, s2, max(K, 5))
    score_launches = ctx.launch_count - l1
    kernel_ms = max_over_ranks(kernel_ms)
    hyp_per_s = world * H / (kernel_ms * 1e-3)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    hbm_peak, peak_src = peaks()
    alg_bytes = 20.0 * len(s2) * H                      # 12 B point + 8 B hash slot per moving point per hypothesis
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    sm_hz = 1e6 * (clocks["sm_max_mhz"] if clocks and clocks.get("sm_max_mhz") else 1965.0)
    fp32_peak = 148 * 128 * 2 * sm_hz / 1e12        # TFLOP/s at max clock
    fp64_peak = 148 * 64 * 2 * sm_hz / 1e12
    pts_per_s = len(s2) * H / (kernel_ms * 1e-3)

    out = {
        "metric": METRIC, "value": round(ms_per_reg, 5), "unit": "ms/registration", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(B), "points_per_cloud": args.points, "voxel_m": args.leaf, "pairs_per_step_per_gpu": B,
                   "parallelism": "independent pairs per GPU (pair b -> rank b mod N), no data-path collective",
                   "lanes_in_flight": int(ctx.params.batch_lanes) or 8,
                   "l2": "256 MiB L2 flush between timed steps", "timing": "CUDA events spanning each step's batch (recorded by the library on its streams), summed over steps; max over ranks"},
        "registrations_per_s": round(1e3 / ms_per_reg, 2),
        "latency_ms_single_pair": round(float(np.median(lat)), 4),
        "stage_ms_per_registration": {n: round(float(v) / (K * B), 4) for n, v in zip(
            ["voxelgrid_main", "voxelgrid_pipeline", "planes", "hypotheses", "cluster", "quick_verify", "fine_verify_fuse"], stage[:7])},
        "hypotheses_scored_per_s": round(hyp_per_s, 1),
        "scoring": {"hypotheses_per_gpu": H, "static_points": int(len(s1)), "moving_points": int(len(s2)), "kernel_ms": round(kernel_ms, 5),
                    "moving_points_per_s": round(pts_per_s, 1),
                    "fp32_frac_of_peak": round(18.0 * pts_per_s / 1e12 / fp32_peak, 5), "fp64_frac_of_peak": round(9.0 * pts_per_s / 1e12 / fp64_peak, 5),
                    "note": "18 FP32 + 9 FP64 flop per moving point per hypothesis (SURVEY.md §8d) against 148 SM x 128 (64) lanes x 2 x max SM clock"},
        "wall_s_timed_region": round(t_wall, 3),
        "e2e": {"value": round(e2e_ms_per_reg, 5), "unit": "ms/registration", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                "h2d_ms_per_step": round(h2d_ms_step, 3), "api": "fccf_register_batch (pinned host buffers in, 4x4 out)", "timing": "host wall clock around the call, synchronised on both sides"},
        "gpu_launches": int(launches + score_launches),
        "roofline": {"kernel": "score_warp_kernel (fine_verify-equivalent hypothesis scoring, one hypothesis per warp)", "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak,
                     "unit": "GB/s", "frac": round(achieved / hbm_peak, 5), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "note": "20 B x moving points x hypotheses per launch; the clouds stay L2/shared-memory resident across hypotheses, so issue rate (scoring.fp32/fp64 fractions), not DRAM, is what binds"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(pairs, args, hyps, s1, s2)
    if rank == 0:
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_baseline(pairs, args, hyps, s1, s2, budget_s=12.0):
    """The CPU oracle (port of the reference's algorithm, single thread like the reference) on a bounded
    sample of the same workload: the same pairs, repeated until ~budget_s of CPU work."""
    from oracle.oracle import Oracle

    o = Oracle()
    o.keep_blobs(False)
    t0 = time.perf_counter(); n = 0; pipe = 0.0
    while True:
        for src, tar, _ in pairs:
            o.register(src, tar, args.leaf)
            pipe += o.time_pipeline
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    nrep = 64
    sec = o.bench_fine_verify(hyps[:64], s1, s2, nrep)
    return {"value": round(1e3 * dt / n, 3), "unit": "ms/registration", "cores": 1, "kind": "port",
            "sample": "%d registrations of the bench's own pairs (host points in -> 4x4 out), %.1f s; reference's clock() region alone %.3f ms/registration" % (n, dt, 1e3 * pipe / n),
            "hypotheses_scored_per_s": round(nrep / sec, 1), "host_cpus": os.cpu_count()}


_REF = {}
_REAL_STDOUT = 1


def _ref_init(leaf):
    from oracle.oracle import Oracle

    _REF["o"] = Oracle()
    _REF["o"].keep_blobs(False)
    _REF["leaf"] = leaf


def _ref_one(i):
    src, tar, _ = _REF["pairs"][i]        # generated in the parent before the fork
    t0 = time.perf_counter()
    _REF["o"].register(src, tar, _REF["leaf"])
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on all host cores.  The reference
    (FCCF.cpp) cannot be built here (PCL / Eigen / Ceres / FLANN absent), so this is the oracle port,
    one single-threaded registration per core (the reference is single-threaded), `cores` pairs per step."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    K, W = args.steps, args.warmup
    from fccf_pcr_b200 import scenes

    _REF["pairs"] = [scenes.make_pair(KIND, args.points, SEED0 + i) for i in range(procs)]
    seeds = list(range(procs))
    with mp.get_context("fork").Pool(procs, initializer=_ref_init, initargs=(args.leaf,)) as pool:
        for _ in range(W):
            pool.map(_ref_one, seeds, chunksize=1)
        t0 = time.perf_counter()
        per = []
        for _ in range(K):
            per += pool.map(_ref_one, seeds, chunksize=1)
        dt = time.perf_counter() - t0
    nreg = K * procs
    ms = 1e3 * dt / nreg
    out = {"impl": "reference", "metric": METRIC, "value": round(ms, 4), "unit": "ms/registration", "n_gpus": args.gpus, "steps": K, "warmup": W,
           "ms_per_step": round(1e3 * dt / K, 3), "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(procs), "points_per_cloud": args.points, "voxel_m": args.leaf},
           "cpu_baseline": {"value": round(ms, 4), "unit": "ms/registration", "cores": procs, "kind": "port",
                            "sample": "%d steps x %d registrations (one single-threaded oracle process per host core); mean single registration %.2f ms" % (K, procs, 1e3 * float(np.mean(per)))},
           "e2e": {"value": round(ms, 4), "unit": "ms/registration", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def emit(obj):
    """The ONE JSON line of the contract, written to the process's real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def main():
    # libraries (NCCL's version banner, torch.distributed warnings) write to fd 1: keep the real stdout for
    # the JSON line and send everything else to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=32, help="scan pairs per step per GPU")
    ap.add_argument("--lanes", type=int, default=16, help="registrations kept in flight per GPU")
    ap.add_argument("--points", type=int, default=NPTS)
    ap.add_argument("--leaf", type=float, default=LEAF)
    ap.add_argument("--score-hyps", type=int, default=9472, help="hypotheses scored per GPU (9472 = 2 x 148 SMs x 32 warps: one hypothesis per warp)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
