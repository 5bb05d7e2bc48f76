#!/usr/bin/env python
"""
bench.py — the headline benchmark of the receiver-grid path-tracing hot path.

Workload (BASELINE.json configs[2] / north_star): the example.geojson city scene (28 walls, 2 of
zero length; TX = bbox NW corner), ImagePath orders 0-2 (785 candidates), smooth logic
(hard_sigmoid, alpha = 100), forward power map + VJP (cotangents of the receiver coordinates, object
vertices, TX position and alpha), on a receiver grid of 1024 x 1024 points PER GPU (weak scaling:
N GPUs trace a (1024 N) x 1024 grid, row-sharded, one NCCL all-reduce of the scene-parameter
cotangents per step).  A "step" = one forward launch + one backward launch, as jax.vjp runs them: the forward
also writes the activity mask (1 bit per warp x candidate, 3 MB), the backward re-traces the paths it marks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--coords raw|normalised]

Metric: Rx x candidate paths per second (R * T * C / step time), whole job.
`--impl reference` times the restated reference (oracle/ref_torch.py: torch-CPU fp32 + autograd,
all host threads — JAX is not installable in this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rx_x_candidate_paths_per_s"
UNIT = "paths/s"
GRID_PER_GPU = (1024, 1024)
MAX_ORDER = 2
ALPHA = 100.0
MODE = "hard_sigmoid"
FLUSH_BYTES = 160 << 20  # > the 126 MB L2


def load_scene(coords: str):
    import differt2d_b200 as d
    from tests import helpers as H

    sc = d.Scene.from_geojson(H.geojson_text())
    if coords == "normalised":
        sc = H.normalised(sc)
    return sc


def flops_per_path(k: int, n: int, smooth: bool) -> float:
    """F_fwd(k,N) = F_seg (k+1) N + F_int k + 8 (k+1) + 5."""
    f_seg, f_int = (37.0, 62.0) if smooth else (17.0, 52.0)
    return f_seg * (k + 1) * n + f_int * k + 8.0 * (k + 1) + 5.0


def flops_per_receiver(n: int, max_order: int, smooth: bool) -> float:
    total = 0.0
    for k in range(max_order + 1):
        c = 1 if k == 0 else n * (n - 1) ** (k - 1)
        total += c * flops_per_path(k, n, smooth)
    return total


class ClockSampler:
    """
    SM clock / throttle-reason samples taken DURING the timed region.  The region lasts tens of milliseconds, far
    below what `nvidia-smi -lms` can resolve, so NVML is polled in-process (nvidia_ml_py) from a thread, about
    every millisecond, between mark_start() and mark_stop(); `nvidia-smi` is the fallback when NVML cannot load.
    """

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index: int, uuid: str | None = None):
        self.index, self.uuid = index, uuid
        self.sm, self.bits, self.power = [], 0, []
        self.sm_max = None
        self.live = False
        self.done = False
        self.thread = None
        self.source = None
        self.h = None

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = N.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
                except Exception:
                    h = None
            self.h = h if h is not None else N.nvmlDeviceGetHandleByIndex(self.index)
            self.N = N
            self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.h = None
            self.source = "nvidia-smi"

    def _poll(self):
        N = self.N
        while not self.done:
            if self.live:
                try:
                    self.sm.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
                    self.bits |= int(N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.power.append(N.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                except Exception:
                    pass
            time.sleep(0.002)

    def mark_start(self):
        self.live = True

    def mark_stop(self):
        self.live = False

    def _smi_once(self) -> dict:
        try:
            out = subprocess.run(
                ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                    "source": "nvidia-smi single query right after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}

    def stop(self) -> dict:
        self.done = True
        if self.h is None:
            return self._smi_once()
        if self.thread is not None:
            self.thread.join(timeout=1)
        reasons = sorted(nm for nm, bit in self.REASONS if self.bits & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None,
                "source": "NVML polled in-process during the timed region"}


def bench_reference(args, rank: int, world: int) -> None:
    """The restated reference on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import torch

    from oracle import ref_torch as R
    from tests import helpers as H

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sc = load_scene(args.coords)
    osc = H.oracle_scene_from_product(sc)
    n_s, m_s, stride = 64, 64, 16
    X, Y = sc.grid(GRID_PER_GPU[1], GRID_PER_GPU[0])
    X, Y = X[:: GRID_PER_GPU[0] // n_s, :: GRID_PER_GPU[1] // m_s], Y[:: GRID_PER_GPU[0] // n_s, :: GRID_PER_GPU[1] // m_s]
    cands = R.all_path_candidates(osc.n, 0, MAX_ORDER)
    n_c = len(cands[::stride])

    def step():
        with R.clean_gradients():
            R.power_map_and_vjp(osc, X, Y, None, max_order=MAX_ORDER, approx=True, alpha=ALPHA, function=MODE,
                                cand_stride=stride)

    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    paths = X.size * n_c
    value = paths / dt
    sample = (f"{n_s}x{m_s} receivers (every {GRID_PER_GPU[0] // n_s}th row/col of the 1024x1024 grid) x every "
              f"{stride}th of {len(cands)} candidates, forward + autograd VJP, torch-CPU fp32 eager")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "restated reference (JAX unavailable in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def workload_config(args, world):
    return {
        "workload": f"example.geojson city scene ({args.coords} coordinates), ImagePath orders 0-{MAX_ORDER} "
                    f"(785 candidates), {MODE} alpha={ALPHA:g}, forward + VJP, "
                    f"{GRID_PER_GPU[0]}x{GRID_PER_GPU[1]} receivers per GPU",
        "grid_global": [GRID_PER_GPU[0] * world, GRID_PER_GPU[1]],
        "sharding": f"receiver-grid rows over {world} GPU(s), bands of 8 rows dealt round robin; NCCL all-reduce of scene-parameter cotangents",
        "l2": f"L2 flushed between timed steps ({FLUSH_BYTES >> 20} MiB memset, inside the timed region)",
    }


def cpu_baseline_sample(sc, args) -> dict:
    """Bounded CPU sample of the same workload, on rank 0 at N=1 only (about 10-30 s)."""
    import torch

    from oracle import c_oracle as CO
    from oracle import ref_torch as R
    from tests import helpers as H

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    osc = H.oracle_scene_from_product(sc)
    n_s, m_s, stride = 64, 64, 16
    X, Y = sc.grid(GRID_PER_GPU[1], GRID_PER_GPU[0])
    Xs, Ys = X[:: GRID_PER_GPU[0] // n_s, :: GRID_PER_GPU[1] // m_s], Y[:: GRID_PER_GPU[0] // n_s, :: GRID_PER_GPU[1] // m_s]
    cands = R.all_path_candidates(osc.n, 0, MAX_ORDER)
    n_c = len(cands[::stride])
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 and time.perf_counter() - t0 < 20.0:
        with R.clean_gradients():
            R.power_map_and_vjp(osc, Xs, Ys, None, max_order=MAX_ORDER, approx=True, alpha=ALPHA, function=MODE,
                                cand_stride=stride)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    out = {"value": Xs.size * n_c / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{n_s}x{m_s} receivers x every {stride}th of {len(cands)} candidates, forward + autograd VJP, "
                     f"torch-CPU fp32 (restated reference; JAX unavailable), {reps} rep(s)"}
    # context: the scalar C port of the FORWARD only, OpenMP over receivers
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    Xc, Yc = X[::8, ::8], Y[::8, ::8]
    grid = np.stack([Xc, Yc], -1).reshape(-1, 2).astype(np.float32)
    t0 = time.perf_counter()
    CO.power_map(xys, fixed, grid, max_order=MAX_ORDER, mode=MODE, alpha=ALPHA)
    dtc = time.perf_counter() - t0
    out["forward_only_c_port"] = {"value": grid.shape[0] * len(cands) / dtc, "unit": UNIT, "cores": CO.num_threads(),
                                  "sample": "128x128 receivers x 785 candidates, forward only, scalar C + OpenMP"}
    return out


def _emit(line: dict) -> None:
    """The ONE JSON line on the real stdout (fd 1 is pointed at stderr while the benchmark runs: NCCL and other
    native libraries print banners on stdout)."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


_REAL_STDOUT = None


def _quiet_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--coords", default="raw", choices=["raw", "normalised"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    _quiet_stdout()
    if args.impl == "reference":
        bench_reference(args, rank, world)
        return

    import torch

    import differt2d_b200 as d  # noqa: F401
    from differt2d_b200 import _lib as L
    from differt2d_b200 import distributed as D
    from differt2d_b200 import functional as F

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = D.init(world, rank) if world > 1 else None

    sc = load_scene(args.coords)
    xys, kinds, phis = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    n_rows, n_cols = GRID_PER_GPU[0] * world, GRID_PER_GPU[1]
    X, Y = sc.grid(n_cols, n_rows)
    rows = D.row_tiles_cyclic(n_rows, world, rank)  # bands of 8 rows, round robin: equal work on every rank
    grid_h = np.stack([X[rows], Y[rows]], -1).reshape(-1, 2).astype(np.float32)
    R = grid_h.shape[0]
    cfg = F.TraceConfig(mode=MODE, max_order=MAX_ORDER, reduce_all=True, grid_cols=n_cols)
    n_obj = xys.shape[0]
    n_cand = sum(1 if k == 0 else n_obj * (n_obj - 1) ** (k - 1) for k in range(MAX_ORDER + 1))
    T = fixed.shape[0]

    rng = np.random.default_rng(1234 + rank)
    zbar_h = rng.standard_normal(R).astype(np.float32)
    grid = torch.from_numpy(grid_h).to(dev)
    zbar = torch.from_numpy(zbar_h).to(dev)
    xys_d = torch.from_numpy(xys).to(dev)
    fixed_d = torch.from_numpy(fixed).to(dev)

    # device-resident step through the C ABI (device pointers, caller's stream)
    pk = F._Packed(cfg, xys_d, None, None, fixed_d, grid, ALPHA, None, dev)
    # the VJP's only residual: one activity bit per (warp of 32 receivers, candidate), written by the forward
    # launch of the step and read by its backward launch (the custom_vjp residual)
    mask = pk.new_mask()
    Z = torch.empty(R, device=dev)
    gbar = torch.empty(R, 2, device=dev)
    pbar = torch.zeros(n_obj * 4 + n_obj + 2 * T + 1, device=dev)  # objects | phis | fixed | alpha, one NCCL buffer
    o_obj, o_phi, o_fix, o_alpha = 0, n_obj * 4, n_obj * 5, n_obj * 5 + 2 * T
    base = pbar.data_ptr()
    lib = L.lib()
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    def step(i=None):
        flush.zero_()
        if i is not None:
            ev[i][0].record(stream)
        L.check(lib.d2d_power_fwd(C.byref(pk.p), Z.data_ptr(), None, stream.cuda_stream), "d2d_power_fwd")
        if i is not None:
            ev[i][1].record(stream)
        L.check(lib.d2d_power_bwd(C.byref(pk.p), zbar.data_ptr(), None, gbar.data_ptr(), base + 4 * o_obj,
                                  base + 4 * o_phi, base + 4 * o_fix, base + 4 * o_alpha, stream.cuda_stream),
                "d2d_power_bwd")
        if i is not None:
            ev[i][2].record(stream)
        if dist is not None:
            D.allreduce_sum_(pbar)
            if i is not None:
                ev[i][3].record(stream)

    # FP32 peak (roofline denominator): register-resident FMA chains on every SM, timed alone
    sink = torch.zeros(4, device=dev)
    flops_c = C.c_double(0.0)
    peak_tf = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        L.check(lib.d2d_fma_peak_launch(sink.data_ptr(), 20000, C.byref(flops_c), stream.cuda_stream), "fma_peak")
        e1.record(stream)
        torch.cuda.synchronize(dev)
        peak_tf = max(peak_tf, flops_c.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    if dist is not None:
        D.barrier()
    try:
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
        sampler.mark_start()
    launches0 = F.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    t_start.record(stream)
    for i in range(args.steps):
        step(i)
    t_end.record(stream)
    torch.cuda.synchronize(dev)
    if dist is not None:
        D.barrier()
    sampler.mark_stop()
    launches = F.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t_start.elapsed_time(t_end)
    if dist is not None:
        elapsed_ms = D.max_over_ranks(elapsed_ms, dev)
    ms_per_step = elapsed_ms / args.steps
    fwd_ms = float(np.mean([ev[i][0].elapsed_time(ev[i][1]) for i in range(args.steps)]))
    bwd_ms = float(np.mean([ev[i][1].elapsed_time(ev[i][2]) for i in range(args.steps)]))
    per_rank = None
    if dist is not None:  # where a multi-GPU step goes, per rank: kernels, all-reduce (incl. waiting for the slowest rank)
        ar_ms = float(np.mean([ev[i][2].elapsed_time(ev[i][3]) for i in range(args.steps)]))
        mine = torch.tensor([fwd_ms, bwd_ms, ar_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"fwd_ms": float(t[0]), "bwd_ms": float(t[1]), "allreduce_and_wait_ms": float(t[2])} for t in allr]

    # end to end through the host-buffer C ABI entry (pinned host memory in, host memory out)
    grid_p = torch.from_numpy(grid_h).pin_memory()
    zbar_p = torch.from_numpy(zbar_h).pin_memory()
    Z_p = torch.empty(R).pin_memory()
    gbar_p = torch.empty(R, 2).pin_memory()
    obar_p = torch.empty(n_obj, 4).pin_memory()
    fbar_p = torch.empty(T, 2).pin_memory()
    abar_p = torch.empty(1).pin_memory()
    hp = L.new_problem()
    hp.n_objects, hp.objects_xys = n_obj, xys.ctypes.data
    hp.n_fixed, hp.fixed_xy = T, fixed.ctypes.data
    hp.n_grid, hp.grid_xy = R, grid_p.data_ptr()
    hp.max_order, hp.mode, hp.alpha, hp.reduce_all = MAX_ORDER, L.MODE_HARD_SIGMOID, ALPHA, 1
    hp.grid_cols = n_cols

    def e2e_step():
        L.check(lib.d2d_power_host(C.byref(hp), zbar_p.data_ptr(), Z_p.data_ptr(), gbar_p.data_ptr(),
                                   obar_p.data_ptr(), None, fbar_p.data_ptr(), abar_p.data_ptr(), local),
                "d2d_power_host")
        if dist is not None:
            D.allreduce_sum_(pbar)

    for _ in range(max(1, args.warmup)):
        e2e_step()
    torch.cuda.synchronize(dev)
    if dist is not None:
        D.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        D.barrier()
        e2e_s = D.max_over_ranks(e2e_s, dev)
    e2e_ms = e2e_s * 1e3 / args.steps
    h2d = int(grid_h.nbytes + zbar_h.nbytes + xys.nbytes + fixed.nbytes)
    d2h = int(R * 4 + R * 8 + n_obj * 16 + T * 8 + 4)

    if rank != 0:
        if dist is not None:
            D.shutdown()
        return

    paths_step = float(R) * world * T * n_cand
    f_rx = flops_per_receiver(n_obj, MAX_ORDER, smooth=True)
    kernels = {
        "power_fwd_kernel": {"ms": fwd_ms, "algorithmic_flop": f_rx * R * T},
        "power_bwd_kernel": {"ms": bwd_ms, "algorithmic_flop": 2.0 * f_rx * R * T},
    }
    for k in kernels.values():
        k["achieved_tflops"] = k["algorithmic_flop"] / (k["ms"] * 1e-3) / 1e12
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    # the executed picture of the same command under `ncu --set full` (profiles/, committed; cold-cache capture)
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics.json")))["kernels"]
    except Exception:
        pass
    traffic = None
    executed = None
    if dom in ncu:
        traffic = ncu[dom]["dram_bytes_read"] + ncu[dom]["dram_bytes_write"]
        executed = {k: ncu[dom][k] for k in ("issue_slots_busy_pct", "fp32_lanes_busy_pct", "executed_fp32_flop",
                                             "warp_instructions", "achieved_occupancy_pct")}
        ex_tf = ncu[dom]["executed_fp32_flop"] / (ncu[dom]["duration_ms_under_ncu"] * 1e-3) / 1e12
        executed["fp32_tflops"] = ex_tf
        executed["fp32_frac_of_peak"] = ex_tf / peak_tf if peak_tf else None
        executed["source"] = "profiles/ncu_metrics.json (ncu --set full of this command, per launch)"
    algo_bytes = R * (8 + 4 + 4 + 8)  # grid in, Zbar in, Z out... per launch of the dominant kernel
    line = {
        "metric": METRIC, "value": paths_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "power_map_plus_vjp_ms": ms_per_step,
        "e2e": {"value": paths_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "d2d_power_host (fused value+VJP, pinned host buffers in/out, sync)"},
        "gpu_launches": int(launches),
        "per_rank": per_rank,
        "clocks": clocks,
        "roofline": {
            "bound": "fp32", "kernel": dom, "achieved": kernels[dom]["achieved_tflops"], "peak": peak_tf,
            "unit": "TFLOP/s", "frac": kernels[dom]["achieved_tflops"] / peak_tf if peak_tf else None,
            "traffic": traffic,
            "executed": executed,
            "peak_source": "FP32 FMA-chain microbenchmark (d2d_fma_peak_launch) timed in this run; "
                           "MEASURED_PEAKS.json holds no FP32 CUDA-core figure",
            "convention": "ALGORITHMIC flop of SURVEY §8(d) (pruned work still counted; VJP = 2 x forward): exact "
                          "pruning (tile/warp culls, early outs, activity mask) removes most of it, hence frac > 1; "
                          "`executed` is what the SMs actually issued",
            "kernels": kernels,
            "hbm": {"algorithmic_bytes_per_launch": algo_bytes,
                    "achieved_gbs": algo_bytes / (kernels[dom]["ms"] * 1e-3) / 1e9,
                    "peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0},
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(sc, args)
    _emit(line)
    if dist is not None:
        D.shutdown()


if __name__ == "__main__":
    main()
