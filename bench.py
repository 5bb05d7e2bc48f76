#!/usr/bin/env python
"""
bench.py — the headline benchmark of the receiver-grid path-tracing hot path.

Headline workload (BASELINE.json configs[2] / north_star): the example.geojson city scene (28 walls, 2 of zero length;
TX = bbox NW corner), ImagePath orders 0-2 (785 candidates), smooth logic (hard_sigmoid, alpha = 100), forward power
map + VJP (cotangents of the receiver coordinates, object vertices, TX position and alpha).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]
                    [--workload city|p2p500] [--coords raw|normalised]

* `--scaling weak` (default, the headline line): 1024 x 1024 receivers PER GPU (N GPUs trace a (1024 N) x 1024 grid).
  `--scaling strong`: BASELINE config 3, a 2048 x 2048 grid in total, row-sharded over the N GPUs.  The default run
  also times the strong-scaling configuration for a few steps and reports it under the extra key "strong".
* A "step" = L2 flush + forward launch (writes the activity mask, the VJP's only residual) + backward launch + (N > 1)
  ONE NCCL all-reduce of the 143 scene-parameter cotangents, captured ONCE in a CUDA graph and replayed K times: one
  host call per step, so that host-side jitter of a rank cannot delay the collective of all the others.
* "roofline": the headline scene is pruned almost entirely by exact culls, so its algorithmic-flop figure (SURVEY §8d,
  kept under roofline.algorithmic) says nothing about the hardware.  roofline.frac is measured on a second timed leg
  where NOTHING can be pruned (BASELINE config 2: obstacle scene, sigmoid, alpha = 1, 1024 x 1024, forward + VJP: every
  one of the 65 x 2^20 paths has a non-zero validity and is traced and reversed): algorithmic flop / time / FP32 peak.
* "parity_spotcheck": after the timed loop, a strided subset of the receivers of the benchmark's OWN outputs (Z of the
  last launch, grid_bar) is checked against the CPU oracle.

`--impl reference` times the restated reference on the host cores (oracle/d2d_oracle_ad.cpp: compiled scalar forward +
dual-number VJP, OpenMP — JAX is not installable in this image, probed at every run) on a bounded sample of the same
workload: ALL 785 candidates on a strided 128 x 128 subset of the receivers (stated in config.workload).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rx_x_candidate_paths_per_s"
UNIT = "paths/s"
GRID_PER_GPU = (1024, 1024)      # weak scaling: rows x cols per GPU
GRID_STRONG = (2048, 2048)       # strong scaling: BASELINE config 3, whole job
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


def n_candidates(n: int, lo: int, hi: int) -> int:
    return sum(1 if k == 0 else n * (n - 1) ** (k - 1) for k in range(lo, hi + 1))


# ---- clocks: sampled DURING the timed region by a separate process ------------------------------------------------------
_SAMPLER_SRC = r"""
import sys, time
idx, uuid = int(sys.argv[1]), sys.argv[2]
try:
    import pynvml as N
    N.nvmlInit()
    h = None
    if uuid and uuid != "-":
        try:
            h = N.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            h = None
    if h is None:
        h = N.nvmlDeviceGetHandleByIndex(idx)
    print("max", float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)), flush=True)
    while True:
        t = time.monotonic()
        try:
            print("s", t, float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)),
                  int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h)), N.nvmlDeviceGetPowerUsage(h) / 1e3, flush=True)
        except Exception:
            pass
        time.sleep(0.001)
except Exception as e:
    print("err", repr(e), flush=True)
"""


class ClockSampler:
    """
    SM clock / throttle-reason samples taken DURING the timed region.  The region lasts tens of milliseconds, far below
    what `nvidia-smi -lms` resolves, so NVML is polled about every millisecond — by a SEPARATE PROCESS (round 1 polled
    from a thread of rank 0's interpreter; under the GIL that made rank 0 arrive ~0.25 ms late at every all-reduce).
    Samples carry time.monotonic() stamps (system-wide on Linux); the parent keeps those inside [start, stop].
    """

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index: int, uuid: str | None = None):
        self.index, self.uuid = index, uuid
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self.index), self.uuid or "-"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.5)  # NVML initialisation, before the timed region
        except Exception:
            self.proc = None

    def mark_start(self):
        self.t0 = time.monotonic()

    def mark_stop(self):
        self.t1 = time.monotonic()

    def _smi_once(self, why: str) -> dict:
        try:
            out = subprocess.run(
                ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                    "source": f"nvidia-smi single query right after the timed region ({why})"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}

    def stop(self) -> dict:
        if self.proc is None:
            return self._smi_once("sampler process could not start")
        time.sleep(0.005)
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            text = ""
        sm, bits, power, sm_max = [], 0, [], None
        for ln in text.splitlines():
            f = ln.split()
            try:  # (the last line may be cut short: the sampler is terminated while it writes)
                if len(f) == 2 and f[0] == "max":
                    sm_max = float(f[1])
                elif len(f) == 5 and f[0] == "s" and self.t0 is not None and self.t0 <= float(f[1]) <= self.t1:
                    c, b, w = float(f[2]), int(f[3]), float(f[4])
                    sm.append(c)
                    bits |= b
                    power.append(w)
            except ValueError:
                continue
        if not sm:
            return self._smi_once("no NVML sample fell inside the timed region")
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": sm_max,
                "reasons": sorted(nm for nm, bit in self.REASONS if bits & bit), "samples": len(sm),
                "power_w_max": max(power),
                "source": "NVML polled every ~1 ms by a separate process during the timed region"}


# ---- CPU reference sample (both arms print the same description; `--impl reference` times K steps of it) ------------------
def jax_probe() -> str:
    try:
        import differt2d  # noqa: F401
        import jax  # noqa: F401

        return "import jax, differt2d: OK"
    except Exception as e:  # noqa: BLE001
        return f"import jax, differt2d failed ({type(e).__name__}: {e}) -> restated reference (oracle/ref_torch.py)"


class CpuReference:
    """
    The CPU arm: ALL 785 candidates of the workload on a strided subset of its receivers, forward + VJP.
    Timed implementation: oracle/d2d_oracle_ad.cpp — the compiled scalar port (every path evaluated literally and
    completely as the reference does under vmap, nothing pruned; paths with a non-zero validity re-evaluated with dual
    numbers for the cotangents), OpenMP over the receivers on every host core.  It is the FASTEST CPU statement of the
    path available here, i.e. the conservative denominator for a speed-up (JAX is not installable: probed every run).
    Context figure (reference arm only): the torch-CPU eager port that mirrors the reference's own design (array ops
    vectorised over the receivers like jax.vmap, Python loop over the candidates, reverse-mode tape).
    """

    SAMPLE = (128, 128)

    def __init__(self, coords: str):
        from oracle import c_oracle as CO

        self.CO = CO
        self.threads = CO.num_threads()
        self.sc = load_scene(coords)
        self.xys, _, _ = self.sc.packed_objects()
        self.fixed = np.stack([p.xy for p in self.sc.transmitters.values()])
        n_s, m_s = self.SAMPLE
        X, Y = self.sc.grid(GRID_PER_GPU[1], GRID_PER_GPU[0])
        sr, scl = GRID_PER_GPU[0] // n_s, GRID_PER_GPU[1] // m_s
        self.grid = np.stack([X[sr // 2:: sr, scl // 2:: scl], Y[sr // 2:: sr, scl // 2:: scl]], -1).reshape(-1, 2).astype(np.float32)
        self.zbar = np.random.default_rng(99).standard_normal(self.grid.shape[0]).astype(np.float32)
        self.n_cand = n_candidates(self.xys.shape[0], 0, MAX_ORDER)
        self.paths = self.grid.shape[0] * self.n_cand
        self.sample = (f"{n_s}x{m_s} receivers (every {sr}th row / {scl}th column of the 1024x1024 grid) x ALL "
                       f"{self.n_cand} candidates, forward + VJP, compiled scalar C++ port with dual numbers "
                       f"(oracle/d2d_oracle_ad.cpp), OpenMP")

    def step(self):
        return self.CO.power_vjp(self.xys, self.fixed, self.grid, self.zbar, max_order=MAX_ORDER, mode=MODE, alpha=ALPHA)

    def torch_port(self) -> dict:
        """one repetition of the torch-CPU eager port (oracle/ref_torch.py) on 16 x 16 receivers x all candidates"""
        import torch

        from oracle import ref_torch as R
        from tests import helpers as H

        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        osc = H.oracle_scene_from_product(self.sc)
        X, Y = self.sc.grid(GRID_PER_GPU[1], GRID_PER_GPU[0])
        Xs, Ys = X[32::64, 32::64], Y[32::64, 32::64]
        t0 = time.perf_counter()
        with R.clean_gradients():
            R.power_map_and_vjp(osc, Xs, Ys, None, max_order=MAX_ORDER, approx=True, alpha=ALPHA, function=MODE)
        dt = time.perf_counter() - t0
        return {"value": Xs.size * self.n_cand / dt, "unit": UNIT, "cores": threads,
                "sample": f"{Xs.shape[0]}x{Xs.shape[1]} receivers x ALL {self.n_cand} candidates, forward + autograd VJP, "
                          "torch-CPU fp32 eager (dispatch-bound: ~300 array ops per candidate), 1 repetition"}

    def describe(self, value: float, reps: int) -> dict:
        return {"value": value, "unit": UNIT, "cores": self.threads, "kind": "port", "sample": self.sample,
                "reps": reps, "note": "compiled restated reference; " + jax_probe()}


def workload_config(args, world):
    if args.scaling == "strong":
        rows = f"{GRID_STRONG[0]}x{GRID_STRONG[1]} receivers in total (BASELINE config 3), row-sharded"
        gg = list(GRID_STRONG)
    else:
        rows = f"{GRID_PER_GPU[0]}x{GRID_PER_GPU[1]} receivers per GPU"
        gg = [GRID_PER_GPU[0] * world, GRID_PER_GPU[1]]
    return {
        "workload": f"example.geojson city scene ({args.coords} coordinates), ImagePath orders 0-{MAX_ORDER} "
                    f"(785 candidates), {MODE} alpha={ALPHA:g}, forward + VJP, {rows}",
        "grid_global": gg,
        "sharding": f"receiver-grid rows over {world} GPU(s), bands of 8 rows dealt round robin; NCCL all-reduce of "
                    "scene-parameter cotangents",
        "l2": f"L2 flushed between timed steps ({FLUSH_BYTES >> 20} MiB memset, inside the timed region)",
        "step": "one CUDA-graph replay = L2 flush + d2d_power_fwd + d2d_power_bwd (+ all-reduce)",
    }


def bench_reference(args, rank: int, world: int) -> None:
    """The restated reference on the host cores (rank 0 only)."""
    if rank != 0:
        return
    ref = CpuReference(args.coords)
    for _ in range(max(1, min(args.warmup, 1))):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = (time.perf_counter() - t0) / args.steps
    value = ref.paths / dt
    cfg = workload_config(args, world)
    cfg["workload"] += f" — CPU arm timed on a bounded sample: {ref.sample}"
    base = ref.describe(value, args.steps)
    try:
        base["torch_port"] = ref.torch_port()
    except Exception as e:  # noqa: BLE001
        base["torch_port"] = {"error": repr(e)[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def cpu_baseline_sample(args) -> dict:
    """Bounded CPU sample of the same workload, on rank 0 at N = 1 only: the SAME sample and code as `--impl reference`."""
    ref = CpuReference(args.coords)
    ref.step()  # warm-up (thread pool)
    t0 = time.perf_counter()
    reps = 0
    while reps < 8 and time.perf_counter() - t0 < 12.0:
        ref.step()
        reps += 1
    return ref.describe(ref.paths * reps / (time.perf_counter() - t0), reps)


def _emit(line: dict) -> None:
    """The ONE JSON line on the real stdout (fd 1 is pointed at stderr while the benchmark runs: NCCL and other
    native libraries print banners on stdout)."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


_REAL_STDOUT = None


def _shutdown(D, torch, dev, graphs=()) -> None:
    """Leaves the process group without ever hanging the launcher: CUDA graphs that captured NCCL kernels are released
    FIRST (destroying a communicator that live graphs still reference blocked for the whole time limit on a 2-GPU
    run), and a watchdog ends the process if the teardown itself does not return."""
    import threading

    for g in graphs:
        try:
            if g is not None:
                g.reset()
        except Exception:  # noqa: BLE001
            pass
    try:
        torch.cuda.synchronize(dev)
    except Exception:  # noqa: BLE001
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(30.0, lambda: os._exit(0)).start()  # (daemon-like: only fires if destroy_process_group blocks)
    D.shutdown()
    os._exit(0)  # the JSON line is out (rank 0) and every rank has left the group: nothing else to run


def _quiet_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


# ---- device-resident job: one rank's share of a row-sharded grid ----------------------------------------------------------
class CityJob:
    """This rank's rows of an (n_rows x n_cols) receiver grid over the city scene, buffers resident in HBM, the step
    issued through the C ABI (device pointers, the current stream)."""

    def __init__(self, torch, L, F, D, sc, n_rows, n_cols, world, rank, dev, scene_arrays=None, mode=MODE, alpha=ALPHA,
                 seed=1234):
        self.torch, self.L, self.F, self.D = torch, L, F, D
        self.dev, self.world, self.rank = dev, world, rank
        xys, kinds, phis = sc.packed_objects()
        self.xys, self.fixed = xys, np.stack([p.xy for p in sc.transmitters.values()])
        X, Y = sc.grid(n_cols, n_rows)
        self.rows = D.row_tiles_cyclic(n_rows, world, rank)  # bands of 8 rows, round robin: equal work on every rank
        self.n_cols = n_cols
        self.grid_h = np.stack([X[self.rows], Y[self.rows]], -1).reshape(-1, 2).astype(np.float32)
        self.R = self.grid_h.shape[0]
        self.N, self.T = xys.shape[0], self.fixed.shape[0]
        self.n_cand = n_candidates(self.N, 0, MAX_ORDER)
        self.cfg = F.TraceConfig(mode=mode, max_order=MAX_ORDER, reduce_all=True, grid_cols=n_cols)
        self.zbar_h = np.random.default_rng(seed + rank).standard_normal(self.R).astype(np.float32)
        self.grid = torch.from_numpy(self.grid_h).to(dev)
        self.zbar = torch.from_numpy(self.zbar_h).to(dev)
        self.pk = F._Packed(self.cfg, torch.from_numpy(xys).to(dev), None, None, torch.from_numpy(self.fixed).to(dev),
                            self.grid, alpha, None, dev)
        self.mask = self.pk.new_mask()  # the VJP's only residual, written by the forward launch of the step
        self.Z = torch.empty(self.R, device=dev)
        self.gbar = torch.empty(self.R, 2, device=dev)
        n = self.N
        self.pbar = torch.zeros(n * 4 + n + 2 * self.T + 1, device=dev)  # objects | phis | fixed | alpha: one NCCL buffer
        self.off = (0, n * 4, n * 5, n * 5 + 2 * self.T)
        self.lib = L.lib()
        self.graph = None

    def launch(self, stream, events=None):
        """forward + backward (+ all-reduce) on `stream`; optional (e0, e1, e2, e3) events around the three parts"""
        L, lib, pk = self.L, self.lib, self.pk
        base, (o_obj, o_phi, o_fix, o_alpha) = self.pbar.data_ptr(), self.off
        if events:
            events[0].record(stream)
        L.check(lib.d2d_power_fwd(C.byref(pk.p), self.Z.data_ptr(), None, stream.cuda_stream), "d2d_power_fwd")
        if events:
            events[1].record(stream)
        L.check(lib.d2d_power_bwd(C.byref(pk.p), self.zbar.data_ptr(), None, self.gbar.data_ptr(), base + 4 * o_obj,
                                  base + 4 * o_phi, base + 4 * o_fix, base + 4 * o_alpha, stream.cuda_stream),
                "d2d_power_bwd")
        if events:
            events[2].record(stream)
        if self.world > 1:
            self.D.allreduce_sum_(self.pbar)
        if events:
            events[3].record(stream)

    def step_eager(self, flush, events=None):
        flush.zero_()
        self.launch(self.torch.cuda.current_stream(self.dev), events)

    def capture(self, flush):
        """One step as a CUDA graph (flush + forward + backward + all-reduce); False when capture is not possible."""
        torch = self.torch
        try:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(g):
                flush.zero_()
                self.launch(torch.cuda.current_stream(self.dev))
            self.graph = g
            return True
        except Exception as e:  # noqa: BLE001
            self.graph = None
            self.graph_error = repr(e)[:200]
            try:
                torch.cuda.synchronize(self.dev)
            except Exception:  # noqa: BLE001
                pass
            return False

    def step(self, flush):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.step_eager(flush)

    def timed(self, flush, steps, warmup, barrier=None):
        """ms per step over `steps` replays, CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            self.step(flush)
        torch.cuda.synchronize(self.dev)
        if barrier:
            barrier()
        stream = torch.cuda.current_stream(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            self.step(flush)
        e1.record(stream)
        torch.cuda.synchronize(self.dev)
        if barrier:
            barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            ms = self.D.max_over_ranks(ms, self.dev)
        return ms / steps


def dense_leg(torch, L, F, dev, flush, steps, peak_tf):
    """
    The dense regime (BASELINE config 2, examples/plot_power_profiles.py:118 alpha sweep): obstacle scene (8 walls),
    orders 0-2 (65 candidates), sigmoid, alpha = 1, 1024 x 1024 receivers, forward + VJP.  sigmoid(alpha x) is never
    exactly 0 or 1 at alpha = 1, so no cull, early exit or activity mask removes a path: every (receiver, candidate)
    path is constructed, tested, valued and accumulated by the forward launch, and re-traced and reversed by the
    backward launch.  The one thing the kernels still avoid (exactly) is the occlusion fold of a path whose validity
    cannot depend on it (fold_skip_bound, csrc/d2d_device.cuh: min(a_on, a_l) <= 1 - act(0.51)): ~3/4 of the paths.
    Algorithmic flop of SURVEY §8(d) (VJP = 2 x forward; the skipped folds still count, as everywhere else) /
    measured time / measured FP32 FMA peak.
    """
    import differt2d_b200 as d

    sc = d.Scene.square_scene_with_obstacle()
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    n = 1024
    X, Y = sc.grid(n, n)
    grid = torch.from_numpy(np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)).to(dev)
    R = grid.shape[0]
    cfg = F.TraceConfig(mode="sigmoid", max_order=2, reduce_all=True, grid_cols=n)
    pk = F._Packed(cfg, torch.from_numpy(xys).to(dev), None, None, torch.from_numpy(fixed).to(dev), grid, 1.0, None, dev)
    pk.new_mask()
    Z, gbar = torch.empty(R, device=dev), torch.empty(R, 2, device=dev)
    ob, fb, ab = torch.empty(8, 2, 2, device=dev), torch.empty(1, 2, device=dev), torch.empty(1, device=dev)
    zbar = torch.from_numpy(np.random.default_rng(5).standard_normal(R).astype(np.float32)).to(dev)
    lib = L.lib()
    stream = torch.cuda.current_stream(dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]

    def one(e=None):
        flush.zero_()
        if e:
            e[0].record(stream)
        L.check(lib.d2d_power_fwd(C.byref(pk.p), Z.data_ptr(), None, stream.cuda_stream), "d2d_power_fwd (dense leg)")
        if e:
            e[1].record(stream)
        L.check(lib.d2d_power_bwd(C.byref(pk.p), zbar.data_ptr(), None, gbar.data_ptr(), ob.data_ptr(), None, fb.data_ptr(),
                                  ab.data_ptr(), stream.cuda_stream), "d2d_power_bwd (dense leg)")
        if e:
            e[2].record(stream)

    for _ in range(3):
        one()
    torch.cuda.synchronize(dev)
    for i in range(steps):
        one(ev[i])
    torch.cuda.synchronize(dev)
    fwd = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    bwd = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    f_fwd = flops_per_receiver(8, 2, smooth=True) * R
    alive = float((Z != 0).float().mean())
    out = {
        "workload": "BASELINE config 2: square_scene_with_obstacle, ImagePath orders 0-2 (65 candidates), sigmoid "
                    "alpha=1, 1024x1024 receivers, forward + VJP (receivers, vertices, TX, alpha); every path is traced, valued "
                    "and reversed (validity never 0); the occlusion fold is skipped where it provably cannot change the validity",
        "steps": steps, "fwd_ms": fwd, "bwd_ms": bwd, "ms": fwd + bwd,
        "algorithmic_flop": {"fwd": f_fwd, "vjp": 2.0 * f_fwd, "total": 3.0 * f_fwd},
        "achieved_tflops": 3.0 * f_fwd / ((fwd + bwd) * 1e-3) / 1e12,
        "fwd_tflops": f_fwd / (fwd * 1e-3) / 1e12, "bwd_tflops": 2.0 * f_fwd / (bwd * 1e-3) / 1e12,
        "paths_per_s": R * 65 / ((fwd + bwd) * 1e-3),
        "receivers_with_nonzero_map": alive,
    }
    out["frac_of_fp32_fma_peak"] = out["achieved_tflops"] / peak_tf if peak_tf else None
    return out


def parity_spotcheck(job, Z, gbar) -> dict:
    """A strided 64 x 64 subset of THIS run's outputs against the CPU oracle: Z of the last forward launch (scalar C port)
    and grid_bar of the last backward launch (dual-number VJP of the scalar C++ port, fp32 and fp64)."""
    from oracle import c_oracle as CO

    n_rows, n_cols = job.R // job.n_cols, job.n_cols
    Zh = Z.cpu().numpy().reshape(n_rows, n_cols)
    gh = gbar.cpu().numpy().reshape(n_rows, n_cols, 2)
    G = job.grid_h.reshape(n_rows, n_cols, 2)
    ri, ci = np.arange(3, n_rows, max(n_rows // 64, 1))[:64], np.arange(5, n_cols, max(n_cols // 64, 1))[:64]
    sub = G[np.ix_(ri, ci)].reshape(-1, 2)
    Zo = CO.power_map(job.xys, job.fixed, sub, max_order=MAX_ORDER, mode=job.cfg.mode, alpha=ALPHA, reduce_all=True)
    got = Zh[np.ix_(ri, ci)].reshape(-1)
    denom = np.maximum(np.abs(Zo), 1e-6 * max(np.abs(Zo).max(), 1e-30))
    z_rel = float((np.abs(got - Zo) / denom).max())
    # grid_bar of the last backward launch on the same receivers: dual-number VJP of the scalar port (fp32), with the
    # fp64 evaluation of the same function telling which entries are meaningful in fp32 at all on these coordinates
    zb = job.zbar_h.reshape(n_rows, n_cols)[np.ix_(ri, ci)].reshape(-1)
    a32 = CO.power_vjp(job.xys, job.fixed, sub, zb, max_order=MAX_ORDER, mode=job.cfg.mode, alpha=ALPHA)
    a64 = CO.power_vjp(job.xys, job.fixed, sub, zb, max_order=MAX_ORDER, mode=job.cfg.mode, alpha=ALPHA, real64=True)
    gg = gh[np.ix_(ri, ci)].reshape(-1, 2).astype(np.float64)
    w32, w64 = a32["grid"], a64["grid"]
    scale = max(np.abs(w32).max(), 1e-30)
    # well conditioned in fp32: the oracle's own fp32 evaluation is within 1e-5 of its fp64 value (an order below the
    # bar, so that a second fp32 evaluation of the same size of error — the kernel's — can be held to rtol 1e-4)
    agree = np.abs(w32 - w64) <= 1e-5 * np.abs(w64) + 1e-7 * scale
    tol = 1e-4 * np.abs(w32) + 1e-6 * scale  # the elementwise bar of the tests (tests/test_gpu_parity._close)
    err = np.abs(gg - w32) / tol
    noise = np.abs(w32 - w64)
    within_noise = np.abs(gg - w32) <= tol + 4.0 * noise + 4.0 * np.percentile(noise, 90)
    g_rel = float(err[agree].max()) if agree.any() else None
    return {"points": int(got.size), "max_rel": z_rel, "nonzero_points": int((Zo != 0).sum()),
            "z": "Z of the timed loop's last forward launch vs oracle/d2d_oracle.c on a strided 64x64 subset",
            "grid_bar": {"entries": int(w32.size), "nonzero": int((w32 != 0).sum()),
                         "well_conditioned_entries": int(agree.sum()), "max_err_over_tol_on_those": g_rel,
                         "tol": "|kernel - oracle32| <= 1e-4 |oracle32| + 1e-6 max|oracle32| (elementwise)",
                         "all_within_oracle_fp32_noise": bool(within_noise.all()),
                         "vs": "dual-number VJP of oracle/d2d_oracle_ad.cpp in fp32; 'well conditioned' = its fp32 and fp64 "
                               "evaluations agree to 1e-5 (raw lon/lat coordinates put fp32 noise on the reference's own "
                               "cotangents, tests/test_gpu_parity_round2.py); everywhere: |kernel - oracle32| <= rtol 1e-4 "
                               "+ 4 |oracle32 - oracle64| + 4 P90"},
            "ok": bool(z_rel <= 1e-5 and (g_rel is None or g_rel <= 1.0) and within_noise.all())}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--coords", default="raw", choices=["raw", "normalised"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--workload", default="city", choices=["city", "p2p500"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the dense leg, the strong-scaling leg and the spot check")
    ap.add_argument("--only-dense", action="store_true", help="profiling aid: run the dense leg alone and print it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
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
    barrier = D.barrier if dist is not None else None
    if args.workload == "p2p500":
        bench_p2p500(args, torch, L, F, D, dev, world, rank, dist)
        return

    if args.only_dense:
        flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
        _emit(dense_leg(torch, L, F, dev, flush, min(args.steps, 10), 72.5))
        return

    sc = load_scene(args.coords)
    n_rows, n_cols = (GRID_STRONG if args.scaling == "strong" else (GRID_PER_GPU[0] * world, GRID_PER_GPU[1]))
    job = CityJob(torch, L, F, D, sc, n_rows, n_cols, world, rank, dev)
    lib = L.lib()
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)

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

    # eager warm-up (also initialises NCCL), then the per-kernel split of a step from an eager pass with events
    for _ in range(args.warmup):
        job.step_eager(flush)
    torch.cuda.synchronize(dev)
    n_split = min(args.steps, 10)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_split)]
    if barrier:
        barrier()
    for i in range(n_split):
        job.step_eager(flush, ev[i])
    torch.cuda.synchronize(dev)
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    bwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    per_rank = None
    if dist is not None:  # where a multi-GPU step goes, per rank (eager pass): kernels, all-reduce incl. waiting for the slowest
        ar_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
        mine = torch.tensor([fwd_ms, bwd_ms, ar_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"fwd_ms": float(t[0]), "bwd_ms": float(t[1]), "allreduce_and_wait_ms": float(t[2])} for t in allr]

    graphed = (not args.no_graph) and job.capture(flush)
    for _ in range(args.warmup):
        job.step(flush)
    torch.cuda.synchronize(dev)
    if barrier:
        barrier()
    try:
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:  # noqa: BLE001
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    if barrier:
        barrier()
    launches0 = F.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    sampler.mark_start()
    t_start.record(stream)
    for _ in range(args.steps):
        job.step(flush)
    t_end.record(stream)
    torch.cuda.synchronize(dev)
    sampler.mark_stop()
    if barrier:
        barrier()
    # kernels of this library inside the timed region: 2 per step (forward, backward); a graph replay launches the
    # captured kernels without passing through the library's counter
    launches = 2 * args.steps if graphed else F.launch_count() - launches0
    clocks = None
    if rank == 0:
        try:
            clocks = sampler.stop()
        except Exception as e:  # noqa: BLE001  (the sampler must never take the measurement down)
            clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"clock sampler failed: {e!r}"[:200]], "samples": 0}
    elapsed_ms = t_start.elapsed_time(t_end)
    if dist is not None:
        elapsed_ms = D.max_over_ranks(elapsed_ms, dev)
    ms_per_step = elapsed_ms / args.steps

    # end to end through the host-buffer C ABI entry (pinned host memory in, host memory out)
    R, T, n_obj = job.R, job.T, job.N
    grid_p = torch.from_numpy(job.grid_h).pin_memory()
    zbar_p = torch.from_numpy(job.zbar_h).pin_memory()
    Z_p = torch.empty(R).pin_memory()
    gbar_p = torch.empty(R, 2).pin_memory()
    par_p = torch.zeros(job.pbar.numel()).pin_memory()  # objects | phis | fixed | alpha, as the device buffer
    o_obj, o_phi, o_fix, o_alpha = job.off
    hp = L.new_problem()
    hp.n_objects, hp.objects_xys = n_obj, job.xys.ctypes.data
    hp.n_fixed, hp.fixed_xy = T, job.fixed.ctypes.data
    hp.n_grid, hp.grid_xy = R, grid_p.data_ptr()
    hp.max_order, hp.mode, hp.alpha, hp.reduce_all = MAX_ORDER, L.MODE_HARD_SIGMOID, ALPHA, 1
    hp.grid_cols = n_cols
    pb = par_p.data_ptr()
    par_d = torch.zeros_like(job.pbar)

    def e2e_step():
        L.check(lib.d2d_power_host(C.byref(hp), zbar_p.data_ptr(), Z_p.data_ptr(), gbar_p.data_ptr(), pb + 4 * o_obj,
                                   None, pb + 4 * o_fix, pb + 4 * o_alpha, local), "d2d_power_host")
        if dist is not None:  # the cotangents this call produced, summed over the ranks, back on the host
            par_d.copy_(par_p, non_blocking=True)
            D.allreduce_sum_(par_d)
            par_p.copy_(par_d, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()

    for _ in range(max(1, args.warmup)):
        e2e_step()
    torch.cuda.synchronize(dev)
    if barrier:
        barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        D.barrier()
        e2e_s = D.max_over_ranks(e2e_s, dev)
    e2e_ms = e2e_s * 1e3 / args.steps
    par_bytes = int(par_p.numel() * 4) if dist is not None else 0
    h2d = int(job.grid_h.nbytes + job.zbar_h.nbytes + job.xys.nbytes + job.fixed.nbytes) + par_bytes
    d2h = int(R * 4 + R * 8 + n_obj * 16 + T * 8 + 4) + par_bytes
    # the host entry and the device entries must agree (same kernels): Z bit for bit
    job.step_eager(flush)
    torch.cuda.synchronize(dev)
    e2e_equal = bool(np.array_equal(Z_p.numpy(), job.Z.cpu().numpy()))

    extras = {}
    graphs = [job.graph]
    if not args.no_extras:
        # strong-scaling leg (BASELINE config 3: 2048 x 2048 in total), every rank takes part
        if args.scaling == "weak":  # (every rank takes part: collectives inside; errors here are fatal on purpose)
            sjob = CityJob(torch, L, F, D, sc, GRID_STRONG[0], GRID_STRONG[1], world, rank, dev)
            for _ in range(2):
                sjob.step_eager(flush)
            if not args.no_graph:
                sjob.capture(flush)
                graphs.append(sjob.graph)
            s_ms = sjob.timed(flush, min(args.steps, 10), 3, barrier)
            extras["strong"] = {
                "grid_global": list(GRID_STRONG), "n_gpus": world, "ms_per_step": s_ms,
                "value": float(GRID_STRONG[0]) * GRID_STRONG[1] * sjob.T * sjob.n_cand / (s_ms * 1e-3), "unit": UNIT,
                "steps": min(args.steps, 10), "scaling": "strong",
                "note": "BASELINE config 3 (2048x2048 receivers in total, row-sharded); same step as the headline"}
            del sjob
        if rank == 0:
            try:
                extras["dense"] = dense_leg(torch, L, F, dev, flush, min(args.steps, 10), peak_tf)
            except Exception as e:  # noqa: BLE001
                extras["dense_error"] = repr(e)[:300]
            try:
                extras["parity_spotcheck"] = parity_spotcheck(job, job.Z, job.gbar)
            except Exception as e:  # noqa: BLE001
                extras["parity_spotcheck"] = {"error": repr(e)[:300], "ok": False}
        if barrier:
            barrier()

    if rank != 0:
        if dist is not None:
            _shutdown(D, torch, dev, graphs)
        return

    paths_step = float(n_rows) * n_cols * T * job.n_cand  # whole job, all ranks
    f_rx = flops_per_receiver(n_obj, MAX_ORDER, smooth=True)
    kernels = {
        "power_fwd_kernel": {"ms": fwd_ms, "algorithmic_flop": f_rx * R * T},
        "power_bwd_kernel": {"ms": bwd_ms, "algorithmic_flop": 2.0 * f_rx * R * T},
    }
    for k in kernels.values():
        k["algorithmic_tflops"] = k["algorithmic_flop"] / (k["ms"] * 1e-3) / 1e12
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    # what the SMs issued for the same command under `ncu --set full` (committed capture; NOT measured by this run)
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics.json")))
    except Exception:  # noqa: BLE001
        pass
    nk, nd = ncu.get("kernels", {}), ncu.get("dense", {})
    traffic = None
    capture = None
    keys = ("duration_ms_under_ncu", "issue_slots_busy_pct", "fp32_lanes_busy_pct", "executed_fp32_flop", "warp_instructions",
            "achieved_occupancy_pct", "registers_per_thread", "stall_no_instruction_per_issue", "dram_bytes_read", "dram_bytes_write")
    if nd:  # DRAM bytes of one step of the dense leg (the leg roofline.frac is measured on), per launch pair
        traffic = sum(v["dram_bytes_read"] + v["dram_bytes_write"] for v in nd.values())
    if nk or nd:
        capture = {"headline_step": {k: {a: v[a] for a in keys if a in v} for k, v in nk.items()},
                   "dense_leg": {k: {a: v[a] for a in keys if a in v} for k, v in nd.items()},
                   "source": ("profiles/ncu_metrics.json: a COMMITTED `ncu --set full` capture of this command "
                              f"({ncu.get('capture', 'see profiles/README.md')}); constants, NOT measured by this run")}
    dense = extras.get("dense")
    algo_bytes = R * (8 + 4 + 4 + 8)
    hbm_peak = 6650.0
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {
        "bound": "fp32",
        "kernel": "power_fwd_kernel + power_bwd_kernel on the dense leg (BASELINE config 2, sigmoid alpha=1)",
        "achieved": dense["achieved_tflops"] if dense else None, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": dense["frac_of_fp32_fma_peak"] if dense else None,
        "traffic": traffic,
        "peak_source": "FP32 FMA-chain microbenchmark (d2d_fma_peak_launch) timed in this run; MEASURED_PEAKS.json holds "
                       "no FP32 CUDA-core figure",
        "ceiling": "the denominator counts 2 flop per lane per cycle (FFMA).  The bit-exactness contract compiles the "
                   "canonical geometry with -fmad=false (one IEEE op per reference op), so its multiply-add pairs issue as "
                   "FMUL + FADD (1 flop per slot) and every IEEE division / sqrt costs 8-10 slots for the 1 flop the "
                   "convention credits: the attainable fraction for this arithmetic contract is below 0.5",
        "frac_definition": "ALGORITHMIC flop of SURVEY §8(d) (forward + VJP = 3 x F_fwd) of the dense leg / its measured "
                           "time / peak.  No path is pruned there (every path is constructed, tested, valued, accumulated and "
                           "reversed); the occlusion fold — about half of F_fwd — is skipped, exactly, for the ~3/4 of the "
                           "paths whose validity cannot depend on it, and still counted (SURVEY §8(d): no early-out "
                           "credit).  What the SMs executed is under ncu_capture.dense_leg (executed_fp32_flop, "
                           "fp32_lanes_busy_pct, issue_slots_busy_pct)",
        "dense_leg": dense,
        "algorithmic": {
            "note": "SURVEY §8(d) convention on the HEADLINE workload (pruned work still counted; VJP = 2 x forward): exact "
                    "pruning (macro / tile / warp culls, early exits, activity mask) removes > 99 % of it, so these figures "
                    "exceed the peak by construction — a pruning ratio, not a utilisation",
            "dominant_kernel": dom, "kernels": kernels,
            "frac_of_peak": kernels[dom]["algorithmic_tflops"] / peak_tf if peak_tf else None},
        "ncu_capture": capture,
        "hbm": {"algorithmic_bytes_per_launch": algo_bytes,
                "achieved_gbs": algo_bytes / (kernels[dom]["ms"] * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "note": "HBM is idle: 24 B per receiver against ~1e6 flop"},
    }
    line = {
        "metric": METRIC, "value": paths_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "power_map_plus_vjp_ms": ms_per_step,
        "cuda_graph": bool(graphed),
        "e2e": {"value": paths_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "equals_device_entry": e2e_equal,
                "call": "d2d_power_host (fused value+VJP, pinned host buffers in/out, sync)"
                        + ("; then the produced parameter cotangents are all-reduced (NCCL) and read back" if world > 1 else "")},
        "gpu_launches": int(launches),
        "kernel_split_eager_pass": {"fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "steps": n_split},
        "per_rank": per_rank,
        "clocks": clocks,
        "roofline": roofline,
    }
    if not graphed and not args.no_graph:
        line["cuda_graph_error"] = getattr(job, "graph_error", None)
    for k in ("strong", "parity_spotcheck", "dense_error"):
        if k in extras:
            line[k] = extras[k]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args)
    _emit(line)
    if dist is not None:
        _shutdown(D, torch, dev, graphs)


def bench_p2p500(args, torch, L, F, D, dev, world, rank, dist) -> None:
    """
    BASELINE config 5 (second half): synthetic 500-wall scene (random_uniform_scene layout, numpy default_rng(1234)),
    ImagePath order 3 — 124 500 500 candidates per link, 1 TX x 2 RX — forward + VJP, the CANDIDATE list sharded over
    the GPUs, ONE all-reduce of [Z | receiver cotangents | scene-parameter cotangents] per step.
    Strong scaling: the job is the same at every N.
    """
    rng = np.random.default_rng(1234)
    pts = rng.random((1 + 2 * 500 + 2, 2), dtype=np.float32)  # scene.py:718-733 layout
    fixed = pts[:1]
    xys = pts[1:1001].reshape(500, 2, 2)
    grid = torch.from_numpy(pts[-2:][::-1].copy()).to(dev)
    cfg = F.TraceConfig(mode=MODE, min_order=3, max_order=3, cand_shard=(rank, world))
    n_cand = 500 * 499 * 499
    want = ("grid", "objects", "fixed", "alpha")

    def step():
        out = F.power_value_and_vjp(cfg, xys, fixed, grid, None, alpha=ALPHA, want=want, device=dev)
        buf = torch.cat([out[k].reshape(-1) for k in ("Z", *want)])
        if dist is not None:
            D.allreduce_sum_(buf)
        return buf

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    if dist is not None:
        D.barrier()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = F.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        buf = step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    launches = F.launch_count() - l0
    ms = e0.elapsed_time(e1)
    if dist is not None:
        D.barrier()
        ms = D.max_over_ranks(ms, dev)
    ms /= args.steps
    if rank == 0:
        _emit({"metric": METRIC, "value": 2.0 * n_cand / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "gpu_launches": int(launches),
               "config": {"workload": "BASELINE config 5: 500 random walls, ImagePath order 3 (124 500 500 candidates per "
                                      f"link), 1 TX x 2 RX, {MODE} alpha={ALPHA:g}, forward + VJP",
                          "sharding": f"candidate list dealt to {world} GPU(s) in chunks of 128; one NCCL all-reduce of "
                                      f"{int(buf.numel())} floats per step"},
               "checksum": {"Z": [float(v) for v in buf[:2].cpu()]}})
    if dist is not None:
        _shutdown(D, torch, dev)


if __name__ == "__main__":
    main()
