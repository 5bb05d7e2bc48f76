"""
Builds libdiffert2d_b200.so (hand-written sm_100a CUDA + the C ABI of include/differt2d_b200.h)
IN-TREE with nvcc, so that the shared object travels to the GPU box with the repository snapshot.

    python -m differt2d_b200.build [--force] [--verbose]

The kernel translation units (one per source and logic mode) are compiled with `-fmad=false` (see csrc/d2d_device.cuh,
"Arithmetic discipline"); FMAs are requested explicitly where wanted.
"""

from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libdiffert2d_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]
# (source, object, extra flags): the kernel sources are compiled once per logic mode, in parallel
UNITS = [("d2d_forward.cu", f"d2d_forward_m{m}.o", ["-fmad=false", f"-DD2D_TU_MODE={m}"]) for m in (1, 0, 2)]
UNITS += [("d2d_backward.cu", f"d2d_backward_m{m}.o", ["-fmad=false", f"-DD2D_TU_MODE={m}"]) for m in (1, 0, 2)]
# reverse mode through the Adam scan of FermatPath / MinPath: the slowest units, listed first
UNITS = [("d2d_backward.cu", f"d2d_backward_solver_m{m}.o", ["-fmad=false", f"-DD2D_TU_MODE={m}", "-DD2D_TU_SOLVER=1"])
         for m in (1, 0, 2)] + UNITS
UNITS += [("d2d_paths.cu", f"d2d_paths_m{m}.o", ["-fmad=false", f"-DD2D_TU_MODE={m}"]) for m in (1, 0, 2)]
UNITS += [("d2d_nan.cu", "d2d_nan.o", ["-fmad=false"])]
UNITS += [("d2d_sanitise.cu", "d2d_sanitise.o", ["-fmad=false"])]
UNITS += [("d2d_paths_bwd.cu", "d2d_paths_bwd.o", ["-fmad=false"])]
UNITS += [("d2d_abi.cu", "d2d_abi.o", [])]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _deps():
    out = []
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force: bool = False, verbose: bool = False, debug_counters: bool = False) -> str:
    """debug_counters: diagnostic library libdiffert2d_b200_dbg.so (forward units only count; select it at run time
    with D2D_B200_LIB=<path>); never the default."""
    if debug_counters:
        return _build(verbose, ["-DD2D_DEBUG_COUNTERS"], "dbg_", os.path.join(OUT_DIR, "libdiffert2d_b200_dbg.so"))
    if not force and not needs_build():
        return LIB
    return _build(verbose, [], "", LIB)


def _build(verbose: bool, defs, prefix: str, lib_path: str) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def compile_one(item):
        src, objname, extra = item
        obj = os.path.join(OUT_DIR, prefix + objname)
        tune = os.environ.get("D2D_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DD2D_BWD_MIN_CTAS=4
        cmd = [nvcc, *ARCH, *COMMON, *extra, *defs, *tune, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, flush=True)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, UNITS))
    cmd = [nvcc, *ARCH, "-shared", "-o", lib_path, *objs, "--cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


def build_jax_ffi(verbose: bool = False) -> str:
    """Compiles jax_ffi/d2d_xla_ffi.cc (the XLA-FFI handlers over the C ABI) against jax.ffi.include_dir().
    Needs `import jax` to work: not the case in the image this repository was developed in."""
    try:
        import jax
    except ImportError as e:  # loud, never a silent skip
        raise RuntimeError("--jax-ffi needs JAX (jax.ffi.include_dir() holds the XLA-FFI headers)") from e
    build()
    out = os.path.join(OUT_DIR, "libdiffert2d_b200_xla.so")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include") if os.path.isabs(_nvcc()) \
        else "/usr/local/cuda/include"
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + jax.ffi.include_dir(), "-I" + cuda_inc,
           os.path.join(HERE, "jax_ffi", "d2d_xla_ffi.cc"), "-L" + OUT_DIR, "-ldiffert2d_b200",
           "-Wl,-rpath,$ORIGIN", "-o", out]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for d2d_xla_ffi.cc:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    if "--jax-ffi" in sys.argv:
        print(build_jax_ffi(verbose="--verbose" in sys.argv))
        sys.exit(0)
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                 debug_counters="--debug-counters" in sys.argv)
    print(path)
