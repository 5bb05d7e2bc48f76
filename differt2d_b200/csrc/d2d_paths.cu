// differt2d_b200 — path materialisation kernel: Scene.all_paths / all_valid_paths (scene.py:1156-1248) and
// the generic-`fun` escape hatch.
//
// Same CTA driver as the forward kernel (tile cull, warp cull, trace), but instead of accumulating
// valid * fun it appends one fixed-size record per surviving (fixed point, grid point, candidate) to a
// caller-owned array through a global atomic counter: indices, validity, loss, the fused `fun` value and the
// k + 2 path vertices.  Two modes:
//   * emit_all = 0: records with validity > min_valid only (all_valid_paths; min_valid = 0.5 is logic.is_true);
//     dead paths leave through the same exact early-outs as in the forward kernel;
//   * emit_all = 1: every triple, as all_paths yields them — no cull, no early-out, the complete path and its loss
//     even when the path is invalid.
// Records land in arrival order; the host sorts them by (fixed, grid, candidate) when list order matters.
// With records == nullptr the kernel only counts (sizing pass).
#include "d2d_driver.cuh"
#include "d2d_launch.h"
#include "d2d_solver.cuh"
#include "d2d_newton.cuh"

namespace d2d {

template <int MODE, int METHOD, int K, bool TXGRID>
__device__ __forceinline__ void run_order_paths(const SceneTab& T, const KParams& p, const Tile& tile, DriverShared& sh,
                                                const float alpha, const int t, const float2 fx, const float2 g,
                                                const long long col0, int& buf, const PathsOut out) {
    const float2 tx = TXGRID ? g : fx;
    const float2 rx = TXGRID ? fx : g;
    for_each_candidate<MODE, METHOD, K, TXGRID>(
        T, p, tile, sh, alpha, t, fx, col0, nullptr, buf, [&](const Cand<K>& cd, const long long col, const float2 apex) {
            if (!tile.active) return;
            float2 X[K + 2];
            float valid, loss = 0.0f;
            constexpr bool kSolverLoss = (METHOD == D2D_METHOD_MINPATH) && K > 0;
            if (out.emit_all) {
                construct_path<METHOD, K>(T, p, cd, tx, rx, col, X, loss);
                if (!kSolverLoss) loss = path_loss<K>(T, cd, X);  // geometry.py:1077-1084, 1202-1204
                valid = validity<MODE, K, false>(T, p, alpha, cd, X, loss, &sh.hint[threadIdx.x >> 5]);
            } else {
                if constexpr (METHOD == D2D_METHOD_IMAGE) {
                    float onx;
                    const float2 ap = TXGRID ? image_apex<K>(T, cd, tx) : apex;
                    if (!image_path_on<MODE, K>(T, cd, tx, rx, ap, alpha, X, onx)) return;
                    valid = validity_from_onx<MODE, K, true>(T, p, alpha, cd, X, 0.0f, onx, &sh.hint[threadIdx.x >> 5]);
                } else {
                    construct_path<METHOD, K>(T, p, cd, tx, rx, col, X, loss);
                    valid = validity<MODE, K, !kSolverLoss>(T, p, alpha, cd, X, loss, &sh.hint[threadIdx.x >> 5]);
                }
                if (!(valid > out.min_valid)) return;
                if (!kSolverLoss) loss = path_loss<K>(T, cd, X);
            }
            const unsigned long long slot = atomicAdd(out.count, 1ULL);
            if (out.records == nullptr || slot >= (unsigned long long)out.capacity) return;
            float r;
            const float value = path_value<K>(p, X, r);
            D2DPathRecord rec;
            rec.fixed = t;
            rec.order = K;
            rec.grid = tile.r;
            rec.candidate = col;
            rec.valid = valid;
            rec.loss = loss;
            rec.value = value;
            rec.length = r;
#pragma unroll
            for (int i = 0; i < D2D_MAX_ORDER + 2; ++i) {
                rec.xys[2 * i] = i < K + 2 ? X[i < K + 2 ? i : 0].x : 0.0f;
                rec.xys[2 * i + 1] = i < K + 2 ? X[i < K + 2 ? i : 0].y : 0.0f;
            }
            out.records[slot] = rec;
        });
}

template <int MODE, int METHOD, bool TXGRID>
__global__ void __launch_bounds__(kBlock) paths_kernel(const KParams p, const PathsOut out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ DriverShared sh;
    SceneTab T = carve_tab(smem, p.N);
    build_tab(T, p, &sh.count);
    const Tile tile = make_tile(p, T, sh);
    const float alpha = p.alpha_dev ? *p.alpha_dev : p.alpha;
    if constexpr (METHOD == D2D_METHOD_IMAGE) {
        if (p.macro) macro_prologue<MODE, TXGRID>(T, p, tile, sh, alpha);
    }
    const float2 g = tile.active ? reinterpret_cast<const float2*>(p.grid)[tile.r] : make_float2(0.f, 0.f);
    int buf = 0;
    for (int t = 0; t < p.T; ++t) {
        const float2 fx = reinterpret_cast<const float2*>(p.fixed)[t];
        long long col0 = 0;
        for (int k = p.min_order; k <= p.max_order; ++k) {
            switch (k) {
                case 0: run_order_paths<MODE, METHOD, 0, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, out); break;
                case 1: run_order_paths<MODE, METHOD, 1, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, out); break;
                case 2: run_order_paths<MODE, METHOD, 2, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, out); break;
                case 3: run_order_paths<MODE, METHOD, 3, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, out); break;
                case 4: run_order_paths<MODE, METHOD, 4, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, out); break;
                default: break;
            }
            col0 += order_count(k, T.n_allowed);
        }
    }
}

template <int MODE, int METHOD, bool TXGRID>
static int launch_paths_one(const KParams& p, const PathsOut& out, cudaStream_t stream) {
    const size_t smem = scene_tab_bytes(p.N);
    auto kern = paths_kernel<MODE, METHOD, TXGRID>;
    if (smem > 32 * 1024) {  // static + dynamic > 48 KB needs the opt-in; the static part is ~11 KB
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    KParams q = p;
    q.macro = (host_macro_ok(p) && METHOD == D2D_METHOD_IMAGE) ? 1 : 0;  // both grid roles
    const cudaError_t e = launch_tiles(kern, q, q.macro != 0, smem, stream, q, out);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

template <int MODE, int METHOD>
static int launch_paths_role(const KParams& p, int grid_role, const PathsOut& out, cudaStream_t s) {
    return grid_role == D2D_GRID_TRANSMITTERS ? launch_paths_one<MODE, METHOD, true>(p, out, s)
                                              : launch_paths_one<MODE, METHOD, false>(p, out, s);
}

#ifndef D2D_TU_MODE
#error "compile with -DD2D_TU_MODE=<D2D_MODE_*> (differt2d_b200/build.py)"
#endif

template <>
int launch_paths_mode<D2D_TU_MODE>(const KParams& p, int grid_role, int method, const PathsOut& out, cudaStream_t s) {
    switch (method) {
        case D2D_METHOD_IMAGE: return launch_paths_role<D2D_TU_MODE, D2D_METHOD_IMAGE>(p, grid_role, out, s);
        case D2D_METHOD_FERMAT: return launch_paths_role<D2D_TU_MODE, D2D_METHOD_FERMAT>(p, grid_role, out, s);
        case D2D_METHOD_MINPATH: return launch_paths_role<D2D_TU_MODE, D2D_METHOD_MINPATH>(p, grid_role, out, s);
    }
    return (int)cudaErrorInvalidValue;
}

#if D2D_TU_MODE == D2D_MODE_HARD
int launch_paths(const KParams& p0, int mode, int grid_role, int method, const PathsOut& out, cudaStream_t stream,
                 long long* launches) {
    cudaError_t e = cudaMemsetAsync(out.count, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return (int)e;
    if (p0.R <= 0) return 0;
    KParams p = p0;
    p.mask = nullptr;
    if (out.emit_all) p.cull = 0;  // all_paths yields every candidate, valid or not
    int rc;
    switch (mode) {
        case D2D_MODE_HARD: rc = launch_paths_mode<D2D_MODE_HARD>(p, grid_role, method, out, stream); break;
        case D2D_MODE_HARD_SIGMOID: rc = launch_paths_mode<D2D_MODE_HARD_SIGMOID>(p, grid_role, method, out, stream); break;
        case D2D_MODE_SIGMOID: rc = launch_paths_mode<D2D_MODE_SIGMOID>(p, grid_role, method, out, stream); break;
        default: return (int)cudaErrorInvalidValue;
    }
    if (launches) *launches += 1;
    return rc;
}
#endif

}  // namespace d2d
