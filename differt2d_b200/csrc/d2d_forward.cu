// differt2d_b200 — forward power-map kernel.
//
// Mapping: one thread per grid point, one CTA per compact tile of 128 points (d2d_driver.cuh).  The
// CTA walks the candidate list in list order (orders ascending, lexicographic inside an order —
// scene.py:166-175), so the per-point accumulation order is the reference's (scene.py:1893-1918).
// Every lane works on the same candidate at the same time: object-table reads are shared-memory
// broadcasts.  HBM traffic is the grid point (8 B) in and the map value (4 B) out per thread; the
// kernel is FP32-issue bound (DESIGN.md "Roofline").
#include "d2d_driver.cuh"
#include "d2d_launch.h"
#include "d2d_solver.cuh"
#include "d2d_newton.cuh"

namespace d2d {

template <int MODE, int METHOD, int K, bool TXGRID>
__device__ __forceinline__ void run_order(const SceneTab& T, const KParams& p, const Tile& tile, DriverShared& sh,
                                          const float alpha, const int t, const float2 fx, const float2 g,
                                          const long long col0, int& buf, float& acc, float* vrow, uint32_t* mrow) {
    const float2 tx = TXGRID ? g : fx;
    const float2 rx = TXGRID ? fx : g;
    for_each_candidate<MODE, METHOD, K, TXGRID>(
        T, p, tile, sh, alpha, t, fx, col0, nullptr, buf, [&](const Cand<K>& cd, const long long col, const float2 apex) {
            float2 X[K + 2];
            float valid = 0.0f;
            if (tile.active) {
                if constexpr (METHOD == D2D_METHOD_IMAGE) {
                    // construction fused with on_objects: most paths that reach this point die at an interaction
                    float onx;
                    const float2 ap = TXGRID ? image_apex<K>(T, cd, tx) : apex;
                    D2D_COUNT(16);
                    if (image_path_on<MODE, K>(T, cd, tx, rx, ap, alpha, X, onx)) {
                        D2D_COUNT(17);
                        bool dg = false;
                        for (int i = 0; i < K; ++i) dg = dg || (T.w0[cd.c[i]].z == 0.f && T.w0[cd.c[i]].w == 0.f);
                        if (dg) D2D_COUNT(18);
                        valid = validity_from_onx<MODE, K, true>(T, p, alpha, cd, X, 0.0f, onx, &sh.hint[threadIdx.x >> 5]);
                    }
                } else {
                    float loss;
                    construct_path<METHOD, K>(T, p, cd, tx, rx, col, X, loss);
                    valid = validity<MODE, K, (METHOD != D2D_METHOD_MINPATH) || K == 0>(T, p, alpha, cd, X, loss, &sh.hint[threadIdx.x >> 5]);
                }
                if (valid != 0.0f) {
                    D2D_COUNT(19);
                    float r;
                    acc = acc + valid * path_value<K>(p, X, r);  // scene.py:1909
                    if (vrow) vrow[col] = valid;                 // (valid_out is zero-filled by the launcher)
                }
            }
            if (mrow) {  // activity bit of (this warp, candidate) for the backward kernel; zero-filled by the launcher
                if (__any_sync(0xffffffffu, valid != 0.0f) && (threadIdx.x & 31) == 0)
                    atomicOr(&mrow[col >> 5], 1u << (col & 31));
            }
        });
}

template <int MODE, int METHOD, bool TXGRID>
__global__ void __launch_bounds__(kBlock, D2D_FWD_MIN_CTAS) power_fwd_kernel(const KParams p, float* __restrict__ Z,
                                                           float* __restrict__ valid_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ DriverShared sh;
    SceneTab T = carve_tab(smem, p.N);
    build_tab(T, p, &sh.count);
    const Tile tile = make_tile(p, T, sh);
    const float alpha = p.alpha_dev ? *p.alpha_dev : p.alpha;
    if (MODE != D2D_MODE_HARD && !(alpha > 0.0f)) {
        // a traced alpha cannot be validated on the host; the folds (act(min x) = min act(x)), the dead-path exits and
        // the culls all assume a non-decreasing activation: poison the outputs instead of returning wrong numbers
        // (uniform over the grid: every CTA of a cluster leaves here)
        if (tile.active && blockIdx.y == 0)
            for (int t = 0; t < (p.reduce_all ? 1 : p.T); ++t) Z[(long long)t * p.R + tile.r] = CUDART_NAN_F;
        return;
    }
    if (D2D_FOLD_SKIP) T.fold_skip = fold_skip_bound<MODE>(alpha);
    if constexpr (METHOD == D2D_METHOD_IMAGE) {
        if (p.macro) macro_prologue<MODE, TXGRID>(T, p, tile, sh, alpha);
    }
    const float2 g = tile.active ? reinterpret_cast<const float2*>(p.grid)[tile.r] : make_float2(0.f, 0.f);
    float zsum = 0.0f;
    int buf = 0;
    for (int t = 0; t < p.T; ++t) {
        const float2 fx = reinterpret_cast<const float2*>(p.fixed)[t];
        float acc = 0.0f;  // scene.py:1893
        long long col0 = 0;
        float* vrow = (valid_out && tile.active) ? valid_out + ((long long)t * p.R + tile.r) * p.C_total : nullptr;
        uint32_t* mrow = p.mask ? p.mask + ((long long)t * gridDim.x * (kBlock / 32) + (long long)blockIdx.x * (kBlock / 32) +
                                            (threadIdx.x >> 5)) * p.mask_wpw
                                : nullptr;
        for (int k = p.min_order; k <= p.max_order; ++k) {
            switch (k) {
                case 0: run_order<MODE, METHOD, 0, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, acc, vrow, mrow); break;
                case 1: run_order<MODE, METHOD, 1, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, acc, vrow, mrow); break;
                case 2: run_order<MODE, METHOD, 2, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, acc, vrow, mrow); break;
                case 3: run_order<MODE, METHOD, 3, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, acc, vrow, mrow); break;
                case 4: run_order<MODE, METHOD, 4, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, acc, vrow, mrow); break;
                default: break;
            }
            col0 += order_count(k, T.n_allowed);
        }
        if (tile.active) {
            if (p.reduce_all) zsum = zsum + acc;  // scene.py:1939-1952 (0.0 + p0 + p1 ...)
            else if (gridDim.y == 1) Z[(long long)t * p.R + tile.r] = acc;
            else if (acc != 0.0f) atomicAdd(&Z[(long long)t * p.R + tile.r], acc);  // Z zeroed by the launcher
        }
    }
    if (tile.active && p.reduce_all) {
        if (gridDim.y == 1) Z[tile.r] = zsum;
        else if (zsum != 0.0f) atomicAdd(&Z[tile.r], zsum);
    }
}

template <int MODE, int METHOD, bool TXGRID>
static int launch_one(const KParams& p, float* Z, float* valid_out, cudaStream_t stream) {
    const size_t smem = scene_tab_bytes(p.N);
    auto kern = power_fwd_kernel<MODE, METHOD, TXGRID>;
    if (smem > 32 * 1024) {  // static + dynamic > 48 KB needs the opt-in; the static part is ~11 KB
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    // thread-block clusters (macro-tile cull through distributed shared memory) where there is something to cull
    KParams q = p;
    q.macro = (host_macro_ok(p) && METHOD == D2D_METHOD_IMAGE) ? 1 : 0;  // both grid roles
    const cudaError_t e = launch_tiles(kern, q, q.macro != 0, smem, stream, q, Z, valid_out);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

template <int MODE, int METHOD>
static int launch_role(const KParams& p, int grid_role, float* Z, float* valid_out, cudaStream_t s) {
    return grid_role == D2D_GRID_TRANSMITTERS ? launch_one<MODE, METHOD, true>(p, Z, valid_out, s)
                                              : launch_one<MODE, METHOD, false>(p, Z, valid_out, s);
}

template <int MODE>
static int launch_method(const KParams& p, int grid_role, int method, float* Z, float* valid_out,
                         cudaStream_t s) {
    switch (method) {
        case D2D_METHOD_IMAGE: return launch_role<MODE, D2D_METHOD_IMAGE>(p, grid_role, Z, valid_out, s);
        case D2D_METHOD_FERMAT: return launch_role<MODE, D2D_METHOD_FERMAT>(p, grid_role, Z, valid_out, s);
        case D2D_METHOD_MINPATH: return launch_role<MODE, D2D_METHOD_MINPATH>(p, grid_role, Z, valid_out, s);
    }
    return (int)cudaErrorInvalidValue;
}

#ifndef D2D_TU_MODE
#error "compile with -DD2D_TU_MODE=<D2D_MODE_*> (differt2d_b200/build.py)"
#endif

template <>
int launch_fwd_mode<D2D_TU_MODE>(const KParams& p, int grid_role, int method, float* Z, float* valid_out,
                                 cudaStream_t s) {
    return launch_method<D2D_TU_MODE>(p, grid_role, method, Z, valid_out, s);
}

#ifdef D2D_DEBUG_COUNTERS
template <>
void debug_counters_mode<D2D_TU_MODE>(unsigned long long* out, int reset) {
    if (reset) {
        unsigned long long z[32] = {0};
        cudaMemcpyToSymbol(d2d_dbg, z, sizeof(z));
    } else {
        cudaMemcpyFromSymbol(out, d2d_dbg, sizeof(unsigned long long) * 32);
    }
}
#endif

#if D2D_TU_MODE == D2D_MODE_HARD
long long num_tile_blocks(const KParams& p) { return host_tile_blocks(p); }

int launch_power_fwd(const KParams& p, int mode, int grid_role, int method, float* Z, float* valid_out,
                     cudaStream_t stream, long long* launches) {
    if (p.R <= 0) return 0;
    if (p.slices > 1) {  // partial sums of the candidate slices are combined with atomics
        const cudaError_t e = cudaMemsetAsync(Z, 0, sizeof(float) * (size_t)(p.reduce_all ? 1 : p.T) * p.R, stream);
        if (e != cudaSuccess) return (int)e;
    }
    if (valid_out) {
        const cudaError_t e = cudaMemsetAsync(valid_out, 0, sizeof(float) * (size_t)p.T * p.R * p.C_total, stream);
        if (e != cudaSuccess) return (int)e;
    }
    if (p.mask) {
        const size_t words = (size_t)p.T * (size_t)host_tile_blocks(p) * (kBlock / 32) * (size_t)p.mask_wpw;
        const cudaError_t e = cudaMemsetAsync(p.mask, 0, sizeof(uint32_t) * words, stream);
        if (e != cudaSuccess) return (int)e;
    }
    int e;
    switch (mode) {
        case D2D_MODE_HARD: e = launch_fwd_mode<D2D_MODE_HARD>(p, grid_role, method, Z, valid_out, stream); break;
        case D2D_MODE_HARD_SIGMOID:
            e = launch_fwd_mode<D2D_MODE_HARD_SIGMOID>(p, grid_role, method, Z, valid_out, stream);
            break;
        case D2D_MODE_SIGMOID: e = launch_fwd_mode<D2D_MODE_SIGMOID>(p, grid_role, method, Z, valid_out, stream); break;
        default: return (int)cudaErrorInvalidValue;
    }
    if (launches) *launches += 1;
    return e;
}
#endif

}  // namespace d2d
