// differt2d_b200 — forward power-map kernel.
//
// Mapping: one thread per grid point; the CTA walks the candidate list in list order (orders
// ascending, lexicographic inside an order — scene.py:166-175) so that the per-point accumulation
// order is the reference's (scene.py:1893-1918).  Every lane of a warp works on the same candidate
// at the same time: object-table reads are shared-memory broadcasts and the candidate odometer
// lives in uniform registers.  Nothing but the grid point (8 B) is read from and the map value
// (4 B) written to HBM per thread; the kernel is FP32-issue bound (DESIGN.md "Roofline").
#include "d2d_launch.h"
#include "d2d_trace.cuh"
#include "d2d_solver.cuh"

namespace d2d {

template <int MODE, int METHOD, int K>
__device__ __forceinline__ void run_order(const SceneTab& T, const KParams& p, const float alpha,
                                          const float2 tx, const float2 rx, float& acc, float* vrow,
                                          long long& col) {
    Odometer<K> od;
    if (!od.first(T.n_allowed)) return;
    do {
        Cand<K> cd;
#pragma unroll
        for (int i = 0; i < K; ++i) cd.c[i] = T.allowed[od.pos[i]];
        float2 X[K + 2];
        float loss;
        construct_path<METHOD, K>(T, p, cd, tx, rx, col, X, loss);
        const float valid = validity<MODE, K, (METHOD != D2D_METHOD_MINPATH) || K == 0>(T, p, alpha, cd, X, loss);
        if (valid != 0.0f) {
            float r;
            acc = acc + valid * path_value<K>(p, X, r);  // scene.py:1909
        }
        if (vrow) vrow[col] = valid;
        ++col;
    } while (od.next(T.n_allowed));
}

template <int MODE, int METHOD, bool TXGRID>
__global__ void __launch_bounds__(128) power_fwd_kernel(const KParams p, float* __restrict__ Z,
                                                        float* __restrict__ valid_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_count;
    SceneTab T = carve_tab(smem, p.N);
    build_tab(T, p, &s_count);
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    const float alpha = p.alpha_dev ? *p.alpha_dev : p.alpha;
    const float2 g = reinterpret_cast<const float2*>(p.grid)[r];
    float zsum = 0.0f;
    for (int t = 0; t < p.T; ++t) {
        const float2 fx = reinterpret_cast<const float2*>(p.fixed)[t];
        const float2 tx = TXGRID ? g : fx;
        const float2 rx = TXGRID ? fx : g;
        float acc = 0.0f;  // scene.py:1893
        long long col = 0;
        float* vrow = valid_out ? valid_out + ((long long)t * p.R + r) * p.C_total : nullptr;
        for (int k = p.min_order; k <= p.max_order; ++k) {
            switch (k) {
                case 0: run_order<MODE, METHOD, 0>(T, p, alpha, tx, rx, acc, vrow, col); break;
                case 1: run_order<MODE, METHOD, 1>(T, p, alpha, tx, rx, acc, vrow, col); break;
                case 2: run_order<MODE, METHOD, 2>(T, p, alpha, tx, rx, acc, vrow, col); break;
                case 3: run_order<MODE, METHOD, 3>(T, p, alpha, tx, rx, acc, vrow, col); break;
                case 4: run_order<MODE, METHOD, 4>(T, p, alpha, tx, rx, acc, vrow, col); break;
                default: break;
            }
        }
        if (p.reduce_all) zsum = zsum + acc;  // scene.py:1939-1952 (0.0 + p0 + p1 ...)
        else Z[(long long)t * p.R + r] = acc;
    }
    if (p.reduce_all) Z[r] = zsum;
}

template <int MODE, int METHOD, bool TXGRID>
static int launch_one(const KParams& p, float* Z, float* valid_out, cudaStream_t stream) {
    const int block = 128;
    const long long nblk = (p.R + block - 1) / block;
    const size_t smem = scene_tab_bytes(p.N);
    auto kern = power_fwd_kernel<MODE, METHOD, TXGRID>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<(unsigned)nblk, block, smem, stream>>>(p, Z, valid_out);
    return (int)cudaGetLastError();
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

int launch_power_fwd(const KParams& p, int mode, int grid_role, int method, float* Z, float* valid_out,
                     cudaStream_t stream, long long* launches) {
    if (p.R <= 0) return 0;
    int e;
    switch (mode) {
        case D2D_MODE_HARD: e = launch_method<D2D_MODE_HARD>(p, grid_role, method, Z, valid_out, stream); break;
        case D2D_MODE_HARD_SIGMOID:
            e = launch_method<D2D_MODE_HARD_SIGMOID>(p, grid_role, method, Z, valid_out, stream);
            break;
        case D2D_MODE_SIGMOID: e = launch_method<D2D_MODE_SIGMOID>(p, grid_role, method, Z, valid_out, stream); break;
        default: return (int)cudaErrorInvalidValue;
    }
    if (launches) *launches += 1;
    return e;
}

}  // namespace d2d
