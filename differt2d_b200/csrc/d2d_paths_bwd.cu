// differt2d_b200 — reverse mode of the path materialisation kernel (gradients of a generic `fun`).
//
// The reference differentiates `acc = sum_c valid_c * fun(tx, rx, path_c, objects_c)` for ANY python `fun`
// (scene.py:1892-1925: jax.grad / value_and_grad over facc).  For the two fused functions the backward kernel does it
// in one launch; for an arbitrary `fun` the escape hatch materialises the paths (d2d_paths), the host framework
// evaluates `fun` on the batched vertices AND differentiates it (torch autograd), which yields for every emitted record
//     valid_bar = d acc / d valid   (= fun)          xys_bar = d acc / d xys   (= valid * d fun / d xys)
// and this kernel pulls those two cotangents back through the path construction and the validity logic to the grid
// points, the fixed points, the object vertices, the RIS angles and alpha — thread = record, ImagePath, the same
// tracked re-trace and reverse sweep as power_bwd_kernel (d2d_image_bwd.cuh).
#include "d2d_driver.cuh"
#include "d2d_image_bwd.cuh"
#include "d2d_launch.h"

namespace d2d {

struct RecordsIn {
    const int32_t* fixed;          // [n]
    const long long* grid;         // [n]
    const long long* candidate;    // [n] column in the problem's candidate list
    const float* valid_bar;        // [n]
    const float* xys_bar;          // [n, D2D_MAX_ORDER + 2, 2]
    long long n;
};

template <int MODE, int K, bool TXGRID>
__device__ __forceinline__ void record_vjp(const SceneTab& T, const KParams& p, const float alpha, const RecordsIn& in,
                                           const long long rec, const long long idx, const BwdOut& out, float* s_obj,
                                           float* s_phi, float& alpha_acc) {
    constexpr int KK = K > 0 ? K : 1;
    const int t = in.fixed[rec];
    const long long r = in.grid[rec];
    const float2 fx = reinterpret_cast<const float2*>(p.fixed)[t];
    const float2 g = reinterpret_cast<const float2*>(p.grid)[r];
    const float2 tx = TXGRID ? g : fx, rx = TXGRID ? fx : g;
    // candidate index -> objects (lexicographic, no equal neighbours; as test_candidate_inline)
    Cand<K> cd;
    cd.c[0] = 0;
    if constexpr (K > 0) {
        const int m = T.n_allowed;
        long long rem = idx;
        int dig[KK];
#pragma unroll
        for (int i = K - 1; i >= 1; --i) {
            dig[i] = (int)(rem % (m - 1));
            rem /= (m - 1);
        }
        dig[0] = (int)rem;
        int prev = -1;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            int pos = dig[i];
            if (i > 0 && pos >= prev) ++pos;
            cd.c[i] = T.allowed[pos];
            prev = pos;
        }
    }
    ImageTrace<K> tr;
    const float2 apex = image_apex<K>(T, cd, tx);
    if (!trace_image_tracked<MODE, K>(T, p, alpha, cd, tx, rx, apex, tr, nullptr)) return;  // validity exactly 0: constant
    float2 Xb[K + 2];
#pragma unroll
    for (int i = 0; i < K + 2; ++i)
        Xb[i] = reinterpret_cast<const float2*>(in.xys_bar)[rec * (D2D_MAX_ORDER + 2) + i];
    float2 txb, rxb;
    float ab;
    ObjAdj oa[KK];
    int occ_j;
    float4 occ_bar = make_float4(0.f, 0.f, 0.f, 0.f);
    image_reverse_general<MODE, K>(T, p, alpha, cd, tr, in.valid_bar[rec], Xb, txb, rxb, ab, oa, occ_j, occ_bar);
    const float2 gb = TXGRID ? txb : rxb, fb = TXGRID ? rxb : txb;
    if (out.grid_bar) {
        atomicAdd(&out.grid_bar[2 * ((long long)t * p.R + r) + 0], gb.x);
        atomicAdd(&out.grid_bar[2 * ((long long)t * p.R + r) + 1], gb.y);
    }
    if (out.fixed_bar) {
        atomicAdd(&out.fixed_bar[2 * t + 0], fb.x);
        atomicAdd(&out.fixed_bar[2 * t + 1], fb.y);
    }
    alpha_acc += ab;
    if (s_obj) {
        if (occ_j >= 0) {
            atomicAdd(&s_obj[4 * occ_j + 0], occ_bar.x); atomicAdd(&s_obj[4 * occ_j + 1], occ_bar.y);
            atomicAdd(&s_obj[4 * occ_j + 2], occ_bar.z); atomicAdd(&s_obj[4 * occ_j + 3], occ_bar.w);
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float4 v = oa[i].to_vertices(T.w0[cd.c[i]], T.w1[cd.c[i]]);
            atomicAdd(&s_obj[4 * cd.c[i] + 0], v.x); atomicAdd(&s_obj[4 * cd.c[i] + 1], v.y);
            atomicAdd(&s_obj[4 * cd.c[i] + 2], v.z); atomicAdd(&s_obj[4 * cd.c[i] + 3], v.w);
            if (T.kind[cd.c[i]] == D2D_KIND_RIS) atomicAdd(&s_phi[cd.c[i]], oa[i].phi);
        }
    }
}

template <int MODE, bool TXGRID>
__global__ void __launch_bounds__(kBlock) paths_vjp_kernel(const KParams p, const RecordsIn in, const BwdOut out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_count;
    __shared__ float s_alpha[kBlock / 32];
    SceneTab T = carve_tab(smem, p.N);
    float* s_obj = reinterpret_cast<float*>(smem + ((scene_tab_bytes(p.N) + 15) / 16) * 16);
    float* s_phi = s_obj + 4 * p.N;
    for (int j = threadIdx.x; j < 5 * p.N; j += blockDim.x) s_obj[j] = 0.f;
    build_tab(T, p, &s_count);
    const float alpha = p.alpha_dev ? *p.alpha_dev : p.alpha;
    float alpha_acc = 0.f;
    if (MODE == D2D_MODE_HARD || alpha > 0.0f) {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long rec = (long long)blockIdx.x * blockDim.x + threadIdx.x; rec < in.n; rec += stride) {
            long long col = in.candidate[rec];
            int k = p.min_order;
            for (; k < p.max_order; ++k) {  // order of this column
                const long long c = order_count(k, T.n_allowed);
                if (col < c) break;
                col -= c;
            }
            switch (k) {
                case 0: record_vjp<MODE, 0, TXGRID>(T, p, alpha, in, rec, col, out, s_obj, s_phi, alpha_acc); break;
                case 1: record_vjp<MODE, 1, TXGRID>(T, p, alpha, in, rec, col, out, s_obj, s_phi, alpha_acc); break;
                case 2: record_vjp<MODE, 2, TXGRID>(T, p, alpha, in, rec, col, out, s_obj, s_phi, alpha_acc); break;
                case 3: record_vjp<MODE, 3, TXGRID>(T, p, alpha, in, rec, col, out, s_obj, s_phi, alpha_acc); break;
                case 4: record_vjp<MODE, 4, TXGRID>(T, p, alpha, in, rec, col, out, s_obj, s_phi, alpha_acc); break;
                default: break;
            }
        }
    }
    // CTA reduction of alpha_bar, then the CTA's object cotangents
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) alpha_acc += __shfl_xor_sync(0xffffffffu, alpha_acc, o);
    if ((threadIdx.x & 31) == 0) s_alpha[threadIdx.x >> 5] = alpha_acc;
    __syncthreads();
    if (threadIdx.x == 0 && out.alpha_bar) {
        float s = 0.f;
        for (int w = 0; w < kBlock / 32; ++w) s += s_alpha[w];
        if (s != 0.f) atomicAdd(out.alpha_bar, s);
    }
    for (int j = threadIdx.x; j < 4 * p.N; j += blockDim.x)
        if (out.objects_bar && s_obj[j] != 0.f) atomicAdd(&out.objects_bar[j], s_obj[j]);
    for (int j = threadIdx.x; j < p.N; j += blockDim.x)
        if (out.phis_bar && s_phi[j] != 0.f) atomicAdd(&out.phis_bar[j], s_phi[j]);
}

template <int MODE>
static int launch_paths_vjp_mode(const KParams& p, int grid_role, const RecordsIn& in, const BwdOut& out, cudaStream_t stream) {
    const size_t smem = ((scene_tab_bytes(p.N) + 15) / 16) * 16 + (size_t)5 * p.N * sizeof(float);
    long long blocks = (in.n + kBlock - 1) / kBlock;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    cudaError_t e = cudaSuccess;
    if (grid_role == D2D_GRID_TRANSMITTERS) {
        auto kern = paths_vjp_kernel<MODE, true>;
        if (smem > 40 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<(unsigned)blocks, kBlock, smem, stream>>>(p, in, out);
    } else {
        auto kern = paths_vjp_kernel<MODE, false>;
        if (smem > 40 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<(unsigned)blocks, kBlock, smem, stream>>>(p, in, out);
    }
    return (int)cudaGetLastError();
}

int launch_paths_vjp(const KParams& p, int mode, int grid_role, long long n, const int32_t* rec_fixed,
                     const long long* rec_grid, const long long* rec_candidate, const float* valid_bar,
                     const float* xys_bar, const BwdOut& out, cudaStream_t stream, long long* launches) {
    cudaError_t e;
    if (out.grid_bar && (e = cudaMemsetAsync(out.grid_bar, 0, sizeof(float) * 2 * (size_t)p.T * p.R, stream)) != cudaSuccess) return (int)e;
    if (out.objects_bar && (e = cudaMemsetAsync(out.objects_bar, 0, sizeof(float) * 4 * p.N, stream)) != cudaSuccess) return (int)e;
    if (out.phis_bar && (e = cudaMemsetAsync(out.phis_bar, 0, sizeof(float) * p.N, stream)) != cudaSuccess) return (int)e;
    if (out.fixed_bar && (e = cudaMemsetAsync(out.fixed_bar, 0, sizeof(float) * 2 * p.T, stream)) != cudaSuccess) return (int)e;
    if (out.alpha_bar && (e = cudaMemsetAsync(out.alpha_bar, 0, sizeof(float), stream)) != cudaSuccess) return (int)e;
    if (n <= 0) return 0;
    const RecordsIn in{rec_fixed, rec_grid, rec_candidate, valid_bar, xys_bar, n};
    int rc;
    switch (mode) {
        case D2D_MODE_HARD: rc = launch_paths_vjp_mode<D2D_MODE_HARD>(p, grid_role, in, out, stream); break;
        case D2D_MODE_HARD_SIGMOID: rc = launch_paths_vjp_mode<D2D_MODE_HARD_SIGMOID>(p, grid_role, in, out, stream); break;
        case D2D_MODE_SIGMOID: rc = launch_paths_vjp_mode<D2D_MODE_SIGMOID>(p, grid_role, in, out, stream); break;
        default: return (int)cudaErrorInvalidValue;
    }
    if (launches) *launches += 1;
    return rc;
}

}  // namespace d2d
