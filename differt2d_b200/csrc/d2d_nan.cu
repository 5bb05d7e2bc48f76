// differt2d_b200 — D2D_GRAD_NAN_PARITY: where the reference's reverse mode yields NaN.
//
// The backward kernels compute the CLEAN gradient (masked branches are constants).  jax.grad over the reference's
// literal op graph instead returns NaN wherever a masked branch divides by zero, because a ZERO cotangent still
// flows into the branch that was not taken (0 * inf).  Three places on the ImagePath route do that:
//
//   (a) geometry.py:1105   inc = where(un == 0, 0, vn * u / un)          single `where`: NaN into u, vn, un whenever
//                          un == 0 — i.e. into the grid point, the fixed point and every object of the candidate
//                          (all of them are upstream of u = point - image_i);
//   (b) geometry.py:227-230 normalize(v) = v / where(|v| == 0, 1, |v|)   d|v| at v = 0 is 0/0: the normal of a
//                          zero-length wall (always together with (a): n = 0 makes un = 0), and — smooth logic only,
//                          the loss of hard logic ends in a comparison — a zero-length path segment inside
//                          evaluate_cartesian (:641-650): same upstream set as (a);
//   (c) geometry.py:163-171 t = where(d == 0, inf, num / den_safe)       the division is guarded, but smooth logic
//                          evaluates act(alpha * (t + tol)) at t = inf: d/d(alpha) = act'(inf) * inf = 0 * inf —
//                          NaN into alpha only.
//
// Validity does not matter: acc + valid * fun sends the cotangent valid * g = 0 into the path of an invalid
// candidate, and 0 * inf is NaN all the same.  So the events have to be looked for on EVERY (fixed point, grid point,
// candidate), without any of the culls: this pass is a diagnostic mode (about the cost of an unculled forward launch),
// run after the clean backward kernel, overwriting the affected outputs with NaN (tests/test_gpu_parity.py:
// test_nan_parity_gradient_mode compares the NaN pattern with autograd over the literal graph).
#include "d2d_launch.h"
#include "d2d_trace.cuh"

namespace d2d {

namespace {

struct PoisonAcc {
    bool path;    // events (a) / (b) on some candidate of the current (fixed point, grid point)
    bool alpha;   // event (c) anywhere
};

template <int K, bool TXGRID>
__device__ __forceinline__ void poison_order(const SceneTab& T, const KParams& p, const bool smooth, const bool want_alpha,
                                             const float2 fx, const float2 g, uint32_t* s_obj, PoisonAcc& acc) {
    constexpr int KK = K > 0 ? K : 1;
    const float2 tx = TXGRID ? g : fx;
    const float2 rx = TXGRID ? fx : g;
    const int m = T.n_allowed;
    Odometer<K> od;
    if (!od.first(m)) return;
    do {
        int c[KK];
#pragma unroll
        for (int i = 0; i < K; ++i) c[i] = T.allowed[od.pos[i]];
        float2 X[K + 2];
        X[0] = tx;
        X[K + 1] = rx;
        bool ev = false;
        if constexpr (K > 0) {
            float2 I[K + 1];
            I[0] = tx;
#pragma unroll
            for (int i = 0; i < K; ++i) I[i + 1] = mirror(I[i], T.w0[c[i]], T.w1[c[i]]);
            float2 q = rx;
#pragma unroll
            for (int i = K - 1; i >= 0; --i) {
                const float4 w0 = T.w0[c[i]], w1 = T.w1[c[i]];
                const float ux = q.x - I[i + 1].x, uy = q.y - I[i + 1].y;
                const float un = ux * w1.x + uy * w1.y;  // geometry.py:1102 (same operations as back_project)
                if (un == 0.0f && T.kind[c[i]] != D2D_KIND_VERTEX) ev = true;  // (a)
                q = back_project(q, I[i + 1], w0, w1);
                X[i + 1] = q;
            }
            if (smooth) {  // (b): every segment of the path is normalised by evaluate_cartesian
#pragma unroll
                for (int i = 0; i <= K; ++i)
                    if (X[i + 1].x - X[i].x == 0.0f && X[i + 1].y - X[i].y == 0.0f) ev = true;
            }
            if (ev) {
                acc.path = true;
#pragma unroll
                for (int i = 0; i < K; ++i) atomicOr(&s_obj[c[i] >> 5], 1u << (c[i] & 31));
            }
        }
        if (smooth && want_alpha && !acc.alpha) {  // (c): a segment parallel to a tested object, d == 0 exactly
            for (int j = 0; j < p.N && !acc.alpha; ++j) {
                if (T.kind[j] == D2D_KIND_VERTEX) continue;  // never tested (geometry.py:405-414)
                const float4 w = T.w2[j];
#pragma unroll
                for (int i = 0; i <= K; ++i) {
                    const int sa = (i > 0) ? c[i > 0 ? i - 1 : 0] : -1;
                    const int sb = (i < K) ? c[i < K ? i : 0] : -1;
                    if (j == sa || j == sb) continue;  // geometry.py:887-904
                    const float Bx = X[i].x - X[i + 1].x, By = X[i].y - X[i + 1].y;
                    const float d = w.w * Bx - w.z * By;  // geometry.py:159
                    if (d == 0.0f) acc.alpha = true;
                }
            }
        }
    } while (od.next(m));
}

template <bool TXGRID>
__global__ void __launch_bounds__(128) nan_poison_kernel(const KParams p, const int smooth, const BwdOut out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_count;
    __shared__ uint32_t s_obj[D2D_MAX_OBJECTS / 32];
    SceneTab T = carve_tab(smem, p.N);
    for (int j = threadIdx.x; j < D2D_MAX_OBJECTS / 32; j += blockDim.x) s_obj[j] = 0u;
    build_tab(T, p, &s_count);
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = r < p.R;
    const float2 g = active ? reinterpret_cast<const float2*>(p.grid)[r] : make_float2(0.f, 0.f);
    const float nan = CUDART_NAN_F;
    const bool want_alpha = out.alpha_bar != nullptr;
    PoisonAcc acc;
    acc.alpha = false;
    bool any_t = false;
    for (int t = 0; t < p.T; ++t) {
        const float2 fx = reinterpret_cast<const float2*>(p.fixed)[t];
        acc.path = false;
        if (active) {
            for (int k = p.min_order; k <= p.max_order; ++k) {
                switch (k) {
                    case 0: poison_order<0, TXGRID>(T, p, smooth != 0, want_alpha, fx, g, s_obj, acc); break;
                    case 1: poison_order<1, TXGRID>(T, p, smooth != 0, want_alpha, fx, g, s_obj, acc); break;
                    case 2: poison_order<2, TXGRID>(T, p, smooth != 0, want_alpha, fx, g, s_obj, acc); break;
                    case 3: poison_order<3, TXGRID>(T, p, smooth != 0, want_alpha, fx, g, s_obj, acc); break;
                    case 4: poison_order<4, TXGRID>(T, p, smooth != 0, want_alpha, fx, g, s_obj, acc); break;
                    default: break;
                }
            }
        }
        any_t = any_t || acc.path;
        if (acc.path && out.grid_bar && !p.reduce_all) {
            out.grid_bar[2 * ((long long)t * p.R + r) + 0] = nan;
            out.grid_bar[2 * ((long long)t * p.R + r) + 1] = nan;
        }
        if (__syncthreads_or(acc.path) && threadIdx.x == 0 && out.fixed_bar) {
            out.fixed_bar[2 * t + 0] = nan;
            out.fixed_bar[2 * t + 1] = nan;
        }
    }
    if (any_t && out.grid_bar && p.reduce_all) {  // the reduced map's cotangent sums over the fixed points
        out.grid_bar[2 * r + 0] = nan;
        out.grid_bar[2 * r + 1] = nan;
    }
    if (__syncthreads_or(acc.alpha) && threadIdx.x == 0 && out.alpha_bar) *out.alpha_bar = nan;
    if (out.objects_bar) {
        for (int j = threadIdx.x; j < p.N; j += blockDim.x) {
            if ((s_obj[j >> 5] >> (j & 31)) & 1u) {
#pragma unroll
                for (int q = 0; q < 4; ++q) out.objects_bar[4 * j + q] = nan;
            }
        }
    }
}

}  // namespace

int launch_nan_poison(const KParams& p, int mode, int grid_role, const BwdOut& out, cudaStream_t stream,
                      long long* launches) {
    if (p.R <= 0 || p.T <= 0) return 0;
    const size_t smem = scene_tab_bytes(p.N);
    const unsigned nblk = (unsigned)((p.R + 127) / 128);
    const int smooth = mode != D2D_MODE_HARD;
    cudaError_t e;
    if (grid_role == D2D_GRID_TRANSMITTERS) {
        auto kern = nan_poison_kernel<true>;
        if (smem > 32 * 1024 &&
            (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return (int)e;
        kern<<<nblk, 128, smem, stream>>>(p, smooth, out);
    } else {
        auto kern = nan_poison_kernel<false>;
        if (smem > 32 * 1024 &&
            (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return (int)e;
        kern<<<nblk, 128, smem, stream>>>(p, smooth, out);
    }
    if (launches) *launches += 1;
    return (int)cudaGetLastError();
}

}  // namespace d2d
