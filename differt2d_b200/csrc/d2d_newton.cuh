// differt2d_b200 — D2D_OPT_NEWTON: a NON-PARITY fast mode for FermatPath / MinPath.
//
// The reference minimises the path loss with `steps` iterations of a first-order optax optimiser (optimize.py:44-97;
// 100 Adam steps by default, 1000 in examples/plot_ris_power_map.py:72).  The losses are smooth functions of at most
// D2D_MAX_ORDER unknowns, so a damped Newton iteration reaches the same stationary point in a handful of steps:
//     (H + lam diag(H)) delta = -g ,   theta <- theta + delta   when the loss does not increase, else lam grows
// with g and the Hessian H obtained from the hand-written reverse sweep of the loss run in dual-number arithmetic
// (d2d_solver_adj.cuh: one dual evaluation per unknown gives one Hessian column and the gradient).
//
// Reverse mode: instead of unrolling the iterations, the converged point is differentiated IMPLICITLY.  theta*(q)
// solves g(theta*, q) = 0 (q = transmitter, receiver, object parameters), so d theta* = -H^-1 (dg/dq) dq, and for a
// cotangent theta_bar the parameters receive  lam^T dg/dq  with  H lam = -theta_bar : one K x K solve and ONE dual
// evaluation of the loss gradient along lam, whatever the number of iterations.  (MinPath's `loss` output is
// differentiated at theta*: its theta-part is g = 0.)
#pragma once

#include "d2d_solver.cuh"
#include "d2d_solver_adj.cuh"

namespace d2d {

// Solves A x = b for the K x K symmetric system (no pivoting: A is a damped Hessian), in registers.
// Returns false when a pivot is not safely positive (the caller then increases the damping).
template <int K>
__device__ __forceinline__ bool solve_small(float (&A)[K][K], float (&b)[K], float (&x)[K]) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const float piv = A[c][c];
        ok = ok && (piv > 1e-30f);
        const float inv = 1.0f / piv;
#pragma unroll
        for (int r = c + 1; r < K; ++r) {
            const float f = A[r][c] * inv;
#pragma unroll
            for (int k = c; k < K; ++k) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
#pragma unroll
    for (int r = K - 1; r >= 0; --r) {
        float acc = b[r];
#pragma unroll
        for (int k = r + 1; k < K; ++k) acc -= A[r][k] * x[k];
        x[r] = acc / A[r][r];
    }
    return ok;
}

// gradient g and Hessian H of the loss w.r.t. the unknowns at theta (rows / columns of objects without an unknown —
// vertices — are the identity / zero)
template <int METHOD, int K>
__device__ __forceinline__ void loss_grad_hessian(const SceneTab& T, const Cand<K>& cd, const float (&th)[K > 0 ? K : 1],
                                                  const float2 tx, const float2 rx, float (&g)[K > 0 ? K : 1],
                                                  float (&H)[K > 0 ? K : 1][K > 0 ? K : 1]) {
    constexpr int KK = K > 0 ? K : 1;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const bool active = T.kind[cd.c[j]] != D2D_KIND_VERTEX;
        Dual thd[KK], thb[KK];
#pragma unroll
        for (int i = 0; i < KK; ++i) thd[i] = mk(th[i], (i == j && active) ? 1.0f : 0.0f);
        V2<Dual> txb, rxb;
        ObjAdjS<Dual> ob[KK];
        loss_grad_full<METHOD, K, Dual>(T, cd, thd, tx, rx, 1.0f, thb, txb, rxb, ob);
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const bool ai = T.kind[cd.c[i]] != D2D_KIND_VERTEX;
            H[i][j] = (ai && active) ? thb[i].d : (i == j ? 1.0f : 0.0f);
            if (i == j) g[i] = ai ? thb[i].v : 0.0f;
        }
    }
}

template <int METHOD, int K>
__device__ __forceinline__ float loss_only(const SceneTab& T, const Cand<K>& cd, const float (&th)[K > 0 ? K : 1],
                                           const float2 tx, const float2 rx) {
    float2 X[K + 2], G[K + 2];
    X[0] = tx;
    X[K + 1] = rx;
    place_points<K>(T, cd, th, X);
    return solver_loss_grad<METHOD, K>(T, cd, X, G);
}

template <int METHOD, int K>
__device__ float newton_solve(const SceneTab& T, const KParams& p, const Cand<K>& cd, const float2 tx, const float2 rx,
                              float (&th)[K > 0 ? K : 1]) {
    constexpr int KK = K > 0 ? K : 1;
    float L = loss_only<METHOD, K>(T, cd, th, tx, rx);
    float lam = 1e-3f;
    for (int it = 0; it < p.steps; ++it) {
        float g[KK], H[KK][KK];
        loss_grad_hessian<METHOD, K>(T, cd, th, tx, rx, g, H);
        float A[KK][KK], b[KK], dlt[KK];
#pragma unroll
        for (int i = 0; i < KK; ++i) {
#pragma unroll
            for (int j = 0; j < KK; ++j) A[i][j] = H[i][j];
            // modified Newton + Levenberg-Marquardt: |H_ii| keeps the pivots positive where MinPath's loss is concave,
            // the damping is relative to the curvature and bounded away from zero
            A[i][i] = fabsf(H[i][i]) * (1.0f + lam) + 1e-6f * (1.0f + lam);
            b[i] = -g[i];
            dlt[i] = 0.f;
        }
        const bool ok = solve_small<KK>(A, b, dlt);
        // trust region: the unknowns are wall parameters of order one, and both losses are built from NORMS, whose
        // quadratic model is poor far from the solution (the Hessian vanishes like 1 / distance): a step is never longer
        // than half a wall
        float big = 0.f;
#pragma unroll
        for (int i = 0; i < KK; ++i) big = fmaxf(big, fabsf(dlt[i]));
        const float shrink = big > 0.5f ? 0.5f / big : 1.0f;
        float cand[KK];
#pragma unroll
        for (int i = 0; i < KK; ++i) cand[i] = th[i] + shrink * dlt[i];
        const float Lc = (ok && big == big) ? loss_only<METHOD, K>(T, cd, cand, tx, rx) : CUDART_INF_F;
        if (Lc <= L + 1e-7f * fabsf(L)) {  // (NaN compares false: rejected; equal up to rounding: accepted)
#pragma unroll
            for (int i = 0; i < KK; ++i) th[i] = cand[i];
            L = Lc;
            lam = fmaxf(lam * 0.25f, 1e-7f);
        } else {
            lam = fminf(lam * 8.0f, 1e6f);
        }
    }
    return L;
}

// Implicit reverse mode at the converged point: adds to tx_bar / rx_bar / oa the cotangents that theta_bar (cotangent of
// the returned unknowns) and loss_bar (cotangent of the returned loss, MinPath) induce on the parameters.
template <int METHOD, int K>
__device__ __forceinline__ void newton_reverse(const SceneTab& T, const Cand<K>& cd, const float (&th)[K > 0 ? K : 1],
                                               const float2 tx, const float2 rx, const float (&th_bar)[K > 0 ? K : 1],
                                               const float loss_bar, float2& tx_bar, float2& rx_bar,
                                               ObjAdj (&oa)[K > 0 ? K : 1]) {
    constexpr int KK = K > 0 ? K : 1;
    if (METHOD == D2D_METHOD_MINPATH && loss_bar != 0.f) {  // d loss / d q at fixed theta (the theta-part is g = 0)
        float thb[KK];
        V2<float> txb, rxb;
        ObjAdjS<float> ob[KK];
        loss_grad_full<METHOD, K, float>(T, cd, th, tx, rx, loss_bar, thb, txb, rxb, ob);
        tx_bar.x += txb.x; tx_bar.y += txb.y;
        rx_bar.x += rxb.x; rx_bar.y += rxb.y;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            oa[i].p1.x += ob[i].p1.x; oa[i].p1.y += ob[i].p1.y;
            oa[i].t.x += ob[i].t.x; oa[i].t.y += ob[i].t.y;
            oa[i].n.x += ob[i].n.x; oa[i].n.y += ob[i].n.y;
            oa[i].phi += ob[i].phi;
        }
    }
    bool any = false;
#pragma unroll
    for (int i = 0; i < K; ++i) any = any || th_bar[i] != 0.f;
    if (!any) return;
    float g[KK], H[KK][KK], b[KK], lam[KK];
    loss_grad_hessian<METHOD, K>(T, cd, th, tx, rx, g, H);
#pragma unroll
    for (int i = 0; i < KK; ++i) {
        H[i][i] += 1e-7f * (fabsf(H[i][i]) + 1e-6f);  // (keeps a flat direction from blowing the solve up)
        b[i] = i < K ? -th_bar[i] : 0.f;
        lam[i] = 0.f;
    }
    if (!solve_small<KK>(H, b, lam)) return;  // singular curvature: no well-defined implicit derivative
    Dual thd[KK], thb[KK];
#pragma unroll
    for (int i = 0; i < KK; ++i) thd[i] = mk(th[i], (i < K && T.kind[cd.c[i < K ? i : 0]] != D2D_KIND_VERTEX) ? lam[i] : 0.f);
    V2<Dual> txb, rxb;
    ObjAdjS<Dual> ob[KK];
    loss_grad_full<METHOD, K, Dual>(T, cd, thd, tx, rx, 1.0f, thb, txb, rxb, ob);
    tx_bar.x += txb.x.d; tx_bar.y += txb.y.d;
    rx_bar.x += rxb.x.d; rx_bar.y += rxb.y.d;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        oa[i].p1.x += ob[i].p1.x.d; oa[i].p1.y += ob[i].p1.y.d;
        oa[i].t.x += ob[i].t.x.d; oa[i].t.y += ob[i].t.y.d;
        oa[i].n.x += ob[i].n.x.d; oa[i].n.y += ob[i].n.y.d;
        oa[i].phi += ob[i].phi.d;
    }
}

}  // namespace d2d
