// differt2d_b200 — path construction dispatch and the in-register Adam solver of FermatPath / MinPath.
//
// FermatPath.from_tx_objects_rx  geometry.py:1121-1204 : minimise path_length over the parametric
//                                coordinates of the interacting objects;
// MinPath.from_tx_objects_rx     geometry.py:1211-1288 : minimise the sum of interaction residuals;
// both through optimize.minimize (optimize.py:44-97): `steps` iterations of optax.adam(lr) started
// at x0 (optimize.py:132), whole state (theta, mu, nu : <= 3*K floats) in registers.
#pragma once

#include "d2d_adjoint.cuh"
#include "d2d_trace.cuh"

namespace d2d {

// parametric_to_cartesian — geometry.py:976-1010, Wall :581-587, Vertex :383-389
template <int K>
__device__ __forceinline__ void place_points(const SceneTab& T, const Cand<K>& cd, const float (&th)[K > 0 ? K : 1],
                                             float2 (&X)[K + 2]) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float4 w0 = T.w0[cd.c[i]];
        if (T.kind[cd.c[i]] == D2D_KIND_VERTEX) X[i + 1] = make_float2(w0.x, w0.y);
        else X[i + 1] = make_float2(w0.x + th[i] * w0.z, w0.y + th[i] * w0.w);
    }
}

// loss_fun value and gradient w.r.t. the interior points X[1..K]
template <int METHOD, int K>
__device__ __forceinline__ float solver_loss_grad(const SceneTab& T, const Cand<K>& cd, const float2 (&X)[K + 2],
                                                  float2 (&G)[K + 2]) {
#pragma unroll
    for (int i = 0; i < K + 2; ++i) G[i] = make_float2(0.f, 0.f);
    if (METHOD == D2D_METHOD_FERMAT) {
        float total = 0.0f;
#pragma unroll
        for (int i = 0; i <= K; ++i) {  // path_length geometry.py:176-203
            const float dx = (X[i + 1].x - X[i].x) + kEps32;
            const float dy = (X[i + 1].y - X[i].y) + kEps32;
            const float len = sqrtf(dx * dx + dy * dy);
            total = (i == 0) ? len : total + len;
            const float inv = 1.0f / len;
            G[i + 1].x += dx * inv; G[i + 1].y += dy * inv;
            G[i].x -= dx * inv; G[i].y -= dy * inv;
        }
        return total;
    }
    float loss = 0.0f;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int j = cd.c[i];
        loss = loss + residual_value_and_grad(T.kind[j], X[i], X[i + 1], X[i + 2], T.w1[j], T.sc[j], G[i], G[i + 1],
                                              G[i + 2]);
    }
    return loss;
}

// D2D_OPT_NEWTON (d2d_newton.cuh, included by every kernel translation unit after this header): damped Newton
// iterations from th (in / out); returns the loss at the returned point.
template <int METHOD, int K>
__device__ float newton_solve(const SceneTab& T, const KParams& p, const Cand<K>& cd, const float2 tx, const float2 rx,
                              float (&th)[K > 0 ? K : 1]);

// One optimiser update of unknown i for the loss gradient g (optimize.py:87-93 with the optimizer of p.opt):
// optax.adam: mu, nu moments, bias correction with b1^s / b2^s (bc1, bc2); optax.sgd: mu is the momentum trace.
__device__ __forceinline__ void optimizer_update(const KParams& p, const float g, const float bc1, const float bc2,
                                                 float& th, float& mu, float& nu) {
    if (p.opt == D2D_OPT_SGD) {
        mu = g + p.b1 * mu;  // optax.trace(decay = momentum)
        th = th + (-p.lr) * mu;
        return;
    }
    mu = (1.0f - p.b1) * g + p.b1 * mu;
    nu = (1.0f - p.b2) * (g * g) + p.b2 * nu;
    const float mh = mu / bc1, nh = nu / bc2;
    th = th + (-p.lr) * (mh / (sqrtf(nh) + p.opt_eps));
}

template <int METHOD, int K>
__device__ __forceinline__ void construct_path(const SceneTab& T, const KParams& p, const Cand<K>& cd,
                                               const float2 tx, const float2 rx, const long long col,
                                               float2 (&X)[K + 2], float& loss) {
    // `loss` is only produced here for MinPath (the solver's own value, optimize.py:96-97); for the
    // other methods it is path_loss(X), evaluated lazily by validity<.., LAZY_LOSS = true>.
    loss = 0.0f;
    if (METHOD == D2D_METHOD_IMAGE || K == 0) {
        image_path<K>(T, cd, tx, rx, X);
        return;
    }
    constexpr int KK = K > 0 ? K : 1;
    float th[KK], mu[KK], nu[KK];
    float best_th[KK];
    float best = 0.0f, last = 0.0f;
    X[0] = tx;
    X[K + 1] = rx;
    const float b1 = p.b1, b2 = p.b2;  // optax.adam defaults 0.9 / 0.999 / 1e-8 unless the caller chose otherwise
    // minimize_many_random_uniform (optimize.py:142-182): `many` independent scans, keep argmin of the final losses
    for (int r = 0; r < p.many; ++r) {
        int u = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            mu[i] = 0.f;
            nu[i] = 0.f;
            th[i] = 0.f;
            if (T.kind[cd.c[i]] != D2D_KIND_VERTEX) {
                th[i] = p.x0 ? p.x0[(col * p.many + r) * p.max_order + u] : 0.5f;
                ++u;
            }
        }
        if (p.opt == D2D_OPT_NEWTON) {
            last = newton_solve<METHOD, K>(T, p, cd, tx, rx, th);
            break;  // (no restarts in this mode)
        }
        float b1p = 1.0f, b2p = 1.0f;
        for (int s = 0; s < p.steps; ++s) {
            place_points<K>(T, cd, th, X);
            float2 G[K + 2];
            last = solver_loss_grad<METHOD, K>(T, cd, X, G);
            b1p *= b1;
            b2p *= b2;
            const float bc1 = 1.0f - b1p, bc2 = 1.0f - b2p;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                if (T.kind[cd.c[i]] == D2D_KIND_VERTEX) continue;
                const float4 w0 = T.w0[cd.c[i]];
                const float g = G[i + 1].x * w0.z + G[i + 1].y * w0.w;
                optimizer_update(p, g, bc1, bc2, th[i], mu[i], nu[i]);
            }
        }
        if (p.many == 1) break;
        // jnp.argmin: first minimum; a NaN loss wins (NaN propagates through argmin)
        if (r == 0 || (last < best && best == best) || (last != last && best == best)) {
            best = last;
#pragma unroll
            for (int i = 0; i < K; ++i) best_th[i] = th[i];
        }
    }
    if (p.many > 1) {
        last = best;
#pragma unroll
        for (int i = 0; i < K; ++i) th[i] = best_th[i];
    }
    place_points<K>(T, cd, th, X);
    if (METHOD == D2D_METHOD_FERMAT) {
        // geometry.py:1202-1204: loss = path_loss(xys), left to validity()
    } else {
        if (p.steps <= 0 && p.opt != D2D_OPT_NEWTON) {
            float2 G[K + 2];
            last = solver_loss_grad<METHOD, K>(T, cd, X, G);
        }
        loss = last;  // optimize.py:96-97 losses[-1]
    }
}

}  // namespace d2d
