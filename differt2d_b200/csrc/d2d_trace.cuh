// differt2d_b200 — one (transmitter, candidate, receiver) path: construction, validity, value.
// Forward-only pieces; the reverse sweep lives in d2d_adjoint.cuh.
#pragma once

#include "d2d_device.cuh"

namespace d2d {

template <int K>
struct Cand {
    int c[K > 0 ? K : 1];
};

// ImagePath.from_tx_objects_rx — geometry.py:1017-1114.  Fills X[0..K+1].
template <int K>
__device__ __forceinline__ void image_path(const SceneTab& T, const Cand<K>& cd, const float2 tx,
                                           const float2 rx, float2 (&X)[K + 2]) {
    X[0] = tx;
    X[K + 1] = rx;
    if (K == 0) return;
    float2 I[K + 1];
    I[0] = tx;
#pragma unroll
    for (int i = 0; i < K; ++i) I[i + 1] = mirror(I[i], T.w0[cd.c[i]], T.w1[cd.c[i]]);
    float2 q = rx;
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        q = back_project(q, I[i + 1], T.w0[cd.c[i]], T.w1[cd.c[i]]);
        X[i + 1] = q;
    }
}

// Path.on_objects in pre-activation form — geometry.py:821-854, 589-621.  +inf when no object
// contributes a comparison (order 0, or only vertices): `true_value`.
template <int K>
__device__ __forceinline__ float on_objects_x(const SceneTab& T, const Cand<K>& cd, const float2 (&X)[K + 2]) {
    float onx = CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int j = cd.c[i];
        if (T.kind[j] == D2D_KIND_VERTEX) continue;
        const float s = to_parametric(X[i + 1], T.w0[j], T.w1[j]);
        float x = fminf(s, 1.0f - s);  // ge(s, 0) -> s - 0 ; le(s, 1) -> 1 - s
        if (s != s) x = -CUDART_INF_F;
        onx = fminf(onx, x);
    }
    return onx;
}

// sum of interaction residuals — geometry.py:1077-1084
template <int K>
__device__ __forceinline__ float path_loss(const SceneTab& T, const Cand<K>& cd, const float2 (&X)[K + 2]) {
    float loss = 0.0f;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int j = cd.c[i];
        loss = loss + residual(T.kind[j], X[i], X[i + 1], X[i + 2], T.w1[j], T.sc[j]);
    }
    return loss;
}

// Occlusion of one segment by objects j in [j0, j1) — geometry.py:887-904 + :153-173.
// Fast path per test: canonical a, b, d; approximate parameters through one MUFU.RCP and two FMAs;
// the exact IEEE divisions only run when the test could raise the running maximum `interx`.
template <int MODE, bool TRACK>
__device__ __forceinline__ void occlude_range(const SceneTab& T, const int j0, const int j1, const float2 P,
                                              const float2 B, const float alpha, const float xz,
                                              float& interx, float& cthr, bool& alive, int& arg_j) {
#pragma unroll 4
    for (int j = j0; j < j1 && alive; ++j) {
        const float4 w = T.w2[j];
        const float Cx = w.x - P.x, Cy = w.y - P.y;
        const float a = B.y * Cx - B.x * Cy;
        const float b = w.z * Cy - w.w * Cx;
        const float d = w.w * B.x - w.z * B.y;
        const float r = rcp_approx(d);
        const float qa = fmaf(a, r, -0.5f);
        const float qb = fmaf(b, r, -0.5f);
        const float m = fmaxf(fabsf(qa), fabsf(qb));
        if (!(m >= cthr)) {
            const float hx = hit_exact(a, b, d);
            if (hx > interx) {
                interx = hx;
                if (TRACK) arg_j = j;
                cthr = filter_threshold(fmaxf(hx, xz));
                // the path is dead once `intersects` is exactly true / 1.0
                if (MODE == D2D_MODE_HARD) alive = !(hx >= 0.0f);
                else alive = !(act<MODE>(hx, alpha) == 1.0f);
            }
        }
    }
}

// Path.intersects_with_objects in pre-activation form.  Returns interx (-inf: `false_value`).
template <int MODE, int K, bool TRACK>
__device__ __forceinline__ float intersects_x(const SceneTab& T, const int N, const Cand<K>& cd,
                                              const float2 (&X)[K + 2], const float alpha, bool& alive,
                                              int& arg_seg, int& arg_j) {
    float interx = -CUDART_INF_F;
    const float xz = x_zero<MODE>(alpha);
    float cthr = filter_threshold(xz);
#pragma unroll
    for (int i = 0; i <= K; ++i) {
        const int sa = (i > 0) ? cd.c[i - 1] : -1;
        const int sb = (i < K) ? cd.c[i] : -1;
        const int lo = min(sa, sb), hi = max(sa, sb);  // lo may be -1; sa != sb unless both -1
        const float2 P = X[i];
        const float2 B = make_float2(X[i].x - X[i + 1].x, X[i].y - X[i + 1].y);
        const float before = interx;
        occlude_range<MODE, TRACK>(T, 0, lo < 0 ? 0 : lo, P, B, alpha, xz, interx, cthr, alive, arg_j);
        occlude_range<MODE, TRACK>(T, lo + 1, hi < 0 ? 0 : hi, P, B, alpha, xz, interx, cthr, alive, arg_j);
        occlude_range<MODE, TRACK>(T, hi + 1, N, P, B, alpha, xz, interx, cthr, alive, arg_j);
        if (TRACK && interx != before) arg_seg = i;
        if (!alive) break;
    }
    return interx;
}

// utils.received_power (utils.py:16-54) / path.length()**2
template <int K>
__device__ __forceinline__ float path_value(const KParams& p, const float2 (&X)[K + 2], float& r) {
    r = path_length<K + 2>(X);
    if (p.fun == D2D_FUN_RECEIVED_POWER) return p.rc_pow[K] / (p.h2 + r * r);
    return r * r;
}

// Path.is_valid (geometry.py:908-963) for an already constructed path.
// Returns the validity (0/1 in hard mode).  The three conjuncts commute, so the cheap ones run first and
// the path loss (two normalisations per interaction) is only evaluated for paths that lie on their objects:
// LAZY_LOSS = true computes it here from X (Image / Fermat), false takes the solver's value (MinPath).
template <int MODE, int K, bool LAZY_LOSS>
__device__ __forceinline__ float validity(const SceneTab& T, const KParams& p, const float alpha,
                                          const Cand<K>& cd, const float2 (&X)[K + 2], float loss) {
    // 1. on_objects
    const float onx = on_objects_x<K>(T, cd, X);
    float a_on = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(onx >= 0.0f)) return 0.0f;
    } else if (onx != CUDART_INF_F) {
        a_on = act<MODE>(onx, alpha);
        if (a_on == 0.0f) return 0.0f;
    }
    // 2. loss below tolerance
    if (LAZY_LOSS) loss = path_loss<K>(T, cd, X);
    const float lx = p.tol - loss;
    float a_l = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(lx > 0.0f)) return 0.0f;
    } else {
        if (lx != lx) return 0.0f;  // nan_to_num
        a_l = act<MODE>(lx, alpha);
        if (a_l == 0.0f) return 0.0f;
    }
    // 3. occlusion
    bool alive = true;
    int seg = 0, jj = 0;
    const float interx = intersects_x<MODE, K, false>(T, p.N, cd, X, alpha, alive, seg, jj);
    if (!alive) return 0.0f;
    if (MODE == D2D_MODE_HARD) return 1.0f;
    const float a_in = (interx == -CUDART_INF_F) ? 0.0f : act<MODE>(interx, alpha);
    return fminf(fminf(a_on, 1.0f - a_in), a_l);
}

}  // namespace d2d
