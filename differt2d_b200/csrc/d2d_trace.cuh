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

// ImagePath construction fused with Path.on_objects, last interaction first (the order of the reference's own
// backward scan, geometry.py:1093-1107): the walk stops as soon as one interaction point lies so far off its
// object that act(s) or act(1 - s) is exactly 0 — is_valid is then exactly 0 whatever the rest of the path
// does.  Same operations, same order, same roundings as image_path() + on_objects_x() for the paths that
// survive.  `apex` = image of the transmitter through all K objects (I[K]); the earlier images are only
// built for paths that pass the last interaction.  Returns false for a dead path; X and onx are then partial.
template <int MODE, int K>
__device__ __forceinline__ bool image_path_on(const SceneTab& T, const Cand<K>& cd, const float2 tx, const float2 rx,
                                              const float2 apex, const float alpha, float2 (&X)[K + 2], float& onx) {
    X[0] = tx;
    X[K + 1] = rx;
    onx = CUDART_INF_F;
    if constexpr (K == 0) return true;
    float2 q = rx;
    {
        const int j = cd.c[K > 0 ? K - 1 : 0];
        const float4 w0 = T.w0[j], w1 = T.w1[j];
        q = back_project(q, apex, w0, w1);
        X[K] = q;
        if (T.kind[j] != D2D_KIND_VERTEX) {
            const float s = to_parametric(q, w0, w1);
            float x = fminf(s, 1.0f - s);
            if (s != s) x = -CUDART_INF_F;
            if (act_is_zero<MODE>(x, alpha)) return false;
            onx = x;
        }
    }
    if constexpr (K > 1) {
        float2 I[K];
        I[0] = tx;
#pragma unroll
        for (int i = 0; i + 1 < K; ++i) I[i + 1] = mirror(I[i], T.w0[cd.c[i]], T.w1[cd.c[i]]);
#pragma unroll
        for (int i = K - 2; i >= 0; --i) {
            const int j = cd.c[i];
            const float4 w0 = T.w0[j], w1 = T.w1[j];
            q = back_project(q, I[i + 1], w0, w1);
            X[i + 1] = q;
            if (T.kind[j] == D2D_KIND_VERTEX) continue;
            const float s = to_parametric(q, w0, w1);
            float x = fminf(s, 1.0f - s);
            if (s != s) x = -CUDART_INF_F;
            if (act_is_zero<MODE>(x, alpha)) return false;
            onx = fminf(onx, x);
        }
    }
    return true;
}

// image of `tx` through all K objects of the candidate (the apex seen from the receiver side)
template <int K>
__device__ __forceinline__ float2 image_apex(const SceneTab& T, const Cand<K>& cd, const float2 tx) {
    float2 a = tx;
#pragma unroll
    for (int i = 0; i < K; ++i) a = mirror(a, T.w0[cd.c[i]], T.w1[cd.c[i]]);
    return a;
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

// sum of interaction residuals — geometry.py:1077-1084.  Interaction i uses i_hat = normalize(X[i+1] - X[i]) and
// r_hat = normalize(X[i+2] - X[i+1]) (geometry.py:641-650): the r_hat of one interaction is the i_hat of the next
// (same function of the same inputs, hence the same bits), so the K + 1 segment directions are normalised once.
template <int K>
__device__ __forceinline__ float path_loss(const SceneTab& T, const Cand<K>& cd, const float2 (&X)[K + 2]) {
    float loss = 0.0f;
    if (K == 0) return loss;
    float l;
    float2 dir = normalize2(make_float2(X[1].x - X[0].x, X[1].y - X[0].y), l);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int j = cd.c[i];
        const float2 r = normalize2(make_float2(X[i + 2].x - X[i + 1].x, X[i + 2].y - X[i + 1].y), l);
        loss = loss + residual_dirs(T.kind[j], dir, r, T.w1[j], T.sc[j]);
        dir = r;
    }
    return loss;
}

// path_loss() that also hands out what it normalised: the K + 1 unit directions and lengths (1 for a zero-length
// segment, geometry.py:227-230).  Same operations, same order, same bits; the reverse sweep of the loss takes them from
// here instead of normalising every segment again — twice per interaction, plus the square roots of normalize_adj: 8
// IEEE square roots and 12 IEEE divisions per order-2 path, a third of the sweep's instruction footprint.
template <int K>
__device__ __forceinline__ float path_loss_dirs(const SceneTab& T, const Cand<K>& cd, const float2 (&X)[K + 2],
                                                float2 (&U)[K + 1], float (&Ls)[K + 1]) {
    float loss = 0.0f;
    if (K == 0) return loss;
    U[0] = normalize2(make_float2(X[1].x - X[0].x, X[1].y - X[0].y), Ls[0]);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int j = cd.c[i];
        U[i + 1] = normalize2(make_float2(X[i + 2].x - X[i + 1].x, X[i + 2].y - X[i + 1].y), Ls[i + 1]);
        loss = loss + residual_dirs(T.kind[j], U[i], U[i + 1], T.w1[j], T.sc[j]);
    }
    return loss;
}

// Path.intersects_with_objects in pre-activation form (geometry.py:887-904 + :153-173).  Returns interx
// (-inf: `false_value`).  The double loop runs object-major: one shared-memory read of the object serves the
// K + 1 segments, whose tests are independent instruction streams (ILP), and the whole fold is ONE loop body
// (the segment-major form needed 3 (K + 1) unrolled range loops: 9x the code at K = 2).  max is exact and
// commutative, so the fold order does not matter; for the reverse sweep (TRACK) ties between tests keep the
// first arg-max in the reference's (segment, object) order.
// Fast path per test: canonical a, b, d; approximate parameters through one MUFU.RCP and two FMAs; the exact
// IEEE divisions (hit_exact, out of line) only run when the test could raise the running maximum `interx`.
// `hint` (optional): a per-warp slot holding the object that blocked the warp's previous
// path.  The fold then starts there and wraps around: neighbouring candidates and receivers are mostly blocked by the
// same wall, and a blocked path leaves the loop at its blocker (95 % of the paths that reach the fold are blocked; in
// list order they scan half the scene first).  max is exact and commutative, so the value does not depend on the order.
template <int MODE, int K, bool TRACK>
__device__ __forceinline__ float intersects_x(const SceneTab& T, const int N, const Cand<K>& cd,
                                              const float2 (&X)[K + 2], const float alpha, const float xz,
                                              bool& alive, int& arg_seg, int& arg_j, int* hint = nullptr) {
    // `xz`: tests with hx <= xz do not matter (x_zero<MODE>(alpha), or fold_start<MODE>() for this path)
    float interx = -CUDART_INF_F;
    float cthr = filter_threshold(xz);
    float2 B[K + 1];
    int sa[K + 1], sb[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) {
        B[i] = make_float2(X[i].x - X[i + 1].x, X[i].y - X[i + 1].y);
        sa[i] = (i > 0) ? cd.c[i > 0 ? i - 1 : 0] : -1;  // the segment's own end objects are skipped
        sb[i] = (i < K) ? cd.c[i < K ? i : 0] : -1;
    }
    int j0 = 0, blocker = -1;
    if (hint) {
        j0 = *reinterpret_cast<volatile int*>(hint);
        if (j0 >= N) j0 = 0;
    }
#pragma unroll 1
    for (int jj = 0; jj < N; ++jj) {
        int j = jj + j0;
        if (j >= N) j -= N;
        const float4 w = T.w2[j];
#pragma unroll
        for (int i = 0; i <= K; ++i) {
            if (j == sa[i] || j == sb[i]) continue;
            const float Cx = w.x - X[i].x, Cy = w.y - X[i].y;
            const float a = B[i].y * Cx - B[i].x * Cy;
            const float b = w.z * Cy - w.w * Cx;
            const float d = w.w * B[i].x - w.z * B[i].y;
            const float r = rcp_approx(d);
            const float qa = fmaf(a, r, -0.5f);
            const float qb = fmaf(b, r, -0.5f);
            const float m = fmaxf(fabsf(qa), fabsf(qb));
            if (!TRACK) D2D_COUNT(27);
            if (!(m >= cthr)) {
                if (!TRACK) D2D_COUNT(28);
                const float hx = hit_exact(a, b, d);
                // ties between tests (TRACK: the reverse sweep follows ONE arg-max): the first in the reference's
                // (segment, object) order wins, whatever order this fold visits the objects in
                if (hx > interx || (TRACK && hx == interx && (i < arg_seg || (i == arg_seg && j < arg_j)))) {
                    interx = hx;
                    if (TRACK) { arg_j = j; arg_seg = i; }
                    cthr = filter_threshold(fmaxf(hx, xz));
                    // the path is dead once `intersects` is exactly true / 1.0
                    if (MODE == D2D_MODE_HARD) alive = !(hx >= 0.0f);
                    else alive = !act_is_one<MODE>(hx, alpha);
                }
            }
        }
        if (!alive) {
            blocker = j;
            if (!TRACK) { D2D_COUNT(26); if (jj == 0) D2D_COUNT(25); }
            break;
        }
    }
    if (hint && blocker >= 0) *reinterpret_cast<volatile int*>(hint) = blocker;  // (any lane's: a heuristic)
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
__device__ __forceinline__ float validity_from_onx(const SceneTab& T, const KParams& p, const float alpha,
                                                   const Cand<K>& cd, const float2 (&X)[K + 2], float loss,
                                                   const float onx, int* hint = nullptr);

template <int MODE, int K, bool LAZY_LOSS>
__device__ __forceinline__ float validity(const SceneTab& T, const KParams& p, const float alpha,
                                          const Cand<K>& cd, const float2 (&X)[K + 2], float loss, int* hint = nullptr) {
    return validity_from_onx<MODE, K, LAZY_LOSS>(T, p, alpha, cd, X, loss, on_objects_x<K>(T, cd, X), hint);
}

template <int MODE, int K, bool LAZY_LOSS>
__device__ __forceinline__ float validity_from_onx(const SceneTab& T, const KParams& p, const float alpha,
                                                   const Cand<K>& cd, const float2 (&X)[K + 2], float loss,
                                                   const float onx, int* hint) {
    // 1. on_objects
    float a_on = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(onx >= 0.0f)) return 0.0f;
    } else if (onx != CUDART_INF_F) {
        a_on = act<MODE>(onx, alpha);
        if (a_on == 0.0f) return 0.0f;
    }
    // 2. loss below tolerance
    // (Measured and removed: deferring the canonical loss behind the fold — an estimate through MUFU.RSQ first, to drop
    // the wrong-side paths, the IEEE square roots and divisions only for paths nothing occludes.  Bit-identical, 184
    // parity tests green, and the raw city forward went from 0.945 to 1.026 ms: 6 % MORE warp instructions
    // (profiles/r02w), the estimate's lanes diverge from the fold's instead of leaving the warp early.)
    D2D_COUNT(22);
    if (LAZY_LOSS) loss = path_loss<K>(T, cd, X);
    const float lx = p.tol - loss;
    float a_l = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(lx > 0.0f)) { D2D_COUNT(23); return 0.0f; }
    } else {
        if (lx != lx) return 0.0f;  // nan_to_num
        a_l = act<MODE>(lx, alpha);
        if (a_l == 0.0f) { D2D_COUNT(23); return 0.0f; }
    }
    // 3. occlusion — unless it cannot matter (fold_skip_bound): is_valid is then min(a_on, a_l) exactly
    float xz = x_zero<MODE>(alpha);
    if (MODE != D2D_MODE_HARD) {
        const float v0 = fminf(a_on, a_l);
        if (v0 <= fold_skip_of<MODE>(T, alpha)) return v0;
        if (MODE == D2D_MODE_SIGMOID && T.fold_skip > -CUDART_INF_F) xz = fold_start<MODE>(v0, alpha, xz);
    }
    D2D_COUNT(24);
    bool alive = true;
    int seg = 0, jj = 0;
    const float interx = intersects_x<MODE, K, false>(T, p.N, cd, X, alpha, xz, alive, seg, jj, hint);
    if (!alive) return 0.0f;
    if (MODE == D2D_MODE_HARD) return 1.0f;
    const float a_in = (interx == -CUDART_INF_F) ? 0.0f : act<MODE>(interx, alpha);
    return fminf(fminf(a_on, 1.0f - a_in), a_l);
}

}  // namespace d2d
