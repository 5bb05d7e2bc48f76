// differt2d_b200 — ImagePath: ONE tracked re-trace + the reverse sweep, fully inlined (register resident).
//
// Round 1's backward kernel re-traced a path lightly, and for the paths with a non-zero validity called an
// out-of-line path_vjp() that traced it AGAIN (with arg-min / arg-max tracking) and returned its cotangents through
// reference arguments — i.e. through a 900-byte stack frame per thread: 3.8e6 local-memory stores per launch whose
// footprint (resident threads x frame) did not fit the L2 and showed up as 139 MB of DRAM writes against 8 MB of
// algorithmic output (profiles/r02b_ncu_summary.md), and where nothing is prunable (sigmoid, small alpha: every path
// is valid) the backward cost 4.7 forward launches.  Here the trace tracks what the reverse sweep needs while it
// runs, the sweep starts from the live registers, and nothing crosses a call boundary.
//
// Arithmetic: the re-trace is the forward kernel's (canonical fp32, bit for bit: it decides which paths are alive and
// which comparison carries the cotangent).  The sweep itself feeds no predicate; it uses explicit FMAs and the fast
// reciprocal where that helps (gradients are compared at rtol 1e-4).
#pragma once

#include "d2d_adjoint.cuh"
#include "d2d_trace.cuh"

namespace d2d {

__device__ __forceinline__ float fdiv(const float a, const float b) { return __fdividef(a, b); }

template <int MODE>
__device__ __forceinline__ float act_and_dz(const float x, const float alpha, float& dz);

// act(x) and d act / d z at z = alpha x in one evaluation (the sigmoid's exponential is shared)
template <int MODE>
__device__ __forceinline__ float act_and_dz(const float x, const float alpha, float& dz) {
    const float z = alpha * x;
    if (MODE == D2D_MODE_SIGMOID) {
        const float s = 1.0f / (1.0f + expf(-z));  // == act<MODE>(x, alpha), bit for bit
        dz = s * (1.0f - s);
        return s;
    }
    const float v = z + 3.0f;
    float g = (v > 0.0f && v < 6.0f) ? 1.0f : 0.0f;
    if (v == 0.0f || v == 6.0f) g = 0.5f;  // jnp.maximum / jnp.minimum split ties 1/2 - 1/2
    dz = g * (1.0f / 6.0f);
    if (!(v > 0.0f)) return 0.0f;
    if (v >= 6.0f) return 1.0f;
    return v / 6.0f;
}

template <int K>
struct ImageTrace {
    float2 X[K + 2];
    float2 U[K + 1];   // unit directions of the segments, as path_loss normalised them
    float Ls[K + 1];   // ... and their lengths (1 for a zero-length segment)
    float on_s, a_on, a_l, a_in, lx, interx, valid, val, r;
    int on_i, seg, jj;
};

// d act / d z at z = alpha x from the activation the trace already evaluated there (sigmoid: s (1 - s), no second
// exponential; same bits as act_and_dz, which forms s the same way)
template <int MODE>
__device__ __forceinline__ float dz_from_act(const float a, const float x, const float alpha) {
    if (MODE == D2D_MODE_SIGMOID) return a * (1.0f - a);
    float dz;
    act_and_dz<MODE>(x, alpha, dz);
    return dz;
}

// The forward kernel's path evaluation (image_path_on + validity_from_onx, same operations in the same order) that
// also records the arg-min interaction of on_objects and the arg-max (segment, object) test of the occlusion fold.
// Returns false when the validity is exactly 0.
template <int MODE, int K>
__device__ __forceinline__ bool trace_image_tracked(const SceneTab& T, const KParams& p, const float alpha,
                                                    const Cand<K>& cd, const float2 tx, const float2 rx,
                                                    const float2 apex, ImageTrace<K>& tr, int* hint) {
    tr.X[0] = tx;
    tr.X[K + 1] = rx;
    float onx = CUDART_INF_F;
    tr.on_i = -1;
    tr.on_s = 0.f;
    if constexpr (K > 0) {
        float2 q = rx;
        {
            const int j = cd.c[K - 1];
            const float4 w0 = T.w0[j], w1 = T.w1[j];
            q = back_project(q, apex, w0, w1);
            tr.X[K] = q;
            if (T.kind[j] != D2D_KIND_VERTEX) {
                const float s = to_parametric(q, w0, w1);
                float x = fminf(s, 1.0f - s);
                if (s != s) x = -CUDART_INF_F;
                if (act_is_zero<MODE>(x, alpha)) return false;
                onx = x; tr.on_i = K - 1; tr.on_s = s;
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
                tr.X[i + 1] = q;
                if (T.kind[j] == D2D_KIND_VERTEX) continue;
                const float s = to_parametric(q, w0, w1);
                float x = fminf(s, 1.0f - s);
                if (s != s) x = -CUDART_INF_F;
                if (act_is_zero<MODE>(x, alpha)) return false;
                if (x <= onx) { onx = x; tr.on_i = i; tr.on_s = s; }  // ties: the first interaction in list order
            }
        }
    }
    tr.a_on = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(onx >= 0.0f)) return false;
    } else if (onx != CUDART_INF_F) {
        tr.a_on = act<MODE>(onx, alpha);
        if (tr.a_on == 0.0f) return false;
    }
    // sigmoid: (nearly) every path goes on to the reverse sweep, which takes the normalised segments from here.  The
    // other logics reverse a path in a hundred: keeping 3 (K + 1) more values alive through the fold cost their
    // un-masked re-trace (the fused value + VJP launch) 7 % at 64 registers; the sweep normalises again there.
    const float loss = (MODE == D2D_MODE_SIGMOID) ? path_loss_dirs<K>(T, cd, tr.X, tr.U, tr.Ls) : path_loss<K>(T, cd, tr.X);
    tr.lx = p.tol - loss;
    tr.a_l = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(tr.lx > 0.0f)) return false;
    } else {
        if (tr.lx != tr.lx) return false;  // nan_to_num
        tr.a_l = act<MODE>(tr.lx, alpha);
        if (tr.a_l == 0.0f) return false;
    }
    bool alive = true;
    tr.seg = -1;
    tr.jj = -1;
    tr.interx = -CUDART_INF_F;
    tr.valid = 1.0f;
    tr.a_in = 0.0f;
    float xz = x_zero<MODE>(alpha);
    bool fold = true;
    if (MODE != D2D_MODE_HARD) {
        // the fold cannot matter (fold_skip_bound): validity = min(a_on, a_l) exactly, 1 - a_in is strictly above it
        // (no tie, the cotangent follows a_on / a_l); otherwise it only has to look for tests above fold_start
        const float v0 = fminf(tr.a_on, tr.a_l);
        if (v0 <= fold_skip_of<MODE>(T, alpha)) { fold = false; tr.valid = v0; }
        else if (MODE == D2D_MODE_SIGMOID && T.fold_skip > -CUDART_INF_F) xz = fold_start<MODE>(v0, alpha, xz);
    }
    if (fold) tr.interx = intersects_x<MODE, K, true>(T, p.N, cd, tr.X, alpha, xz, alive, tr.seg, tr.jj, hint);
    if (!alive) return false;
    if (MODE != D2D_MODE_HARD && fold) {
        tr.a_in = (tr.interx == -CUDART_INF_F) ? 0.0f : act<MODE>(tr.interx, alpha);
        tr.valid = fminf(fminf(tr.a_on, 1.0f - tr.a_in), tr.a_l);
        if (tr.valid == 0.0f) return false;
    }
    tr.val = path_value<K>(p, tr.X, tr.r);
    return true;
}

// Reverse sweep of one traced ImagePath (clean gradients) for the cotangents of its two outputs: the
// validity (valid_bar) and the path vertices (Xb: in = cotangent of tr.X, used as the accumulator).  Fills tx_bar,
// rx_bar, alpha_bar, oa[] (per interacting object) and the occluder's vertex cotangent (occ_j, occ_bar).
template <int MODE, int K>
__device__ __forceinline__ void image_reverse_general(const SceneTab& T, const KParams& p, const float alpha,
                                                      const Cand<K>& cd, const ImageTrace<K>& tr, const float valid_bar,
                                                      float2 (&Xb)[K + 2], float2& tx_bar, float2& rx_bar,
                                                      float& alpha_bar, ObjAdj (&oa)[K > 0 ? K : 1], int& occ_j,
                                                      float4& occ_bar) {
    const float2(&X)[K + 2] = tr.X;
#pragma unroll
    for (int i = 0; i < (K > 0 ? K : 1); ++i) oa[i].zero();
    alpha_bar = 0.f;
    occ_j = -1;
    // (2) validity (smooth logic only): geometry.py:947-963, logic.py:511-512
    if (MODE != D2D_MODE_HARD) {
        const float v1 = tr.a_on, v2 = 1.0f - tr.a_in, v3 = tr.a_l;
        const int cnt = (v1 == tr.valid) + (v2 == tr.valid) + (v3 == tr.valid);
        const float share = valid_bar * (cnt == 1 ? 1.0f : (cnt == 2 ? 0.5f : (1.0f / 3.0f)));  // jnp.min: even split
        if (v1 == tr.valid && tr.on_i >= 0) {
            // contains_parametric = minimum(act(s - 0), act(1 - s)) (geometry.py:608-621); jnp.minimum's tie rule
            // (1/2, 1/2) applies to the ACTIVATED values, which tie far more often than the pre-activations do
            const float xg = tr.on_s, xl = 1.0f - tr.on_s;
            float dg, dl, wg;
            // a_on IS the activation of the smaller of the two (onx = min(xg, xl), monotone map); the other one only
            // matters if it ties with it after rounding, which a sigmoid away from saturation cannot do once the
            // pre-activations differ by 1e-3 / alpha (slope >= 1e-3 there: 1e-6 apart, 8 ulp): no exponential at all
            // for nearly every path of a soft-activation run
            const float xgap = fabsf(xg - xl);
            if (MODE == D2D_MODE_SIGMOID && tr.a_on < 0.999f && tr.a_on > 1e-30f && alpha * xgap > 1e-3f) {
                const float dm = tr.a_on * (1.0f - tr.a_on);
                wg = xg < xl ? 1.0f : 0.0f;
                dg = dm; dl = dm;  // (the one with weight 0 is not used)
            } else {
                const float Ag = act_and_dz<MODE>(xg, alpha, dg), Al = act_and_dz<MODE>(xl, alpha, dl);
                wg = Ag < Al ? 1.0f : (Ag == Al ? 0.5f : 0.0f);
            }
            const float wl = 1.0f - wg;
            const float s_bar = share * alpha * (wg * dg - wl * dl);
            alpha_bar = fmaf(share, wg * xg * dg + wl * xl * dl, alpha_bar);
            if (s_bar != 0.f) {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (i != tr.on_i) continue;
                    const float4 w0 = T.w0[cd.c[i]];
                    const float4 w1 = T.w1[cd.c[i]];
                    const float wx = X[i + 1].x - w0.x, wy = X[i + 1].y - w0.y;
                    const float k = fdiv(s_bar, w1.z);
                    Xb[i + 1].x = fmaf(k, w0.z, Xb[i + 1].x); Xb[i + 1].y = fmaf(k, w0.w, Xb[i + 1].y);
                    oa[i].p1.x -= k * w0.z; oa[i].p1.y -= k * w0.w;
                    oa[i].t.x += k * wx; oa[i].t.y += k * wy;
                    oa[i].tt += -k * tr.on_s;
                }
            }
        }
        if (v3 == tr.valid) {
            const float dz = dz_from_act<MODE>(tr.a_l, tr.lx, alpha);
            const float loss_bar = -share * alpha * dz;
            alpha_bar = fmaf(share, tr.lx * dz, alpha_bar);
            if (loss_bar != 0.f) {
                float2 U[K + 1];
                float Ls[K + 1];
                if (MODE == D2D_MODE_SIGMOID) {
#pragma unroll
                    for (int i = 0; i <= K; ++i) { U[i] = tr.U[i]; Ls[i] = tr.Ls[i]; }
                } else {
                    path_loss_dirs<K>(T, cd, X, U, Ls);  // the trace's own bits again (see trace_image_tracked)
                }
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const int j = cd.c[i];
                    residual_adj_dirs(T.kind[j], U[i], Ls[i], U[i + 1], Ls[i + 1], T.w1[j], T.sc[j], loss_bar,
                                      Xb[i], Xb[i + 1], Xb[i + 2], oa[i].n, oa[i].phi);
                }
            }
        }
        if (v2 == tr.valid && tr.jj >= 0 && tr.interx != -CUDART_INF_F) {
            const float gsh = -share;  // d valid / d a_in = -1
            float2 P = X[0], Q = X[1];
#pragma unroll
            for (int i = 0; i <= K; ++i)
                if (i == tr.seg) { P = X[i]; Q = X[i + 1]; }
            const float4 w = T.w2[tr.jj];
            const float2 B = make_float2(P.x - Q.x, P.y - Q.y);
            const float Cx = w.x - P.x, Cy = w.y - P.y;
            const float a = B.y * Cx - B.x * Cy;
            const float b = w.z * Cy - w.w * Cx;
            const float d = w.w * B.x - w.z * B.y;
            const float ta = a / d, tb = b / d;  // (canonical: these decide the tie weights below)
            // hit = minimum(minimum(ge_a, le_a), minimum(ge_b, le_b)) on ACTIVATED values (geometry.py:167-173)
            const float x1 = ta + kTolSeg, x2 = kHiSeg - ta, x3 = tb + kTolSeg, x4 = kHiSeg - tb;
            float d1, d2, d3, d4;
            const float A1 = act_and_dz<MODE>(x1, alpha, d1), A2 = act_and_dz<MODE>(x2, alpha, d2);
            const float A3 = act_and_dz<MODE>(x3, alpha, d3), A4 = act_and_dz<MODE>(x4, alpha, d4);
            const float Ta = fminf(A1, A2), Tb = fminf(A3, A4);
            const float wa = Ta < Tb ? 1.0f : (Ta == Tb ? 0.5f : 0.0f), wb = 1.0f - wa;
            const float w1 = wa * (A1 < A2 ? 1.0f : (A1 == A2 ? 0.5f : 0.0f)), w2 = wa - w1;
            const float w3 = wb * (A3 < A4 ? 1.0f : (A3 == A4 ? 0.5f : 0.0f)), w4 = wb - w3;
            alpha_bar = fmaf(gsh, w1 * x1 * d1 + w2 * x2 * d2 + w3 * x3 * d3 + w4 * x4 * d4, alpha_bar);
            const float ta_bar = gsh * alpha * (w1 * d1 - w2 * d2);
            const float tb_bar = gsh * alpha * (w3 * d3 - w4 * d4);
            if (ta_bar != 0.f || tb_bar != 0.f) {
                const float rd = fdiv(1.0f, d);
                const float a_bar = ta_bar * rd, b_bar = tb_bar * rd;
                const float d_bar = -(ta_bar * ta + tb_bar * tb) * rd;
                float2 Bb, Cb, Ab;
                Bb.y = a_bar * Cx - d_bar * w.z;
                Bb.x = -a_bar * Cy + d_bar * w.w;
                Cb.x = a_bar * B.y - b_bar * w.w;
                Cb.y = -a_bar * B.x + b_bar * w.z;
                Ab.x = b_bar * Cy - d_bar * B.y;
                Ab.y = -b_bar * Cx + d_bar * B.x;
                const float2 Pb = make_float2(Bb.x - Cb.x, Bb.y - Cb.y);
#pragma unroll
                for (int i = 0; i <= K; ++i)
                    if (i == tr.seg) {
                        Xb[i].x += Pb.x; Xb[i].y += Pb.y;
                        Xb[i + 1].x -= Bb.x; Xb[i + 1].y -= Bb.y;
                    }
                // P1' = (1+patch) P1 - patch P2 ; P2' = (1+patch) P2 - patch P1 ; A = P2' - P1' ; C = P1' - P
                const float2 p1p = make_float2(Cb.x - Ab.x, Cb.y - Ab.y);
                const float2 p2p = Ab;
                const float q = p.patch;
                occ_j = tr.jj;
                occ_bar = make_float4((1.0f + q) * p1p.x - q * p2p.x, (1.0f + q) * p1p.y - q * p2p.y,
                                      (1.0f + q) * p2p.x - q * p1p.x, (1.0f + q) * p2p.y - q * p1p.y);
            }
        }
    }
    // (3) backward scan of the image method: geometry.py:1093-1107 (clean `where`), then (4) the mirror chain
    // geometry.py:1086-1091, 652-670.  The images are rebuilt here (K mirrors) instead of being kept alive.
    float2 I[K + 1];
    I[0] = tr.X[0];
#pragma unroll
    for (int i = 0; i < K; ++i) I[i + 1] = mirror(I[i], T.w0[cd.c[i]], T.w1[cd.c[i]]);
    float2 Ib[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) Ib[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float4 w0 = T.w0[cd.c[i]];
        const float4 w1 = T.w1[cd.c[i]];
        const float2 qb = Xb[i + 1];
        const float2 pt = X[i + 2];
        const float ux = pt.x - I[i + 1].x, uy = pt.y - I[i + 1].y;
        const float vx = w0.x - pt.x, vy = w0.y - pt.y;
        const float un = ux * w1.x + uy * w1.y;  // (canonical: decides the masked branch exactly like the forward)
        const float vn = vx * w1.x + vy * w1.y;
        Xb[i + 2].x += qb.x; Xb[i + 2].y += qb.y;
        if (un != 0.0f) {
            const float run = fdiv(1.0f, un);
            const float g = vn * run;
            float ubx = g * qb.x, uby = g * qb.y;
            const float gb = fmaf(qb.x, ux, qb.y * uy);
            const float vnb = gb * run;
            const float unb = -gb * g * run;
            ubx = fmaf(unb, w1.x, ubx); uby = fmaf(unb, w1.y, uby);
            oa[i].n.x += fmaf(unb, ux, vnb * vx);
            oa[i].n.y += fmaf(unb, uy, vnb * vy);
            const float vbx = vnb * w1.x, vby = vnb * w1.y;
            Xb[i + 2].x += ubx - vbx; Xb[i + 2].y += uby - vby;
            Ib[i + 1].x -= ubx; Ib[i + 1].y -= uby;
            oa[i].p1.x += vbx; oa[i].p1.y += vby;
        }
    }
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        const float4 w0 = T.w0[cd.c[i]];
        const float4 w1 = T.w1[cd.c[i]];
        const float wx = I[i].x - w0.x, wy = I[i].y - w0.y;
        const float cc = 2.0f * fmaf(wx, w1.x, wy * w1.y);
        const float2 ib = Ib[i + 1];
        const float ccb = -fmaf(ib.x, w1.x, ib.y * w1.y);
        const float dotb = 2.0f * ccb;
        oa[i].n.x += fmaf(-cc, ib.x, dotb * wx);
        oa[i].n.y += fmaf(-cc, ib.y, dotb * wy);
        const float wbx = dotb * w1.x, wby = dotb * w1.y;
        Ib[i].x += ib.x + wbx; Ib[i].y += ib.y + wby;
        oa[i].p1.x -= wbx; oa[i].p1.y -= wby;
    }
    tx_bar = make_float2(Xb[0].x + Ib[0].x, Xb[0].y + Ib[0].y);
    rx_bar = Xb[K + 1];
}

// The fused `fun` (received_power / length**2, utils.py:52-54; path_length geometry.py:199-203) for the upstream
// cotangent zbar of Z: d(valid * val) = val dvalid + valid dval.
template <int MODE, int K>
__device__ __forceinline__ void image_reverse(const SceneTab& T, const KParams& p, const float alpha, const Cand<K>& cd,
                                              const ImageTrace<K>& tr, const float zbar, float2& tx_bar, float2& rx_bar,
                                              float& alpha_bar, ObjAdj (&oa)[K > 0 ? K : 1], int& occ_j, float4& occ_bar) {
    const float2(&X)[K + 2] = tr.X;
    float2 Xb[K + 2];
#pragma unroll
    for (int i = 0; i < K + 2; ++i) Xb[i] = make_float2(0.f, 0.f);
    const float val_bar = zbar * tr.valid;
    float r_bar;
    if (p.fun == D2D_FUN_RECEIVED_POWER) r_bar = -val_bar * tr.val * 2.0f * fdiv(tr.r, fmaf(tr.r, tr.r, p.h2));
    else r_bar = val_bar * 2.0f * tr.r;
#pragma unroll
    for (int i = 0; i <= K; ++i) {
        const float dx = (X[i + 1].x - X[i].x) + kEps32;
        const float dy = (X[i + 1].y - X[i].y) + kEps32;
        const float sq = fmaf(dx, dx, dy * dy);
        if (sq > 0.0f) {
            const float c = r_bar * rsqrtf(sq);
            Xb[i + 1].x = fmaf(c, dx, Xb[i + 1].x); Xb[i + 1].y = fmaf(c, dy, Xb[i + 1].y);
            Xb[i].x = fmaf(-c, dx, Xb[i].x); Xb[i].y = fmaf(-c, dy, Xb[i].y);
        }
    }
    image_reverse_general<MODE, K>(T, p, alpha, cd, tr, zbar * tr.val, Xb, tx_bar, rx_bar, alpha_bar, oa, occ_j, occ_bar);
}

}  // namespace d2d
