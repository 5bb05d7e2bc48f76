// differt2d_b200 — reverse mode through the Adam iterations of FermatPath / MinPath.
//
// jax.grad over Scene.accumulate_* with path_cls=FermatPath|MinPath differentiates the whole
// `lax.scan` of optimize.minimize (optimize.py:85-97): theta_S depends on the transmitter, the
// receiver and the interacting objects through every iterate.  The VJP of one Adam step
//     g = grad_theta L(theta, q)            (q = tx, rx, object parameters)
//     mu' = (1-b1) g + b1 mu ;  nu' = (1-b2) g^2 + b2 nu
//     theta' = theta - lr (mu'/bc1) / (sqrt(nu'/bc2) + eps)
// needs the VJP of g, i.e. for a cotangent lam_g the gradient of  Phi = lam_g . grad_theta L  w.r.t.
// (theta, q).  Phi is the directional derivative of L along lam_g in theta-space, so its gradient is the
// directional derivative of the FULL gradient grad_(theta,q) L — obtained here by running the hand-written
// reverse sweep of L in dual-number arithmetic (theta + e lam_g) and reading the e-parts
// ("forward over reverse").  No Hessian is formed; everything stays in registers.
//
// The scan stores nothing in the forward kernel.  The reverse sweep re-runs the iterations once keeping a
// checkpoint of (theta, mu, nu, b1^s, b2^s) every kBlk steps, then walks the checkpoints backwards,
// re-running each block while keeping its kBlk states (thread-local memory), and reverses step by step:
// 3 loss gradients + 1 dual gradient per step and path, only for the paths whose validity is non-zero.
#pragma once

#include "d2d_adjoint.cuh"
#include "d2d_trace.cuh"

namespace d2d {

// ---- dual numbers -------------------------------------------------------------------------------
struct Dual {
    float v, d;
};
__device__ __forceinline__ Dual mk(float v, float d) { Dual r; r.v = v; r.d = d; return r; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return mk(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return mk(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return mk(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
    const float q = a.v / b.v;
    return mk(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ Dual operator+(Dual a, float b) { return mk(a.v + b, a.d); }
__device__ __forceinline__ Dual operator+(float a, Dual b) { return mk(a + b.v, b.d); }
__device__ __forceinline__ Dual operator-(Dual a, float b) { return mk(a.v - b, a.d); }
__device__ __forceinline__ Dual operator-(float a, Dual b) { return mk(a - b.v, -b.d); }
__device__ __forceinline__ Dual operator*(Dual a, float b) { return mk(a.v * b, a.d * b); }
__device__ __forceinline__ Dual operator*(float a, Dual b) { return mk(a * b.v, a * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, float b) { return mk(a.v / b, a.d / b); }
__device__ __forceinline__ Dual operator/(float a, Dual b) {
    const float q = a / b.v;
    return mk(q, -q * b.d / b.v);
}
__device__ __forceinline__ Dual operator-(Dual a) { return mk(-a.v, -a.d); }

__device__ __forceinline__ float prim(float a) { return a; }
__device__ __forceinline__ float prim(Dual a) { return a.v; }
__device__ __forceinline__ float tang(float) { return 0.f; }
__device__ __forceinline__ float tang(Dual a) { return a.d; }
__device__ __forceinline__ float gsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ Dual gsqrt(Dual a) {
    const float s = sqrtf(a.v);
    return mk(s, s > 0.f ? 0.5f * a.d / s : 0.f);
}
template <class S> __device__ __forceinline__ S lift(float a);
template <> __device__ __forceinline__ float lift<float>(float a) { return a; }
template <> __device__ __forceinline__ Dual lift<Dual>(float a) { return mk(a, 0.f); }

template <class S>
struct V2 {
    S x, y;
};
template <class S> __device__ __forceinline__ V2<S> v2(S x, S y) { V2<S> r; r.x = x; r.y = y; return r; }
template <class S> __device__ __forceinline__ V2<S> v2f(float2 a) { return v2<S>(lift<S>(a.x), lift<S>(a.y)); }
template <class S> __device__ __forceinline__ S gdot(V2<S> a, V2<S> b) { return a.x * b.x + a.y * b.y; }

// normalize (geometry.py:206-230): v / |v|, |v| == 0 -> 1
template <class S>
__device__ __forceinline__ V2<S> gnormalize(V2<S> v, S& len, bool& zero) {
    len = gsqrt(v.x * v.x + v.y * v.y);
    zero = prim(len) == 0.0f;
    if (zero) len = lift<S>(1.0f);
    return v2<S>(v.x / len, v.y / len);
}
// VJP of normalize (clean: |v| == 0 is the constant branch v / 1)
template <class S>
__device__ __forceinline__ V2<S> gnormalize_adj(V2<S> vh, S len, bool zero, V2<S> vh_bar) {
    if (zero) return vh_bar;
    const S pr = gdot(vh_bar, vh);
    return v2<S>((vh_bar.x - pr * vh.x) / len, (vh_bar.y - pr * vh.y) / len);
}

// per-object adjoints touched by the solver's loss: origin, direction, unit normal, RIS angle
template <class S>
struct ObjAdjS {
    V2<S> p1, t, n;
    S phi;
};

// Gradient of  w * loss_fun(theta)  w.r.t. theta, tx, rx and the interacting objects' table entries, in
// S arithmetic.  loss_fun: FermatPath geometry.py:1185-1188 (path_length), MinPath :1270-1276 (sum of
// evaluate_cartesian), both through parametric_to_cartesian (:976-1010).  Outputs are OVERWRITTEN.
template <int METHOD, int K, class S>
__device__ __forceinline__ void loss_grad_full(const SceneTab& T, const Cand<K>& cd, const S (&th)[K > 0 ? K : 1],
                                               const float2 tx, const float2 rx, const float w,
                                               S (&th_bar)[K > 0 ? K : 1], V2<S>& tx_bar, V2<S>& rx_bar,
                                               ObjAdjS<S> (&oa)[K > 0 ? K : 1]) {
    V2<S> X[K + 2], Xb[K + 2];
    X[0] = v2f<S>(tx);
    X[K + 1] = v2f<S>(rx);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float4 w0 = T.w0[cd.c[i]];
        if (T.kind[cd.c[i]] == D2D_KIND_VERTEX) X[i + 1] = v2<S>(lift<S>(w0.x), lift<S>(w0.y));
        else X[i + 1] = v2<S>(w0.x + th[i] * w0.z, w0.y + th[i] * w0.w);
    }
#pragma unroll
    for (int i = 0; i < K + 2; ++i) Xb[i] = v2<S>(lift<S>(0.f), lift<S>(0.f));
#pragma unroll
    for (int i = 0; i < K; ++i) {
        oa[i].p1 = oa[i].t = oa[i].n = v2<S>(lift<S>(0.f), lift<S>(0.f));
        oa[i].phi = lift<S>(0.f);
    }
    if (METHOD == D2D_METHOD_FERMAT) {
#pragma unroll
        for (int i = 0; i <= K; ++i) {  // path_length geometry.py:176-203
            const S dx = (X[i + 1].x - X[i].x) + kEps32;
            const S dy = (X[i + 1].y - X[i].y) + kEps32;
            const S len = gsqrt(dx * dx + dy * dy);
            if (prim(len) > 0.0f) {
                const S c = w / len;
                Xb[i + 1].x = Xb[i + 1].x + c * dx; Xb[i + 1].y = Xb[i + 1].y + c * dy;
                Xb[i].x = Xb[i].x - c * dx; Xb[i].y = Xb[i].y - c * dy;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) {  // evaluate_cartesian: Wall geometry.py:641-650, RIS :698-711
            const int j = cd.c[i];
            const int kind = T.kind[j];
            if (kind == D2D_KIND_VERTEX) continue;
            const float4 w1 = T.w1[j];
            const V2<S> n = v2<S>(lift<S>(w1.x), lift<S>(w1.y));
            const V2<S> rv = v2<S>(X[i + 2].x - X[i + 1].x, X[i + 2].y - X[i + 1].y);
            S rl; bool rz;
            const V2<S> r = gnormalize(rv, rl, rz);
            V2<S> r_bar;
            if (kind == D2D_KIND_WALL) {
                const V2<S> iv = v2<S>(X[i + 1].x - X[i].x, X[i + 1].y - X[i].y);
                S il; bool iz;
                const V2<S> ih = gnormalize(iv, il, iz);
                const S c2 = 2.0f * gdot(ih, n);
                const V2<S> e = v2<S>(r.x - (ih.x - c2 * n.x), r.y - (ih.y - c2 * n.y));
                const V2<S> e_bar = v2<S>((2.0f * w) * e.x, (2.0f * w) * e.y);
                const S en = gdot(e_bar, n);
                r_bar = e_bar;
                const V2<S> i_bar = v2<S>(2.0f * en * n.x - e_bar.x, 2.0f * en * n.y - e_bar.y);
                oa[i].n.x = oa[i].n.x + (c2 * e_bar.x + 2.0f * en * ih.x);
                oa[i].n.y = oa[i].n.y + (c2 * e_bar.y + 2.0f * en * ih.y);
                const V2<S> iv_bar = gnormalize_adj(ih, il, iz, i_bar);
                Xb[i + 1].x = Xb[i + 1].x + iv_bar.x; Xb[i + 1].y = Xb[i + 1].y + iv_bar.y;
                Xb[i].x = Xb[i].x - iv_bar.x; Xb[i].y = Xb[i].y - iv_bar.y;
            } else {  // RIS
                const float2 sc = T.sc[j];
                const S mx = -r.x, my = -r.y;
                const S sin_a = mx * n.y - my * n.x;
                const S cos_a = mx * n.x + my * n.y;
                const S ds = (2.0f * w) * (sin_a - sc.x);
                const S dc = (2.0f * w) * (cos_a - sc.y);
                oa[i].phi = oa[i].phi + (dc * sc.x - ds * sc.y);
                const S mbx = ds * n.y + dc * n.x;
                const S mby = dc * n.y - ds * n.x;
                r_bar = v2<S>(-mbx, -mby);
                oa[i].n.x = oa[i].n.x + (dc * mx - ds * my);
                oa[i].n.y = oa[i].n.y + (ds * mx + dc * my);
            }
            const V2<S> rv_bar = gnormalize_adj(r, rl, rz, r_bar);
            Xb[i + 2].x = Xb[i + 2].x + rv_bar.x; Xb[i + 2].y = Xb[i + 2].y + rv_bar.y;
            Xb[i + 1].x = Xb[i + 1].x - rv_bar.x; Xb[i + 1].y = Xb[i + 1].y - rv_bar.y;
        }
    }
    // parametric_to_cartesian: X = P1 + theta t (Wall geometry.py:581-587), X = xy (Vertex :383-389)
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float4 w0 = T.w0[cd.c[i]];
        oa[i].p1.x = oa[i].p1.x + Xb[i + 1].x;
        oa[i].p1.y = oa[i].p1.y + Xb[i + 1].y;
        if (T.kind[cd.c[i]] == D2D_KIND_VERTEX) {
            th_bar[i] = lift<S>(0.f);
        } else {
            th_bar[i] = Xb[i + 1].x * w0.z + Xb[i + 1].y * w0.w;
            oa[i].t.x = oa[i].t.x + th[i] * Xb[i + 1].x;
            oa[i].t.y = oa[i].t.y + th[i] * Xb[i + 1].y;
        }
    }
    tx_bar = Xb[0];
    rx_bar = Xb[K + 1];
}

// ---- the Adam scan, with checkpoints -----------------------------------------------------------------
constexpr int kBlk = 32;   // states kept per block / steps between checkpoints (times `mult`)
constexpr int kNck = 32;   // checkpoints

template <int K>
struct AdamState {
    float th[K > 0 ? K : 1], mu[K > 0 ? K : 1], nu[K > 0 ? K : 1];
    float b1p, b2p;
};


// One iteration of optimize.py:87-93, identical (bit for bit) to the loop of construct_path (d2d_solver.cuh).
// Returns the loss at the iterate BEFORE the update.
template <int METHOD, int K>
__device__ __forceinline__ float adam_step(const SceneTab& T, const KParams& p, const Cand<K>& cd, const float2 tx,
                                           const float2 rx, AdamState<K>& st) {
    float2 X[K + 2], G[K + 2];
    X[0] = tx;
    X[K + 1] = rx;
    place_points<K>(T, cd, st.th, X);
    const float loss = solver_loss_grad<METHOD, K>(T, cd, X, G);
    st.b1p *= p.b1;
    st.b2p *= p.b2;
    const float bc1 = 1.0f - st.b1p, bc2 = 1.0f - st.b2p;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        if (T.kind[cd.c[i]] == D2D_KIND_VERTEX) continue;
        const float4 w0 = T.w0[cd.c[i]];
        const float g = G[i + 1].x * w0.z + G[i + 1].y * w0.w;
        optimizer_update(p, g, bc1, bc2, st.th[i], st.mu[i], st.nu[i]);
    }
    return loss;
}

template <int K>
__device__ __forceinline__ void adam_init(const SceneTab& T, const KParams& p, const Cand<K>& cd, const long long col,
                                          const int restart, AdamState<K>& st) {
    int u = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        st.mu[i] = 0.f;
        st.nu[i] = 0.f;
        st.th[i] = 0.f;
        if (T.kind[cd.c[i]] != D2D_KIND_VERTEX) {
            st.th[i] = p.x0 ? p.x0[(col * p.many + restart) * p.max_order + u] : 0.5f;
            ++u;
        }
    }
    st.b1p = 1.0f;
    st.b2p = 1.0f;
}

// Forward scan that keeps checkpoints.  ck[c] = state before step c * stride.  Returns losses[-1].
template <int METHOD, int K>
__device__ __forceinline__ float adam_scan_ckpt(const SceneTab& T, const KParams& p, const Cand<K>& cd, const float2 tx,
                                                const float2 rx, const long long col, const int restart,
                                                const int stride, AdamState<K> (&ck)[kNck], AdamState<K>& st) {
    adam_init<K>(T, p, cd, col, restart, st);
    float last = 0.f;
    int c = 0, next = 0;
    for (int s = 0; s < p.steps; ++s) {
        if (s == next) {
            ck[c++] = st;
            next += stride;
        }
        last = adam_step<METHOD, K>(T, p, cd, tx, rx, st);
    }
    return last;
}

// minimize_many_random_uniform (optimize.py:171-182): index of the restart whose final loss is the smallest
// (first one on ties, NaN wins: jnp.argmin).  argmin is piecewise constant, so the cotangent flows through the
// selected scan only; the selection itself is re-derived here by plain scans (nothing is stored by the forward).
template <int METHOD, int K>
__device__ __forceinline__ int best_restart(const SceneTab& T, const KParams& p, const Cand<K>& cd, const float2 tx,
                                            const float2 rx, const long long col) {
    int arg = 0;
    float best = 0.0f;
    for (int r = 0; r < p.many; ++r) {
        AdamState<K> st;
        adam_init<K>(T, p, cd, col, r, st);
        float last = 0.0f;
        for (int s = 0; s < p.steps; ++s) last = adam_step<METHOD, K>(T, p, cd, tx, rx, st);
        if (r == 0 || (last < best && best == best) || (last != last && best == best)) {
            best = last;
            arg = r;
        }
    }
    return arg;
}

__host__ __device__ inline int adam_ckpt_stride(int steps) {
    const int per = kBlk * kNck;
    return kBlk * ((steps + per - 1) / per > 0 ? (steps + per - 1) / per : 1);
}

// Reverse of the whole scan.  th_bar: cotangent of theta_S (in), loss_bar: cotangent of losses[-1] (MinPath's
// `loss`, optimize.py:96-97; 0 otherwise).  Accumulates into tx_bar, rx_bar and oa[].
template <int METHOD, int K>
__device__ __forceinline__ void adam_scan_reverse(const SceneTab& T, const KParams& p, const Cand<K>& cd, const float2 tx,
                                                  const float2 rx, const int stride, const AdamState<K> (&ck)[kNck],
                                                  const float (&th_bar_in)[K > 0 ? K : 1], const float loss_bar,
                                                  float2& tx_bar, float2& rx_bar, ObjAdj (&oa)[K > 0 ? K : 1]) {
    constexpr int KK = K > 0 ? K : 1;
    float lth[KK], lmu[KK], lnu[KK];
#pragma unroll
    for (int i = 0; i < KK; ++i) { lth[i] = th_bar_in[i]; lmu[i] = 0.f; lnu[i] = 0.f; }
    AdamState<K> blk[kBlk];
    const int S = p.steps;
    const int n_ck = (S + stride - 1) / stride;
    for (int c = n_ck - 1; c >= 0; --c) {
        const int seg0 = c * stride;
        const int seg1 = min(S, seg0 + stride);
        const int n_blk = (seg1 - seg0 + kBlk - 1) / kBlk;
        for (int b = n_blk - 1; b >= 0; --b) {
            const int b0 = seg0 + b * kBlk, b1 = min(seg1, b0 + kBlk);
            AdamState<K> st = ck[c];
            for (int s = seg0; s < b0; ++s) adam_step<METHOD, K>(T, p, cd, tx, rx, st);  // only when stride > kBlk
            for (int s = b0; s < b1; ++s) {
                blk[s - b0] = st;
                adam_step<METHOD, K>(T, p, cd, tx, rx, st);
            }
            for (int s = b1 - 1; s >= b0; --s) {
                const AdamState<K> pre = blk[s - b0];
                // recompute the step's intermediates
                float2 X[K + 2], G[K + 2];
                X[0] = tx;
                X[K + 1] = rx;
                place_points<K>(T, cd, pre.th, X);
                solver_loss_grad<METHOD, K>(T, cd, X, G);
                const float bc1 = 1.0f - pre.b1p * p.b1, bc2 = 1.0f - pre.b2p * p.b2;
                float lg[KK];
                bool any = false;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    lg[i] = 0.f;
                    if (T.kind[cd.c[i]] == D2D_KIND_VERTEX) continue;
                    const float4 w0 = T.w0[cd.c[i]];
                    const float g = G[i + 1].x * w0.z + G[i + 1].y * w0.w;
                    if (p.opt == D2D_OPT_SGD) {  // trace' = g + m trace ; theta' = theta - lr trace'
                        const float ltr_s = lmu[i] + (-p.lr) * lth[i];
                        lg[i] = ltr_s;
                        lmu[i] = p.b1 * ltr_s;
                        any = any || lg[i] != 0.f;
                        continue;
                    }
                    const float mu = (1.0f - p.b1) * g + p.b1 * pre.mu[i];
                    const float nu = (1.0f - p.b2) * (g * g) + p.b2 * pre.nu[i];
                    const float mh = mu / bc1, nh = nu / bc2;
                    const float sq = sqrtf(nh), den = sq + p.opt_eps;
                    const float lupd = -p.lr * lth[i];
                    const float lmh = lupd / den;
                    const float lden = -lupd * mh / (den * den);
                    const float lnh = sq > 0.f ? 0.5f * lden / sq : 0.f;  // clean: d sqrt(0) is a masked constant
                    const float lmu_s = lmu[i] + lmh / bc1;
                    const float lnu_s = lnu[i] + lnh / bc2;
                    lg[i] = (1.0f - p.b1) * lmu_s + (1.0f - p.b2) * 2.0f * g * lnu_s;
                    lmu[i] = p.b1 * lmu_s;
                    lnu[i] = p.b2 * lnu_s;
                    any = any || lg[i] != 0.f;
                }
                if (METHOD == D2D_METHOD_MINPATH && s == S - 1 && loss_bar != 0.f) {
                    // losses[-1] = loss_fun(theta_{S-1})
                    float thb[KK];
                    V2<float> txb, rxb;
                    ObjAdjS<float> ob[KK];
                    loss_grad_full<METHOD, K, float>(T, cd, pre.th, tx, rx, loss_bar, thb, txb, rxb, ob);
                    tx_bar.x += txb.x; tx_bar.y += txb.y;
                    rx_bar.x += rxb.x; rx_bar.y += rxb.y;
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        lth[i] += thb[i];
                        oa[i].p1.x += ob[i].p1.x; oa[i].p1.y += ob[i].p1.y;
                        oa[i].t.x += ob[i].t.x; oa[i].t.y += ob[i].t.y;
                        oa[i].n.x += ob[i].n.x; oa[i].n.y += ob[i].n.y;
                        oa[i].phi += ob[i].phi;
                    }
                }
                if (any) {
                    Dual th[KK], thb[KK];
#pragma unroll
                    for (int i = 0; i < KK; ++i) th[i] = mk(pre.th[i], i < K ? lg[i] : 0.f);
                    V2<Dual> txb, rxb;
                    ObjAdjS<Dual> ob[KK];
                    loss_grad_full<METHOD, K, Dual>(T, cd, th, tx, rx, 1.0f, thb, txb, rxb, ob);
                    tx_bar.x += txb.x.d; tx_bar.y += txb.y.d;
                    rx_bar.x += rxb.x.d; rx_bar.y += rxb.y.d;
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        lth[i] += thb[i].d;
                        oa[i].p1.x += ob[i].p1.x.d; oa[i].p1.y += ob[i].p1.y.d;
                        oa[i].t.x += ob[i].t.x.d; oa[i].t.y += ob[i].t.y.d;
                        oa[i].n.x += ob[i].n.x.d; oa[i].n.y += ob[i].n.y.d;
                        oa[i].phi += ob[i].phi.d;
                    }
                }
            }
        }
    }
}

}  // namespace d2d
