// differt2d_b200 — reverse-mode (VJP) kernel, by recomputation.
//
// Same mapping as the forward kernel (one thread per grid point, candidates walked in list order
// by the whole CTA).  Nothing is stored by the forward pass: each path is re-traced, and only paths
// whose validity is non-zero run the reverse sweep.  In smooth mode the validity is
// min(act(onx), 1 - act(interx), act(lx)) and every fold is a min/max, so the cotangent follows a
// single chain: the arg-min comparison of on_objects, OR the arg-max (segment, object) occlusion
// test, OR the loss — the reverse sweep of the occlusion double loop costs one test, not (k+1)N.
//
// Per-grid-point cotangents are written directly (disjoint).  Scene-parameter cotangents are
// reduced warp (shuffle) -> CTA (shared-memory atomics) -> device (global fp32 atomics); the
// multi-GPU all-reduce of these few KB happens in the host layer (differt2d_b200/distributed.py).
#include "d2d_driver.cuh"
#include "d2d_image_bwd.cuh"
#include "d2d_launch.h"
#include "d2d_solver.cuh"
#include "d2d_solver_adj.cuh"
#include "d2d_newton.cuh"

namespace d2d {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums N (a power of two <= 32) per-lane values over the warp in one butterfly: at every stage a lane keeps one half of
// its values, hands the other half to its partner and adds what it receives, so N values cost N - 1 + (5 - log2 N)
// shuffles instead of 5 N (8 values: 9 instead of 40).  Returns, in EVERY lane l, the warp total of value number
// l >> (5 - log2 N).  Must be called by the whole warp.
template <int N>
__device__ __forceinline__ float warp_sum_many(float (&v)[N]) {
    static_assert(N >= 1 && N <= 32 && (N & (N - 1)) == 0, "N must be a power of two");
    const int lane = threadIdx.x & 31;
    int stride = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const bool upper = (lane & stride) != 0;
#pragma unroll
        for (int k = 0; k < n / 2; ++k) {
            const float send = upper ? v[k] : v[k + n / 2];
            const float keep = upper ? v[k + n / 2] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, stride);
        }
        stride >>= 1;
    }
    float r = v[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o <= stride) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

// Reverse sweep of one path (ImagePath: closed form; FermatPath / MinPath: through the Adam scan, see
// d2d_solver_adj.cuh).  Returns valid * fun; when the path carries gradient, `has` is set and tx_bar / rx_bar /
// alpha_bar / oa[] / occ_* are filled (not accumulated).
template <int MODE, int METHOD, int K>
__device__ __noinline__ float path_vjp(const SceneTab& T, const KParams& p, const float alpha,
                                          const Cand<K>& cd, const float2 tx, const float2 rx, const long long col,
                                          const float zbar, bool& has, float2& tx_bar, float2& rx_bar,
                                          float& alpha_bar, ObjAdj (&oa)[K > 0 ? K : 1], int& occ_j,
                                          float4& occ_bar) {
    constexpr bool kSolver = (METHOD != D2D_METHOD_IMAGE) && K > 0;
    has = false;
    occ_j = -1;
    // ---- recompute the forward, keeping what the reverse sweep needs --------------------------
    float2 X[K + 2];
    float2 I[K + 1];
    X[0] = tx;
    X[K + 1] = rx;
    I[0] = tx;
    AdamState<K> ck[kSolver ? kNck : 1];
    AdamState<K> th_final;
    float solver_loss = 0.f;
    const int stride = adam_ckpt_stride(p.steps);
    if constexpr (kSolver) {
        if (p.opt == D2D_OPT_NEWTON) {  // the forward's Newton iterations again; differentiated implicitly below
            adam_init<K>(T, p, cd, col, 0, th_final);
            solver_loss = newton_solve<METHOD, K>(T, p, cd, tx, rx, th_final.th);
        } else {
            const int restart = p.many > 1 ? best_restart<METHOD, K>(T, p, cd, tx, rx, col) : 0;
            solver_loss = adam_scan_ckpt<METHOD, K>(T, p, cd, tx, rx, col, restart, stride, ck, th_final);
        }
        place_points<K>(T, cd, th_final.th, X);
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) I[i + 1] = mirror(I[i], T.w0[cd.c[i]], T.w1[cd.c[i]]);
        float2 q = rx;
#pragma unroll
        for (int i = K - 1; i >= 0; --i) {
            q = back_project(q, I[i + 1], T.w0[cd.c[i]], T.w1[cd.c[i]]);
            X[i + 1] = q;
        }
    }
    float onx = CUDART_INF_F;
    int on_i = -1;
    float on_s = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int j = cd.c[i];
        if (T.kind[j] == D2D_KIND_VERTEX) continue;
        const float s = to_parametric(X[i + 1], T.w0[j], T.w1[j]);
        float x = fminf(s, 1.0f - s);
        if (s != s) x = -CUDART_INF_F;
        if (x < onx) { onx = x; on_i = i; on_s = s; }
    }
    float a_on = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(onx >= 0.0f)) return 0.0f;
    } else if (onx != CUDART_INF_F) {
        a_on = act<MODE>(onx, alpha);
        if (a_on == 0.0f) return 0.0f;
    }
    // FermatPath's loss is path_loss(xys) (geometry.py:1202-1204), MinPath's is the scan's last loss (:1286-1288)
    const float loss = (kSolver && METHOD == D2D_METHOD_MINPATH) ? solver_loss : path_loss<K>(T, cd, X);
    const float lx = p.tol - loss;
    float a_l = 1.0f;
    if (MODE == D2D_MODE_HARD) {
        if (!(lx > 0.0f)) return 0.0f;
    } else {
        if (lx != lx) return 0.0f;
        a_l = act<MODE>(lx, alpha);
        if (a_l == 0.0f) return 0.0f;
    }
    bool alive = true;
    int seg = -1, jj = -1;
    const float interx = intersects_x<MODE, K, true>(T, p.N, cd, X, alpha, x_zero<MODE>(alpha), alive, seg, jj);
    if (!alive) return 0.0f;
    float valid = 1.0f, a_in = 0.0f;
    if (MODE != D2D_MODE_HARD) {
        a_in = (interx == -CUDART_INF_F) ? 0.0f : act<MODE>(interx, alpha);
        valid = fminf(fminf(a_on, 1.0f - a_in), a_l);
        if (valid == 0.0f) return 0.0f;
    }
    float r;
    const float val = path_value<K>(p, X, r);
    const float contrib = valid * val;
    if (zbar == 0.0f) return contrib;
    has = true;

    // ---- reverse sweep ----------------------------------------------------------------------------
    float2 Xb[K + 2];
#pragma unroll
    for (int i = 0; i < K + 2; ++i) Xb[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < K; ++i) oa[i].zero();
    alpha_bar = 0.f;
    float solver_loss_bar = 0.f;

    // (1) fun(path) : utils.py:52-54 / length**2 ; path_length geometry.py:199-203
    {
        const float val_bar = zbar * valid;
        float r_bar;
        if (p.fun == D2D_FUN_RECEIVED_POWER) r_bar = -val_bar * val * 2.0f * r / (p.h2 + r * r);
        else r_bar = val_bar * 2.0f * r;
#pragma unroll
        for (int i = 0; i <= K; ++i) {
            const float dx = (X[i + 1].x - X[i].x) + kEps32;
            const float dy = (X[i + 1].y - X[i].y) + kEps32;
            const float len = sqrtf(dx * dx + dy * dy);
            if (len > 0.0f) {
                const float c = r_bar / len;
                Xb[i + 1].x += c * dx; Xb[i + 1].y += c * dy;
                Xb[i].x -= c * dx; Xb[i].y -= c * dy;
            }
        }
    }

    // (2) validity (smooth logic only) : geometry.py:947-963, logic.py:511-512
    if (MODE != D2D_MODE_HARD) {
        const float v1 = a_on, v2 = 1.0f - a_in, v3 = a_l;
        const int cnt = (v1 == valid) + (v2 == valid) + (v3 == valid);
        const float share = zbar * val / (float)cnt;  // jnp.min splits evenly among ties
        if (v1 == valid && on_i >= 0) {
            // contains_parametric = minimum(act(s - 0), act(1 - s)) (geometry.py:608-621): the tie rule of
            // jnp.minimum (1/2, 1/2) applies to the ACTIVATED values, which tie far more often than the
            // pre-activations do (the activation is not injective in fp32).
            const float xg = on_s, xl = 1.0f - on_s;
            const float Ag = act<MODE>(xg, alpha), Al = act<MODE>(xl, alpha);
            const float wg = Ag < Al ? 1.0f : (Ag == Al ? 0.5f : 0.0f), wl = 1.0f - wg;
            const float dg = act_dz<MODE>(xg, alpha), dl = act_dz<MODE>(xl, alpha);
            const float s_bar = share * alpha * (wg * dg - wl * dl);
            alpha_bar += share * (wg * xg * dg + wl * xl * dl);
            if (s_bar != 0.f) {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (i != on_i) continue;
                    const float4 w0 = T.w0[cd.c[i]];
                    const float4 w1 = T.w1[cd.c[i]];
                    const float wx = X[i + 1].x - w0.x, wy = X[i + 1].y - w0.y;
                    const float k = s_bar / w1.z;
                    Xb[i + 1].x += k * w0.z; Xb[i + 1].y += k * w0.w;
                    oa[i].p1.x -= k * w0.z; oa[i].p1.y -= k * w0.w;
                    oa[i].t.x += k * wx; oa[i].t.y += k * wy;
                    oa[i].tt += -k * on_s;
                }
            }
        }
        if (v3 == valid) {
            const float dz = act_dz<MODE>(lx, alpha);
            const float loss_bar = -share * alpha * dz;
            alpha_bar += share * lx * dz;
            if (kSolver && METHOD == D2D_METHOD_MINPATH) {
                solver_loss_bar = loss_bar;  // flows into loss_fun(theta_{S-1}), inside the scan
            } else if (loss_bar != 0.f) {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const int j = cd.c[i];
                    residual_adj(T.kind[j], X[i], X[i + 1], X[i + 2], T.w1[j], T.sc[j], loss_bar, Xb[i], Xb[i + 1],
                                 Xb[i + 2], oa[i].n, oa[i].phi);
                }
            }
        }
        if (v2 == valid && jj >= 0 && interx != -CUDART_INF_F) {
            const float gsh = -share;  // d valid / d a_in = -1
            {
                float2 P = X[0], Q = X[1];
#pragma unroll
                for (int i = 0; i <= K; ++i)
                    if (i == seg) { P = X[i]; Q = X[i + 1]; }
                const float4 w = T.w2[jj];
                const float2 B = make_float2(P.x - Q.x, P.y - Q.y);
                const float Cx = w.x - P.x, Cy = w.y - P.y;
                const float a = B.y * Cx - B.x * Cy;
                const float b = w.z * Cy - w.w * Cx;
                const float d = w.w * B.x - w.z * B.y;
                const float ta = a / d, tb = b / d;
                // hit = minimum(minimum(ge_a, le_a), minimum(ge_b, le_b)) on ACTIVATED values (geometry.py:167-173)
                const float x1 = ta + kTolSeg, x2 = kHiSeg - ta, x3 = tb + kTolSeg, x4 = kHiSeg - tb;
                const float A1 = act<MODE>(x1, alpha), A2 = act<MODE>(x2, alpha);
                const float A3 = act<MODE>(x3, alpha), A4 = act<MODE>(x4, alpha);
                const float Ta = fminf(A1, A2), Tb = fminf(A3, A4);
                const float wa = Ta < Tb ? 1.0f : (Ta == Tb ? 0.5f : 0.0f), wb = 1.0f - wa;
                const float w1 = wa * (A1 < A2 ? 1.0f : (A1 == A2 ? 0.5f : 0.0f)), w2 = wa - w1;
                const float w3 = wb * (A3 < A4 ? 1.0f : (A3 == A4 ? 0.5f : 0.0f)), w4 = wb - w3;
                const float d1 = act_dz<MODE>(x1, alpha), d2 = act_dz<MODE>(x2, alpha);
                const float d3 = act_dz<MODE>(x3, alpha), d4 = act_dz<MODE>(x4, alpha);
                alpha_bar += gsh * (w1 * x1 * d1 + w2 * x2 * d2 + w3 * x3 * d3 + w4 * x4 * d4);
                const float ta_bar = gsh * alpha * (w1 * d1 - w2 * d2);
                const float tb_bar = gsh * alpha * (w3 * d3 - w4 * d4);
                const float a_bar = ta_bar / d, b_bar = tb_bar / d;
                const float d_bar = -(ta_bar * ta + tb_bar * tb) / d;
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
                    if (i == seg) {
                        Xb[i].x += Pb.x; Xb[i].y += Pb.y;
                        Xb[i + 1].x -= Bb.x; Xb[i + 1].y -= Bb.y;
                    }
                // P1' = (1+patch) P1 - patch P2 ; P2' = (1+patch) P2 - patch P1 ; A = P2' - P1' ; C = P1' - P
                const float2 p1p = make_float2(Cb.x - Ab.x, Cb.y - Ab.y);
                const float2 p2p = Ab;
                const float q = p.patch;
                occ_j = jj;
                occ_bar = make_float4((1.0f + q) * p1p.x - q * p2p.x, (1.0f + q) * p1p.y - q * p2p.y,
                                      (1.0f + q) * p2p.x - q * p1p.x, (1.0f + q) * p2p.y - q * p1p.y);
            }
        }
    }

    if constexpr (kSolver) {
        // (3') xys = parametric_to_cartesian(objects, theta_S) (geometry.py:1200, 1284), then the Adam scan
        constexpr int KK = K > 0 ? K : 1;
        float thb[KK];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float4 w0 = T.w0[cd.c[i]];
            oa[i].p1.x += Xb[i + 1].x; oa[i].p1.y += Xb[i + 1].y;
            thb[i] = 0.f;
            if (T.kind[cd.c[i]] != D2D_KIND_VERTEX) {
                thb[i] = Xb[i + 1].x * w0.z + Xb[i + 1].y * w0.w;
                oa[i].t.x += th_final.th[i] * Xb[i + 1].x;
                oa[i].t.y += th_final.th[i] * Xb[i + 1].y;
            }
        }
        tx_bar = Xb[0];
        rx_bar = Xb[K + 1];
        if (p.opt == D2D_OPT_NEWTON)
            newton_reverse<METHOD, K>(T, cd, th_final.th, tx, rx, thb, solver_loss_bar, tx_bar, rx_bar, oa);
        else
            adam_scan_reverse<METHOD, K>(T, p, cd, tx, rx, stride, ck, thb, solver_loss_bar, tx_bar, rx_bar, oa);
        return contrib;
    }
    // (3) backward scan of the image method : geometry.py:1093-1107 (clean `where`)
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
        const float un = ux * w1.x + uy * w1.y;
        const float vn = vx * w1.x + vy * w1.y;
        Xb[i + 2].x += qb.x; Xb[i + 2].y += qb.y;
        if (un != 0.0f) {
            const float g = vn / un;
            float ubx = g * qb.x, uby = g * qb.y;
            const float gb = qb.x * ux + qb.y * uy;
            const float vnb = gb / un;
            const float unb = -gb * g / un;
            ubx += unb * w1.x; uby += unb * w1.y;
            oa[i].n.x += unb * ux + vnb * vx;
            oa[i].n.y += unb * uy + vnb * vy;
            const float vbx = vnb * w1.x, vby = vnb * w1.y;
            Xb[i + 2].x += ubx - vbx; Xb[i + 2].y += uby - vby;
            Ib[i + 1].x -= ubx; Ib[i + 1].y -= uby;
            oa[i].p1.x += vbx; oa[i].p1.y += vby;
        }
    }
    // (4) forward scan (mirror chain) : geometry.py:1086-1091, 652-670
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        const float4 w0 = T.w0[cd.c[i]];
        const float4 w1 = T.w1[cd.c[i]];
        const float wx = I[i].x - w0.x, wy = I[i].y - w0.y;
        const float cc = 2.0f * (wx * w1.x + wy * w1.y);
        const float2 ib = Ib[i + 1];
        const float ccb = -(ib.x * w1.x + ib.y * w1.y);
        const float dotb = 2.0f * ccb;
        oa[i].n.x += -cc * ib.x + dotb * wx;
        oa[i].n.y += -cc * ib.y + dotb * wy;
        const float wbx = dotb * w1.x, wby = dotb * w1.y;
        Ib[i].x += ib.x + wbx; Ib[i].y += ib.y + wby;
        oa[i].p1.x -= wbx; oa[i].p1.y -= wby;
    }
    tx_bar = make_float2(Xb[0].x + Ib[0].x, Xb[0].y + Ib[0].y);
    rx_bar = Xb[K + 1];
    return contrib;
}

struct BwdAcc {
    float2 grid_bar;   // this thread's grid point, current fixed point
    float2 fixed_bar;  // current fixed point (to be reduced over the grid)
    float alpha_bar;
    float acc;
};

template <int MODE, int METHOD, int K, bool TXGRID>
__device__ __forceinline__ void run_order_bwd(const SceneTab& T, const KParams& p, const Tile& tile, DriverShared& sh,
                                              const float alpha, const int t, const float2 fx, const float2 g,
                                              const long long col0, int& buf, const float zbar, BwdAcc& A, float* s_obj,
                                              float* s_phi, const uint32_t* mread) {
    constexpr int KK = K > 0 ? K : 1;
    const float2 tx = TXGRID ? g : fx;
    const float2 rx = TXGRID ? fx : g;
    for_each_candidate<MODE, METHOD, K, TXGRID>(
        T, p, tile, sh, alpha, t, fx, col0, mread, buf, [&](const Cand<K>& cd, const long long col, const float2 apex) {
            bool has = false;
            float2 txb = make_float2(0.f, 0.f), rxb = make_float2(0.f, 0.f);
            float ab = 0.f;
            ObjAdj oa[KK];
            int occ_j = -1;
            float4 occ_bar = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tile.active) {
                if constexpr (METHOD == D2D_METHOD_IMAGE) {
                    // ONE tracked re-trace (the forward kernel's arithmetic, plus the arg-min / arg-max bookkeeping of
                    // the reverse sweep), then the sweep from the live registers (d2d_image_bwd.cuh)
                    ImageTrace<K> tr;
                    const float2 ap = TXGRID ? image_apex<K>(T, cd, tx) : apex;
                    if (trace_image_tracked<MODE, K>(T, p, alpha, cd, tx, rx, ap, tr, &sh.hint[threadIdx.x >> 5])) {
                        A.acc = A.acc + tr.valid * tr.val;
                        if (zbar != 0.0f) {
                            has = true;
                            image_reverse<MODE, K>(T, p, alpha, cd, tr, zbar, txb, rxb, ab, oa, occ_j, occ_bar);
                        }
                    }
                } else {
                    // light re-trace first (same code as the forward kernel); the reverse sweep through the Adam scan
                    // is out of line and only runs for the paths whose validity is non-zero
                    float2 X[K + 2];
                    float loss;
                    construct_path<METHOD, K>(T, p, cd, tx, rx, col, X, loss);
                    const float valid = validity<MODE, K, (METHOD != D2D_METHOD_MINPATH) || K == 0>(
                        T, p, alpha, cd, X, loss, &sh.hint[threadIdx.x >> 5]);
                    if (valid != 0.0f) {
                        const float c = path_vjp<MODE, METHOD, K>(T, p, alpha, cd, tx, rx, col, zbar, has, txb, rxb, ab,
                                                                  oa, occ_j, occ_bar);
                        A.acc = A.acc + c;
                    }
                }
            }
            if (has) {
                const float2 gb = TXGRID ? txb : rxb;
                const float2 fb = TXGRID ? rxb : txb;
                A.grid_bar.x += gb.x; A.grid_bar.y += gb.y;
                A.fixed_bar.x += fb.x; A.fixed_bar.y += fb.y;
                A.alpha_bar += ab;
            }
            if (s_obj) {
                // occluders' vertex cotangents: the lanes of a warp mostly share their arg-max occluder (neighbouring
                // receivers, same candidate), and 32 shared-memory atomics on one address serialise: one warp reduction
                // per DISTINCT occluder instead (uniform over the warp: the visit is)
                unsigned pend = __ballot_sync(0xffffffffu, has && occ_j >= 0);
                while (pend) {
                    const int j0 = __shfl_sync(0xffffffffu, occ_j, __ffs(pend) - 1);
                    const bool mine = has && occ_j == j0;
                    float v[4] = {mine ? occ_bar.x : 0.f, mine ? occ_bar.y : 0.f, mine ? occ_bar.z : 0.f,
                                  mine ? occ_bar.w : 0.f};
                    const float tot = warp_sum_many<4>(v);  // lane l: component l >> 3
                    if ((threadIdx.x & 7) == 0) atomicAdd(&s_obj[4 * j0 + ((threadIdx.x & 31) >> 3)], tot);
                    pend &= ~__ballot_sync(0xffffffffu, mine);
                }
            }
            if (K > 0 && s_obj && __any_sync(0xffffffffu, has)) {
                // every lane holds the same candidate: the 4 K vertex cotangents of its interacting objects go through
                // ONE butterfly (warp_sum_many) and one shared-memory atomic instruction (this epilogue was 350 of the
                // 1600 instructions of a visit on the un-prunable leg: 5 shuffles per value, 8 values at order 2)
                constexpr int NV = K <= 1 ? 4 : (K == 2 ? 8 : 16);
                float v[NV];
#pragma unroll
                for (int q = 0; q < NV; ++q) v[q] = 0.f;
                bool any_ris = false;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (has) {
                        const float4 u = oa[i].to_vertices(T.w0[cd.c[i]], T.w1[cd.c[i]]);
                        v[4 * i + 0] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
                    }
                    any_ris = any_ris || T.kind[cd.c[i]] == D2D_KIND_RIS;
                }
                const float tot = warp_sum_many<NV>(v);
                constexpr int kShift = NV == 4 ? 3 : (NV == 8 ? 2 : 1);  // lane l holds value l >> kShift
                const int lane = threadIdx.x & 31, q = lane >> kShift;
                if ((lane & ((1 << kShift) - 1)) == 0 && q < 4 * K) {
                    int obj = cd.c[0];
#pragma unroll
                    for (int i = 1; i < K; ++i)
                        if ((q >> 2) == i) obj = cd.c[i];
                    atomicAdd(&s_obj[4 * obj + (q & 3)], tot);
                }
                if (any_ris) {  // (uniform over the warp)
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        if (T.kind[cd.c[i]] != D2D_KIND_RIS) continue;
                        const float ph = warp_sum(has ? oa[i].phi : 0.f);
                        if (lane == 0) atomicAdd(&s_phi[cd.c[i]], ph);
                    }
                }
            }
        });
}

template <int MODE, int METHOD, bool TXGRID>
__global__ void __launch_bounds__(kBlock, D2D_BWD_MIN_CTAS) power_bwd_kernel(const KParams p, const float* __restrict__ Zbar,
                                                           const BwdOut out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ DriverShared sh;
    __shared__ float s_red[4][4];
    SceneTab T = carve_tab(smem, p.N);
    const bool want_obj = out.objects_bar != nullptr || out.phis_bar != nullptr;
    float* s_obj = nullptr;
    float* s_phi = nullptr;
    if (want_obj) {
        s_obj = reinterpret_cast<float*>(smem + ((scene_tab_bytes(p.N) + 15) / 16) * 16);
        s_phi = s_obj + 4 * p.N;
        for (int j = threadIdx.x; j < 5 * p.N; j += blockDim.x) s_obj[j] = 0.f;
    }
    build_tab(T, p, &sh.count);
    const Tile tile = make_tile(p, T, sh);
    const long long r = tile.r;
    const bool active = tile.active;
    const float alpha = p.alpha_dev ? *p.alpha_dev : p.alpha;
    if (MODE != D2D_MODE_HARD && !(alpha > 0.0f)) {  // see power_fwd_kernel: every output becomes NaN
        if (active && blockIdx.y == 0)
            for (int t = 0; t < (p.reduce_all ? 1 : p.T); ++t) {
                if (out.Z) out.Z[(long long)t * p.R + r] = CUDART_NAN_F;
                if (out.grid_bar) reinterpret_cast<float2*>(out.grid_bar)[(long long)t * p.R + r] = make_float2(CUDART_NAN_F, CUDART_NAN_F);
            }
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            for (int j = threadIdx.x; j < 4 * p.N; j += blockDim.x) if (out.objects_bar) out.objects_bar[j] = CUDART_NAN_F;
            for (int j = threadIdx.x; j < p.N; j += blockDim.x) if (out.phis_bar) out.phis_bar[j] = CUDART_NAN_F;
            for (int j = threadIdx.x; j < 2 * p.T; j += blockDim.x) if (out.fixed_bar) out.fixed_bar[j] = CUDART_NAN_F;
            if (threadIdx.x == 0 && out.alpha_bar) out.alpha_bar[0] = CUDART_NAN_F;
        }
        return;
    }
    if (D2D_FOLD_SKIP) T.fold_skip = fold_skip_bound<MODE>(alpha);
    if constexpr (METHOD == D2D_METHOD_IMAGE) {
        if (p.macro) macro_prologue<MODE, TXGRID>(T, p, tile, sh, alpha);
    }
    if (p.mask && mask_bitmap_fits(p)) mask_prologue(p, sh);
    const float2 g = active ? reinterpret_cast<const float2*>(p.grid)[r] : make_float2(0.f, 0.f);
    float zsum = 0.0f;
    float2 gsum = make_float2(0.f, 0.f);
    float alpha_total = 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int buf = 0;
    for (int t = 0; t < p.T; ++t) {
        const float2 fx = reinterpret_cast<const float2*>(p.fixed)[t];
        float zbar = 0.f;
        if (active) zbar = Zbar ? Zbar[p.reduce_all ? r : (long long)t * p.R + r] : 1.0f;
        const uint32_t* mread =
            p.mask ? p.mask + ((long long)t * gridDim.x * (kBlock / 32) + (long long)blockIdx.x * (kBlock / 32)) * p.mask_wpw
                   : nullptr;
        BwdAcc A;
        A.grid_bar = A.fixed_bar = make_float2(0.f, 0.f);
        A.alpha_bar = 0.f;
        A.acc = 0.f;
        long long col0 = 0;
        for (int k = p.min_order; k <= p.max_order; ++k) {
            switch (k) {
                case 0: run_order_bwd<MODE, METHOD, 0, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, zbar, A, s_obj, s_phi, mread); break;
                case 1: run_order_bwd<MODE, METHOD, 1, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, zbar, A, s_obj, s_phi, mread); break;
                case 2: run_order_bwd<MODE, METHOD, 2, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, zbar, A, s_obj, s_phi, mread); break;
                case 3: run_order_bwd<MODE, METHOD, 3, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, zbar, A, s_obj, s_phi, mread); break;
                case 4: run_order_bwd<MODE, METHOD, 4, TXGRID>(T, p, tile, sh, alpha, t, fx, g, col0, buf, zbar, A, s_obj, s_phi, mread); break;
                default: break;
            }
            col0 += order_count(k, T.n_allowed);
        }
        if (active) {
            if (p.reduce_all) {
                zsum = zsum + A.acc;
                gsum.x += A.grid_bar.x; gsum.y += A.grid_bar.y;
            } else if (gridDim.y == 1) {
                if (out.Z) out.Z[(long long)t * p.R + r] = A.acc;
                if (out.grid_bar) reinterpret_cast<float2*>(out.grid_bar)[(long long)t * p.R + r] = A.grid_bar;
            } else {  // candidate slices: outputs zeroed by the launcher
                if (out.Z && A.acc != 0.f) atomicAdd(&out.Z[(long long)t * p.R + r], A.acc);
                if (out.grid_bar) {
                    if (A.grid_bar.x != 0.f) atomicAdd(&out.grid_bar[2 * ((long long)t * p.R + r) + 0], A.grid_bar.x);
                    if (A.grid_bar.y != 0.f) atomicAdd(&out.grid_bar[2 * ((long long)t * p.R + r) + 1], A.grid_bar.y);
                }
            }
        }
        alpha_total += A.alpha_bar;
        if (out.fixed_bar) {  // CTA reduction of this fixed point's cotangent
            const float fxs = warp_sum(A.fixed_bar.x), fys = warp_sum(A.fixed_bar.y);
            __syncthreads();
            if (lane == 0) { s_red[warp][0] = fxs; s_red[warp][1] = fys; }
            __syncthreads();
            if (threadIdx.x == 0) {
                float sx = 0.f, sy = 0.f;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sx += s_red[w][0]; sy += s_red[w][1]; }
                if (sx != 0.f) atomicAdd(&out.fixed_bar[2 * t + 0], sx);
                if (sy != 0.f) atomicAdd(&out.fixed_bar[2 * t + 1], sy);
            }
        }
    }
    if (active && p.reduce_all) {
        if (gridDim.y == 1) {
            if (out.Z) out.Z[r] = zsum;
            if (out.grid_bar) reinterpret_cast<float2*>(out.grid_bar)[r] = gsum;
        } else {
            if (out.Z && zsum != 0.f) atomicAdd(&out.Z[r], zsum);
            if (out.grid_bar) {
                if (gsum.x != 0.f) atomicAdd(&out.grid_bar[2 * r + 0], gsum.x);
                if (gsum.y != 0.f) atomicAdd(&out.grid_bar[2 * r + 1], gsum.y);
            }
        }
    }
    if (out.alpha_bar) {
        const float as = warp_sum(alpha_total);
        __syncthreads();
        if (lane == 0) s_red[warp][2] = as;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_red[w][2];
            if (s != 0.f) atomicAdd(out.alpha_bar, s);
        }
    }
    if (want_obj) {
        __syncthreads();
        for (int j = threadIdx.x; j < 4 * p.N; j += blockDim.x)
            if (out.objects_bar && s_obj[j] != 0.f) atomicAdd(&out.objects_bar[j], s_obj[j]);
        for (int j = threadIdx.x; j < p.N; j += blockDim.x)
            if (out.phis_bar && s_phi[j] != 0.f) atomicAdd(&out.phis_bar[j], s_phi[j]);
    }
}

template <int MODE, int METHOD, bool TXGRID>
static int launch_bwd_one(const KParams& p, const float* Zbar, const BwdOut& out, cudaStream_t stream) {
    size_t smem = ((scene_tab_bytes(p.N) + 15) / 16) * 16;
    if (out.objects_bar || out.phis_bar) smem += (size_t)5 * p.N * sizeof(float);
    auto kern = power_bwd_kernel<MODE, METHOD, TXGRID>;
    if (smem > 32 * 1024) {  // static + dynamic > 48 KB needs the opt-in; the static part is ~11 KB
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    // without an activity mask the backward re-runs the culls: clusters + macro-tile cull as in the forward kernel
    KParams q = p;
    q.macro = (host_macro_ok(p) && !p.mask && METHOD == D2D_METHOD_IMAGE) ? 1 : 0;  // both grid roles
    const cudaError_t e = launch_tiles(kern, q, q.macro != 0, smem, stream, q, Zbar, out);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

#ifndef D2D_TU_MODE
#error "compile with -DD2D_TU_MODE=<D2D_MODE_*> (differt2d_b200/build.py)"
#endif
#ifndef D2D_TU_SOLVER
#define D2D_TU_SOLVER 0  // 0: ImagePath kernels; 1: FermatPath / MinPath kernels (separate translation unit)
#endif

template <int MODE, int METHOD>
static int launch_bwd_role(const KParams& p, int grid_role, const float* Zbar, const BwdOut& out, cudaStream_t stream) {
    return grid_role == D2D_GRID_TRANSMITTERS ? launch_bwd_one<MODE, METHOD, true>(p, Zbar, out, stream)
                                              : launch_bwd_one<MODE, METHOD, false>(p, Zbar, out, stream);
}

#if D2D_TU_SOLVER
template <>
int launch_bwd_solver_mode<D2D_TU_MODE>(const KParams& p, int grid_role, int method, const float* Zbar,
                                        const BwdOut& out, cudaStream_t stream) {
    if (method == D2D_METHOD_FERMAT) return launch_bwd_role<D2D_TU_MODE, D2D_METHOD_FERMAT>(p, grid_role, Zbar, out, stream);
    if (method == D2D_METHOD_MINPATH) return launch_bwd_role<D2D_TU_MODE, D2D_METHOD_MINPATH>(p, grid_role, Zbar, out, stream);
    return (int)cudaErrorInvalidValue;
}
#else
template <>
int launch_bwd_mode<D2D_TU_MODE>(const KParams& p, int grid_role, int method, const float* Zbar, const BwdOut& out,
                                 cudaStream_t stream) {
    if (method != D2D_METHOD_IMAGE) return launch_bwd_solver_mode<D2D_TU_MODE>(p, grid_role, method, Zbar, out, stream);
    return launch_bwd_role<D2D_TU_MODE, D2D_METHOD_IMAGE>(p, grid_role, Zbar, out, stream);
}

#if D2D_TU_MODE == D2D_MODE_HARD
int launch_power_bwd(const KParams& p, int mode, int grid_role, int method, const float* Zbar, const BwdOut& out,
                     cudaStream_t stream, long long* launches) {
    cudaError_t e;
    if (out.objects_bar && (e = cudaMemsetAsync(out.objects_bar, 0, sizeof(float) * 4 * p.N, stream)) != cudaSuccess) return (int)e;
    if (out.phis_bar && (e = cudaMemsetAsync(out.phis_bar, 0, sizeof(float) * p.N, stream)) != cudaSuccess) return (int)e;
    if (out.fixed_bar && (e = cudaMemsetAsync(out.fixed_bar, 0, sizeof(float) * 2 * p.T, stream)) != cudaSuccess) return (int)e;
    if (out.alpha_bar && (e = cudaMemsetAsync(out.alpha_bar, 0, sizeof(float), stream)) != cudaSuccess) return (int)e;
    if (p.R <= 0) return 0;
    if (p.slices > 1) {
        const size_t nz = (size_t)(p.reduce_all ? 1 : p.T) * p.R;
        if (out.Z && (e = cudaMemsetAsync(out.Z, 0, sizeof(float) * nz, stream)) != cudaSuccess) return (int)e;
        if (out.grid_bar && (e = cudaMemsetAsync(out.grid_bar, 0, sizeof(float) * 2 * nz, stream)) != cudaSuccess) return (int)e;
    }
    int rc;
    switch (mode) {
        case D2D_MODE_HARD: rc = launch_bwd_mode<D2D_MODE_HARD>(p, grid_role, method, Zbar, out, stream); break;
        case D2D_MODE_HARD_SIGMOID: rc = launch_bwd_mode<D2D_MODE_HARD_SIGMOID>(p, grid_role, method, Zbar, out, stream); break;
        case D2D_MODE_SIGMOID: rc = launch_bwd_mode<D2D_MODE_SIGMOID>(p, grid_role, method, Zbar, out, stream); break;
        default: return (int)cudaErrorInvalidValue;
    }
    if (launches) *launches += 1;
    return rc;
}
#endif
#endif

}  // namespace d2d
