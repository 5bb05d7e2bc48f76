// differt2d_b200 — hand-written reverse-mode pieces (VJPs) of the path geometry.
//
// Gradient semantics are the "clean" ones (DESIGN.md "NaN semantics"): a branch that the forward
// masks with a `where` (zero-length normalisation, un == 0 in the image back-projection, parallel
// segments) is a constant with zero cotangent, instead of the NaN the reference's single-`where`
// produces (geometry.py:1105, :227-230).  None of this feeds a hard predicate, so FMA contraction
// is harmless here; with -fmad=false the compiler will not contract, fmaf() is used where it pays.
#pragma once

#include "d2d_device.cuh"

namespace d2d {

// plain (uncontracted) on purpose: the loss cotangent is proportional to the residual vector e, which is
// pure rounding noise for specular paths; recomputing it canonically keeps it equal to the forward's.
__device__ __forceinline__ float dot2(const float2 a, const float2 b) { return a.x * b.x + a.y * b.y; }

// v_hat = v / len with len = |v| (1 when |v| == 0).  Returns v_bar for a given v_hat_bar.
__device__ __forceinline__ float2 normalize_adj(const float2 v, const float2 vh_bar) {
    const float sq = v.x * v.x + v.y * v.y;
    const float len = sqrtf(sq);
    if (len == 0.0f) return vh_bar;  // v / 1
    const float inv = 1.0f / len;
    const float2 vh = make_float2(v.x * inv, v.y * inv);
    const float pr = dot2(vh_bar, vh);
    return make_float2((vh_bar.x - pr * vh.x) * inv, (vh_bar.y - pr * vh.y) * inv);
}

// VJP of residual() (Wall geometry.py:641-650, RIS :698-711) for an upstream cotangent g.
// Accumulates into a_bar, b_bar, c_bar (points), n_bar (unit normal), phi_bar.
__device__ __forceinline__ void residual_adj(const int kind, const float2 a, const float2 b, const float2 c,
                                             const float4 w1, const float2 sc, const float g, float2& a_bar,
                                             float2& b_bar, float2& c_bar, float2& n_bar, float& phi_bar) {
    if (kind == D2D_KIND_VERTEX) return;
    const float2 n = make_float2(w1.x, w1.y);
    const float2 rv = make_float2(c.x - b.x, c.y - b.y);
    float l;
    const float2 r = normalize2(rv, l);
    float2 r_bar;
    if (kind == D2D_KIND_WALL) {
        const float2 iv = make_float2(b.x - a.x, b.y - a.y);
        const float2 i = normalize2(iv, l);
        const float c2 = 2.0f * dot2(i, n);
        const float2 e = make_float2(r.x - (i.x - c2 * n.x), r.y - (i.y - c2 * n.y));
        const float2 e_bar = make_float2(2.0f * g * e.x, 2.0f * g * e.y);
        const float en = dot2(e_bar, n);
        r_bar = e_bar;
        const float2 i_bar = make_float2(-e_bar.x + 2.0f * en * n.x, -e_bar.y + 2.0f * en * n.y);
        n_bar.x += c2 * e_bar.x + 2.0f * en * i.x;
        n_bar.y += c2 * e_bar.y + 2.0f * en * i.y;
        const float2 iv_bar = normalize_adj(iv, i_bar);
        b_bar.x += iv_bar.x; b_bar.y += iv_bar.y;
        a_bar.x -= iv_bar.x; a_bar.y -= iv_bar.y;
    } else {  // RIS
        const float mx = -r.x, my = -r.y;
        const float sin_a = mx * n.y - my * n.x;
        const float cos_a = mx * n.x + my * n.y;
        const float ds = 2.0f * g * (sin_a - sc.x);
        const float dc = 2.0f * g * (cos_a - sc.y);
        phi_bar += -ds * sc.y + dc * sc.x;
        const float mbx = ds * n.y + dc * n.x;
        const float mby = -ds * n.x + dc * n.y;
        r_bar = make_float2(-mbx, -mby);
        n_bar.x += -ds * my + dc * mx;
        n_bar.y += ds * mx + dc * my;
    }
    const float2 rv_bar = normalize_adj(rv, r_bar);
    c_bar.x += rv_bar.x; c_bar.y += rv_bar.y;
    b_bar.x -= rv_bar.x; b_bar.y -= rv_bar.y;
}

// VJP of normalize (v_hat = v / len) from the unit vector and the length the forward produced: no square root, one fast
// reciprocal (the sweep feeds no predicate).  len == 1 with v_hat == v for a zero-length segment (v / 1): the
// projection term then vanishes with |v|^2 and the cotangent passes through, like the `where` of geometry.py:227-230.
__device__ __forceinline__ float2 normalize_adj_hat(const float2 vh, const float len, const float2 vh_bar) {
    const float inv = __fdividef(1.0f, len);
    const float pr = dot2(vh_bar, vh);
    return make_float2((vh_bar.x - pr * vh.x) * inv, (vh_bar.y - pr * vh.y) * inv);
}

// residual_adj() from the unit directions i = normalize(b - a), r = normalize(c - b) and lengths the trace kept
// (path_loss_dirs): the residual vector e is formed from the forward's own bits, as in residual_adj.
__device__ __forceinline__ void residual_adj_dirs(const int kind, const float2 i, const float li, const float2 r,
                                                  const float lr, const float4 w1, const float2 sc, const float g,
                                                  float2& a_bar, float2& b_bar, float2& c_bar, float2& n_bar,
                                                  float& phi_bar) {
    if (kind == D2D_KIND_VERTEX) return;
    const float2 n = make_float2(w1.x, w1.y);
    float2 r_bar;
    if (kind == D2D_KIND_WALL) {
        const float c2 = 2.0f * dot2(i, n);
        const float2 e = make_float2(r.x - (i.x - c2 * n.x), r.y - (i.y - c2 * n.y));
        const float2 e_bar = make_float2(2.0f * g * e.x, 2.0f * g * e.y);
        const float en = dot2(e_bar, n);
        r_bar = e_bar;
        const float2 i_bar = make_float2(-e_bar.x + 2.0f * en * n.x, -e_bar.y + 2.0f * en * n.y);
        n_bar.x += c2 * e_bar.x + 2.0f * en * i.x;
        n_bar.y += c2 * e_bar.y + 2.0f * en * i.y;
        const float2 iv_bar = normalize_adj_hat(i, li, i_bar);
        b_bar.x += iv_bar.x; b_bar.y += iv_bar.y;
        a_bar.x -= iv_bar.x; a_bar.y -= iv_bar.y;
    } else {  // RIS
        const float mx = -r.x, my = -r.y;
        const float sin_a = mx * n.y - my * n.x;
        const float cos_a = mx * n.x + my * n.y;
        const float ds = 2.0f * g * (sin_a - sc.x);
        const float dc = 2.0f * g * (cos_a - sc.y);
        phi_bar += -ds * sc.y + dc * sc.x;
        const float mbx = ds * n.y + dc * n.x;
        const float mby = -ds * n.x + dc * n.y;
        r_bar = make_float2(-mbx, -mby);
        n_bar.x += -ds * my + dc * mx;
        n_bar.y += ds * mx + dc * my;
    }
    const float2 rv_bar = normalize_adj_hat(r, lr, r_bar);
    c_bar.x += rv_bar.x; c_bar.y += rv_bar.y;
    b_bar.x -= rv_bar.x; b_bar.y -= rv_bar.y;
}

// normalize_adj with the length normalize2 already produced (same values: len = sqrtf(v.v), 1 when v == 0)
__device__ __forceinline__ float2 normalize_adj_len(const float2 v, const float len, const bool zero, const float2 vh_bar) {
    if (zero) return vh_bar;  // v / 1
    const float inv = 1.0f / len;
    const float2 vh = make_float2(v.x * inv, v.y * inv);
    const float pr = dot2(vh_bar, vh);
    return make_float2((vh_bar.x - pr * vh.x) * inv, (vh_bar.y - pr * vh.y) * inv);
}

// residual() and residual_adj(..., g = 1) in one pass: the value and the cotangents of the three points, with every
// segment normalised ONCE (the two separate calls each normalise both segments, and normalize_adj takes the square
// roots a third time: 6 IEEE square roots and 10 divisions per wall interaction instead of 2 and 6).  Same operations
// on the same operands in the same order as the separate calls, hence the same bits — this is the inner loop of the
// MinPath solver (d2d_solver.cuh), whose iterates must not move.
__device__ __forceinline__ float residual_value_and_grad(const int kind, const float2 a, const float2 b, const float2 c,
                                                         const float4 w1, const float2 sc, float2& a_bar, float2& b_bar,
                                                         float2& c_bar) {
    if (kind == D2D_KIND_VERTEX) return 0.0f;
    const float2 n = make_float2(w1.x, w1.y);
    const float2 rv = make_float2(c.x - b.x, c.y - b.y);
    float lr;
    const float2 r = normalize2(rv, lr);
    const bool rzero = rv.x * rv.x + rv.y * rv.y == 0.0f;
    float2 r_bar;
    float value;
    if (kind == D2D_KIND_WALL) {
        const float2 iv = make_float2(b.x - a.x, b.y - a.y);
        float li;
        const float2 i = normalize2(iv, li);
        const bool izero = iv.x * iv.x + iv.y * iv.y == 0.0f;
        const float c2 = 2.0f * dot2(i, n);
        const float2 e = make_float2(r.x - (i.x - c2 * n.x), r.y - (i.y - c2 * n.y));
        value = e.x * e.x + e.y * e.y;
        const float2 e_bar = make_float2(2.0f * e.x, 2.0f * e.y);
        const float en = dot2(e_bar, n);
        r_bar = e_bar;
        const float2 i_bar = make_float2(-e_bar.x + 2.0f * en * n.x, -e_bar.y + 2.0f * en * n.y);
        const float2 iv_bar = normalize_adj_len(iv, li, izero, i_bar);
        b_bar.x += iv_bar.x; b_bar.y += iv_bar.y;
        a_bar.x -= iv_bar.x; a_bar.y -= iv_bar.y;
    } else {  // RIS
        const float mx = -r.x, my = -r.y;
        const float sin_a = mx * n.y - my * n.x;
        const float cos_a = mx * n.x + my * n.y;
        const float dsv = sin_a - sc.x, dcv = cos_a - sc.y;
        value = dsv * dsv + dcv * dcv;
        const float ds = 2.0f * dsv;
        const float dc = 2.0f * dcv;
        const float mbx = ds * n.y + dc * n.x;
        const float mby = -ds * n.x + dc * n.y;
        r_bar = make_float2(-mbx, -mby);
    }
    const float2 rv_bar = normalize_adj_len(rv, lr, rzero, r_bar);
    c_bar.x += rv_bar.x; c_bar.y += rv_bar.y;
    b_bar.x -= rv_bar.x; b_bar.y -= rv_bar.y;
    return value;
}

// Adjoints of the per-object table entries of one interacting object.
struct ObjAdj {
    float2 p1;   // origin
    float2 t;    // P2 - P1
    float2 n;    // unit normal
    float tt;    // t.t
    float phi;
    __device__ __forceinline__ void zero() {
        p1 = t = n = make_float2(0.f, 0.f);
        tt = 0.f;
        phi = 0.f;
    }
    // Folds (n, tt, t, p1) adjoints back to the raw vertices (P1, P2): returns (P1x, P1y, P2x, P2y).
    __device__ __forceinline__ float4 to_vertices(const float4 w0, const float4 w1) const {
        float2 tb = t;
        // n = m / L, m = (t.y, -t.x), L = |m| (w1.w, 1 if zero)
        const float2 nn = make_float2(w1.x, w1.y);
        const bool degenerate = (w0.z == 0.0f && w0.w == 0.0f);
        float2 mb;
        if (degenerate) mb = n;
        else {
            const float pr = dot2(n, nn);
            const float inv = 1.0f / w1.w;
            mb = make_float2((n.x - pr * nn.x) * inv, (n.y - pr * nn.y) * inv);
        }
        tb.y += mb.x;
        tb.x -= mb.y;
        if (!degenerate) {  // tt = t.t (constant 1 when zero)
            tb.x += 2.0f * tt * w0.z;
            tb.y += 2.0f * tt * w0.w;
        }
        return make_float4(p1.x - tb.x, p1.y - tb.y, tb.x, tb.y);
    }
};

}  // namespace d2d
