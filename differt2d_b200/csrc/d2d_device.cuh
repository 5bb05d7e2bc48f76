// differt2d_b200 — device-side building blocks of the path-tracing kernels (sm_100a, FP32 CUDA cores).
//
// Arithmetic discipline.  The translation units that include this header are compiled with
// `-fmad=false -prec-div=true -prec-sqrt=true -ftz=false`: every `*`, `+`, `-`, `/`, `sqrtf`
// below is one IEEE-754 binary32 operation in the order written, which is the order of the
// reference's Python (cited per function).  That makes the geometric quantities that feed the
// hard-logic predicates bit-reproducible.  Where an approximate value is enough (the occlusion
// pre-filter) fused multiply-adds are requested explicitly with fmaf() and the result is only
// used to decide whether the exact expression has to be evaluated at all.
//
// Validity in "pre-activation" form.  The reference evaluates an activation per comparison and
// folds them with min/max (logic.py:315-358).  Every activation is the same non-decreasing map
// act(x) = f(alpha*x) (alpha > 0) applied to a difference x, and rounding-to-nearest keeps
// monotone maps monotone, so   min_i act(x_i) == act(min_i x_i)   and   max_i act(x_i) ==
// act(max_i x_i)   hold bit-for-bit.  The kernels therefore fold the *differences*
//     onx     = min_i min(s_i - 0, 1 - s_i)                                   (geometry.py:841-854, 600-621)
//     interx  = max_{seg,obj} min(ta + tol, (1+tol) - ta, tb + tol, (1+tol) - tb)   (geometry.py:153-173, 887-904)
//     lx      = tol_loss - loss                                               (geometry.py:960)
// and apply the activation three times per path instead of 2k + 4(k+1)N + 1 times.  In hard mode
// the comparisons `x >= y` / `x <= y` / `x < y` are equivalent to sign tests of the rounded
// differences (fl(a-b) has the sign of a-b), so the same folds serve both logics.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/differt2d_b200.h"

namespace d2d {

// Optional census of the culls and of the trace (diagnostic builds only: python -m differt2d_b200.build --debug-counters)
#ifdef D2D_DEBUG_COUNTERS
__device__ unsigned long long d2d_dbg[32];
#define D2D_COUNT(i) atomicAdd(&d2d_dbg[i], 1ULL)
#else
#define D2D_COUNT(i) ((void)0)
#endif

constexpr float kEps32 = 1.1920928955078125e-07f;  // geometry.py:200
constexpr float kTolSeg = 0.005f;                  // geometry.py:89
constexpr float kHiSeg = 1.0f + 0.005f;            // geometry.py:169  fl(1.0 + tol)
constexpr float kHalfSpan = 0.505f;                // (hi - lo)/2 : hx ~= kHalfSpan - max(|ta-.5|,|tb-.5|)
constexpr int kMaxOrder = D2D_MAX_ORDER;

// ---- kernel parameters (by value) -----------------------------------------------------------
struct KParams {
    const float* xys;      // [N,2,2]
    const uint8_t* kinds;  // [N] or nullptr
    const float* phis;     // [N] or nullptr
    const float* fixed;    // [T,2]
    const float* grid;     // [R,2]
    const float* x0;       // [C,max_order] or nullptr
    const float* alpha_dev;
    long long R;
    int grid_cols;         // row length of the grid when it is a row-major n x m mesh (0: unknown)
    int grid_rows;         // R / grid_cols, set by the launcher with grid_cols (the kernels map CTAs to tiles in 32-bit
                           // arithmetic: four 64-bit divisions per thread were 2 % of the forward launch)
    int cull;              // tile-level candidate culling on/off (results are identical either way)
    int slices;            // gridDim.y: CTAs sharing one tile, each walking every slices-th chunk of candidates
    int shard_index, shard_count;  // multi-GPU candidate sharding: this launch owns virtual slices shard_index * slices + y
    int tile_points;       // grid points per CTA when the tile is 1-D (<= kBlock; 1 for point-to-point links)
    int cluster;           // 8: 2-D tiles are numbered cluster by cluster (8 CTAs = 2 x 4 tiles = 32 x 32 points); 0: row-major
    int macro;             // set by a launcher that starts the kernel as clusters of 8 CTAs: macro-tile cull stage on
    int N, T;
    int min_order, max_order;
    int steps;
    int many;              // restarts of the Fermat/MinPath scan (>= 1)
    int opt;               // D2D_OPT_*
    float b1, b2, opt_eps; // Adam b1 / b2 / eps; SGD momentum in b1
    int fun, reduce_all;
    long long C_total;     // candidates over all orders (columns of valid_out)
    float alpha, tol, patch, lr;
    float h2;                      // (float)(height*height), folded in double like Python does
    float rc_pow[kMaxOrder + 1];   // (float)(r_coef**k)
    uint32_t blocked[D2D_MAX_OBJECTS / 32];  // bit j set: object j is never visited (filter_objects)
    // activity mask (optional): bit (t, warp, candidate column) = "some lane of the warp has validity != 0".
    // The forward kernel writes it, the backward kernel then only re-traces the set bits (no cull, no dead paths).
    uint32_t* mask;            // [T, gridDim.x * 4 warps, mask_wpw] words, or nullptr
    long long mask_wpw;        // words per (fixed point, warp) = ceil(C_total / 32)
};

// ---- shared-memory scene table ---------------------------------------------------------------
// One entry per object, derived once per CTA from the raw vertices with the reference's formulas.
struct SceneTab {
    float4* w0;    // P1.x, P1.y, t.x, t.y                       Ray.origin/t   geometry.py:458-487
    float4* w1;    // n.x, n.y, tt (0 -> 1), len(n_raw) (0 -> 1) Wall.normal :561-573, :596-597
    float4* w2;    // P1'.x, P1'.y, A.x, A.y  (patched wall; A = 0 for a Vertex)   :632-635
    float2* sc;    // sin(phi), cos(phi)                         RIS :707-708
    int* kind;     // D2D_KIND_*
    short* allowed;  // ascending list of visitable objects
    int n_allowed;
    float fold_skip;  // fold_skip_bound<MODE>(alpha), set by the kernels that want the shortcut (-inf: never)
};

__host__ __device__ inline size_t scene_tab_bytes(int N) {
    return (size_t)N * (3 * sizeof(float4) + sizeof(float2) + sizeof(int) + sizeof(short)) + 64;
}

__device__ __forceinline__ SceneTab carve_tab(unsigned char* smem, int N) {
    SceneTab T;
    T.w0 = reinterpret_cast<float4*>(smem);
    T.w1 = T.w0 + N;
    T.w2 = T.w1 + N;
    T.sc = reinterpret_cast<float2*>(T.w2 + N);
    T.kind = reinterpret_cast<int*>(T.sc + N);
    T.allowed = reinterpret_cast<short*>(T.kind + N);
    T.n_allowed = 0;
    T.fold_skip = -CUDART_INF_F;
    return T;
}

// Builds the table cooperatively; must be called by every thread of the CTA.
__device__ inline void build_tab(SceneTab& T, const KParams& p, int* s_count) {
    const int N = p.N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float4 raw = reinterpret_cast<const float4*>(p.xys)[j];  // P1.x P1.y P2.x P2.y
        const int kind = p.kinds ? (int)p.kinds[j] : D2D_KIND_WALL;
        const float tx = raw.z - raw.x, ty = raw.w - raw.y;
        // Wall.normal: n = normalize((t.y, -t.x))  geometry.py:568-572, 227-230
        const float mx = ty, my = -tx;
        float len = sqrtf(mx * mx + my * my);
        if (len == 0.0f) len = 1.0f;
        float tt = tx * tx + ty * ty;  // geometry.py:596-597
        if (tt == 0.0f) tt = 1.0f;
        T.w0[j] = make_float4(raw.x, raw.y, tx, ty);
        T.w1[j] = make_float4(mx / len, my / len, tt, len);
        // intersects_cartesian: origin - patch*t, dest + patch*t  geometry.py:632-635
        const float p1x = raw.x - p.patch * tx, p1y = raw.y - p.patch * ty;
        const float p2x = raw.z + p.patch * tx, p2y = raw.w + p.patch * ty;
        float ax = p2x - p1x, ay = p2y - p1y;
        if (kind == D2D_KIND_VERTEX) { ax = 0.0f; ay = 0.0f; }  // never occludes (geometry.py:405-414)
        T.w2[j] = make_float4(p1x, p1y, ax, ay);
        // (sinf(0) = 0 and cosf(0) = 1 exactly; walls-only scenes skip the range reduction: build_tab was 3 % of the
        // forward kernel's instructions, per CTA)
        const float phi = p.phis ? p.phis[j] : 0.0f;
        T.sc[j] = phi == 0.0f ? make_float2(phi, 1.0f) : make_float2(sinf(phi), cosf(phi));
        T.kind[j] = kind;
    }
    if (threadIdx.x == 0) {
        int m = 0;
        for (int j = 0; j < N; ++j)
            if (!((p.blocked[j >> 5] >> (j & 31)) & 1u)) T.allowed[m++] = (short)j;
        *s_count = m;
    }
    __syncthreads();
    T.n_allowed = *s_count;
}

// ---- logic ----------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float act(float x, float alpha) {
    const float z = alpha * x;
    if (MODE == D2D_MODE_SIGMOID) {
        return 1.0f / (1.0f + expf(-z));  // logic.py:235
    }
    const float v = z + 3.0f;  // logic.py:255  relu6(z + 3) / 6
    // the saturated branches are returned directly: 0/6 and 6/6 are exact, and an IEEE division with a zero
    // numerator takes the 115-instruction slow path of div.rn (14 % of the forward kernel before this)
    if (!(v > 0.0f)) return 0.0f;  // also NaN: relu6 via fmax/fmin maps NaN to 0
    if (v >= 6.0f) return 1.0f;
    return v / 6.0f;
}

// act(x) == 0 exactly?  (the path is dead: a conjunct of is_valid vanishes)  Hard logic: the comparison is false.
template <int MODE>
__device__ __forceinline__ bool act_is_zero(float x, float alpha) {
    if (MODE == D2D_MODE_HARD) return !(x >= 0.0f);
    if (MODE == D2D_MODE_HARD_SIGMOID) return !(alpha * x + 3.0f > 0.0f);
    const float z = alpha * x;
    if (z > -87.0f) return false;  // expf(87) is finite: 1/(1+e) > 0
    return act<MODE>(x, alpha) == 0.0f;
}

// act(x) == 1 for sure?  (`intersects` is then exactly true / 1.0: the path is dead.)  Conservative on purpose: a
// "no" only means that the fold goes on and the exact validity is formed at the end, so the results do not depend on
// it — while evaluating the activation itself on every update of the running maximum cost an expf and an IEEE
// division in the sigmoid kernels (the un-prunable regime spends 70 % of its time in this fold).
template <int MODE>
__device__ __forceinline__ bool act_is_one(float x, float alpha) {
    const float z = alpha * x;
    if (MODE == D2D_MODE_SIGMOID) return z >= 18.0f;  // expf(-18) < 2^-25: 1 + e rounds to 1 and 1 / 1 == 1
    return z + 3.0f >= 6.0f;                          // relu6 saturates exactly
}

// d/dz of f at z = alpha*x (the VJP rules of jnp.minimum/maximum give 1/2 at the kinks)
template <int MODE>
__device__ __forceinline__ float act_dz(float x, float alpha) {
    const float z = alpha * x;
    if (MODE == D2D_MODE_SIGMOID) {
        const float s = 1.0f / (1.0f + expf(-z));
        return s * (1.0f - s);
    }
    const float v = z + 3.0f;
    float g = (v > 0.0f && v < 6.0f) ? 1.0f : 0.0f;
    if (v == 0.0f || v == 6.0f) g = 0.5f;
    return g / 6.0f;
}

// Threshold on the occlusion pre-filter measure m: any test with m >= cthr(interx) cannot raise
// interx (hx ~= kHalfSpan - m up to a few ulp; the margin is generous on purpose).
__device__ __forceinline__ float filter_threshold(float interx) {
    const float c = kHalfSpan - interx;
    return fmaf(fabsf(c), 4e-6f, c + 4e-6f);
}

// Smallest pre-activation that still matters: act(x) == 0 exactly for every x <= x_zero.
template <int MODE>
__device__ __forceinline__ float x_zero(float alpha) {
    if (MODE == D2D_MODE_HARD) return 0.0f;  // only hx >= 0 matters (with margin from filter_threshold)
    if (MODE == D2D_MODE_HARD_SIGMOID) return -3.0f / alpha - fabsf(4.0f / alpha) * 1e-5f;
    // sigmoid: 1/(1+expf(-z)) is exactly 0 once expf(-z) overflows (-z > 88.73): z <= -89 is safely there
    return -89.5f / alpha;
}

// ---- when the occlusion fold cannot matter (smooth logic) -----------------------------------------
// is_valid = min(a_on, 1 - a_in, a_l) (geometry.py:947-963) with a_in = act(interx), and every test of the fold is
// hx = min(ta + tol, (1 + tol) - ta, tb + tol, (1 + tol) - tb) <= (1 + 2 tol) / 2 = 0.505 (+ an ulp) whatever ta is:
// a_in <= act(0.51), hence 1 - a_in >= 1 - act(0.51).  A path whose v0 = min(a_on, a_l) is not above that has
// validity v0 EXACTLY, whatever the fold returns — for a soft activation (sigmoid, small alpha) that is most of the
// list: act(0.51) = 0.62 at alpha = 1, and a path that misses its wall or breaks the reflection law has v0 < 0.38.
// The 4e-6 keeps the shortcut away from the last-bit behaviour of expf (2 ulp) and of the two roundings after it: a
// skipped fold would have returned 1 - a_in > v0 strictly, so the min, its arg-min and its tie count are unchanged.
template <int MODE>
__device__ __forceinline__ float fold_skip_bound(const float alpha) {
    if (MODE == D2D_MODE_HARD) return -CUDART_INF_F;
    return (1.0f - act<MODE>(0.51f, alpha)) - 4e-6f;
}

// The bound at the point of use.  sigmoid: the value the kernel computed once (an exponential); hard_sigmoid: recomputed
// from alpha (a few uniform operations, and no register held across the whole kernel: as a SceneTab member it cost the
// 64-register headline forward 1.3 %).  The shortcut is exact, so every kernel that folds may take it.
template <int MODE>
__device__ __forceinline__ float fold_skip_of(const SceneTab& T, const float alpha) {
#if defined(D2D_FOLD_SKIP) && !D2D_FOLD_SKIP
    return -CUDART_INF_F;
#else
    if (MODE == D2D_MODE_HARD_SIGMOID) return fold_skip_bound<MODE>(alpha);
    return T.fold_skip;
#endif
}

// Where the fold has to start caring: the largest x' >= xz with 1 - act(x') >= v0 + 2e-6, VERIFIED with the
// canonical activation (the closed form / fast logarithm below only proposes it).  Tests with hx <= x' cannot bring
// 1 - a_in down to v0: they are filtered like the ones below x_zero, and the exact divisions run for real
// occluders only.  Falls back to xz.  (Used by the sigmoid kernels only: in the hard_sigmoid kernels a per-thread
// threshold instead of the uniform x_zero cost the headline forward 3 % in registers for paths that are nearly all at
// v0 = 1 there, where fold_start == x_zero anyway.)
template <int MODE>
__device__ __forceinline__ float fold_start(const float v0, const float alpha, const float xz) {
    if (MODE == D2D_MODE_HARD) return xz;
    const float tgt = (1.0f - v0) - 4e-6f;  // wanted: act(x') <= tgt
    if (!(tgt > 1e-6f) || !(tgt < 1.0f)) return xz;
    float x;
    if (MODE == D2D_MODE_SIGMOID) {
        x = __logf(__fdividef(tgt, 1.0f - tgt));  // logit
        x = __fdividef(x - 1e-3f * (1.0f + fabsf(x)), alpha);
    } else {
        x = __fdividef(6.0f * tgt - 3.0f, alpha);
        x = x - 1e-5f * (fabsf(x) + __fdividef(3.0f, alpha));
    }
    if (!(x > xz)) return xz;
    const float a = act<MODE>(x, alpha);
    return (1.0f - a >= v0 + 2e-6f) ? x : xz;
}

// ---- geometry -------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Wall.image_of — geometry.py:652-670
__device__ __forceinline__ float2 mirror(const float2 p, const float4 w0, const float4 w1) {
    const float ix = p.x - w0.x, iy = p.y - w0.y;
    const float cc = 2.0f * (ix * w1.x + iy * w1.y);
    return make_float2(p.x - cc * w1.x, p.y - cc * w1.y);
}

// one step of the backward scan of ImagePath — geometry.py:1093-1107
__device__ __forceinline__ float2 back_project(const float2 point, const float2 image, const float4 w0,
                                               const float4 w1) {
    const float ux = point.x - image.x, uy = point.y - image.y;
    const float vx = w0.x - point.x, vy = w0.y - point.y;
    const float un = ux * w1.x + uy * w1.y;
    const float vn = vx * w1.x + vy * w1.y;
    float2 q = point;
    if (un != 0.0f) {
        q.x = point.x + (vn * ux) / un;
        q.y = point.y + (vn * uy) / un;
    }
    return q;
}

// normalize — geometry.py:206-230
__device__ __forceinline__ float2 normalize2(const float2 v, float& len) {
    const float sq = v.x * v.x + v.y * v.y;
    if (sq == 0.0f) {  // |v| == 0 -> v / 1 == v exactly; taken apart because sqrt.rn(0) and div.rn(0, .) both
        len = 1.0f;    // leave the fast path of their IEEE expansions (zero-length closure walls hit this often)
        return v;
    }
    len = sqrtf(sq);
    return make_float2(v.x / len, v.y / len);
}

// Interactable.evaluate_cartesian — Wall geometry.py:641-650, RIS :698-711, Vertex :416-419 — from the unit
// directions i = normalize(b - a) and r = normalize(c - b)
__device__ __forceinline__ float residual_dirs(const int kind, const float2 i, const float2 r, const float4 w1,
                                               const float2 sc) {
    if (kind == D2D_KIND_VERTEX) return 0.0f;
    if (kind == D2D_KIND_WALL) {
        const float c2 = 2.0f * (i.x * w1.x + i.y * w1.y);
        const float ex = r.x - (i.x - c2 * w1.x);
        const float ey = r.y - (i.y - c2 * w1.y);
        return ex * ex + ey * ey;
    }
    const float mx = -r.x, my = -r.y;
    const float sin_a = mx * w1.y - my * w1.x;
    const float cos_a = mx * w1.x + my * w1.y;
    const float ds = sin_a - sc.x, dc = cos_a - sc.y;
    return ds * ds + dc * dc;
}

__device__ __forceinline__ float residual(const int kind, const float2 a, const float2 b, const float2 c,
                                          const float4 w1, const float2 sc) {
    if (kind == D2D_KIND_VERTEX) return 0.0f;
    float l;
    const float2 r = normalize2(make_float2(c.x - b.x, c.y - b.y), l);
    float2 i = make_float2(0.f, 0.f);
    if (kind == D2D_KIND_WALL) i = normalize2(make_float2(b.x - a.x, b.y - a.y), l);
    return residual_dirs(kind, i, r, w1, sc);
}

// Wall.cartesian_to_parametric — geometry.py:589-598
__device__ __forceinline__ float to_parametric(const float2 x, const float4 w0, const float4 w1) {
    const float ox = x.x - w0.x, oy = x.y - w0.y;
    const float num = w0.z * ox + w0.w * oy;
    return num == 0.0f ? num : num / w1.z;  // (+-0 / tt == +-0 exactly, tt > 0; avoids div.rn's slow path)
}

// path_length — geometry.py:176-203
template <int NP>
__device__ __forceinline__ float path_length(const float2 (&X)[NP]) {
    float total = 0.0f;
#pragma unroll
    for (int i = 0; i + 1 < NP; ++i) {
        const float dx = (X[i + 1].x - X[i].x) + kEps32;
        const float dy = (X[i + 1].y - X[i].y) + kEps32;
        const float len = sqrtf(dx * dx + dy * dy);
        total = (i == 0) ? len : total + len;
    }
    return total;
}

// Exact pre-activation of one segment/object test — geometry.py:153-173 with tol = 0.005
static __device__ __forceinline__ float hit_exact(const float a, const float b, const float d) {
    if (d == 0.0f) return -CUDART_INF_F;  // t = +inf  ->  (1+tol) - inf
    const float ta = a / d, tb = b / d;
    const float h = fminf(fminf(ta + kTolSeg, kHiSeg - ta), fminf(tb + kTolSeg, kHiSeg - tb));
    // fminf drops NaNs; a NaN parameter makes every comparison false in the reference
    return (ta != ta || tb != tb) ? -CUDART_INF_F : h;
}

// ---- candidate odometer (lexicographic, no equal neighbours) over positions in `allowed` ------
template <int K>
struct Odometer {
    int pos[K > 0 ? K : 1];
    __device__ __forceinline__ bool first(int m) {
        if (K == 0) return true;
        if (m <= 0 || (K > 1 && m < 2)) return false;
#pragma unroll
        for (int i = 0; i < K; ++i) pos[i] = (i & 1);  // 0,1,0,1,...
        return true;
    }
    __device__ __forceinline__ bool next(int m) {
        if (K == 0) return false;
#pragma unroll
        for (int i = K - 1; i >= 0; --i) {
            int v = pos[i] + 1;
            if (i > 0 && v == pos[i - 1]) ++v;
            if (v < m) {
                pos[i] = v;
#pragma unroll
                for (int j = i + 1; j < K; ++j) pos[j] = (pos[j - 1] == 0) ? 1 : 0;
                return true;
            }
        }
        return false;
    }
};

}  // namespace d2d
