// differt2d_b200 — the per-CTA candidate driver shared by the forward and backward kernels.
//
// A CTA owns a compact TILE of grid points (16 x 8 when the caller declares the grid's row length,
// 128 consecutive points otherwise).  For every order k it walks the candidate list in list order
// (scene.py:166-175) in chunks of one candidate per thread:
//
//   1. CULL (one thread = one candidate, integer decode of its index): a conservative, tile-level
//      necessary condition for a non-zero validity — "some point of the tile's bounding box can have its
//      last interaction point on the last object" — evaluated from the four corners of the box
//      (the parametric coordinate is a linear-fractional function of the grid point, so its extrema over
//      the box are at the corners whenever the denominator keeps its sign).  The tolerance covers the fp32
//      error of the exact per-thread evaluation, so a culled candidate has validity EXACTLY 0 for every
//      point of the tile: results are unchanged, bit for bit (tests/test_gpu_parity.py).
//   2. ordered compaction of the survivors into shared memory (per-warp segments, list order kept);
//   3. TRACE: every thread evaluates the surviving candidates for its own grid point, exactly.
//
// The culled work still counts as algorithmic work ("no early-out credit").
#pragma once

#include "d2d_trace.cuh"

namespace d2d {

constexpr int kBlock = 128;
constexpr int kTileCols = 16;
constexpr int kTileRows = 8;

struct Tile {
    long long r;      // grid-point index of this thread (row-major), valid when `active`
    bool active;
    float4 bbox;      // xmin, ymin, xmax, ymax over the tile's active points
    float scale;      // max |coordinate| over tile, fixed points and objects (for error bounds)
};

struct DriverShared {
    int count;               // n_allowed
    float red[4][4];         // per-warp partials
    int wcount[2][4];        // survivors per warp segment, double buffered
    int4 list[2][kBlock];    // packed survivors: (c0 | c1 << 16, c2 | c3 << 16, index lo, index hi)
};

__device__ __forceinline__ long long tile_grid_blocks(const long long R, const int cols) {
    if (cols > 0 && R % cols == 0) {
        const long long rows = R / cols;
        return ((cols + kTileCols - 1) / kTileCols) * ((rows + kTileRows - 1) / kTileRows);
    }
    return (R + kBlock - 1) / kBlock;
}

// Maps (blockIdx, threadIdx) to a grid point; computes the tile's bounding box (all threads must call).
__device__ inline Tile make_tile(const KParams& p, const SceneTab& T, DriverShared& sh) {
    Tile t;
    const int tid = threadIdx.x;
    if (p.grid_cols > 0 && p.R % p.grid_cols == 0) {
        const long long rows = p.R / p.grid_cols;
        const int tiles_x = (p.grid_cols + kTileCols - 1) / kTileCols;
        const long long by = blockIdx.x / tiles_x;
        const int bx = (int)(blockIdx.x % tiles_x);
        const int col = bx * kTileCols + (tid & (kTileCols - 1));
        const long long row = by * kTileRows + (tid / kTileCols);
        t.active = col < p.grid_cols && row < rows;
        t.r = row * p.grid_cols + col;
    } else {
        t.r = (long long)blockIdx.x * kBlock + tid;
        t.active = t.r < p.R;
    }
    float xmin = CUDART_INF_F, ymin = CUDART_INF_F, xmax = -CUDART_INF_F, ymax = -CUDART_INF_F;
    if (t.active) {
        const float2 g = reinterpret_cast<const float2*>(p.grid)[t.r];
        xmin = xmax = g.x;
        ymin = ymax = g.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) {
        sh.red[warp][0] = xmin; sh.red[warp][1] = ymin; sh.red[warp][2] = xmax; sh.red[warp][3] = ymax;
    }
    __syncthreads();
    xmin = fminf(fminf(sh.red[0][0], sh.red[1][0]), fminf(sh.red[2][0], sh.red[3][0]));
    ymin = fminf(fminf(sh.red[0][1], sh.red[1][1]), fminf(sh.red[2][1], sh.red[3][1]));
    xmax = fmaxf(fmaxf(sh.red[0][2], sh.red[1][2]), fmaxf(sh.red[2][2], sh.red[3][2]));
    ymax = fmaxf(fmaxf(sh.red[0][3], sh.red[1][3]), fmaxf(sh.red[2][3], sh.red[3][3]));
    t.bbox = make_float4(xmin, ymin, xmax, ymax);
    float s = fmaxf(fmaxf(fabsf(xmin), fabsf(xmax)), fmaxf(fabsf(ymin), fabsf(ymax)));
    if (!(s < CUDART_INF_F)) s = 0.0f;  // empty tile
    for (int j = 0; j < p.N; ++j) {    // uniform, N is small next to the candidate count
        const float4 w = T.w0[j];
        s = fmaxf(s, fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fmaxf(fabsf(w.x + w.z), fabsf(w.y + w.w))));
    }
    for (int f = 0; f < p.T; ++f) s = fmaxf(s, fmaxf(fabsf(p.fixed[2 * f]), fabsf(p.fixed[2 * f + 1])));
    t.scale = s;
    __syncthreads();
    return t;
}

// Conservative tile test for an ImagePath candidate on a receivers grid.  `apex` is the last image of
// the transmitter; the last interaction point of a receiver q is X = q + (vn/un) u (geometry.py:1093-1107)
// and must satisfy min(s, 1 - s) > xz with s its parametric coordinate on the last object (geometry.py:
// 595-621), xz = x_zero<MODE>.  Returns false only if NO point of the box can satisfy it.
__device__ __forceinline__ bool tile_may_reach(const float2 apex, const float4 w0, const float4 w1, const int kind,
                                               const float4 bbox, const float scale, const float xz, const float p_tolscale) {
    if (kind == D2D_KIND_VERTEX) return true;
    if (!(xz > -CUDART_INF_F)) return true;
    const float qx[4] = {bbox.x, bbox.z, bbox.x, bbox.z};
    const float qy[4] = {bbox.y, bbox.y, bbox.w, bbox.w};
    float smin = CUDART_INF_F, smax = -CUDART_INF_F, unmin = CUDART_INF_F, gmax = 0.f, umax = 0.f, Gmax = 0.f;
    int pos = 0, neg = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float ux = qx[c] - apex.x, uy = qy[c] - apex.y;
        const float vx = w0.x - qx[c], vy = w0.y - qy[c];
        const float un = fmaf(ux, w1.x, uy * w1.y);
        const float vn = fmaf(vx, w1.x, vy * w1.y);
        pos += un > 0.f;
        neg += un < 0.f;
        const float g = vn / un;
        const float Xx = fmaf(g, ux, qx[c]), Xy = fmaf(g, uy, qy[c]);
        const float s = fmaf(w0.z, Xx - w0.x, w0.w * (Xy - w0.y)) / w1.z;
        smin = fminf(smin, s);
        smax = fmaxf(smax, s);
        const float ul = sqrtf(fmaf(ux, ux, uy * uy));
        unmin = fminf(unmin, fabsf(un));
        gmax = fmaxf(gmax, fabsf(g));
        umax = fmaxf(umax, ul);
        Gmax = fmaxf(Gmax, fabsf(g) * ul);
    }
    if (!(pos == 4 || neg == 4)) return true;   // the denominator may vanish inside the box
    if (!(smin == smin) || !(smax == smax)) return true;
    // fp32 error of the exact evaluation, first order: the image carries ~2K eps S, so
    // err(X) <~ 4 eps S (|g| + |u|/|un| + |g||u|/|un|) + eps S; 16 eps S (...) is used, times p_tolscale (4)
    const float eps = 5.9604645e-8f;
    const float tolX = 16.0f * eps * scale * (gmax + umax / unmin + Gmax / unmin) + 4.0f * eps * (scale + Gmax);
    const float tol = 1e-4f + p_tolscale * tolX / sqrtf(w1.z);
    if (!(tol < CUDART_INF_F)) return true;
    const float lo = xz - tol, hi = 1.0f - xz + tol;
    return !(smax < lo || smin > hi);
}

// number of candidates of order K over m visitable objects
__device__ __forceinline__ long long order_count(const int K, const int m) {
    if (K == 0) return 1;
    if (m <= 0 || (K > 1 && m < 2)) return 0;
    long long c = m;
    for (int i = 1; i < K; ++i) c *= (m - 1);
    return c;
}

// Walks all candidates of order K for the fixed point `fx`; calls visit(cd, col) — uniformly over the
// CTA — for every candidate that survives the tile cull, in list order.  `col0` = column of the first
// candidate of this order in the global list.
template <int MODE, int METHOD, int K, bool TXGRID, class Visit>
__device__ __forceinline__ void for_each_candidate(const SceneTab& T, const KParams& p, const Tile& tile,
                                                   DriverShared& sh, const float alpha, const float2 fx,
                                                   const long long col0, int& buf, Visit&& visit) {
    const int m = T.n_allowed;
    const long long Ck = order_count(K, m);
    if (Ck == 0) return;
    if constexpr (K == 0) {
        Cand<0> cd;
        cd.c[0] = 0;
        visit(cd, col0);
        return;
    }
    constexpr int KK = K > 0 ? K : 1;
    const bool cull = (METHOD == D2D_METHOD_IMAGE) && !TXGRID && p.cull;
    const float xz = x_zero<MODE>(alpha);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (long long base = 0; base < Ck; base += kBlock) {
        const long long idx = base + tid;
        bool keep = false;
        int c[KK];
#pragma unroll
        for (int i = 0; i < KK; ++i) c[i] = 0;
        if (idx < Ck) {
            // index -> positions in `allowed` (lexicographic, no equal neighbours)
            long long rem = idx;
            int dig[KK];
#pragma unroll
            for (int i = K - 1; i >= 1; --i) {
                dig[i] = (int)(rem % (m - 1));
                rem /= (m - 1);
            }
            dig[0] = (int)rem;
            int prev = -1;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                int pos = dig[i];
                if (i > 0 && pos >= prev) ++pos;
                c[i] = T.allowed[pos];
                prev = pos;
            }
            keep = true;
            if (cull) {
                float2 I = fx;
#pragma unroll
                for (int i = 0; i < K; ++i) I = mirror(I, T.w0[c[i]], T.w1[c[i]]);
                const int j = c[K - 1];
                keep = tile_may_reach(I, T.w0[j], T.w1[j], T.kind[j], tile.bbox, tile.scale, xz, (float)p.cull);
            }
        }
        // ordered compaction: per-warp segments keep list order
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int off = __popc(ballot & ((1u << lane) - 1u));
            sh.list[buf][warp * 32 + off] =
                make_int4(c[0] | ((K > 1 ? c[K > 1 ? 1 : 0] : 0) << 16),
                          (K > 2 ? c[K > 2 ? 2 : 0] : 0) | ((K > 3 ? c[K > 3 ? 3 : 0] : 0) << 16),
                          (int)(idx & 0xffffffffLL), (int)(idx >> 32));
        }
        if (lane == 0) sh.wcount[buf][warp] = __popc(ballot);
        __syncthreads();
#pragma unroll 1
        for (int w = 0; w < kBlock / 32; ++w) {
            const int n = sh.wcount[buf][w];
#pragma unroll 1
            for (int q = 0; q < n; ++q) {
                const int4 e = sh.list[buf][w * 32 + q];
                Cand<K> cd;
                cd.c[0] = e.x & 0xffff;
                if (K > 1) cd.c[K > 1 ? 1 : 0] = (e.x >> 16) & 0xffff;
                if (K > 2) cd.c[K > 2 ? 2 : 0] = e.y & 0xffff;
                if (K > 3) cd.c[K > 3 ? 3 : 0] = (e.y >> 16) & 0xffff;
                const long long ci = ((long long)(unsigned)e.z) | ((long long)e.w << 32);
                visit(cd, col0 + ci);
            }
        }
        buf ^= 1;  // the next chunk fills the other buffer: one barrier per chunk
    }
}

}  // namespace d2d
