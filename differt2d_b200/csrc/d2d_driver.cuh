// differt2d_b200 — the per-CTA candidate driver shared by the forward and backward kernels.
//
// A CTA owns a compact TILE of grid points (16 x 8 when the caller declares the grid's row length,
// 128 consecutive points otherwise).  For every order k it walks the candidate list in list order
// (scene.py:166-175) in chunks of one candidate per thread:
//
//   0. MACRO CULL (2-D grids; the kernel then runs as thread-block CLUSTERS of 8 CTAs = 2 x 4 tiles = one macro tile
//      of 32 x 32 points), once, before anything is traced: the cluster's 32 warps split the whole (fixed point,
//      candidate) list, every thread tests one candidate against the MACRO tile's bounding box with the same
//      conservative test as step 1, and each warp pushes its ballot — one word of the survivor bitmap — into the
//      shared memory of all 8 CTAs (distributed shared memory).  After ONE cluster barrier every CTA owns the complete
//      bitmap and never talks to its neighbours again (no barrier at which a lightly loaded tile would wait for a
//      heavy one; a first version with per-order survivor lists and two cluster barriers per order lost on the raw
//      scene what the cull saved).  Steps 1-4 then only see the survivors, enumerated in list order through
//      per-word prefix counts.  The tile-level test costs ~700 warp instructions per 128 candidates and used to run
//      on every candidate in every CTA (12 % of the forward kernel on the raw city scene).
//   1. CULL (one thread = one candidate, integer decode of its index): a conservative, tile-level
//      necessary condition for a non-zero validity — "some point of the tile's bounding box can have its
//      last interaction point on the last object" — evaluated from the four corners of the box
//      (the parametric coordinate is a linear-fractional function of the grid point, so its extrema over
//      the box are at the corners whenever the denominator keeps its sign).  The tolerance covers the fp32
//      error of the exact per-thread evaluation, so a culled candidate has validity EXACTLY 0 for every
//      point of the tile: results are unchanged, bit for bit (tests/test_gpu_parity.py).
//   2. ordered compaction of the survivors into shared memory (per-warp segments, list order kept);
//   3. WARP CULL: each warp re-tests the survivors (one per lane) against the bounding box of ITS 32 points
//      (16 x 2 in a tiled grid) with the s-range rule on the last interaction only, reusing the error bound the
//      tile-level test established (it holds for every point of the tile, hence for the sub-box's corners);
//      the warp then walks the set bits of the ballot, still in list order;
//   4. TRACE: every thread evaluates the remaining candidates for its own grid point, exactly.
//
// The culled work still counts as algorithmic work ("no early-out credit").
#pragma once

#include <cooperative_groups.h>

#include "d2d_trace.cuh"

namespace d2d {

namespace cg = cooperative_groups;

constexpr int kBlock = 128;
// build switch of the fold shortcut (1: on; the 0 build exists for A/B timing, results are identical)
#ifndef D2D_FOLD_SKIP
#define D2D_FOLD_SKIP 1
#endif
// resident CTAs per SM the kernels are compiled for (register caps 64 / 96 per thread): occupancy hides the
// instruction-fetch and fixed-latency stalls that dominate these branchy FP32 kernels (profiles/)
#ifndef D2D_FWD_MIN_CTAS
#define D2D_FWD_MIN_CTAS 8
#endif
#ifndef D2D_BWD_MIN_CTAS
#if defined(D2D_TU_SOLVER) && D2D_TU_SOLVER
#define D2D_BWD_MIN_CTAS 5  // FermatPath / MinPath: the reverse sweep through the scan keeps far more state alive
#elif defined(D2D_TU_MODE) && D2D_TU_MODE == D2D_MODE_SIGMOID
// sigmoid never saturates: every path runs the reverse sweep, whose live state does not fit 64 registers.  At 8
// CTAs/SM the dense leg is 2 % faster (4.98 vs 5.10 ms; 7: 5.02, 6: 5.23, r02t) but its spills no longer fit the L2:
// 1.08 GB of DRAM writes per launch (r02y) against 19 MB at 96 registers.  Not worth 2 %.
#define D2D_BWD_MIN_CTAS 5
#else
// hard / hard_sigmoid: the sweep runs for the few paths with a non-zero validity, the re-trace wants occupancy:
// 64 registers; measured 8 > 6 > 5 > 4 > 3 CTAs/SM (r02g, r02j)
#define D2D_BWD_MIN_CTAS 8
#endif
#endif
constexpr int kTileCols = 16;
constexpr int kTileRows = 8;
constexpr int kCluster = 8;        // CTAs per cluster = tiles per macro tile
constexpr int kMacroTilesX = 2;    // macro tile = 2 x 4 tiles = 32 x 32 grid points
constexpr int kMacroTilesY = 4;
constexpr int kMacroWords = 512;   // survivor bitmap: up to 16384 (fixed point, candidate) pairs, else no macro stage


struct Tile {
    long long r;      // grid-point index of this thread (row-major), valid when `active`
    bool active;
    float4 bbox;      // xmin, ymin, xmax, ymax over the tile's active points
    float4 mbox;      // same over the 8 tiles of the cluster (p.macro only)
    float4 wbox;      // same over this thread's warp (inverted / infinite when the warp has no active point)
    float scale;      // max |coordinate| over tile, fixed points and objects (for error bounds)
    float scale_x, scale_y;  // the same per component (lon/lat scenes: |x| ~ 5, |y| ~ 50 — the fp32 lattice differs 8x)
    float diam;       // extent of the bounding box of tile (macro tile), fixed points and objects: bounds every |P1 - X|
};

struct DriverShared {
    int count;               // n_allowed
    int hint[kBlock / 32];   // per warp: the object that blocked its previous path (intersects_x starts its fold there).
                             // (One slot per LAST OBJECT of the candidate instead — 8 or 32 per warp — was measured: the
                             // most recent blocker of ANY candidate predicts better, +8 % / +16 % on the forward launch.)
    float red[4][4];         // per-warp partials
    float ext[4][6];         // per-warp partials of the scene extent (make_tile)
    int wcount[2][4];        // survivors per warp segment, double buffered
    int4 list[2][kBlock];    // packed survivors: (c0 | c1 << 16, c2 | c3 << 16, index lo, index hi)
    float4 aux[2][kBlock];   // per survivor: apex (image of the fixed point through all objects) x, y; s-tolerance of
                             // the last interaction for the warp-level test (+inf: no test possible); unused
    // macro cull (written by all CTAs of the cluster through distributed shared memory)
    float4 tbox;                             // this tile's bounding box (read by the neighbours)
    uint32_t mbits[kMacroWords + 1];         // bit (t * C_total + column): the candidate may be valid in the macro tile
    unsigned short mpref[kMacroWords + 1];   // set bits in the words before word w
};

// CTAs along x for a problem (host side)
inline long long host_tile_blocks(const KParams& p) {
    if (p.grid_cols > 0 && p.R % p.grid_cols == 0) {
        const long long rows = p.R / p.grid_cols;
        const long long tx = (p.grid_cols + kTileCols - 1) / kTileCols, ty = (rows + kTileRows - 1) / kTileRows;
        if (p.cluster)  // whole macro tiles: CTAs beyond the grid's edge only help with the macro cull
            return ((tx + kMacroTilesX - 1) / kMacroTilesX) * ((ty + kMacroTilesY - 1) / kMacroTilesY) * kCluster;
        return tx * ty;
    }
    return (p.R + p.tile_points - 1) / p.tile_points;
}

// The macro stage needs clusters (2-D tiled grid, no candidate slices) and a bitmap that fits.
inline bool host_macro_ok(const KParams& p) {
    const long long bits = (long long)p.T * p.C_total;
    return p.cluster && p.cull && bits > 0 && bits <= 32LL * kMacroWords;
}

// Launches a tile kernel: gridDim = (tiles, candidate slices), kBlock threads; as clusters of 8 CTAs when `clusters`.
template <class... KArgs, class... Args>
inline cudaError_t launch_tiles(void (*kern)(KArgs...), const KParams& p, const bool clusters, const size_t smem,
                                cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)host_tile_blocks(p), (unsigned)p.slices, 1);
    cfg.blockDim = dim3(kBlock, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (clusters) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// Maps (blockIdx, threadIdx) to a grid point; computes the tile's bounding box (all threads must call).
__device__ inline Tile make_tile(const KParams& p, const SceneTab& T, DriverShared& sh) {
    Tile t;
    const int tid = threadIdx.x;
    if (p.grid_cols > 0) {  // (the launcher sets grid_cols / grid_rows for a whole n x m mesh only)
        const int rows = p.grid_rows;
        const unsigned tiles_x = (unsigned)(p.grid_cols + kTileCols - 1) / kTileCols;
        unsigned by = blockIdx.x / tiles_x;
        unsigned bx = blockIdx.x - by * tiles_x;
        if (p.cluster) {  // 8 consecutive CTAs (one cluster) = 2 x 4 neighbouring tiles
            const unsigned macro_x = (tiles_x + kMacroTilesX - 1) / kMacroTilesX;
            const unsigned cid = blockIdx.x / kCluster, rk = blockIdx.x % kCluster;
            const unsigned cy = cid / macro_x, cx = cid - cy * macro_x;
            bx = kMacroTilesX * cx + (rk % kMacroTilesX);
            by = kMacroTilesY * cy + (rk / kMacroTilesX);
        }
        const int col = (int)(bx * kTileCols) + (tid & (kTileCols - 1));
        const long long row = (long long)by * kTileRows + (tid / kTileCols);
        t.active = col < p.grid_cols && row < rows;
        t.r = row * p.grid_cols + col;
    } else {
        t.r = (long long)blockIdx.x * p.tile_points + tid;
        t.active = tid < p.tile_points && t.r < p.R;
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
    t.wbox = make_float4(xmin, ymin, xmax, ymax);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) {
        sh.red[warp][0] = xmin; sh.red[warp][1] = ymin; sh.red[warp][2] = xmax; sh.red[warp][3] = ymax;
    }
    if (lane == 0) sh.hint[warp] = 0;
    __syncthreads();
    xmin = fminf(fminf(sh.red[0][0], sh.red[1][0]), fminf(sh.red[2][0], sh.red[3][0]));
    ymin = fminf(fminf(sh.red[0][1], sh.red[1][1]), fminf(sh.red[2][1], sh.red[3][1]));
    xmax = fmaxf(fmaxf(sh.red[0][2], sh.red[1][2]), fmaxf(sh.red[2][2], sh.red[3][2]));
    ymax = fmaxf(fmaxf(sh.red[0][3], sh.red[1][3]), fmaxf(sh.red[2][3], sh.red[3][3]));
    t.bbox = make_float4(xmin, ymin, xmax, ymax);
    t.mbox = t.bbox;
    if (p.macro) {  // bounding box of the cluster's 8 tiles (empty tiles hold inverted boxes)
        cg::cluster_group cl = cg::this_cluster();
        if (tid == 0) sh.tbox = t.bbox;
        cl.sync();
#pragma unroll
        for (int r = 0; r < kCluster; ++r) {
            const float4 b = *cl.map_shared_rank(&sh.tbox, r);
            xmin = fminf(xmin, b.x); ymin = fminf(ymin, b.y);
            xmax = fmaxf(xmax, b.z); ymax = fmaxf(ymax, b.w);
        }
        t.mbox = make_float4(xmin, ymin, xmax, ymax);
        cl.sync();  // nobody leaves (or reuses tbox) while a neighbour may still read it
    }
    float sx = fmaxf(fabsf(xmin), fabsf(xmax)), sy = fmaxf(fabsf(ymin), fabsf(ymax));
    if (!(sx < CUDART_INF_F)) sx = 0.0f;  // empty tile
    if (!(sy < CUDART_INF_F)) sy = 0.0f;
    float ex0 = xmin, ex1 = xmax, ey0 = ymin, ey1 = ymax;  // extent of everything in play (inverted for an empty tile)
    {
        // objects and fixed points: strided over the CTA, then min / max through shuffles and shared memory (exact and
        // order-free, so every thread ends up with the same bits as the serial loop over all N objects that each
        // thread used to run: 3 % of the forward kernel on the city scene)
        float a0 = 0.f, a1 = 0.f;                                             // max |x|, max |y|
        float b0 = CUDART_INF_F, b1 = -CUDART_INF_F, c0 = CUDART_INF_F, c1 = -CUDART_INF_F;  // x range, y range
        for (int j = tid; j < p.N; j += kBlock) {
            const float4 w = T.w0[j];
            a0 = fmaxf(a0, fmaxf(fabsf(w.x), fabsf(w.x + w.z)));
            a1 = fmaxf(a1, fmaxf(fabsf(w.y), fabsf(w.y + w.w)));
            b0 = fminf(b0, fminf(w.x, w.x + w.z)); b1 = fmaxf(b1, fmaxf(w.x, w.x + w.z));
            c0 = fminf(c0, fminf(w.y, w.y + w.w)); c1 = fmaxf(c1, fmaxf(w.y, w.y + w.w));
        }
        for (int f = tid; f < p.T; f += kBlock) {
            const float fxx = p.fixed[2 * f], fyy = p.fixed[2 * f + 1];
            a0 = fmaxf(a0, fabsf(fxx));
            a1 = fmaxf(a1, fabsf(fyy));
            b0 = fminf(b0, fxx); b1 = fmaxf(b1, fxx);
            c0 = fminf(c0, fyy); c1 = fmaxf(c1, fyy);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 = fmaxf(a0, __shfl_xor_sync(0xffffffffu, a0, o)); a1 = fmaxf(a1, __shfl_xor_sync(0xffffffffu, a1, o));
            b0 = fminf(b0, __shfl_xor_sync(0xffffffffu, b0, o)); b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, o));
            c0 = fminf(c0, __shfl_xor_sync(0xffffffffu, c0, o)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, o));
        }
        if (lane == 0) {
            sh.ext[warp][0] = a0; sh.ext[warp][1] = a1; sh.ext[warp][2] = b0;
            sh.ext[warp][3] = b1; sh.ext[warp][4] = c0; sh.ext[warp][5] = c1;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) {
            sx = fmaxf(sx, sh.ext[w][0]); sy = fmaxf(sy, sh.ext[w][1]);
            ex0 = fminf(ex0, sh.ext[w][2]); ex1 = fmaxf(ex1, sh.ext[w][3]);
            ey0 = fminf(ey0, sh.ext[w][4]); ey1 = fmaxf(ey1, sh.ext[w][5]);
        }
    }
    t.scale_x = sx;
    t.scale_y = sy;
    t.scale = fmaxf(sx, sy);
    t.diam = (ex1 >= ex0 && ey1 >= ey0) ? ((ex1 - ex0) + (ey1 - ey0)) * 1.000001f : 0.0f;
    __syncthreads();
    return t;
}

// Conservative tile test for an ImagePath candidate on a receivers grid: returns false only if the validity
// is EXACTLY 0 for every grid point of the tile, so that skipping the candidate changes nothing, bit for bit.
//
// The per-thread (canonical fp32) evaluation walks the interactions from the last to the first
// (geometry.py:1093-1107): from the current point p (the receiver, then the previous interaction point) and
// the apex A = I_{i+1} (image of the transmitter; computed here with the same `mirror`, hence bit-identical)
//     u = p - A, v = P1 - p, g = (v.n)/(u.n), X = p + g u, s = t.(X - P1)/tt            (geometry.py:589-598)
// and the path is dead (a_on == 0) unless min(s, 1 - s) > xz for every interaction.  s is a linear-fractional
// function of p, so over a convex set of p on which u.n keeps its sign its extrema are at the extreme points:
// the 4 corners of the tile's box for the last interaction, then the 2 end points of the reachable part of
// the wall for the earlier ones.  Three exact rules are applied:
//   (1) s-range: [smin, smax] misses [xz, 1 - xz] by more than the error bound `tol`;
//   (2) wrong side: for walls, the residual of a specular interaction is e = (sign(-g_i) - sign(-g_{i-1})
//       sign(1 + g_i)) u_hat (geometry.py:641-650), i.e. all residuals vanish iff every g_i lies in (-1, 0), and
//       otherwise one of them is |e|^2 = 4: when some g-range is entirely outside [-1, 0] and every segment is
//       long enough (>= Lmin) for the directions to be accurate, loss >= 3.9 > tol - xz and the path is dead;
//   (3) zero-length walls (polygon closure walls of from_geojson, scene.py:646-652): n = 0, X = p, the
//       residual is |i_hat|^2 = 1 unless the previous point coincides bit for bit: dead when that segment is
//       provably longer than Lmin.
// Error bound of s (first order, eps = 2^-24, S = largest coordinate magnitude in play).  Differences of
// fp32 inputs carry RELATIVE errors (u, v exact up to eps|u|, eps|v|), the only absolute term is the final
// rounding of X onto the fp32 lattice, half an ulp of its binade: eps 2^floor(log2(S + |g||u|)) <= eps (S + |g||u|):
//     dX_c <= eps [S + |g||u| + 4|g||u_c| + 4.4 (|u_c|/|un|)(|v|_1 + |g||u|_1)] + L dev
//     ds   <= (|t_x| dX_x + |t_y| dX_y)/tt + 4 eps |s|
// with L = (1 + |g|)(1 + |u|/|un|) the Lipschitz constant of p -> X and dev the distance of the thread's
// computed p from the convex set used here.  The evaluation at the extreme points below forms X - P1 = g u - v
// directly (differences only), so it carries the relative terms but NOT the lattice term: tol = 1.25 (ds_threads +
// ds_here) + 1e-6, about half of what two lattice roundings would need on lon/lat coordinates.
// largest power of two <= x (x > 0, normal)
__device__ __forceinline__ float pow2_floor(const float x) { return __int_as_float(__float_as_int(x) & 0x7f800000); }

template <int K>
__device__ __forceinline__ bool tile_may_be_valid(const SceneTab& T, const int (&c)[K > 0 ? K : 1],
                                                  const float2 (&I)[K + 1], const float4 bbox, const float scale,
                                                  const float scale_x, const float scale_y, const float xz, const float loss_dead /* tol_loss - xz */,
                                                  float& tol_last) {
    tol_last = CUDART_INF_F;
    if (!(xz > -CUDART_INF_F)) return true;
    const float eps = 5.9604645e-8f;
    const float S = scale;
    const float Lmin = fmaxf(4096.0f * eps * S, 1e-15f);
    float2 pts[4] = {make_float2(bbox.x, bbox.y), make_float2(bbox.z, bbox.y), make_float2(bbox.x, bbox.w),
                     make_float2(bbox.z, bbox.w)};
    int npts = 4;
    float dev = 0.0f;
    bool all_walls = true, any_deg = false;
    bool g_out = false, lens_ok = true;
    bool deg_pending = false, deg_dead = false;
    // The stage loop and the corner loop are ROLLED on purpose: these kernels are instruction-fetch bound
    // (profiles/), and the unrolled form made the cull a 24 KB straight line that every warp streamed through
    // once per candidate.  c[i] / I[i+1] are picked with selects so that the arrays stay in registers.
#pragma unroll 1
    for (int i = K - 1; i >= 0; --i) {
        int j = c[0];
        float2 A = I[1];
#pragma unroll
        for (int u = 1; u < K; ++u)
            if (i == u) { j = c[u]; A = I[u + 1]; }
        const float4 w0 = T.w0[j];
        const float4 w1 = T.w1[j];
        const int kind = T.kind[j];
        if (kind != D2D_KIND_WALL) all_walls = false;
        if (kind == D2D_KIND_VERTEX) continue;  // X = p, always on the object (geometry.py:387-403)
        if (w0.z == 0.0f && w0.w == 0.0f) {      // zero-length object: n = 0, un == 0, X = p, s = 0
            any_deg = true;
            deg_pending = true;
            // s = 0 passes on_objects iff act(0) != 0, i.e. xz < 0 (hard: 0 >= 0 is true)
            continue;
        }
        float smin = CUDART_INF_F, smax = -CUDART_INF_F, gmin = CUDART_INF_F, gmax = -CUDART_INF_F;
        float unmin = CUDART_INF_F, U1 = 0.f, V1 = 0.f, uxm = 0.f, uym = 0.f, u2m = 0.f;
        int pos = 0, neg = 0;
        const float rtt = rcp_approx(w1.z);
#pragma unroll 1
        for (int q = 0; q < npts; ++q) {
            const float2 pq = q == 0 ? pts[0] : (q == 1 ? pts[1] : (q == 2 ? pts[2] : pts[3]));
            const float ux = pq.x - A.x, uy = pq.y - A.y;
            const float vx = w0.x - pq.x, vy = w0.y - pq.y;
            const float un = fmaf(ux, w1.x, uy * w1.y);
            const float vn = fmaf(vx, w1.x, vy * w1.y);
            pos += un > 0.f;
            neg += un < 0.f;
            // MUFU-based quotients (relative error < 2^-22 = 4 eps): accounted for in xmag and ds below
            const float g = vn * rcp_approx(un);
            // X - P1 = g u - v, never formed in absolute coordinates: THIS evaluation then carries relative errors
            // only, and the lattice rounding of X enters the tolerance once (the threads' evaluation), not twice
            const float Dx = fmaf(g, ux, -vx), Dy = fmaf(g, uy, -vy);
            const float s = fmaf(w0.z, Dx, w0.w * Dy) * rtt;
            smin = fminf(smin, s); smax = fmaxf(smax, s);
            gmin = fminf(gmin, g); gmax = fmaxf(gmax, g);
            unmin = fminf(unmin, fabsf(un));
            U1 = fmaxf(U1, fabsf(ux) + fabsf(uy));
            V1 = fmaxf(V1, fabsf(vx) + fabsf(vy));
            uxm = fmaxf(uxm, fabsf(ux)); uym = fmaxf(uym, fabsf(uy));
            u2m = fmaxf(u2m, fmaf(ux, ux, uy * uy));
        }
        u2m = sqrt_approx(u2m) * 1.000001f;  // max |u| over the point set, rounded up
        if (!(pos == npts || neg == npts)) { D2D_COUNT(8); return true; }  // u.n may vanish inside the set
        if (!(smin == smin) || !(smax == smax) || !(gmin == gmin) || !(gmax == gmax)) { D2D_COUNT(9); return true; }
        if (!(unmin > 64.0f * eps * U1 + 4.0f * dev)) { D2D_COUNT(10); return true; }  // u.n not reliably away from zero
        const float gabs = fmaxf(fabsf(gmin), fabsf(gmax)) * 1.000001f;
        const float run = rcp_approx(unmin) * 1.000001f;  // 1 / min |u.n|, rounded up
        const float lip = (1.0f + gabs) * (1.0f + u2m * run);
        const float amp = 4.4f * (V1 + gabs * U1) * run;
        const float gu5 = 5.0f * gabs * u2m;  // (4 of the 5: the approximate quotient g at the extreme points)
        // lattice rounding of X per component: |X_c| <= S_c + |g||u|, and rounding a value of that binade moves it by
        // at most half an ulp = 2^-24 * 2^floor(log2 |X_c|) (not 2^-24 |X_c|: up to 2x tighter, 1.6x at latitude 50.7)
        const float gu1 = gabs * u2m;
        // relative-type terms (both evaluations) and the lattice term (the threads' evaluation only)
        // (the deviation `dev` of the thread's previous point from the point set belongs to the threads' evaluation
        // only — the evaluation here starts ON the set — so it is counted once as well)
        const float relx = eps * (0.8f * gu5 + 4.0f * gabs * uxm + amp * uxm + 2.0f * V1);
        const float rely = eps * (0.8f * gu5 + 4.0f * gabs * uym + amp * uym + 2.0f * V1);
        const float devterm = 1.5f * lip * dev;
        const float dXx = eps * pow2_floor(scale_x + gu1) + relx + devterm;
        const float dXy = eps * pow2_floor(scale_y + gu1) + rely + devterm;
        const float smag = fmaxf(fabsf(smin), fabsf(smax));
        const float ds = (fabsf(w0.z) * (dXx + relx) + fabsf(w0.w) * (dXy + rely)) * (rtt * 1.000001f) + 16.0f * eps * smag;
        const float tol = 1.25f * ds + 1e-6f;  // ds = threads' bound + this evaluation's bound
        if (!(tol < CUDART_INF_F)) return true;
        if (i == K - 1) tol_last = tol;  // valid for every point of the tile: reused by warp_may_be_valid
        const float lo = xz - tol, hi = 1.0f - xz + tol;
        if (smax < lo || smin > hi) { D2D_COUNT(i == K - 1 ? 1 : 2); return false; }  // rule (1)
        // g-range and segment lengths for rules (2) and (3)
        const float dg = 1.5f * dev * (1.0f + gabs) * run + 12.0f * eps * (gabs + (V1 + gabs * U1) * run);
        const float glo = gmin - dg, ghi = gmax + dg;
        if (ghi < -1.0f || glo > 0.0f) g_out = true;
        const float g_abs_min = glo > 0.0f ? glo : (ghi < 0.0f ? -ghi : 0.0f);
        const float g1_abs_min = (1.0f + glo) > 0.0f ? (1.0f + glo) : ((1.0f + ghi) < 0.0f ? -(1.0f + ghi) : 0.0f);
        const float un_eff = unmin - 1.5f * dev - 8.0f * eps * U1;
        const float seg_pb = g_abs_min * un_eff;    // |p - X| >= |g| |u.n|
        const float seg_bA = g1_abs_min * un_eff;   // |X - A| >= |1 + g| |u.n|
        // (An error-scaled requirement — 8 x the end points' own error bounds instead of the fixed 4096 eps S, which
        // never holds on lon/lat coordinates — was measured: rule (2) then fires on the raw city scene, and the forward
        // launch got 1 % SLOWER on both coordinate variants, profiles/r02p_ab.txt; the candidates it removes are cheap
        // ones.  Removed.)
        if (!(seg_pb >= Lmin && seg_bA >= Lmin)) lens_ok = false;
        if (deg_pending) {  // the zero-length object(s) between p and this X: previous point is X, distance |p - X|
            // rule (3) only needs the two COMPUTED points to differ (any non-zero vector normalises to |i_hat|^2 =
            // 1 +- 4 ulp): true distance minus both evaluation errors, not the direction accuracy of rule (2)
            if (seg_pb > dXx + dXy + 16.0f * eps * S + 1e-15f) deg_dead = true;
            deg_pending = false;
        }
        // the reachable part of this object becomes the point set of the next (earlier) interaction
        const float a = fmaxf(smin - tol, lo), b = fminf(smax + tol, hi);
        pts[0] = make_float2(fmaf(a, w0.z, w0.x), fmaf(a, w0.w, w0.y));
        pts[1] = make_float2(fmaf(b, w0.z, w0.x), fmaf(b, w0.w, w0.y));
        npts = 2;
        dev = 1.25f * fmaxf(dXx, dXy) + 1.5f * eps * pow2_floor(S);  // + rounding the two points onto the lattice
    }
    // (A direct test of the FIRST interaction through the unfolded path — s_1 over the images of the box corners —
    // was measured on the bench scenes: it rejected 2 of 100 000 candidates the stage loop had kept, i.e. the
    // propagated point sets lose nothing there; removed.)
    if (deg_pending) {  // zero-length object(s) right after the transmitter: previous point is tx = I[0]
        float xl = CUDART_INF_F, xh = -CUDART_INF_F, yl = CUDART_INF_F, yh = -CUDART_INF_F;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q >= npts) break;
            xl = fminf(xl, pts[q].x); xh = fmaxf(xh, pts[q].x);
            yl = fminf(yl, pts[q].y); yh = fmaxf(yh, pts[q].y);
        }
        const float dx = fmaxf(fmaxf(xl - I[0].x, I[0].x - xh), 0.0f);
        const float dy = fmaxf(fmaxf(yl - I[0].y, I[0].y - yh), 0.0f);
        if (fmaxf(dx, dy) - dev > 16.0f * eps * S + 1e-15f) deg_dead = true;
    }
    if (all_walls) {
        if (any_deg) {
            if (deg_dead && 0.98f >= loss_dead) { D2D_COUNT(4); return false; }  // rule (3)
            D2D_COUNT(7);
        } else if (g_out && lens_ok && 3.9f >= loss_dead) {
            D2D_COUNT(5);
            return false;                                           // rule (2)
        }
    }
    D2D_COUNT(6);
    return true;
}

// The same question for a TRANSMITTERS grid (scene.py:1489-1648): the tile holds transmitter positions, `rx` is the fixed
// receiver.  The per-thread evaluation mirrors ITS transmitter through the candidate, I_0 = tx, I_{i+1} = image(I_i, c_i),
// and back-projects from the receiver: X_{K+1} = rx, X_{i+1} = line(X_{i+2}, I_{i+1}) ^ wall c_i, i = K-1 ... 0.
// In exact arithmetic X_{i+1} is also the intersection of wall c_i with the line from I_{i+1}(tx) to the receiver
// UNFOLDED back through the later walls, R_i = image(... image(rx, c_{K-1}) ..., c_{i+1}) (R_{K-1} = rx): a fixed centre
// and an apex that is an AFFINE function of tx.  The parametric coordinate s_i is therefore a linear-fractional
// function of tx for every interaction, and over the tile's box its extrema are at the images of the four corners as
// long as u.n keeps its sign there.  Only rule (1) is applied (s-range misses [xz, 1 - xz] by more than the error
// bound), interaction by interaction in the THREAD'S order, so that the error recursion of the thread's chain is
// available when a stage needs it:
//     E_{i+1} <= 1.5 L_i (E_{i+2} + EI_{i+1}) + own_i,   E_{K+1} = 0,   L_i = (1 + |g|)(1 + |u| / |u.n|)
// with EI_m the rounding of an m-fold image (one lattice rounding per mirror plus relative terms) and own_i the
// rounding of one back-projection (lattice + relative terms, as in tile_may_be_valid; the thread's |P1 - X| is bounded
// by the scene's extent `diam`, its |g||u| and |u| by this frame's, the ratio |u| / |u.n| is the same in both frames).
// This evaluation carries the relative terms and the rounding of ITS images (corner images and R_i).
template <int K>
__device__ __forceinline__ bool tile_may_be_valid_tx(const SceneTab& T, const int (&c)[K > 0 ? K : 1], const float2 rx,
                                                     const float4 bbox, const float scale_x, const float scale_y,
                                                     const float diam, const float xz) {
    constexpr int KK = K > 0 ? K : 1;
    if (!(xz > -CUDART_INF_F)) return true;
    const float eps = 5.9604645e-8f;
    float2 R[KK];
    R[K - 1] = rx;
#pragma unroll
    for (int i = K - 2; i >= 0; --i) R[i] = mirror(R[i + 1], T.w0[c[i + 1]], T.w1[c[i + 1]]);
    float2 Ic[4][KK];  // Ic[q][i]: corner q mirrored through c[0 .. i]
    float simx = scale_x, simy = scale_y;  // largest image coordinates (the lattice the images are rounded onto)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 pt = make_float2((q & 1) ? bbox.z : bbox.x, (q & 2) ? bbox.w : bbox.y);
#pragma unroll
        for (int i = 0; i < K; ++i) {
            pt = mirror(pt, T.w0[c[i]], T.w1[c[i]]);
            Ic[q][i] = pt;
            simx = fmaxf(simx, fabsf(pt.x)); simy = fmaxf(simy, fabsf(pt.y));
        }
    }
#pragma unroll
    for (int i = 0; i < K; ++i) { simx = fmaxf(simx, fabsf(R[i].x)); simy = fmaxf(simy, fabsf(R[i].y)); }
    // rounding of ONE mirror (p - 2 ((p - P1).n) n with |p - P1| <= (2 K + 1) diam): lattice + relative terms
    const float e1 = 1.5f * eps * pow2_floor(fmaxf(simx, simy)) + 16.0f * eps * (2.0f * K + 1.0f) * diam;
    // Frames.  At the stage next to the receiver (i = K - 1) the thread's back-projection IS this evaluation (same centre
    // rx, same apex up to rounding).  At an earlier stage the thread projects from ITS previous point X_{i+2}, which lies
    // on this evaluation's line at the fraction (1 + g') of the way from the apex to the unfolded receiver (g' = this
    // evaluation's g at stage i + 1): the thread's u and u.n are this frame's times (1 + g'), its g is (g - g') / (1 + g').
    // When 1 + g' approaches 0 (the previous point falls onto the image itself: e.g. a transmitter ON the next wall's
    // line) the thread's u.n vanishes, its masked branch (geometry.py:1105) or an unbounded error takes over, and
    // nothing can be said from here: keep the candidate.
    float Enext = 0.0f;
    float opg_prev = 1.0f, gabs_prev = 0.0f;
    bool first = true;
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
        const int j = c[i];
        const float4 w0 = T.w0[j];
        const float4 w1 = T.w1[j];
        const int kind = T.kind[j];
        // (a Vertex or a zero-length object leaves the previous point where it is: X = previous point in both frames;
        // the frame relation above would need the stage before it, so the cull stops refining here)
        if (kind == D2D_KIND_VERTEX || (w0.z == 0.0f && w0.w == 0.0f)) break;
        const float EI = (float)(i + 1) * e1;              // the thread's (and this evaluation's) image of level i + 1
        const float ER = (float)(K - 1 - i) * e1;          // this evaluation's unfolded receiver
        float smin = CUDART_INF_F, smax = -CUDART_INF_F, gabs = 0.f, opg = CUDART_INF_F;
        float unmin = CUDART_INF_F, U1 = 0.f, V1 = 0.f, uxm = 0.f, uym = 0.f, u2m = 0.f;
        int pos = 0, neg = 0;
        const float rtt = rcp_approx(w1.z);
        const float2 pc = R[i];
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const float2 A = q == 0 ? Ic[0][i] : (q == 1 ? Ic[1][i] : (q == 2 ? Ic[2][i] : Ic[3][i]));
            const float ux = pc.x - A.x, uy = pc.y - A.y;
            const float vx = w0.x - pc.x, vy = w0.y - pc.y;
            const float un = fmaf(ux, w1.x, uy * w1.y);
            const float vn = fmaf(vx, w1.x, vy * w1.y);
            pos += un > 0.f;
            neg += un < 0.f;
            const float g = vn * rcp_approx(un);
            const float Dx = fmaf(g, ux, -vx), Dy = fmaf(g, uy, -vy);  // X - P1 in relative form
            const float sq = fmaf(w0.z, Dx, w0.w * Dy) * rtt;
            smin = fminf(smin, sq); smax = fmaxf(smax, sq);
            gabs = fmaxf(gabs, fabsf(g));
            opg = fminf(opg, fabsf(1.0f + g));
            unmin = fminf(unmin, fabsf(un));
            U1 = fmaxf(U1, fabsf(ux) + fabsf(uy));
            V1 = fmaxf(V1, fabsf(vx) + fabsf(vy));
            uxm = fmaxf(uxm, fabsf(ux)); uym = fmaxf(uym, fabsf(uy));
            u2m = fmaxf(u2m, fmaf(ux, ux, uy * uy));
        }
        u2m = sqrt_approx(u2m) * 1.000001f;
        if (!(pos == 4 || neg == 4)) return true;                          // u.n may vanish inside the tile
        if (!(smin == smin) || !(smax == smax) || !(gabs == gabs) || !(opg == opg)) return true;
        gabs *= 1.000001f;
        // the thread's frame at this stage
        const float f = first ? 1.0f : opg_prev;
        const float unminT = f * unmin;
        if (!(unmin > 64.0f * eps * U1 + 4.0f * (EI + ER))) return true;           // this frame's u.n not reliable
        if (!(unminT > 64.0f * eps * U1 * fmaxf(f, 1.0f) + 8.0f * (EI + Enext))) return true;  // the thread's u.n not reliable
        const float run = rcp_approx(unmin) * 1.000001f;
        const float rho = u2m * run;                                               // |u| / |u.n|: the same in both frames
        const float gsum = first ? gabs : (gabs + gabs_prev);
        const float gabsT = first ? gabs : gsum * rcp_approx(f) * 1.000001f;       // |g_thread| <= (|g| + |g'|) / |1 + g'|
        const float ellT = gsum * u2m;                                             // |g_thread| |u_thread|
        const float lipT = (1.0f + gabsT) * (1.0f + rho);
        const float lipC = (1.0f + gabs) * (1.0f + rho);
        // the thread's back-projection: |v|_1 <= 2 diam, |u_c| / |u.n| frame invariant
        const float ampT = 4.4f * (2.0f * diam + 1.5f * ellT) * run;
        const float relTx = eps * (8.0f * ellT + ampT * uxm + 4.0f * diam);
        const float relTy = eps * (8.0f * ellT + ampT * uym + 4.0f * diam);
        const float devT = 1.5f * lipT * (Enext + EI);
        const float dTx = eps * pow2_floor(scale_x + ellT) + relTx + devT;
        const float dTy = eps * pow2_floor(scale_y + ellT) + relTy + devT;
        // this evaluation: relative terms with ITS v, and the rounding of its two images
        const float gu1 = gabs * u2m;
        const float ampC = 4.4f * (V1 + gabs * U1) * run;
        const float relCx = eps * (4.0f * gu1 + 4.0f * gabs * uxm + ampC * uxm + 2.0f * V1);
        const float relCy = eps * (4.0f * gu1 + 4.0f * gabs * uym + ampC * uym + 2.0f * V1);
        const float devC = 1.5f * lipC * (EI + ER);
        const float smag = fmaxf(fabsf(smin), fabsf(smax));
        const float ds = (fabsf(w0.z) * (dTx + relCx + devC) + fabsf(w0.w) * (dTy + relCy + devC)) * (rtt * 1.000001f) +
                         16.0f * eps * smag;
        const float tol = 2.0f * (1.25f * ds + 1e-6f);  // (x 2: this bound has had less mileage than the receivers-grid one)
        if (!(tol < CUDART_INF_F)) return true;
        if (smax < xz - tol || smin > 1.0f - xz + tol) { D2D_COUNT(i == K - 1 ? 1 : 2); return false; }  // rule (1)
        Enext = fmaxf(dTx, dTy) * 1.25f;  // the thread's X_{i+1}
        // |1 + g| over the tile, less what the rounding of g can take away
        opg_prev = opg - (8.0f * eps * (1.0f + gabs) + 2.0f * (EI + ER) * run * (1.0f + gabs));
        gabs_prev = gabs;
        first = false;
    }
    D2D_COUNT(6);
    return true;
}

// Warp-level refinement of rule (1) for the LAST interaction: the s-range over the warp's own bounding box
// (a sub-box of the tile's), evaluated exactly like the tile-level test at its corners; `tol` is the tile-level
// error bound (every quantity it is built from is a maximum / minimum over the whole tile, and the tile-level
// test has already established that u.n keeps its sign there).  false => validity exactly 0 for all 32 points.
// (A LANE-level version of the same test — the approximate s at each thread's own point — was measured on the bench
// scene: it kept 94 % of the lanes and emptied 105 of 2.6 M warp visits.  What reaches a warp lies inside the band
// |s - {0,1}| < tol that no approximate evaluation can decide, or dies at an earlier interaction; removed.)
__device__ __forceinline__ bool warp_may_be_valid(const float4 w0, const float4 w1, const float2 A, const float4 box,
                                                  const float tol, const float xz) {
    if (!(tol < CUDART_INF_F)) return true;
    float smin = CUDART_INF_F, smax = -CUDART_INF_F;
    bool nan = false;
    const float rtt = rcp_approx(w1.z);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float px = (q & 1) ? box.z : box.x, py = (q & 2) ? box.w : box.y;
        const float ux = px - A.x, uy = py - A.y;
        const float vx = w0.x - px, vy = w0.y - py;
        const float un = fmaf(ux, w1.x, uy * w1.y);
        const float vn = fmaf(vx, w1.x, vy * w1.y);
        const float g = vn * rcp_approx(un);  // (MUFU quotients: their 2^-22 relative error is a term of `tol`)
        const float Dx = fmaf(g, ux, -vx), Dy = fmaf(g, uy, -vy);  // X - P1 in relative form, as in the tile-level test
        const float s = fmaf(w0.z, Dx, w0.w * Dy) * rtt;
        nan = nan || !(s == s);
        smin = fminf(smin, s);
        smax = fmaxf(smax, s);
    }
    if (nan) return true;
    return !(smax < xz - tol || smin > 1.0f - xz + tol);
}

// One candidate: decode + image chain + tile test.
struct CandTest {
    int c01, c23;     // objects, 16 bits each
    float2 apex;      // image of the fixed point through all objects (ImagePath on a receivers grid)
    float tol_last;   // error bound of the last interaction's s (+inf: none)
    bool keep;
};

template <int K>
__device__ __forceinline__ CandTest test_candidate_inline(const SceneTab& T, const int m, const long long idx,
                                                          const float2 fx, const bool apex_wanted, const bool cull,
                                                          const float4 box, const float scale, const float scale_x,
                                                          const float scale_y, const float xz, const float loss_dead,
                                                          const bool txgrid = false, const float diam = 0.0f) {
    constexpr int KK = K > 0 ? K : 1;
    CandTest o;
    o.apex = fx;
    o.tol_last = CUDART_INF_F;
    o.keep = true;
    // index -> positions in `allowed` (lexicographic, no equal neighbours)
    int c[KK];
    int dig[KK];
    if ((idx >> 32) == 0) {  // (uniform in practice) 32-bit quotients: a 64-bit division is ~5x the instructions
        unsigned rem = (unsigned)idx;
        const unsigned dv = (unsigned)(m - 1);
#pragma unroll
        for (int i = K - 1; i >= 1; --i) {
            const unsigned q = rem / dv;
            dig[i] = (int)(rem - q * dv);
            rem = q;
        }
        dig[0] = (int)rem;
    } else {
        long long rem = idx;
#pragma unroll
        for (int i = K - 1; i >= 1; --i) {
            dig[i] = (int)(rem % (m - 1));
            rem /= (m - 1);
        }
        dig[0] = (int)rem;
    }
    int prev = -1;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        int pos = dig[i];
        if (i > 0 && pos >= prev) ++pos;
        c[i] = T.allowed[pos];
        prev = pos;
    }
    o.c01 = c[0] | ((K > 1 ? c[K > 1 ? 1 : 0] : 0) << 16);
    o.c23 = (K > 2 ? c[K > 2 ? 2 : 0] : 0) | ((K > 3 ? c[K > 3 ? 3 : 0] : 0) << 16);
    if (txgrid) {  // transmitters grid: the images depend on the grid point; fx is the fixed RECEIVER
        if (cull) {
            D2D_COUNT(0);
            if constexpr (K > 0) o.keep = tile_may_be_valid_tx<K>(T, c, fx, box, scale_x, scale_y, diam, xz);
        }
    } else if (apex_wanted) {
        float2 I[K + 1];
        I[0] = fx;
#pragma unroll
        for (int i = 0; i < K; ++i) I[i + 1] = mirror(I[i], T.w0[c[i]], T.w1[c[i]]);
        o.apex = I[K];
        if (cull) {
            D2D_COUNT(0);
            o.keep = tile_may_be_valid<K>(T, c, I, box, scale, scale_x, scale_y, xz, loss_dead, o.tol_last);
        }
    }
    return o;
}

// The out-of-line copy serves the macro stage (a prologue that runs once per kernel): the chunk loop keeps its own
// inlined copy, so the hot loop's instruction footprint is what it was before the macro stage existed.
template <int K>
__device__ __noinline__ CandTest test_candidate(unsigned char* smem, const int N, const int m, const long long idx,
                                                const float2 fx, const bool apex_wanted, const bool cull,
                                                const float4 box, const float scale, const float scale_x,
                                                const float scale_y, const float xz, const float loss_dead,
                                                const bool txgrid, const float diam) {
    const SceneTab T = carve_tab(smem, N);
    return test_candidate_inline<K>(T, m, idx, fx, apex_wanted, cull, box, scale, scale_x, scale_y, xz, loss_dead,
                                    txgrid, diam);
}

// number of candidates of order K over m visitable objects
__device__ __forceinline__ long long order_count(const int K, const int m) {
    if (K == 0) return 1;
    if (m <= 0 || (K > 1 && m < 2)) return 0;
    long long c = m;
    for (int i = 1; i < K; ++i) c *= (m - 1);
    return c;
}

// set bits before each word of sh.mbits[0 .. nwords) -> sh.mpref (warp 0; every thread of the CTA must call)
__device__ __forceinline__ void bitmap_prefix(DriverShared& sh, const int nwords) {
    const int lane = threadIdx.x & 31;
    if ((threadIdx.x >> 5) == 0) {
        int run = 0;
        for (int w0 = 0; w0 < nwords; w0 += 32) {
            const int c = (w0 + lane < nwords) ? __popc(sh.mbits[w0 + lane]) : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (w0 + lane < nwords) sh.mpref[w0 + lane] = (unsigned short)(run + inc - c);
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            sh.mpref[nwords] = (unsigned short)run;
            sh.mbits[nwords] = 0u;
        }
    }
    __syncthreads();
}

// Macro stage (p.macro: the kernel runs as clusters of 8 CTAs): the survivor bitmap of the cluster's macro tile, built
// cooperatively and pushed into every CTA's shared memory; must be called by every thread of every CTA, once, after
// make_tile().  Word wd of the bitmap is the ballot of one warp over the pairs 32 wd ... 32 wd + 31, pair b = fixed point
// b / C_total, candidate column b % C_total (orders ascending, list order inside an order).
template <int MODE, bool TXGRID = false>
__device__ __forceinline__ void macro_prologue(const SceneTab& T, const KParams& p, const Tile& tile, DriverShared& sh,
                                               const float alpha) {
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long Ct = p.C_total, nbits = (long long)p.T * Ct;
    const int nwords = (int)((nbits + 31) >> 5);
    const float xz = x_zero<MODE>(alpha);
    const int m = T.n_allowed;
    unsigned char* const smem_tab = reinterpret_cast<unsigned char*>(T.w0);
#pragma unroll 1
    for (int wd = rank * (kBlock / 32) + warp; wd < nwords; wd += kCluster * (kBlock / 32)) {
        const long long b = ((long long)wd << 5) + lane;
        bool k = false;
        if (b < nbits) {
            const int t = (int)(b / Ct);
            long long col = b - (long long)t * Ct;
            const float2 fx = reinterpret_cast<const float2*>(p.fixed)[t];
            int ord = p.min_order;
            for (; ord < p.max_order; ++ord) {  // order of column `col`
                const long long c = order_count(ord, m);
                if (col < c) break;
                col -= c;
            }
            D2D_COUNT(20);
            const float ld = p.tol - xz;
            switch (ord) {
                case 1: k = test_candidate<1>(smem_tab, p.N, m, col, fx, !TXGRID, true, tile.mbox, tile.scale, tile.scale_x, tile.scale_y, xz, ld, TXGRID, tile.diam).keep; break;
                case 2: k = test_candidate<2>(smem_tab, p.N, m, col, fx, !TXGRID, true, tile.mbox, tile.scale, tile.scale_x, tile.scale_y, xz, ld, TXGRID, tile.diam).keep; break;
                case 3: k = test_candidate<3>(smem_tab, p.N, m, col, fx, !TXGRID, true, tile.mbox, tile.scale, tile.scale_x, tile.scale_y, xz, ld, TXGRID, tile.diam).keep; break;
                case 4: k = test_candidate<4>(smem_tab, p.N, m, col, fx, !TXGRID, true, tile.mbox, tile.scale, tile.scale_x, tile.scale_y, xz, ld, TXGRID, tile.diam).keep; break;
                default: k = true; break;  // order 0: line of sight, never culled
            }
            if (k) D2D_COUNT(21);
        }
        const unsigned word = __ballot_sync(0xffffffffu, k);
        if (lane < kCluster) cl.map_shared_rank(&sh.mbits[0], lane)[wd] = word;  // one copy per CTA of the cluster
    }
    cl.sync();  // the bitmap is complete in every CTA; no remote access after this point
    bitmap_prefix(sh, nwords);
}

// Backward kernel after a forward that wrote the activity mask: the same bitmap, filled from the mask rows of this CTA's
// four warps (bit = some lane of the CTA found the candidate valid), so that the chunk loop walks the handful of marked
// candidates instead of testing every column of the list.  Rows are word aligned: bit (t * 32 wpw + column).
__device__ __forceinline__ bool mask_bitmap_fits(const KParams& p) { return (long long)p.T * p.mask_wpw <= kMacroWords; }

__device__ __forceinline__ void mask_prologue(const KParams& p, DriverShared& sh) {
    const int wpw = (int)p.mask_wpw, nwords = p.T * wpw;
    for (int w = threadIdx.x; w < nwords; w += kBlock) {
        const int t = w / wpw, ww = w - t * wpw;
        const uint32_t* row = p.mask + ((long long)t * gridDim.x * (kBlock / 32) + (long long)blockIdx.x * (kBlock / 32)) * wpw + ww;
        sh.mbits[w] = row[0] | row[wpw] | row[2 * wpw] | row[3 * wpw];
    }
    __syncthreads();
    bitmap_prefix(sh, nwords);
}

// Walks all candidates of order K for the fixed point `fx`; calls visit(cd, col, apex) — uniformly over a
// WARP — for every candidate that survives the tile and warp culls, in list order.  `col0` = column of the first
// candidate of this order in the global list; `apex` = image of fx through the candidate's objects (ImagePath on a
// receivers grid only, otherwise unspecified).
// `t` = index of the fixed point (rows of the macro bitmap).
// `mread` (backward kernel after a forward that wrote the activity mask): rows of this CTA's four warps for the
// current fixed point; the stored bits then REPLACE both culls — only candidates some warp of the CTA found
// valid are decoded, and each warp only visits its own set bits.
template <int MODE, int METHOD, int K, bool TXGRID, class Visit>
__device__ __forceinline__ void for_each_candidate(const SceneTab& T, const KParams& p, const Tile& tile,
                                                   DriverShared& sh, const float alpha, const int t, const float2 fx,
                                                   const long long col0, const uint32_t* __restrict__ mread, int& buf,
                                                   Visit&& visit) {
    const int m = T.n_allowed;
    const long long Ck = order_count(K, m);
    if (Ck == 0) return;
    if constexpr (K == 0) {
        Cand<0> cd;
        cd.c[0] = 0;
        if (blockIdx.y == 0 && p.shard_index == 0) visit(cd, col0, fx);  // uniform over the CTA
        return;
    }
    constexpr int KK = K > 0 ? K : 1;
    constexpr bool kApex = (METHOD == D2D_METHOD_IMAGE) && !TXGRID;
    constexpr bool kTxCull = (METHOD == D2D_METHOD_IMAGE) && TXGRID;  // transmitters grid: tile_may_be_valid_tx
    const bool cull = (kApex || kTxCull) && p.cull && !mread;
    const long long wpw = p.mask_wpw;
    const float xz = x_zero<MODE>(alpha);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // a survivor bitmap (macro cull, or the activity mask in the backward kernel) replaces the walk over all columns
    const bool use_macro = (cull && p.macro) || (mread && mask_bitmap_fits(p));  // uniform over the CTA

    // One chunk of at most kBlock candidates, one per thread (`pre` = this thread has one), in list order across the
    // CTA: tile cull, ordered compaction, warp cull, visits.  Must be called by every thread of the CTA.
    auto chunk = [&](const long long idx, bool pre) {
        CandTest ct;
        ct.c01 = ct.c23 = 0;
        ct.apex = fx;
        ct.tol_last = CUDART_INF_F;
        ct.keep = false;
        if (pre && mread && !use_macro) {
            const long long wd = (col0 + idx) >> 5;
            const uint32_t any4 = mread[wd] | mread[wpw + wd] | mread[2 * wpw + wd] | mread[3 * wpw + wd];
            pre = (any4 >> ((col0 + idx) & 31)) & 1u;
        }
        if (pre)
            ct = test_candidate_inline<K>(T, m, idx, fx, kApex, cull, tile.bbox, tile.scale, tile.scale_x, tile.scale_y,
                                          xz, p.tol - xz, kTxCull, tile.diam);
        const bool keep = ct.keep;
        // ordered compaction: per-warp segments keep list order
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int off = __popc(ballot & ((1u << lane) - 1u));
            sh.list[buf][warp * 32 + off] = make_int4(ct.c01, ct.c23, (int)(idx & 0xffffffffLL), (int)(idx >> 32));
            if (kApex) sh.aux[buf][warp * 32 + off] = make_float4(ct.apex.x, ct.apex.y, ct.tol_last, 0.f);
        }
        if (lane == 0) sh.wcount[buf][warp] = __popc(ballot);
        __syncthreads();
        // survivors of the chunk, addressed contiguously across the four per-warp segments (list order)
        const int n0 = sh.wcount[buf][0], n1 = n0 + sh.wcount[buf][1], n2 = n1 + sh.wcount[buf][2],
                  n3 = n2 + sh.wcount[buf][3];
        auto slot_of = [&](const int q) {
            const int w = (q >= n0) + (q >= n1) + (q >= n2);
            return w * 32 + (q - (w == 0 ? 0 : (w == 1 ? n0 : (w == 2 ? n1 : n2))));
        };
#pragma unroll 1
        for (int q0 = 0; q0 < n3; q0 += 32) {
            const int nq = min(32, n3 - q0);
            unsigned todo = nq >= 32 ? 0xffffffffu : ((1u << nq) - 1u);
            const int myslot = lane < nq ? slot_of(q0 + lane) : 0;  // lane q holds the slot of survivor q0 + q
            if (mread) {  // this warp's own bits
                bool wk = false;
                if (lane < nq) {
                    const int4 e = sh.list[buf][myslot];
                    const long long col = col0 + (((long long)(unsigned)e.z) | ((long long)e.w << 32));
                    wk = (mread[warp * wpw + (col >> 5)] >> (col & 31)) & 1u;
                }
                todo = __ballot_sync(0xffffffffu, wk);
            } else if (cull && kApex) {  // warp-level refinement: lane q tests survivor q0 + q against this warp's box
                // (Running the COMPLETE conservative test again on the warp's own 16 x 2 box — every stage, every rule,
                // through the out-of-line copy the macro stage uses — was measured: with a handful of survivors per
                // chunk its lanes are mostly idle, and it cost 0.17 ms per launch on either bench leg for 0.02 ms of
                // visits saved on the raw city scene, profiles/r02p_ab.txt; removed.)
                bool wk = false;
                if (lane < nq && tile.wbox.x <= tile.wbox.z) {  // (a warp without active points skips everything)
                    const int sl = myslot;
                    const int4 e = sh.list[buf][sl];
                    const float4 a = sh.aux[buf][sl];
                    int jl = e.x & 0xffff;
                    if (K == 2) jl = (e.x >> 16) & 0xffff;
                    if (K == 3) jl = e.y & 0xffff;
                    if (K == 4) jl = (e.y >> 16) & 0xffff;
                    wk = warp_may_be_valid(T.w0[jl], T.w1[jl], make_float2(a.x, a.y), tile.wbox, a.z, xz);
                }
                todo = __ballot_sync(0xffffffffu, wk);
#ifdef D2D_DEBUG_COUNTERS
                if (lane == 0) { atomicAdd(&d2d_dbg[13], (unsigned long long)nq); atomicAdd(&d2d_dbg[14], (unsigned long long)__popc(todo)); }
#endif
            }
#pragma unroll 1
            while (todo) {
                const int q = __ffs(todo) - 1;
                todo &= todo - 1;
                const int sl = __shfl_sync(0xffffffffu, myslot, q);
                const int4 e = sh.list[buf][sl];
                Cand<K> cd;
                cd.c[0] = e.x & 0xffff;
                if (K > 1) cd.c[K > 1 ? 1 : 0] = (e.x >> 16) & 0xffff;
                if (K > 2) cd.c[K > 2 ? 2 : 0] = e.y & 0xffff;
                if (K > 3) cd.c[K > 3 ? 3 : 0] = (e.y >> 16) & 0xffff;
                const long long ci = ((long long)(unsigned)e.z) | ((long long)e.w << 32);
                float2 ax = fx;
                if (kApex) {
                    const float4 a = sh.aux[buf][sl];
                    ax = make_float2(a.x, a.y);
                }
                visit(cd, col0 + ci, ax);
            }
        }
        buf ^= 1;  // the next chunk fills the other buffer: one barrier per chunk
    };

    // With a bitmap (macro_prologue / mask_prologue) the chunk loop walks the survivors of this (fixed point, order) only: the
    // g-th set bit of the range [B0, B0 + Ck), found through the per-word prefix counts.  (ONE loop serves both forms,
    // so that chunk() — and with it the whole trace — is instantiated once.)
    const long long B0 = (long long)t * (mread ? 32 * wpw : p.C_total) + col0;
    auto set_before = [&](const long long b) {  // set bits of the bitmap below bit b
        const int w = (int)(b >> 5);
        return (int)sh.mpref[w] + __popc(sh.mbits[w] & ((1u << (b & 31)) - 1u));
    };
    long long G = Ck;  // length of the list the chunk loop walks
    int r0 = 0;
    if (use_macro) {
        r0 = set_before(B0);
        G = (tile.bbox.x <= tile.bbox.z) ? set_before(B0 + Ck) - r0 : 0;  // (a CTA beyond the grid's edge traces nothing)
    }
    // candidate slices (point-to-point links with huge candidate lists): CTA y walks chunks y, y + slices, ...;
    // candidate SHARDS (one per GPU) deal the chunks round robin one level up: virtual slice shard * slices + y of
    // shards * slices.  The activity mask of a sharded forward only holds this shard's bits, so the mask-driven
    // backward walks all of them.
    const long long vs = mread ? (long long)blockIdx.y : (long long)p.shard_index * gridDim.y + blockIdx.y;
    const long long nvs = mread ? (long long)gridDim.y : (long long)p.shard_count * gridDim.y;
#pragma unroll 1
    for (long long g0 = vs * kBlock; g0 < G; g0 += (long long)kBlock * nvs) {
        const long long g = g0 + tid;
        long long idx = g;
        if (use_macro && g < G) {
            const int target = r0 + (int)g;
            int lo = (int)(B0 >> 5), hi = (int)((B0 + Ck - 1) >> 5);
            while (lo < hi) {  // first word whose cumulative count exceeds the target
                const int mid = (lo + hi) >> 1;
                if ((int)sh.mpref[mid + 1] > target) hi = mid;
                else lo = mid + 1;
            }
            const unsigned bit = __fns(sh.mbits[lo], 0u, target - (int)sh.mpref[lo] + 1);
            idx = (((long long)lo << 5) + bit) - B0;
        }
        chunk(idx, g < G);
    }
}

}  // namespace d2d
