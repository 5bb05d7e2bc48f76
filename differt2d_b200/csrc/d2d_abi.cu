// differt2d_b200 — the C ABI (include/differt2d_b200.h): argument checking, parameter packing,
// candidate enumeration (host + integer CUDA kernel), host-buffer staging, FP32 peak probe.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "d2d_launch.h"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int cuda_fail(int e, const char* where) {
    return fail(D2D_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString((cudaError_t)e));
}

// allowed (visitable) objects in ascending order — scene.py:158-160 disconnects the filtered nodes
int build_blocked(int n, const int32_t* filter, int n_filter, uint32_t* blocked /*[D2D_MAX_OBJECTS/32]*/) {
    std::memset(blocked, 0, sizeof(uint32_t) * (D2D_MAX_OBJECTS / 32));
    if (n_filter < 0 || (n_filter > 0 && !filter)) return -1;
    for (int i = 0; i < n_filter; ++i) {
        const int j = filter[i];
        if (j < 0 || j >= n) continue;  // disconnecting a node that does not exist is a no-op
        blocked[j >> 5] |= 1u << (j & 31);
    }
    int m = 0;
    for (int j = 0; j < n; ++j)
        if (!((blocked[j >> 5] >> (j & 31)) & 1u)) ++m;
    return m;
}

long long count_for(int m, int order) {
    if (order == 0) return 1;
    if (m <= 0) return 0;
    long long c = m;
    for (int i = 1; i < order; ++i) {
        c *= (m - 1);
        if (c > (1LL << 62) / (m > 1 ? m : 2)) return -2;  // overflow guard
    }
    return c;
}

struct BlockedMask {
    uint32_t w[D2D_MAX_OBJECTS / 32];
};

// Integer kernel: candidate index -> object-index sequence (lexicographic, no equal neighbours).
// idx = sum_i digit_i * (m-1)^(k-1-i), digit_0 in [0,m), digit_i in [0,m-1);
// position_i = digit_i + (digit_i >= position_{i-1}); object = allowed[position_i].
__global__ void candidates_kernel(const BlockedMask mask, const int n, const int m, const int order,
                                  const long long count, int32_t* __restrict__ out) {
    __shared__ short allowed[D2D_MAX_OBJECTS];
    if (threadIdx.x == 0) {
        int q = 0;
        for (int j = 0; j < n; ++j)
            if (!((mask.w[j >> 5] >> (j & 31)) & 1u)) allowed[q++] = (short)j;
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += stride) {
        long long rem = idx;
        int digits[D2D_MAX_ORDER];
        for (int i = order - 1; i >= 1; --i) {
            digits[i] = (int)(rem % (m - 1));
            rem /= (m - 1);
        }
        digits[0] = (int)rem;
        int prev = -1;
        for (int i = 0; i < order; ++i) {
            int pos = digits[i];
            if (i > 0 && pos >= prev) ++pos;
            out[idx * order + i] = allowed[pos];
            prev = pos;
        }
    }
}

// register-resident FMA chains: 16 independent accumulators per thread
__global__ void __launch_bounds__(256) fma_peak_kernel(float* sink, int iters) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i);
    const float x = 0.999f + 1e-7f * (float)blockIdx.x, y = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) sink[0] = s;  // never true in practice; keeps the chains alive
}

int pack(const D2DProblem* p, d2d::KParams& k) {
    if (!p) return fail(D2D_ERR_INVALID_ARGUMENT, "problem is NULL");
    if (p->n_objects < 0 || p->n_objects > D2D_MAX_OBJECTS)
        return fail(D2D_ERR_UNSUPPORTED, "n_objects outside [0, D2D_MAX_OBJECTS]");
    if (p->n_objects > 0 && !p->objects_xys) return fail(D2D_ERR_INVALID_ARGUMENT, "objects_xys is NULL");
    if (p->min_order < 0 || p->max_order < p->min_order)
        return fail(D2D_ERR_INVALID_ARGUMENT, "need 0 <= min_order <= max_order");
    if (p->max_order > D2D_MAX_ORDER) return fail(D2D_ERR_UNSUPPORTED, "max_order > D2D_MAX_ORDER");
    if (p->n_fixed < 0 || p->n_grid < 0) return fail(D2D_ERR_INVALID_ARGUMENT, "negative sizes");
    if ((p->n_fixed > 0 && !p->fixed_xy) || (p->n_grid > 0 && !p->grid_xy))
        return fail(D2D_ERR_INVALID_ARGUMENT, "fixed_xy / grid_xy is NULL");
    if (p->grid_role != D2D_GRID_RECEIVERS && p->grid_role != D2D_GRID_TRANSMITTERS)
        return fail(D2D_ERR_INVALID_ARGUMENT, "bad grid_role");
    if (p->mode < D2D_MODE_HARD || p->mode > D2D_MODE_SIGMOID) return fail(D2D_ERR_INVALID_ARGUMENT, "bad mode");
    if (p->method < D2D_METHOD_IMAGE || p->method > D2D_METHOD_MINPATH)
        return fail(D2D_ERR_INVALID_ARGUMENT, "bad method");
    if (p->fun != D2D_FUN_RECEIVED_POWER && p->fun != D2D_FUN_LENGTH_SQUARED)
        return fail(D2D_ERR_UNSUPPORTED, "fun must be received_power or length_squared");
    if (p->mode != D2D_MODE_HARD && !p->alpha_dev && !(p->alpha > 0.0f))
        return fail(D2D_ERR_INVALID_ARGUMENT, "alpha must be > 0 (activations must be non-decreasing)");
    if (p->method != D2D_METHOD_IMAGE && p->steps < 1)
        return fail(D2D_ERR_INVALID_ARGUMENT, "steps must be >= 1 (optimize.py:96 indexes losses[-1])");
    if (p->optimizer < D2D_OPT_ADAM || p->optimizer > D2D_OPT_NEWTON) return fail(D2D_ERR_INVALID_ARGUMENT, "bad optimizer");
    if (p->optimizer == D2D_OPT_NEWTON && p->many > 1)
        return fail(D2D_ERR_UNSUPPORTED, "D2D_OPT_NEWTON does not support restarts (many > 1)");
    if (p->grad_mode != D2D_GRAD_CLEAN && p->grad_mode != D2D_GRAD_NAN_PARITY)
        return fail(D2D_ERR_INVALID_ARGUMENT, "bad grad_mode");
    if (p->grad_mode == D2D_GRAD_NAN_PARITY && p->method != D2D_METHOD_IMAGE)
        return fail(D2D_ERR_UNSUPPORTED, "D2D_GRAD_NAN_PARITY covers ImagePath only");
    std::memset(&k, 0, sizeof(k));
    k.xys = p->objects_xys;
    k.kinds = p->object_kinds;
    k.phis = p->object_phis;
    k.fixed = p->fixed_xy;
    k.grid = p->grid_xy;
    k.x0 = p->x0;
    k.alpha_dev = p->alpha_dev;
    k.R = p->n_grid;
    k.grid_cols = (p->grid_cols > 0 && p->n_grid % p->grid_cols == 0) ? p->grid_cols : 0;
    if (k.grid_cols > 0 && p->n_grid / k.grid_cols > 0x3fffffffLL) k.grid_cols = 0;  // (rows as an int in the kernels)
    k.grid_rows = k.grid_cols > 0 ? (int)(p->n_grid / k.grid_cols) : 0;
    k.cull = p->no_cull ? 0 : 4;  // non-zero: safety factor applied to the cull's first-order error bound
    k.N = p->n_objects;
    k.T = p->n_fixed;
    k.min_order = p->min_order;
    k.max_order = p->max_order;
    k.steps = p->steps;
    k.many = p->many > 1 ? p->many : 1;
    k.fun = p->fun;
    k.reduce_all = p->reduce_all ? 1 : 0;
    k.alpha = p->alpha;
    k.tol = p->tol;
    k.patch = p->patch;
    k.lr = p->lr;
    k.opt = p->optimizer;
    k.b1 = p->opt_b1;
    k.b2 = p->opt_b2;
    k.opt_eps = p->opt_eps;
    // utils.py:52-54 — Python folds r_coef**n and height*height in double, JAX then casts to f32
    k.h2 = (float)(p->height * p->height);
    for (int i = 0; i <= d2d::kMaxOrder; ++i) k.rc_pow[i] = (float)std::pow(p->r_coef, (double)i);
    const int m = build_blocked(p->n_objects, p->filter_nodes, p->n_filter, k.blocked);
    if (m < 0) return fail(D2D_ERR_INVALID_ARGUMENT, "bad filter_nodes");
    long long total = 0;
    for (int o = p->min_order; o <= p->max_order; ++o) {
        const long long c = count_for(m, o);
        if (c < 0) return fail(D2D_ERR_UNSUPPORTED, "candidate count overflows int64");
        total += c;
    }
    k.C_total = total;
    // Launch geometry.  Links with very few grid points (point-to-point, scene.py:1272-1334) get one point per
    // CTA (the tile cull is then exact up to rounding) and share their candidate list among `slices` CTAs.
    k.tile_points = 128;
    k.slices = 1;
    if (k.grid_cols == 0 && p->n_grid <= 32 && total >= 4096) k.tile_points = 1;
    if (p->candidate_slices > 0) {
        k.slices = p->candidate_slices;
    } else if (total >= 16384) {
        long long nblk = k.grid_cols > 0 ? ((k.grid_cols + 15) / 16) * ((p->n_grid / k.grid_cols + 7) / 8)
                                         : (p->n_grid + k.tile_points - 1) / k.tile_points;
        if (nblk < 1) nblk = 1;
        long long want = (148LL * 8) / nblk;                      // fill the machine about 8 CTAs deep
        const int shards = p->cand_shard_count > 1 ? p->cand_shard_count : 1;
        const long long chunks = (total + 127) / 128 / shards;    // (this GPU's share when the list is sharded)
        if (want > chunks / 4) want = chunks / 4;                  // at least 4 chunks per CTA
        if (want > 65535) want = 65535;
        if (want > 1) k.slices = (int)want;
    }
    if (k.slices > 65535) return fail(D2D_ERR_INVALID_ARGUMENT, "candidate_slices must be <= 65535");
    k.shard_count = p->cand_shard_count > 1 ? p->cand_shard_count : 1;
    k.shard_index = p->cand_shard_count > 1 ? p->cand_shard_index : 0;
    if (k.shard_index < 0 || k.shard_index >= k.shard_count)
        return fail(D2D_ERR_INVALID_ARGUMENT, "need 0 <= cand_shard_index < cand_shard_count");
    // 2-D tiled grids are laid out for thread-block clusters: 8 CTAs = 2 x 4 tiles = one 32 x 32 macro tile whose
    // candidates are culled once, cooperatively, through distributed shared memory (csrc/d2d_driver.cuh)
    k.cluster = (k.grid_cols > 0 && k.slices == 1) ? 8 : 0;
    k.mask = p->active_mask;
    k.mask_wpw = (total + 31) / 32;
    return D2D_OK;
}

}  // namespace

extern "C" {
#ifdef D2D_DEBUG_COUNTERS
int d2d_debug_counters(int32_t mode, uint64_t* out32, int32_t reset);
#endif

void d2d_problem_defaults(D2DProblem* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->grid_role = D2D_GRID_RECEIVERS;
    p->min_order = 0;
    p->max_order = 1;  // scene.py:1817
    p->method = D2D_METHOD_IMAGE;
    p->steps = 100;  // optimize.py:49
    p->lr = 0.1f;    // optimize.py:83
    p->optimizer = D2D_OPT_ADAM;
    p->opt_b1 = 0.9f;  // optax.adam defaults
    p->opt_b2 = 0.999f;
    p->opt_eps = 1e-8f;
    p->mode = D2D_MODE_HARD;
    p->alpha = 100.0f;  // defaults.py:3
    p->tol = 1e-2f;     // geometry.py:915
    p->patch = 0.0f;    // defaults.py:7
    p->fun = D2D_FUN_RECEIVED_POWER;
    p->r_coef = 0.5;  // defaults.py:12
    p->height = 0.1;  // defaults.py:15
    p->grad_mode = D2D_GRAD_CLEAN;
}

int64_t d2d_candidates_count(int32_t n_objects, int32_t order, const int32_t* filter_nodes, int32_t n_filter) {
    if (n_objects < 0 || n_objects > D2D_MAX_OBJECTS || order < 0) return -1;
    uint32_t blocked[D2D_MAX_OBJECTS / 32];
    const int m = build_blocked(n_objects, filter_nodes, n_filter, blocked);
    if (m < 0) return -1;
    return count_for(m, order);
}

int d2d_candidates_host(int32_t n_objects, int32_t order, const int32_t* filter_nodes, int32_t n_filter,
                        int32_t* out) {
    if (n_objects < 0 || n_objects > D2D_MAX_OBJECTS || order < 0 || order > D2D_MAX_ORDER)
        return fail(D2D_ERR_INVALID_ARGUMENT, "bad n_objects / order");
    uint32_t blocked[D2D_MAX_OBJECTS / 32];
    const int m = build_blocked(n_objects, filter_nodes, n_filter, blocked);
    if (m < 0) return fail(D2D_ERR_INVALID_ARGUMENT, "bad filter_nodes");
    if (order == 0 || !out) return D2D_OK;
    if (m == 0 || (order > 1 && m < 2)) return D2D_OK;
    std::vector<int> allowed;
    for (int j = 0; j < n_objects; ++j)
        if (!((blocked[j >> 5] >> (j & 31)) & 1u)) allowed.push_back(j);
    // odometer over positions in `allowed`
    int pos[D2D_MAX_ORDER];
    for (int i = 0; i < order; ++i) pos[i] = i & 1;
    long long row = 0;
    for (;;) {
        for (int i = 0; i < order; ++i) out[row * order + i] = allowed[pos[i]];
        ++row;
        int i = order - 1;
        for (; i >= 0; --i) {
            int v = pos[i] + 1;
            if (i > 0 && v == pos[i - 1]) ++v;
            if (v < m) {
                pos[i] = v;
                for (int j = i + 1; j < order; ++j) pos[j] = (pos[j - 1] == 0) ? 1 : 0;
                break;
            }
        }
        if (i < 0) break;
    }
    return D2D_OK;
}

int d2d_candidates_device(int32_t n_objects, int32_t order, const int32_t* filter_nodes, int32_t n_filter,
                          int32_t* out, void* stream) {
    if (n_objects < 0 || n_objects > D2D_MAX_OBJECTS || order < 0 || order > D2D_MAX_ORDER)
        return fail(D2D_ERR_INVALID_ARGUMENT, "bad n_objects / order");
    BlockedMask mask;
    const int m = build_blocked(n_objects, filter_nodes, n_filter, mask.w);
    if (m < 0) return fail(D2D_ERR_INVALID_ARGUMENT, "bad filter_nodes");
    const long long count = count_for(m, order);
    if (count < 0) return fail(D2D_ERR_UNSUPPORTED, "candidate count overflows int64");
    if (order == 0 || count == 0) return D2D_OK;
    if (!out) return fail(D2D_ERR_INVALID_ARGUMENT, "out is NULL");
    const int block = 256;
    long long nblk = (count + block - 1) / block;
    if (nblk > 148 * 16) nblk = 148 * 16;
    candidates_kernel<<<(unsigned)nblk, block, 0, (cudaStream_t)stream>>>(mask, n_objects, m, order, count, out);
    g_launches += 1;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? D2D_OK : cuda_fail(e, "candidates_kernel");
}

int64_t d2d_problem_num_candidates(const D2DProblem* p) {
    d2d::KParams k;
    if (pack(p, k) != D2D_OK) return -1;
    return k.C_total;
}

int64_t d2d_active_mask_words(const D2DProblem* p) {
    d2d::KParams k;
    if (pack(p, k) != D2D_OK) return -1;
    return (int64_t)k.T * d2d::num_tile_blocks(k) * 4 * k.mask_wpw;
}

int d2d_power_fwd(const D2DProblem* p, float* Z, float* valid_out, void* stream) {
    d2d::KParams k;
    const int rc = pack(p, k);
    if (rc != D2D_OK) return rc;
    if (!Z && p->n_grid > 0 && p->n_fixed > 0) return fail(D2D_ERR_INVALID_ARGUMENT, "Z is NULL");
    if (p->method != D2D_METHOD_IMAGE && !p->x0 && p->max_order > 0)
        return fail(D2D_ERR_INVALID_ARGUMENT, "Fermat/MinPath need x0 (initial guesses per candidate)");
    long long n = 0;
    const int e = d2d::launch_power_fwd(k, p->mode, p->grid_role, p->method, Z, valid_out, (cudaStream_t)stream, &n);
    g_launches += n;
    return e == 0 ? D2D_OK : cuda_fail(e, "power_fwd_kernel");
}

int d2d_power_bwd(const D2DProblem* p, const float* Zbar, float* Z_out, float* grid_bar, float* objects_bar,
                  float* phis_bar, float* fixed_bar, float* alpha_bar, void* stream) {
    d2d::KParams k;
    const int rc = pack(p, k);
    if (rc != D2D_OK) return rc;
    if (p->method != D2D_METHOD_IMAGE && !p->x0 && p->max_order > 0)
        return fail(D2D_ERR_INVALID_ARGUMENT, "Fermat/MinPath need x0 (initial guesses per candidate)");
    d2d::BwdOut out{Z_out, grid_bar, objects_bar, phis_bar, fixed_bar, alpha_bar};
    long long n = 0;
    int e = d2d::launch_power_bwd(k, p->mode, p->grid_role, p->method, Zbar, out, (cudaStream_t)stream, &n);
    if (e == 0 && p->grad_mode == D2D_GRAD_NAN_PARITY) {
        if (k.C_total > (1LL << 22))
            return fail(D2D_ERR_UNSUPPORTED, "D2D_GRAD_NAN_PARITY visits every candidate per grid point: lists up to 2^22");
        e = d2d::launch_nan_poison(k, p->mode, p->grid_role, out, (cudaStream_t)stream, &n);
    }
    g_launches += n;
    return e == 0 ? D2D_OK : cuda_fail(e, "power_bwd_kernel");
}

int d2d_paths(const D2DProblem* p, float min_valid, int32_t emit_all, D2DPathRecord* records, int64_t capacity,
              unsigned long long* count, void* stream) {
    d2d::KParams k;
    const int rc = pack(p, k);
    if (rc != D2D_OK) return rc;
    if (!count) return fail(D2D_ERR_INVALID_ARGUMENT, "count is NULL");
    if (records && capacity < 0) return fail(D2D_ERR_INVALID_ARGUMENT, "negative capacity");
    if (p->method != D2D_METHOD_IMAGE && !p->x0 && p->max_order > 0)
        return fail(D2D_ERR_INVALID_ARGUMENT, "Fermat/MinPath need x0 (initial guesses per candidate)");
    d2d::PathsOut out{records, records ? (long long)capacity : 0, count, min_valid, emit_all ? 1 : 0};
    long long n = 0;
    const int e = d2d::launch_paths(k, p->mode, p->grid_role, p->method, out, (cudaStream_t)stream, &n);
    g_launches += n;
    return e == 0 ? D2D_OK : cuda_fail(e, "paths_kernel");
}

int d2d_paths_vjp(const D2DProblem* p, int64_t n_records, const int32_t* rec_fixed, const int64_t* rec_grid,
                  const int64_t* rec_candidate, const float* valid_bar, const float* xys_bar, float* grid_bar,
                  float* objects_bar, float* phis_bar, float* fixed_bar, float* alpha_bar, void* stream) {
    d2d::KParams k;
    const int rc = pack(p, k);
    if (rc != D2D_OK) return rc;
    if (p->method != D2D_METHOD_IMAGE) return fail(D2D_ERR_UNSUPPORTED, "d2d_paths_vjp covers ImagePath only");
    if (n_records < 0 || (n_records > 0 && (!rec_fixed || !rec_grid || !rec_candidate || !valid_bar || !xys_bar)))
        return fail(D2D_ERR_INVALID_ARGUMENT, "record arrays / cotangents are NULL");
    d2d::BwdOut out{nullptr, grid_bar, objects_bar, phis_bar, fixed_bar, alpha_bar};
    long long nl = 0;
    const int e = d2d::launch_paths_vjp(k, p->mode, p->grid_role, n_records, rec_fixed, (const long long*)rec_grid,
                                        (const long long*)rec_candidate, valid_bar, xys_bar, out, (cudaStream_t)stream, &nl);
    g_launches += nl;
    return e == 0 ? D2D_OK : cuda_fail(e, "paths_vjp_kernel");
}

int d2d_sanitise_scene(const float* xys, const uint8_t* kinds, const float* phis, int32_t n, const float* points,
                       int64_t n_points, int32_t drop_zero_length, int32_t normalise, float* xys_out, uint8_t* kinds_out,
                       float* phis_out, int32_t* kept_index, int32_t* n_kept, uint8_t* flags, double* affine, void* stream) {
    if (n < 0 || n > D2D_MAX_OBJECTS) return fail(D2D_ERR_UNSUPPORTED, "n outside [0, D2D_MAX_OBJECTS]");
    if ((n > 0 && (!xys || !xys_out)) || n_points < 0 || (n_points > 0 && !points))
        return fail(D2D_ERR_INVALID_ARGUMENT, "xys / xys_out / points is NULL");
    if (normalise && !affine) return fail(D2D_ERR_INVALID_ARGUMENT, "normalise needs the affine output (the map back)");
    long long nl = 0;
    const int e = d2d::launch_sanitise(xys, kinds, phis, n, points, n_points, drop_zero_length ? 1 : 0, normalise ? 1 : 0,
                                       xys_out, kinds_out, phis_out, kept_index, n_kept, flags, affine,
                                       (cudaStream_t)stream, &nl);
    g_launches += nl;
    return e == 0 ? D2D_OK : cuda_fail(e, "sanitise_objects_kernel");
}

int d2d_affine_points(const float* points, int64_t n_points, const double* affine, float* points_out, void* stream) {
    if (n_points < 0 || (n_points > 0 && (!points || !points_out || !affine)))
        return fail(D2D_ERR_INVALID_ARGUMENT, "points / points_out / affine is NULL");
    long long nl = 0;
    const int e = d2d::launch_affine_points(points, n_points, affine, points_out, (cudaStream_t)stream, &nl);
    g_launches += nl;
    return e == 0 ? D2D_OK : cuda_fail(e, "affine_points_kernel");
}

// ---- host-buffer entry ---------------------------------------------------------------------------
namespace {
constexpr int kHostStreams = 3;
constexpr int kHostMaxChunks = 16;

// Staging state of ONE calling thread on ONE device: arenas grown on demand, three streams, one event.  Nothing here is
// shared between threads (the entry point is re-entrant without a lock); freed by d2d_host_release() or at thread exit.
struct HostCtx {
    int dev = -1;
    void* din = nullptr;      size_t din_cap = 0;
    void* dout = nullptr;     size_t dout_cap = 0;
    void* partials = nullptr; size_t partials_cap = 0;  // pinned
    cudaStream_t streams[kHostStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t ready = nullptr;

    cudaError_t init(int device) {
        dev = device;
        cudaError_t e;
        for (int i = 0; i < kHostStreams; ++i)
            if ((e = cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
        return cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
    }
    static cudaError_t grow(void** p, size_t* cap, size_t n, bool pinned) {
        if (n < 256) n = 256;
        if (*p && n <= *cap) return cudaSuccess;
        if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); }
        *p = nullptr;
        *cap = 0;
        const cudaError_t e = pinned ? cudaMallocHost(p, n) : cudaMalloc(p, n);
        if (e == cudaSuccess) *cap = n;
        return e;
    }
    void release() {
        if (dev < 0) return;
        int cur = -1;
        const bool sw = cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess;
        for (auto& s : streams) if (s) { cudaStreamDestroy(s); s = nullptr; }
        if (ready) { cudaEventDestroy(ready); ready = nullptr; }
        if (din) cudaFree(din);
        if (dout) cudaFree(dout);
        if (partials) cudaFreeHost(partials);
        din = dout = partials = nullptr;
        din_cap = dout_cap = partials_cap = 0;
        if (sw) cudaSetDevice(cur);
        dev = -1;
    }
    ~HostCtx() { release(); }  // (at process exit the driver may already be gone: the calls then fail harmlessly)
};

struct HostCtxTable {
    std::vector<HostCtx*> per_dev;
    ~HostCtxTable() { for (HostCtx* c : per_dev) delete c; }
};
thread_local HostCtxTable t_host;

HostCtx* host_ctx(int device, cudaError_t* err) {
    if (device < 0 || device >= 1024) { *err = cudaErrorInvalidDevice; return nullptr; }
    if ((size_t)device >= t_host.per_dev.size()) t_host.per_dev.resize((size_t)device + 1, nullptr);
    HostCtx*& c = t_host.per_dev[(size_t)device];
    if (!c) {
        c = new HostCtx();
        if ((*err = c->init(device)) != cudaSuccess) { delete c; c = nullptr; return nullptr; }
    }
    *err = cudaSuccess;
    return c;
}
}  // namespace

void d2d_host_release(void) {
    for (HostCtx*& c : t_host.per_dev) { delete c; c = nullptr; }
}

// Large 2-D grids are traced in (eight) row chunks on three streams: the upload of chunk c + 1 and the download of chunk c - 1
// overlap the kernels of chunk c (the copies are a quarter of a step otherwise: 12.6 MB each way for 1024^2 points).
// Every chunk is an independent d2d_power_fwd / d2d_power_bwd pair on whole macro tiles (rows in multiples of 32);
// the scene-parameter cotangents of the chunks are partial sums, added on the host.
int d2d_power_host(const D2DProblem* hp, const float* Zbar, float* Z, float* grid_bar, float* objects_bar,
                   float* phis_bar, float* fixed_bar, float* alpha_bar, int32_t device) {
    if (!hp) return fail(D2D_ERR_INVALID_ARGUMENT, "problem is NULL");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    HostCtx* cx = host_ctx(device, &e);
    if (!cx) return cuda_fail(e, "d2d_power_host: stream / event creation");
    const size_t N = (size_t)hp->n_objects, T = (size_t)hp->n_fixed, R = (size_t)hp->n_grid;
    const size_t Tout = hp->reduce_all ? 1 : T;
    D2DProblem dp = *hp;
    dp.alpha_dev = nullptr;
    if (hp->alpha_dev) dp.alpha = *hp->alpha_dev;  // host scalar in this entry (validated like `alpha` by pack())
    const long long C = d2d_problem_num_candidates(&dp);
    if (C < 0) return D2D_ERR_INVALID_ARGUMENT;
    const bool want_bwd = grid_bar || objects_bar || phis_bar || fixed_bar || alpha_bar;
    // chunking: rows in multiples of 32 (whole macro tiles), about an eighth of the grid each
    int n_chunks = 1;
    size_t chunk_pts = R;
    const size_t cols = (hp->grid_cols > 0 && R % (size_t)hp->grid_cols == 0) ? (size_t)hp->grid_cols : 0;
    if (cols > 0 && R >= (size_t)1 << 18) {
        const size_t rows = R / cols;
        size_t want = 8;  // chunks (measured at 1024^2: 1 -> 2.16 ms, 4 -> 1.94, 6 -> 1.83, 8 -> 1.82, 12 -> 1.86;
                          // tuning knob: D2D_HOST_CHUNKS=1..16)
        if (const char* ev = std::getenv("D2D_HOST_CHUNKS")) {
            const long v = std::strtol(ev, nullptr, 10);
            if (v >= 1 && v <= kHostMaxChunks) want = (size_t)v;
        }
        size_t rows_c = ((rows + want - 1) / want + 31) / 32 * 32;
        if (rows_c < rows) {
            chunk_pts = rows_c * cols;
            n_chunks = (int)((R + chunk_pts - 1) / chunk_pts);
            if (n_chunks > kHostMaxChunks) { n_chunks = 1; chunk_pts = R; }
        }
    }
    auto al = [](size_t n) { return (n + 255) / 256 * 256; };
    const size_t x0_bytes = hp->x0 ? (size_t)C * (hp->many > 1 ? hp->many : 1) * hp->max_order * 4 : 0;
    // input arena: scene | grid (whole) | Zbar (chunk-local [Tout, Rc] blocks, one per chunk)
    const size_t o_xys = 0, o_kind = o_xys + al(N * 16), o_phi = o_kind + al(N), o_fix = o_phi + al(N * 4),
                 o_x0 = o_fix + al(T * 8), o_grid = o_x0 + al(x0_bytes), o_zbar = o_grid + al(R * 8),
                 in_total = o_zbar + (Zbar ? (size_t)n_chunks * al(Tout * chunk_pts * 4) : 0);
    if ((e = HostCtx::grow(&cx->din, &cx->din_cap, in_total, false)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(inputs)");
    char* din = (char*)cx->din;
    // output arena per chunk: Z | grid_bar | objects_bar | phis_bar | fixed_bar | alpha_bar | mask
    D2DProblem probe = dp;
    probe.n_grid = (int64_t)chunk_pts;
    probe.objects_xys = (const float*)din;  // (non-NULL placeholders: only sizes matter for the mask size)
    probe.fixed_xy = (const float*)din;
    probe.grid_xy = (const float*)din;
    const long long mask_words = want_bwd ? d2d_active_mask_words(&probe) : 0;
    if (mask_words < 0) return D2D_ERR_INVALID_ARGUMENT;
    const size_t q_z = 0, q_gb = q_z + al(Tout * chunk_pts * 4), q_ob = q_gb + al(grid_bar ? Tout * chunk_pts * 8 : 0),
                 q_pb = q_ob + al(N * 16), q_fb = q_pb + al(N * 4), q_ab = q_fb + al(T * 8), q_mask = q_ab + 256,
                 q_total = q_mask + al((size_t)mask_words * 4) + 256;
    if ((e = HostCtx::grow(&cx->dout, &cx->dout_cap, q_total * n_chunks, false)) != cudaSuccess)
        return cuda_fail(e, "cudaMalloc(outputs)");
    char* dout = (char*)cx->dout;
    // pinned staging for the per-chunk parameter cotangents
    const size_t part_floats = 5 * N + 2 * T + 1;
    if (n_chunks > 1 && want_bwd) {
        if ((e = HostCtx::grow(&cx->partials, &cx->partials_cap, (size_t)n_chunks * part_floats * 4, true)) != cudaSuccess)
            return cuda_fail(e, "cudaMallocHost(partials)");
    }
    // every asynchronous call is checked; after the first failure nothing more is enqueued, the streams are drained
    // (the caller's buffers must not be in flight when we return) and the error is reported
    cudaError_t first = cudaSuccess;
    const char* where = "";
    auto ck = [&](cudaError_t err, const char* w) {
        if (err != cudaSuccess && first == cudaSuccess) { first = err; where = w; }
        return first == cudaSuccess;
    };
    cudaStream_t s0 = cx->streams[0];
    auto h2d = [&](cudaStream_t s, size_t off, const void* src, size_t n) {
        if (src && n && first == cudaSuccess)
            ck(cudaMemcpyAsync(din + off, src, n, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync(host to device)");
    };
    // scene tables on stream 0; the other streams wait for them
    h2d(s0, o_xys, hp->objects_xys, N * 16);
    h2d(s0, o_kind, hp->object_kinds, N);
    h2d(s0, o_phi, hp->object_phis, N * 4);
    h2d(s0, o_fix, hp->fixed_xy, T * 8);
    if (hp->x0) h2d(s0, o_x0, hp->x0, x0_bytes);
    if (first == cudaSuccess) ck(cudaEventRecord(cx->ready, s0), "cudaEventRecord");
    dp.objects_xys = (const float*)(din + o_xys);
    dp.object_kinds = hp->object_kinds ? (const uint8_t*)(din + o_kind) : nullptr;
    dp.object_phis = hp->object_phis ? (const float*)(din + o_phi) : nullptr;
    dp.fixed_xy = (const float*)(din + o_fix);
    dp.x0 = hp->x0 ? (const float*)(din + o_x0) : nullptr;
    int rc = D2D_OK;
    for (int c = 0; c < n_chunks && first == cudaSuccess && rc == D2D_OK; ++c) {
        cudaStream_t s = cx->streams[c % kHostStreams];
        if (s != s0 || c >= kHostStreams) { if (!ck(cudaStreamWaitEvent(s, cx->ready, 0), "cudaStreamWaitEvent")) break; }
        const size_t r0 = (size_t)c * chunk_pts, Rc = (r0 + chunk_pts <= R) ? chunk_pts : R - r0;
        char* qo = dout + (size_t)c * q_total;
        const size_t zb_off = o_zbar + (size_t)c * al(Tout * chunk_pts * 4);
        h2d(s, o_grid + r0 * 8, hp->grid_xy ? (const char*)hp->grid_xy + r0 * 8 : nullptr, Rc * 8);
        if (Zbar)
            for (size_t t = 0; t < Tout; ++t) h2d(s, zb_off + t * Rc * 4, Zbar + t * R + r0, Rc * 4);
        if (first != cudaSuccess) break;
        D2DProblem cp = dp;
        cp.n_grid = (int64_t)Rc;
        cp.grid_xy = (const float*)(din + o_grid + r0 * 8);
        if (want_bwd) {
            // value, then the VJP over the paths the forward found alive (the activity mask is the only residual)
            cp.active_mask = (uint32_t*)(qo + q_mask);
            rc = d2d_power_fwd(&cp, (float*)(qo + q_z), nullptr, s);
            if (rc != D2D_OK) break;
            rc = d2d_power_bwd(&cp, Zbar ? (const float*)(din + zb_off) : nullptr, nullptr,
                               grid_bar ? (float*)(qo + q_gb) : nullptr, objects_bar ? (float*)(qo + q_ob) : nullptr,
                               phis_bar ? (float*)(qo + q_pb) : nullptr, fixed_bar ? (float*)(qo + q_fb) : nullptr,
                               alpha_bar ? (float*)(qo + q_ab) : nullptr, s);
        } else {
            cp.active_mask = nullptr;
            rc = d2d_power_fwd(&cp, (float*)(qo + q_z), nullptr, s);
        }
        if (rc != D2D_OK) break;
        auto d2h = [&](void* dst, size_t off, size_t n) {
            if (dst && n && first == cudaSuccess)
                ck(cudaMemcpyAsync(dst, qo + off, n, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(device to host)");
        };
        for (size_t t = 0; t < Tout; ++t) {
            if (Z) d2h(Z + t * R + r0, q_z + t * Rc * 4, Rc * 4);
            if (grid_bar) d2h(grid_bar + 2 * (t * R + r0), q_gb + t * Rc * 8, Rc * 8);
        }
        if (n_chunks == 1) {
            d2h(objects_bar, q_ob, N * 16);
            d2h(phis_bar, q_pb, N * 4);
            d2h(fixed_bar, q_fb, T * 8);
            d2h(alpha_bar, q_ab, 4);
        } else if (want_bwd) {
            float* part = (float*)cx->partials + (size_t)c * part_floats;
            if (objects_bar) d2h(part, q_ob, N * 16);
            if (phis_bar) d2h(part + 4 * N, q_pb, N * 4);
            if (fixed_bar) d2h(part + 5 * N, q_fb, T * 8);
            if (alpha_bar) d2h(part + 5 * N + 2 * T, q_ab, 4);
        }
    }
    for (int i = 0; i < kHostStreams; ++i) ck(cudaStreamSynchronize(cx->streams[i]), "cudaStreamSynchronize");
    if (rc != D2D_OK) return rc;  // (d2d_power_fwd / d2d_power_bwd have set the message)
    if (first != cudaSuccess) return cuda_fail(first, where);
    if (n_chunks > 1 && want_bwd) {  // partial sums of the chunks, in chunk order
        auto add = [&](float* dst, size_t off, size_t n) {
            if (!dst) return;
            for (size_t i = 0; i < n; ++i) {
                float acc = 0.0f;
                for (int c = 0; c < n_chunks; ++c) acc += ((const float*)cx->partials)[(size_t)c * part_floats + off + i];
                dst[i] = acc;
            }
        };
        add(objects_bar, 0, 4 * N);
        add(phis_bar, 4 * N, N);
        add(fixed_bar, 5 * N, 2 * T);
        add(alpha_bar, 5 * N + 2 * T, 1);
    }
    return D2D_OK;
}

int64_t d2d_launch_count(void) { return g_launches.load(); }

int d2d_fma_peak_launch(float* sink, int32_t iters, double* flops, void* stream) {
    if (!sink || iters <= 0) return fail(D2D_ERR_INVALID_ARGUMENT, "bad arguments");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    fma_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters);
    g_launches += 1;
    if (flops) *flops = (double)blocks * threads * (double)iters * 4.0 * 16.0 * 2.0;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? D2D_OK : cuda_fail(e, "fma_peak_kernel");
}

#ifdef D2D_DEBUG_COUNTERS
// diagnostic builds only: census of the forward kernel's culls for one logic mode (reset != 0 clears it)
int d2d_debug_counters(int32_t mode, uint64_t* out32, int32_t reset) {
    switch (mode) {
        case D2D_MODE_HARD: d2d::debug_counters_mode<D2D_MODE_HARD>((unsigned long long*)out32, reset); break;
        case D2D_MODE_HARD_SIGMOID: d2d::debug_counters_mode<D2D_MODE_HARD_SIGMOID>((unsigned long long*)out32, reset); break;
        default: d2d::debug_counters_mode<D2D_MODE_SIGMOID>((unsigned long long*)out32, reset); break;
    }
    return D2D_OK;
}
#endif

const char* d2d_last_error(void) { return g_err.c_str(); }
int32_t d2d_abi_version(void) { return D2D_ABI_VERSION; }

}  // extern "C"
