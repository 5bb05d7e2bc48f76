// differt2d_b200 — GPU-side scene sanitiser.
//
// Scene.from_geojson (scene.py:628-663) emits `Wall([coords[i-1], coords[i]])` for every i of a polygon's ring, so a
// CLOSED ring (first vertex repeated at the end) yields one zero-length wall per polygon: n = 0, every candidate
// through it is invalid (residual 1), and the reference's reverse mode turns NaN for the whole map (geometry.py:1105).
// Raw lon/lat coordinates (|y| ~ 50) additionally leave ~3 % of a wall length of fp32 lattice noise on every
// parametric coordinate.  This kernel builds, on the device, the object table a well-conditioned trace wants:
//   * flags the zero-length Wall / RIS objects and (drop != 0) compacts the others, order kept, with the index map;
//   * (normalise != 0) shifts / scales all coordinates to the unit square — origin and extent of the bounding box of
//     the objects and the given points, computed and applied in binary64, then rounded once to binary32.
// With drop == 0 and normalise == 0 it is the identity (the PARITY switch: the raw scene stays the reference case).
// One CTA (n <= D2D_MAX_OBJECTS = 1024 objects); points (transmitters / receivers / a whole grid) are transformed by
// a grid-stride loop of the same launch's second kernel.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/differt2d_b200.h"

namespace {

constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads) sanitise_objects_kernel(
    const float* __restrict__ xys, const uint8_t* __restrict__ kinds, const float* __restrict__ phis, const int n,
    const float* __restrict__ points, const long long n_points, const int drop, const int normalise,
    float* __restrict__ xys_out, uint8_t* __restrict__ kinds_out, float* __restrict__ phis_out,
    int32_t* __restrict__ kept_index, int32_t* __restrict__ n_kept, uint8_t* __restrict__ flags, double* __restrict__ affine) {
    __shared__ int s_warp[kThreads / 32];
    __shared__ double s_box[4][kThreads / 32];
    __shared__ double s_aff[3];
    const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    int kind = D2D_KIND_WALL;
    bool zero = false, keep = false;
    if (j < n) {
        w = reinterpret_cast<const float4*>(xys)[j];
        kind = kinds ? (int)kinds[j] : D2D_KIND_WALL;
        zero = kind != D2D_KIND_VERTEX && w.x == w.z && w.y == w.w;  // a Vertex stores its point twice: not a wall
        keep = !(drop && zero);
        if (flags) flags[j] = zero ? 1 : 0;
    }
    // ordered compaction of the kept objects
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    // bounding box of the KEPT objects and of the points handed in (scene.py:1023-1036: transmitters, receivers, objects)
    double xmin = CUDART_INF, ymin = CUDART_INF, xmax = -CUDART_INF, ymax = -CUDART_INF;
    if (keep) {
        xmin = fmin((double)w.x, (double)w.z); xmax = fmax((double)w.x, (double)w.z);
        ymin = fmin((double)w.y, (double)w.w); ymax = fmax((double)w.y, (double)w.w);
    }
    for (long long q = j; q < n_points; q += kThreads) {
        const double px = points[2 * q], py = points[2 * q + 1];
        xmin = fmin(xmin, px); xmax = fmax(xmax, px);
        ymin = fmin(ymin, py); ymax = fmax(ymax, py);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) { s_box[0][warp] = xmin; s_box[1][warp] = ymin; s_box[2][warp] = xmax; s_box[3][warp] = ymax; }
    __syncthreads();
    int before = 0, total = 0;
    for (int q = 0; q < kThreads / 32; ++q) {
        if (q < warp) before += s_warp[q];
        total += s_warp[q];
    }
    if (j == 0) {
        for (int q = 1; q < kThreads / 32; ++q) {
            xmin = fmin(xmin, s_box[0][q]); ymin = fmin(ymin, s_box[1][q]);
            xmax = fmax(xmax, s_box[2][q]); ymax = fmax(ymax, s_box[3][q]);
        }
        double ox = 0.0, oy = 0.0, sc = 1.0;
        if (normalise && xmax >= xmin) {
            ox = xmin; oy = ymin;
            sc = fmax(xmax - xmin, ymax - ymin);
            if (!(sc > 0.0)) sc = 1.0;
        }
        s_aff[0] = ox; s_aff[1] = oy; s_aff[2] = sc;
        if (affine) { affine[0] = ox; affine[1] = oy; affine[2] = sc; }
        if (n_kept) *n_kept = total;
    }
    __syncthreads();
    if (keep) {
        const int slot = before + __popc(ballot & ((1u << lane) - 1u));
        const double ox = s_aff[0], oy = s_aff[1], sc = s_aff[2];
        float4 o = w;
        if (normalise) {
            o = make_float4((float)(((double)w.x - ox) / sc), (float)(((double)w.y - oy) / sc),
                            (float)(((double)w.z - ox) / sc), (float)(((double)w.w - oy) / sc));
        }
        reinterpret_cast<float4*>(xys_out)[slot] = o;
        if (kinds_out) kinds_out[slot] = (uint8_t)kind;
        if (phis_out) phis_out[slot] = phis ? phis[j] : 0.0f;
        if (kept_index) kept_index[slot] = j;
    }
}

__global__ void affine_points_kernel(const float* __restrict__ in, const long long n, const double* __restrict__ affine,
                                     float* __restrict__ out) {
    const double ox = affine[0], oy = affine[1], sc = affine[2];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
        const float2 p = reinterpret_cast<const float2*>(in)[q];
        reinterpret_cast<float2*>(out)[q] = make_float2((float)(((double)p.x - ox) / sc), (float)(((double)p.y - oy) / sc));
    }
}

}  // namespace

namespace d2d {

int launch_sanitise(const float* xys, const uint8_t* kinds, const float* phis, int n, const float* points,
                    long long n_points, int drop, int normalise, float* xys_out, uint8_t* kinds_out, float* phis_out,
                    int32_t* kept_index, int32_t* n_kept, uint8_t* flags, double* affine, cudaStream_t stream,
                    long long* launches) {
    sanitise_objects_kernel<<<1, kThreads, 0, stream>>>(xys, kinds, phis, n, points, n_points, drop, normalise, xys_out,
                                                        kinds_out, phis_out, kept_index, n_kept, flags, affine);
    if (launches) *launches += 1;
    return (int)cudaGetLastError();
}

int launch_affine_points(const float* in, long long n, const double* affine, float* out, cudaStream_t stream,
                         long long* launches) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    affine_points_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, n, affine, out);
    if (launches) *launches += 1;
    return (int)cudaGetLastError();
}

}  // namespace d2d
