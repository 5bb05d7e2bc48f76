// differt2d_b200 — host-side launcher declarations shared by the translation units.
#pragma once
#include <cuda_runtime.h>

#include "d2d_device.cuh"

namespace d2d {

struct BwdOut {
    float* Z;            // optional
    float* grid_bar;     // optional
    float* objects_bar;  // optional [N,4]
    float* phis_bar;     // optional [N]
    float* fixed_bar;    // optional [T,2]
    float* alpha_bar;    // optional [1]
};

struct PathsOut {
    D2DPathRecord* records;  // optional (nullptr: count only)
    long long capacity;
    unsigned long long* count;
    float min_valid;
    int emit_all;
};

long long num_tile_blocks(const KParams& p);
int launch_paths(const KParams& p, int mode, int grid_role, int method, const PathsOut& out, cudaStream_t stream,
                 long long* launches);
template <int MODE>
int launch_paths_mode(const KParams& p, int grid_role, int method, const PathsOut& out, cudaStream_t s);
template <> int launch_paths_mode<D2D_MODE_HARD>(const KParams&, int, int, const PathsOut&, cudaStream_t);
template <> int launch_paths_mode<D2D_MODE_HARD_SIGMOID>(const KParams&, int, int, const PathsOut&, cudaStream_t);
template <> int launch_paths_mode<D2D_MODE_SIGMOID>(const KParams&, int, int, const PathsOut&, cudaStream_t);

// returns cudaError_t as int; *launches incremented by the number of kernels launched
int launch_power_fwd(const KParams& p, int mode, int grid_role, int method, float* Z, float* valid_out,
                     cudaStream_t stream, long long* launches);
int launch_power_bwd(const KParams& p, int mode, int grid_role, int method, const float* Zbar, const BwdOut& out,
                     cudaStream_t stream, long long* launches);

// reverse mode of the path materialisation (d2d_paths_bwd.cu): cotangents of (valid, xys) per record -> inputs
int launch_paths_vjp(const KParams& p, int mode, int grid_role, long long n, const int32_t* rec_fixed,
                     const long long* rec_grid, const long long* rec_candidate, const float* valid_bar,
                     const float* xys_bar, const BwdOut& out, cudaStream_t stream, long long* launches);

// scene sanitiser (d2d_sanitise.cu)
int launch_sanitise(const float* xys, const uint8_t* kinds, const float* phis, int n, const float* points,
                    long long n_points, int drop, int normalise, float* xys_out, uint8_t* kinds_out, float* phis_out,
                    int32_t* kept_index, int32_t* n_kept, uint8_t* flags, double* affine, cudaStream_t stream,
                    long long* launches);
int launch_affine_points(const float* in, long long n, const double* affine, float* out, cudaStream_t stream,
                         long long* launches);

// D2D_GRAD_NAN_PARITY: overwrites with NaN the cotangents the reference's literal graph poisons (d2d_nan.cu)
int launch_nan_poison(const KParams& p, int mode, int grid_role, const BwdOut& out, cudaStream_t stream,
                      long long* launches);

// One translation unit per logic mode (compiled from the same source with -DD2D_TU_MODE=<mode>, in parallel).
template <int MODE>
int launch_fwd_mode(const KParams& p, int grid_role, int method, float* Z, float* valid_out, cudaStream_t s);
template <int MODE>
int launch_bwd_mode(const KParams& p, int grid_role, int method, const float* Zbar, const BwdOut& out, cudaStream_t s);
template <int MODE>
int launch_bwd_solver_mode(const KParams& p, int grid_role, int method, const float* Zbar, const BwdOut& out, cudaStream_t s);
template <> int launch_bwd_solver_mode<D2D_MODE_HARD>(const KParams&, int, int, const float*, const BwdOut&, cudaStream_t);
template <> int launch_bwd_solver_mode<D2D_MODE_HARD_SIGMOID>(const KParams&, int, int, const float*, const BwdOut&, cudaStream_t);
template <> int launch_bwd_solver_mode<D2D_MODE_SIGMOID>(const KParams&, int, int, const float*, const BwdOut&, cudaStream_t);
template <> int launch_fwd_mode<D2D_MODE_HARD>(const KParams&, int, int, float*, float*, cudaStream_t);
template <> int launch_fwd_mode<D2D_MODE_HARD_SIGMOID>(const KParams&, int, int, float*, float*, cudaStream_t);
template <> int launch_fwd_mode<D2D_MODE_SIGMOID>(const KParams&, int, int, float*, float*, cudaStream_t);
template <> int launch_bwd_mode<D2D_MODE_HARD>(const KParams&, int, int, const float*, const BwdOut&, cudaStream_t);
template <> int launch_bwd_mode<D2D_MODE_HARD_SIGMOID>(const KParams&, int, int, const float*, const BwdOut&, cudaStream_t);
template <> int launch_bwd_mode<D2D_MODE_SIGMOID>(const KParams&, int, int, const float*, const BwdOut&, cudaStream_t);

#ifdef D2D_DEBUG_COUNTERS
template <int MODE> void debug_counters_mode(unsigned long long* out, int reset);
template <> void debug_counters_mode<D2D_MODE_HARD>(unsigned long long*, int);
template <> void debug_counters_mode<D2D_MODE_HARD_SIGMOID>(unsigned long long*, int);
template <> void debug_counters_mode<D2D_MODE_SIGMOID>(unsigned long long*, int);
#endif

}  // namespace d2d
