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

long long num_tile_blocks(const KParams& p);

// returns cudaError_t as int; *launches incremented by the number of kernels launched
int launch_power_fwd(const KParams& p, int mode, int grid_role, int method, float* Z, float* valid_out,
                     cudaStream_t stream, long long* launches);
int launch_power_bwd(const KParams& p, int mode, int grid_role, int method, const float* Zbar, const BwdOut& out,
                     cudaStream_t stream, long long* launches);

}  // namespace d2d
