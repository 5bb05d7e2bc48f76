// differt2d_b200 — XLA-FFI handlers over the C ABI of include/differt2d_b200.h: the `jax.ffi` custom calls that
// differt2d_b200/jax_binding.py wraps in jax.custom_vjp (BASELINE north_star: "thin jax.ffi C-ABI custom call").
//
// NOT compiled in the build image: JAX / jaxlib and the XLA-FFI headers (xla/ffi/api/ffi.h, shipped inside jaxlib at
// jax.ffi.include_dir()) are not installable there.  `python -m differt2d_b200.build --jax-ffi` compiles this file when
// `import jax` works:   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())")
//                            -I<cuda>/include d2d_xla_ffi.cc -L_lib -ldiffert2d_b200 -o _lib/libdiffert2d_b200_xla.so
// The handlers only unpack buffers / attributes into a D2DProblem and forward: no device allocation, no
// synchronisation, the launch goes to the stream XLA hands over.  Static configuration travels as attributes (part of
// the XLA compile-cache key, like the Python scalars eqx.filter_jit treats as static — SURVEY §3.1), arrays and the
// traced alpha as buffers.
#include <cstdint>

#include "xla/ffi/api/ffi.h"

#include <cuda_runtime_api.h>

#include "../../include/differt2d_b200.h"

namespace ffi = xla::ffi;

namespace {

struct Static {  // attributes shared by the three handlers
    int32_t grid_role, grid_cols, min_order, max_order, method, steps, many, mode, fun, grad_mode, candidate_slices;
    float lr, tol, patch;
    double r_coef, height;
    bool reduce_all, cull;
};

void fill(D2DProblem& p, const Static& s, ffi::Buffer<ffi::F32>& xys, ffi::Buffer<ffi::U8>& kinds,
          ffi::Buffer<ffi::F32>& phis, ffi::Buffer<ffi::F32>& fixed, ffi::Buffer<ffi::F32>& grid,
          ffi::Buffer<ffi::F32>& alpha, ffi::Buffer<ffi::F32>& x0, ffi::Span<const int32_t> filter_nodes) {
    d2d_problem_defaults(&p);
    p.n_objects = static_cast<int32_t>(xys.element_count() / 4);
    p.objects_xys = xys.typed_data();
    p.object_kinds = kinds.element_count() ? kinds.typed_data() : nullptr;
    p.object_phis = phis.element_count() ? phis.typed_data() : nullptr;
    p.n_fixed = static_cast<int32_t>(fixed.element_count() / 2);
    p.fixed_xy = fixed.typed_data();
    p.n_grid = static_cast<int64_t>(grid.element_count() / 2);
    p.grid_xy = grid.typed_data();
    p.grid_role = s.grid_role;
    p.grid_cols = s.grid_cols;
    p.min_order = s.min_order;
    p.max_order = s.max_order;
    p.filter_nodes = filter_nodes.size() ? filter_nodes.begin() : nullptr;  // host memory (attribute)
    p.n_filter = static_cast<int32_t>(filter_nodes.size());
    p.method = s.method;
    p.steps = s.steps;
    p.many = s.many;
    p.lr = s.lr;
    p.x0 = x0.element_count() ? x0.typed_data() : nullptr;
    p.mode = s.mode;
    p.alpha_dev = alpha.typed_data();  // traced scalar: stays on the device (and gets a cotangent)
    p.tol = s.tol;
    p.patch = s.patch;
    p.fun = s.fun;
    p.r_coef = s.r_coef;
    p.height = s.height;
    p.reduce_all = s.reduce_all ? 1 : 0;
    p.grad_mode = s.grad_mode;
    p.no_cull = s.cull ? 0 : 1;
    p.candidate_slices = s.candidate_slices;
}

ffi::Error status(int rc) {
    if (rc == D2D_OK) return ffi::Error::Success();
    const auto code = rc == D2D_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument
                    : rc == D2D_ERR_UNSUPPORTED      ? ffi::ErrorCode::kUnimplemented
                                                     : ffi::ErrorCode::kInternal;
    return ffi::Error(code, d2d_last_error());
}

#define D2D_STATIC_PARAMS                                                                                             \
    int32_t grid_role, int32_t grid_cols, int32_t min_order, int32_t max_order, int32_t method, int32_t steps,        \
        int32_t many, int32_t mode, int32_t fun, int32_t grad_mode, int32_t candidate_slices, float lr, float tol,    \
        float patch, double r_coef, double height, bool reduce_all, bool cull
#define D2D_STATIC_VALUE                                                                                              \
    Static { grid_role, grid_cols, min_order, max_order, method, steps, many, mode, fun, grad_mode, candidate_slices, \
             lr, tol, patch, r_coef, height, reduce_all, cull }

// Z = power map [T, R] (or [R]); `mask` (optional result, may be empty) = the custom_vjp residual.
ffi::Error PowerFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> xys, ffi::Buffer<ffi::U8> kinds,
                        ffi::Buffer<ffi::F32> phis, ffi::Buffer<ffi::F32> fixed, ffi::Buffer<ffi::F32> grid,
                        ffi::Buffer<ffi::F32> alpha, ffi::Buffer<ffi::F32> x0, ffi::Span<const int32_t> filter_nodes,
                        D2D_STATIC_PARAMS, ffi::ResultBuffer<ffi::F32> Z, ffi::ResultBuffer<ffi::U32> mask) {
    D2DProblem p;
    fill(p, D2D_STATIC_VALUE, xys, kinds, phis, fixed, grid, alpha, x0, filter_nodes);
    if (mask->element_count() > 0) {
        if (static_cast<int64_t>(mask->element_count()) < d2d_active_mask_words(&p))
            return ffi::Error(ffi::ErrorCode::kInvalidArgument, "activity mask result is too small");
        p.active_mask = mask->typed_data();
    }
    return status(d2d_power_fwd(&p, Z->typed_data(), nullptr, stream));
}

// cotangents of the grid points, object vertices, RIS angles, fixed points and alpha for a given Zbar
ffi::Error PowerBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> xys, ffi::Buffer<ffi::U8> kinds,
                        ffi::Buffer<ffi::F32> phis, ffi::Buffer<ffi::F32> fixed, ffi::Buffer<ffi::F32> grid,
                        ffi::Buffer<ffi::F32> alpha, ffi::Buffer<ffi::F32> x0, ffi::Buffer<ffi::F32> Zbar,
                        ffi::Buffer<ffi::U32> mask, ffi::Span<const int32_t> filter_nodes, D2D_STATIC_PARAMS,
                        ffi::ResultBuffer<ffi::F32> grid_bar, ffi::ResultBuffer<ffi::F32> objects_bar,
                        ffi::ResultBuffer<ffi::F32> phis_bar, ffi::ResultBuffer<ffi::F32> fixed_bar,
                        ffi::ResultBuffer<ffi::F32> alpha_bar) {
    D2DProblem p;
    fill(p, D2D_STATIC_VALUE, xys, kinds, phis, fixed, grid, alpha, x0, filter_nodes);
    p.active_mask = mask.element_count() ? mask.typed_data() : nullptr;  // (read-only in the backward)
    return status(d2d_power_bwd(&p, Zbar.typed_data(), /*Z_out=*/nullptr, grid_bar->typed_data(),
                                objects_bar->typed_data(), phis_bar->typed_data(), fixed_bar->typed_data(),
                                alpha_bar->typed_data(), stream));
}

#define D2D_BIND_STATIC(b)                                                                                            \
    b.Attr<int32_t>("grid_role").Attr<int32_t>("grid_cols").Attr<int32_t>("min_order").Attr<int32_t>("max_order")     \
        .Attr<int32_t>("method").Attr<int32_t>("steps").Attr<int32_t>("many").Attr<int32_t>("mode")                   \
        .Attr<int32_t>("fun").Attr<int32_t>("grad_mode").Attr<int32_t>("candidate_slices").Attr<float>("lr")          \
        .Attr<float>("tol").Attr<float>("patch").Attr<double>("r_coef").Attr<double>("height")                        \
        .Attr<bool>("reduce_all").Attr<bool>("cull")

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    D2dPowerFwd, PowerFwdImpl,
    D2D_BIND_STATIC(ffi::Ffi::Bind()
                        .Ctx<ffi::PlatformStream<cudaStream_t>>()
                        .Arg<ffi::Buffer<ffi::F32>>()   // xys    [N,2,2]
                        .Arg<ffi::Buffer<ffi::U8>>()    // kinds  [N] or [0]
                        .Arg<ffi::Buffer<ffi::F32>>()   // phis   [N] or [0]
                        .Arg<ffi::Buffer<ffi::F32>>()   // fixed  [T,2]
                        .Arg<ffi::Buffer<ffi::F32>>()   // grid   [R,2]
                        .Arg<ffi::Buffer<ffi::F32>>()   // alpha  [1]
                        .Arg<ffi::Buffer<ffi::F32>>()   // x0     [C,many,max_order] or [0]
                        .Attr<ffi::Span<const int32_t>>("filter_nodes"))
        .Ret<ffi::Buffer<ffi::F32>>()    // Z
        .Ret<ffi::Buffer<ffi::U32>>());  // activity mask (d2d_active_mask_words words, or [0])

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    D2dPowerBwd, PowerBwdImpl,
    D2D_BIND_STATIC(ffi::Ffi::Bind()
                        .Ctx<ffi::PlatformStream<cudaStream_t>>()
                        .Arg<ffi::Buffer<ffi::F32>>()   // xys
                        .Arg<ffi::Buffer<ffi::U8>>()    // kinds
                        .Arg<ffi::Buffer<ffi::F32>>()   // phis
                        .Arg<ffi::Buffer<ffi::F32>>()   // fixed
                        .Arg<ffi::Buffer<ffi::F32>>()   // grid
                        .Arg<ffi::Buffer<ffi::F32>>()   // alpha
                        .Arg<ffi::Buffer<ffi::F32>>()   // x0
                        .Arg<ffi::Buffer<ffi::F32>>()   // Zbar   [T,R] or [R]
                        .Arg<ffi::Buffer<ffi::U32>>()   // activity mask of the forward call, or [0]
                        .Attr<ffi::Span<const int32_t>>("filter_nodes"))
        .Ret<ffi::Buffer<ffi::F32>>()    // grid_bar    [T,R,2] or [R,2]
        .Ret<ffi::Buffer<ffi::F32>>()    // objects_bar [N,2,2]
        .Ret<ffi::Buffer<ffi::F32>>()    // phis_bar    [N]
        .Ret<ffi::Buffer<ffi::F32>>()    // fixed_bar   [T,2]
        .Ret<ffi::Buffer<ffi::F32>>());  // alpha_bar   [1]
