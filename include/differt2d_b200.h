/*
 * differt2d_b200 — C ABI of the B200-native receiver-grid path-tracing hot path of DiffeRT2d.
 *
 * This header is the drop-in boundary.  The reference (pure Python on JAX) has no FFI of its
 * own; what a maintainer would bind is a `jax.ffi` custom call per entry point below (see
 * INTEGRATION.md for the XLA-FFI handler and the `jax.custom_vjp` wiring) replacing, on CUDA
 * devices, the bodies of
 *     Scene.accumulate_on_receivers_grid_over_paths      differt2d/scene.py:1803-1953
 *     Scene.accumulate_on_transmitters_grid_over_paths   differt2d/scene.py:1489-1648
 *     Scene.accumulate_over_paths (point-to-point)       differt2d/scene.py:1272-1334
 *     all_path_candidates                                differt2d/scene.py:122-175
 * Plain C types only: pointers, sizes, scalars.  Unless a comment says "host", every pointer is a
 * DEVICE pointer owned by the caller; the library never allocates or frees device memory on these
 * paths, launches asynchronously on the caller's stream, and never synchronises the device.
 * All floating point is IEEE binary32, all indices int32 (scene.py:167).
 *
 * Error convention: every function returns 0 on success, a D2D_ERR_* code otherwise, and never
 * throws; d2d_last_error() returns a thread-local message for the last failure.
 */
#ifndef DIFFERT2D_B200_H_
#define DIFFERT2D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D2D_ABI_VERSION 3

/* limits of this build */
#define D2D_MAX_ORDER 4      /* interactions per path (reference examples use <= 3) */
#define D2D_MAX_OBJECTS 1024 /* scene objects (walls + RIS + vertices) */

/* object kinds — geometry.py: Wall :540-680, RIS :683-721, Vertex :352-431 */
#define D2D_KIND_WALL 0
#define D2D_KIND_RIS 1
#define D2D_KIND_VERTEX 2

/* which end of the link the grid samples */
#define D2D_GRID_RECEIVERS 0    /* scene.py:1803 — fixed points are transmitters */
#define D2D_GRID_TRANSMITTERS 1 /* scene.py:1489 — fixed points are receivers    */

/* path construction — geometry.py: ImagePath :1013-1114, FermatPath :1117-1204, MinPath :1207-1288 */
#define D2D_METHOD_IMAGE 0
#define D2D_METHOD_FERMAT 1
#define D2D_METHOD_MINPATH 2

/* logic.py:218-617 — `approx=False` is HARD; `approx=True` picks the activation `function` */
#define D2D_MODE_HARD 0
#define D2D_MODE_HARD_SIGMOID 1 /* logic.py:238-255, the default activation (logic.py:266) */
#define D2D_MODE_SIGMOID 2      /* logic.py:218-235 */

/* the accumulated path function `fun` */
#define D2D_FUN_RECEIVED_POWER 0 /* utils.py:16-54 : r_coef**k / (height**2 + length**2) */
#define D2D_FUN_LENGTH_SQUARED 1 /* tests/test_scene.py:444,488 : path.length()**2 */

/* gradient semantics of masked branches (DESIGN.md "NaN semantics") */
#define D2D_GRAD_CLEAN 0 /* masked / saturated branches have zero cotangent */
#define D2D_GRAD_NAN_PARITY 1 /* the clean gradient, then NaN wherever jax.grad over the reference's literal graph    */
                              /* yields NaN (geometry.py:1105 single `where`, :227-230 normalize(0), :163-171 alpha *  */
                              /* inf): ImagePath only, every candidate visited without culling (diagnostic mode;       */
                              /* csrc/d2d_nan.cu)                                                                      */

/* the optimiser behind FermatPath / MinPath — optimize.minimize(..., optimizer=...) (optimize.py:44-97) */
#define D2D_OPT_ADAM 0   /* optax.adam(lr, b1 = opt_b1, b2 = opt_b2, eps = opt_eps): the reference's default (optimize.py:83) */
#define D2D_OPT_SGD 1    /* optax.sgd(lr, momentum = opt_b1) (0 = plain gradient descent)                                    */
#define D2D_OPT_NEWTON 2 /* NOT a reference optimiser — a fast mode: damped Newton iterations on the same loss (`steps` of   */
                         /* them, 8-12 suffice); the reverse mode differentiates the CONVERGED point implicitly               */
                         /* (d theta / d q = -H^-1 d grad / d q) instead of unrolling the scan; many > 1 is not supported      */

#define D2D_OK 0
#define D2D_ERR_INVALID_ARGUMENT 1
#define D2D_ERR_UNSUPPORTED 2
#define D2D_ERR_CUDA 3

typedef struct D2DProblem {
    /* ---- scene objects: Scene.objects (scene.py:178-192) ---------------------------------- */
    int32_t n_objects;
    const float *objects_xys;    /* [n_objects,2,2]; a Vertex stores its xy in both rows          */
    const uint8_t *object_kinds; /* [n_objects] D2D_KIND_* or NULL (= all walls)                  */
    const float *object_phis;    /* [n_objects] RIS reflection angle (geometry.py:692) or NULL    */
    /* ---- link end points ------------------------------------------------------------------- */
    int32_t n_fixed;       /* transmitters (receivers grid) or receivers (transmitters grid)      */
    const float *fixed_xy; /* [n_fixed,2], in dictionary order (scene.py:1072-1087)               */
    int64_t n_grid;        /* R = n*m grid points                                                 */
    const float *grid_xy;  /* [R,2] = dstack((X,Y)).reshape(-1,2) (scene.py:1932), row-major n×m  */
    int32_t grid_role;     /* D2D_GRID_*                                                          */
    int32_t grid_cols;     /* optional: m when the grid is a row-major n x m mesh (X.shape[1]); lets the  */
                           /* kernels use compact 16 x 8 tiles.  0 = unknown (128 consecutive points).     */
    /* ---- candidates: all_path_candidates (scene.py:122-175) --------------------------------- */
    int32_t min_order, max_order;
    const int32_t *filter_nodes; /* HOST pointer: object indices never visited (scene.py:158-160) */
    int32_t n_filter;
    /* ---- path construction ------------------------------------------------------------------ */
    int32_t method;  /* D2D_METHOD_*                                                              */
    int32_t steps;   /* Adam iterations, optimize.py:49 (default 100)                             */
    float lr;        /* Adam learning rate, optimize.py:83 (0.1)                                  */
    const float *x0; /* [C,many,max_order] initial guesses per candidate and restart (optimize.py:132, :173-177) */
    /* ---- validity logic: Path.is_valid (geometry.py:908-963) --------------------------------- */
    int32_t mode;           /* D2D_MODE_*                                                         */
    float alpha;            /* activation slope (defaults.py:3); must be > 0                      */
    const float *alpha_dev; /* optional device scalar overriding `alpha` (a traced jnp scalar).   */
                            /* It cannot be validated on the host without a synchronisation: a    */
                            /* value that is not > 0 makes every output of the call NaN (the folds */
                            /* and culls assume a non-decreasing activation)                       */
    float tol;              /* loss tolerance of is_valid (geometry.py:915), default 1e-2         */
    float patch;            /* wall lengthening for the occlusion test (geometry.py:632-635)      */
    /* ---- accumulated function --------------------------------------------------------------- */
    int32_t fun;    /* D2D_FUN_*                                                                  */
    double r_coef;  /* defaults.py:12 — Python floats: r_coef**k and height*height are folded in    */
    double height;  /* defaults.py:15   double precision and then cast to f32 (utils.py:52-54)      */
    int32_t reduce_all; /* sum over the fixed points (scene.py:1939-1952)                          */
    int32_t grad_mode;  /* D2D_GRAD_*                                                              */
    int32_t no_cull;    /* 1 disables the tile-level candidate culling (identical results, slower)        */
    int32_t candidate_slices; /* 0 = automatic (lists of >= 16384 candidates on few tiles).  > 1: that many CTAs share  */
                              /* each tile's candidate list and combine their partial sums with fp32 atomics (point-to- */
                              /* point links with huge lists, e.g. 500 objects at order 3): the summation order of Z    */
                              /* and grid_bar is then NOT the list order and NOT reproducible from run to run (last-bit */
                              /* differences); 1 forces the ordered, bit-reproducible single-CTA walk.                  */
    uint32_t *active_mask; /* optional DEVICE buffer of d2d_active_mask_words(p) words: the custom_vjp RESIDUAL.       */
                           /* d2d_power_fwd fills it with one bit per (fixed point, warp of 32 grid points, candidate):  */
                           /* "some path of this group has a non-zero validity".  d2d_power_bwd, given the buffer a    */
                           /* forward over the SAME inputs filled, re-traces only the set bits instead of the whole     */
                           /* candidate list (identical results; the reference keeps a full tape of every intermediate */
                           /* at [n,m] size instead, scene.py:1920-1952).  NULL: the backward re-traces everything.     */
    int32_t cand_shard_index; /* multi-GPU sharding of the CANDIDATE list (point-to-point links with huge lists, SURVEY    */
    int32_t cand_shard_count; /* §8e): this call walks chunk (of 128 candidates) number q of every order only when      */
                              /* q % (slices * cand_shard_count) falls into shard cand_shard_index (chunks are dealt     */
                              /* round robin: equal work per rank); order 0 belongs to shard 0.  Z and every cotangent  */
                              /* are then PARTIAL sums: the caller adds them over the shards (one all-reduce).           */
                              /* 0 or 1 = the whole list.                                                                */
    int32_t optimizer;  /* D2D_OPT_*                                                                  */
    float opt_b1;       /* Adam b1 (0.9) / SGD momentum                                               */
    float opt_b2;       /* Adam b2 (0.999)                                                            */
    float opt_eps;      /* Adam eps (1e-8)                                                            */
    int32_t many; /* Fermat/MinPath restarts, optimize.py:142-182 (minimize_many_random_uniform): the scan runs    */
                  /* `many` times from x0[c, 0..many-1] and the iterate with the smallest final loss is kept       */
                  /* (first one on ties, jnp.argmin).  0 or 1 = a single run (the path classes' default, :1198).   */
} D2DProblem;

/* Fills a problem with the reference's defaults (defaults.py, geometry.py:915, optimize.py:49,83). */
void d2d_problem_defaults(D2DProblem *p);

/* ---- candidates (replaces differt_core.rt.CompleteGraph/DiGraph.all_paths at scene.py:154-174) -- */
/* Number of candidates of exactly `order` interactions; -1 on invalid arguments.                    */
int64_t d2d_candidates_count(int32_t n_objects, int32_t order, const int32_t *filter_nodes /*host*/,
                             int32_t n_filter);
/* Lexicographic list, row-major [count, order] int32, into HOST memory.                              */
int d2d_candidates_host(int32_t n_objects, int32_t order, const int32_t *filter_nodes /*host*/,
                        int32_t n_filter, int32_t *out /*host*/);
/* Same list written by an integer CUDA kernel (closed-form index -> sequence decode) into DEVICE memory. */
int d2d_candidates_device(int32_t n_objects, int32_t order, const int32_t *filter_nodes /*host*/,
                          int32_t n_filter, int32_t *out /*device*/, void *stream);
/* Total over [min_order, max_order] as used by a problem (= columns of `valid_out`). */
int64_t d2d_problem_num_candidates(const D2DProblem *p);
/* Size, in 32-bit words, of the `active_mask` residual of a problem; -1 on invalid arguments. */
int64_t d2d_active_mask_words(const D2DProblem *p);

/*
 * Forward map.  Z: [n_fixed, R] (or [R] when reduce_all), Z[t,r] = sum over candidates, in list
 * order, of valid * fun (scene.py:1892-1918).  valid_out: optional [n_fixed, R, C] float32
 * validity of every (fixed point, grid point, candidate) — 0/1 in hard mode — for parity runs.
 */
int d2d_power_fwd(const D2DProblem *p, float *Z, float *valid_out, void *stream);

/*
 * Reverse mode by recomputation (what jax.vjp of the forward yields).
 * Zbar: cotangent of Z, same shape as Z, or NULL (= ones: jax.grad of the sum, scene.py:1920-1923).
 * Every output is optional (NULL = not wanted) and is OVERWRITTEN:
 *   Z_out        same shape as Z (fused value_and_grad)
 *   grid_bar     [n_fixed, R, 2] (or [R,2] when reduce_all): d/d(grid point), per fixed point
 *   objects_bar  [n_objects,2,2]   phis_bar [n_objects]   fixed_bar [n_fixed,2]   alpha_bar [1]
 * Scene-parameter cotangents are reduced over the whole grid with fp32 atomics (order not fixed).
 */
int d2d_power_bwd(const D2DProblem *p, const float *Zbar, float *Z_out, float *grid_bar,
                  float *objects_bar, float *phis_bar, float *fixed_bar, float *alpha_bar,
                  void *stream);

/*
 * Path materialisation — Scene.all_paths / all_valid_paths (scene.py:1156-1248) and the escape hatch for an
 * arbitrary `fun` evaluated by the host framework on the path vertices.
 * One record per emitted (fixed point, grid point, candidate), in arrival order (sort by the three indices
 * for list order).  `xys` holds the order + 2 vertices TX, interaction points..., RX (scene.py's Path.xys).
 */
typedef struct D2DPathRecord {
    int32_t fixed;     /* index into fixed_xy                                                        */
    int32_t order;     /* number of interactions k                                                   */
    int64_t grid;      /* index into grid_xy                                                         */
    int64_t candidate; /* column in the candidate list of the problem (orders ascending, lexicographic) */
    float valid;       /* Path.is_valid: 0/1 (hard) or the smooth truth value                        */
    float loss;        /* Path.loss (geometry.py:1077-1084, :1202-1204, :1286-1288)                 */
    float value;       /* the fused `fun` of the problem (received_power / length**2)                */
    float length;      /* Path.length() (geometry.py:811-819)                                        */
    float xys[(D2D_MAX_ORDER + 2) * 2];
} D2DPathRecord;
/*
 * emit_all = 0: records with valid > min_valid only (all_valid_paths: min_valid = 0.5 is logic.is_true's
 * threshold, logic.py:542-556; hard logic: any min_valid in [0,1)).  emit_all = 1: every triple, as all_paths
 * yields them (complete path and loss even when invalid; min_valid ignored).
 * records: DEVICE array of `capacity` records or NULL (count only).  count: DEVICE counter, overwritten with the
 * number of records the problem emits (it may exceed capacity: the surplus is dropped — size with a NULL pass).
 */
int d2d_paths(const D2DProblem *p, float min_valid, int32_t emit_all, D2DPathRecord *records, int64_t capacity,
              unsigned long long *count, void *stream);

/*
 * Reverse mode of d2d_paths — gradients of a GENERIC `fun` (scene.py:1892-1925 differentiates acc = sum valid * fun(...)
 * for any python callable).  The host framework evaluates and differentiates `fun` on the materialised vertices, which
 * gives, per emitted record, valid_bar = d acc / d valid and xys_bar = d acc / d xys ([n, D2D_MAX_ORDER + 2, 2], rows
 * beyond order + 2 ignored); this call pulls them back through path construction and validity logic (ImagePath only:
 * D2D_ERR_UNSUPPORTED otherwise).  rec_fixed / rec_grid / rec_candidate: the `fixed`, `grid`, `candidate` fields of the
 * records (device arrays, any order).  Outputs as d2d_power_bwd, OVERWRITTEN, each optional; grid_bar is always
 * [n_fixed, R, 2] (per fixed point).  Records whose validity is exactly 0 contribute nothing (clean gradients).
 */
int d2d_paths_vjp(const D2DProblem *p, int64_t n_records, const int32_t *rec_fixed, const int64_t *rec_grid,
                  const int64_t *rec_candidate, const float *valid_bar, const float *xys_bar, float *grid_bar,
                  float *objects_bar, float *phis_bar, float *fixed_bar, float *alpha_bar, void *stream);

/*
 * Scene sanitiser on the device.  Scene.from_geojson (scene.py:628-663) emits one ZERO-LENGTH closure wall
 * per closed polygon ring (coords[i-1] wraps to the duplicated vertex): every candidate through it is invalid, and the
 * reference's own reverse mode turns NaN for the whole map (geometry.py:1105); raw lon/lat coordinates add ~3 % of a wall
 * length of fp32 lattice noise to every parametric coordinate.  This call builds the object table a well-conditioned
 * trace wants.  With drop_zero_length == 0 and normalise == 0 it copies the scene unchanged (the PARITY switch: results on
 * a sanitised scene differ from the reference's on the raw one, which stays the parity case).
 *   flags      [n] u8 or NULL: 1 for a zero-length Wall / RIS (a Vertex is never flagged)
 *   xys_out    [n,2,2]; kinds_out [n] / phis_out [n] / kept_index [n] i32 or NULL: the kept objects, compacted, order kept;
 *              kept_index[q] = index of kept object q in the input (to carry cotangents back); *n_kept (device i32)
 *   points     [n_points,2] or NULL: transmitters / receivers taking part in the bounding box (scene.py:1023-1036)
 *   affine     [3] device doubles {origin x, origin y, scale}: x' = (x - origin) / scale, computed AND applied in binary64,
 *              rounded once to binary32 ({0, 0, 1} without normalise); d/d raw = d/d sanitised / scale
 * n <= D2D_MAX_OBJECTS.  d2d_affine_points applies the same map to any point set (transmitters, receivers, a grid).
 */
int d2d_sanitise_scene(const float *xys, const uint8_t *kinds, const float *phis, int32_t n, const float *points,
                       int64_t n_points, int32_t drop_zero_length, int32_t normalise, float *xys_out,
                       uint8_t *kinds_out, float *phis_out, int32_t *kept_index, int32_t *n_kept, uint8_t *flags,
                       double *affine, void *stream);
int d2d_affine_points(const float *points, int64_t n_points, const double *affine /*device*/, float *points_out,
                      void *stream);

/*
 * Host-buffer entry (numpy users of the reference; `bench.py`'s end-to-end leg): every pointer of `p` and every
 * output is a HOST pointer (pinned memory lets the copies overlap); the call stages the inputs, runs the forward
 * kernel and — when any *_bar output is non-NULL — the backward kernel over the activity mask, copies the results
 * back and synchronises before returning.  2-D grids of >= 2^18 points are traced in row chunks (whole macro tiles,
 * eight by default, D2D_HOST_CHUNKS=1..16) on three internal streams, so that uploads and downloads overlap the
 * kernels; Z and grid_bar are bit-identical to the device entries, the scene-parameter cotangents are the sum of the
 * chunks' partial sums.  `device`: CUDA ordinal.
 * Re-entrant and thread-safe: the staging arenas, streams and the event belong to the CALLING THREAD (one set per
 * thread and device, grown on demand and reused by that thread's later calls); no lock is taken and no state is
 * shared between threads.  d2d_host_release() frees the calling thread's arenas (they are also freed at thread exit).
 */
int d2d_power_host(const D2DProblem *p, const float *Zbar, float *Z, float *grid_bar,
                   float *objects_bar, float *phis_bar, float *fixed_bar, float *alpha_bar,
                   int32_t device);

void d2d_host_release(void);

/* Kernel launches issued by this library since load (for the bench's gpu_launches claim). */
int64_t d2d_launch_count(void);

/* FP32 FMA-chain microbenchmark used as the roofline denominator: runs `iters` dependent-free
 * FMA rounds on every SM and writes the executed flop count to *flops (host).  Asynchronous. */
int d2d_fma_peak_launch(float *sink /*device, >= 1 float*/, int32_t iters, double *flops /*host*/,
                        void *stream);

const char *d2d_last_error(void);
int32_t d2d_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DIFFERT2D_B200_H_ */
