/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not linked into, imported by, or shipped with the product.
 *
 * Plain-C, scalar, IEEE-fp32 restatement of the forward receiver-grid hot path of DiffeRT2d
 * (ImagePath only; Fermat/MinPath live in oracle/ref_torch.py).  It follows the reference
 * operation by operation and recomputes everything per path (normals included) exactly as the
 * reference does; nothing is precomputed, nothing is pruned.  Build with
 *     gcc -O2 -std=c11 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC
 * (see oracle/Makefile) so that no product is ever contracted into an FMA: this file and
 * oracle/ref_torch.py define the canonical rounding of the hard-logic validity masks.
 *
 * Citations are to /root/reference/differt2d/<file>:<line>.
 * Used by: tests/ (parity checker), __graft_entry__.smoke(), bench.py cpu_baseline and
 * `--impl reference` (timed CPU port, OpenMP over grid points).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KIND_WALL 0
#define KIND_RIS 1
#define KIND_VERTEX 2
#define MODE_HARD 0
#define MODE_HARD_SIGMOID 1
#define MODE_SIGMOID 2
#define FUN_RECEIVED_POWER 0
#define FUN_LENGTH_SQUARED 1
#define MAX_ORDER 8

static const float EPS32 = 1.1920928955078125e-07f; /* geometry.py:200 jnp.finfo(float32).eps */
static const float TOL_SEG = 0.005f;                /* geometry.py:89 */

typedef struct { float x, y; } v2;

typedef struct {
    int mode;
    float alpha;
} logic_t;

/* ---- logic.py ---------------------------------------------------------------------- */
static float activation(const logic_t *L, float x) { /* logic.py:218-312 */
    float z = L->alpha * x;
    if (L->mode == MODE_SIGMOID) return 1.0f / (1.0f + expf(-z));
    float v = z + 3.0f; /* jax.nn.hard_sigmoid = relu6(z + 3) / 6 */
    v = v > 0.0f ? v : 0.0f;
    v = v < 6.0f ? v : 6.0f;
    return v / 6.0f;
}
/* truthy values are carried as float: hard -> exactly 0.0f / 1.0f */
static float l_true(const logic_t *L) { (void)L; return 1.0f; }
static float l_false(const logic_t *L) { (void)L; return 0.0f; }
static float l_and(const logic_t *L, float x, float y) { /* logic.py:338-358 */
    if (L->mode == MODE_HARD) return (x != 0.0f && y != 0.0f) ? 1.0f : 0.0f;
    return x < y ? x : y;
}
static float l_or(const logic_t *L, float x, float y) { /* logic.py:315-335 */
    if (L->mode == MODE_HARD) return (x != 0.0f || y != 0.0f) ? 1.0f : 0.0f;
    return x > y ? x : y;
}
static float l_not(const logic_t *L, float x) { /* logic.py:361-377 */
    if (L->mode == MODE_HARD) return x != 0.0f ? 0.0f : 1.0f;
    return 1.0f - x;
}
static float l_ge(const logic_t *L, float x, float y) { /* logic.py:407-433 */
    if (L->mode == MODE_HARD) return x >= y ? 1.0f : 0.0f;
    return activation(L, x - y);
}
static float l_le(const logic_t *L, float x, float y) { /* logic.py:463-487 */
    if (L->mode == MODE_HARD) return x <= y ? 1.0f : 0.0f;
    return activation(L, y - x);
}
static float l_lt(const logic_t *L, float x, float y) { /* logic.py:436-460 */
    if (L->mode == MODE_HARD) return x < y ? 1.0f : 0.0f;
    return activation(L, y - x);
}

/* ---- geometry.py helpers --------------------------------------------------------------- */
static float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static v2 sub(v2 a, v2 b) { v2 r = {a.x - b.x, a.y - b.y}; return r; }

static v2 normalize(v2 v) { /* geometry.py:206-230 */
    float len = sqrtf(v.x * v.x + v.y * v.y);
    if (len == 0.0f) len = 1.0f;
    v2 r = {v.x / len, v.y / len};
    return r;
}

typedef struct {
    int n;
    const float *xys; /* [n,2,2] */
    const uint8_t *kinds;
    const float *phis;
} scene_t;

static v2 origin(const scene_t *s, int j) { v2 r = {s->xys[4 * j + 0], s->xys[4 * j + 1]}; return r; }
static v2 dest(const scene_t *s, int j) { v2 r = {s->xys[4 * j + 2], s->xys[4 * j + 3]}; return r; }
static v2 tvec(const scene_t *s, int j) { return sub(dest(s, j), origin(s, j)); }

static v2 normal(const scene_t *s, int j) { /* geometry.py:561-573 */
    v2 t = tvec(s, j);
    v2 n = {t.y, -t.x};
    return normalize(n);
}

static v2 image_of(const scene_t *s, int j, v2 p) { /* geometry.py:652-670 */
    v2 i = sub(p, origin(s, j));
    v2 n = normal(s, j);
    float c = 2.0f * dot(i, n);
    v2 r = {p.x - c * n.x, p.y - c * n.y};
    return r;
}

static float cartesian_to_parametric(const scene_t *s, int j, v2 p) { /* geometry.py:589-598 */
    v2 other = sub(p, origin(s, j));
    v2 t = tvec(s, j);
    float sq = dot(t, t);
    if (sq == 0.0f) sq = 1.0f;
    return dot(t, other) / sq;
}

static float seg_test(const logic_t *L, float num, float den) { /* geometry.py:161-171 */
    float t;
    if (den == 0.0f) t = INFINITY; else t = num / den;
    return l_and(L, l_ge(L, t, -TOL_SEG), l_le(L, t, 1.0f + TOL_SEG));
}

static float segments_intersect(const logic_t *L, v2 P1, v2 P2, v2 P3, v2 P4) { /* geometry.py:82-173 */
    v2 A = sub(P2, P1), B = sub(P3, P4), C = sub(P1, P3);
    float a = B.y * C.x - B.x * C.y;
    float b = A.x * C.y - A.y * C.x;
    float d = A.y * B.x - A.x * B.y;
    return l_and(L, seg_test(L, a, d), seg_test(L, b, d));
}

static float intersects_cartesian(const scene_t *s, const logic_t *L, int j, v2 r0, v2 r1, float patch) {
    /* geometry.py:623-639 ; Vertex :405-414 */
    if (s->kinds[j] == KIND_VERTEX) return l_false(L);
    v2 t = tvec(s, j), o = origin(s, j), d = dest(s, j);
    v2 P1 = {o.x - patch * t.x, o.y - patch * t.y};
    v2 P2 = {d.x + patch * t.x, d.y + patch * t.y};
    return segments_intersect(L, P1, P2, r0, r1);
}

static float evaluate_cartesian(const scene_t *s, int j, v2 a, v2 b, v2 c) {
    /* Wall geometry.py:641-650 ; RIS :698-711 ; Vertex :416-419 */
    if (s->kinds[j] == KIND_VERTEX) return 0.0f;
    v2 n = normal(s, j);
    if (s->kinds[j] == KIND_WALL) {
        v2 i = normalize(sub(b, a));
        v2 r = normalize(sub(c, b));
        float c2 = 2.0f * dot(i, n);
        v2 e = {r.x - (i.x - c2 * n.x), r.y - (i.y - c2 * n.y)};
        return dot(e, e);
    }
    v2 r = normalize(sub(c, b));
    v2 mr = {-r.x, -r.y};
    float sin_a = mr.x * n.y - mr.y * n.x;
    float cos_a = dot(mr, n);
    float sin_p = sinf(s->phis[j]), cos_p = cosf(s->phis[j]);
    float ds = sin_a - sin_p, dc = cos_a - cos_p;
    return ds * ds + dc * dc;
}

/* ImagePath.from_tx_objects_rx — geometry.py:1017-1114.  xys has k+2 entries. */
static float image_path(const scene_t *s, v2 tx, const int *cand, int k, v2 rx, v2 *xys) {
    xys[0] = tx;
    xys[k + 1] = rx;
    if (k == 0) return 0.0f;
    v2 images[MAX_ORDER];
    v2 image = tx;
    for (int i = 0; i < k; ++i) { image = image_of(s, cand[i], image); images[i] = image; }
    v2 point = rx;
    for (int i = k - 1; i >= 0; --i) {
        int j = cand[i];
        v2 p = origin(s, j), n = normal(s, j);
        v2 u = sub(point, images[i]);
        v2 v = sub(p, point);
        float un = dot(u, n), vn = dot(v, n);
        v2 inc = {0.0f, 0.0f};
        if (!(un == 0.0f)) { inc.x = vn * u.x / un; inc.y = vn * u.y / un; }
        point.x = point.x + inc.x;
        point.y = point.y + inc.y;
        xys[i + 1] = point;
    }
    float loss = 0.0f;
    for (int i = 0; i < k; ++i) loss = loss + evaluate_cartesian(s, cand[i], xys[i], xys[i + 1], xys[i + 2]);
    return loss;
}

static float on_objects(const scene_t *s, const logic_t *L, const int *cand, int k, const v2 *xys) {
    /* geometry.py:821-854 */
    float contains = l_true(L);
    for (int i = 0; i < k; ++i) {
        int j = cand[i];
        float c;
        if (s->kinds[j] == KIND_VERTEX) c = l_true(L);
        else {
            float p = cartesian_to_parametric(s, j, xys[i + 1]);
            c = l_and(L, l_ge(L, p, 0.0f), l_le(L, p, 1.0f)); /* geometry.py:600-621 */
        }
        contains = l_and(L, contains, c);
    }
    return contains;
}

static float intersects_with_objects(const scene_t *s, const logic_t *L, const int *cand, int k,
                                     const v2 *xys, float patch) { /* geometry.py:856-906 */
    float intersects = l_false(L);
    for (int i = 0; i < k + 1; ++i) {
        int before = i == 0 ? -1 : cand[i - 1];
        int after = i == k ? -1 : cand[i];
        for (int j = 0; j < s->n; ++j) {
            if (j == before || j == after) continue;
            intersects = l_or(L, intersects, intersects_cartesian(s, L, j, xys[i], xys[i + 1], patch));
        }
    }
    return intersects;
}

static float is_valid(const scene_t *s, const logic_t *L, const int *cand, int k, const v2 *xys, float loss,
                      float tol, float patch) { /* geometry.py:908-963 */
    float a = on_objects(s, L, cand, k, xys);
    float b = l_not(L, intersects_with_objects(s, L, cand, k, xys, patch));
    float c = l_lt(L, loss, tol);
    float v;
    if (L->mode == MODE_HARD) v = (a != 0.0f && b != 0.0f && c != 0.0f) ? 1.0f : 0.0f;
    else {
        if (isnan(a) || isnan(b) || isnan(c)) v = NAN;
        else { v = a < b ? a : b; v = v < c ? v : c; }
        if (isnan(v)) v = 0.0f; /* jnp.nan_to_num */
    }
    return v;
}

static float path_length(const v2 *xys, int npts) { /* geometry.py:176-203 */
    float total = 0.0f;
    for (int i = 0; i + 1 < npts; ++i) {
        float dx = (xys[i + 1].x - xys[i].x) + EPS32;
        float dy = (xys[i + 1].y - xys[i].y) + EPS32;
        float len = sqrtf(dx * dx + dy * dy);
        total = i == 0 ? len : total + len;
    }
    return total;
}

/* ---- candidates: scene.py:122-175 (differt-core CompleteGraph/DiGraph.all_paths restated) ---- */
/* depth-first, ascending neighbours, no self loops; nodes flagged in `blocked` are never visited */
static int64_t enumerate(int n, int k, const uint8_t *blocked, int32_t *out) {
    if (k == 0) return 1;
    int seq[MAX_ORDER];
    int64_t count = 0;
    int depth = 0;
    seq[0] = -1;
    while (depth >= 0) {
        int v = seq[depth] + 1;
        while (v < n && (blocked[v] || (depth > 0 && v == seq[depth - 1]))) ++v;
        if (v >= n) { --depth; continue; }
        seq[depth] = v;
        if (depth == k - 1) {
            if (out) for (int i = 0; i < k; ++i) out[count * k + i] = seq[i];
            ++count;
        } else {
            ++depth;
            seq[depth] = -1;
        }
    }
    return count;
}

int64_t orc_candidates(int n, int order, const int32_t *filter_nodes, int n_filter, int32_t *out) {
    uint8_t *blocked = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int i = 0; i < n_filter; ++i)
        if (filter_nodes[i] >= 0 && filter_nodes[i] < n) blocked[filter_nodes[i]] = 1;
    int64_t c = enumerate(n, order, blocked, out);
    free(blocked);
    return c;
}

/*
 * Scene.accumulate_on_receivers_grid_over_paths (scene.py:1803-1953; grid_role 0) and
 * Scene.accumulate_on_transmitters_grid_over_paths (scene.py:1489-1648; grid_role 1),
 * fun = received_power (utils.py:16-54) or path.length()**2 (tests/test_scene.py:488).
 * Z: [T,R] (or [R] when reduce_all).  valid_out / fun_out: optional [T,R,C].
 * rcoef_pow[k] = (float)(r_coef ** k) and h2 = (float)(height*height) are folded on the host in
 * double precision, as Python does for the reference's default arguments (utils.py:52-54).
 */
int orc_power_map(const float *xys, const uint8_t *kinds, const float *phis, int n_objects,
                  const float *fixed, int n_fixed, const float *grid, int64_t n_grid, int grid_role,
                  int min_order, int max_order, const int32_t *filter_nodes, int n_filter, int mode,
                  float alpha, float tol_loss, float patch, int fun, const float *rcoef_pow, float h2,
                  int reduce_all, float *Z, float *valid_out, float *fun_out, int nthreads) {
    if (max_order > MAX_ORDER || min_order < 0) return 1;
    scene_t s = {n_objects, xys, kinds, phis};
    logic_t L = {mode, alpha};
    /* materialise the candidate list, in list order (orders ascending) */
    int64_t total = 0;
    int64_t offs[MAX_ORDER + 2];
    int32_t *lists[MAX_ORDER + 1];
    for (int k = min_order; k <= max_order; ++k) {
        int64_t c = orc_candidates(n_objects, k, filter_nodes, n_filter, NULL);
        lists[k] = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c * (k > 0 ? k : 1) + 1));
        orc_candidates(n_objects, k, filter_nodes, n_filter, lists[k]);
        offs[k] = total;
        total += c;
    }
    offs[max_order + 1] = total;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int t = 0; t < n_fixed; ++t) {
        v2 fx = {fixed[2 * t], fixed[2 * t + 1]};
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_grid; ++r) {
            v2 g = {grid[2 * r], grid[2 * r + 1]};
            v2 tx = grid_role == 0 ? fx : g;
            v2 rx = grid_role == 0 ? g : fx;
            float acc = 0.0f; /* scene.py:1893 */
            v2 pts[MAX_ORDER + 2];
            for (int k = min_order; k <= max_order; ++k) {
                int64_t cnt = offs[k + 1] - offs[k];
                for (int64_t c = 0; c < cnt; ++c) {
                    const int32_t *cand = lists[k] + c * k;
                    float loss = image_path(&s, tx, cand, k, rx, pts);
                    float valid = is_valid(&s, &L, cand, k, pts, loss, tol_loss, patch);
                    float len = path_length(pts, k + 2);
                    float val = fun == FUN_RECEIVED_POWER ? rcoef_pow[k] / (h2 + len * len) : len * len;
                    acc = acc + valid * val; /* scene.py:1909 */
                    if (valid_out) valid_out[((int64_t)t * n_grid + r) * total + offs[k] + c] = valid;
                    if (fun_out) fun_out[((int64_t)t * n_grid + r) * total + offs[k] + c] = val;
                }
            }
            if (reduce_all) {
                /* scene.py:1939-1952: Z = 0.0 ; Z = Z + p  (per fixed point, in order) */
                Z[r] = t == 0 ? 0.0f + acc : Z[r] + acc;
            } else {
                Z[(int64_t)t * n_grid + r] = acc;
            }
        }
    }
    for (int k = min_order; k <= max_order; ++k) free(lists[k]);
    return 0;
}

/* single path probe for KATs: xys_out [k+2,2], returns loss; valid/len through pointers */
float orc_image_path(const float *xys, const uint8_t *kinds, const float *phis, int n_objects, const float *tx,
                     const float *rx, const int32_t *cand, int k, int mode, float alpha, float tol_loss,
                     float patch, float *xys_out, float *valid_out, float *len_out) {
    scene_t s = {n_objects, xys, kinds, phis};
    logic_t L = {mode, alpha};
    v2 pts[MAX_ORDER + 2];
    v2 t = {tx[0], tx[1]}, r = {rx[0], rx[1]};
    float loss = image_path(&s, t, cand, k, r, pts);
    for (int i = 0; i < k + 2; ++i) { xys_out[2 * i] = pts[i].x; xys_out[2 * i + 1] = pts[i].y; }
    if (valid_out) *valid_out = is_valid(&s, &L, cand, k, pts, loss, tol_loss, patch);
    if (len_out) *len_out = path_length(pts, k + 2);
    return loss;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
