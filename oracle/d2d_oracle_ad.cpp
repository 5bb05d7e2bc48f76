/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not linked into, imported by, or shipped with the product.
 *
 * Forward-mode automatic differentiation (dual numbers) over a scalar restatement of the ImagePath hot path of
 * DiffeRT2d: value AND reverse-mode products (what jax.vjp of Scene.accumulate_on_*_grid_over_paths(reduce_all=True)
 * delivers) for the cotangent Zbar, in binary32 (the reference's arithmetic, docs/source/contributing/internals.md:7-10)
 * or in binary64 (the third leg of the parity triangulation: the same function of the same fp32 inputs).
 *
 * Why a second gradient oracle next to oracle/ref_torch.py: (1) it is INDEPENDENT of both torch autograd and the CUDA
 * kernels' hand-written adjoints — dual numbers through the literal formulas; (2) it is compiled, so it checks
 * gradients at the benchmark's own scale (bench.py "parity_spotcheck") and serves as the compiled-CPU forward + VJP
 * baseline (bench.py --impl reference).
 *
 * Method.  Every path is first evaluated with plain scalars, literally and completely (nothing pruned), exactly as
 * oracle/d2d_oracle.c does (same operation order; built with -ffp-contract=off).  Only paths with a non-zero validity
 * are re-evaluated with dual numbers Dual<Real, ND> whose tangent directions are: receiver (2), transmitter (2), alpha
 * (1) and, per tracked object, its four vertex coordinates and its RIS angle (5).  Tracked objects = the candidate's own
 * objects plus the objects whose occlusion test attains the running maximum of Path.intersects_with_objects
 * (geometry.py:887-904): a test below the maximum has no cotangent under jnp.maximum's VJP, so it may be treated as a
 * constant.  min / max follow JAX's rules: ties split 1/2 - 1/2 (lax.min / lax.max JVP), the 3-way jnp.min of
 * logic.all splits evenly (logic.py:511-512).  Masked branches are constants ("clean" gradients, DESIGN.md: the
 * reference's own reverse mode yields NaN there): normalize(0) (geometry.py:227-230), un == 0 (:1105), d == 0 (:163-171).
 *
 * Citations are to /root/reference/differt2d/<file>:<line>.
 * Build: oracle/Makefile (g++ -O2 -std=c++17 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int KIND_WALL = 0, KIND_RIS = 1, KIND_VERTEX = 2;
constexpr int MODE_HARD = 0, MODE_HARD_SIGMOID = 1, MODE_SIGMOID = 2;
constexpr int FUN_RECEIVED_POWER = 0;
constexpr int MAX_ORDER = 8;
constexpr int MAX_TRACKED = 11;  // objects with tangent directions per path: candidate objects + tied occluders
constexpr int ND = 5 + 5 * MAX_TRACKED;

// ---- dual numbers --------------------------------------------------------------------------------------------------
template <class R>
struct Dual {
    R v;
    R d[ND];
    int n;  // directions in use (the others are zero and never touched)
};

template <class R> inline R val(const R& x) { return x; }
template <class R> inline R val(const Dual<R>& x) { return x.v; }

template <class R> struct Ops;  // scalar-type traits

template <> struct Ops<float> { using Real = float; static constexpr bool dual = false; };
template <> struct Ops<double> { using Real = double; static constexpr bool dual = false; };
template <class R> struct Ops<Dual<R>> { using Real = R; static constexpr bool dual = true; };

template <class R> inline Dual<R> dconst(R v, int n) { Dual<R> r; r.v = v; r.n = n; for (int i = 0; i < n; ++i) r.d[i] = 0; return r; }

#define D2D_BIN(op, expr_v, expr_d)                                                              \
    template <class R> inline Dual<R> operator op(const Dual<R>& a, const Dual<R>& b) {          \
        Dual<R> r; r.n = a.n; r.v = expr_v; for (int i = 0; i < a.n; ++i) r.d[i] = expr_d; return r; }
D2D_BIN(+, a.v + b.v, a.d[i] + b.d[i])
D2D_BIN(-, a.v - b.v, a.d[i] - b.d[i])
D2D_BIN(*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
#undef D2D_BIN
template <class R> inline Dual<R> operator/(const Dual<R>& a, const Dual<R>& b) {
    Dual<R> r; r.n = a.n; r.v = a.v / b.v;
    for (int i = 0; i < a.n; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
template <class R> inline Dual<R> operator-(const Dual<R>& a) { Dual<R> r; r.n = a.n; r.v = -a.v; for (int i = 0; i < a.n; ++i) r.d[i] = -a.d[i]; return r; }
template <class R> inline Dual<R> operator+(const Dual<R>& a, R b) { Dual<R> r = a; r.v = a.v + b; return r; }
template <class R> inline Dual<R> operator+(R b, const Dual<R>& a) { Dual<R> r = a; r.v = b + a.v; return r; }
template <class R> inline Dual<R> operator-(const Dual<R>& a, R b) { Dual<R> r = a; r.v = a.v - b; return r; }
template <class R> inline Dual<R> operator-(R b, const Dual<R>& a) { Dual<R> r; r.n = a.n; r.v = b - a.v; for (int i = 0; i < a.n; ++i) r.d[i] = -a.d[i]; return r; }
template <class R> inline Dual<R> operator*(const Dual<R>& a, R b) { Dual<R> r; r.n = a.n; r.v = a.v * b; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * b; return r; }
template <class R> inline Dual<R> operator*(R b, const Dual<R>& a) { Dual<R> r; r.n = a.n; r.v = b * a.v; for (int i = 0; i < a.n; ++i) r.d[i] = b * a.d[i]; return r; }
template <class R> inline Dual<R> operator/(const Dual<R>& a, R b) { Dual<R> r; r.n = a.n; r.v = a.v / b; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] / b; return r; }
template <class R> inline Dual<R> operator/(R b, const Dual<R>& a) {
    Dual<R> r; r.n = a.n; r.v = b / a.v;
    for (int i = 0; i < a.n; ++i) r.d[i] = -r.v * a.d[i] / a.v;
    return r;
}

inline float xsqrt(float x) { return std::sqrt(x); }
inline double xsqrt(double x) { return std::sqrt(x); }
template <class R> inline Dual<R> xsqrt(const Dual<R>& a) {
    Dual<R> r; r.n = a.n; r.v = std::sqrt(a.v);
    const R k = R(0.5) / r.v;  // JAX: g * 0.5 / sqrt(x)
    for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * k;
    return r;
}
inline float xexp(float x) { return std::exp(x); }
inline double xexp(double x) { return std::exp(x); }
template <class R> inline Dual<R> xexp(const Dual<R>& a) {
    Dual<R> r; r.n = a.n; r.v = std::exp(a.v);
    for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * r.v;
    return r;
}
// jax.nn.sigmoid = lax.logistic: value 1 / (1 + exp(-z)) (logic.py:235), JVP g * s * (1 - s) (no inf * 0 at saturation)
inline float xlogistic(float z) { return 1.0f / (1.0f + std::exp(-z)); }
inline double xlogistic(double z) { return 1.0 / (1.0 + std::exp(-z)); }
template <class R> inline Dual<R> xlogistic(const Dual<R>& a) {
    Dual<R> r; r.n = a.n; r.v = xlogistic(a.v);
    const R k = r.v * (R(1) - r.v);
    for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * k;
    return r;
}
inline float xsin(float x) { return std::sin(x); }
inline double xsin(double x) { return std::sin(x); }
inline float xcos(float x) { return std::cos(x); }
inline double xcos(double x) { return std::cos(x); }
template <class R> inline Dual<R> xsin(const Dual<R>& a) { Dual<R> r; r.n = a.n; r.v = std::sin(a.v); const R c = std::cos(a.v); for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * c; return r; }
template <class R> inline Dual<R> xcos(const Dual<R>& a) { Dual<R> r; r.n = a.n; r.v = std::cos(a.v); const R s = -std::sin(a.v); for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * s; return r; }

// jnp.minimum / jnp.maximum: NaN-propagating; JVP splits ties 1/2 - 1/2
template <class R> inline R xmin(R a, R b) { return (a != a || b != b) ? std::numeric_limits<R>::quiet_NaN() : (a < b ? a : b); }
template <class R> inline R xmax(R a, R b) { return (a != a || b != b) ? std::numeric_limits<R>::quiet_NaN() : (a > b ? a : b); }
template <class R> inline Dual<R> xmin(const Dual<R>& a, const Dual<R>& b) {
    if (a.v < b.v) return a;
    if (b.v < a.v) return b;
    Dual<R> r; r.n = a.n; r.v = xmin(a.v, b.v);
    for (int i = 0; i < a.n; ++i) r.d[i] = R(0.5) * (a.d[i] + b.d[i]);
    return r;
}
template <class R> inline Dual<R> xmax(const Dual<R>& a, const Dual<R>& b) {
    if (a.v > b.v) return a;
    if (b.v > a.v) return b;
    Dual<R> r; r.n = a.n; r.v = xmax(a.v, b.v);
    for (int i = 0; i < a.n; ++i) r.d[i] = R(0.5) * (a.d[i] + b.d[i]);
    return r;
}

// a constant of the scalar type with the tangent layout of `like`
template <class R> inline R cst(R v, const R&) { return v; }
template <class R> inline Dual<R> cst(R v, const Dual<R>& like) { return dconst<R>(v, like.n); }

template <class S> struct V2 { S x, y; };
template <class S> inline S dot(const V2<S>& a, const V2<S>& b) { return a.x * b.x + a.y * b.y; }
template <class S> inline V2<S> sub(const V2<S>& a, const V2<S>& b) { return {a.x - b.x, a.y - b.y}; }

// ---- scene: object vertices / angles as scalars of type S ------------------------------------------------------------
template <class S>
struct SceneT {
    using R = typename Ops<S>::Real;
    int n;
    const float* xys;
    const uint8_t* kinds;
    const float* phis;
    // tangent layout (dual evaluation only): tracked[q] = object index of slot q; its directions are 5 + 5 q ... + 4
    int n_tracked = 0;
    int tracked[MAX_TRACKED] = {};
    int ndir = 0;
    SceneT(int n_, const float* x, const uint8_t* k, const float* p) : n(n_), xys(x), kinds(k), phis(p) {}

    int slot_of(int j) const {
        for (int q = 0; q < n_tracked; ++q) if (tracked[q] == j) return q;
        return -1;
    }
    S coord(int j, int c) const {  // c: 0 P1.x, 1 P1.y, 2 P2.x, 3 P2.y
        const R v = (R)xys[4 * j + c];
        if constexpr (Ops<S>::dual) {
            S r = dconst<R>(v, ndir);
            const int q = slot_of(j);
            if (q >= 0) r.d[5 + 5 * q + c] = R(1);
            return r;
        } else {
            return v;
        }
    }
    S phi(int j) const {
        const R v = (R)phis[j];
        if constexpr (Ops<S>::dual) {
            S r = dconst<R>(v, ndir);
            const int q = slot_of(j);
            if (q >= 0) r.d[5 + 5 * q + 4] = R(1);
            return r;
        } else {
            return v;
        }
    }
    V2<S> origin(int j) const { return {coord(j, 0), coord(j, 1)}; }
    V2<S> dest(int j) const { return {coord(j, 2), coord(j, 3)}; }
    V2<S> tvec(int j) const { return sub(dest(j), origin(j)); }
};

// geometry.py:206-230 — clean: |v| == 0 -> v / 1 with a constant 1
template <class S> inline V2<S> normalize(const V2<S>& v) {
    using R = typename Ops<S>::Real;
    const S sq = v.x * v.x + v.y * v.y;
    if (val(sq) == R(0)) return v;
    const S len = xsqrt(sq);
    if (val(len) == R(0)) return v;
    return {v.x / len, v.y / len};
}
template <class S> inline V2<S> normal(const SceneT<S>& s, int j) {  // geometry.py:561-573
    const V2<S> t = s.tvec(j);
    return normalize(V2<S>{t.y, -t.x});
}
template <class S> inline V2<S> image_of(const SceneT<S>& s, int j, const V2<S>& p) {  // geometry.py:652-670
    using R = typename Ops<S>::Real;
    const V2<S> i = sub(p, s.origin(j));
    const V2<S> n = normal(s, j);
    const S c = R(2) * dot(i, n);
    return {p.x - c * n.x, p.y - c * n.y};
}
template <class S> inline S cartesian_to_parametric(const SceneT<S>& s, int j, const V2<S>& p) {  // geometry.py:589-598
    using R = typename Ops<S>::Real;
    const V2<S> other = sub(p, s.origin(j));
    const V2<S> t = s.tvec(j);
    const S sq = dot(t, t);
    if (val(sq) == R(0)) return dot(t, other);  // sq -> 1 (a constant)
    return dot(t, other) / sq;
}

template <class S>
struct Logic {
    using R = typename Ops<S>::Real;
    int mode;
    S alpha;
    S act(const S& x) const {  // logic.py:218-255
        const S z = alpha * x;
        if (mode == MODE_SIGMOID) return xlogistic(z);
        S v = z + R(3);  // jax.nn.hard_sigmoid = relu6(z + 3) / 6 ; relu6 = minimum(maximum(x, 0), 6)
        v = xmax(v, cst(R(0), v));
        v = xmin(v, cst(R(6), v));
        return v / R(6);
    }
};

template <class S> inline S evaluate_cartesian(const SceneT<S>& s, int j, const V2<S>& a, const V2<S>& b, const V2<S>& c) {
    using R = typename Ops<S>::Real;  // Wall geometry.py:641-650 ; RIS :698-711 ; Vertex :416-419
    if (s.kinds[j] == KIND_VERTEX) return cst(R(0), a.x);
    const V2<S> n = normal(s, j);
    if (s.kinds[j] == KIND_WALL) {
        const V2<S> i = normalize(sub(b, a));
        const V2<S> r = normalize(sub(c, b));
        const S c2 = R(2) * dot(i, n);
        const V2<S> e = {r.x - (i.x - c2 * n.x), r.y - (i.y - c2 * n.y)};
        return dot(e, e);
    }
    const V2<S> r = normalize(sub(c, b));
    const V2<S> mr = {-r.x, -r.y};
    const S sin_a = mr.x * n.y - mr.y * n.x;
    const S cos_a = dot(mr, n);
    const S ph = s.phi(j);
    const S ds = sin_a - xsin(ph), dc = cos_a - xcos(ph);
    return ds * ds + dc * dc;
}

// ImagePath.from_tx_objects_rx — geometry.py:1017-1114
template <class S>
inline S image_path(const SceneT<S>& s, const V2<S>& tx, const int* cand, int k, const V2<S>& rx, V2<S>* xys) {
    using R = typename Ops<S>::Real;
    xys[0] = tx;
    xys[k + 1] = rx;
    if (k == 0) return cst(R(0), tx.x);
    V2<S> images[MAX_ORDER];
    V2<S> image = tx;
    for (int i = 0; i < k; ++i) { image = image_of(s, cand[i], image); images[i] = image; }
    V2<S> point = rx;
    for (int i = k - 1; i >= 0; --i) {
        const int j = cand[i];
        const V2<S> p = s.origin(j), n = normal(s, j);
        const V2<S> u = sub(point, images[i]);
        const V2<S> v = sub(p, point);
        const S un = dot(u, n), vn = dot(v, n);
        if (!(val(un) == R(0))) {  // geometry.py:1105 (clean: the masked branch is a constant 0)
            point = {point.x + vn * u.x / un, point.y + vn * u.y / un};
        } else {
            point = {point.x + R(0), point.y + R(0)};
        }
        xys[i + 1] = point;
    }
    S loss = cst(R(0), tx.x);
    for (int i = 0; i < k; ++i) loss = loss + evaluate_cartesian(s, cand[i], xys[i], xys[i + 1], xys[i + 2]);
    return loss;
}

constexpr float TOL_SEG = 0.005f;  // geometry.py:89 (fp32 value in both precisions)

// one (segment, object) test of Path.intersects_with_objects, smooth logic — geometry.py:623-639, :82-173
template <class S>
inline S hit_smooth(const SceneT<S>& s, const Logic<S>& L, int j, const V2<S>& r0, const V2<S>& r1, typename Ops<S>::Real patch) {
    using R = typename Ops<S>::Real;
    const V2<S> t = s.tvec(j), o = s.origin(j), d = s.dest(j);
    const V2<S> P1 = {o.x - patch * t.x, o.y - patch * t.y};
    const V2<S> P2 = {d.x + patch * t.x, d.y + patch * t.y};
    const V2<S> A = sub(P2, P1), B = sub(r0, r1), Cc = sub(P1, r0);
    const S a = B.y * Cc.x - B.x * Cc.y;
    const S b = A.x * Cc.y - A.y * Cc.x;
    const S den = A.y * B.x - A.x * B.y;
    if (val(den) == R(0)) return cst(R(0), den);  // t = +inf: ge -> act(+inf) = 1, le -> act(-inf) = 0, and -> 0 (constant)
    const R lo = -(R)TOL_SEG, hi = (R)(1.0f + TOL_SEG);
    const S ta = a / den, tb = b / den;
    const S Ta = xmin(L.act(ta - lo), L.act(hi - ta));
    const S Tb = xmin(L.act(tb - lo), L.act(hi - tb));
    return xmin(Ta, Tb);
}

template <class R> inline bool hit_hard(const SceneT<R>& s, int j, const V2<R>& r0, const V2<R>& r1, R patch) {
    const V2<R> t = s.tvec(j), o = s.origin(j), d = s.dest(j);
    const V2<R> P1 = {o.x - patch * t.x, o.y - patch * t.y};
    const V2<R> P2 = {d.x + patch * t.x, d.y + patch * t.y};
    const V2<R> A = sub(P2, P1), B = sub(r0, r1), Cc = sub(P1, r0);
    const R a = B.y * Cc.x - B.x * Cc.y;
    const R b = A.x * Cc.y - A.y * Cc.x;
    const R den = A.y * B.x - A.x * B.y;
    const R lo = -(R)TOL_SEG, hi = (R)(1.0f + TOL_SEG);
    const R ta = den == R(0) ? std::numeric_limits<R>::infinity() : a / den;
    const R tb = den == R(0) ? std::numeric_limits<R>::infinity() : b / den;
    return (ta >= lo) && (ta <= hi) && (tb >= lo) && (tb <= hi);
}

template <class S> inline S path_length(const V2<S>* xys, int npts) {  // geometry.py:176-203
    using R = typename Ops<S>::Real;
    const R eps = (R)1.1920928955078125e-07f;
    S total = cst(R(0), xys[0].x);
    for (int i = 0; i + 1 < npts; ++i) {
        const S dx = (xys[i + 1].x - xys[i].x) + eps;
        const S dy = (xys[i + 1].y - xys[i].y) + eps;
        const S len = xsqrt(dx * dx + dy * dy);
        total = i == 0 ? len : total + len;
    }
    return total;
}

struct Params {
    int mode, fun;
    float alpha, tol, patch;
    float rc[MAX_ORDER + 1];
    float h2;
};

// valid * fun of one path with scalars of type S.  `only` (dual pass): the occlusion fold visits these objects only
// (the others are strictly below the running maximum: constants).  `ties` (plain pass, smooth logic): receives the
// objects whose test equals the final maximum when that maximum is > 0.
template <class S>
inline S path_contribution(const SceneT<S>& s, const Params& P, const S& alpha, const V2<S>& tx, const V2<S>& rx,
                           const int* cand, int k, const int* only, int n_only, int* ties, int* n_ties,
                           typename Ops<S>::Real* valid_out) {
    using R = typename Ops<S>::Real;
    V2<S> xys[MAX_ORDER + 2];
    const S loss = image_path(s, tx, cand, k, rx, xys);
    Logic<S> L{P.mode, alpha};
    const S zero = cst(R(0), tx.x), one = cst(R(1), tx.x);
    S valid = zero;
    if (P.mode == MODE_HARD) {
        if constexpr (!Ops<S>::dual) {  // the plain pass decides; the dual pass of a valid path takes valid = 1
            bool on = true;
            for (int i = 0; i < k; ++i) {
                const int j = cand[i];
                if (s.kinds[j] == KIND_VERTEX) continue;
                const R p = cartesian_to_parametric(s, j, xys[i + 1]);
                on = on && (p >= R(0)) && (p <= R(1));
            }
            bool inter = false;
            for (int i = 0; i < k + 1; ++i) {
                const int before = i == 0 ? -1 : cand[i - 1], after = i == k ? -1 : cand[i];
                for (int j = 0; j < s.n; ++j) {
                    if (j == before || j == after || s.kinds[j] == KIND_VERTEX) continue;
                    inter = inter || hit_hard<R>(s, j, xys[i], xys[i + 1], (R)P.patch);
                }
            }
            valid = (on && !inter && (loss < (R)P.tol)) ? R(1) : R(0);
        } else {
            valid = one;
        }
    } else {
        // on_objects — geometry.py:821-854, contains_parametric :600-621
        S contains = one;
        for (int i = 0; i < k; ++i) {
            const int j = cand[i];
            S c = one;
            if (s.kinds[j] != KIND_VERTEX) {
                const S p = cartesian_to_parametric(s, j, xys[i + 1]);
                c = xmin(L.act(p - R(0)), L.act(R(1) - p));
            }
            contains = xmin(contains, c);
        }
        // intersects_with_objects — geometry.py:856-906: sequential OR-fold in (segment, object) order
        S inter = zero;
        R hits_max = R(0);
        int n_tied = 0;
        bool tie_overflow = false;
        for (int i = 0; i < k + 1; ++i) {
            const int before = i == 0 ? -1 : cand[i - 1], after = i == k ? -1 : cand[i];
            for (int j = 0; j < s.n; ++j) {
                if (j == before || j == after || s.kinds[j] == KIND_VERTEX) continue;
                if (only) {
                    bool in = false;
                    for (int q = 0; q < n_only; ++q) in = in || only[q] == j;
                    if (!in) continue;
                }
                const S h = hit_smooth(s, L, j, xys[i], xys[i + 1], (R)P.patch);
                if (ties && val(h) > R(0)) {  // objects whose test attains the running maximum (> 0)
                    if (val(h) > hits_max) { hits_max = val(h); n_tied = 0; tie_overflow = false; }
                    if (val(h) == hits_max) {
                        bool have = false;
                        for (int q = 0; q < n_tied; ++q) have = have || ties[q] == j;
                        if (!have) { if (n_tied < MAX_TRACKED) ties[n_tied++] = j; else tie_overflow = true; }
                    }
                }
                inter = xmax(inter, h);
            }
        }
        if (ties) *n_ties = tie_overflow ? -1 : n_tied;
        const S a = contains, b = R(1) - inter, c = L.act((R)P.tol - loss);
        // logic.all: jnp.min over the stacked three (logic.py:511-512): even split among ties; then nan_to_num
        const R av = val(a), bv = val(b), cv = val(c);
        if (av != av || bv != bv || cv != cv) {
            valid = zero;
        } else {
            const R m = av < bv ? (av < cv ? av : cv) : (bv < cv ? bv : cv);
            const int cnt = (av == m) + (bv == m) + (cv == m);
            valid = zero;
            if (av == m) valid = valid + a;
            if (bv == m) valid = valid + b;
            if (cv == m) valid = valid + c;
            valid = valid / (R)cnt;
            if constexpr (!Ops<S>::dual) valid = m;  // (the plain value is the minimum itself, bit for bit)
        }
    }
    if (valid_out) *valid_out = val(valid);
    const S len = path_length(xys, k + 2);
    const S value = P.fun == FUN_RECEIVED_POWER ? (R)P.rc[k] / ((R)P.h2 + len * len) : len * len;  // utils.py:52-54
    return valid * value;  // scene.py:1909
}

long long enumerate(int n, int k, const uint8_t* blocked, int32_t* out) {  // as oracle/d2d_oracle.c (scene.py:122-175)
    if (k == 0) return 1;
    int seq[MAX_ORDER];
    long long count = 0;
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

template <class R>
int run(const float* xys, const uint8_t* kinds, const float* phis, int N, const float* fixed, int T, const float* grid,
        long long Rn, int grid_role, int min_order, int max_order, const int32_t* filter_nodes, int n_filter,
        const Params& P, const float* zbar, double* Z, double* grid_bar, double* objects_bar, double* phis_bar,
        double* fixed_bar, double* alpha_bar, long long* n_valid, int nthreads) {
    if (max_order > MAX_ORDER || min_order < 0) return 1;
    std::vector<uint8_t> blocked((size_t)(N > 0 ? N : 1), 0);
    for (int i = 0; i < n_filter; ++i)
        if (filter_nodes[i] >= 0 && filter_nodes[i] < N) blocked[filter_nodes[i]] = 1;
    std::vector<std::vector<int32_t>> lists(max_order + 1);
    std::vector<long long> counts(max_order + 1, 0);
    for (int k = min_order; k <= max_order; ++k) {
        counts[k] = enumerate(N, k, blocked.data(), nullptr);
        lists[k].resize((size_t)(counts[k] * (k > 0 ? k : 1) + 1));
        enumerate(N, k, blocked.data(), lists[k].data());
    }
    const bool want_grad = grid_bar || objects_bar || phis_bar || fixed_bar || alpha_bar;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    const int nt = omp_get_max_threads();
#else
    const int nt = 1;
#endif
    const size_t np = (size_t)5 * N + 2 * T + 1;  // objects [N,4] | phis [N] | fixed [T,2] | alpha
    std::vector<std::vector<double>> part(nt, std::vector<double>(np, 0.0));
    long long nv_total = 0;
    int err = 0;
    for (long long r = 0; r < Rn; ++r) {
        if (Z) Z[r] = 0.0;
        if (grid_bar) grid_bar[2 * r] = grid_bar[2 * r + 1] = 0.0;
    }
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : nv_total)
    for (long long r = 0; r < Rn; ++r) {
#ifdef _OPENMP
        std::vector<double>& pp = part[omp_get_thread_num()];
#else
        std::vector<double>& pp = part[0];
#endif
        SceneT<R> s(N, xys, kinds, phis);
        const R zb = zbar ? (R)zbar[r] : R(1);
        R zsum = R(0);
        for (int t = 0; t < T; ++t) {
            const V2<R> fx = {(R)fixed[2 * t], (R)fixed[2 * t + 1]};
            const V2<R> g = {(R)grid[2 * r], (R)grid[2 * r + 1]};
            const V2<R> tx = grid_role == 0 ? fx : g, rx = grid_role == 0 ? g : fx;
            R acc = R(0);  // scene.py:1893
            for (int k = min_order; k <= max_order; ++k) {
                for (long long c = 0; c < counts[k]; ++c) {
                    const int32_t* cand = lists[k].data() + c * k;
                    int ties[MAX_TRACKED], n_ties = 0;
                    R valid = R(0);
                    const R contrib = path_contribution<R>(s, P, (R)P.alpha, tx, rx, cand, k, nullptr, 0,
                                                           P.mode == MODE_HARD ? nullptr : ties, &n_ties, &valid);
                    acc = acc + contrib;
                    if (valid == R(0)) continue;
                    nv_total += 1;
                    if (!want_grad || zb == R(0)) continue;
                    // ---- dual pass over this path ----
                    SceneT<Dual<R>> sd(N, xys, kinds, phis);
                    for (int i = 0; i < k; ++i)
                        if (sd.slot_of(cand[i]) < 0) sd.tracked[sd.n_tracked++] = cand[i];
                    int occ[MAX_TRACKED], n_occ = 0;
                    if (n_ties < 0) {  // more tied occluders than slots: reported, the path is skipped
#pragma omp atomic write
                        err = 2;
                        continue;
                    }
                    bool full = false;
                    for (int q = 0; q < n_ties; ++q) {
                        if (sd.slot_of(ties[q]) < 0) {
                            if (sd.n_tracked >= MAX_TRACKED) { full = true; break; }
                            sd.tracked[sd.n_tracked++] = ties[q];
                        }
                        occ[n_occ++] = ties[q];
                    }
                    if (full) {
#pragma omp atomic write
                        err = 2;
                        continue;
                    }
                    sd.ndir = 5 + 5 * sd.n_tracked;
                    const int nd = sd.ndir;
                    auto seed = [&](R v, int dir) { Dual<R> x = dconst<R>(v, nd); x.d[dir] = R(1); return x; };
                    const V2<Dual<R>> gd = {seed(g.x, 0), seed(g.y, 1)}, fd = {seed(fx.x, 2), seed(fx.y, 3)};
                    const Dual<R> ad = seed((R)P.alpha, 4);
                    const V2<Dual<R>> txd = grid_role == 0 ? fd : gd, rxd = grid_role == 0 ? gd : fd;
                    // (an empty `only` list = no occluder carries a cotangent: the fold is a constant)
                    static const int none[1] = {-1};
                    const Dual<R> cd = path_contribution<Dual<R>>(sd, P, ad, txd, rxd, cand, k, n_occ ? occ : none,
                                                                  n_occ ? n_occ : 1, nullptr, nullptr, nullptr);
                    const double w = (double)zb;
                    if (grid_bar) { grid_bar[2 * r] += w * (double)cd.d[0]; grid_bar[2 * r + 1] += w * (double)cd.d[1]; }
                    pp[(size_t)5 * N + 2 * t] += w * (double)cd.d[2];
                    pp[(size_t)5 * N + 2 * t + 1] += w * (double)cd.d[3];
                    pp[(size_t)5 * N + 2 * T] += w * (double)cd.d[4];
                    for (int q = 0; q < sd.n_tracked; ++q) {
                        const int j = sd.tracked[q];
                        for (int cc = 0; cc < 4; ++cc) pp[(size_t)4 * j + cc] += w * (double)cd.d[5 + 5 * q + cc];
                        pp[(size_t)4 * N + j] += w * (double)cd.d[5 + 5 * q + 4];
                    }
                }
            }
            zsum = t == 0 ? R(0) + acc : zsum + acc;  // scene.py:1939-1952
        }
        if (Z) Z[r] = (double)zsum;
    }
    if (n_valid) *n_valid = nv_total;
    for (size_t i = 0; i < np; ++i) {
        double a = 0.0;
        for (int t = 0; t < nt; ++t) a += part[t][i];
        if (i < (size_t)4 * N) { if (objects_bar) objects_bar[i] = a; }
        else if (i < (size_t)5 * N) { if (phis_bar) phis_bar[i - 4 * N] = a; }
        else if (i < (size_t)5 * N + 2 * T) { if (fixed_bar) fixed_bar[i - 5 * N] = a; }
        else if (alpha_bar) alpha_bar[0] = a;
    }
    return err;
}

}  // namespace

extern "C" {

/*
 * Z [R] = sum over the fixed points of the accumulated map (reduce_all, scene.py:1939-1952) and the VJP for the cotangent
 * zbar [R] (NULL: ones): grid_bar [R,2], objects_bar [N,2,2], phis_bar [N], fixed_bar [T,2], alpha_bar [1]; every
 * output is optional and returned in double (binary32 results widened when real64 == 0).  n_valid: paths with a
 * non-zero validity.  Returns 0, 1 (bad orders) or 2 (more tied occluders than MAX_TRACKED).
 */
int orc_ad_power_vjp(int real64, const float* xys, const uint8_t* kinds, const float* phis, int n_objects,
                     const float* fixed, int n_fixed, const float* grid, long long n_grid, int grid_role, int min_order,
                     int max_order, const int32_t* filter_nodes, int n_filter, int mode, float alpha, float tol_loss,
                     float patch, int fun, const float* rcoef_pow, float h2, const float* zbar, double* Z,
                     double* grid_bar, double* objects_bar, double* phis_bar, double* fixed_bar, double* alpha_bar,
                     long long* n_valid, int nthreads) {
    Params P;
    P.mode = mode; P.fun = fun; P.alpha = alpha; P.tol = tol_loss; P.patch = patch; P.h2 = h2;
    for (int k = 0; k <= MAX_ORDER; ++k) P.rc[k] = k <= max_order ? rcoef_pow[k] : 0.0f;
    if (real64)
        return run<double>(xys, kinds, phis, n_objects, fixed, n_fixed, grid, n_grid, grid_role, min_order, max_order,
                           filter_nodes, n_filter, P, zbar, Z, grid_bar, objects_bar, phis_bar, fixed_bar, alpha_bar,
                           n_valid, nthreads);
    return run<float>(xys, kinds, phis, n_objects, fixed, n_fixed, grid, n_grid, grid_role, min_order, max_order,
                      filter_nodes, n_filter, P, zbar, Z, grid_bar, objects_bar, phis_bar, fixed_bar, alpha_bar, n_valid,
                      nthreads);
}

}  // extern "C"
