// klp_oracle.cpp -- CPU ORACLE for the klineprofiles hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a line-by-line *restatement* (not a copy; the reference is Zig) of the
// algorithm of fjebaker/lineprofiles for the kline/kline5/kconv/kconv5/kconv10 path.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it; the product (lineprofiles_b200/) never links, imports or calls it.
//
// PARITY STATUS: the reference toolchain (Zig 0.15) is absent from this environment and the
// reference's own tests assert NO spectrum values (xspec-wrapper.zig:553-585 asserts nothing).
// The only known answers it holds are the index identities of main.zig:19-30 and the
// RangeIterator identities of utils.zig:158-171; those are pinned in tests/.  For flux values:
// **parity unpinned** -- this restatement is the authority, cross-checked by an independently
// written numpy restatement (oracle/numpy_check.py: bit-identical to this file in both modes on
// fine flux, spectra, conv profile and convolved spectrum) and by property tests.
//
// Precision modes (template parameter T):
//   T = float   "faithful": every operation in the type the reference instantiates
//               (LineProfileTable(2, f32), xspec-wrapper.zig:61,347-359), same operation ORDER,
//               no FMA contraction (compile with -ffp-contract=off; Zig's default float mode is
//               strict).  Rebin/normalise/convolution are f64 as in the reference.
//   T = double  "gate": the same code with every f32 replaced by f64 (including the fine
//               g-grid iterator the reference hard-codes to f32, xspec-wrapper.zig:125).
// Transcendentals (cos, pow, log10): Zig ships its own implementations which cannot be run
// here; both the oracle and the CUDA path evaluate them in f64 and round to T.
//
// Cost structure: with faithful_cost=1 the oracle also keeps the reference's *cost* (visit
// all 1999 fine bins per radius, linear scans from index 0, O(n_E^2 * window) convolution) so
// it can stand in as the CPU timing baseline.  faithful_cost=0 skips provably-zero work with
// bit-identical results (asserted in tests/test_oracle.py).
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr int kFineG = 2000;    // FINE_G_BINS, xspec-wrapper.zig:15
constexpr int kRadial = 1500;   // RADIAL_BINS, xspec-wrapper.zig:16
constexpr double kH = 5e-4;     // utils.zig:219
constexpr double kConvRes = 1e-3;  // CONVOLUTION_RESOLUTION, xspec-wrapper.zig:55
constexpr int kConvN = 1999;    // @trunc((2.0-1e-3)/1e-3) in comptime f128 arithmetic = 1999
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kRadPerDeg = 0.017453292519943295769236907684886127;

// Gauss 7 nodes / weights, Integrator.zig:5-16
constexpr double kGaussX[3] = {9.4910791234275852452618968404809e-01,
                               7.415311855993944398638647732811e-01,
                               4.0584515137739716690660641207707e-01};
constexpr double kGaussW[4] = {1.2948496616886969327061143267787e-01,
                               2.797053914892766679014677714229e-01,
                               3.8183005050511894495036977548818e-01,
                               4.1795918367346938775510204081658e-01};

// ---- transcendental policy ---------------------------------------------------------------------
// Default (mode 0): cos / pow / log10 evaluated in f64 and rounded to T -- the policy the CUDA path
// shares.  The reference calls Zig's own std.math.pow(f32, ..), @cos, std.math.log10 on f32
// (emissivity.zig:12,45,52-53,67-75; xspec-wrapper.zig:224-226), which are not correctly rounded
// and cannot be run here.  The other modes exist ONLY for the sensitivity study
// (tools/transcendental_sensitivity.py, tests/test_oracle.py): they bound how far a spectrum can
// move when every transcendental result is off by a few ulps of T.
//   1: +k ulps of T on every result      2: -k ulps      3: pseudo-random in [-k, +k] ulps (hash of the bits)
//   4: the C library's T-precision functions (powf / cosf / log10f for T = float)
//   5: pow by the algorithm of Zig's std/math/pow.zig (a port of Go's math.Pow: frexp, square-and-multiply
//      for the integer part of y, exp(yf * log(x)) for the fraction, all in T; restated from the published
//      algorithm, the Zig sources are not available here), cos / log10 as in mode 4
std::atomic<int> g_tp_mode{0}, g_tp_k{0}, g_tp_mask{7};   // mask: 1 = pow, 2 = cos, 4 = log10 are subject to the policy

template <class T> struct UlpBits;
template <> struct UlpBits<float> { using I = int32_t; };
template <> struct UlpBits<double> { using I = int64_t; };
template <class T> inline T ulp_shift(T v, long k) {
    if (k == 0 || !(v == v) || std::isinf(v) || v == (T)0) return v;
    typename UlpBits<T>::I b;
    std::memcpy(&b, &v, sizeof(T));
    b += (typename UlpBits<T>::I)k;   // sign-magnitude encoding: the magnitude moves by k ulps for either sign
    T r;
    std::memcpy(&r, &b, sizeof(T));
    return r;
}
template <class T> inline T tp_perturb(T v) {
    const int mode = g_tp_mode.load(std::memory_order_relaxed);
    if (mode < 1 || mode > 3) return v;
    const long k = g_tp_k.load(std::memory_order_relaxed);
    long d = k;
    if (mode == 2) d = -k;
    if (mode == 3) {
        typename UlpBits<T>::I b;
        std::memcpy(&b, &v, sizeof(T));
        uint64_t h = (uint64_t)b * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        d = (long)(h % (uint64_t)(2 * k + 1)) - k;
    }
    // move |v| by d ulps (for v < 0 the bit pattern grows with the magnitude as well)
    return ulp_shift<T>(v, d);
}
inline float lib_pow(float x, float y) { return ::powf(x, y); }
inline double lib_pow(double x, double y) { return std::pow(x, y); }
inline float lib_cos(float x) { return ::cosf(x); }
inline double lib_cos(double x) { return std::cos(x); }
inline float lib_log10(float x) { return ::log10f(x); }
inline double lib_log10(double x) { return std::log10(x); }
inline float lib_exp(float x) { return ::expf(x); }
inline double lib_exp(double x) { return std::exp(x); }
inline float lib_log(float x) { return ::logf(x); }
inline double lib_log(double x) { return std::log(x); }
// pow(T, x, y) as Zig's std.math.pow / Go's math.Pow structure it, every operation in T.
template <class T> inline T gostyle_pow(T x, T y) {
    if (y == (T)0 || x == (T)1) return (T)1;
    if (x != x || y != y) return std::numeric_limits<T>::quiet_NaN();
    if (y == (T)1) return x;
    if (x == (T)0 || std::isinf(x) || std::isinf(y)) return lib_pow(x, y);   // special values: any libm agrees
    if (y == (T)0.5) return std::sqrt(x);
    if (y == (T)-0.5) return (T)1 / std::sqrt(x);
    T yi, yf = std::modf(std::fabs(y), &yi);
    if (yf != (T)0 && x < (T)0) return std::numeric_limits<T>::quiet_NaN();
    if (yi >= (T)(sizeof(T) == 4 ? 2147483648.0 : 9223372036854775808.0)) return lib_exp(y * lib_log(x));
    T a1 = (T)1;
    int ae = 0;
    if (yf != (T)0) {
        if (yf > (T)0.5) { yf -= (T)1; yi += (T)1; }
        a1 = lib_exp(yf * lib_log(x));
    }
    int xe = 0;
    T x1 = std::frexp(x, &xe);
    for (long long i = (long long)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }   // over/underflow: left to scalbn below
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < (T)0.5) { x1 += x1; xe -= 1; }
    }
    if (y < (T)0) { a1 = (T)1 / a1; ae = -ae; }
    return std::ldexp(a1, ae);
}
template <class T> inline T t_pow(T x, T y) {
    const int mode = (g_tp_mask.load(std::memory_order_relaxed) & 1) ? g_tp_mode.load(std::memory_order_relaxed) : 0;
    if (mode == 4) return lib_pow(x, y);
    if (mode == 5) return gostyle_pow<T>(x, y);
    return mode ? tp_perturb<T>((T)std::pow((double)x, (double)y)) : (T)std::pow((double)x, (double)y);
}
template <class T> inline T t_cos(T x) {
    const int mode = (g_tp_mask.load(std::memory_order_relaxed) & 2) ? g_tp_mode.load(std::memory_order_relaxed) : 0;
    if (mode >= 4) return lib_cos(x);
    return mode ? tp_perturb<T>((T)std::cos((double)x)) : (T)std::cos((double)x);
}
template <class T> inline T t_log10(T x) {
    const int mode = (g_tp_mask.load(std::memory_order_relaxed) & 4) ? g_tp_mode.load(std::memory_order_relaxed) : 0;
    if (mode >= 4) return lib_log10(x);
    return mode ? tp_perturb<T>((T)std::log10((double)x)) : (T)std::log10((double)x);
}

// Zig's @min/@max (LLVM minnum/maxnum): the non-NaN operand wins.  Spelled out so the CUDA
// path can use the very same expression.
template <class T> inline T t_min(T a, T b) { return (a != a) ? b : ((b < a) ? b : a); }
template <class T> inline T t_max(T a, T b) { return (a != a) ? b : ((a < b) ? b : a); }
// std.math.clamp(val, lower, upper) = @max(lower, @min(val, upper))
template <class T> inline T t_clamp(T v, T lo, T hi) { return t_max<T>(lo, t_min<T>(v, hi)); }

// utils.zig:121-156  RangeIterator: cumulative-add linspace (Q6)
template <class T> struct Range {
    T delta, current;
    size_t remaining;
    Range(T lo, T hi, size_t n) : delta((hi - lo) / (T)(double)(n - 1)), current(lo), remaining(n) {}
    T next() {
        T v = current;
        current += delta;
        --remaining;
        return v;
    }
};

struct Stats {
    std::atomic<uint64_t> bin_visits{0}, active_bins{0}, gauss_bins{0}, edge_evals{0},
        integrand_evals{0}, conv_pairs{0}, conv_terms{0};
};

struct LocalStats {
    uint64_t bin_visits = 0, active_bins = 0, gauss_bins = 0, edge_evals = 0, integrand_evals = 0,
             conv_pairs = 0, conv_terms = 0;
};

// The table as the reference holds it after io.zig:53-87: per frame radii/gmin/gmax[n_r] and
// upper/lower[n_r][n_g]; frames a-major, mu-minor (line-profile.zig:60-70).
struct Table {
    int n_a, n_mu, n_r, n_g;
    std::vector<float> spins, mus;
    std::vector<float> radii, gmin, gmax;  // [F][n_r]
    std::vector<float> upper, lower;       // [F][n_r][n_g]
};

template <class T> struct Frame {
    std::vector<T> radii, gmin, gmax, upper, lower;
    void resize(int n_r, int n_g) {
        radii.resize(n_r); gmin.resize(n_r); gmax.resize(n_r);
        upper.resize((size_t)n_r * n_g); lower.resize((size_t)n_r * n_g);
    }
};

// ---- line-profile.zig ------------------------------------------------------------------------

// utils.zig:92-105  first index with haystack[i] >= item, from `start`
template <class T> inline long find_first_geq(const T* h, size_t n, T item, size_t start) {
    for (size_t i = start; i < n; ++i) if (h[i] >= item) return (long)i;
    return -1;
}
// utils.zig:106-119
template <class T> inline long find_first_leq(const T* h, size_t n, T item, size_t start) {
    for (size_t i = start; i < n; ++i) if (h[i] <= item) return (long)i;
    return -1;
}

// line-profile.zig:60-70 with NParams = 2
inline size_t table_index(const Table& t, size_t ia, size_t imu) { return ia * (size_t)t.n_mu + imu; }

// line-profile.zig:72-93.  last_lower_parameters is never written (Q18), so start is always 0.
template <class T> inline void find_parameter_indices(const Table& t, const T p[2], size_t out[2]) {
    std::vector<T> sp(t.spins.begin(), t.spins.end()), mu(t.mus.begin(), t.mus.end());
    long i0 = find_first_geq<T>(sp.data(), sp.size(), p[0], 0);
    long i1 = find_first_geq<T>(mu.data(), mu.size(), p[1], 0);
    out[0] = i0 < 0 ? 0 : (size_t)i0;   // Q5: above the table -> index 0
    out[1] = i1 < 0 ? 0 : (size_t)i1;
}

// line-profile.zig:209-231 (saturating decrements)
inline void find_tables(const Table& t, const size_t pi[2], size_t out[4]) {
    size_t p0 = pi[0], p1 = pi[1];
    out[1] = table_index(t, p0, p1);
    if (p0 > 0) p0 -= 1;
    out[0] = table_index(t, p0, p1);
    if (p1 > 0) p1 -= 1;
    out[2] = table_index(t, p0, p1);
    if (p0 < pi[0]) p0 += 1;
    out[3] = table_index(t, p0, p1);
}

// utils.zig:43-51
template <class T> inline T lerp_w(T w, T y0, T y1) { return (y1 - y0) * w + y0; }

// line-profile.zig:116-194 + transfer-functions.zig:52-109: bilinear blend (Q14, Q15)
template <class T> void interpolate_parameters(const Table& t, const T p[2], Frame<T>& out, Frame<T>& tmp) {
    size_t pi[2], ti[4];
    find_parameter_indices<T>(t, p, pi);
    find_tables(t, pi, ti);
    bool has_w[2];
    T w[2] = {0, 0};
    for (int k = 0; k < 2; ++k) {  // line-profile.zig:196-207
        has_w[k] = pi[k] != 0;
        if (has_w[k]) {
            const std::vector<float>& arr = k == 0 ? t.spins : t.mus;
            T q0 = (T)arr[pi[k] - 1], q1 = (T)arr[pi[k]];
            w[k] = (p[k] - q0) / (q1 - q0);
        }
    }
    const int n_r = t.n_r, n_g = t.n_g;
    out.resize(n_r, n_g);
    tmp.resize(n_r, n_g);
    auto blend_pair = [&](size_t fa, size_t fb, Frame<T>& dst) {
        const size_t o1 = fa * n_r, o2 = fb * n_r, O1 = fa * (size_t)n_r * n_g, O2 = fb * (size_t)n_r * n_g;
        if (has_w[0]) {
            for (size_t e = 0; e < (size_t)n_r * n_g; ++e) {
                dst.upper[e] = lerp_w<T>(w[0], (T)t.upper[O1 + e], (T)t.upper[O2 + e]);
                dst.lower[e] = lerp_w<T>(w[0], (T)t.lower[O1 + e], (T)t.lower[O2 + e]);
            }
            for (int e = 0; e < n_r; ++e) {
                dst.gmin[e] = lerp_w<T>(w[0], (T)t.gmin[o1 + e], (T)t.gmin[o2 + e]);
                dst.gmax[e] = lerp_w<T>(w[0], (T)t.gmax[o1 + e], (T)t.gmax[o2 + e]);
                dst.radii[e] = lerp_w<T>(w[0], (T)t.radii[o1 + e], (T)t.radii[o2 + e]);
            }
        } else {  // assignFrom(from)
            for (size_t e = 0; e < (size_t)n_r * n_g; ++e) {
                dst.upper[e] = (T)t.upper[O1 + e];
                dst.lower[e] = (T)t.lower[O1 + e];
            }
            for (int e = 0; e < n_r; ++e) {
                dst.gmin[e] = (T)t.gmin[o1 + e];
                dst.gmax[e] = (T)t.gmax[o1 + e];
                dst.radii[e] = (T)t.radii[o1 + e];
            }
        }
    };
    blend_pair(ti[0], ti[1], out);  // cache[0]
    blend_pair(ti[2], ti[3], tmp);  // cache[1]
    // interpolate(&cache[1], &cache[0], weights[1], 0)
    if (has_w[1]) {
        for (size_t e = 0; e < (size_t)n_r * n_g; ++e) {
            out.upper[e] = lerp_w<T>(w[1], tmp.upper[e], out.upper[e]);
            out.lower[e] = lerp_w<T>(w[1], tmp.lower[e], out.lower[e]);
        }
        for (int e = 0; e < n_r; ++e) {
            out.gmin[e] = lerp_w<T>(w[1], tmp.gmin[e], out.gmin[e]);
            out.gmax[e] = lerp_w<T>(w[1], tmp.gmax[e], out.gmax[e]);
            out.radii[e] = lerp_w<T>(w[1], tmp.radii[e], out.radii[e]);
        }
    } else {
        out = tmp;
    }
}

// ---- emissivity.zig ----------------------------------------------------------------------------

template <class T> struct Emissivity {
    int n = 0;          // 0 => PowerLawEmissivity(index)
    T index = 0;        // power-law exponent handed to pow (== -alpha)
    T powers[10] = {}, radii[10] = {}, coeffs[10] = {}, alpha = 0;
    // emissivity.zig:42-62 (+ :26-36)
    void init_knots(int N, const T* w, T rmin, T rmax, T alpha_) {
        n = N;
        alpha = alpha_;
        Range<T> itt(t_log10<T>(rmin), t_log10<T>(rmax), (size_t)N + 1);
        (void)itt.next();
        for (int i = 0; i < N; ++i) { radii[i] = itt.next(); powers[i] = w[i]; }
        coeffs[N - 1] = radii[N - 1] * (powers[N - 1] - alpha);
        for (int j = 0; j < N - 1; ++j) {
            int i = N - j - 2;
            T dp = powers[i] - powers[i + 1];
            coeffs[i] = coeffs[i + 1] + (radii[i] * dp);
        }
        for (int i = 0; i < N; ++i) radii[i] = t_pow<T>((T)10, radii[i]);
        for (int i = 0; i < N; ++i) coeffs[i] = t_pow<T>((T)10, coeffs[i]);
    }
    // emissivity.zig:11-13 / :64-76
    T operator()(T r) const {
        if (n == 0) return t_pow<T>(r, index);
        long i = find_first_geq<T>(radii, (size_t)n, r, 0);
        if (i < 0) return t_pow<T>(r, -alpha);
        return coeffs[i] * t_pow<T>(r, -powers[i]);
    }
};

// ---- transfer-functions.zig: InterpolatingTransferFunction ------------------------------------

template <class T> struct Integrator {
    const Frame<T>* tf;
    int n_r, n_g;
    std::vector<T> cache_upper, cache_lower;
    T cache_gmin = 0, cache_gmax = 0;
    const T* gstars;
    bool faithful_cost;
    LocalStats* st;

    // :135-138
    void stage_index(size_t idx) {
        for (int i = 0; i < n_g; ++i) cache_lower[i] = tf->lower[idx * n_g + i];
        for (int i = 0; i < n_g; ++i) cache_upper[i] = tf->upper[idx * n_g + i];
    }
    // :140-176 (search hint is result-neutral, Q18)
    void stage_radius(T r) {
        long loc = find_first_leq<T>(tf->radii.data(), (size_t)n_r, r, 0);
        if (loc < 0) { stage_index((size_t)n_r - 1); return; }
        if (!(loc > 1)) { stage_index(0); return; }  // Q4: gmin/gmax stay stale
        size_t idx = (size_t)loc, ilow = idx - 1;
        T lo = tf->radii[n_r - 1], hi = tf->radii[0];
        T r_clamped = t_clamp<T>(r, lo, hi);
        T r0 = tf->radii[ilow];
        T factor = (r_clamped - r0) / (tf->radii[idx] - r0);
        // :178-196
        for (int i = 0; i < n_g; ++i) {
            T f0 = tf->lower[ilow * n_g + i];
            cache_lower[i] = (tf->lower[idx * n_g + i] - f0) * factor + f0;
        }
        for (int i = 0; i < n_g; ++i) {
            T f0 = tf->upper[ilow * n_g + i];
            cache_upper[i] = (tf->upper[idx * n_g + i] - f0) * factor + f0;
        }
        T gmax0 = tf->gmax[ilow], gmax1 = tf->gmax[idx];
        cache_gmax = (gmax1 - gmax0) * factor + gmax0;
        T gmin0 = tf->gmin[ilow], gmin1 = tf->gmin[idx];
        cache_gmin = (gmin1 - gmin0) * factor + gmin0;
    }
    // :198-236
    inline void branches_at_gstar(T gstar, T& lower, T& upper) const {
        long idx = find_first_geq<T>(gstars, (size_t)n_g, gstar, 0);
        if (idx < 0) idx = n_g - 1;
        if (!(idx > 1)) { lower = cache_lower[0]; upper = cache_upper[0]; return; }
        T g0 = gstars[idx - 1];
        T factor = (gstar - g0) / (gstars[idx] - g0);
        T f0l = cache_lower[idx - 1];
        lower = factor * (cache_lower[idx] - f0l) + f0l;
        T f0u = cache_upper[idx - 1];
        upper = factor * (cache_upper[idx] - f0u) + f0u;
    }
    // :260-272
    inline T integrand(T g) const {
        st->integrand_evals++;
        T gstar = (g - cache_gmin) / (cache_gmax - cache_gmin);
        T lo, up;
        branches_at_gstar(gstar, lo, up);
        T f = up + lo;
        T denom = std::sqrt(gstar * ((T)1 - gstar)) * (cache_gmax - cache_gmin);
        return f * g * g * g / denom;
    }
    // :329-337 (Q17)
    inline T integrate_edge(T lim, T lim_star) const {
        st->edge_evals++;
        T k = (cache_gmax - cache_gmin) * lim_star + cache_gmin;
        T w = std::fabs(std::sqrt(k) - std::sqrt(lim));
        return (T)2 * w * integrand(k) * (T)0.022360679774997896964091736687313 * (cache_gmax - cache_gmin);
    }
    // Integrator.zig:38-63 (Q1: centre weight unscaled)
    inline T integrate_gauss(T a, T b) const {
        T scale = (b - a) / (T)2;
        T sum = 0;
        for (int k = 0; k < 3; ++k) {
            T xp = (T)(kGaussX[k] + 1.0) * scale + a;
            T xm = (T)(-kGaussX[k] + 1.0) * scale + a;
            T w = (T)kGaussW[k] * scale;
            sum += (integrand(xp) + integrand(xm)) * w;
        }
        T x0 = scale + a;
        sum += integrand(x0) * (T)kGaussW[3];
        return sum;
    }
    // :339-384
    inline T integrate_g_bin(T low, T high) const {
        T flux = 0;
        T glow = t_clamp<T>(low, cache_gmin, cache_gmax);
        T ghigh = t_clamp<T>(high, cache_gmin, cache_gmax);
        if (glow == ghigh) return flux;
        st->active_bins++;
        const T Ht = (T)kH, H1 = (T)(1.0 - kH);
        T gstar_low = (glow - cache_gmin) / (cache_gmax - cache_gmin);
        T gstar_high = (ghigh - cache_gmin) / (cache_gmax - cache_gmin);
        if (gstar_low < Ht) {
            if (gstar_high > Ht) {
                flux += integrate_edge(glow, Ht);
                glow = (cache_gmax - cache_gmin) * Ht + cache_gmin;
            } else {
                flux += integrate_edge(glow, gstar_high);
                return flux;
            }
        }
        if (gstar_high > H1) {
            if (gstar_low < H1) {
                flux += integrate_edge(ghigh, H1);
                ghigh = (cache_gmax - cache_gmin) * H1 + cache_gmin;
            } else {
                flux += integrate_edge(ghigh, gstar_low);
                return flux;
            }
        }
        st->gauss_bins++;
        flux += integrate_gauss(glow, ghigh);
        return flux;
    }
    // :305-327 (Q12)
    void integrate_transfer_function(const T* g_grid, T* out, int n_out, T weight) const {
        int lo = 0, hi = n_out;
        if (!faithful_cost && cache_gmin <= cache_gmax) {
            // bins entirely outside [gmin, gmax] clamp to an empty interval and return 0;
            // out[i] += 0 leaves out[i] unchanged (weight is finite here or the add is of
            // NaN->0).  Restrict the visit to bins that can be non-empty.
            while (lo < n_out && g_grid[lo + 1] <= cache_gmin) ++lo;
            while (hi > lo && g_grid[hi - 1] >= cache_gmax) --hi;
        }
        for (int i = lo; i < hi; ++i) {
            st->bin_visits++;
            T val = integrate_g_bin(g_grid[i], g_grid[i + 1]) * weight;
            if (std::isnan(val) || std::isinf(val)) val = 0;
            out[i] += val;
        }
    }
    // :274-303, Integrator.zig:92-102 (Q11, Q13)
    void integrate(const T* r_grid, int n_radii, const T* g_grid, T* flux, int n_flux, const Emissivity<T>& emis) {
        for (int i = 0; i < n_radii; ++i) {
            T r = r_grid[i];
            if (r < tf->radii[n_r - 1]) continue;
            if (r > tf->radii[0]) continue;
            stage_radius(r);
            T e = emis(r);
            T dr = i == 0 ? r_grid[1] - r_grid[0] : r_grid[i] - r_grid[i - 1];
            T weight = dr * r * e * (T)kPi;
            integrate_transfer_function(g_grid, flux, n_flux, weight);
        }
    }
};

// utils.zig:173-184
template <class T> void inverse_grid_inplace(T* grid, int n, T lo, T hi) {
    Range<T> itt((T)1 / hi, (T)1 / lo, (size_t)n);
    for (int i = 0; i < n; ++i) grid[i] = (T)1 / itt.next();
    for (int i = 0; i < n / 2; ++i) std::swap(grid[i], grid[n - 1 - i]);
}

template <class T> struct Workspace {
    Frame<T> frame, tmp;
    std::vector<T> gstars, g_grid, r_grid, flux_cache;
    T base_lo, base_hi;  // r_grid[0], r_grid[last] of the fresh setup() grid (Q2)
    std::vector<double> conv_energy, conv_flux, flux_copy;
    explicit Workspace(int n_g) {
        // utils.zig:219-232
        gstars.resize(n_g);
        Range<T> it((T)kH, (T)(1.0 - kH), (size_t)n_g);
        for (int i = 0; i < n_g; ++i) gstars[i] = it.next();
        g_grid.resize(kFineG);
        flux_cache.resize(kFineG - 1);
        r_grid.resize(kRadial);
        // xspec-wrapper.zig:84
        inverse_grid_inplace<T>(r_grid.data(), kRadial, (T)1.0, (T)50.0);
        base_lo = r_grid[0];
        base_hi = r_grid[kRadial - 1];
        // xspec-wrapper.zig:98-115 (Q16)
        conv_energy.resize(kConvN);
        Range<double> ce(kConvRes, 2.0, (size_t)kConvN);
        for (int i = 0; i < kConvN; ++i) conv_energy[i] = ce.next();
        conv_flux.resize(kConvN - 1);
    }
};

template <class T> struct SetParams {
    T a, mu, eline, rmin, rmax;
    Emissivity<T> emis;
};

// xspec-wrapper.zig:209-322, 346-359, 395-408
template <class T> SetParams<T> unpack(int model, const double* p) {
    SetParams<T> s;
    const bool conv = model >= 2;
    const int nemis = (model == 0 || model == 2) ? 1 : (model == 4 ? 10 : 5);
    int k = 0;
    s.a = (T)p[k++];
    s.mu = t_cos<T>((T)p[k++] * (T)kRadPerDeg);
    s.eline = conv ? (T)1 : (T)p[k++];
    s.rmin = (T)p[k++];
    s.rmax = (T)p[k++];
    if (nemis == 1) {
        T alpha = (T)p[k++];
        s.emis.n = 0;
        s.emis.index = -alpha;
    } else {
        T rcut = (T)p[k++];
        T alpha = (T)p[k++];
        T w[10];
        for (int i = 0; i < nemis; ++i) w[i] = (T)p[k++];
        s.emis.init_knots(nemis, w, (T)1.0, rcut, alpha);
    }
    return s;
}

// xspec-wrapper.zig:134-183.  lim_lo/lim_hi: the current r_grid end points (Q2 ratchet);
// a fresh process has (base_lo, base_hi).  On return they hold the new end points.
template <class T>
void integrate_lineprofile(const Table& tab, Workspace<T>& ws, const double* energy, int n_flux, double* flux,
                           const SetParams<T>& sp, T& lim_lo, T& lim_hi, bool faithful_cost, LocalStats* st,
                           T* fine_out) {
    // refine_grid :117-132  (iterator type == T in gate mode, f32 in faithful mode: identical
    // because T == f32 there)
    {
        Range<T> itt((T)energy[0], (T)energy[n_flux], (size_t)kFineG);
        T inorm = (T)1 / sp.eline;
        for (int i = 0; i < kFineG; ++i) ws.g_grid[i] = itt.next() * inorm;
    }
    // set_radial_grid :149-153
    inverse_grid_inplace<T>(ws.r_grid.data(), kRadial, t_max<T>(lim_lo, sp.rmin), t_min<T>(sp.rmax, lim_hi));
    lim_lo = ws.r_grid[0];
    lim_hi = ws.r_grid[kRadial - 1];
    for (auto& f : ws.flux_cache) f = 0;

    T p[2] = {sp.a, sp.mu};
    interpolate_parameters<T>(tab, p, ws.frame, ws.tmp);
    Integrator<T> itf;
    itf.tf = &ws.frame;
    itf.n_r = tab.n_r;
    itf.n_g = tab.n_g;
    itf.cache_upper.assign(tab.n_g, 0);
    itf.cache_lower.assign(tab.n_g, 0);
    itf.gstars = ws.gstars.data();
    itf.faithful_cost = faithful_cost;
    itf.st = st;
    itf.integrate(ws.r_grid.data(), kRadial, ws.g_grid.data(), ws.flux_cache.data(), kFineG - 1, sp.emis);
    if (fine_out) std::memcpy(fine_out, ws.flux_cache.data(), sizeof(T) * (kFineG - 1));

    // rebin :162-182 (Q3, Q10).  The reference leaves j unchecked (UB past the grid); we stop at
    // the last fine bin.
    T inorm = (T)1 / sp.eline;
    size_t j = 0;
    double total_flux = 0;
    for (int i = 0; i < n_flux; ++i) {
        flux[i] = 0;
        T summed = 0;
        while (j < (size_t)(kFineG - 1) && (double)ws.g_grid[j] < energy[i] * (double)inorm) {
            summed += ws.flux_cache[j];
            ++j;
        }
        flux[i] += (double)summed;
        double e_mid = 0.5 * (energy[i + 1] + energy[i]);
        flux[i] = flux[i] / e_mid;
        total_flux += flux[i];
    }
    for (int i = 0; i < n_flux; ++i) flux[i] /= total_flux;
}

// convolution.zig:58-102
inline double convolution_weight(size_t start, size_t end, const double* g, const double* a, double low,
                                 double high, bool faithful_cost, LocalStats* st) {
    double weight = 0;
    size_t w0 = start;
    if (!faithful_cost && end > start) {
        // skipped bins (bin_high < low) contribute nothing; jump to the first candidate
        size_t lo = start, hi = end;  // first w with g[w+1] >= low
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (g[mid + 1] < low) lo = mid + 1; else hi = mid;
        }
        w0 = lo;
    }
    for (size_t w = w0; w < end; ++w) {
        double bin_low = g[w], bin_high = g[w + 1];
        double bin_width = bin_high - bin_low;
        if (bin_high < low) continue;
        else if (bin_low > high) break;
        st->conv_terms++;
        double overlap = 1;
        if (bin_low < low && bin_high > high) overlap = (high - low) / bin_width;
        else if (bin_low < low && bin_high < high) overlap = (bin_high - low) / bin_width;
        else if (bin_low > low && bin_high > high) overlap = (high - bin_low) / bin_width;
        weight += a[w] * bin_width * overlap;
    }
    return weight;
}

// convolution.zig:7-56
void convolve(double* output, int n, const double* g, const double* a, int na, const double* x, const double* b,
              bool faithful_cost, LocalStats* st) {
    size_t window_start = 0, window_end = 0;
    for (int i = 0; i < na; ++i) if (a[i] > 0) { window_start = i > 0 ? i - 1 : i; break; }
    for (int i = 0; i < na; ++i) {
        int j = na - i - 1;
        if (a[j] > 0) { window_end = j < na - 1 ? j + 1 : j; break; }
    }
    const double g_min = g[window_start], g_max = g[window_end];
    for (int i = 0; i < n; ++i) output[i] = 0;
    for (int i = 0; i < n; ++i) {
        const double avg = 2 / (x[i + 1] + x[i]);
        for (int j = 0; j < n; ++j) {
            const double low = x[j] * avg, high = x[j + 1] * avg;
            if (high < g_min || low > g_max) continue;
            st->conv_pairs++;
            double weight = convolution_weight(window_start, window_end, g, a, low, high, faithful_cost, st);
            output[j] += weight * b[i];
        }
    }
}

template <class T>
void eval_one(const Table& tab, Workspace<T>& ws, int model, const double* params, const double* energy, int n_flux,
              const double* in_spec, double* out, bool faithful_cost, LocalStats* st, double* lims, T* fine_out,
              double* profile_out) {
    SetParams<T> sp = unpack<T>(model, params);
    T lo = lims ? (T)lims[0] : ws.base_lo, hi = lims ? (T)lims[1] : ws.base_hi;
    if (model < 2) {
        integrate_lineprofile<T>(tab, ws, energy, n_flux, out, sp, lo, hi, faithful_cost, st, fine_out);
    } else {
        // kerr_convolve xspec-wrapper.zig:185-205
        integrate_lineprofile<T>(tab, ws, ws.conv_energy.data(), kConvN - 1, ws.conv_flux.data(), sp, lo, hi,
                                 faithful_cost, st, fine_out);
        if (profile_out) std::memcpy(profile_out, ws.conv_flux.data(), sizeof(double) * (kConvN - 1));
        ws.flux_copy.assign(in_spec, in_spec + n_flux);
        convolve(out, n_flux, ws.conv_energy.data(), ws.conv_flux.data(), kConvN - 1, energy, ws.flux_copy.data(),
                 faithful_cost, st);
    }
    if (lims) { lims[0] = (double)lo; lims[1] = (double)hi; }
}

int model_nparams(int model) {
    switch (model) {
        case 0: return 6;
        case 1: return 12;
        case 2: return 5;
        case 3: return 11;
        case 4: return 16;
    }
    return -1;
}

}  // namespace

extern "C" {

struct orc_table { Table t; };

orc_table* orc_table_create(const float* spins, int n_a, const float* mus, int n_mu, int n_r, int n_g,
                            const float* radii, const float* gmin, const float* gmax, const float* upper,
                            const float* lower) {
    orc_table* h = new orc_table;
    Table& t = h->t;
    t.n_a = n_a; t.n_mu = n_mu; t.n_r = n_r; t.n_g = n_g;
    size_t F = (size_t)n_a * n_mu;
    t.spins.assign(spins, spins + n_a);
    t.mus.assign(mus, mus + n_mu);
    t.radii.assign(radii, radii + F * n_r);
    t.gmin.assign(gmin, gmin + F * n_r);
    t.gmax.assign(gmax, gmax + F * n_r);
    t.upper.assign(upper, upper + F * n_r * n_g);
    t.lower.assign(lower, lower + F * n_r * n_g);
    return h;
}
void orc_table_destroy(orc_table* h) { delete h; }

int orc_model_nparams(int model) { return model_nparams(model); }

// Sensitivity study only (see the policy comment at t_pow): mode 0 restores the default.
void orc_set_transcendental_policy(int mode, int k_ulps, int mask) {
    g_tp_mode.store(mode); g_tp_k.store(k_ulps < 0 ? 0 : k_ulps); g_tp_mask.store(mask & 7);
}
// pow under the current policy, for the unit tests of the policies themselves
double orc_pow_probe(int precision, double x, double y) {
    return precision == 0 ? (double)t_pow<float>((float)x, (float)y) : t_pow<double>(x, y);
}

size_t orc_table_index(const orc_table* h, size_t ia, size_t imu) { return table_index(h->t, ia, imu); }
void orc_find_tables(const orc_table* h, const size_t* pi, size_t* out) { find_tables(h->t, pi, out); }
void orc_find_parameter_indices(const orc_table* h, float a, float mu, size_t* out) {
    float p[2] = {a, mu};
    find_parameter_indices<float>(h->t, p, out);
}
// returns has_weight (0/1); line-profile.zig:196-207
int orc_parameter_weight(const orc_table* h, int which, float v, size_t index, float* w) {
    if (index == 0) return 0;
    const std::vector<float>& arr = which == 0 ? h->t.spins : h->t.mus;
    *w = (v - arr[index - 1]) / (arr[index] - arr[index - 1]);
    return 1;
}

// RangeIterator exposed for the utils.zig:158-171 known-answer test
void orc_range_f32(float lo, float hi, int n, float* out, float* delta) {
    Range<float> r(lo, hi, (size_t)n);
    *delta = r.delta;
    for (int i = 0; i < n; ++i) out[i] = r.next();
}
void orc_range_f64(double lo, double hi, int n, double* out, double* delta) {
    Range<double> r(lo, hi, (size_t)n);
    *delta = r.delta;
    for (int i = 0; i < n; ++i) out[i] = r.next();
}
void orc_gstar_grid(int precision, int n, double* out) {
    if (precision == 0) { Workspace<float> w(n); for (int i = 0; i < n; ++i) out[i] = w.gstars[i]; }
    else { Workspace<double> w(n); for (int i = 0; i < n; ++i) out[i] = w.gstars[i]; }
}
void orc_conv_grid(double* out /*1999*/) {
    Workspace<double> w(2);
    std::memcpy(out, w.conv_energy.data(), sizeof(double) * kConvN);
}
void orc_base_limits(int precision, double* lo, double* hi) {
    if (precision == 0) { Workspace<float> w(2); *lo = w.base_lo; *hi = w.base_hi; }
    else { Workspace<double> w(2); *lo = w.base_lo; *hi = w.base_hi; }
}

// emissivity probe: eps(r) for `n_r` radii.  model_nemis in {1,5,10}.
void orc_emissivity(int precision, int nemis, const double* weights, double rcut, double alpha, const double* r,
                    int n, double* out) {
    auto run = [&](auto tag) {
        using T = decltype(tag);
        Emissivity<T> e;
        if (nemis == 1) { e.n = 0; e.index = -(T)alpha; }
        else { T w[10]; for (int i = 0; i < nemis; ++i) w[i] = (T)weights[i]; e.init_knots(nemis, w, (T)1.0, (T)rcut, (T)alpha); }
        for (int i = 0; i < n; ++i) out[i] = (double)e((T)r[i]);
    };
    if (precision == 0) run(float{}); else run(double{});
}

// Blended frame for (a, incl deg) -- for the K1 parity test.  out: radii,gmin,gmax [n_r], upper,lower [n_r*n_g] as double.
void orc_blend_frame(const orc_table* h, int precision, double a, double incl_deg, double* radii, double* gmin,
                     double* gmax, double* upper, double* lower) {
    auto run = [&](auto tag) {
        using T = decltype(tag);
        Frame<T> f, tmp;
        T p[2] = {(T)a, t_cos<T>((T)incl_deg * (T)kRadPerDeg)};
        interpolate_parameters<T>(h->t, p, f, tmp);
        for (int i = 0; i < h->t.n_r; ++i) { radii[i] = f.radii[i]; gmin[i] = f.gmin[i]; gmax[i] = f.gmax[i]; }
        for (size_t i = 0; i < (size_t)h->t.n_r * h->t.n_g; ++i) { upper[i] = f.upper[i]; lower[i] = f.lower[i]; }
    };
    if (precision == 0) run(float{}); else run(double{});
}

// convolution alone (convolution.zig:7-56)
void orc_convolve(double* output, int n, const double* g, const double* a, int na, const double* x, const double* b,
                  int faithful_cost) {
    LocalStats st;
    convolve(output, n, g, a, na, x, b, faithful_cost != 0, &st);
}

// Batch evaluation.  params [B][P]; in_spec [n_flux] (per_set=0) or [B][n_flux]; out [B][n_flux].
// stats (optional) uint64[7]: bin_visits, active_bins, gauss_bins, edge_evals, integrand_evals, conv_pairs, conv_terms.
// fine_out (optional): [B][1999] doubles of the fine-grid flux before rebinning.
// profile_out (optional, conv models): [B][1998] doubles of the line profile on the conv grid.
// ratchet != 0: emulate the reference's process-global r-grid limits across the B calls (Q2),
// starting from lims[2] (in/out); otherwise each set sees a fresh process.
int orc_eval_batch(const orc_table* h, int model, int precision, long B, const double* params, const double* energy,
                   int n_flux, const double* in_spec, int per_set, double* out, int n_threads, int faithful_cost,
                   uint64_t* stats, double* fine_out, double* profile_out, int ratchet, double* lims) {
    const int P = model_nparams(model);
    if (P < 0 || n_flux < 1) return 1;
    if (model >= 2 && !in_spec) return 2;
    if (n_threads < 1) n_threads = 1;
    if (ratchet) n_threads = 1;
    std::atomic<long> next{0};
    std::vector<LocalStats> lst((size_t)n_threads);
    auto worker = [&](int tid, auto tag) {
        using T = decltype(tag);
        Workspace<T> ws(h->t.n_g);
        std::vector<T> fine(kFineG - 1);
        double rl[2];
        if (ratchet) { rl[0] = lims[0]; rl[1] = lims[1]; }
        for (;;) {
            long s = next.fetch_add(1);
            if (s >= B) break;
            const double* spec = in_spec ? (per_set ? in_spec + (size_t)s * n_flux : in_spec) : nullptr;
            eval_one<T>(h->t, ws, model, params + (size_t)s * P, energy, n_flux, spec, out + (size_t)s * n_flux,
                        faithful_cost != 0, &lst[tid], ratchet ? rl : nullptr, fine_out ? fine.data() : nullptr,
                        profile_out ? profile_out + (size_t)s * (kConvN - 1) : nullptr);
            if (fine_out) for (int i = 0; i < kFineG - 1; ++i) fine_out[(size_t)s * (kFineG - 1) + i] = (double)fine[i];
        }
        if (ratchet) { lims[0] = rl[0]; lims[1] = rl[1]; }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) {
        if (precision == 0) th.emplace_back([&, t] { worker(t, float{}); });
        else th.emplace_back([&, t] { worker(t, double{}); });
    }
    for (auto& t : th) t.join();
    if (stats) {
        for (int i = 0; i < 7; ++i) stats[i] = 0;
        for (auto& s : lst) {
            stats[0] += s.bin_visits; stats[1] += s.active_bins; stats[2] += s.gauss_bins; stats[3] += s.edge_evals;
            stats[4] += s.integrand_evals; stats[5] += s.conv_pairs; stats[6] += s.conv_terms;
        }
    }
    return 0;
}

}  // extern "C"
