// Microbenchmark: the packed Gauss body of k_quad (gauss_share_x2) in isolation -- cycles per call per SM sub-partition
// with 8 resident warps per SMSP, against the scalar SHARE body.  Tells how much of a visit's ~560 cycles is the rule itself.
#include <cstdio>
#include "../../lineprofiles_b200/csrc/klp_kernels.cuh"
using namespace klp;


// ---- experimental variants of the Gauss body (NOT product code; results of some are deliberately wrong) ----
template <int V> __device__ __forceinline__ float gauss_var(const RadCtx<float>& c, const float glow, const float ghigh) {
    const float xp[3] = {(float)(9.4910791234275852452618968404809e-01 + 1.0), (float)(7.415311855993944398638647732811e-01 + 1.0),
                         (float)(4.0584515137739716690660641207707e-01 + 1.0)};
    const float xm[3] = {(float)(-9.4910791234275852452618968404809e-01 + 1.0), (float)(-7.415311855993944398638647732811e-01 + 1.0),
                         (float)(-4.0584515137739716690660641207707e-01 + 1.0)};
    const float wk[3] = {(float)1.2948496616886969327061143267787e-01, (float)2.797053914892766679014677714229e-01,
                         (float)3.8183005050511894495036977548818e-01};
    const float scale = __fmul_rn(__fsub_rn(ghigh, glow), 0.5f);
    const float2 dg2 = P2::bc(c.dg), ydg2 = P2::bc(c.ydg), gmin2 = P2::bc(c.gmin), glow2 = P2::bc(glow), scale2 = P2::bc(scale);
    float2 x[3], s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x[k] = P2::add_prod(P2::mul(make_float2(xp[k], xm[k]), scale2), glow2, c.one);
        s[k] = P2::div_by(P2::sub(x[k], gmin2), dg2, ydg2);
    }
    const float x0 = __fadd_rn(scale, glow);
    const float s0 = div_by<float, true>(__fsub_rn(x0, c.gmin), c.dg, c.ydg);
    const int idx_lo = knot_index_fast<float>(c, s[0].y);
    const float gs_lo = c.gs[idx_lo];
    const int idx_hi = idx_lo + ((gs_lo < s[0].x && idx_lo < c.n_g - 1) ? 1 : 0);
    const float4 kn_lo = c.knots[idx_lo], e_lo = c.row[idx_lo];
    const float4 kn_hi = c.knots[idx_hi], e_hi = c.row[idx_hi];
    float sum = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float2 v;
        if (V == 1) {   // no selects at all (wrong when a bin straddles a knot): cost of the FSEL/MOV pair assembly
            v = integrand_at_x2(c.dg, c.one, x[k], s[k], kn_lo, e_lo, kn_lo, e_lo);
        } else {
            const bool p1 = k == 0 ? false : gs_lo >= s[k].x;
            const bool p2 = k == 0 ? true : gs_lo >= s[k].y;
            v = integrand_at_x2(c.dg, c.one, x[k], s[k], sel4(p1, kn_lo, kn_hi), sel4(p1, e_lo, e_hi), sel4(p2, kn_lo, kn_hi), sel4(p2, e_lo, e_hi));
        }
        sum = __fadd_rn(sum, __fmul_rn(__fadd_rn(v.x, v.y), __fmul_rn(wk[k], scale)));
    }
    if (V != 2) {   // V == 2: no centre node (wrong): cost of the scalar centre evaluation
        const bool p0 = gs_lo >= s0;
        const float4 kn = V == 1 ? kn_lo : sel4(p0, kn_lo, kn_hi), e = V == 1 ? e_lo : sel4(p0, e_lo, e_hi);
        const float factor = div_by<float, true>(__fsub_rn(s0, kn.x), kn.y, kn.z);
        const float lower = __fadd_rn(__fmul_rn(factor, e.y), e.x);
        const float upper = __fadd_rn(__fmul_rn(factor, e.w), e.z);
        const float f = __fadd_rn(upper, lower);
        const float denom = __fmul_rn(sqrt_gen<true>(__fmul_rn(s0, __fsub_rn(1.0f, s0))), c.dg);
        const float v0 = div_gen<true>(__fmul_rn(__fmul_rn(__fmul_rn(f, x0), x0), x0), denom);
        sum = __fadd_rn(sum, __fmul_rn(v0, (float)4.1795918367346938775510204081658e-01));
    }
    return sum;
}

// variant: SoA tables, every packed operand loaded directly as two 32-bit shared loads (no FSEL / MOV assembly)
struct SoA { const float *g0, *d, *y, *l0, *dl, *u0, *du; };
__device__ __forceinline__ float2 integrand_soa_x2(const float dg, const float one, const float2 g, const float2 gstar, const SoA& t, int i1, int i2) {
    const float2 factor = P2::div_by(P2::sub(gstar, make_float2(t.g0[i1], t.g0[i2])), make_float2(t.d[i1], t.d[i2]), make_float2(t.y[i1], t.y[i2]));
    const float2 lower = P2::add_prod(P2::mul(factor, make_float2(t.dl[i1], t.dl[i2])), make_float2(t.l0[i1], t.l0[i2]), one);
    const float2 upper = P2::add_prod(P2::mul(factor, make_float2(t.du[i1], t.du[i2])), make_float2(t.u0[i1], t.u0[i2]), one);
    const float2 f = P2::add(upper, lower);
    const float2 denom = P2::mul(P2::sqrt_gen(P2::mul(gstar, P2::sub(P2::bc(1.0f), gstar))), P2::bc(dg));
    return P2::div_gen(P2::mul(P2::mul(P2::mul(f, g), g), g), denom);
}
__device__ __forceinline__ float gauss_soa(const RadCtx<float>& c, const SoA& t, const float glow, const float ghigh) {
    const float xp[3] = {(float)(9.4910791234275852452618968404809e-01 + 1.0), (float)(7.415311855993944398638647732811e-01 + 1.0),
                         (float)(4.0584515137739716690660641207707e-01 + 1.0)};
    const float xm[3] = {(float)(-9.4910791234275852452618968404809e-01 + 1.0), (float)(-7.415311855993944398638647732811e-01 + 1.0),
                         (float)(-4.0584515137739716690660641207707e-01 + 1.0)};
    const float wk[3] = {(float)1.2948496616886969327061143267787e-01, (float)2.797053914892766679014677714229e-01,
                         (float)3.8183005050511894495036977548818e-01};
    const float scale = __fmul_rn(__fsub_rn(ghigh, glow), 0.5f);
    const float2 dg2 = P2::bc(c.dg), ydg2 = P2::bc(c.ydg), gmin2 = P2::bc(c.gmin), glow2 = P2::bc(glow), scale2 = P2::bc(scale);
    float2 x[3], s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x[k] = P2::add_prod(P2::mul(make_float2(xp[k], xm[k]), scale2), glow2, c.one);
        s[k] = P2::div_by(P2::sub(x[k], gmin2), dg2, ydg2);
    }
    const float x0 = __fadd_rn(scale, glow);
    const float s0 = div_by<float, true>(__fsub_rn(x0, c.gmin), c.dg, c.ydg);
    const int idx_lo = knot_index_fast<float>(c, s[0].y);
    const float gs_lo = c.gs[idx_lo];
    const int idx_hi = idx_lo + ((gs_lo < s[0].x && idx_lo < c.n_g - 1) ? 1 : 0);
    float sum = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i1 = k == 0 ? idx_hi : (gs_lo >= s[k].x ? idx_lo : idx_hi);
        const int i2 = k == 0 ? idx_lo : (gs_lo >= s[k].y ? idx_lo : idx_hi);
        const float2 v = integrand_soa_x2(c.dg, c.one, x[k], s[k], t, i1, i2);
        sum = __fadd_rn(sum, __fmul_rn(__fadd_rn(v.x, v.y), __fmul_rn(wk[k], scale)));
    }
    {
        const int i0 = gs_lo >= s0 ? idx_lo : idx_hi;
        const float factor = div_by<float, true>(__fsub_rn(s0, t.g0[i0]), t.d[i0], t.y[i0]);
        const float lower = __fadd_rn(__fmul_rn(factor, t.dl[i0]), t.l0[i0]);
        const float upper = __fadd_rn(__fmul_rn(factor, t.du[i0]), t.u0[i0]);
        const float f = __fadd_rn(upper, lower);
        const float denom = __fmul_rn(sqrt_gen<true>(__fmul_rn(s0, __fsub_rn(1.0f, s0))), c.dg);
        const float v0 = div_gen<true>(__fmul_rn(__fmul_rn(__fmul_rn(f, x0), x0), x0), denom);
        sum = __fadd_rn(sum, __fmul_rn(v0, (float)4.1795918367346938775510204081658e-01));
    }
    return sum;
}

// variant: one per-warp SoA block {g0, d, y, l0, dl, u0, du} x 32 floats: one address per index, immediate offsets
__device__ __forceinline__ float2 integrand_soa1_x2(const float dg, const float one, const float2 g, const float2 gstar, const float* t1, const float* t2) {
    const float2 factor = P2::div_by(P2::sub(gstar, make_float2(t1[0], t2[0])), make_float2(t1[32], t2[32]), make_float2(t1[64], t2[64]));
    const float2 lower = P2::add_prod(P2::mul(factor, make_float2(t1[128], t2[128])), make_float2(t1[96], t2[96]), one);
    const float2 upper = P2::add_prod(P2::mul(factor, make_float2(t1[192], t2[192])), make_float2(t1[160], t2[160]), one);
    const float2 f = P2::add(upper, lower);
    const float2 denom = P2::mul(P2::sqrt_gen(P2::mul(gstar, P2::sub(P2::bc(1.0f), gstar))), P2::bc(dg));
    return P2::div_gen(P2::mul(P2::mul(P2::mul(f, g), g), g), denom);
}
__device__ __forceinline__ float gauss_soa1(const RadCtx<float>& c, const float* tab, const float glow, const float ghigh) {
    const float xp[3] = {(float)(9.4910791234275852452618968404809e-01 + 1.0), (float)(7.415311855993944398638647732811e-01 + 1.0),
                         (float)(4.0584515137739716690660641207707e-01 + 1.0)};
    const float xm[3] = {(float)(-9.4910791234275852452618968404809e-01 + 1.0), (float)(-7.415311855993944398638647732811e-01 + 1.0),
                         (float)(-4.0584515137739716690660641207707e-01 + 1.0)};
    const float wk[3] = {(float)1.2948496616886969327061143267787e-01, (float)2.797053914892766679014677714229e-01,
                         (float)3.8183005050511894495036977548818e-01};
    const float scale = __fmul_rn(__fsub_rn(ghigh, glow), 0.5f);
    const float2 dg2 = P2::bc(c.dg), ydg2 = P2::bc(c.ydg), gmin2 = P2::bc(c.gmin), glow2 = P2::bc(glow), scale2 = P2::bc(scale);
    float2 x[3], s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x[k] = P2::add_prod(P2::mul(make_float2(xp[k], xm[k]), scale2), glow2, c.one);
        s[k] = P2::div_by(P2::sub(x[k], gmin2), dg2, ydg2);
    }
    const float x0 = __fadd_rn(scale, glow);
    const float s0 = div_by<float, true>(__fsub_rn(x0, c.gmin), c.dg, c.ydg);
    const int idx_lo = knot_index_fast<float>(c, s[0].y);
    const float gs_lo = c.gs[idx_lo];
    const float* t_lo = tab + idx_lo;
    const float* t_hi = t_lo + ((gs_lo < s[0].x && idx_lo < c.n_g - 1) ? 1 : 0);
    float sum = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* t1 = k == 0 ? t_hi : (gs_lo >= s[k].x ? t_lo : t_hi);
        const float* t2 = k == 0 ? t_lo : (gs_lo >= s[k].y ? t_lo : t_hi);
        const float2 v = integrand_soa1_x2(c.dg, c.one, x[k], s[k], t1, t2);
        sum = __fadd_rn(sum, __fmul_rn(__fadd_rn(v.x, v.y), __fmul_rn(wk[k], scale)));
    }
    {
        const float* t0 = gs_lo >= s0 ? t_lo : t_hi;
        const float factor = div_by<float, true>(__fsub_rn(s0, t0[0]), t0[32], t0[64]);
        const float lower = __fadd_rn(__fmul_rn(factor, t0[128]), t0[96]);
        const float upper = __fadd_rn(__fmul_rn(factor, t0[192]), t0[160]);
        const float f = __fadd_rn(upper, lower);
        const float denom = __fmul_rn(sqrt_gen<true>(__fmul_rn(s0, __fsub_rn(1.0f, s0))), c.dg);
        const float v0 = div_gen<true>(__fmul_rn(__fmul_rn(__fmul_rn(f, x0), x0), x0), denom);
        sum = __fadd_rn(sum, __fmul_rn(v0, (float)4.1795918367346938775510204081658e-01));
    }
    return sum;
}

// variant: pair tables.  Entry (two float4) of table "same" at k: every value duplicated (v[k], v[k]); of table "step" at k:
// (v[k+1], v[k]).  A packed pair of nodes (higher node in .x) whose intervals are (iH, iL), iH in {iL, iL+1}, reads its
// operands with 4 LDS.128 from ONE address: (iH != iL ? step : same)[iL].  Layout of an entry (16 floats):
//   {g0.x g0.y d.x d.y} {y.x y.y - -} {l0.x l0.y dl.x dl.y} {u0.x u0.y du.x du.y}
__device__ __forceinline__ float2 integrand_pt_x2(const float dg, const float one, const float2 g, const float2 gstar, const float4* e) {
    const float4 a = e[0], b = e[1], c2 = e[2], d2 = e[3];
    const float2 factor = P2::div_by(P2::sub(gstar, make_float2(a.x, a.y)), make_float2(a.z, a.w), make_float2(b.x, b.y));
    const float2 lower = P2::add_prod(P2::mul(factor, make_float2(c2.z, c2.w)), make_float2(c2.x, c2.y), one);
    const float2 upper = P2::add_prod(P2::mul(factor, make_float2(d2.z, d2.w)), make_float2(d2.x, d2.y), one);
    const float2 f = P2::add(upper, lower);
    const float2 denom = P2::mul(P2::sqrt_gen(P2::mul(gstar, P2::sub(P2::bc(1.0f), gstar))), P2::bc(dg));
    return P2::div_gen(P2::mul(P2::mul(P2::mul(f, g), g), g), denom);
}
__device__ __forceinline__ float gauss_pt(const RadCtx<float>& c, const float4* same, const float4* step, const float glow, const float ghigh) {
    const float xp[3] = {(float)(9.4910791234275852452618968404809e-01 + 1.0), (float)(7.415311855993944398638647732811e-01 + 1.0),
                         (float)(4.0584515137739716690660641207707e-01 + 1.0)};
    const float xm[3] = {(float)(-9.4910791234275852452618968404809e-01 + 1.0), (float)(-7.415311855993944398638647732811e-01 + 1.0),
                         (float)(-4.0584515137739716690660641207707e-01 + 1.0)};
    const float wk[3] = {(float)1.2948496616886969327061143267787e-01, (float)2.797053914892766679014677714229e-01,
                         (float)3.8183005050511894495036977548818e-01};
    const float scale = __fmul_rn(__fsub_rn(ghigh, glow), 0.5f);
    const float2 dg2 = P2::bc(c.dg), ydg2 = P2::bc(c.ydg), gmin2 = P2::bc(c.gmin), glow2 = P2::bc(glow), scale2 = P2::bc(scale);
    float2 x[3], s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x[k] = P2::add_prod(P2::mul(make_float2(xp[k], xm[k]), scale2), glow2, c.one);
        s[k] = P2::div_by(P2::sub(x[k], gmin2), dg2, ydg2);
    }
    const float x0 = __fadd_rn(scale, glow);
    const float s0 = div_by<float, true>(__fsub_rn(x0, c.gmin), c.dg, c.ydg);
    const int idx_lo = knot_index_fast<float>(c, s[0].y);
    const float gs_lo = c.gs[idx_lo];
    const bool strad = gs_lo < s[0].x && idx_lo < c.n_g - 1;     // the bin straddles the knot gs_lo
    const float4* e_ll = same + 4 * idx_lo;                       // both nodes in the lower interval
    const float4* e_hl = (strad ? step : same) + 4 * idx_lo;      // higher node above the knot, lower node below
    const float4* e_hh = same + 4 * (idx_lo + (strad ? 1 : 0));   // both above
    float sum = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4* e = k == 0 ? e_hl : (gs_lo >= s[k].x ? e_ll : (gs_lo >= s[k].y ? e_hl : e_hh));
        const float2 v = integrand_pt_x2(c.dg, c.one, x[k], s[k], e);
        sum = __fadd_rn(sum, __fmul_rn(__fadd_rn(v.x, v.y), __fmul_rn(wk[k], scale)));
    }
    {
        const float4* e = gs_lo >= s0 ? e_ll : e_hh;
        const float4 a = e[0], b = e[1], c2 = e[2], d2 = e[3];
        const float factor = div_by<float, true>(__fsub_rn(s0, a.x), a.z, b.x);
        const float lower = __fadd_rn(__fmul_rn(factor, c2.z), c2.x);
        const float upper = __fadd_rn(__fmul_rn(factor, d2.z), d2.x);
        const float f = __fadd_rn(upper, lower);
        const float denom = __fmul_rn(sqrt_gen<true>(__fmul_rn(s0, __fsub_rn(1.0f, s0))), c.dg);
        const float v0 = div_gen<true>(__fmul_rn(__fmul_rn(__fmul_rn(f, x0), x0), x0), denom);
        sum = __fadd_rn(sum, __fmul_rn(v0, (float)4.1795918367346938775510204081658e-01));
    }
    return sum;
}

template <int MODE> __global__ void __launch_bounds__(512, KLP_UB_MINB) k(float* sink, int iters, float one, float dgv) {
    __shared__ float4 s_knots[32];
    __shared__ float s_gs[32];
    __shared__ float4 s_rows[16 * 30];
    __shared__ float s_soa[7 * 32 + 16 * 4 * 32];
    __shared__ float s_soa1[16 * 7 * 32];
    extern __shared__ float4 s_pt[];   // [16 warps][2 tables][30][4]
    const int n_g = 30, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < n_g) {
        const float gs = 5e-4f + tid * (0.999f / 29);
        s_gs[tid] = gs;
        const float gp = 5e-4f + (tid - 1) * (0.999f / 29);
        s_knots[tid] = tid >= 2 ? make_float4(gp, gs - gp, 1.0f / (gs - gp), 0.f) : make_float4(0.f, 1.f, 1.f, 0.f);
    }
    for (int k2 = lane; k2 < n_g; k2 += 32) s_rows[warp * n_g + k2] = make_float4(1.0f + 0.01f * k2, 0.01f, 0.8f + 0.01f * k2, 0.012f);
    if (tid < n_g) {
        const float gs = 5e-4f + tid * (0.999f / 29), gp = 5e-4f + (tid - 1) * (0.999f / 29);
        s_soa[tid] = tid >= 2 ? gp : 0.f; s_soa[32 + tid] = tid >= 2 ? gs - gp : 1.f; s_soa[64 + tid] = tid >= 2 ? 1.0f / (gs - gp) : 1.f;
    }
    for (int k2 = lane; k2 < n_g; k2 += 32) {
        float* r = s_soa + 7 * 32 + warp * 128;
        r[k2] = 1.0f + 0.01f * k2; r[32 + k2] = 0.01f; r[64 + k2] = 0.8f + 0.01f * k2; r[96 + k2] = 0.012f;
    }
    for (int k2 = lane; k2 < n_g; k2 += 32) {
        float* r = s_soa1 + warp * 224;
        const float gs = 5e-4f + k2 * (0.999f / 29), gp = 5e-4f + (k2 - 1) * (0.999f / 29);
        r[k2] = k2 >= 2 ? gp : 0.f; r[32 + k2] = k2 >= 2 ? gs - gp : 1.f; r[64 + k2] = k2 >= 2 ? 1.0f / (gs - gp) : 1.f;
        r[96 + k2] = 1.0f + 0.01f * k2; r[128 + k2] = 0.01f; r[160 + k2] = 0.8f + 0.01f * k2; r[192 + k2] = 0.012f;
    }
    for (int k2 = lane; k2 < n_g; k2 += 32) {
        float4* sm = s_pt + warp * 240 + 4 * k2; float4* st = sm + 120;
        sm[0] = make_float4(0.01f * k2, 0.01f * k2, 0.0344f, 0.0344f); sm[1] = make_float4(29.f, 29.f, 0, 0);
        sm[2] = make_float4(1.f, 1.f, 0.01f, 0.01f); sm[3] = make_float4(0.8f, 0.8f, 0.012f, 0.012f);
        st[0] = sm[0]; st[1] = sm[1]; st[2] = sm[2]; st[3] = sm[3];
    }
    __syncthreads();
    SoA soa;
    soa.g0 = s_soa; soa.d = s_soa + 32; soa.y = s_soa + 64;
    soa.l0 = s_soa + 7 * 32 + warp * 128; soa.dl = soa.l0 + 32; soa.u0 = soa.l0 + 64; soa.du = soa.l0 + 96;
    RadCtx<float> c;
    c.row = s_rows + warp * n_g; c.knots = s_knots; c.gs = s_gs; c.n_g = n_g;
    c.est_inv_delta = 29.0f / 0.999f; c.est_off = -5e-4f * c.est_inv_delta + (0.5f - 2e-4f);
    c.gmin = 0.3f; c.dg = dgv; c.ydg = 1.0f / dgv; c.tA = 0; c.tD = 10; c.one = one;
    float acc = 0;
    float glow = 0.35f + lane * 7.6e-4f;
    for (int it = 0; it < iters; ++it) {
        const float ghigh = glow + 7.6e-4f;
        float v;
        if (MODE == 0) v = gauss_share_x2(c, glow, ghigh);
        else if (MODE == 1) v = integrate_g_bin_active<float, true, true>(c, glow, ghigh);
        else if (MODE == 20) v = gauss_soa(c, soa, glow, ghigh);
        else if (MODE == 21) v = gauss_soa1(c, s_soa1 + warp * 224, glow, ghigh);
        else if (MODE == 22) v = gauss_pt(c, s_pt + warp * 240, s_pt + warp * 240 + 120, glow, ghigh);
        else v = gauss_var<MODE - 10>(c, glow, ghigh);
        acc += v;
        glow += 1e-7f * (v > 1e30f ? 1.f : 0.5f);   // keep the loop-carried dependence light but real
    }
    if (acc == -1.2345f) sink[0] = acc;
}
template <int MODE> void run(const char* name, int ctas_per_sm = 2) {
    float* sink; cudaMalloc(&sink, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 2000, grid = p.multiProcessorCount * ctas_per_sm;   // CTAs x 16 warps: 4 warps per SMSP each
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 240 * 16);
    k<MODE><<<grid, 512, 16 * 240 * 16>>>(sink, 10, 1.0f, 0.6f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<grid, 512, 16 * 240 * 16>>>(sink, iters, 1.0f, 0.6f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-36s %2d warps/SMSP: %8.1f cycles per call per SMSP  %s\n", name, 4 * ctas_per_sm, cyc / (4.0 * ctas_per_sm * iters), cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}
int main() {
    run<0>("packed (product)"); run<1>("integrate_g_bin_active (KLP_PACKED_F32 as compiled)"); run<10>("variant 0 = packed copy");
    run<20>("variant: SoA tables, direct 32-bit loads");
    run<21>("variant: one SoA block per warp, imm offsets");
    run<22>("variant: pair tables (4 LDS.128 per pair)");
    run<11>("variant 1: no selects (wrong)"); run<12>("variant 2: no centre node (wrong)");
    return 0;
}
