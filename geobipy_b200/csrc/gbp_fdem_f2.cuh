// fp32 production instantiation of fdem_eval(): packed fp32x2 arithmetic, two (forward + Jacobian) or four (forward
// only) filter abscissae per lane.
//
// Blackwell (sm_100) has two-wide fp32 instructions (FFMA2 / FMUL2 / FADD2 on 64-bit register pairs, with free
// negation / absolute-value operand modifiers).  Each lane carries the admittance recursion of abscissae l and l + 32
// of a 64-abscissa CHUNK in the two halves of float2 registers; the forward-only pass runs two chunks at once as two
// independent dependency chains (the sampler is bound by dependent-issue latency, not by any pipe: profiles/).
// Special functions (MUFU rsq / sqrt / rcp / ex2 / sin / cos) stay scalar, two per pair.
//
// Same mathematics as the generic template in gbp_fdem.cuh (the fp64 validation path), rearranged so that a layer
// costs ONE complex reciprocal and no tanh: with e = exp(-2 u t), d = y - u (y = admittance of the stack below)
//     y' = u (y + u tanh(ut)) / (u + y tanh(ut)) = u (2y + (e-1) d) / (2u - (e-1) d)
// and, for the Jacobian (the reference's M1_1 form, fdem1d_numba.py:223-301, quirk (viii) of DESIGN.md included),
//     dy'/dy      = 4 u^2 e / den^2,                                   den = 2u - (e-1) d
//     dy'/dln(s)  = (i b / 2) [ s^2 - 4 u e d (1 + t s) + e^2 (4 u^2 - d^2) ] / (u den^2),   s = y + u
// (obtained from the tanh expressions with tanh = (1-e)/(1+e); 1 - tanh^2 = 4e/(1+e)^2).  exp(-2ut) underflows to an
// exact 0 beyond 2 Re(u) t = 87 (ex2.approx.ftz), where tanh = 1 to fp32 precision, and |Im(2ut)| <= Re(2ut), so the
// phase needs no range reduction beyond the one MUFU.SIN/COS apply themselves (error <= x e^-x 2^-24 on the product).
//
// Table layout (gbp_tables.h): per chunk four float4 per lane, quad-major (conflict-free LDS.128):
//   quad 0 = (lam_a, lam_b, u0r_a, u0r_b)  quad 1 = (u0i_a, u0i_b, er_a, er_b)
//   quad 2 = (ei_a, ei_b, cr_a, cr_b)      quad 3 = (ci_a, ci_b, -, -)
#pragma once
#include "gbp_fdem.cuh"

namespace gbp {
namespace f2 {

typedef float2 v2;
struct c2 {
    v2 re, im;
};
__device__ __forceinline__ v2 V(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ v2 S(float a) { return make_float2(a, a); }
__device__ __forceinline__ v2 neg(v2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ v2 mul(v2 a, v2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ v2 add(v2 a, v2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ v2 sub(v2 a, v2 b) { return __fadd2_rn(a, neg(b)); }
__device__ __forceinline__ v2 fma(v2 a, v2 b, v2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ v2 rcp(v2 a) { return V(rt<float>::rcp(a.x), rt<float>::rcp(a.y)); }
__device__ __forceinline__ v2 sqrt(v2 a) { return V(rt<float>::sqrt(a.x), rt<float>::sqrt(a.y)); }
__device__ __forceinline__ float rsqrt1(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ v2 rsqrt(v2 a) { return V(rsqrt1(a.x), rsqrt1(a.y)); }

__device__ __forceinline__ c2 operator+(c2 a, c2 b) { return {add(a.re, b.re), add(a.im, b.im)}; }
__device__ __forceinline__ c2 operator-(c2 a, c2 b) { return {sub(a.re, b.re), sub(a.im, b.im)}; }
__device__ __forceinline__ c2 operator*(c2 a, c2 b)
{
    return {fma(a.re, b.re, neg(mul(a.im, b.im))), fma(a.re, b.im, mul(a.im, b.re))};
}
__device__ __forceinline__ c2 operator*(c2 a, v2 s) { return {mul(a.re, s), mul(a.im, s)}; }
__device__ __forceinline__ c2 csq(c2 a) { return {fma(a.re, a.re, neg(mul(a.im, a.im))), mul(add(a.re, a.re), a.im)}; }
__device__ __forceinline__ c2 cinv(c2 a)
{
    const v2 d = rcp(fma(a.re, a.re, mul(a.im, a.im)));
    return {mul(a.re, d), neg(mul(a.im, d))};
}
// u = sqrt(a + ib), b > 0 (first quadrant), both halves; m = |u|^2 = |a + ib|.  2 MUFU per half:
// with x = (m + |a|)/2:  sqrt(x) = x rsq(x),  b / (2 sqrt(x)) = (b/2) rsq(x)
__device__ __forceinline__ c2 csqrt_q1(v2 a, float b, v2& m)
{
    m = sqrt(fma(a, a, S(b * b)));
    const v2 x = mul(S(0.5f), add(m, V(fabsf(a.x), fabsf(a.y))));
    const v2 r = rsqrt(x);
    const v2 t = mul(x, r);
    const v2 o = mul(S(0.5f * b), r);
    c2 u;
    u.re = V(a.x >= 0.f ? t.x : o.x, a.y >= 0.f ? t.y : o.y);
    u.im = V(a.x >= 0.f ? o.x : t.x, a.y >= 0.f ? o.y : t.y);
    return u;
}
__device__ __forceinline__ c2 csqrt_q1(v2 a, v2 b)   // per-half b (time-domain kernel: lane = frequency)
{
    const v2 m = sqrt(fma(a, a, mul(b, b)));
    const v2 x = mul(S(0.5f), add(m, V(fabsf(a.x), fabsf(a.y))));
    const v2 r = rsqrt(x);
    const v2 t = mul(x, r);
    const v2 o = mul(mul(S(0.5f), b), r);
    c2 u;
    u.re = V(a.x >= 0.f ? t.x : o.x, a.y >= 0.f ? t.y : o.y);
    u.im = V(a.x >= 0.f ? o.x : t.x, a.y >= 0.f ? o.y : t.y);
    return u;
}
// exp(z); the phase goes to MUFU.SIN / MUFU.COS as it is (callers: |Im z| <= |Re z| or tiny)
__device__ __forceinline__ c2 cexp(c2 z)
{
    const v2 l = mul(z.re, S(1.4426950408889634f));
    v2 e, s, c;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(l.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(l.y));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s.x) : "f"(z.im.x));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s.y) : "f"(z.im.y));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c.x) : "f"(z.im.x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c.y) : "f"(z.im.y));
    return {mul(e, c), mul(e, s)};
}

// one finite layer, forward only: y <- admittance at the top of the layer (b = omega mu sigma, t2 = -2 thickness)
__device__ __forceinline__ void layer_fwd(const v2 a, const float b, const float t2, c2& y)
{
    v2 m;
    const c2 u = csqrt_q1(a, b, m);
    const c2 e = cexp(c2{mul(S(t2), u.re), mul(S(t2), u.im)});
    const c2 p = c2{add(e.re, S(-1.f)), e.im} * (y - u);                             // (e - 1)(y - u)
    const c2 den = {fma(S(2.f), u.re, neg(p.re)), fma(S(2.f), u.im, neg(p.im))};     // (1+e) u + (1-e) y
    const c2 num = {fma(S(2.f), y.re, p.re), fma(S(2.f), y.im, p.im)};               // (1+e) y + (1-e) u
    y = (u * num) * cinv(den);
}

// the air-earth interface of one abscissa pair: rTE * c * exp(e_j hDiff); `is` = 1 / (u0 + y1) for the Jacobian
struct TopOut {
    c2 term, K, is, u0;
};
__device__ __forceinline__ TopOut top_term(const float4 q0, const float4 q1, const float4 q2, const float4 q3, const c2 y,
                                           const float hd)
{
    TopOut o;
    o.u0 = c2{V(q0.z, q0.w), V(q1.x, q1.y)};
    o.is = cinv(o.u0 + y);
    const c2 rte = (o.u0 - y) * o.is;
    o.K = c2{V(q2.z, q2.w), V(q3.x, q3.y)} * cexp(c2{mul(V(q1.z, q1.w), S(hd)), mul(V(q2.x, q2.y), S(hd))});
    o.term = rte * o.K;
    return o;
}

// the Hankel sum of one frequency is complete: both warp sums in one folded reduction (same bits as two warp_sum calls)
__device__ __forceinline__ void close_frequency(const c2& acc, float* __restrict__ pred, const int F, const int f, const int lane)
{
    const float v = warp_sum2(acc.re.x + acc.re.y, acc.im.x + acc.im.y, lane);
    if ((lane & 15) == 0) pred[(lane >> 4) * F + f] = v;
}

}  // namespace f2

// ---------------------------------------------------------------- forward only: two chunks per pass
// chunks [c0, c1) only (a whole forward: 0, Q.n_chunks; one half of it: Q.half_begin[h], Q.half_begin[h + 1])
__device__ __noinline__ void fdem_fwd_f2(const SysShared<float>& Q, const float* __restrict__ tab, float alt, int L,
                                         const float* __restrict__ msig, const float* __restrict__ mthk,
                                         float* __restrict__ pred, const int c0, const int c1)
{
    using namespace f2;
    __builtin_assume(__isShared(&Q));
    __builtin_assume(__isShared(tab));
    __builtin_assume(__isShared(msig));
    __builtin_assume(__isShared(mthk));
    __builtin_assume(__isShared(pred));
    const int lane = threadIdx.x & 31;
    const int F = Q.n_freq, NCH = c1;
    if (c0 >= c1) return;
    const float4* T4 = reinterpret_cast<const float4*>(tab) + lane;
    const float sL = msig[L - 1];
    int cur_f = Q.chunk_freq[c0];
    c2 acc = {S(0.f), S(0.f)};
#pragma unroll 1
    for (int c = c0; c < NCH; c += 2) {
        const bool hasB = c + 1 < NCH;
        const int cB = hasB ? c + 1 : c;
        const int fA = Q.chunk_freq[c], fB = Q.chunk_freq[cB];
        const float4 qa0 = T4[(c * 4 + 0) * 32], qb0 = T4[(cB * 4 + 0) * 32];
        const float omuA = Q.omu[fA], omuB = Q.omu[fB];
        const v2 lamA = V(qa0.x, qa0.y), lamB = V(qb0.x, qb0.y);
        const v2 aA = fma(lamA, lamA, S(Q.k2re[fA])), aB = fma(lamB, lamB, S(Q.k2re[fB]));   // Re(u^2) of every earth layer
        v2 mA, mB;
        c2 yA = csqrt_q1(aA, omuA * sL, mA), yB = csqrt_q1(aB, omuB * sL, mB);               // basement: y_L = u_L
#pragma unroll 1
        for (int k = L - 2; k >= 0; --k) {
            const float sg = msig[k], t2 = -2.f * mthk[k];
            layer_fwd(aA, omuA * sg, t2, yA);
            layer_fwd(aB, omuB * sg, t2, yB);
        }
        const TopOut tA = top_term(qa0, T4[(c * 4 + 1) * 32], T4[(c * 4 + 2) * 32], T4[(c * 4 + 3) * 32], yA,
                                   Q.hd0[fA] - 2.f * alt);
        const TopOut tB = top_term(qb0, T4[(cB * 4 + 1) * 32], T4[(cB * 4 + 2) * 32], T4[(cB * 4 + 3) * 32], yB,
                                   Q.hd0[fB] - 2.f * alt);
        // chunks are ordered by frequency: close a frequency when the next chunk belongs to another one
        if (fA != cur_f) {
            close_frequency(acc, pred, F, cur_f, lane);
            acc = c2{S(0.f), S(0.f)};
            cur_f = fA;
        }
        acc = acc + tA.term;
        if (hasB) {
            if (fB != cur_f) {
                close_frequency(acc, pred, F, cur_f, lane);
                acc = c2{S(0.f), S(0.f)};
                cur_f = fB;
            }
            acc = acc + tB.term;
        }
    }
    close_frequency(acc, pred, F, cur_f, lane);
    __syncwarp();
}

// ---------------------------------------------------------------- forward + Jacobian: one chunk per pass
__device__ __noinline__ void fdem_sens_f2(const SysShared<float>& Q, const float* __restrict__ tab, float alt, int L,
                                          const float* __restrict__ msig, const float* __restrict__ mthk,
                                          float* __restrict__ pred, float* __restrict__ J, const int c0, const int c1)
{
    using namespace f2;
    __builtin_assume(__isShared(&Q));
    __builtin_assume(__isShared(tab));
    __builtin_assume(__isShared(msig));
    __builtin_assume(__isShared(mthk));
    __builtin_assume(__isShared(pred));
    const int lane = threadIdx.x & 31;
    const int F = Q.n_freq, NCH = c1;
    const float4* T4 = reinterpret_cast<const float4*>(tab) + lane;

    // thread-local scratch of the chain-rule pass, one float2 (two abscissae) per layer:
    // D_k = dy_k/dy_{k+1}, l_k = dy_k/dln(sigma_k); jr/ji accumulate the Jacobian of the current frequency
    v2 Dr[KS], Di[KS], lr[KS], li[KS];
    v2 jr[KS], ji[KS];

    int cur_f = -1;
    c2 acc = {S(0.f), S(0.f)};
#pragma unroll 1
    for (int c = c0; c <= NCH; ++c) {
        const int f = c < NCH ? (int)Q.chunk_freq[c] : -2;
        if (f != cur_f) {
            if (cur_f >= 0) {  // close frequency cur_f: 8 warp sums per folded reduction (same bits as warp_sum), the
                // response and the first three layers in the first one, then four layers at a time; value i = 2 j + part
                // (j = 0: response, j >= 1: layer k0 + j - 1) lands in lane vl, which stores it
                const int vi = ((lane >> 4) & 1) | ((lane >> 2) & 2) | (lane & 4), part = vi & 1, slot = vi >> 1;
                const bool writer = (lane & 3) == 0;
                {
                    const v2 z = S(0.f);
                    const v2 a0 = L > 0 ? jr[0] : z, b0 = L > 0 ? ji[0] : z, a1 = L > 1 ? jr[1] : z, b1 = L > 1 ? ji[1] : z;
                    const v2 a2 = L > 2 ? jr[2] : z, b2 = L > 2 ? ji[2] : z;
                    const float v = warp_sum8(acc.re.x + acc.re.y, acc.im.x + acc.im.y, a0.x + a0.y, b0.x + b0.y, a1.x + a1.y,
                                              b1.x + b1.y, a2.x + a2.y, b2.x + b2.y, lane);
                    if (writer) {
                        if (slot == 0) pred[part * F + cur_f] = v;
                        else if (slot <= L) J[(part * F + cur_f) * KS + slot - 1] = v;
                    }
                }
#pragma unroll 1
                for (int k0 = 3; k0 < L; k0 += 4) {
                    const v2 z = S(0.f);
                    const v2 a0 = jr[k0], b0 = ji[k0], a1 = k0 + 1 < L ? jr[k0 + 1] : z, b1 = k0 + 1 < L ? ji[k0 + 1] : z;
                    const v2 a2 = k0 + 2 < L ? jr[k0 + 2] : z, b2 = k0 + 2 < L ? ji[k0 + 2] : z;
                    const v2 a3 = k0 + 3 < L ? jr[k0 + 3] : z, b3 = k0 + 3 < L ? ji[k0 + 3] : z;
                    const float v = warp_sum8(a0.x + a0.y, b0.x + b0.y, a1.x + a1.y, b1.x + b1.y, a2.x + a2.y, b2.x + b2.y,
                                              a3.x + a3.y, b3.x + b3.y, lane);
                    if (writer && k0 + slot < L) J[(part * F + cur_f) * KS + k0 + slot] = v;
                }
            }
            if (c >= NCH) break;
            cur_f = f;
            acc = c2{S(0.f), S(0.f)};
#pragma unroll 1
            for (int k = 0; k < L; ++k) {
                jr[k] = S(0.f);
                ji[k] = S(0.f);
            }
        }
        const float omu = Q.omu[f];
        const float4 q0 = T4[(c * 4 + 0) * 32];
        const v2 lam = V(q0.x, q0.y);
        const v2 a = fma(lam, lam, S(Q.k2re[f]));  // Re(u^2) of every earth layer
        // basement: y_L = u_L, dy_L/dln(sigma_L) = i b / (2 u_L)
        float b = omu * msig[L - 1];
        v2 m;
        c2 u = csqrt_q1(a, b, m);
        c2 y = u;
        {
            const v2 hb = mul(S(0.5f * b), rcp(m));    // (b/2) / |u|^2 ;  i (b/2) conj(u) / |u|^2
            lr[L - 1] = mul(hb, u.im);
            li[L - 1] = mul(hb, u.re);
        }
#pragma unroll 1
        for (int k = L - 2; k >= 0; --k) {
            b = omu * msig[k];
            const float t = mthk[k];
            u = csqrt_q1(a, b, m);
            const c2 e = cexp(c2{mul(S(-2.f * t), u.re), mul(S(-2.f * t), u.im)});
            const c2 d = y - u, s = y + u;
            const c2 p = c2{add(e.re, S(-1.f)), e.im} * d;                                   // (e - 1)(y - u)
            const c2 den = {fma(S(2.f), u.re, neg(p.re)), fma(S(2.f), u.im, neg(p.im))};
            const c2 num = {fma(S(2.f), y.re, p.re), fma(S(2.f), y.im, p.im)};
            const c2 inv = cinv(den);
            const c2 inv2 = csq(inv);
            const c2 ue = u * e;
            const c2 dk = ((u * ue) * inv2) * S(4.f);                                        // 4 u^2 e / den^2
            Dr[k] = dk.re;
            Di[k] = dk.im;
            // N = s^2 - 4 u e d (1 + t s) + e^2 (4 u^2 - d^2)
            const c2 d2 = csq(d);
            const c2 g = (ue * d) * c2{fma(S(t), s.re, S(1.f)), mul(S(t), s.im)};
            const c2 h = csq(e) * c2{fma(S(4.f), a, neg(d2.re)), add(S(4.f * b), neg(d2.im))};
            const c2 s2 = csq(s);
            const c2 N = {add(fma(S(-4.f), g.re, s2.re), h.re), add(fma(S(-4.f), g.im, s2.im), h.im)};
            // q = N / (u den^2) = N inv2 conj(u) / |u|^2 ;  l = (i b / 2) q
            const c2 q = (N * inv2) * c2{u.re, neg(u.im)};
            const v2 hb = mul(S(0.5f * b), rcp(m));
            lr[k] = neg(mul(hb, q.im));
            li[k] = mul(hb, q.re);
            y = (u * num) * inv;
        }
        const TopOut tp = top_term(q0, T4[(c * 4 + 1) * 32], T4[(c * 4 + 2) * 32], T4[(c * 4 + 3) * 32], y,
                                   Q.hd0[f] - 2.f * alt);
        acc = acc + tp.term;
        c2 P = ((tp.u0 * csq(tp.is)) * S(-2.f)) * tp.K;  // d rTE/dy1 * K
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            const c2 v = P * c2{lr[k], li[k]};
            jr[k] = add(jr[k], v.re);
            ji[k] = add(ji[k], v.im);
            if (k < L - 1) P = P * c2{Dr[k], Di[k]};
        }
    }
    __syncwarp();
}

// dispatch on the arithmetic type: fp32 -> the packed bodies above, fp64 -> the generic template
template <typename T>
__device__ __forceinline__ void fdem_run(const SysShared<T>& Q, const T* __restrict__ tab, T alt, int L,
                                         const T* __restrict__ msig, const T* __restrict__ mthk, T* __restrict__ pred,
                                         T* __restrict__ J, const bool sens)
{
    if constexpr (sizeof(T) == 4) {
        if (sens) fdem_sens_f2(Q, tab, alt, L, msig, mthk, pred, J, 0, Q.n_chunks);
        else fdem_fwd_f2(Q, tab, alt, L, msig, mthk, pred, 0, Q.n_chunks);
    } else {
        fdem_eval<T>(Q, tab, alt, L, msig, mthk, pred, J, sens);
    }
}

}  // namespace gbp
