// fp32 production instantiation of fdem_eval(): TWO filter abscissae per lane, packed fp32x2 arithmetic.
//
// Blackwell (sm_100) has two-wide fp32 instructions (FFMA2 / FMUL2 / FADD2 on 64-bit register pairs,
// PTX fma.rn.f32x2 ...).  Each lane carries the admittance recursion of abscissae j and j+32 in the two
// halves of float2 registers: the instruction count of the complex arithmetic halves, and the two
// independent recursions give every warp instruction-level parallelism (the sampler is bound by
// dependent-issue latency, not by any pipe: ncu profiles/).  Special functions (MUFU rcp/sqrt/ex2/sin/cos)
// stay scalar, two per pair.  Same mathematics as the generic template in gbp_fdem.cuh (which remains the
// fp64 validation path); results differ only by summation order.
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

__device__ __forceinline__ c2 operator+(c2 a, c2 b) { return {add(a.re, b.re), add(a.im, b.im)}; }
__device__ __forceinline__ c2 operator-(c2 a, c2 b) { return {sub(a.re, b.re), sub(a.im, b.im)}; }
__device__ __forceinline__ c2 operator*(c2 a, c2 b)
{
    return {fma(a.re, b.re, neg(mul(a.im, b.im))), fma(a.re, b.im, mul(a.im, b.re))};
}
__device__ __forceinline__ c2 operator*(c2 a, v2 s) { return {mul(a.re, s), mul(a.im, s)}; }
__device__ __forceinline__ c2 cinv(c2 a)
{
    const v2 d = rcp(fma(a.re, a.re, mul(a.im, a.im)));
    return {mul(a.re, d), neg(mul(a.im, d))};
}
// sqrt of a + ib with b >= 0 (first-quadrant result), both halves
__device__ __forceinline__ c2 csqrt_q1(v2 a, v2 b)
{
    const v2 m = sqrt(fma(a, a, mul(b, b)));
    const v2 t = sqrt(mul(S(0.5f), add(m, V(fabsf(a.x), fabsf(a.y)))));
    const v2 o = mul(b, rcp(add(t, t)));
    c2 r;
    r.re = V(a.x >= 0.f ? t.x : o.x, a.y >= 0.f ? t.y : o.y);
    r.im = V(a.x >= 0.f ? o.x : t.x, a.y >= 0.f ? o.y : t.y);
    return r;
}
// exp(z), |Im z| <= ~64
__device__ __forceinline__ c2 cexp(c2 z)
{
    const v2 l = mul(z.re, S(1.4426950408889634f));
    v2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(l.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(l.y));
    const v2 q = mul(z.im, S(0.15915494309189535f));
    const v2 n = V(rintf(q.x), rintf(q.y));
    v2 r = fma(n, S(-6.2831854820251465f), z.im);
    r = fma(n, S(1.7484556000744883e-07f), r);
    v2 s, c;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s.x) : "f"(r.x));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s.y) : "f"(r.y));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c.x) : "f"(r.x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c.y) : "f"(r.y));
    return {mul(e, c), mul(e, s)};
}

}  // namespace f2

template <>
__device__ __noinline__ void fdem_eval<float>(const SysShared<float>& Q, const float* __restrict__ tab, float alt, int L,
                                              const float* __restrict__ msig, const float* __restrict__ mthk,
                                              float* __restrict__ pred, float* __restrict__ J, const bool sens)
{
    using namespace f2;
    __builtin_assume(__isShared(&Q));
    __builtin_assume(__isShared(tab));
    __builtin_assume(__isShared(msig));
    __builtin_assume(__isShared(mthk));
    __builtin_assume(__isShared(pred));
    const int lane = threadIdx.x & 31;
    const int F = Q.n_freq;
    const int ts = Q.tab_stride;
    const float* t_lam = tab;
    const float* t_u0r = tab + ts;
    const float* t_u0i = tab + 2 * ts;
    const float* t_er = tab + 3 * ts;
    const float* t_ei = tab + 4 * ts;
    const float* t_cr = tab + 5 * ts;
    const float* t_ci = tab + 6 * ts;

    // thread-local scratch of the chain-rule pass (sens only), one float2 (two abscissae) per layer
    v2 Dr[KS], Di[KS], lr[KS], li[KS];
    v2 jr[KS], ji[KS];

    int seg = 0;
#pragma unroll 1
    for (int f = 0; f < F; ++f) {
        const float omu = Q.omu[f];
        const float k2 = Q.k2re[f];
        const float hd = Q.hd0[f] - 2.f * alt;
        c2 acc = {S(0.f), S(0.f)};
        if (sens) {
#pragma unroll 1
            for (int k = 0; k < L; ++k) {
                jr[k] = S(0.f);
                ji[k] = S(0.f);
            }
        }
#pragma unroll 1
        for (; seg < Q.n_seg && Q.seg[seg].freq == f; ++seg) {
            const int s0 = Q.seg[seg].start, cnt = Q.seg[seg].count;
#pragma unroll 1
            for (int j = lane; j < cnt; j += 64) {
                // abscissae j and j + 32 (an out-of-range partner is computed on a valid index with weight 0)
                const int jb = j + 32;
                const bool vb = jb < cnt;
                const int ia = s0 + j, ib = s0 + (vb ? jb : j);
                const v2 lam = V(t_lam[ia], t_lam[ib]);
                const v2 a = fma(lam, lam, S(k2));  // Re(u^2) of every earth layer
                // basement: y_L = u_L
                float b = omu * msig[L - 1];
                c2 u = csqrt_q1(a, S(b));
                c2 y = u;
                if (sens) {  // i*b/(2u)
                    const c2 iu = cinv(u);
                    lr[L - 1] = mul(S(-0.5f * b), iu.im);
                    li[L - 1] = mul(S(0.5f * b), iu.re);
                }
#pragma unroll 1
                for (int k = L - 2; k >= 0; --k) {
                    b = omu * msig[k];
                    const float t = mthk[k];
                    u = csqrt_q1(a, S(b));
                    // tanh(u t) = (1 - e)/(1 + e), e = exp(-2ut); clamp as in the generic template
                    const float two_t = 2.f * t;
                    const v2 lim = mul(S(60.f), rcp(u.re));
                    const v2 sc = V(fminf(two_t, lim.x), fminf(two_t, lim.y));
                    c2 e = cexp(c2{neg(mul(sc, u.re)), neg(mul(sc, u.im))});
                    const v2 ze = V(two_t * u.re.x > 60.f ? 0.f : 1.f, two_t * u.re.y > 60.f ? 0.f : 1.f);
                    e = e * ze;
                    const c2 th = c2{sub(S(1.f), e.re), neg(e.im)} * cinv(c2{add(S(1.f), e.re), e.im});
                    const c2 den = u + y * th;
                    const c2 num = y + u * th;
                    const c2 inv = cinv(den);
                    if (sens) {
                        const c2 u2 = {a, S(b)};
                        const c2 th2 = th * th;
                        const c2 inv2 = inv * inv;
                        const c2 w = y * y - u2;                               // y^2 - u^2
                        const c2 one_m = {sub(S(1.f), th2.re), neg(th2.im)};   // 1 - tanh^2
                        const c2 d = u2 * one_m * inv2;                        // accumulate[] of M1_1
                        Dr[k] = d.re;
                        Di[k] = d.im;
                        // B = 2uy th^2 + (y^2-u^2) th + 2u^2 - t u (y^2-u^2)(1 - th^2)
                        const c2 uy = u * y;
                        const c2 B = (uy * th2) * S(2.f) + w * th + u2 * S(2.f) - ((u * w) * one_m) * S(t);
                        const c2 q = B * inv2 * cinv(u);                       // B / (u den^2)
                        lr[k] = mul(S(-0.5f * b), q.im);                        // * i*b/2
                        li[k] = mul(S(0.5f * b), q.re);
                    }
                    y = u * num * inv;
                }
                const c2 u0 = {V(t_u0r[ia], t_u0r[ib]), V(t_u0i[ia], t_u0i[ib])};
                const c2 is = cinv(u0 + y);
                const c2 rte = (u0 - y) * is;
                const c2 cw = {V(t_cr[ia], vb ? t_cr[ib] : 0.f), V(t_ci[ia], vb ? t_ci[ib] : 0.f)};
                const c2 K = cw * cexp(c2{mul(V(t_er[ia], t_er[ib]), S(hd)), mul(V(t_ei[ia], t_ei[ib]), S(hd))});
                acc = acc + rte * K;
                if (sens) {
                    c2 P = (u0 * is * is) * S(-2.f) * K;  // d rTE/dy1 * K
#pragma unroll 1
                    for (int k = 0; k < L; ++k) {
                        const c2 v = P * c2{lr[k], li[k]};
                        jr[k] = add(jr[k], v.re);
                        ji[k] = add(ji[k], v.im);
                        if (k < L - 1) P = P * c2{Dr[k], Di[k]};
                    }
                }
            }
        }
        const float sr = warp_sum(acc.re.x + acc.re.y), si = warp_sum(acc.im.x + acc.im.y);
        if (lane == 0) {
            pred[f] = sr;
            pred[F + f] = si;
        }
        if (sens) {
#pragma unroll 1
            for (int k = 0; k < L; ++k) {
                const float a = warp_sum(jr[k].x + jr[k].y), b = warp_sum(ji[k].x + ji[k].y);
                if (lane == 0) {
                    J[f * KS + k] = a;
                    J[(F + f) * KS + k] = b;
                }
            }
        }
    }
    __syncwarp();
}

}  // namespace gbp
