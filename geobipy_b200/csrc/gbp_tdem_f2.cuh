// fp32 production instantiation of tdem_eval(): TWO Hankel abscissae per lane, packed fp32x2 arithmetic
// (FFMA2 / FMUL2 / FADD2 on sm_100), exactly as gbp_fdem_f2.cuh does for the frequency-domain forward.  Each lane
// carries the difference-admittance recursion of abscissae j and j + 1 of its own frequency in the two halves of
// float2 registers; the number of abscissae is even by construction (gbp_tdem_tables.h).  Same mathematics as the
// generic template in gbp_tdem.cuh (the fp64 validation path); results differ only by summation order.
#pragma once
#include "gbp_fdem_f2.cuh"
#include "gbp_tdem.cuh"

namespace gbp {

template <>
__device__ __noinline__ void tdem_eval<float>(const TdShared<float>& Q, const float* __restrict__ Mt,
                                              const float* __restrict__ lam, const float* __restrict__ wgt, int L,
                                              const float* __restrict__ msig, const float* __restrict__ mthk,
                                              float* __restrict__ sbuf, float* __restrict__ pred, float* __restrict__ J,
                                              const bool sens, const signed char* ccomp, const int comp)
{
    using namespace f2;
    __builtin_assume(__isShared(&Q));
    __builtin_assume(__isShared(Mt));
    __builtin_assume(__isShared(lam));
    __builtin_assume(__isShared(wgt));
    __builtin_assume(__isShared(msig));
    __builtin_assume(__isShared(mthk));
    __builtin_assume(__isShared(sbuf));
    __builtin_assume(__isShared(pred));
    const int lane = threadIdx.x & 31;
    const int C = Q.C, NL = Q.n_lam;
    const float omu = Q.omu[lane];

    v2 Ddr[KS], Ddi[KS], Gr[KS], Gi[KS];
    v2 jr[KS], ji[KS];
    if (sens) {
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            jr[k] = S(0.f);
            ji[k] = S(0.f);
        }
    }
    c2 acc = {S(0.f), S(0.f)};
    const v2 one = S(1.f), zero = S(0.f);
#pragma unroll 1
    for (int j = 0; j < NL; j += 2) {
        const v2 l = V(lam[j], lam[j + 1]);
        const v2 w = V(wgt[j], wgt[j + 1]);
        const v2 a2 = mul(l, l);
        // basement: D_L = i b / (u + lambda)
        float b = omu * msig[L - 1];
        c2 u = csqrt_q1(a2, S(b));
        c2 ib = {zero, S(b)};
        c2 D = ib * cinv(c2{add(u.re, l), u.im});
        if (sens) {  // G_L = i b / (2u)
            const c2 g = (ib * cinv(u)) * S(0.5f);
            Gr[L - 1] = g.re;
            Gi[L - 1] = g.im;
        }
#pragma unroll 1
        for (int k = L - 2; k >= 0; --k) {
            b = omu * msig[k];
            const float two_h = 2.f * mthk[k];
            u = csqrt_q1(a2, S(b));
            ib = c2{zero, S(b)};
            const c2 a = ib * cinv(c2{add(u.re, l), u.im});  // u - lambda
            // e = exp(-2 u h), clamped as in the generic template
            const v2 lim = mul(S(60.f), rcp(u.re));
            const v2 sc = V(fminf(two_h, lim.x), fminf(two_h, lim.y));
            c2 e = cexp(c2{neg(mul(sc, u.re)), neg(mul(sc, u.im))});
            const v2 ze = V(two_h * u.re.x > 60.f ? 0.f : 1.f, two_h * u.re.y > 60.f ? 0.f : 1.f);
            e = e * ze;
            const c2 E = D - a;                                // Y_{k+1} - u_k
            const c2 Y = {add(l, D.re), D.im};
            const c2 ope = {add(one, e.re), e.im};             // 1 + e
            const c2 q = ope * u + c2{sub(one, e.re), neg(e.im)} * Y;
            const c2 iq = cinv(q);
            const c2 eu = e * u;
            if (sens) {
                const c2 iq2 = iq * iq;
                const c2 dd = ((eu * u) * iq2) * S(4.f);      // 4 e u^2 / q^2
                Ddr[k] = dd.re;
                Ddi[k] = dd.im;
                // dD_k/du = 1 + 2e(-2h u E + E - u)/q - 2 e u E ((1+e) + 2 h e E)/q^2
                const c2 t1 = ((e * ((u * E) * S(-two_h) + E - u)) * iq) * S(2.f);
                const c2 t2 = (((eu * E) * (ope + (e * E) * S(two_h))) * iq2) * S(2.f);
                const c2 dDdu = {add(one, sub(t1.re, t2.re)), sub(t1.im, t2.im)};
                const c2 g = dDdu * ((ib * cinv(u)) * S(0.5f));
                Gr[k] = g.re;
                Gi[k] = g.im;
            }
            D = a + ((eu * E) * iq) * S(2.f);
        }
        const c2 iden = cinv(c2{fma(S(2.f), l, D.re), D.im});
        const c2 rte = D * iden;  // = -rTE
        acc = acc - rte * w;
        if (sens) {
            c2 P = (iden * iden) * mul(mul(S(-2.f), l), w);  // w * d rTE / d D_1
#pragma unroll 1
            for (int k = 0; k < L; ++k) {
                const c2 v = P * c2{Gr[k], Gi[k]};
                jr[k] = add(jr[k], v.re);
                ji[k] = add(ji[k], v.im);
                if (k < L - 1) P = P * c2{Ddr[k], Ddi[k]};
            }
        }
    }
    // windows: lanes own channels
    __syncwarp();
    sbuf[lane] = acc.re.x + acc.re.y;
    sbuf[TD_NF + lane] = acc.im.x + acc.im.y;
    __syncwarp();
#pragma unroll 1
    for (int c = lane; c < C; c += 32) {
        if (ccomp != nullptr && ccomp[c] != comp) continue;
        float d = 0.f;
#pragma unroll 8
        for (int i = 0; i < TD_ROWS; ++i) d = fmaf(Mt[i * TD_CP + c], sbuf[i], d);
        pred[c] = d;
    }
    if (sens) {
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            __syncwarp();
            sbuf[lane] = jr[k].x + jr[k].y;
            sbuf[TD_NF + lane] = ji[k].x + ji[k].y;
            __syncwarp();
#pragma unroll 1
            for (int c = lane; c < C; c += 32) {
                if (ccomp != nullptr && ccomp[c] != comp) continue;
                float d = 0.f;
#pragma unroll 8
                for (int i = 0; i < TD_ROWS; ++i) d = fmaf(Mt[i * TD_CP + c], sbuf[i], d);
                J[c * KS + k] = d;
            }
        }
    }
    __syncwarp();
}

}  // namespace gbp
