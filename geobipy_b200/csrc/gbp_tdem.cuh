// Warp-cooperative airborne time-domain EM forward + analytic Jacobian (one warp = one sounding).
//
// Replaces TdemDataPoint.forward / .sensitivity / .fm_dlogc (geobipy/src/classes/data/datapoint/
// TdemDataPoint.py:997-1055) and what they call: gaTdem1dfwd / ga_fm_dlogc / gaTdem1dsen
// (classes/forwardmodelling/Electromagnetic/TD/tdem1d.py:89-154), i.e. the external gatdaem1d
// (GeoscienceAustralia/ga-aem) forward model for a SkyTEM-type system: layered-earth frequency-domain
// response at log-spaced frequencies -> spline in log-frequency -> waveform spectrum and receiver
// filters -> time domain -> window averages.
//
// B200 design (not a translation of gatdaem1d):
//  * lane = spline-node frequency (GBP_TD_NFREQ = 32 nodes shared by the systems of a dual-moment
//    datapoint: both see the same transmitter/receiver geometry).  Each lane runs the Hankel integral of
//    its own frequency sequentially over the n_lam log-spaced abscissae (trapezoid rule), so the
//    frequency-domain response S_i needs no warp reduction at all;
//  * the admittance recursion is written on the difference D_k = Y_k - lambda, which has no cancellation
//    at low induction numbers - that is what makes the fp32 instantiation accurate at late times;
//  * everything after S_i (spline, waveform Fourier series, low-pass filters, window averages, the sign
//    convention) is linear in S and is folded on the host, in fp64, into ONE matrix per datapoint type
//    (gbp_tdem_tables.h), staged to shared memory with a TMA bulk copy:  d_c = sum_i Mt[i][c] s_i with
//    s = [Re S_0..31, Im S_0..31].  Lanes then own output channels (conflict-free column reads);
//  * the per-sounding abscissae / geometry weights (height dependent) are computed once per chain.
#pragma once
#include "gbp_fdem.cuh"

namespace gbp {

constexpr int TD_NF = GBP_TD_NFREQ;      // spline nodes = lanes
constexpr int TD_CP = GBP_TD_MAXC;       // padded channel stride of Mt
constexpr int TD_ROWS = 2 * TD_NF;       // rows of Mt: Re S_i then Im S_i

// Per-datapoint-type constants, passed by value as a kernel parameter.
struct TdDev {
    int n_sys, n_lam, C, pad;
    int n_win[GBP_TD_MAXSYS];
    double omu[TD_NF];            // omega_i * mu0
    double xi[GBP_TD_MAXLAM];     // ln(lambda ZH / 2)
    double tw[GBP_TD_MAXLAM];     // trapezoid weights * d(xi) * mu0 / (4 pi)
    double rx_r, rx_dz, loop_radius;
    double tsc[GBP_TD_MAXC];      // sqrt(1 ms / t_c): additive-error scaling of channel c (TdemDataPoint.py:369)
    int csys[GBP_TD_MAXC];        // system of channel c
    int ccomp[GBP_TD_MAXC];       // component of channel c: 0 = z, 1 = x
    int has_x, tempest;           // some channel measures the x component; the Tempest datapoint's error model (sampler)
    double rx_cx;                 // dx / r: direction cosine of the receiver offset
    double poff[GBP_TD_MAXC];     // Tempest: predicted primary field of channel c's component (added to the response)
};

template <typename T> struct TdShared {
    int n_lam, C;
    T omu[TD_NF];
    double xi[GBP_TD_MAXLAM], tw[GBP_TD_MAXLAM];   // td_geometry (per sounding; per proposal when the height is sampled)
    double rx_r, rx_dz, loop_radius, rx_cx;
    signed char ccomp[GBP_TD_MAXC];
    T poff[GBP_TD_MAXC];
};
template <typename T> __device__ __forceinline__ void fill_td_shared(const TdDev& S, TdShared<T>& q)
{
    q.rx_cx = S.rx_cx;
    for (int i = 0; i < GBP_TD_MAXC; ++i) {
        q.ccomp[i] = (signed char)S.ccomp[i];
        q.poff[i] = (T)S.poff[i];
    }
    q.n_lam = S.n_lam;
    q.C = S.C;
    for (int i = 0; i < TD_NF; ++i) q.omu[i] = (T)S.omu[i];
    for (int i = 0; i < GBP_TD_MAXLAM; ++i) {
        q.xi[i] = S.xi[i];
        q.tw[i] = S.tw[i];
    }
    q.rx_r = S.rx_r;
    q.rx_dz = S.rx_dz;
    q.loop_radius = S.loop_radius;
}

// Per-sounding Hankel abscissae and geometry weights (one warp; lane j = abscissa j), fp64 then cast:
//   lambda_j = (2/ZH) exp(xi_j),  w_j = tw_j lambda_j^3 exp(-lambda_j ZH) J0(lambda_j r) 2 J1(lambda_j a)/(lambda_j a)
// xcomp: the weights of the horizontal (x) secondary field instead - J1(lambda_j r) dx / r in place of J0(lambda_j r)
template <typename T>
__device__ __noinline__ void td_geometry(const TdShared<T>& S, double altitude, T* lam, T* wgt, const bool xcomp = false)
{
    __builtin_assume(__isShared(&S));
    const int lane = threadIdx.x & 31;
    if (lane < S.n_lam) {
        const double ZH = 2.0 * altitude + S.rx_dz;
        const double l = (2.0 / ZH) * ::exp(S.xi[lane]);
        double w = S.tw[lane] * l * l * l * ::exp(-l * ZH) * (xcomp ? ::j1(l * S.rx_r) * S.rx_cx : ::j0(l * S.rx_r));
        if (S.loop_radius > 0.0) {
            const double x = l * S.loop_radius;
            w *= 2.0 * ::j1(x) / x;
        }
        lam[lane] = (T)l;
        wgt[lane] = (T)w;
    }
    __syncwarp();
}

// Q, Mt, lam, wgt, msig, mthk, sbuf, pred, J live in SHARED memory.  Mt: TD_ROWS rows of TD_CP values.
// sbuf: TD_ROWS values of scratch.  pred: [C].  J: [C][KS] = d pred / d ln(sigma), written only if sens.
// All lanes of the warp must call.
template <typename T>
__device__ __noinline__ void tdem_eval(const TdShared<T>& Q, const T* __restrict__ Mt, const T* __restrict__ lam,
                                       const T* __restrict__ wgt, int L, const T* __restrict__ msig,
                                       const T* __restrict__ mthk, T* __restrict__ sbuf, T* __restrict__ pred,
                                       T* __restrict__ J, const bool sens, const signed char* ccomp = nullptr, const int comp = 0)
{   // ccomp != nullptr: only the channels c with ccomp[c] == comp are written (one call per measured component)
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
    const T omu = Q.omu[lane];

    // thread-local scratch of the chain-rule pass (sens only):
    //   Dd_k = dD_k/dD_{k+1},  G_k = dD_k/d ln(sigma_k);  jr/ji accumulate dS/d ln(sigma_k) over abscissae
    T Ddr[KS], Ddi[KS], Gr[KS], Gi[KS];
    T jr[KS], ji[KS];
    if (sens) {
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            jr[k] = T(0);
            ji[k] = T(0);
        }
    }
    cx<T> acc = {T(0), T(0)};
#pragma unroll 1
    for (int j = 0; j < NL; ++j) {
        const T l = lam[j];
        const T w = wgt[j];
        const T a2 = l * l;
        // basement: D_L = u - lambda = i b / (u + lambda)
        T b = omu * msig[L - 1];
        cx<T> u = csqrt_q1<T>(a2, b);
        cx<T> ib = mk<T>(T(0), b);
        cx<T> D = ib * cinv(mk<T>(u.re + l, u.im));
        if (sens) {  // dD_L/du = 1  ->  G_L = i b / (2u)
            const cx<T> g = ib * cinv(u) * T(0.5);
            Gr[L - 1] = g.re;
            Gi[L - 1] = g.im;
        }
#pragma unroll 1
        for (int k = L - 2; k >= 0; --k) {
            b = omu * msig[k];
            const T h = mthk[k];
            u = csqrt_q1<T>(a2, b);
            ib = mk<T>(T(0), b);
            const cx<T> a = ib * cinv(mk<T>(u.re + l, u.im));   // u - lambda
            // e = exp(-2 u h); beyond Re(2uh) = 60 it is below round-off of 1: clamp the argument (keeps the
            // sin/cos argument small) and set e to exactly 0 there
            const T two_h = T(2) * h;
            const T sc = fmin(two_h, T(60) * rt<T>::rcp(u.re));
            cx<T> e = cexp_<T>(mk<T>(-sc * u.re, -sc * u.im));
            if (two_h * u.re > T(60)) e = mk<T>(T(0), T(0));
            const cx<T> E = D - a;                               // Y_{k+1} - u_k
            const cx<T> Y = mk<T>(l + D.re, D.im);
            const cx<T> q = mk<T>(T(1) + e.re, e.im) * u + mk<T>(T(1) - e.re, -e.im) * Y;
            const cx<T> iq = cinv(q);
            const cx<T> eu = e * u;
            if (sens) {
                const cx<T> iq2 = iq * iq;
                const cx<T> dd = (eu * u) * iq2 * T(4);          // 4 e u^2 / q^2
                Ddr[k] = dd.re;
                Ddi[k] = dd.im;
                // dD_k/du = 1 + 2e(-2h u E + E - u)/q - 2 e u E ((1+e) + 2 h e E)/q^2
                const cx<T> t1 = (e * ((u * E) * (-two_h) + E - u)) * iq * T(2);
                const cx<T> t2 = (eu * E) * (mk<T>(T(1) + e.re, e.im) + (e * E) * two_h) * iq2 * T(2);
                const cx<T> dDdu = mk<T>(T(1) + t1.re - t2.re, t1.im - t2.im);
                const cx<T> g = dDdu * (ib * cinv(u) * T(0.5));  // * d u / d ln(sigma) = i b / (2u)
                Gr[k] = g.re;
                Gi[k] = g.im;
            }
            D = a + (eu * E) * iq * T(2);
        }
        const cx<T> iden = cinv(mk<T>(T(2) * l + D.re, D.im));
        const cx<T> rte = D * iden;   // = -rTE
        acc = acc - rte * w;
        if (sens) {
            cx<T> P = (iden * iden) * (T(-2) * l * w);  // w * d rTE / d D_1
#pragma unroll 1
            for (int k = 0; k < L; ++k) {
                const cx<T> v = P * mk<T>(Gr[k], Gi[k]);
                jr[k] += v.re;
                ji[k] += v.im;
                if (k < L - 1) P = P * mk<T>(Ddr[k], Ddi[k]);
            }
        }
    }
    // windows: lanes own channels
    __syncwarp();
    sbuf[lane] = acc.re;
    sbuf[TD_NF + lane] = acc.im;
    __syncwarp();
#pragma unroll 1
    for (int c = lane; c < C; c += 32) {
        if (ccomp != nullptr && ccomp[c] != comp) continue;
        T d = T(0);
#pragma unroll 8
        for (int i = 0; i < TD_ROWS; ++i) d += Mt[i * TD_CP + c] * sbuf[i];
        pred[c] = d;
    }
    if (sens) {
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            __syncwarp();
            sbuf[lane] = jr[k];
            sbuf[TD_NF + lane] = ji[k];
            __syncwarp();
#pragma unroll 1
            for (int c = lane; c < C; c += 32) {
                if (ccomp != nullptr && ccomp[c] != comp) continue;
                T d = T(0);
#pragma unroll 8
                for (int i = 0; i < TD_ROWS; ++i) d += Mt[i * TD_CP + c] * sbuf[i];
                J[c * KS + k] = d;
            }
        }
    }
    __syncwarp();
}

// Compact single-precision Bessel functions J0, J1 (rational / asymptotic forms of Abramowitz & Stegun 9.4, the
// classic bessj0 / bessj1 coefficients; ~1e-7 absolute in fp32).  Small on purpose: the per-step geometry update below
// must not push the sampler's hot code out of the instruction cache (libdevice's j0f / j1f did: +16 % per step).
__device__ __forceinline__ float bessel_j01_f32(const float x, const bool one)
{
    const float ax = fabsf(x);
    if (ax < 8.f) {
        const float y = x * x;
        float n, d;
        if (!one) {
            n = 57568490574.0f + y * (-13362590354.0f + y * (651619640.7f + y * (-11214424.18f + y * (77392.33017f + y * (-184.9052456f)))));
            d = 57568490411.0f + y * (1029532985.0f + y * (9494680.718f + y * (59272.64853f + y * (267.8532712f + y))));
            return n / d;
        }
        n = x * (72362614232.0f + y * (-7895059235.0f + y * (242396853.1f + y * (-2972611.439f + y * (15704.48260f + y * (-30.16036606f))))));
        d = 144725228442.0f + y * (2300535178.0f + y * (18583304.74f + y * (99447.43394f + y * (376.9991397f + y))));
        return n / d;
    }
    const float z = 8.f / ax, y = z * z;
    const float xx = ax - (one ? 2.356194491f : 0.785398164f);
    float p, q;
    if (!one) {
        p = 1.f + y * (-0.1098628627e-2f + y * (0.2734510407e-4f + y * (-0.2073370639e-5f + y * 0.2093887211e-6f)));
        q = -0.1562499995e-1f + y * (0.1430488765e-3f + y * (-0.6911147651e-5f + y * (0.7621095161e-6f - y * 0.934935152e-7f)));
    } else {
        p = 1.f + y * (0.183105e-2f + y * (-0.3516396496e-4f + y * (0.2457520174e-5f + y * (-0.240337019e-6f))));
        q = 0.04687499995f + y * (-0.2002690873e-3f + y * (0.8449199096e-5f + y * (-0.88228987e-6f + y * 0.105787412e-6f)));
    }
    float sn, cs;
    rt<float>::sincos(xx, &sn, &cs);
    const float r = rt<float>::sqrt(0.636619772f / ax) * (cs * p - z * sn * q);
    return (one && x < 0.f) ? -r : r;
}

// td_geometry in single precision (fp32 sampler with a sampled transmitter height: once per accept_reject step;
// weights to ~1e-6 relative, the forward's fp32 tolerance is 2e-4).  lambda ZH = 2 exp(xi) does not depend on the height.
__device__ __noinline__ void td_geometry_f32(const TdShared<float>& S, float altitude, float* lam, float* wgt)
{
    __builtin_assume(__isShared(&S));
    const int lane = threadIdx.x & 31;
    if (lane < S.n_lam) {
        const float ZH = 2.f * altitude + (float)S.rx_dz;
        const float lz = 2.f * rt<float>::exp((float)S.xi[lane]);   // lambda * ZH
        const float l = lz / ZH;
        float w = (float)S.tw[lane] * l * l * l * rt<float>::exp(-lz) * bessel_j01_f32(l * (float)S.rx_r, false);
        if (S.loop_radius > 0.0) {
            const float x = l * (float)S.loop_radius;
            w *= 2.f * bessel_j01_f32(x, true) / x;
        }
        lam[lane] = l;
        wgt[lane] = w;
    }
    __syncwarp();
}

// standalone operators: one warp per sounding, grid-stride over soundings
template <typename T, bool SENS>
__global__ void __launch_bounds__(256) tdem_kernel(const __grid_constant__ TdDev S, const T* __restrict__ g_Mt, int B,
                                                    int l_stride, const int32_t* __restrict__ nlayers,
                                                    const double* __restrict__ sigma, const double* __restrict__ thickness,
                                                    const double* __restrict__ altitude, double* __restrict__ out,
                                                    double* __restrict__ Jout, const double out_scale)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ TdShared<T> sys_s;
    T* Mt = reinterpret_cast<T*>(smem);
    constexpr uint32_t mt_bytes = (uint32_t)(TD_ROWS * TD_CP * sizeof(T));
    if (threadIdx.x == 0) fill_td_shared<T>(S, sys_s);
    tma_stage(Mt, g_Mt, mt_bytes, &bar);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int C = S.C;
    constexpr int PER_WARP = 2 * KS + 2 * GBP_TD_MAXLAM + TD_ROWS + TD_CP + (SENS ? TD_CP * KS : 0);
    T* base = reinterpret_cast<T*>(smem + mt_bytes) + (size_t)warp * PER_WARP;
    T* msig = base;
    T* mthk = base + KS;
    T* lam = base + 2 * KS;
    T* wgt = lam + GBP_TD_MAXLAM;
    T* sbuf = wgt + GBP_TD_MAXLAM;
    T* pred = sbuf + TD_ROWS;
    T* J = pred + TD_CP;
#pragma unroll 1
    for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
        const int L = nlayers[b];
        if (lane < L) {
            msig[lane] = (T)sigma[(size_t)b * l_stride + lane];
            mthk[lane] = (T)thickness[(size_t)b * l_stride + lane];
        }
        // one pass per measured component: the x channels see the same admittance recursion through other Hankel weights
        // (J1 instead of J0), so a pass evaluates every window and keeps the channels of its component
#pragma unroll 1
        for (int pass = 0; pass <= (S.has_x ? 1 : 0); ++pass) {
            td_geometry<T>(sys_s, altitude[b], lam, wgt, pass == 1);
            tdem_eval<T>(sys_s, Mt, lam, wgt, L, msig, mthk, sbuf, pred, SENS ? J : nullptr, SENS);
#pragma unroll 1
            for (int c = lane; c < C; c += 32)
                if (sys_s.ccomp[c] == pass) out[(size_t)b * C + c] = (double)pred[c] * out_scale;
            if (SENS) {
#pragma unroll 1
                for (int i = lane; i < C * l_stride; i += 32) {
                    const int c = i / l_stride, kk = i % l_stride;
                    if (sys_s.ccomp[c] == pass) Jout[(size_t)b * C * l_stride + i] = (kk < L) ? (double)J[c * KS + kk] * out_scale : 0.0;
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace gbp
