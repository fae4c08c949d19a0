// Warp-cooperative 1D layered-earth FDEM forward + analytic Jacobian (one warp = one sounding).
//
// Replaces nbFdem1dfwd / nbFdem1dsen (geobipy/src/classes/forwardmodelling/Electromagnetic/FD/
// fdem1d_numba.py:25-121, initCoefficients :158, M1_0 :195, M1_1 :223, Hxx..Hzz :307-438, cTanh :442).
//
// Re-design (not a translation):
//  * lane = filter abscissa.  Only the (frequency, filter) pairs a tensor id consumes are evaluated
//    (RESOLVE: 860 points instead of the reference's 1560).
//  * the bottom-up admittance recursion runs in registers on y = i*omega*mu0 * Y (so the reference's
//    1/z-hat factors cancel analytically); per-layer derivative factors needed by the top-down
//    chain-rule pass live in thread-local scratch (L1-resident).
//  * everything that does not depend on the earth model is folded on the host (fp64) into a per
//    abscissa table staged to shared memory with one TMA bulk copy per CTA: lambda, the air-layer
//    wavenumber u0, the exponent of the height term and the complex constant
//    c = 1e6*scale/H0 * (geometry * weight * lambda^n [/u0]).  The response is then
//        d_f = sum_j rTE_j * c_j * exp(e_j * hDiff)
//    i.e. the secondary field is accumulated directly - the reference's (H - H0)/H0 cancellation
//    never happens, which is what makes the fp32 instantiation accurate.
#pragma once
#include "gbp_math.cuh"
#include "../../include/geobipy_b200.h"

namespace gbp {

constexpr int KS = GBP_MAXL;        // layer stride of per-warp arrays
constexpr int TAB_ROWS = 7;         // lam, u0re, u0im, ere, eim, cre, cim
constexpr int MAX_SEG = 2 * GBP_MAXF;
// fp32 path: the abscissae of a system are cut into CHUNKS of 64 (two per lane, packed fp32x2); a chunk never spans
// two frequencies.  120-point J0 filter: 2 chunks, 140-point J1 filter: 3 chunks -> at most 5 per frequency.
constexpr int CHUNK = 64;
constexpr int MAX_CHUNK = 5 * GBP_MAXF;
constexpr int CHUNK_FLOATS = 4 * 32 * 4;   // 4 quads of one float4 per lane

struct Seg {
    int start, count, freq, pad;
};

// Per-system constants, passed by value as a kernel parameter.
struct SysDev {
    int n_freq, n_seg, n_items, tab_stride;  // tab_stride = n_items rounded up (16-byte rows)
    int n_chunks, half_begin[3];             // chunks [half_begin[h], half_begin[h+1]) = half h of a forward
    // work units of a team round (gbp_chain.cuh): unit r = the chunks [unit_begin[r], unit_begin[r] + unit_count[r]) of
    // ONE frequency, frequencies with the most chunks first
    unsigned char unit_begin[GBP_MAXF], unit_count[GBP_MAXF];
    unsigned char chunk_freq[MAX_CHUNK];     // frequency of chunk c (chunks are ordered by frequency)
    Seg seg[MAX_SEG];
    double omu[GBP_MAXF];   // omega * mu0
    double k2re[GBP_MAXF];  // -omega^2 * mu0 * eps0   (Re of y-hat*z-hat, same for every earth layer)
    double hd0[GBP_MAXF];   // hDiff = hd0 - 2*altitude  (fdem1d.py:31-32: rz - 2 tz - 2 z)
};

// ---------------------------------------------------------------- TMA bulk copy (global -> shared)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One elected thread issues cp.async.bulk; everybody waits on the mbarrier.  bytes % 16 == 0.
__device__ __forceinline__ void tma_stage(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    const uint32_t b = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(smem_dst)),
            "l"(gmem_src), "r"(bytes), "r"(b)
            : "memory");
    }
    uint32_t done = 0;
    #pragma unroll 1
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(b)
            : "memory");
    }
}

// Per-CTA copy of the per-frequency constants in the arithmetic type T (shared memory).
template <typename T> struct SysShared {
    int n_freq, n_seg, tab_stride, n_chunks;
    int half_begin[3], pad;
    unsigned char unit_begin[GBP_MAXF], unit_count[GBP_MAXF];
    unsigned char chunk_freq[MAX_CHUNK];
    Seg seg[MAX_SEG];
    T omu[GBP_MAXF], k2re[GBP_MAXF], hd0[GBP_MAXF];
};
template <typename T> __device__ __forceinline__ void fill_sys_shared(const SysDev& S, SysShared<T>& q)
{
    q.n_freq = S.n_freq;
    q.n_seg = S.n_seg;
    q.tab_stride = S.tab_stride;
    q.n_chunks = S.n_chunks;
    for (int i = 0; i < 3; ++i) q.half_begin[i] = S.half_begin[i];
    for (int i = 0; i < GBP_MAXF; ++i) {
        q.unit_begin[i] = S.unit_begin[i];
        q.unit_count[i] = S.unit_count[i];
    }
    for (int i = 0; i < MAX_CHUNK; ++i) q.chunk_freq[i] = S.chunk_freq[i];
    for (int i = 0; i < MAX_SEG; ++i) q.seg[i] = S.seg[i];
    for (int i = 0; i < GBP_MAXF; ++i) {
        q.omu[i] = (T)S.omu[i];
        q.k2re[i] = (T)S.k2re[i];
        q.hd0[i] = (T)S.hd0[i];
    }
}

// bytes of the per-system table of arithmetic type T: fp64 = TAB_ROWS rows of tab_stride values; fp32 = the packed
// chunk layout of gbp_fdem_f2.cuh
template <typename T> __host__ __device__ inline uint32_t fdem_table_bytes(const SysDev& S)
{
    return sizeof(T) == 4 ? (uint32_t)(S.n_chunks * CHUNK_FLOATS * sizeof(float)) : (uint32_t)(TAB_ROWS * S.tab_stride * sizeof(double));
}

// ---------------------------------------------------------------- the operator
// Q, tab, msig, mthk, pred, J all live in SHARED memory (asserted to the compiler so that the loads are
// LDS, not generic).  tab: TAB_ROWS rows of tab_stride values of T.  msig/mthk: per-warp model
// (conductivity, thickness); mthk[L-1] unused.  pred: [2F] (real then imag).  J: [2F][KS], written only
// if sens.  One function body serves forward-only and forward+Jacobian calls (`sens` is warp-uniform) so
// that all warps of an SM share the same instruction-cache lines.  All lanes of the warp must call.
template <typename T>
__device__ __noinline__ void fdem_eval(const SysShared<T>& Q, const T* __restrict__ tab, T alt, int L,
                                       const T* __restrict__ msig, const T* __restrict__ mthk, T* __restrict__ pred,
                                       T* __restrict__ J, const bool sens)
{
    __builtin_assume(__isShared(&Q));
    __builtin_assume(__isShared(tab));
    __builtin_assume(__isShared(msig));
    __builtin_assume(__isShared(mthk));
    __builtin_assume(__isShared(pred));
    const int lane = threadIdx.x & 31;
    const int F = Q.n_freq;
    const int ts = Q.tab_stride;
    const T* t_lam = tab;
    const T* t_u0r = tab + ts;
    const T* t_u0i = tab + 2 * ts;
    const T* t_er = tab + 3 * ts;
    const T* t_ei = tab + 4 * ts;
    const T* t_cr = tab + 5 * ts;
    const T* t_ci = tab + 6 * ts;

    // thread-local scratch of the chain-rule pass (sens only): D_k = dy_k/dy_{k+1}, l_k = dy_k/dln(sigma_k)
    T Dr[KS], Di[KS], lr[KS], li[KS];
    T jr[KS], ji[KS];

    int seg = 0;
#pragma unroll 1
    for (int f = 0; f < F; ++f) {
        const T omu = Q.omu[f];
        const T k2 = Q.k2re[f];
        const T hd = Q.hd0[f] - T(2) * alt;
        cx<T> acc = {T(0), T(0)};
        if (sens) {
#pragma unroll 1
            for (int k = 0; k < L; ++k) {
                jr[k] = T(0);
                ji[k] = T(0);
            }
        }
#pragma unroll 1
        for (; seg < Q.n_seg && Q.seg[seg].freq == f; ++seg) {
            const int s0 = Q.seg[seg].start, cnt = Q.seg[seg].count;
#pragma unroll 1
            for (int j = lane; j < cnt; j += 32) {
                const int i = s0 + j;
                const T lam = t_lam[i];
                const T a = lam * lam + k2;  // Re(u^2) of every earth layer
                // basement: y_L = u_L
                T b = omu * msig[L - 1];
                cx<T> u = csqrt_q1<T>(a, b);
                cx<T> y = u;
                if (sens) {  // i*b/(2u)
                    cx<T> iu = cinv(u);
                    lr[L - 1] = T(-0.5) * b * iu.im;
                    li[L - 1] = T(0.5) * b * iu.re;
                }
#pragma unroll 1
                for (int k = L - 2; k >= 0; --k) {
                    b = omu * msig[k];
                    const T t = mthk[k];
                    u = csqrt_q1<T>(a, b);
                    // tanh(u t) = (1 - e)/(1 + e), e = exp(-2ut), Re(u) > 0 always (cTanh first branch).
                    // |Im(2ut)| <= Re(2ut); beyond Re(2ut) = 60 the term e^-60 is below fp64 round-off of
                    // tanh = 1: clamp so that the sin/cos argument stays small (t may be huge), and set e to
                    // exactly 0 there, as exp() underflows to in the reference
                    const T two_t = T(2) * t;
                    const T sc = fmin(two_t, T(60) * rt<T>::rcp(u.re));
                    cx<T> e = cexp_<T>(mk<T>(-sc * u.re, -sc * u.im));
                    if (two_t * u.re > T(60)) e = mk<T>(T(0), T(0));
                    cx<T> th = mk<T>(T(1) - e.re, -e.im) * cinv(mk<T>(T(1) + e.re, e.im));
                    cx<T> den = u + y * th;
                    cx<T> num = y + u * th;
                    cx<T> inv = cinv(den);
                    if (sens) {
                        cx<T> u2 = mk<T>(a, b);
                        cx<T> th2 = th * th;
                        cx<T> inv2 = inv * inv;
                        cx<T> w = y * y - u2;                              // y^2 - u^2
                        cx<T> one_m = mk<T>(T(1) - th2.re, -th2.im);       // 1 - tanh^2
                        cx<T> d = u2 * one_m * inv2;                       // accumulate[] of M1_1
                        Dr[k] = d.re;
                        Di[k] = d.im;
                        // B = 2uy th^2 + (y^2-u^2) th + 2u^2 - t u (y^2-u^2)(1 - th^2)
                        cx<T> uy = u * y;
                        cx<T> B = (uy * th2) * T(2) + w * th + u2 * T(2) - ((u * w) * one_m) * t;
                        cx<T> q = B * inv2 * cinv(u);                      // B / (u den^2)
                        lr[k] = T(-0.5) * b * q.im;                         // * i*b/2
                        li[k] = T(0.5) * b * q.re;
                    }
                    y = u * num * inv;
                }
                const cx<T> u0 = mk<T>(t_u0r[i], t_u0i[i]);
                const cx<T> is = cinv(u0 + y);
                const cx<T> rte = (u0 - y) * is;
                const cx<T> K = mk<T>(t_cr[i], t_ci[i]) * cexp_<T>(mk<T>(t_er[i] * hd, t_ei[i] * hd));
                acc = acc + rte * K;
                if (sens) {
                    cx<T> P = (u0 * is * is) * T(-2) * K;  // d rTE/dy1 * K
#pragma unroll 1
                    for (int k = 0; k < L; ++k) {
                        cx<T> v = P * mk<T>(lr[k], li[k]);
                        jr[k] += v.re;
                        ji[k] += v.im;
                        if (k < L - 1) P = P * mk<T>(Dr[k], Di[k]);
                    }
                }
            }
        }
        const T sr = warp_sum(acc.re), si = warp_sum(acc.im);
        if (lane == 0) {
            pred[f] = sr;
            pred[F + f] = si;
        }
        if (sens) {
#pragma unroll 1
            for (int k = 0; k < L; ++k) {
                const T a = warp_sum(jr[k]), b = warp_sum(ji[k]);
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
