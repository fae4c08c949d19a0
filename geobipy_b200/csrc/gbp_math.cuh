// Small complex / warp / RNG helpers for the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gbp {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------- real-type dispatch
template <typename T> struct rt;
// fp32: hardware special-function unit (MUFU) approximations, 1-2 instructions each.  Keeps the hot loop
// small (instruction-cache resident) - the accurate libm versions carry slow paths of several hundred
// instructions each.  Accuracy: rcp/sqrt/ex2 ~1-2 ulp; sin/cos 2^-21 absolute after the Cody-Waite reduction.
template <> struct rt<float> {
    static __device__ __forceinline__ float sqrt(float x)
    {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float exp(float x)
    {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
        return r;
    }
    static __device__ __forceinline__ float log(float x)
    {
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r * 0.6931471805599453f;
    }
    static __device__ __forceinline__ float rcp(float x)
    {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    // |x| <= ~64 rad on this path (callers clamp): one Cody-Waite step to [-pi, pi], then MUFU.SIN/COS
    static __device__ __forceinline__ void sincos(float x, float* s, float* c)
    {
        const float n = rintf(x * 0.15915494309189535f);
        float r = fmaf(n, -6.2831854820251465f, x);       // 2*pi (fp32 rounding)
        r = fmaf(n, 1.7484556000744883e-07f, r);           // - (2*pi - fp32(2*pi))
        asm("sin.approx.ftz.f32 %0, %1;" : "=f"(*s) : "f"(r));
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(*c) : "f"(r));
    }
};
__device__ __noinline__ double dexp_(double x) { return ::exp(x); }
__device__ __noinline__ double dlog_(double x) { return ::log(x); }
__device__ __noinline__ void dsincos_(double x, double* s, double* c) { ::sincos(x, s, c); }
template <> struct rt<double> {
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double exp(double x) { return dexp_(x); }
    static __device__ __forceinline__ double log(double x) { return dlog_(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { dsincos_(x, s, c); }
};

// ---------------------------------------------------------------- complex
template <typename T> struct cx {
    T re, im;
};
template <typename T> __device__ __forceinline__ cx<T> mk(T a, T b) { return cx<T>{a, b}; }
template <typename T> __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <typename T> __device__ __forceinline__ cx<T> operator*(cx<T> a, cx<T> b)
{
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename T> __device__ __forceinline__ cx<T> operator*(cx<T> a, T s) { return {a.re * s, a.im * s}; }
template <typename T> __device__ __forceinline__ cx<T> cinv(cx<T> a)
{
    T d = rt<T>::rcp(a.re * a.re + a.im * a.im);
    return {a.re * d, -a.im * d};
}
// sqrt of a + ib with b >= 0 (first-quadrant result)
template <typename T> __device__ __forceinline__ cx<T> csqrt_q1(T a, T b)
{
    T m = rt<T>::sqrt(a * a + b * b);
    T t = rt<T>::sqrt(T(0.5) * (m + fabs(a)));
    T o = b * rt<T>::rcp(t + t);
    return (a >= T(0)) ? cx<T>{t, o} : cx<T>{o, t};
}
// exp(z)
template <typename T> __device__ __forceinline__ cx<T> cexp_(cx<T> z)
{
    T e = rt<T>::exp(z.re), s, c;
    rt<T>::sincos(z.im, &s, &c);
    return {e * c, e * s};
}

// ---------------------------------------------------------------- warp reductions
template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// Several warp sums at once, "folded": at offset o a lane keeps one of two values and sends the other, so two
// registers become one per step - 2 values cost 5 shuffles, 8 values 9 (instead of 10 / 40 in dependent chains of 5).
// The additions pair the same operands as warp_sum's butterfly (fp addition commutes), so every sum has the bits
// warp_sum gives.  warp_sum2: lanes 0-15 hold sum(a), lanes 16-31 sum(b).  warp_sum8: value i ends in the four lanes
// 16 (i & 1) + 8 ((i >> 1) & 1) + 4 (i >> 2) + {0..3}.
__device__ __forceinline__ float warp_fold(float a, float b, int off, int lane)
{
    const bool hi = (lane & off) != 0;
    return (hi ? b : a) + __shfl_xor_sync(FULL, hi ? a : b, off);
}
__device__ __forceinline__ float warp_sum2(float a, float b, int lane)
{
    float v = warp_fold(a, b, 16, lane);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum8(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7, int lane)
{
    const float r01 = warp_fold(v0, v1, 16, lane), r23 = warp_fold(v2, v3, 16, lane);
    const float r45 = warp_fold(v4, v5, 16, lane), r67 = warp_fold(v6, v7, 16, lane);
    const float s0 = warp_fold(r01, r23, 8, lane), s1 = warp_fold(r45, r67, 8, lane);
    float t = warp_fold(s0, s1, 4, lane);
    t += __shfl_xor_sync(FULL, t, 2);
    t += __shfl_xor_sync(FULL, t, 1);
    return t;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
// Everything by value so that the generator state stays in registers: key = 64-bit seed,
// counter = (block, iteration, sounding lo, sounding hi): every accept_reject step of a chain has its own sub-stream
// starting at block 0, so an iteration's numbers do not depend on how many the previous iterations consumed
// (future iterations of a chain can be evaluated speculatively by other warps).
struct Rng {
    uint32_t seed_lo, seed_hi, snd_lo, snd_hi;
    uint32_t block, iter;
};
__device__ __noinline__ uint4 philox4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// one block -> two uniforms in [0,1): 53-bit for double, 24-bit for float
template <typename R> __device__ __forceinline__ void uniforms_of(uint4 x, R* ua, R* ub);
template <> __device__ __forceinline__ void uniforms_of<double>(uint4 x, double* ua, double* ub)
{
    *ua = ((double)(x.x >> 5) * 67108864.0 + (double)(x.y >> 6)) * (1.0 / 9007199254740992.0);
    *ub = ((double)(x.z >> 5) * 67108864.0 + (double)(x.w >> 6)) * (1.0 / 9007199254740992.0);
}
template <> __device__ __forceinline__ void uniforms_of<float>(uint4 x, float* ua, float* ub)
{
    *ua = (float)(x.x >> 8) * (1.0f / 16777216.0f);
    *ub = (float)(x.z >> 8) * (1.0f / 16777216.0f);
}
template <typename R> __device__ __forceinline__ R rng_uniform(Rng& g)
{
    R a, b;
    uniforms_of<R>(philox4(g.block, g.iter, g.snd_lo, g.snd_hi, g.seed_lo, g.seed_hi), &a, &b);
    g.block++;
    return a;
}
// Box-Muller pair of the block at an explicit counter
template <typename R> struct pair_t {
    R a, b;
};
template <typename R> __device__ __noinline__ pair_t<R> normal2_at(uint32_t block, uint32_t iter, uint32_t snd_lo, uint32_t snd_hi, uint32_t k0, uint32_t k1)
{
    R a, b;
    uniforms_of<R>(philox4(block, iter, snd_lo, snd_hi, k0, k1), &a, &b);
    const R r = rt<R>::sqrt(R(-2) * rt<R>::log(R(1) - a));
    R s, c;
    rt<R>::sincos(R(6.283185307179586476925286766559) * b, &s, &c);
    return pair_t<R>{r * c, r * s};
}
template <typename R> __device__ __forceinline__ R rng_normal(Rng& g)
{
    const pair_t<R> z = normal2_at<R>(g.block, g.iter, g.snd_lo, g.snd_hi, g.seed_lo, g.seed_hi);
    g.block++;
    return z.a;
}

}  // namespace gbp
