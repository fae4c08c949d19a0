// C-ABI of geobipy_b200 (see include/geobipy_b200.h).  Host side: table construction, device
// staging, launch configuration.  No CPU fallback: every compute entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "gbp_chain.cuh"
#include "gbp_tables.h"
#include "gbp_tdem_tables.h"

using namespace gbp;

namespace {

thread_local std::string g_err;
std::mutex g_mu;
std::atomic<long long> g_launches{0};
constexpr int MAX_DEV = 64;

// Per-device launch state.  Threading contract (include/geobipy_b200.h): entry points may be called from several
// host threads, on several devices and streams at once - every launch takes its device's mutex for the (short,
// asynchronous) launch sequence, the per-call scratch (Jacobian mirror, work counter) is allocated stream-ordered
// per call, and the timing events form a per-device ring of the last TIMING_RING launches.
constexpr int TIMING_RING = 32;
struct DevState {
    std::mutex mu;
    cudaEvent_t ev0[TIMING_RING] = {nullptr}, ev1[TIMING_RING] = {nullptr};
    long long n_timed = 0;
    bool pool_ready = false;
};
DevState g_dev[MAX_DEV];

int fail(const std::string& m)
{
    g_err = m;
    return 1;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

// Device allocations of one host-pointer call: freed when the scope ends, on the error paths too.
struct DevScope {
    std::vector<void*> ptrs;
    std::vector<cudaEvent_t> events;
    ~DevScope()
    {
        for (void* q : ptrs) cudaFree(q);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
    template <typename U> cudaError_t alloc(U** out, size_t bytes)
    {
        const cudaError_t e = cudaMalloc((void**)out, bytes);
        if (e == cudaSuccess) ptrs.push_back((void*)*out);
        return e;
    }
    cudaError_t event(cudaEvent_t* ev)
    {
        const cudaError_t e = cudaEventCreate(ev);
        if (e == cudaSuccess) events.push_back(*ev);
        return e;
    }
};

// device copies of the filter-point table, cached per (device, system)
struct TableCache {
    int device = -1;
    gbp_fdem_system sys;
    SysHost host;
    float* d_f32 = nullptr;
    double* d_f64 = nullptr;
};
std::vector<TableCache*> g_cache;

int get_tables(const gbp_fdem_system* sys, TableCache** out)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    for (TableCache* c : g_cache)
        if (c->device == dev && std::memcmp(&c->sys, sys, sizeof(*sys)) == 0) {
            *out = c;
            return 0;
        }
    TableCache* c = new TableCache();
    c->device = dev;
    c->sys = *sys;
    if (!build_system_tables(*sys, c->host)) {
        std::string e = c->host.error;
        delete c;
        return fail(e);
    }
    const size_t n = c->host.tab.size();
    const std::vector<float>& f32 = c->host.tab_f32;   // packed chunk layout (gbp_tables.h)
    CK(cudaMalloc(&c->d_f32, f32.size() * sizeof(float)));
    CK(cudaMalloc(&c->d_f64, n * sizeof(double)));
    CK(cudaMemcpy(c->d_f32, f32.data(), f32.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_f64, c->host.tab.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    g_cache.push_back(c);
    *out = c;
    return 0;
}

int current_device(int* dev)
{
    CK(cudaGetDevice(dev));
    if (*dev < 0 || *dev >= MAX_DEV) return fail("device index out of range");
    return 0;
}
// both called with g_dev[dev].mu held
int time_begin(int dev, cudaStream_t st)
{
    DevState& D = g_dev[dev];
    const int slot = (int)(D.n_timed % TIMING_RING);
    if (!D.ev0[slot]) {
        CK(cudaEventCreate(&D.ev0[slot]));
        CK(cudaEventCreate(&D.ev1[slot]));
    }
    CK(cudaEventRecord(D.ev0[slot], st));
    return 0;
}
int time_end(int dev, cudaStream_t st)
{
    DevState& D = g_dev[dev];
    CK(cudaEventRecord(D.ev1[(int)(D.n_timed % TIMING_RING)], st));
    D.n_timed++;
    return 0;
}
// stream-ordered scratch of one call; the device's default memory pool keeps what it frees (no per-call cudaMalloc)
int scratch_alloc(int dev, void** p, size_t bytes, cudaStream_t st)
{
    DevState& D = g_dev[dev];
    if (!D.pool_ready) {
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        D.pool_ready = true;
    }
    CK(cudaMallocAsync(p, bytes, st));
    return 0;
}

template <typename T> const T* tab_ptr(TableCache* c);
template <> const float* tab_ptr<float>(TableCache* c) { return c->d_f32; }
template <> const double* tab_ptr<double>(TableCache* c) { return c->d_f64; }

int sm_count()
{
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

template <typename T, bool SENS>
int launch_fdem(TableCache* tc, int B, int l_stride, const int32_t* nl, const double* sig, const double* thk,
                const double* alt, double* out, double* J, cudaStream_t st)
{
    const int threads = 256, wpb = threads / 32;
    const size_t tab_bytes = ((size_t)fdem_table_bytes<T>(tc->host.dev) + 127) & ~(size_t)127;
    const size_t per_warp = (2 * KS + GBP_MAXC + (SENS ? GBP_MAXC * KS : 0)) * sizeof(T);
    const size_t smem = tab_bytes + wpb * per_warp;
    auto kern = fdem_kernel<T, SENS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int grid = sm_count() * blocks_per_sm;
    const int need = (B + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    int dev;
    if (current_device(&dev)) return 1;
    std::lock_guard<std::mutex> lk(g_dev[dev].mu);
    if (time_begin(dev, st)) return 1;
    kern<<<grid, threads, smem, st>>>(tc->host.dev, tab_ptr<T>(tc), B, l_stride, nl, sig, thk, alt, out, J);
    g_launches++;
    CK(cudaGetLastError());
    return time_end(dev, st);
}

// device copies of the time-domain window operator, cached per (device, survey)
constexpr double TD_F32_SCALE = 1099511627776.0;  // 2^40: fp32 path works in scaled data units (exact in fp64)
struct TdCache {
    int device = -1;
    gbp_tdem_survey sv;
    TdHost host;
    float* d_f32 = nullptr;   // Mt * TD_F32_SCALE
    double* d_f64 = nullptr;
};
std::vector<TdCache*> g_td_cache;

int get_td_tables(const gbp_tdem_survey* sv, TdCache** out, bool need_device)
{
    int dev = -1;
    if (need_device) CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    for (TdCache* c : g_td_cache)
        if (c->device == dev && std::memcmp(&c->sv, sv, sizeof(*sv)) == 0) {
            *out = c;
            return 0;
        }
    TdCache* c = new TdCache();
    c->device = dev;
    c->sv = *sv;
    if (!build_tdem_tables(*sv, c->host)) {
        std::string e = c->host.error;
        delete c;
        return fail(e);
    }
    if (need_device) {
        const size_t n = c->host.Mt.size();
        std::vector<float> f32(n);
        for (size_t i = 0; i < n; ++i) f32[i] = (float)(c->host.Mt[i] * TD_F32_SCALE);
        CK(cudaMalloc(&c->d_f32, n * sizeof(float)));
        CK(cudaMalloc(&c->d_f64, n * sizeof(double)));
        CK(cudaMemcpy(c->d_f32, f32.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_f64, c->host.Mt.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    }
    g_td_cache.push_back(c);
    *out = c;
    return 0;
}
template <typename T> const T* td_ptr(TdCache* c);
template <> const float* td_ptr<float>(TdCache* c) { return c->d_f32; }
template <> const double* td_ptr<double>(TdCache* c) { return c->d_f64; }

template <typename T, bool SENS>
int launch_tdem(TdCache* tc, int B, int l_stride, const int32_t* nl, const double* sig, const double* thk,
                const double* alt, double* out, double* J, double out_scale, cudaStream_t st)
{
    const int threads = 256, wpb = threads / 32;
    const size_t mt_bytes = (size_t)TD_ROWS * TD_CP * sizeof(T);
    const size_t per_warp = (2 * KS + 2 * GBP_TD_MAXLAM + TD_ROWS + TD_CP + (SENS ? TD_CP * KS : 0)) * sizeof(T);
    const size_t smem = mt_bytes + wpb * per_warp;
    auto kern = tdem_kernel<T, SENS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int grid = sm_count() * blocks_per_sm;
    const int need = (B + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    int dev;
    if (current_device(&dev)) return 1;
    std::lock_guard<std::mutex> lk(g_dev[dev].mu);
    if (time_begin(dev, st)) return 1;
    kern<<<grid, threads, smem, st>>>(tc->host.dev, td_ptr<T>(tc), B, l_stride, nl, sig, thk, alt, out, J, out_scale);
    g_launches++;
    CK(cudaGetLastError());
    return time_end(dev, st);
}

unsigned long long* g_finish[MAX_DEV] = {nullptr};  // debug timeline (GBP_DEBUG_TIMELINE)
size_t g_finish_cap[MAX_DEV] = {0};
int g_finish_B[MAX_DEV] = {0};

template <typename R, typename T, int NC, int WARPS, int KIND>
int launch_chain(const typename SysOf<T, KIND>::dev& sysdev, const T* d_tab, size_t tab_bytes_raw, const ChainParams& P,
                 cudaStream_t st)
{
    if (P.C > NC) return fail("datapoint has more channels than this sampler kernel holds");
    const size_t tab_bytes = (tab_bytes_raw + 127) & ~(size_t)127;
    const size_t per_warp = sizeof(WarpState<R, T, NC, KIND>);
    int dev = 0, max_smem = 0;
    if (current_device(&dev)) return 1;
    CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t smem = tab_bytes + (size_t)WARPS * per_warp;
    auto kern = rjmcmc_kernel<R, T, NC, WARPS, KIND>;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    if (smem + fa.sharedSizeBytes + 1024 > (size_t)max_smem) return fail("rjmcmc kernel does not fit in shared memory on this device");
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = sm_count();
    if (grid > P.B) grid = P.B;
    ChainParams Q = P;
    Q.n_warps_total = grid * WARPS;
    {
        const char* e = std::getenv("GBP_SPEC_HELPERS");  // 0 switches speculative evaluation off
        int hmax = e ? std::atoi(e) : 12;
        Q.spec_helpers = hmax < 0 ? 0 : (hmax > WARPS - 1 ? WARPS - 1 : hmax);
        const char* m = std::getenv("GBP_SPEC_MIN_REJECTIONS");
        Q.spec_min_rejections = m ? std::atoi(m) : 24;
        // teams of warps share their forward evaluations (fp32 frequency-domain kernels; gbp_chain.cuh team_round).
        // Measured (profiles/README.md round 2): 8 warps per team, members on different SM sub-partitions, work units of
        // one frequency.  GBP_TEAM=1 switches it off (results are bit-identical either way).
        const char* t = std::getenv("GBP_TEAM");
        Q.team_size = t ? std::atoi(t) : 8;
        const char* ts = std::getenv("GBP_TEAM_SPREAD");
        Q.team_spread = ts ? std::atoi(ts) : 1;
        const char* tu = std::getenv("GBP_TEAM_FREQ_UNITS");
        Q.team_freq_units = tu ? std::atoi(tu) : 1;
    }
    std::lock_guard<std::mutex> lk(g_dev[dev].mu);
    // per-call scratch, stream ordered: [work counter (256 B)] [per-chain Jacobian mirror, L2 resident: 1.4 KB per
    // chain in fp32].  Calls in flight on different streams never share it.
    const size_t jbytes = (size_t)P.B * NC * KS * sizeof(T);
    unsigned char* scratch = nullptr;
    if (scratch_alloc(dev, (void**)&scratch, 256 + jbytes, st)) return 1;
    Q.work_counter = (int*)scratch;
    Q.jstore = scratch + 256;
    Q.finish_ns = nullptr;
    if (std::getenv("GBP_DEBUG_TIMELINE")) {
        const size_t need = ((size_t)(P.B + 1) + (size_t)P.B * 32) * sizeof(unsigned long long);   // + 32 progress stamps per chain
        if (g_finish_cap[dev] < need) {
            if (g_finish[dev]) CK(cudaFree(g_finish[dev]));
            g_finish[dev] = nullptr;
            g_finish_cap[dev] = 0;
            CK(cudaMalloc(&g_finish[dev], need));
            g_finish_cap[dev] = need;
        }
        CK(cudaMemsetAsync(g_finish[dev], 0xff, need, st));
        g_finish_B[dev] = P.B;
        Q.finish_ns = g_finish[dev];
    }
    // device-side work counter: chains beyond the first wave are claimed dynamically
#ifdef GBP_STATIC_FIRST_WAVE
    CK(cudaMemcpyAsync(Q.work_counter, &Q.n_warps_total, sizeof(int), cudaMemcpyHostToDevice, st));
#else
    CK(cudaMemsetAsync(Q.work_counter, 0, sizeof(int), st));
#endif
    if (time_begin(dev, st)) return 1;
    kern<<<grid, WARPS * 32, smem, st>>>(sysdev, d_tab, Q);
    g_launches++;
    CK(cudaGetLastError());
    if (time_end(dev, st)) return 1;
    CK(cudaFreeAsync(scratch, st));
    return 0;
}

// ---- microbenchmarks of the two pipes that bound this path (the roofline denominators for scalar
// fp32 issue and special-function throughput must be measured on the same box)
__global__ void __launch_bounds__(256) peak_fma_kernel(float* out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void __launch_bounds__(256) peak_mufu_kernel(float* out, int iters)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + .1f, x2 = x0 + .2f, x3 = x0 + .3f, x4 = x0 + .4f, x5 = x0 + .5f, x6 = x0 + .6f, x7 = x0 + .7f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x3));
            asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x4));
            asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x5));
            asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(x6));
            asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(x7));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---- posterior summaries: per depth cell mean and percentiles of ln(sigma) from a hitmap
// [B][n_sig][n_depth] (depth contiguous).  One thread per (sounding, depth cell): consecutive threads read
// consecutive depth cells of one sigma row, so every load is coalesced; HBM bound (the hitmaps are read once from
// HBM, the second pass over a column mostly hits L2).  Mesh._mean / Mesh._percentile (classes/mesh/Mesh.py:80,
// :173-217): the percentile is the centre of the first bin whose cumulative count reaches p % of the column total.
constexpr int SUMM_MAXP = 8;
struct SummParams {
    int B, n_sig, n_depth, n_pct;
    double dx;
    double pct[SUMM_MAXP];      // percentiles; when a credible range is asked for, the last two are its bounds
    int want_range;
};
// smallest integer count c with (double)c / (double)tot >= p, p = percent * 0.01 in fp64: exactly the reference's
// rule (Mesh._percentile: searchsorted(cumsum / total, percent * 0.01), Mesh.py:196-208) including its round-off -
// 95.0 * 0.01 = 0.9500000000000001 > 0.95, so a cumulative fraction of exactly 0.95 goes to the NEXT bin
__device__ __forceinline__ long long percentile_threshold(double percent, long long tot)
{
    if (tot <= 0) return 1;   // empty column: every bin is "below" -> the last bin, as the reference's clip
    const double p = percent * 0.01, t = (double)tot;
    long long c = (long long)ceil(p * t);
    if (c < 0) c = 0;
    while (c > 0 && (double)(c - 1) / t >= p) --c;
    while ((double)c / t < p) ++c;
    return c;
}
template <int NP>
__global__ void __launch_bounds__(256) summarise_kernel(const int32_t* __restrict__ hitmap, const double* __restrict__ sig_lo,
                                                         const __grid_constant__ SummParams P, double* __restrict__ mean,
                                                         double* __restrict__ pct, double* __restrict__ mode,
                                                         int32_t* __restrict__ range_bins)
{
    const long long n = (long long)P.B * P.n_depth;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / P.n_depth), j = (int)(i % P.n_depth);
        const int32_t* col = hitmap + (size_t)b * P.n_sig * P.n_depth + j;
        // integer inner loops (fp64 compares / conversions per element made the first version compute bound)
        long long tot = 0;
#pragma unroll 10
        for (int s = 0; s < P.n_sig; ++s) tot += col[(size_t)s * P.n_depth];  // independent loads: 10 in flight per thread
        const double totd = tot > 0 ? (double)tot : 1.0, lo = sig_lo[b];
        long long thr[NP];
        int idx[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            thr[q] = percentile_threshold(P.pct[q], tot);
            idx[q] = 0;
        }
        long long cs = 0, wsum = 0;
        int vmax = -1, imax = 0;   // mode: the first fullest bin (Mesh._mode: argmax, Mesh.py:138-165)
#pragma unroll 10
        for (int s = 0; s < P.n_sig; ++s) {
            const int v32 = col[(size_t)s * P.n_depth];
            const long long v = v32;
            cs += v;
            wsum += v * s;
            if (v32 > vmax) {
                vmax = v32;
                imax = s;
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) idx[q] += (cs < thr[q]) ? 1 : 0;
        }
        const double acc = (double)tot * (lo + 0.5 * P.dx) + (double)wsum * P.dx;
        mean[i] = acc / totd;
        const int n_out = P.want_range ? NP - 2 : NP;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int k = idx[q] < P.n_sig - 1 ? idx[q] : P.n_sig - 1;
            idx[q] = k;
            if (q < n_out) pct[(size_t)q * n + i] = lo + ((double)k + 0.5) * P.dx;
        }
        if (mode) mode[i] = lo + ((double)imax + 0.5) * P.dx;
        // credible range in BINS (Mesh._credible_range, Mesh.py:58-78: |log10 hi - log10 lo| = bins * dx / ln 10)
        if (P.want_range && range_bins) range_bins[i] = idx[NP - 1] - idx[NP - 2];
    }
}

// Opacity, depth of investigation and opacity level from credible ranges (in bins) [B][n_depth]: one warp per sounding.
// Histogram.transparency (Histogram.py:509-541): (range - min) / (max - min) over the group the sounding belongs to - the
// sounding itself (Histogram.opacity of one hitmap) or its whole flight line (Inference2D.compute_opacity :1011-1023);
// compute_doi (:493-532): from the bottom up, the first cell whose opacity reaches doi_p, never above cell 1's
// predecessor test (j >= 1); Histogram.opacity_level (:356-367): from the bottom up while transparency > level_p.
__global__ void __launch_bounds__(256) range_minmax_kernel(const int32_t* __restrict__ range_bins, const int32_t* __restrict__ group,
                                                            int B, int n_depth, int32_t* __restrict__ gmin, int32_t* __restrict__ gmax)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    int mn = 0x7fffffff, mx = -0x7fffffff;
    for (int j = lane; j < n_depth; j += 32) {
        const int v = range_bins[(size_t)warp * n_depth + j];
        mn = min(mn, v);
        mx = max(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(FULL, mn, o));
        mx = max(mx, __shfl_xor_sync(FULL, mx, o));
    }
    if (lane == 0) {
        const int g = group ? group[warp] : warp;
        atomicMin(&gmin[g], mn);
        atomicMax(&gmax[g], mx);
    }
}
__global__ void __launch_bounds__(256) opacity_doi_kernel(const int32_t* __restrict__ range_bins, const int32_t* __restrict__ group,
                                                           int B, int n_depth, const int32_t* __restrict__ gmin,
                                                           const int32_t* __restrict__ gmax, double doi_p, double level_p,
                                                           double* __restrict__ opacity, int32_t* __restrict__ doi_cell,
                                                           int32_t* __restrict__ level_cell)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int g = group ? group[warp] : warp;
    const int mn = gmin[g], mx = gmax[g];
    const double span = (double)(mx - mn);
    int jd = 0, jl = -1;
    for (int j = lane; j < n_depth; j += 32) {
        const double r = (double)(range_bins[(size_t)warp * n_depth + j] - mn);
        const double t = span > 0.0 ? r / span : r;   // transparency
        const double op = 1.0 - t;
        if (opacity) opacity[(size_t)warp * n_depth + j] = op;
        if (op >= doi_p && j >= 1) jd = max(jd, j);   // loop of compute_doi: stops at the deepest cell with opacity >= p, or at 0
        if (!(t > level_p)) jl = max(jl, j);           // opacity_level: stops at the deepest cell with transparency <= p, or at -1
    }
    for (int o = 16; o > 0; o >>= 1) {
        jd = max(jd, __shfl_xor_sync(FULL, jd, o));
        jl = max(jl, __shfl_xor_sync(FULL, jl, o));
    }
    if (lane == 0) {
        if (doi_cell) doi_cell[warp] = jd;
        if (level_cell) level_cell[warp] = jl < 0 ? n_depth - 1 : jl;   // index -1 = the last cell, as in the reference
    }
}

int check_options(const gbp_options* o)
{
    if (o->max_layers < 1 || o->max_layers > GBP_MAXL) return fail("max_layers must be in [1, 30]");
    if (o->n_markov_chains < 1) return fail("n_markov_chains must be >= 1");
    if (o->update_plot_every < 1) return fail("update_plot_every must be >= 1");
    if (!(o->min_width > 0.0) || !(o->max_edge > o->min_edge) || !(o->min_edge > 0.0))
        return fail("invalid depth limits");
    if (!(o->min_width * o->max_layers < o->max_edge))
        return fail("min_width * max_layers must be < max_edge (RectilinearMesh1D.set_priors)");
    if (o->n_sigma_bins < 1 || o->n_err_bins < 1) return fail("invalid bin counts");
    if (o->solve_height && (!(o->max_height_change > 0.0) || !(o->height_prop_var > 0.0)))
        return fail("solve_height needs max_height_change > 0 and height_prop_var > 0");
    return 0;
}

}  // namespace

extern "C" {

const char* gbp_version(void) { return "geobipy_b200 0.1.0 (sm_100a)"; }
const char* gbp_last_error(void) { return g_err.c_str(); }

int gbp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int gbp_n_depth(const gbp_options* o)
{
    const double stop = 1.1 * o->max_edge, step = 0.5 * o->min_width;
    return (int)std::ceil(stop / step) - 1;
}

int gbp_filter_points(const gbp_fdem_system* sys)
{
    int n = 0;
    for (int f = 0; f < sys->n_freq; ++f) {
        const int t = sys->tid[f];
        if (t == 1 || t == 9) n += GBP_NJ0;
        if (t == 1 || t == 3 || t == 7) n += GBP_NJ1;
    }
    return n;
}

// SURVEY.md section 8(d) count convention: per abscissa 17(L+1) + 58(L-1) + 15 + 65 = 75L + 39 flops
double gbp_flops_per_forward(const gbp_fdem_system* sys, int L)
{
    return (double)gbp_filter_points(sys) * (75.0 * (double)L + 39.0);
}

int64_t gbp_launch_count(void) { return g_launches.load(); }

int gbp_last_kernel_ms(float* ms)
{
    int n = 0;
    return gbp_kernel_ms_stats(1, ms, &n);
}

int gbp_kernel_ms_stats(int last_n, float* mean_ms, int* counted)
{
    int dev;
    if (current_device(&dev)) return 1;
    DevState& D = g_dev[dev];
    std::lock_guard<std::mutex> lk(D.mu);
    if (D.n_timed == 0) return fail("no kernel has been timed yet on this device");
    long long n = last_n < 1 ? 1 : last_n;
    if (n > D.n_timed) n = D.n_timed;
    if (n > TIMING_RING) n = TIMING_RING;
    double sum = 0.0;
    for (long long i = D.n_timed - n; i < D.n_timed; ++i) {
        const int slot = (int)(i % TIMING_RING);
        float ms = 0.f;
        CK(cudaEventSynchronize(D.ev1[slot]));
        CK(cudaEventElapsedTime(&ms, D.ev0[slot], D.ev1[slot]));
        sum += ms;
    }
    if (mean_ms) *mean_ms = (float)(sum / (double)n);
    if (counted) *counted = (int)n;
    return 0;
}

int gbp_fdem_forward(const gbp_fdem_system* sys, int B, int l_stride, const int32_t* d_nlayers, const double* d_sigma,
                     const double* d_thickness, const double* d_altitude, double* d_out, int precision, void* stream)
{
    if (B <= 0) return 0;
    if (l_stride < 1 || l_stride > GBP_MAXL) return fail("l_stride must be in [1, 30]");
    TableCache* tc;
    if (get_tables(sys, &tc)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == GBP_PRECISION_F32)
        return launch_fdem<float, false>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, nullptr, st);
    if (precision == GBP_PRECISION_F64)
        return launch_fdem<double, false>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, nullptr, st);
    return fail("precision must be 32 or 64");
}

int gbp_fdem_sensitivity(const gbp_fdem_system* sys, int B, int l_stride, const int32_t* d_nlayers,
                         const double* d_sigma, const double* d_thickness, const double* d_altitude, double* d_out,
                         double* d_J, int precision, void* stream)
{
    if (B <= 0) return 0;
    if (l_stride < 1 || l_stride > GBP_MAXL) return fail("l_stride must be in [1, 30]");
    TableCache* tc;
    if (get_tables(sys, &tc)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == GBP_PRECISION_F32)
        return launch_fdem<float, true>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, d_J, st);
    if (precision == GBP_PRECISION_F64)
        return launch_fdem<double, true>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, d_J, st);
    return fail("precision must be 32 or 64");
}

static int fdem_host(bool sens, const gbp_fdem_system* sys, int B, int l_stride, const int32_t* nlayers,
                     const double* sigma, const double* thickness, const double* altitude, double* out, double* J,
                     int precision, int device)
{
    if (B <= 0) return 0;
    for (int b = 0; b < B; ++b)
        if (nlayers[b] < 1 || nlayers[b] > l_stride) return fail("nlayers out of range");
    CK(cudaSetDevice(device));
    const int C = 2 * sys->n_freq;
    int32_t* d_nl = nullptr;
    double *d_s = nullptr, *d_t = nullptr, *d_a = nullptr, *d_o = nullptr, *d_J = nullptr;
    const size_t nm = (size_t)B * l_stride;
    DevScope ds;
    CK(ds.alloc(&d_nl, B * sizeof(int32_t)));
    CK(ds.alloc(&d_s, nm * sizeof(double)));
    CK(ds.alloc(&d_t, nm * sizeof(double)));
    CK(ds.alloc(&d_a, B * sizeof(double)));
    CK(ds.alloc(&d_o, (size_t)B * C * sizeof(double)));
    if (sens) CK(ds.alloc(&d_J, (size_t)B * C * l_stride * sizeof(double)));
    CK(cudaMemcpy(d_nl, nlayers, B * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_s, sigma, nm * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_t, thickness, nm * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_a, altitude, B * sizeof(double), cudaMemcpyHostToDevice));
    int rc = sens ? gbp_fdem_sensitivity(sys, B, l_stride, d_nl, d_s, d_t, d_a, d_o, d_J, precision, nullptr)
                  : gbp_fdem_forward(sys, B, l_stride, d_nl, d_s, d_t, d_a, d_o, precision, nullptr);
    if (!rc) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(std::string("kernel: ") + cudaGetErrorString(e));
    }
    if (!rc) {
        CK(cudaMemcpy(out, d_o, (size_t)B * C * sizeof(double), cudaMemcpyDeviceToHost));
        if (sens) CK(cudaMemcpy(J, d_J, (size_t)B * C * l_stride * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return rc;
}

int gbp_fdem_forward_host(const gbp_fdem_system* sys, int B, int l_stride, const int32_t* nlayers, const double* sigma,
                          const double* thickness, const double* altitude, double* out, int precision, int device)
{
    return fdem_host(false, sys, B, l_stride, nlayers, sigma, thickness, altitude, out, nullptr, precision, device);
}

int gbp_fdem_sensitivity_host(const gbp_fdem_system* sys, int B, int l_stride, const int32_t* nlayers,
                              const double* sigma, const double* thickness, const double* altitude, double* out,
                              double* J, int precision, int device)
{
    return fdem_host(true, sys, B, l_stride, nlayers, sigma, thickness, altitude, out, J, precision, device);
}

int gbp_rjmcmc_run(const gbp_fdem_system* sys, const gbp_options* opt, int B, const double* d_data,
                   const double* d_altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                   const gbp_chain_buffers* d_buf, int precision, void* stream)
{
    if (B <= 0) return 0;
    if (check_options(opt)) return 1;
    if (!d_buf || !d_buf->scalars) return fail("gbp_chain_buffers.scalars is required");
    TableCache* tc;
    if (get_tables(sys, &tc)) return 1;
    ChainParams P;
    std::memset(&P, 0, sizeof(P));
    P.opt = *opt;
    P.B = B;
    P.n_depth = gbp_n_depth(opt);
    P.C = 2 * sys->n_freq;
    P.data = d_data;
    P.altitude = d_altitude;
    P.seed = seed;
    P.first_index = first_index;
    P.max_iterations = max_iterations;
    P.out = *d_buf;
    P.data_scale = 1.0;
    if (opt->n_systems > 1) return fail("an FDEM datapoint has one system (gbp_options.n_systems must be 0 or 1)");
    cudaStream_t st = (cudaStream_t)stream;
    const bool small = P.C <= 12;
    const SysDev& sd = tc->host.dev;
    if (opt->solve_height) {  // sampled sensor height: its own kernels (the fixed-height ones stay what they were)
        if (!small) return fail("solve_height: systems with more than 6 frequencies are not built");
        if (precision == GBP_PRECISION_F32)
            return launch_chain<float, float, 12, 16, KIND_FDEM_Z>(sd, tc->d_f32, (size_t)fdem_table_bytes<float>(sd), P, st);
        if (precision == GBP_PRECISION_F64)
            return launch_chain<double, double, 12, 8, KIND_FDEM_Z>(sd, tc->d_f64, (size_t)fdem_table_bytes<double>(sd), P, st);
        return fail("precision must be 32 or 64");
    }
    // fp32: 16 chains per SM; fp64: 8 per SM
    if (precision == GBP_PRECISION_F32) {
        const size_t tb = (size_t)fdem_table_bytes<float>(sd);
        // resident chains per SM: 16 (128 registers).  The SM's throughput saturates at ~16 warps (full waves of
        // equal-length chains: 16 -> 34.7 M, 28 -> 36.6 M evals/s), and fewer, faster chains shorten the tail of real
        // batches: 4096 soundings to termination 16: 1936 ms, 20: 2019, 24: 2086, 28: 2163 (profiles/README.md);
        // 8192 and 16384 soundings: equal (measured with the 20 / 24 / 28-warp builds of round 1, since removed).
        return small ? launch_chain<float, float, 12, 16, KIND_FDEM>(sd, tc->d_f32, tb, P, st)
                     : launch_chain<float, float, GBP_MAXC, 16, KIND_FDEM>(sd, tc->d_f32, tb, P, st);
    }
    if (precision == GBP_PRECISION_F64) {
        const size_t tb = (size_t)fdem_table_bytes<double>(sd);
        return small ? launch_chain<double, double, 12, 8, KIND_FDEM>(sd, tc->d_f64, tb, P, st)
                     : launch_chain<double, double, GBP_MAXC, 8, KIND_FDEM>(sd, tc->d_f64, tb, P, st);
    }
    return fail("precision must be 32 or 64");
}

}  // extern "C"

// Device buffers of the host-pointer sampler entry points are kept between calls (per device, one slot per result
// array): allocating and freeing 2.5 GB of posterior arrays cost more per call than copying them.  Two arenas per device,
// each with its own stream: two host threads can have a call in flight at once, and the second call's chains start on
// the SMs the first call's tail has already left while the first call's results are still on their way to the host.
struct HostArena {
    void* ptr[16];
    size_t cap[16];
    cudaStream_t stream;
    std::mutex mu;
};
constexpr int N_ARENAS = 2;
static HostArena g_arena[MAX_DEV][N_ARENAS];
static int arena_get(HostArena& a, int slot, size_t bytes, void** out)
{
    if (a.cap[slot] < bytes) {
        if (a.ptr[slot]) cudaFree(a.ptr[slot]);
        a.ptr[slot] = nullptr;
        a.cap[slot] = 0;
        CK(cudaMalloc(&a.ptr[slot], bytes));
        a.cap[slot] = bytes;
    }
    *out = a.ptr[slot];
    return 0;
}

// shared body of the *_rjmcmc_run_host entry points; `run` launches on the device buffers
template <typename RunFn>
static int rjmcmc_host_impl(int C, const gbp_options* opt, int B, const double* data, const double* altitude,
                            const gbp_chain_buffers* h, int device, RunFn run)
{
    if (B <= 0) return 0;
    if (check_options(opt)) return 1;
    if (!h || !h->scalars) return fail("gbp_chain_buffers.scalars is required");
    CK(cudaSetDevice(device));
    const int nd = gbp_n_depth(opt), ml = opt->max_layers, nsys = opt->n_systems > 1 ? 2 : 1;
    const size_t N2 = 2 * (size_t)opt->n_markov_chains;
    struct Item {
        void* const* host;
        void** dev;
        size_t bytes;
    };
    gbp_chain_buffers d;
    std::memset(&d, 0, sizeof(d));
    const Item items[] = {
        {(void* const*)&h->hitmap, (void**)&d.hitmap, (size_t)B * opt->n_sigma_bins * nd * sizeof(int32_t)},
        {(void* const*)&h->edges_hist, (void**)&d.edges_hist, (size_t)B * nd * sizeof(int32_t)},
        {(void* const*)&h->ncells_hist, (void**)&d.ncells_hist, (size_t)B * (ml + 1) * sizeof(int32_t)},
        {(void* const*)&h->rel_hist, (void**)&d.rel_hist, (size_t)B * nsys * opt->n_err_bins * sizeof(int32_t)},
        {(void* const*)&h->add_hist, (void**)&d.add_hist, (size_t)B * nsys * opt->n_err_bins * sizeof(int32_t)},
        {(void* const*)&h->misfit_trace, (void**)&d.misfit_trace, (size_t)B * N2 * sizeof(double)},
        {(void* const*)&h->accept_trace, (void**)&d.accept_trace, (size_t)B * N2},
        {(void* const*)&h->best_sigma, (void**)&d.best_sigma, (size_t)B * ml * sizeof(double)},
        {(void* const*)&h->best_edges, (void**)&d.best_edges, (size_t)B * (ml + 1) * sizeof(double)},
        {(void* const*)&h->cur_sigma, (void**)&d.cur_sigma, (size_t)B * ml * sizeof(double)},
        {(void* const*)&h->cur_edges, (void**)&d.cur_edges, (size_t)B * (ml + 1) * sizeof(double)},
        {(void* const*)&h->scalars, (void**)&d.scalars, (size_t)B * GBP_NSCALARS * sizeof(double)},
        {(void* const*)&h->height_hist, (void**)&d.height_hist, (size_t)B * opt->n_err_bins * sizeof(int32_t)},
    };
    if (device < 0 || device >= MAX_DEV) return fail("device index out of range");
    // the first free arena of the device (a third concurrent caller waits for the second one)
    HostArena* A = &g_arena[device][0];
    std::unique_lock<std::mutex> lk(A->mu, std::try_to_lock);
    if (!lk.owns_lock()) {
        A = &g_arena[device][1];
        lk = std::unique_lock<std::mutex>(A->mu);
    }
    if (!A->stream) CK(cudaStreamCreateWithFlags(&A->stream, cudaStreamNonBlocking));
    cudaStream_t st = A->stream;
    double *d_data = nullptr, *d_alt = nullptr;
    int rc = 0, slot = 0;
    for (const Item& it : items) {
        if (*it.host) {
            if (arena_get(*A, slot, it.bytes, it.dev)) return 1;
            CK(cudaMemsetAsync(*it.dev, 0, it.bytes, st));
        }
        ++slot;
    }
    if (arena_get(*A, 14, (size_t)B * C * sizeof(double), (void**)&d_data)) return 1;
    if (arena_get(*A, 15, (size_t)B * sizeof(double), (void**)&d_alt)) return 1;
    CK(cudaMemcpyAsync(d_data, data, (size_t)B * C * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_alt, altitude, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = run(d_data, d_alt, &d, (void*)st);
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(std::string("kernel: ") + cudaGetErrorString(e));
    }
    if (!rc) {
        for (const Item& it : items)
            if (*it.host) CK(cudaMemcpyAsync(*it.host, *it.dev, it.bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return rc;
}

extern "C" {

int gbp_rjmcmc_run_host(const gbp_fdem_system* sys, const gbp_options* opt, int B, const double* data,
                        const double* altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                        const gbp_chain_buffers* h, int precision, int device)
{
    return rjmcmc_host_impl(2 * sys->n_freq, opt, B, data, altitude, h, device,
                            [&](const double* d_data, const double* d_alt, const gbp_chain_buffers* d, void* st) {
                                return gbp_rjmcmc_run(sys, opt, B, d_data, d_alt, seed, first_index, max_iterations, d,
                                                      precision, st);
                            });
}

static int summarise_impl(const int32_t* d_hitmap, int B, int n_sig, int n_depth, const double* d_sig_lo, double dx,
                          const double* percentiles, int n_pct, double credible_percent, double* d_mean, double* d_pct,
                          double* d_mode, int32_t* d_range_bins, void* stream)
{
    if (B <= 0) return 0;
    const int want_range = d_range_bins != nullptr;
    const int np_total = n_pct + (want_range ? 2 : 0);
    if (n_pct < (want_range ? 0 : 1) || np_total < 1 || np_total > SUMM_MAXP) return fail("1 to 8 percentiles (a credible range takes two of them)");
    if (n_sig < 1 || n_depth < 1) return fail("invalid hitmap shape");
    if (want_range && !(credible_percent > 0.0 && credible_percent < 100.0)) return fail("credible_percent must be in (0, 100)");
    SummParams P;
    P.B = B;
    P.n_sig = n_sig;
    P.n_depth = n_depth;
    P.n_pct = n_pct;
    P.dx = dx;
    P.want_range = want_range;
    for (int q = 0; q < SUMM_MAXP; ++q) P.pct[q] = q < n_pct ? percentiles[q] : 0.0;
    if (want_range) {   // Mesh._credible_range :74-75: percent = 0.5 * min(percent, 100 - percent); bounds percent, 100 - percent
        const double h = 0.5 * (credible_percent < 100.0 - credible_percent ? credible_percent : 100.0 - credible_percent);
        P.pct[n_pct] = h;
        P.pct[n_pct + 1] = 100.0 - h;
    }
    const long long n = (long long)B * n_depth;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    int dev;
    if (current_device(&dev)) return 1;
    std::lock_guard<std::mutex> lk(g_dev[dev].mu);
    if (time_begin(dev, st)) return 1;
    switch (np_total) {
#define GBP_SUMM_CASE(NP) case NP: summarise_kernel<NP><<<(int)blocks, 256, 0, st>>>(d_hitmap, d_sig_lo, P, d_mean, d_pct, d_mode, d_range_bins); break;
        GBP_SUMM_CASE(1) GBP_SUMM_CASE(2) GBP_SUMM_CASE(3) GBP_SUMM_CASE(4) GBP_SUMM_CASE(5) GBP_SUMM_CASE(6) GBP_SUMM_CASE(7)
        GBP_SUMM_CASE(8)
#undef GBP_SUMM_CASE
        default: break;
    }
    g_launches++;
    CK(cudaGetLastError());
    return time_end(dev, st);
}

int gbp_summarise_hitmap(const int32_t* d_hitmap, int B, int n_sig, int n_depth, const double* d_sig_lo, double dx,
                         const double* percentiles, int n_pct, double* d_mean, double* d_pct, void* stream)
{
    return summarise_impl(d_hitmap, B, n_sig, n_depth, d_sig_lo, dx, percentiles, n_pct, 0.0, d_mean, d_pct, nullptr, nullptr, stream);
}

int gbp_summarise_posterior(const int32_t* d_hitmap, int B, int n_sig, int n_depth, const double* d_sig_lo, double dx,
                            const double* percentiles, int n_pct, double credible_percent, double* d_mean, double* d_pct,
                            double* d_mode, int32_t* d_range_bins, void* stream)
{
    return summarise_impl(d_hitmap, B, n_sig, n_depth, d_sig_lo, dx, percentiles, n_pct, credible_percent, d_mean, d_pct, d_mode,
                          d_range_bins, stream);
}

int gbp_opacity_doi(const int32_t* d_range_bins, int B, int n_depth, const int32_t* d_group, int n_groups, double doi_percent,
                    double level_percent, int32_t* d_group_minmax, double* d_opacity, int32_t* d_doi_cell, int32_t* d_level_cell,
                    void* stream)
{
    if (B <= 0) return 0;
    if (n_depth < 1) return fail("invalid shape");
    if (!d_group_minmax) return fail("gbp_opacity_doi: d_group_minmax (2 x n_groups int32 of scratch) is required");
    const int ng = d_group ? n_groups : B;
    if (ng < 1) return fail("n_groups must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    int dev;
    if (current_device(&dev)) return 1;
    std::lock_guard<std::mutex> lk(g_dev[dev].mu);
    // min <- 0x7f7f7f7f, max <- 0x80808080 (byte fills): above / below every possible bin difference
    CK(cudaMemsetAsync(d_group_minmax, 0x7f, (size_t)ng * sizeof(int32_t), st));
    CK(cudaMemsetAsync(d_group_minmax + ng, 0x80, (size_t)ng * sizeof(int32_t), st));
    const int blocks = (B * 32 + 255) / 256;
    if (time_begin(dev, st)) return 1;
    range_minmax_kernel<<<blocks, 256, 0, st>>>(d_range_bins, d_group, B, n_depth, d_group_minmax, d_group_minmax + ng);
    opacity_doi_kernel<<<blocks, 256, 0, st>>>(d_range_bins, d_group, B, n_depth, d_group_minmax, d_group_minmax + ng,
                                               0.01 * doi_percent, 0.01 * level_percent, d_opacity, d_doi_cell, d_level_cell);
    g_launches += 2;
    CK(cudaGetLastError());
    return time_end(dev, st);
}

int gbp_release_host_buffers(void)
{
    for (int dev = 0; dev < MAX_DEV; ++dev)
        for (int a = 0; a < N_ARENAS; ++a) {
            HostArena& A = g_arena[dev][a];
            std::lock_guard<std::mutex> lk(A.mu);
            for (int s = 0; s < 16; ++s)
                if (A.ptr[s]) {
                    cudaSetDevice(dev);
                    cudaFree(A.ptr[s]);
                    A.ptr[s] = nullptr;
                    A.cap[s] = 0;
                }
        }
    return 0;
}

int gbp_debug_finish_times(double* out_ms, int n)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (!g_finish[dev] || n > g_finish_B[dev]) return fail("no timeline recorded (set GBP_DEBUG_TIMELINE before the run)");
    std::vector<unsigned long long> h((size_t)g_finish_B[dev] + 1);
    CK(cudaMemcpy(h.data(), g_finish[dev], h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const unsigned long long t0 = h[g_finish_B[dev]];
    for (int i = 0; i < n; ++i) out_ms[i] = (double)(h[i] - t0) * 1e-6;
    return 0;
}

int gbp_debug_progress_times(double* out_ms, int n)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (!g_finish[dev] || n > g_finish_B[dev]) return fail("no timeline recorded (set GBP_DEBUG_TIMELINE before the run)");
    const size_t B = (size_t)g_finish_B[dev];
    std::vector<unsigned long long> h(B + 1 + 32 * B);
    CK(cudaMemcpy(h.data(), g_finish[dev], h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const unsigned long long t0 = h[B];
    for (size_t i = 0; i < (size_t)n * 32; ++i) {
        const unsigned long long v = h[B + 1 + i];
        out_ms[i] = (v == ~0ull) ? -1.0 : (double)(v - t0) * 1e-6;
    }
    return 0;
}

int gbp_debug_counters(unsigned long long* out16, int reset)
{
    if (out16) CK(cudaMemcpyFromSymbol(out16, g_diag, 16 * sizeof(unsigned long long)));
    if (reset) {
        unsigned long long z[16];
        std::memset(z, 0, sizeof(z));
        CK(cudaMemcpyToSymbol(g_diag, z, sizeof(z)));
    }
    return 0;
}

int gbp_measure_peaks(double* fp32_tflops, double* mufu_gops)
{
    const int blocks = sm_count() * 8, threads = 256, iters = 4096;
    float* d = nullptr;
    DevScope ds;
    CK(ds.alloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    CK(ds.event(&e0));
    CK(ds.event(&e1));
    double best_f = 0.0, best_m = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        float ms = 0.f;
        CK(cudaEventRecord(e0));
        peak_fma_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double f = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep && f > best_f) best_f = f;
        CK(cudaEventRecord(e0));
        peak_mufu_kernel<<<blocks, threads>>>(d, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double m = 32.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e9;
        if (rep && m > best_m) best_m = m;
        g_launches += 2;
    }
    CK(cudaDeviceSynchronize());
    if (fp32_tflops) *fp32_tflops = best_f;
    if (mufu_gops) *mufu_gops = best_m;
    return 0;
}

// MUFU-class (special function unit) operations of one forward: per abscissa 3 (basement sqrt, sqrt, rcp) + 9 per
// finite layer (3 complex sqrt, 1 clamp rcp, 3 exp/sin/cos, 2 complex reciprocals) + 4 (reflection coefficient, height term)
double gbp_mufu_per_forward(const gbp_fdem_system* sys, int L)
{
    return (double)gbp_filter_points(sys) * (9.0 * (double)L - 2.0);
}
double gbp_tdem_mufu_per_forward(const gbp_tdem_survey* sv, int L)
{
    TdCache* tc;
    if (get_td_tables(sv, &tc, false)) return 0.0;
    // per (frequency, abscissa): basement 3 + 1, per finite layer 3 + 1 + 1 + 3 + 1, top 1
    return (double)TD_NF * tc->host.dev.n_lam * (9.0 * (double)L - 4.0);
}

// ------------------------------------------------------------------------------------------ time domain
int gbp_tdem_n_channels(const gbp_tdem_survey* sv)
{
    return td_n_channels(*sv);
}

int gbp_tdem_primary_field(const gbp_tdem_survey* sv, double* out)
{
    // B of a vertical dipole of moment m = PeakCurrent x (turns x area = 1) at the receiver, z up:
    //   Bx = mu0 m / (4 pi) 3 x z / R^5,  Bz = mu0 m / (4 pi) (3 z^2 / R^5 - 1 / R^3); the reference stacks PX, -PZ
    if (!sv || !out) return -1;
    const double x = sv->rx_dx, y = sv->rx_dy, z = sv->rx_dz, R = std::sqrt(x * x + y * y + z * z);
    if (!(R > 0.0)) {
        fail("primary field: the receiver sits on the transmitter");
        return -1;
    }
    int n = 0;
    for (int s = 0; s < sv->n_systems && s < GBP_TD_MAXSYS; ++s) {
        const TdOutput o = td_output(sv->sys[s]);
        const double sx = sv->sys[s].x_scaling, sz = (sv->sys[s].x_scaling == 0.0 && sv->sys[s].z_scaling == 0.0) ? 1.0 : sv->sys[s].z_scaling;
        for (int q = 0; q < o.n_comp; ++q)
            out[n++] = o.comp[q] == 1 ? 1e-7 * o.peak * 3.0 * x * z / std::pow(R, 5) * sx
                                      : -1e-7 * o.peak * (3.0 * z * z / std::pow(R, 5) - 1.0 / std::pow(R, 3)) * sz;
    }
    return n;
}

int gbp_tdem_window_operator(const gbp_tdem_survey* sv, double* freq, double* MR, double* MI, double* t_centre)
{
    TdCache* tc;
    if (get_td_tables(sv, &tc, false)) return 1;
    const int C = tc->host.dev.C;
    if (freq) std::memcpy(freq, tc->host.freq, sizeof(double) * TD_NF);
    for (int c = 0; c < C; ++c)
        for (int i = 0; i < TD_NF; ++i) {
            if (MR) MR[c * TD_NF + i] = tc->host.Mt[(size_t)i * TD_CP + c];
            if (MI) MI[c * TD_NF + i] = tc->host.Mt[(size_t)(TD_NF + i) * TD_CP + c];
        }
    if (t_centre) {
        int c = 0;
        for (int s = 0; s < sv->n_systems; ++s)
            for (int i = 0; i < sv->sys[s].n_windows; ++i, ++c)
                t_centre[c] = 0.5 * (sv->sys[s].window_start[i] + sv->sys[s].window_end[i]);
    }
    return 0;
}

double gbp_tdem_flops_per_forward(const gbp_tdem_survey* sv, int L)
{
    TdCache* tc;
    if (get_td_tables(sv, &tc, false)) return 0.0;
    return (double)TD_NF * tc->host.dev.n_lam * (75.0 * (double)L + 39.0) + 2.0 * tc->host.dev.C * TD_ROWS;
}

static int tdem_dev(bool sens, const gbp_tdem_survey* sv, int B, int l_stride, const int32_t* d_nlayers,
                    const double* d_sigma, const double* d_thickness, const double* d_altitude, double* d_out, double* d_J,
                    int precision, void* stream)
{
    if (B <= 0) return 0;
    if (l_stride < 1 || l_stride > GBP_MAXL) return fail("l_stride must be in [1, 30]");
    TdCache* tc;
    if (get_td_tables(sv, &tc, true)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == GBP_PRECISION_F32)
        return sens ? launch_tdem<float, true>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, d_J, 1.0 / TD_F32_SCALE, st)
                    : launch_tdem<float, false>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, nullptr, 1.0 / TD_F32_SCALE, st);
    if (precision == GBP_PRECISION_F64)
        return sens ? launch_tdem<double, true>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, d_J, 1.0, st)
                    : launch_tdem<double, false>(tc, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, nullptr, 1.0, st);
    return fail("precision must be 32 or 64");
}

int gbp_tdem_forward(const gbp_tdem_survey* sv, int B, int l_stride, const int32_t* d_nlayers, const double* d_sigma,
                     const double* d_thickness, const double* d_altitude, double* d_out, int precision, void* stream)
{
    return tdem_dev(false, sv, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, nullptr, precision, stream);
}

int gbp_tdem_sensitivity(const gbp_tdem_survey* sv, int B, int l_stride, const int32_t* d_nlayers, const double* d_sigma,
                         const double* d_thickness, const double* d_altitude, double* d_out, double* d_J, int precision,
                         void* stream)
{
    return tdem_dev(true, sv, B, l_stride, d_nlayers, d_sigma, d_thickness, d_altitude, d_out, d_J, precision, stream);
}

static int tdem_host(bool sens, const gbp_tdem_survey* sv, int B, int l_stride, const int32_t* nlayers,
                     const double* sigma, const double* thickness, const double* altitude, double* out, double* J,
                     int precision, int device)
{
    if (B <= 0) return 0;
    for (int b = 0; b < B; ++b) {
        if (nlayers[b] < 1 || nlayers[b] > l_stride) return fail("nlayers out of range");
        if (!(altitude[b] > 0.0)) return fail("Sensor altitude must be above the top of the model");  // tdem1d.py:32
    }
    CK(cudaSetDevice(device));
    const int C = gbp_tdem_n_channels(sv);
    int32_t* d_nl = nullptr;
    double *d_s = nullptr, *d_t = nullptr, *d_a = nullptr, *d_o = nullptr, *d_J = nullptr;
    const size_t nm = (size_t)B * l_stride;
    DevScope ds;
    CK(ds.alloc(&d_nl, B * sizeof(int32_t)));
    CK(ds.alloc(&d_s, nm * sizeof(double)));
    CK(ds.alloc(&d_t, nm * sizeof(double)));
    CK(ds.alloc(&d_a, B * sizeof(double)));
    CK(ds.alloc(&d_o, (size_t)B * C * sizeof(double)));
    if (sens) CK(ds.alloc(&d_J, (size_t)B * C * l_stride * sizeof(double)));
    CK(cudaMemcpy(d_nl, nlayers, B * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_s, sigma, nm * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_t, thickness, nm * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_a, altitude, B * sizeof(double), cudaMemcpyHostToDevice));
    int rc = tdem_dev(sens, sv, B, l_stride, d_nl, d_s, d_t, d_a, d_o, d_J, precision, nullptr);
    if (!rc) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(std::string("kernel: ") + cudaGetErrorString(e));
    }
    if (!rc) {
        CK(cudaMemcpy(out, d_o, (size_t)B * C * sizeof(double), cudaMemcpyDeviceToHost));
        if (sens) CK(cudaMemcpy(J, d_J, (size_t)B * C * l_stride * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return rc;
}

int gbp_tdem_forward_host(const gbp_tdem_survey* sv, int B, int l_stride, const int32_t* nlayers, const double* sigma,
                          const double* thickness, const double* altitude, double* out, int precision, int device)
{
    return tdem_host(false, sv, B, l_stride, nlayers, sigma, thickness, altitude, out, nullptr, precision, device);
}

int gbp_tdem_sensitivity_host(const gbp_tdem_survey* sv, int B, int l_stride, const int32_t* nlayers, const double* sigma,
                              const double* thickness, const double* altitude, double* out, double* J, int precision,
                              int device)
{
    return tdem_host(true, sv, B, l_stride, nlayers, sigma, thickness, altitude, out, J, precision, device);
}

int gbp_tdem_rjmcmc_run(const gbp_tdem_survey* sv, const gbp_options* opt, int B, const double* d_data,
                        const double* d_altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                        const gbp_chain_buffers* d_buf, int precision, void* stream)
{
    if (B <= 0) return 0;
    if (check_options(opt)) return 1;
    if (!d_buf || !d_buf->scalars) return fail("gbp_chain_buffers.scalars is required");
    const bool tempest = sv->error_model == 1;
    if (!tempest && (opt->n_systems > 1 ? 2 : 1) != sv->n_systems)
        return fail("gbp_options.n_systems must equal the number of systems of the datapoint type");
    // the sampler kernels hold GBP_TD_SAMPLER_MAXC channels per chain in shared memory (the forward / Jacobian
    // operators take GBP_TD_MAXC); checked before any device work
    if (gbp_tdem_n_channels(sv) > GBP_TD_SAMPLER_MAXC)
        return fail("time-domain sampler: at most 48 data channels per datapoint (GBP_TD_SAMPLER_MAXC)");
    for (int s = 0; s < sv->n_systems && s < GBP_TD_MAXSYS; ++s) {
        const TdOutput o = td_output(sv->sys[s]);
        if (!tempest && (o.n_comp != 1 || o.comp[0] != 0 || o.b_field))
            return fail("time-domain sampler: a system that measures X or reports B needs the Tempest error model "
                        "(gbp_tdem_survey.error_model = 1 with the additive level of every channel)");
        if (tempest && (opt->n_systems > 1 ? 2 : 1) != o.n_comp)
            return fail("Tempest error model: gbp_options.n_systems must equal the number of measured components");
    }
    if (tempest && opt->solve_height) return fail("Tempest error model: the transmitter height is not sampled (solve_height)");
    TdCache* tc;
    if (get_td_tables(sv, &tc, true)) return 1;
    ChainParams P;
    std::memset(&P, 0, sizeof(P));
    P.opt = *opt;
    P.B = B;
    P.n_depth = gbp_n_depth(opt);
    P.C = tc->host.dev.C;
    P.data = d_data;
    P.altitude = d_altitude;
    P.seed = seed;
    P.first_index = first_index;
    P.max_iterations = max_iterations;
    P.out = *d_buf;
    P.data_scale = 1.0;
    cudaStream_t st = (cudaStream_t)stream;
    const TdDev& sd = tc->host.dev;
    if (tempest) {
        // the additive-error unknowns are dimensionless multipliers: in the scaled data units of the fp32 kernel it is the
        // additive LEVELS and the primary field that scale with the data
        TdDev sq = sd;
        if (precision == GBP_PRECISION_F32) {
            P.data_scale = TD_F32_SCALE;
            for (int c = 0; c < GBP_TD_MAXC; ++c) {
                sq.tsc[c] *= TD_F32_SCALE;
                sq.poff[c] *= TD_F32_SCALE;
            }
            return launch_chain<float, float, GBP_TD_SAMPLER_MAXC, 12, KIND_TEMPEST>(sq, tc->d_f32, (size_t)TD_ROWS * TD_CP * sizeof(float), P, st);
        }
        if (precision == GBP_PRECISION_F64)
            return launch_chain<double, double, GBP_TD_SAMPLER_MAXC, 8, KIND_TEMPEST>(sq, tc->d_f64, (size_t)TD_ROWS * TD_CP * sizeof(double), P, st);
        return fail("precision must be GBP_PRECISION_F32 or GBP_PRECISION_F64");
    }
    if (precision == GBP_PRECISION_F32) {
        // scaled data units: additive errors scale with the data
        P.data_scale = TD_F32_SCALE;
        P.opt.add_init *= TD_F32_SCALE;
        P.opt.add_min *= TD_F32_SCALE;
        P.opt.add_max *= TD_F32_SCALE;
        P.opt.add_init2 *= TD_F32_SCALE;
        P.opt.add_min2 *= TD_F32_SCALE;
        P.opt.add_max2 *= TD_F32_SCALE;
        // 12 chains per SM (to termination, scripts/gpu_tdem_warps.py: 4096 soundings 12: 3157 ms, 16: 3437, 8: 3656;
        // 8192 soundings 12: 5312 ms, 16: 5505, 8: 6243; 18 chains per SM at 113 registers spill and are slower still)
        // solve_height (the options file's solve_transmitter_z): its own instantiation, the fixed-height kernel stays what it was
        if (opt->solve_height)
            return launch_chain<float, float, GBP_TD_SAMPLER_MAXC, 12, KIND_TDEM_Z>(sd, tc->d_f32, (size_t)TD_ROWS * TD_CP * sizeof(float), P, st);
        return launch_chain<float, float, GBP_TD_SAMPLER_MAXC, 12, KIND_TDEM>(sd, tc->d_f32, (size_t)TD_ROWS * TD_CP * sizeof(float), P, st);
    }
    if (precision == GBP_PRECISION_F64) {
        if (opt->solve_height)
            return launch_chain<double, double, GBP_TD_SAMPLER_MAXC, 8, KIND_TDEM_Z>(sd, tc->d_f64, (size_t)TD_ROWS * TD_CP * sizeof(double), P, st);
        return launch_chain<double, double, GBP_TD_SAMPLER_MAXC, 8, KIND_TDEM>(sd, tc->d_f64, (size_t)TD_ROWS * TD_CP * sizeof(double), P, st);
    }
    return fail("precision must be 32 or 64");
}

int gbp_tdem_rjmcmc_run_host(const gbp_tdem_survey* sv, const gbp_options* opt, int B, const double* data,
                             const double* altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                             const gbp_chain_buffers* h, int precision, int device)
{
    for (int b = 0; b < B; ++b)
        if (!(altitude[b] > 0.0)) return fail("Sensor altitude must be above the top of the model");
    return rjmcmc_host_impl(gbp_tdem_n_channels(sv), opt, B, data, altitude, h, device,
                            [&](const double* d_data, const double* d_alt, const gbp_chain_buffers* d, void* st) {
                                return gbp_tdem_rjmcmc_run(sv, opt, B, d_data, d_alt, seed, first_index, max_iterations, d,
                                                           precision, st);
                            });
}

}  // extern "C"
