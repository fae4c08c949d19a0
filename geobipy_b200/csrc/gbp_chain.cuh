// One warp = one Markov chain: the fused trans-dimensional MCMC sampler for one FDEM sounding.
//
// Replaces the Python loop Inference1D.initialize + infer (geobipy/src/inversion/Inference1D.py:353-464,
// :537-631 accept_reject, :633-688 infer, :705-790 update) together with everything it calls:
//   RectilinearMesh1D.perturb            classes/mesh/RectilinearMesh1D.py:993-1120
//   Model.stochastic_newton_perturbation classes/model/Model.py:368-419 (+ :250-272, :347-357, :421-430)
//   Model.probability / gradient_probability / proposal_probabilities  Model.py:533-575, :213-234, :577-660
//   DataPoint.std / data_misfit / likelihood / probability / perturb   classes/data/datapoint/DataPoint.py
//   Model.update_parameter_posterior, RectilinearMesh1D.update_posteriors, EmDataPoint.update_posteriors
//
// Design
//  * chain state lives in shared memory (one WarpState per warp, double-buffered current/proposed model)
//    and registers; nothing but the posterior arrays and traces ever touches HBM;
//  * control flow is warp-uniform: every lane draws the same Philox numbers, lanes split the array work
//    (remapping, width checks, Hessian rows, Cholesky rows, depth cells);
//  * the k x k Gauss-Newton system is factorised in-warp (packed Cholesky, lane = row) instead of the
//    reference's inv() + SVD;
//  * logs are carried, not recomputed: ln(sigma), ln(thickness), ln(errors) are state;
//  * posterior histograms are updated with a dwell count, i.e. only when the model changes (accept):
//    identical counts, ~1/acceptance of the HBM read-modify-writes;
//  * R = arithmetic type of the sampler.  R = double: trajectory twin of the CPU oracle (validation).
//    R = float: production path (forward/Jacobian AND statistics in fp32, MUFU log/exp); counters and
//    histograms are integers in both.
//  * the code is split into one non-inlined function per phase so that the kernel stays small: the first
//    version (everything inlined, 532 KB of SASS) was bound by instruction-cache misses.
#pragma once
#include "gbp_fdem.cuh"

namespace gbp {

constexpr int NPACK = GBP_MAXL * (GBP_MAXL + 1) / 2;
enum { ACT_BIRTH = 0, ACT_DEATH = 1, ACT_MOVE = 2, ACT_NONE = 3 };

template <typename R> struct MeshBuf {
    R edges[GBP_MAXL + 2];  // edges[0] = 0, edges[k] = inf
    R lnh[GBP_MAXL];        // ln(thickness_i)           (gradient prior, RectilinearMesh1D.py:713)
    R t2[GBP_MAXL];         // gradient-operator weights (RectilinearMesh1D.py:747-786) / g^2
};
template <typename R> struct ValBuf {
    R sig[GBP_MAXL];        // conductivity
    R ls[GBP_MAXL];         // ln(conductivity)
};

template <typename R, typename T, int NC> struct __align__(16) WarpState {
    R A[NPACK];             // packed lower triangle: Gauss-Newton matrix, then its Cholesky factor
    MeshBuf<R> mesh[2];
    ValBuf<R> val[2];
    R ls_r[GBP_MAXL];       // ln(sigma) of the remapped model
    R vec[GBP_MAXL + 2];
    R data[NC], ivar[NC];   // observed data (0 where inactive), 1/variance (0 where inactive)
    T J[2][NC * KS];
    T pred[2][NC];
    T msig[KS], mthk[KS];
    int sbin[GBP_MAXL + 2];
};

// Option-derived constants, computed once per CTA in fp64 and shared by its warps.
template <typename R> struct Consts {
    R cum0, cum1, cum2;                       // cumulative event probabilities
    R ln_min_edge, ln_edge_span, min_edge, max_edge, min_width;
    R inv_s2, inv_g2, alpha;                  // 1/ln(1+factor)^2, 1/grad_std^2, covariance_scaling
    R lp_k;                                   // -ln(kmax - 1)
    R c_grad, c_val;                          // per-dimension constants of the gradient / value prior
    R half_log2pi;
    R rel_lnmin, rel_lnmax, rel_sd, rel_lp, rel_ln0;
    R add_lnmin, add_lnmax, add_sd, add_lp, add_ln0;
    R sig_halfspan, sig_dx, rel_dx, add_dx, depth_step, depth_max;
    R ln_half, ln_3half;
};

struct ChainParams {
    gbp_options opt;
    int B, n_depth, C, n_warps_total;
    const double* data;      // [B][C]
    const double* altitude;  // [B]
    unsigned long long seed, first_index;
    long long max_iterations;
    gbp_chain_buffers out;
    int* work_counter;
};

template <typename R> __device__ __noinline__ void make_consts(const gbp_options& o, int n_depth, Consts<R>& c)
{
    c.cum0 = (R)o.p_birth;
    c.cum1 = (R)(o.p_birth + o.p_death);
    c.cum2 = (R)(o.p_birth + o.p_death + o.p_move);
    c.ln_min_edge = (R)dlog_(o.min_edge);
    c.ln_edge_span = (R)(dlog_(o.max_edge) - dlog_(o.min_edge));
    c.min_edge = (R)o.min_edge;
    c.max_edge = (R)o.max_edge;
    c.min_width = (R)o.min_width;
    const double s = dlog_(1.0 + o.factor), g2 = o.gradient_std * o.gradient_std;
    c.inv_s2 = (R)(1.0 / (s * s));
    c.inv_g2 = (R)(1.0 / g2);
    c.alpha = (R)o.covariance_scaling;
    c.lp_k = (R)(-dlog_((double)o.max_layers - 1.0));
    const double l2pi = 1.8378770664093454835606594728112;
    c.c_grad = (R)(-0.5 * l2pi - 0.5 * dlog_(g2));
    c.c_val = (R)(-0.5 * l2pi - 0.5 * dlog_(s * s));
    c.half_log2pi = (R)(0.5 * l2pi);
    c.rel_lnmin = (R)dlog_(o.rel_min);
    c.rel_lnmax = (R)dlog_(o.rel_max);
    c.rel_sd = (R)::sqrt(o.rel_prop_var);
    c.rel_lp = (R)(-dlog_(dlog_(o.rel_max) - dlog_(o.rel_min)));
    c.rel_ln0 = (R)dlog_(o.rel_init);
    c.add_lnmin = (R)dlog_(o.add_min);
    c.add_lnmax = (R)dlog_(o.add_max);
    c.add_sd = (R)::sqrt(o.add_prop_var);
    c.add_lp = (R)(-dlog_(dlog_(o.add_max) - dlog_(o.add_min)));
    c.add_ln0 = (R)dlog_(o.add_init);
    c.sig_halfspan = (R)(o.sigma_bins_nstd * s);
    c.sig_dx = (R)(2.0 * o.sigma_bins_nstd * s / (double)o.n_sigma_bins);
    c.rel_dx = (R)((dlog_(o.rel_max) - dlog_(o.rel_min)) / (double)o.n_err_bins);
    c.add_dx = (R)((dlog_(o.add_max) - dlog_(o.add_min)) / (double)o.n_err_bins);
    c.depth_step = (R)(0.5 * o.min_width);
    c.depth_max = (R)((double)n_depth * 0.5 * o.min_width);
    c.ln_half = (R)dlog_(0.5);
    c.ln_3half = (R)dlog_(1.5);
}

__device__ __forceinline__ int pk(int i, int j) { return i * (i + 1) / 2 + j; }

// searchsorted(edges, v, 'right') - 1 clipped, uniform edges lo + i*dx
template <typename R> __device__ __noinline__ int uniform_bin(R v, R lo, R dx, int n)
{
    R f = floor((v - lo) / dx);
    int i = (f < R(0)) ? 0 : (f > (R)(n - 1) ? n - 1 : (int)f);
#pragma unroll 1
    while (i + 1 < n && v >= lo + (R)(i + 1) * dx) ++i;
#pragma unroll 1
    while (i > 0 && v < lo + (R)i * dx) --i;
    return i;
}

template <typename R> __device__ __forceinline__ R warp_min(R v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// tell the compiler that the chain state and the constants live in shared memory (LDS/STS, not generic)
#define GBP_SHARED_STATE              \
    __builtin_assume(__isShared(&w)); \
    __builtin_assume(__isShared(&K))

template <typename R, typename T, int NC> struct Chain {
    typedef WarpState<R, T, NC> WS;
    WS& w;
    const Consts<R>& K;
    const SysShared<T>& S;
    const T* tab;
    const ChainParams& P;
    const int lane;
    const int C;
    // ---- per-chain scalars
    Rng rng;
    int chain;
    T alt;
    int k;                       // layers of the current model
    int mcur, vcur, jcur, pcur;  // which buffer holds the current mesh / values / Jacobian / predicted data
    R ln_rel, ln_add, rel, add, ln_ref;
    double sigma_ref;
    R misfit, prior, likelihood, posterior, best_posterior;
    R sig_lo;                    // lower edge of the conductivity bins (ln)
    int iteration, burned_in_iter, best_iter, n_accept, n_forward, n_sens;
    int n_act[4];
    int burned_in, n_zero, n_resets, limiters, n_active, acc_win, dwell, best_k;
    R best_rel, best_add;

    __device__ Chain(WS& w_, const Consts<R>& K_, const SysShared<T>& S_, const T* tab_, const ChainParams& P_)
        : w(w_), K(K_), S(S_), tab(tab_), P(P_), lane(threadIdx.x & 31), C(P_.C)
    {
    }

    // ------------------------------------------------------------ forward wrapper
    // J == nullptr: forward only.  Otherwise forward + Jacobian in one pass (FdemDataPoint.fm_dlogc :535)
    __device__ __noinline__ void forward(int kk, const R* sig, const R* edges, T* pred, T* J)
    {
        GBP_SHARED_STATE;
        if (lane < kk) {
            w.msig[lane] = (T)sig[lane];
            w.mthk[lane] = (T)(edges[lane + 1] - edges[lane]);
        }
        __syncwarp();
        n_forward++;
        if (J) n_sens++;
        fdem_eval<T>(S, tab, alt, kk, w.msig, w.mthk, pred, J, J != nullptr);
    }

    // ------------------------------------------------------------ data terms
    // DataPoint.std :268-282 -> 1/variance per active channel (EmDataPoint.active :44-56)
    __device__ __noinline__ void set_ivar(R r, R a)
    {
        GBP_SHARED_STATE;
        if (lane < C) {
            R d = w.data[lane];
            R s = r * d;
            w.ivar[lane] = (d > R(0)) ? R(1) / (s * s + a * a) : R(0);
        }
        __syncwarp();
    }
    // misfit (DataPoint.py:502-525) and Gaussian log-likelihood (MvNormalDistribution.py:209-216)
    __device__ __noinline__ void misfit_likelihood(const T* pred, R* mis, R* like)
    {
        GBP_SHARED_STATE;
        R q = R(0), ld = R(0);
        if (lane < C) {
            R iv = w.ivar[lane];
            if (iv > R(0)) {
                R r = (R)pred[lane] - w.data[lane];
                q = r * r * iv;
                ld = -rt<R>::log(iv);
            }
        }
        q = warp_sum(q);
        ld = warp_sum(ld);
        *mis = q;
        *like = -(R)n_active * K.half_log2pi - R(0.5) * ld - R(0.5) * q;
    }
    // Uniform(log=True) priors on the errors (DataPoint.probability :351-395), arguments in ln space
    __device__ __forceinline__ R datapoint_probability(R lr, R la)
    {
        R p = R(0);
        if (P.opt.solve_relative_error) p += (lr < K.rel_lnmin || lr > K.rel_lnmax) ? (R)-INFINITY : K.rel_lp;
        if (P.opt.solve_additive_error) p += (la < K.add_lnmin || la > K.add_lnmax) ? (R)-INFINITY : K.add_lp;
        return p;
    }
    // Model.probability :533-575 (value_bounds = None); ls = ln sigma, lnh = ln thickness
    __device__ __noinline__ R model_probability(int kk, const R* ls, const R* lnh)
    {
        GBP_SHARED_STATE;
        const gbp_options& o = P.opt;
        R p = (kk >= 1 && kk <= o.max_layers) ? K.lp_k : (R)-INFINITY;
        if (o.solve_parameter) {
            R q = R(0);
            if (lane < kk) {
                R d = ls[lane] - ln_ref;
                q = d * d * K.inv_s2;
            }
            p += (R)kk * K.c_val - R(0.5) * warp_sum(q);
        }
        if (o.solve_gradient) {
            if (kk == 1) {
                p += K.c_grad;  // Model.py:230-232: a virtual 2-layer model with equal values
            } else {
                R q = R(0);
                if (lane < kk - 1) {
                    R g = (ls[lane + 1] - ls[lane]) / lnh[lane];
                    q = g * g * K.inv_g2;
                }
                p += (R)(kk - 1) * K.c_grad - R(0.5) * warp_sum(q);
            }
        }
        return p;
    }

    // ------------------------------------------------------------ mesh-derived quantities
    // lnh[i] = ln(thickness_i), t2[i] = 1/(g^2 (c2c_i (k-1))^2)  (RectilinearMesh1D.gradient_operator :747-786)
    __device__ __noinline__ void mesh_setup(int kk, MeshBuf<R>& m)
    {
        GBP_SHARED_STATE;
        if (kk >= 2 && lane < kk - 1) {
            const R* e = m.edges;
            R x0 = e[lane + 1] - e[lane];
            R x1;
            if (lane + 1 < kk - 1) x1 = e[lane + 2] - e[lane + 1];
            else x1 = (kk == 2) ? x0 : (e[kk - 1] - e[kk - 2]) + (e[kk - 1] - e[0]);
            R c2c = R(0.5) * (x0 + x1);
            R t = R(1) / (c2c * (R)(kk - 1));
            m.t2[lane] = t * t * K.inv_g2;
            m.lnh[lane] = rt<R>::log(x0);
        }
        __syncwarp();
    }
    __device__ __forceinline__ R prior_op(int kk, const R* t2, int i, int j) const
    {
        if (kk == 1) return K.inv_s2 + K.inv_g2;  // gradient_operator = ones((1,1))
        if (i == j) {
            R d = K.inv_s2;
            if (i > 0) d += t2[i - 1];
            if (i < kk - 1) d += t2[i];
            return d;
        }
        if (i == j + 1) return -t2[j];
        if (j == i + 1) return -t2[i];
        return R(0);
    }
    // gradient of lane i: Wm'Wm (ln s - ln ref) + J' Wd'Wd (pred - d)   (Model.local_gradient :347-357)
    __device__ __noinline__ R gradient_lane(int kk, const R* t2, const R* ls, const T* J, const T* pred)
    {
        GBP_SHARED_STATE;
        R g = R(0);
        if (lane < kk) {
            g = prior_op(kk, t2, lane, lane) * (ls[lane] - ln_ref);
            if (lane > 0) g += prior_op(kk, t2, lane, lane - 1) * (ls[lane - 1] - ln_ref);
            if (lane < kk - 1) g += prior_op(kk, t2, lane, lane + 1) * (ls[lane + 1] - ln_ref);
#pragma unroll 1
            for (int c = 0; c < C; ++c) g += (R)J[c * KS + lane] * (((R)pred[c] - w.data[c]) * w.ivar[c]);
        }
        return g;
    }
    // A = Wm'Wm + J' Wd'Wd J (Model.local_precision :250-272), packed lower triangle
    __device__ __noinline__ void assemble(int kk, const R* t2, const T* J)
    {
        GBP_SHARED_STATE;
#pragma unroll 1
        for (int i = 0; i < kk; ++i) {
            if (lane <= i) {
                R s = prior_op(kk, t2, i, lane);
#pragma unroll 1
                for (int c = 0; c < C; ++c) s += (R)J[c * KS + i] * w.ivar[c] * (R)J[c * KS + lane];
                w.A[pk(i, lane)] = s;
            }
        }
        __syncwarp();
    }
    // in-place packed Cholesky, lane = row.  Returns false if the matrix is not positive definite.
    __device__ __noinline__ bool cholesky(int kk)
    {
        GBP_SHARED_STATE;
        bool ok = true;
#pragma unroll 1
        for (int j = 0; j < kk; ++j) {
            R s = R(0);
            if (lane >= j && lane < kk) {
                s = w.A[pk(lane, j)];
#pragma unroll 1
                for (int p = 0; p < j; ++p) s -= w.A[pk(lane, p)] * w.A[pk(j, p)];
            }
            const R d = __shfl_sync(FULL, s, j);
            if (!(d > R(0))) ok = false;
            const R dj = rt<R>::sqrt(d);
            if (lane == j) w.A[pk(j, j)] = dj;
            else if (lane > j && lane < kk) w.A[pk(lane, j)] = s / dj;
            __syncwarp();
        }
        return ok;
    }
    // lane i holds b_i; returns lane i of L^-1 b
    __device__ __noinline__ R solve_L(int kk, R x)
    {
        GBP_SHARED_STATE;
#pragma unroll 1
        for (int j = 0; j < kk; ++j) {
            const R yj = __shfl_sync(FULL, x, j) / w.A[pk(j, j)];
            if (lane == j) x = yj;
            else if (lane > j && lane < kk) x -= w.A[pk(lane, j)] * yj;
        }
        return x;
    }
    __device__ __noinline__ R solve_LT(int kk, R x)
    {
        GBP_SHARED_STATE;
#pragma unroll 1
        for (int j = kk - 1; j >= 0; --j) {
            const R xj = __shfl_sync(FULL, x, j) / w.A[pk(j, j)];
            if (lane == j) x = xj;
            else if (lane < j) x -= w.A[pk(j, lane)] * xj;
        }
        return x;
    }
    // v' A v = |L' v|^2 with v_i held by lane i
    __device__ __noinline__ R quad(int kk, R v)
    {
        GBP_SHARED_STATE;
        if (lane < kk) w.vec[lane] = v;
        __syncwarp();
        R s = R(0);
        if (lane < kk) {
#pragma unroll 1
            for (int i = lane; i < kk; ++i) s += w.A[pk(i, lane)] * w.vec[i];
        }
        __syncwarp();
        return warp_sum(s * s);
    }

    // ------------------------------------------------------------ structure proposal (warp-uniform)
    // RectilinearMesh1D.perturb :993-1120.  Writes the proposed mesh / remapped values into the spare
    // buffers (mesh[mcur^1], val[vcur^1], ls_r) and returns the action; for ACT_NONE only ls_r is filled.
    __device__ __noinline__ int perturb_structure(int* knew)
    {
        GBP_SHARED_STATE;
        const MeshBuf<R>& m0 = w.mesh[mcur];
        const ValBuf<R>& v0 = w.val[vcur];
        MeshBuf<R>& m1 = w.mesh[mcur ^ 1];
        ValBuf<R>& v1 = w.val[vcur ^ 1];
        const int kmax = P.opt.max_layers;
#pragma unroll 1
        for (;;) {
            int event;
#pragma unroll 1
            for (;;) {  // Categorical.rng: searchsorted(cumsum(p), U), re-drawn while illegal (:1041-1049)
                const R u = rng_uniform<R>(rng);
                event = (u <= K.cum0) ? 0 : (u <= K.cum1) ? 1 : (u <= K.cum2) ? 2 : 3;
                if (k == 1 && (event == 1 || event == 2)) continue;
                if (k == kmax && event == 0) continue;
                break;
            }
            if (event == ACT_NONE) {
                if (lane < k) w.ls_r[lane] = v0.ls[lane];
                __syncwarp();
                *knew = k;
                return ACT_NONE;
            }
            if (event == ACT_BIRTH) {  // :1061-1081
                bool ok = false;
                int pos = 0;
                R e = R(0);
#pragma unroll 1
                for (int tries = 1; tries <= 10; ++tries) {
                    e = rt<R>::exp(K.ln_min_edge + K.ln_edge_span * rng_uniform<R>(rng));
                    pos = __popc(__ballot_sync(FULL, lane <= k && m0.edges[lane] < e));  // searchsorted (left)
                    R d = INFINITY;  // widths after insertion: cell pos-1 is split in two
                    if (lane < k)
                        d = (lane == pos - 1) ? fmin(e - m0.edges[lane], m0.edges[lane + 1] - e)
                                              : m0.edges[lane + 1] - m0.edges[lane];
                    const R h = warp_min(d);
                    if (tries == 10) break;  // the 10th try always restarts (:1078-1080)
                    if (h > K.min_width) {
                        ok = true;
                        break;
                    }
                }
                if (!ok) continue;
                if (lane <= k + 1) m1.edges[lane] = (lane < pos) ? m0.edges[lane] : (lane == pos ? e : m0.edges[lane - 1]);
                if (lane <= k) {  // values.insert(pos, values[pos-1]) (:835)
                    const int src = (lane < pos) ? lane : lane - 1;
                    v1.sig[lane] = v0.sig[src];
                    w.ls_r[lane] = v0.ls[src];
                }
                __syncwarp();
                *knew = k + 1;
                return ACT_BIRTH;
            }
            if (event == ACT_DEATH) {  // :1083-1087, delete_edge :643-689
                const int i = (int)(rng_uniform<R>(rng) * (R)(k - 1)) + 1;
                if (lane <= k - 1) m1.edges[lane] = m0.edges[lane + (lane >= i ? 1 : 0)];
                if (lane < k - 1) {
                    const int src = lane + (lane >= i ? 1 : 0);
                    R s = v0.sig[src], l = v0.ls[src];
                    if (lane == i - 1) {
                        s = R(0.5) * (v0.sig[i - 1] + v0.sig[i]);
                        l = rt<R>::log(s);
                    }
                    v1.sig[lane] = s;
                    w.ls_r[lane] = l;
                }
                __syncwarp();
                *knew = k - 1;
                return ACT_DEATH;
            }
            {  // ACT_MOVE :1088-1118
                bool ok = false;
                int i = 1;
                R dz = R(0);
#pragma unroll 1
                for (int tries = 1; tries <= 10; ++tries) {
                    i = (int)(R(1) + ((R)k - R(1)) * rng_uniform<R>(rng));
                    const R zn = rng_normal<R>(rng);
                    const R sgn = (zn > R(0)) ? R(1) : (zn < R(0) ? R(-1) : R(0));
                    dz = sgn * K.min_width * rng_uniform<R>(rng);
                    R d = INFINITY;
                    if (lane < k)
                        d = (m0.edges[lane + 1] + (lane + 1 == i ? dz : R(0))) - (m0.edges[lane] + (lane == i ? dz : R(0)));
                    const R h = warp_min(d);
                    const R z1 = m0.edges[1] + (i == 1 ? dz : R(0));
                    const R zl = m0.edges[k - 1] + (i == k - 1 ? dz : R(0));
                    if (tries == 10) break;
                    if (h > K.min_width && z1 > K.min_edge && zl < K.max_edge) {
                        ok = true;
                        break;
                    }
                }
                if (!ok) continue;
                if (lane <= k) m1.edges[lane] = m0.edges[lane] + (lane == i ? dz : R(0));
                if (lane < k) {
                    v1.sig[lane] = v0.sig[lane];
                    w.ls_r[lane] = v0.ls[lane];
                }
                __syncwarp();
                *knew = k;
                return ACT_MOVE;
            }
        }
    }

    // StatArray.propose(imposePrior=True) for a 1-D log-normal random walk (StatArray.py:578-638), in ln space
    __device__ __noinline__ R propose_ln_error(R ln_cur, R sd, R lnmin, R lnmax)
    {
        R x = ln_cur + sd * rng_normal<R>(rng);
        int tries = 0;
#pragma unroll 1
        while (x < lnmin || x > lnmax) {
            x = ln_cur + sd * rng_normal<R>(rng);
            tries++;
            if (tries == 10) return ln_cur;
        }
        return x;
    }

    // ------------------------------------------------------------ posterior accumulators
    // add `count` visits of the CURRENT model / errors to every histogram
    __device__ __noinline__ void flush(int count)
    {
        GBP_SHARED_STATE;
        if (count <= 0) return;
        const gbp_chain_buffers& o = P.out;
        const int nd = P.n_depth, nsb = P.opt.n_sigma_bins, neb = P.opt.n_err_bins;
        const MeshBuf<R>& m = w.mesh[mcur];
        const ValBuf<R>& v = w.val[vcur];
        if (lane == 0) {
            if (o.ncells_hist) o.ncells_hist[(size_t)chain * (P.opt.max_layers + 1) + k] += count;
            if (o.rel_hist && P.opt.solve_relative_error)
                o.rel_hist[(size_t)chain * neb + uniform_bin<R>(ln_rel, K.rel_lnmin, K.rel_dx, neb)] += count;
            if (o.add_hist && P.opt.solve_additive_error)
                o.add_hist[(size_t)chain * neb + uniform_bin<R>(ln_add, K.add_lnmin, K.add_dx, neb)] += count;
        }
        // per-layer conductivity bin; interface histogram (RectilinearMesh1D.update_posteriors :1594-1610):
        // ratio sigma_i / sigma_{i-1} <= 0.5 or >= 1.5
        if (lane < k) w.sbin[lane] = uniform_bin<R>(v.ls[lane], sig_lo, K.sig_dx, nsb);
        if (lane >= 1 && lane < k && o.edges_hist) {
            const R dl = v.ls[lane] - v.ls[lane - 1];
            const R d = m.edges[lane];
            if ((dl <= K.ln_half || dl >= K.ln_3half) && d >= R(0) && d < K.depth_max)
                atomicAdd(&o.edges_hist[(size_t)chain * nd + uniform_bin<R>(d, R(0), K.depth_step, nd)], count);
        }
        __syncwarp();
        // hitmap (Model.update_parameter_posterior :819-847; staircase interp RectilinearMesh1D.py:1148-1158)
        if (o.hitmap) {
            int32_t* hm = o.hitmap + (size_t)chain * nsb * nd;
#pragma unroll 1
            for (int j = lane; j < nd; j += 32) {
                const R y = ((R)j + R(0.5)) * K.depth_step;
                int b = w.sbin[k - 1];
#pragma unroll 1
                for (int i = 1; i < k; ++i) {
                    const R e = m.edges[i];
                    if (y < e) {
                        b = w.sbin[i - 1];
                        break;
                    }
                    const R e2 = e * R(1.000001);
                    if (y < e2) {
                        const R t = (y - e) / (e2 - e);
                        const R s = v.sig[i - 1] + t * (v.sig[i] - v.sig[i - 1]);
                        b = uniform_bin<R>(rt<R>::log(s), sig_lo, K.sig_dx, nsb);
                        break;
                    }
                }
                atomicAdd(&hm[(size_t)b * nd + j], count);  // RED.ADD, coalesced along depth
            }
        }
        __syncwarp();
    }

    __device__ __noinline__ void zero_posteriors()
    {
        const gbp_chain_buffers& o = P.out;
        const int nd = P.n_depth;
        if (o.hitmap) {
            int32_t* hm = o.hitmap + (size_t)chain * P.opt.n_sigma_bins * nd;
            const size_t n = (size_t)P.opt.n_sigma_bins * nd;
#pragma unroll 1
            for (size_t i = lane; i < n; i += 32) hm[i] = 0;
        }
        if (o.edges_hist) {
#pragma unroll 1
            for (int i = lane; i < nd; i += 32) o.edges_hist[(size_t)chain * nd + i] = 0;
        }
        if (o.ncells_hist) {
#pragma unroll 1
            for (int i = lane; i <= P.opt.max_layers; i += 32) o.ncells_hist[(size_t)chain * (P.opt.max_layers + 1) + i] = 0;
        }
        if (o.rel_hist) {
#pragma unroll 1
            for (int i = lane; i < P.opt.n_err_bins; i += 32) o.rel_hist[(size_t)chain * P.opt.n_err_bins + i] = 0;
        }
        if (o.add_hist) {
#pragma unroll 1
            for (int i = lane; i < P.opt.n_err_bins; i += 32) o.add_hist[(size_t)chain * P.opt.n_err_bins + i] = 0;
        }
        __syncwarp();
    }

    __device__ __noinline__ void write_model(double* sig_out, double* edges_out)
    {
        const int ml = P.opt.max_layers;
        const MeshBuf<R>& m = w.mesh[mcur];
        const ValBuf<R>& v = w.val[vcur];
        if (sig_out && lane < ml) sig_out[(size_t)chain * ml + lane] = lane < k ? (double)v.sig[lane] : NAN;
        if (edges_out) {
#pragma unroll 1
            for (int i = lane; i <= ml; i += 32) edges_out[(size_t)chain * (ml + 1) + i] = i <= k ? (double)m.edges[i] : NAN;
        }
    }
    __device__ __noinline__ void save_best()
    {
        write_model(P.out.best_sigma, P.out.best_edges);
        best_k = k;
        best_rel = rel;
        best_add = add;
        best_posterior = posterior;
        best_iter = iteration;
    }

    // ------------------------------------------------------------ Inference1D.initialize
    __device__ __noinline__ void initialize(bool first)
    {
        const gbp_options& o = P.opt;
        mcur = vcur = jcur = pcur = 0;
        ln_rel = K.rel_ln0;
        ln_add = K.add_ln0;
        rel = (R)o.rel_init;
        add = (R)o.add_init;
        set_ivar(rel, add);
        // EmDataPoint.find_best_halfspace :148-186: argmin misfit over logspace(-4, 4, 100)
        MeshBuf<R>& m = w.mesh[0];
        ValBuf<R>& v = w.val[0];
        if (lane == 0) {
            m.edges[0] = R(0);
            m.edges[1] = INFINITY;
        }
        __syncwarp();
        R best = INFINITY;
        double best_c = 0.0;
#pragma unroll 1
        for (int i = 0; i < 100; ++i) {
            const double e = (i == 99) ? 4.0 : -4.0 + (double)i * (8.0 / 99.0);
            const double c = dexp_(e * 2.302585092994045684017991454684);
            if (lane == 0) v.sig[0] = (R)c;
            __syncwarp();
            forward(1, v.sig, m.edges, w.pred[0], nullptr);
            R mis, like;
            misfit_likelihood(w.pred[0], &mis, &like);
            if (mis < best) {
                best = mis;
                best_c = c;
            }
        }
        sigma_ref = best_c;
        ln_ref = (R)dlog_(best_c);
        k = 1;
        if (lane == 0) {
            v.sig[0] = (R)best_c;
            v.ls[0] = ln_ref;
        }
        __syncwarp();
        forward(1, v.sig, m.edges, w.pred[0], w.J[0]);
        sig_lo = ln_ref - K.sig_halfspan;  // Model.set_posteriors :666-684
        if (!first) {                       // reset(): posteriors and traces are re-created
            zero_posteriors();
            const size_t N2 = 2 * (size_t)o.n_markov_chains;
            if (P.out.misfit_trace) {
#pragma unroll 1
                for (size_t i = lane; i < N2; i += 32) P.out.misfit_trace[(size_t)chain * N2 + i] = 0.0;
            }
            if (P.out.accept_trace) {
#pragma unroll 1
                for (size_t i = lane; i < N2; i += 32) P.out.accept_trace[(size_t)chain * N2 + i] = 0;
            }
            __syncwarp();
        }
        misfit_likelihood(w.pred[0], &misfit, &likelihood);
        prior = model_probability(1, v.ls, m.lnh) + datapoint_probability(ln_rel, ln_add);
        posterior = likelihood + prior;
        burned_in = 0;
        burned_in_iter = 0;
        iteration = 0;
        if (P.out.misfit_trace && lane == 0) P.out.misfit_trace[(size_t)chain * 2 * o.n_markov_chains] = (double)misfit;
        save_best();
        n_zero = 0;
        acc_win = 0;
        dwell = 0;
    }

    // ------------------------------------------------------------ Inference1D.accept_reject
    // returns true if the chain failed (Gauss-Newton matrix not positive definite)
    __device__ __noinline__ bool step(bool* accepted_out)
    {
        GBP_SHARED_STATE;
        const gbp_options& o = P.opt;
        *accepted_out = false;
        int kn;
        const int action = perturb_structure(&kn);
        n_act[action]++;
        const bool changed = action != ACT_NONE;
        const int mp = changed ? (mcur ^ 1) : mcur;  // proposed mesh buffer
        const int vp = vcur ^ 1;                     // proposed values buffer
        MeshBuf<R>& mesh_p = w.mesh[mp];
        ValBuf<R>& val_p = w.val[vp];

        const T* Jh = w.J[jcur];
        const T* ph = w.pred[pcur];
        T* pred_t = w.pred[pcur ^ 1];
        T* J_t = w.J[jcur ^ 1];
        if (changed) {  // observation.fm_dlogc(remapped_model): J and predicted data of the test datapoint
            mesh_setup(kn, mesh_p);
            forward(kn, val_p.sig, mesh_p.edges, pred_t, J_t);
            Jh = J_t;
            ph = pred_t;
        }
        set_ivar(rel, add);
        const R ln_r = (lane < kn) ? w.ls_r[lane] : R(0);
        const R g = gradient_lane(kn, mesh_p.t2, w.ls_r, Jh, ph);
        assemble(kn, mesh_p.t2, Jh);
        if (!cholesky(kn)) return true;
        const R stepv = solve_LT(kn, solve_L(kn, g));   // H * dfk
        const R mean = ln_r - K.alpha * stepv;           // ln sigma + alpha * pk, pk = -H dfk
        // sigma' ~ exp(N(mean, H)),  H = (L L')^-1  ->  mean + L^-T z
        R z0 = R(0), z1 = R(0);
        const int npair = (kn + 1) / 2;
        if (lane < npair) normal2_at<R>(rng, rng.block + (unsigned long long)lane, &z0, &z1);
        rng.block += (unsigned long long)npair;
        const R za = __shfl_sync(FULL, z0, lane >> 1), zb = __shfl_sync(FULL, z1, lane >> 1);
        const R zi = (lane < kn) ? ((lane & 1) ? zb : za) : R(0);
        const R ln_t = mean + solve_LT(kn, zi);
        if (lane < kn) {
            val_p.ls[lane] = ln_t;
            val_p.sig[lane] = rt<R>::exp(ln_t);
        }
        __syncwarp();

        // test_datapoint.perturb() (DataPoint.py:531-573)
        R lr_t = ln_rel, la_t = ln_add;
        if (o.solve_relative_error) lr_t = propose_ln_error(ln_rel, K.rel_sd, K.rel_lnmin, K.rel_lnmax);
        if (o.solve_additive_error) la_t = propose_ln_error(ln_add, K.add_sd, K.add_lnmin, K.add_lnmax);
        const R rel_t = (lr_t == ln_rel) ? rel : rt<R>::exp(lr_t);
        const R add_t = (la_t == ln_add) ? add : rt<R>::exp(la_t);

        const bool jump = (action == ACT_BIRTH || action == ACT_DEATH);
        // forward at the candidate; for birth/death the Jacobian at the candidate is needed as well
        // (Model.proposal_probabilities :619) - fused into the same pass.
        forward(kn, val_p.sig, mesh_p.edges, pred_t, jump ? J_t : nullptr);
        set_ivar(rel_t, add_t);
        R t_misfit, t_like;
        misfit_likelihood(pred_t, &t_misfit, &t_like);
        R t_prior = datapoint_probability(lr_t, la_t);
        if (t_prior == (R)-INFINITY) return false;
        t_prior += model_probability(kn, val_p.ls, mesh_p.lnh);
        if (t_prior == (R)-INFINITY) return false;

        R proposal = R(1), proposal1 = R(1);
        if (jump) {
            const R g2 = gradient_lane(kn, mesh_p.t2, val_p.ls, J_t, pred_t);
            const R s2 = solve_LT(kn, solve_L(kn, g2));   // H dfk'
            const R lv = ln_t + K.alpha * s2;              // Model.py:626 (sign as in the reference)
            const R mv = rt<R>::exp(lv);
            const int bad = __any_sync(FULL, lane < kn && (mv == (R)INFINITY || mv == R(0)));
            const R q_r = quad(kn, (lane < kn) ? (ln_r - lv) : R(0));
            const R q_f = quad(kn, (lane < kn) ? (ln_t - ln_r) : R(0));
            const R logdetL = warp_sum((lane < kn) ? rt<R>::log(w.A[pk(lane, lane)]) : R(0));
            if (bad) {
                proposal = (R)-INFINITY;
                proposal1 = (R)-INFINITY;
            } else {
                proposal = -(R)kn * K.half_log2pi + logdetL - R(0.5) * q_r;
                proposal1 = -(R)kn * K.half_log2pi + logdetL - R(0.5) * q_f;
            }
        }
        const R log_alpha = (t_prior - prior) + (t_like - likelihood) + (proposal - proposal1);
        const R u = rng_uniform<R>(rng);
        const bool acc = rt<R>::exp(log_alpha) > u;
        if (acc) {
            flush(dwell);  // the outgoing model's visits
            dwell = 0;
            misfit = t_misfit;
            prior = t_prior;
            likelihood = t_like;
            posterior = t_prior + t_like;
            k = kn;
            rel = rel_t;
            add = add_t;
            ln_rel = lr_t;
            ln_add = la_t;
            mcur = mp;
            vcur = vp;
            pcur ^= 1;
            if (changed) jcur ^= 1;  // action none keeps the (stale) Jacobian, as the reference does
            n_accept++;
        }
        *accepted_out = acc;
        return false;
    }

    // ------------------------------------------------------------ Inference1D.update; returns true on reset
    __device__ __noinline__ bool update(bool accepted)
    {
        const gbp_options& o = P.opt;
        const long long N2 = 2 * (long long)o.n_markov_chains;
        bool do_reset = false;
        iteration++;
        if (P.out.misfit_trace && lane == 0 && iteration - 1 < N2)
            P.out.misfit_trace[(size_t)chain * N2 + (iteration - 1)] = (double)misfit;
        if (!burned_in && iteration > o.burn_in_min_iter && misfit < (R)n_active) {
            burned_in = 1;
            burned_in_iter = iteration;
            save_best();
            zero_posteriors();
            dwell = 0;
        }
        if (posterior > best_posterior) save_best();
        if (P.out.accept_trace && lane == 0 && iteration < N2)
            P.out.accept_trace[(size_t)chain * N2 + iteration] = accepted ? 1 : 0;
        if (iteration % o.update_plot_every == 0) {
            // acceptance over acceptance_v[it-upe : it] (Inference1D.py:125-131): excludes this iteration
            const int s = acc_win;
            acc_win = 0;
            if (o.update_plot_every > 1) {
                if (!burned_in) {
                    if (s == 0) {
                        n_zero++;
                        if (n_zero == o.reset_limit) {
                            do_reset = true;
                            n_zero = 0;
                        }
                    } else n_zero = 0;
                } else if (s == 0) limiters = 0;
            }
        }
        acc_win += accepted ? 1 : 0;
        if (do_reset) return true;
        dwell++;
        return false;
    }

    // ------------------------------------------------------------ Inference1D.infer
    __device__ __noinline__ void run(int chain_)
    {
        const gbp_options& o = P.opt;
        chain = chain_;
        const unsigned long long snd = P.first_index + (unsigned long long)chain;
        rng.seed_lo = (uint32_t)P.seed;
        rng.seed_hi = (uint32_t)(P.seed >> 32);
        rng.snd_lo = (uint32_t)snd;
        rng.snd_hi = (uint32_t)(snd >> 32);
        rng.block = 0;
        alt = (T)P.altitude[chain];
        int act = 0;
        if (lane < C) {
            double d = P.data[(size_t)chain * C + lane];
            act = d > 0.0;               // EmDataPoint.active: observed > 0 and not NaN
            w.data[lane] = act ? (R)d : R(0);
        }
        n_active = warp_sum_i(act);
        n_accept = n_forward = n_sens = 0;
        n_act[0] = n_act[1] = n_act[2] = n_act[3] = 0;
        n_resets = 0;
        limiters = 0;
        __syncwarp();
        initialize(true);

        bool failed = (n_active == 0);
        bool go = !failed;
        long long total = 0;
        const int N = o.n_markov_chains;
#pragma unroll 1
        while (go) {
            bool accepted;
            failed = step(&accepted);
            const bool reset = update(accepted);
            total++;
            if (reset) {
                n_resets++;
                initialize(false);
                dwell = 1;  // update() continues on the re-initialised state
            }
            go = !failed && (iteration <= N + burned_in_iter);
            if (!failed && !burned_in) {
                go = iteration < N;
                if (!go) failed = true;
            }
            if (n_resets == 3 && !burned_in) {
                if (!limiters) {
                    limiters = 1;
                    n_resets = 1;
                    initialize(false);
                } else {
                    go = false;
                    failed = true;
                }
            }
            if (P.max_iterations > 0 && total >= P.max_iterations) go = false;
        }
        flush(dwell);
        dwell = 0;

        write_model(P.out.cur_sigma, P.out.cur_edges);
        if (lane == 0) {
            double* s = P.out.scalars + (size_t)chain * GBP_NSCALARS;
#pragma unroll 1
            for (int i = 0; i < GBP_NSCALARS; ++i) s[i] = 0.0;
            s[GBP_S_ITER] = (double)iteration;
            s[GBP_S_BURNED_IN] = burned_in;
            s[GBP_S_BURNED_IN_ITER] = (double)burned_in_iter;
            s[GBP_S_BEST_ITER] = (double)best_iter;
            s[GBP_S_BEST_K] = best_k;
            s[GBP_S_CUR_K] = k;
            s[GBP_S_HALFSPACE] = sigma_ref;
            s[GBP_S_FAILED] = failed ? 1.0 : 0.0;
            s[GBP_S_N_ACCEPT] = (double)n_accept;
            s[GBP_S_N_FORWARD] = (double)n_forward;
            s[GBP_S_N_SENS] = (double)n_sens;
            s[GBP_S_BEST_POSTERIOR] = (double)best_posterior;
            s[GBP_S_CUR_REL] = (double)rel;
            s[GBP_S_CUR_ADD] = (double)add;
            s[GBP_S_CUR_MISFIT] = (double)misfit;
            s[GBP_S_CUR_PRIOR] = (double)prior;
            s[GBP_S_CUR_LIKELIHOOD] = (double)likelihood;
            s[GBP_S_BEST_REL] = (double)best_rel;
            s[GBP_S_BEST_ADD] = (double)best_add;
            s[GBP_S_N_RESETS] = n_resets;
            s[GBP_S_N_BIRTH] = (double)n_act[0];
            s[GBP_S_N_DEATH] = (double)n_act[1];
            s[GBP_S_N_MOVE] = (double)n_act[2];
            s[GBP_S_N_NONE] = (double)n_act[3];
        }
        __syncwarp();
    }
};

// ---------------------------------------------------------------- kernels
// R = sampler arithmetic, T = forward/Jacobian arithmetic, NC = channel capacity, WARPS = warps per CTA
template <typename R, typename T, int NC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
    rjmcmc_kernel(const __grid_constant__ SysDev S, const T* __restrict__ g_tab, const __grid_constant__ ChainParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ Consts<R> consts;
    __shared__ SysShared<T> sys_s;
    T* tab = reinterpret_cast<T*>(smem);
    const uint32_t tab_bytes = (uint32_t)(TAB_ROWS * S.tab_stride * sizeof(T));
    if (threadIdx.x == 0) {
        make_consts<R>(P.opt, P.n_depth, consts);
        fill_sys_shared<T>(S, sys_s);
    }
    tma_stage(tab, g_tab, tab_bytes, &bar);
    __syncthreads();
    const uint32_t tab_pad = (tab_bytes + 127u) & ~127u;
    const int warp = threadIdx.x >> 5;
    WarpState<R, T, NC>* ws = reinterpret_cast<WarpState<R, T, NC>*>(smem + tab_pad) + warp;
    Chain<R, T, NC> ch(*ws, consts, sys_s, tab, P);
    // persistent: the first wave is assigned statically, later chains come from a device-side counter
    int c = blockIdx.x * WARPS + warp;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    while (c < P.B) {
        ch.run(c);
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(P.work_counter, 1);
        c = __shfl_sync(FULL, nxt, 0);
    }
}

// standalone operators: one warp per sounding, grid-stride over soundings
template <typename T, bool SENS>
__global__ void __launch_bounds__(256) fdem_kernel(const __grid_constant__ SysDev S, const T* __restrict__ g_tab, int B,
                                                    int l_stride, const int32_t* __restrict__ nlayers,
                                                    const double* __restrict__ sigma, const double* __restrict__ thickness,
                                                    const double* __restrict__ altitude, double* __restrict__ out,
                                                    double* __restrict__ Jout)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ SysShared<T> sys_s;
    T* tab = reinterpret_cast<T*>(smem);
    const uint32_t tab_bytes = (uint32_t)(TAB_ROWS * S.tab_stride * sizeof(T));
    if (threadIdx.x == 0) fill_sys_shared<T>(S, sys_s);
    tma_stage(tab, g_tab, tab_bytes, &bar);
    __syncthreads();
    const uint32_t tab_pad = (tab_bytes + 127u) & ~127u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int C = 2 * S.n_freq;
    constexpr int PER_WARP = 2 * KS + GBP_MAXC + (SENS ? GBP_MAXC * KS : 0);
    T* base = reinterpret_cast<T*>(smem + tab_pad) + (size_t)warp * PER_WARP;
    T* msig = base;
    T* mthk = base + KS;
    T* pred = base + 2 * KS;
    T* J = base + 2 * KS + GBP_MAXC;
#pragma unroll 1
    for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
        const int L = nlayers[b];
        if (lane < L) {
            msig[lane] = (T)sigma[(size_t)b * l_stride + lane];
            mthk[lane] = (T)thickness[(size_t)b * l_stride + lane];
        }
        __syncwarp();
        fdem_eval<T>(sys_s, tab, (T)altitude[b], L, msig, mthk, pred, SENS ? J : nullptr, SENS);
        if (lane < C) out[(size_t)b * C + lane] = (double)pred[lane];
        if (SENS) {
#pragma unroll 1
            for (int i = lane; i < C * l_stride; i += 32) {
                const int c = i / l_stride, kk = i % l_stride;
                Jout[(size_t)b * C * l_stride + i] = (kk < L) ? (double)J[c * KS + kk] : 0.0;
            }
        }
        __syncwarp();
    }
}

}  // namespace gbp
