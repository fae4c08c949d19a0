// One warp = one Markov chain: the fused trans-dimensional MCMC sampler for one FDEM sounding.
//
// Replaces the Python loop Inference1D.initialize + infer (geobipy/src/inversion/Inference1D.py:353-464,
// :537-631 accept_reject, :633-688 infer, :705-790 update) together with everything it calls:
//   RectilinearMesh1D.perturb           classes/mesh/RectilinearMesh1D.py:993-1120
//   Model.stochastic_newton_perturbation classes/model/Model.py:368-419 (+ :250-272, :347-357, :421-430)
//   Model.probability / gradient_probability / proposal_probabilities  Model.py:533-575, :213-234, :577-660
//   DataPoint.std / data_misfit / likelihood / probability / perturb   classes/data/datapoint/DataPoint.py
//   Model.update_parameter_posterior, RectilinearMesh1D.update_posteriors, EmDataPoint.update_posteriors
//
// Design: chain state lives in shared memory (one WarpState per warp) and registers; the k x k
// Gauss-Newton system is factorised in-warp (packed Cholesky, lane = row) instead of the reference's
// inv() + SVD; random numbers come from a counter-based Philox4x32-10 stream (key = seed,
// counter = (block, sounding)); posterior histograms are updated with a dwell count, i.e. only
// when the model changes (accept) - identical counts, ~1/acceptance fewer HBM read-modify-writes.
#pragma once
#include "gbp_fdem.cuh"

namespace gbp {

constexpr int NPACK = GBP_MAXL * (GBP_MAXL + 1) / 2;
constexpr double LOG2PI = 1.8378770664093454835606594728112;
enum { ACT_BIRTH = 0, ACT_DEATH = 1, ACT_MOVE = 2, ACT_NONE = 3 };

template <typename T, int NC> struct __align__(16) WarpState {
    double A[NPACK];                       // packed lower triangle: Hessian, then its Cholesky factor
    double edges_c[GBP_MAXL + 2];          // current model
    double edges_p[GBP_MAXL + 2];          // proposed (remapped == test) edges
    double sig_c[GBP_MAXL], sig_r[GBP_MAXL], sig_t[GBP_MAXL];
    double lns[GBP_MAXL];                  // scratch: ln(sigma) - ln(sigma_ref)
    double t2[GBP_MAXL];                   // gradient-operator weights
    double vec[GBP_MAXL + 2];              // scratch vector
    double data[NC], ivar[NC];             // observed data, 1/variance (0 for inactive channels)
    T Jc[NC * KS], Jt[NC * KS];
    T pred_c[NC], pred_t[NC];
    T msig[KS], mthk[KS];
    int sbin[GBP_MAXL + 2];
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

__device__ __forceinline__ int pk(int i, int j) { return i * (i + 1) / 2 + j; }

// searchsorted(edges, v, 'right') - 1 clipped, uniform edges lo + i*dx
__device__ __noinline__ int uniform_bin(double v, double lo, double dx, int n)
{
    double f = floor((v - lo) / dx);
    int i = (f < 0.0) ? 0 : (f > (double)(n - 1) ? n - 1 : (int)f);
    #pragma unroll 1
    while (i + 1 < n && v >= lo + (double)(i + 1) * dx) ++i;
    #pragma unroll 1
    while (i > 0 && v < lo + (double)i * dx) --i;
    return i;
}

__device__ __noinline__ double log_uniform_logpdf(double x, double mn, double mx)
{
    double lx = dlog_(x), a = dlog_(mn), b = dlog_(mx);
    if (lx < a || lx > b) return -INFINITY;
    return -dlog_(b - a);
}

// StatArray.propose(imposePrior=True) for a 1-D log-normal random walk (StatArray.py:578-638)
__device__ __noinline__ double propose_error(Rng& g, double cur, double prop_var, double mn, double mx)
{
    const double sd = sqrt(prop_var);
    double x = dexp_(dlog_(cur) + sd * rng_normal(g));
    int tries = 0;
    #pragma unroll 1
    while (log_uniform_logpdf(x, mn, mx) == -INFINITY) {
        x = dexp_(dlog_(cur) + sd * rng_normal(g));
        tries++;
        if (tries == 10) return cur;
    }
    return x;
}

template <typename T, int NC> struct Chain {
    typedef WarpState<T, NC> WS;
    WS& w;
    const SysDev& S;
    const T* tab;
    const ChainParams& P;
    const int lane;
    const int C;
    // registers
    Rng rng;
    int chain;
    T alt;
    int k;  // current layers
    double rel, add, sigma_ref, ln_ref;
    double misfit, prior, likelihood, posterior, best_posterior;
    long long iteration, burned_in_iter, best_iter, n_accept, n_forward, n_sens;
    long long n_act0, n_act1, n_act2, n_act3;
    int burned_in, n_zero, n_resets, limiters, n_active, acc_win, dwell, best_k;
    double best_rel, best_add;
    double sig_lo, sig_dx, rel_lo, rel_dx, add_lo, add_dx, depth_step;

    __device__ Chain(WS& w_, const SysDev& S_, const T* tab_, const ChainParams& P_)
        : w(w_), S(S_), tab(tab_), P(P_), lane(threadIdx.x & 31), C(P_.C)
    {
    }

    // ------------------------------------------------------------ forward wrappers
    __device__ __forceinline__ void load_model(int kk, const double* sig, const double* edges)
    {
        if (lane < kk) {
            w.msig[lane] = (T)sig[lane];
            w.mthk[lane] = (T)(edges[lane + 1] - edges[lane]);
        }
        __syncwarp();
    }
    __device__ __noinline__ void forward(int kk, const double* sig, const double* edges, T* pred)
    {
        load_model(kk, sig, edges);
        fdem_eval<T, false>(S, tab, alt, kk, w.msig, w.mthk, pred, nullptr);
        n_forward++;
    }
    // forward + Jacobian in one pass (FdemDataPoint.fm_dlogc, FdemDataPoint.py:535)
    __device__ __noinline__ void forward_sens(int kk, const double* sig, const double* edges, T* pred, T* J)
    {
        load_model(kk, sig, edges);
        fdem_eval<T, true>(S, tab, alt, kk, w.msig, w.mthk, pred, J);
        n_forward++;
        n_sens++;
    }

    // ------------------------------------------------------------ data terms
    // DataPoint.std :268-282 -> 1/variance per active channel (EmDataPoint.active :44-56)
    __device__ __noinline__ void set_ivar(double r, double a)
    {
        if (lane < C) {
            double d = w.data[lane];
            double s = r * d;
            w.ivar[lane] = (d > 0.0) ? 1.0 / (s * s + a * a) : 0.0;
        }
        __syncwarp();
    }
    // misfit (DataPoint.py:502-525) and Gaussian log-likelihood (MvNormalDistribution.py:209-216)
    __device__ __noinline__ void misfit_likelihood(const T* pred, double* mis, double* like)
    {
        double q = 0.0, ld = 0.0;
        if (lane < C) {
            double iv = w.ivar[lane];
            if (iv > 0.0) {
                double r = (double)pred[lane] - w.data[lane];
                q = r * r * iv;
                ld = -dlog_(iv);
            }
        }
        q = warp_sum(q);
        ld = warp_sum(ld);
        *mis = q;
        *like = -(0.5 * (double)n_active) * LOG2PI - 0.5 * ld - 0.5 * q;
    }
    __device__ __noinline__ double datapoint_probability(double r, double a)
    {
        double p = 0.0;
        if (P.opt.solve_relative_error) p += log_uniform_logpdf(r, P.opt.rel_min, P.opt.rel_max);
        if (P.opt.solve_additive_error) p += log_uniform_logpdf(a, P.opt.add_min, P.opt.add_max);
        return p;
    }
    // Model.probability :533-575 (value_bounds = None)
    __device__ __noinline__ double model_probability(int kk, const double* sig, const double* edges)
    {
        const gbp_options& o = P.opt;
        double p = (kk >= 1 && kk <= o.max_layers) ? -dlog_((double)o.max_layers - 1.0) : -INFINITY;
        if (o.solve_parameter) {
            double s2 = dlog_(1.0 + o.factor);
            s2 *= s2;
            double q = 0.0;
            if (lane < kk) {
                double d = dlog_(sig[lane]) - ln_ref;
                q = d * d / s2;
            }
            q = warp_sum(q);
            p += -(0.5 * kk) * LOG2PI - 0.5 * kk * dlog_(s2) - 0.5 * q;
        }
        if (o.solve_gradient) {
            const double g2 = o.gradient_std * o.gradient_std;
            if (kk == 1) {
                p += -0.5 * LOG2PI - 0.5 * dlog_(g2);
            } else {
                const int n = kk - 1;
                double q = 0.0;
                if (lane < n) {
                    double g = (dlog_(sig[lane + 1]) - dlog_(sig[lane])) / dlog_(edges[lane + 1] - edges[lane]);
                    q = g * g / g2;
                }
                q = warp_sum(q);
                p += -(0.5 * n) * LOG2PI - 0.5 * n * dlog_(g2) - 0.5 * q;
            }
        }
        return p;
    }

    // ------------------------------------------------------------ Gauss-Newton system
    // t2[i] = 1/(g^2 (c2c_i (k-1))^2)  (RectilinearMesh1D.gradient_operator :747-786), lns = ln sigma - ln ref
    __device__ __noinline__ void prior_setup(int kk, const double* sig, const double* edges)
    {
        const double g2 = P.opt.gradient_std * P.opt.gradient_std;
        if (lane < kk) w.lns[lane] = dlog_(sig[lane]) - ln_ref;
        if (kk >= 2 && lane < kk - 1) {
            double x0 = edges[lane + 1] - edges[lane];
            double x1;
            if (lane + 1 < kk - 1) x1 = edges[lane + 2] - edges[lane + 1];
            else x1 = (kk == 2) ? x0 : (edges[kk - 1] - edges[kk - 2]) + (edges[kk - 1] - edges[0]);
            double c2c = 0.5 * (x0 + x1);
            double t = 1.0 / (c2c * (double)(kk - 1));
            w.t2[lane] = t * t / g2;
        }
        __syncwarp();
    }
    __device__ __forceinline__ double prior_op(int kk, int i, int j) const
    {
        double s2 = dlog_(1.0 + P.opt.factor);
        s2 *= s2;
        if (kk == 1) return 1.0 / s2 + 1.0 / (P.opt.gradient_std * P.opt.gradient_std);
        if (i == j) {
            double d = 1.0 / s2;
            if (i > 0) d += w.t2[i - 1];
            if (i < kk - 1) d += w.t2[i];
            return d;
        }
        if (i == j + 1) return -w.t2[j];
        if (j == i + 1) return -w.t2[i];
        return 0.0;
    }
    // gradient of lane i: Wm'Wm (ln s - ln ref) + J' Wd'Wd (pred - d)   (Model.local_gradient :347-357)
    __device__ __noinline__ double gradient_lane(int kk, const T* J, const T* pred)
    {
        double g = 0.0;
        if (lane < kk) {
            g = prior_op(kk, lane, lane) * w.lns[lane];
            if (lane > 0) g += prior_op(kk, lane, lane - 1) * w.lns[lane - 1];
            if (lane < kk - 1) g += prior_op(kk, lane, lane + 1) * w.lns[lane + 1];
            #pragma unroll 1
            for (int c = 0; c < C; ++c) {
                double iv = w.ivar[c];
                g += (double)J[c * KS + lane] * (((double)pred[c] - w.data[c]) * iv);
            }
        }
        return g;
    }
    // A = Wm'Wm + J' Wd'Wd J (Model.local_precision :250-272), packed lower triangle
    __device__ __noinline__ void assemble(int kk, const T* J)
    {
        #pragma unroll 1
        for (int i = 0; i < kk; ++i) {
            if (lane <= i) {
                double s = prior_op(kk, i, lane);
                #pragma unroll 1
                for (int c = 0; c < C; ++c) s += (double)J[c * KS + i] * w.ivar[c] * (double)J[c * KS + lane];
                w.A[pk(i, lane)] = s;
            }
        }
        __syncwarp();
    }
    // in-place packed Cholesky, lane = row.  Returns false if not positive definite.  logdetL = sum ln L_jj
    __device__ __noinline__ bool cholesky(int kk, double* logdetL)
    {
        double ld = 0.0;
        bool ok = true;
        #pragma unroll 1
        for (int j = 0; j < kk; ++j) {
            double s = 0.0;
            if (lane >= j && lane < kk) {
                s = w.A[pk(lane, j)];
                #pragma unroll 1
                for (int p = 0; p < j; ++p) s -= w.A[pk(lane, p)] * w.A[pk(j, p)];
            }
            double d = __shfl_sync(FULL, s, j);
            if (!(d > 0.0)) ok = false;
            double dj = sqrt(d);
            ld += dlog_(dj);
            if (lane == j) w.A[pk(j, j)] = dj;
            else if (lane > j && lane < kk) w.A[pk(lane, j)] = s / dj;
            __syncwarp();
        }
        *logdetL = ld;
        return ok;
    }
    // lane i holds b_i; returns (L L')^-1 b for lane i
    __device__ __noinline__ double solve_L(int kk, double x)
    {
        #pragma unroll 1
        for (int j = 0; j < kk; ++j) {
            double yj = __shfl_sync(FULL, x, j) / w.A[pk(j, j)];
            if (lane == j) x = yj;
            else if (lane > j && lane < kk) x -= w.A[pk(lane, j)] * yj;
        }
        return x;
    }
    __device__ __noinline__ double solve_LT(int kk, double x)
    {
        #pragma unroll 1
        for (int j = kk - 1; j >= 0; --j) {
            double xj = __shfl_sync(FULL, x, j) / w.A[pk(j, j)];
            if (lane == j) x = xj;
            else if (lane < j) x -= w.A[pk(j, lane)] * xj;
        }
        return x;
    }
    // v' A v = |L' v|^2 with v_i held by lane i
    __device__ __noinline__ double quad(int kk, double v)
    {
        if (lane < kk) w.vec[lane] = v;
        __syncwarp();
        double s = 0.0;
        if (lane < kk)
            #pragma unroll 1
            for (int i = lane; i < kk; ++i) s += w.A[pk(i, lane)] * w.vec[i];
        __syncwarp();
        return warp_sum(s * s);
    }

    // ------------------------------------------------------------ structure proposal (lane 0, serial)
    // RectilinearMesh1D.perturb :993-1120.  Writes edges_p / sig_r, returns action and new k via shuffle.
    __device__ __noinline__ int perturb_structure(int* knew)
    {
        int action = 0, kn = k;
        unsigned long long blk = rng.block;
        if (lane == 0) {
            const gbp_options& o = P.opt;
            Rng g = rng;
            const double cum0 = o.p_birth, cum1 = cum0 + o.p_death, cum2 = cum1 + o.p_move;
            const double* e0 = w.edges_c;
            const double* s0 = w.sig_c;
            double* z = w.edges_p;
            double* sr = w.sig_r;
            #pragma unroll 1
            for (;;) {
                int event;
                #pragma unroll 1
                for (;;) {
                    double u = rng_uniform(g);
                    event = (u <= cum0) ? 0 : (u <= cum1) ? 1 : (u <= cum2) ? 2 : 3;
                    if (k == 1 && (event == 1 || event == 2)) continue;
                    if (k == o.max_layers && event == 0) continue;
                    break;
                }
                if (event == ACT_NONE) {
                    #pragma unroll 1
                    for (int i = 0; i <= k; ++i) z[i] = e0[i];
                    #pragma unroll 1
                    for (int i = 0; i < k; ++i) sr[i] = s0[i];
                    action = ACT_NONE;
                    kn = k;
                    break;
                }
                if (event == ACT_BIRTH) {
                    bool ok = false;
                    int pos = 0;
                    const double lo = dlog_(o.min_edge), hi = dlog_(o.max_edge);
                    #pragma unroll 1
                    for (int tries = 1; tries <= 10; ++tries) {
                        double e = dexp_(lo + (hi - lo) * rng_uniform(g));
                        pos = 0;
                        #pragma unroll 1
                        while (pos <= k && e0[pos] < e) ++pos;
                        // min width of the edges with e inserted at pos: only the two new cells can shrink
                        double h = INFINITY;
                        #pragma unroll 1
                        for (int i = 0; i + 1 <= k; ++i) {
                            if (i + 1 == pos) continue;
                            double d = e0[i + 1] - e0[i];
                            if (d < h) h = d;
                        }
                        if (pos >= 1) h = fmin(h, e - e0[pos - 1]);
                        if (pos <= k) h = fmin(h, e0[pos] - e);
                        if (tries == 10) break;
                        if (h > o.min_width) {
                            ok = true;
                            #pragma unroll 1
                            for (int i = 0; i < pos; ++i) z[i] = e0[i];
                            z[pos] = e;
                            #pragma unroll 1
                            for (int i = pos; i <= k; ++i) z[i + 1] = e0[i];
                            break;
                        }
                    }
                    if (!ok) continue;
                    #pragma unroll 1
                    for (int i = 0; i < pos; ++i) sr[i] = s0[i];
                    sr[pos] = s0[pos - 1];
                    #pragma unroll 1
                    for (int i = pos; i < k; ++i) sr[i + 1] = s0[i];
                    action = ACT_BIRTH;
                    kn = k + 1;
                    break;
                }
                if (event == ACT_DEATH) {
                    int i = (int)(rng_uniform(g) * (double)(k - 1)) + 1;
                    #pragma unroll 1
                    for (int j = 0; j < i; ++j) z[j] = e0[j];
                    #pragma unroll 1
                    for (int j = i + 1; j <= k; ++j) z[j - 1] = e0[j];
                    double val = 0.5 * (s0[i - 1] + s0[i]);
                    #pragma unroll 1
                    for (int j = 0; j < i; ++j) sr[j] = s0[j];
                    #pragma unroll 1
                    for (int j = i + 1; j < k; ++j) sr[j - 1] = s0[j];
                    sr[i - 1] = val;
                    action = ACT_DEATH;
                    kn = k - 1;
                    break;
                }
                {  // ACT_MOVE
                    bool ok = false;
                    #pragma unroll 1
                    for (int tries = 1; tries <= 10; ++tries) {
                        #pragma unroll 1
                        for (int i = 0; i <= k; ++i) z[i] = e0[i];
                        int i = (int)(1.0 + ((double)k - 1.0) * rng_uniform(g));
                        double zn = rng_normal(g);
                        double sgn = (zn > 0.0) ? 1.0 : (zn < 0.0 ? -1.0 : 0.0);
                        double dz = sgn * o.min_width * rng_uniform(g);
                        z[i] += dz;
                        double h = INFINITY;
                        #pragma unroll 1
                        for (int q = 0; q + 1 <= k; ++q) {
                            double d = z[q + 1] - z[q];
                            if (d < h) h = d;
                        }
                        if (tries == 10) break;
                        if (h > o.min_width && z[1] > o.min_edge && z[k - 1] < o.max_edge) {
                            ok = true;
                            break;
                        }
                    }
                    if (!ok) continue;
                    #pragma unroll 1
                    for (int i = 0; i < k; ++i) sr[i] = s0[i];
                    action = ACT_MOVE;
                    kn = k;
                    break;
                }
            }
            blk = g.block;
        }
        action = __shfl_sync(FULL, action, 0);
        kn = __shfl_sync(FULL, kn, 0);
        rng.block = __shfl_sync(FULL, blk, 0);
        __syncwarp();
        *knew = kn;
        return action;
    }

    // ------------------------------------------------------------ posterior accumulators
    __device__ __forceinline__ size_t nsig() const { return (size_t)P.opt.n_sigma_bins; }

    // add `count` visits of the CURRENT model / errors to every histogram
    __device__ __noinline__ void flush(int count)
    {
        if (count <= 0) return;
        const gbp_chain_buffers& o = P.out;
        const int nd = P.n_depth;
        if (lane == 0) {
            if (o.ncells_hist) o.ncells_hist[(size_t)chain * (P.opt.max_layers + 1) + k] += count;
            if (o.rel_hist && P.opt.solve_relative_error)
                o.rel_hist[(size_t)chain * P.opt.n_err_bins + uniform_bin(dlog_(rel), rel_lo, rel_dx, P.opt.n_err_bins)] += count;
            if (o.add_hist && P.opt.solve_additive_error)
                o.add_hist[(size_t)chain * P.opt.n_err_bins + uniform_bin(dlog_(add), add_lo, add_dx, P.opt.n_err_bins)] += count;
        }
        // per-layer conductivity bin; interface histogram (RectilinearMesh1D.update_posteriors :1594-1610)
        if (lane < k) w.sbin[lane] = uniform_bin(dlog_(w.sig_c[lane]), sig_lo, sig_dx, P.opt.n_sigma_bins);
        if (lane >= 1 && lane < k && o.edges_hist) {
            double r = dexp_(dlog_(w.sig_c[lane]) - dlog_(w.sig_c[lane - 1]));
            double d = w.edges_c[lane];
            if ((r <= 0.5 || r >= 1.5) && d >= 0.0 && d < (double)nd * depth_step)
                atomicAdd(&o.edges_hist[(size_t)chain * nd + uniform_bin(d, 0.0, depth_step, nd)], count);
        }
        __syncwarp();
        // hitmap (Model.update_parameter_posterior :819-847; staircase interp RectilinearMesh1D.py:1148-1158)
        if (o.hitmap) {
            int32_t* hm = o.hitmap + (size_t)chain * nsig() * nd;
            #pragma unroll 1
            for (int j = lane; j < nd; j += 32) {
                const double y = ((double)j + 0.5) * depth_step;
                int b = w.sbin[k - 1];
                #pragma unroll 1
                for (int i = 1; i < k; ++i) {
                    const double e = w.edges_c[i];
                    if (y < e) {
                        b = w.sbin[i - 1];
                        break;
                    }
                    const double e2 = e * 1.000001;
                    if (y < e2) {
                        double t = (y - e) / (e2 - e);
                        double v = w.sig_c[i - 1] + t * (w.sig_c[i] - w.sig_c[i - 1]);
                        b = uniform_bin(dlog_(v), sig_lo, sig_dx, P.opt.n_sigma_bins);
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
            int32_t* hm = o.hitmap + (size_t)chain * nsig() * nd;
            const size_t n = nsig() * nd;
            #pragma unroll 1
            for (size_t i = lane; i < n; i += 32) hm[i] = 0;
        }
        if (o.edges_hist)
            #pragma unroll 1
            for (int i = lane; i < nd; i += 32) o.edges_hist[(size_t)chain * nd + i] = 0;
        if (o.ncells_hist)
            #pragma unroll 1
            for (int i = lane; i <= P.opt.max_layers; i += 32) o.ncells_hist[(size_t)chain * (P.opt.max_layers + 1) + i] = 0;
        if (o.rel_hist)
            #pragma unroll 1
            for (int i = lane; i < P.opt.n_err_bins; i += 32) o.rel_hist[(size_t)chain * P.opt.n_err_bins + i] = 0;
        if (o.add_hist)
            #pragma unroll 1
            for (int i = lane; i < P.opt.n_err_bins; i += 32) o.add_hist[(size_t)chain * P.opt.n_err_bins + i] = 0;
        __syncwarp();
    }

    __device__ __noinline__ void save_best()
    {
        const gbp_chain_buffers& o = P.out;
        const int ml = P.opt.max_layers;
        if (o.best_sigma && lane < ml) o.best_sigma[(size_t)chain * ml + lane] = lane < k ? w.sig_c[lane] : NAN;
        if (o.best_edges)
            #pragma unroll 1
            for (int i = lane; i <= ml; i += 32) o.best_edges[(size_t)chain * (ml + 1) + i] = i <= k ? w.edges_c[i] : NAN;
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
        rel = o.rel_init;
        add = o.add_init;
        set_ivar(rel, add);
        // EmDataPoint.find_best_halfspace :148-186: argmin misfit over logspace(-4, 4, 100)
        if (lane == 0) {
            w.edges_c[0] = 0.0;
            w.edges_c[1] = INFINITY;
        }
        __syncwarp();
        double best = INFINITY, best_c = 0.0;
        #pragma unroll 1
        for (int i = 0; i < 100; ++i) {
            double e = (i == 99) ? 4.0 : -4.0 + (double)i * (8.0 / 99.0);
            double c = dexp_(e * 2.302585092994045684017991454684);
            if (lane == 0) w.sig_c[0] = c;
            __syncwarp();
            forward(1, w.sig_c, w.edges_c, w.pred_c);
            double mis, like;
            misfit_likelihood(w.pred_c, &mis, &like);
            if (mis < best) {
                best = mis;
                best_c = c;
            }
        }
        sigma_ref = best_c;
        ln_ref = dlog_(sigma_ref);
        k = 1;
        if (lane == 0) w.sig_c[0] = sigma_ref;
        __syncwarp();
        forward_sens(1, w.sig_c, w.edges_c, w.pred_c, w.Jc);
        // posterior grids (Model.set_posteriors :666-684, DataPoint.set_*_error_posterior :668-695)
        const double s = dlog_(1.0 + o.factor);
        sig_lo = ln_ref - o.sigma_bins_nstd * s;
        sig_dx = 2.0 * o.sigma_bins_nstd * s / (double)o.n_sigma_bins;
        rel_lo = dlog_(o.rel_min);
        rel_dx = (dlog_(o.rel_max) - dlog_(o.rel_min)) / (double)o.n_err_bins;
        add_lo = dlog_(o.add_min);
        add_dx = (dlog_(o.add_max) - dlog_(o.add_min)) / (double)o.n_err_bins;
        depth_step = 0.5 * o.min_width;
        if (!first) {  // reset(): posteriors and traces are re-created
            zero_posteriors();
            const size_t N2 = 2 * (size_t)o.n_markov_chains;
            if (P.out.misfit_trace)
                #pragma unroll 1
                for (size_t i = lane; i < N2; i += 32) P.out.misfit_trace[(size_t)chain * N2 + i] = 0.0;
            if (P.out.accept_trace)
                #pragma unroll 1
                for (size_t i = lane; i < N2; i += 32) P.out.accept_trace[(size_t)chain * N2 + i] = 0;
            __syncwarp();
        }
        misfit_likelihood(w.pred_c, &misfit, &likelihood);
        prior = model_probability(1, w.sig_c, w.edges_c) + datapoint_probability(rel, add);
        posterior = likelihood + prior;
        burned_in = 0;
        burned_in_iter = 0;
        iteration = 0;
        if (P.out.misfit_trace && lane == 0) P.out.misfit_trace[(size_t)chain * 2 * o.n_markov_chains] = misfit;
        save_best();
        n_zero = 0;
        acc_win = 0;
        dwell = 0;
    }

    // ------------------------------------------------------------ Inference1D.accept_reject
    // returns true if the chain failed (Gauss-Newton matrix not positive definite)
    __device__ __noinline__ bool step(bool* accepted_out)
    {
        const gbp_options& o = P.opt;
        *accepted_out = false;
        int kn;
        const int action = perturb_structure(&kn);
        if (action == 0) n_act0++;
        else if (action == 1) n_act1++;
        else if (action == 2) n_act2++;
        else n_act3++;

        const T* Jh = w.Jc;
        const T* ph = w.pred_c;
        if (action != ACT_NONE) {  // observation.fm_dlogc(remapped_model)
            forward_sens(kn, w.sig_r, w.edges_p, w.pred_t, w.Jt);
            Jh = w.Jt;
            ph = w.pred_t;
        }
        set_ivar(rel, add);
        prior_setup(kn, w.sig_r, w.edges_p);
        const double ln_r = (lane < kn) ? dlog_(w.sig_r[lane]) : 0.0;
        const double g = gradient_lane(kn, Jh, ph);
        assemble(kn, Jh);
        double logdetL;
        if (!cholesky(kn, &logdetL)) return true;
        const double stepv = solve_LT(kn, solve_L(kn, g));       // H * dfk
        const double mean = ln_r - o.covariance_scaling * stepv;  // ln sigma + alpha * pk, pk = -H dfk
        // sigma' ~ dexp_(N(mean, H)),  H = (L L')^-1  ->  mean + L^-T z
        double z0 = 0.0, z1 = 0.0;
        const int npair = (kn + 1) / 2;
        if (lane < npair) normal2_at(rng, rng.block + (unsigned long long)lane, &z0, &z1);
        rng.block += (unsigned long long)npair;
        const double za = __shfl_sync(FULL, z0, lane >> 1), zb = __shfl_sync(FULL, z1, lane >> 1);
        const double zi = (lane < kn) ? ((lane & 1) ? zb : za) : 0.0;
        const double dx = solve_LT(kn, zi);
        const double ln_t = mean + dx;
        if (lane < kn) w.sig_t[lane] = dexp_(ln_t);
        __syncwarp();

        // test_datapoint.perturb() (DataPoint.py:531-573)
        double rel_t = rel, add_t = add;
        if (o.solve_relative_error) rel_t = propose_error(rng, rel, o.rel_prop_var, o.rel_min, o.rel_max);
        if (o.solve_additive_error) add_t = propose_error(rng, add, o.add_prop_var, o.add_min, o.add_max);

        const bool jump = (action == ACT_BIRTH || action == ACT_DEATH);
        // forward at the candidate; for birth/death the Jacobian at the candidate is needed as well
        // (Model.proposal_probabilities :619) - fused into the same pass.
        if (jump) forward_sens(kn, w.sig_t, w.edges_p, w.pred_t, w.Jt);
        else forward(kn, w.sig_t, w.edges_p, w.pred_t);
        set_ivar(rel_t, add_t);
        double t_misfit, t_like;
        misfit_likelihood(w.pred_t, &t_misfit, &t_like);
        double t_prior = datapoint_probability(rel_t, add_t);
        if (t_prior == -INFINITY) return false;
        t_prior += model_probability(kn, w.sig_t, w.edges_p);
        if (t_prior == -INFINITY) return false;

        double proposal = 1.0, proposal1 = 1.0;
        if (jump) {
            prior_setup(kn, w.sig_t, w.edges_p);  // lns <- ln sigma' - ln ref (t2 unchanged: same mesh)
            const double g2 = gradient_lane(kn, w.Jt, w.pred_t);
            const double s2 = solve_LT(kn, solve_L(kn, g2));  // H dfk'
            const double lv = ln_t + o.covariance_scaling * s2;  // Model.py:626 (sign as in the reference)
            const double mv = dexp_(lv);
            const int bad = __any_sync(FULL, lane < kn && (mv == INFINITY || mv == 0.0));
            const double q_r = quad(kn, (lane < kn) ? (ln_r - lv) : 0.0);
            const double q_f = quad(kn, (lane < kn) ? (ln_t - ln_r) : 0.0);
            if (bad) {
                proposal = -INFINITY;
                proposal1 = -INFINITY;
            } else {
                proposal = -(0.5 * kn) * LOG2PI + logdetL - 0.5 * q_r;
                proposal1 = -(0.5 * kn) * LOG2PI + logdetL - 0.5 * q_f;
            }
        }
        const double log_alpha = (t_prior - prior) + (t_like - likelihood) + (proposal - proposal1);
        const double u = rng_uniform(rng);
        const bool acc = dexp_(log_alpha) > u;
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
            if (lane < kn) w.sig_c[lane] = w.sig_t[lane];
            if (lane <= kn) w.edges_c[lane] = w.edges_p[lane];
            if (lane < C) w.pred_c[lane] = w.pred_t[lane];
            if (action != ACT_NONE)
                #pragma unroll 1
                for (int i = lane; i < C * KS; i += 32) w.Jc[i] = w.Jt[i];
            __syncwarp();
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
            P.out.misfit_trace[(size_t)chain * N2 + (iteration - 1)] = misfit;
        if (!burned_in && iteration > o.burn_in_min_iter && misfit < (double)n_active) {
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
    __device__ void run(int chain_)
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
            w.data[lane] = act ? d : 0.0;
        }
        n_active = warp_sum_i(act);
        n_accept = n_forward = n_sens = 0;
        n_act0 = n_act1 = n_act2 = n_act3 = 0;
        n_resets = 0;
        limiters = 0;
        __syncwarp();
        initialize(true);

        bool failed = (n_active == 0);
        bool go = !failed;
        long long total = 0;
        const long long N = o.n_markov_chains;
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

        const gbp_chain_buffers& ob = P.out;
        const int ml = o.max_layers;
        if (ob.cur_sigma && lane < ml) ob.cur_sigma[(size_t)chain * ml + lane] = lane < k ? w.sig_c[lane] : NAN;
        if (ob.cur_edges)
            #pragma unroll 1
            for (int i = lane; i <= ml; i += 32) ob.cur_edges[(size_t)chain * (ml + 1) + i] = i <= k ? w.edges_c[i] : NAN;
        if (lane == 0) {
            double* s = ob.scalars + (size_t)chain * GBP_NSCALARS;
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
            s[GBP_S_BEST_POSTERIOR] = best_posterior;
            s[GBP_S_CUR_REL] = rel;
            s[GBP_S_CUR_ADD] = add;
            s[GBP_S_CUR_MISFIT] = misfit;
            s[GBP_S_CUR_PRIOR] = prior;
            s[GBP_S_CUR_LIKELIHOOD] = likelihood;
            s[GBP_S_BEST_REL] = best_rel;
            s[GBP_S_BEST_ADD] = best_add;
            s[GBP_S_N_RESETS] = n_resets;
            s[GBP_S_N_BIRTH] = (double)n_act0;
            s[GBP_S_N_DEATH] = (double)n_act1;
            s[GBP_S_N_MOVE] = (double)n_act2;
            s[GBP_S_N_NONE] = (double)n_act3;
        }
        __syncwarp();
    }
};

// ---------------------------------------------------------------- kernels
template <typename T, int NC>
__global__ void __launch_bounds__(512, 1) rjmcmc_kernel(const __grid_constant__ SysDev S, const T* __restrict__ g_tab,
                                                         const __grid_constant__ ChainParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    T* tab = reinterpret_cast<T*>(smem);
    const uint32_t tab_bytes = (uint32_t)(TAB_ROWS * S.tab_stride * sizeof(T));
    tma_stage(tab, g_tab, tab_bytes, &bar);
    const uint32_t tab_pad = (tab_bytes + 127u) & ~127u;
    const int warp = threadIdx.x >> 5;
    WarpState<T, NC>* ws = reinterpret_cast<WarpState<T, NC>*>(smem + tab_pad) + warp;
    Chain<T, NC> ch(*ws, S, tab, P);
    // persistent: the first wave is assigned statically, later chains come from a device-side counter
    int c = blockIdx.x * (blockDim.x >> 5) + warp;
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
    T* tab = reinterpret_cast<T*>(smem);
    const uint32_t tab_bytes = (uint32_t)(TAB_ROWS * S.tab_stride * sizeof(T));
    tma_stage(tab, g_tab, tab_bytes, &bar);
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
        fdem_eval<T, SENS>(S, tab, (T)altitude[b], L, msig, mthk, pred, SENS ? J : nullptr);
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
