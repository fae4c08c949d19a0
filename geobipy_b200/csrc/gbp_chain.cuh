// One warp = one Markov chain: the fused trans-dimensional MCMC sampler for one sounding - a frequency-domain
// datapoint (FdemDataPoint, KIND_FDEM) or a time-domain one with one or two systems (TdemDataPoint, KIND_TDEM).
//
// Replaces the Python loop Inference1D.initialize + infer (geobipy/src/inversion/Inference1D.py:353-464,
// :537-631 accept_reject, :633-688 infer, :705-790 update) together with everything it calls:
//   RectilinearMesh1D.perturb            classes/mesh/RectilinearMesh1D.py:993-1120
//   Model.stochastic_newton_perturbation classes/model/Model.py:368-419 (+ :250-272, :347-357, :421-430)
//   Model.probability / gradient_probability / proposal_probabilities  Model.py:533-575, :213-234, :577-660
//   DataPoint.std / data_misfit / likelihood / probability / perturb   classes/data/datapoint/DataPoint.py
//   TdemDataPoint.std (per-system errors, additive error x (t / 1 ms)^-1/2)  classes/data/datapoint/TdemDataPoint.py:329-379
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
//  * instruction footprint is a first-class concern (ncu: the first version, 532 KB of SASS, and every
//    later one were bound by instruction fetch, `stall_no_instruction`): helpers that are called from
//    several places are single non-inlined copies taking shared-memory pointers; the once-per-iteration
//    control code is inlined into the kernel so that its state stays in registers and the options come
//    from the constant bank; cold per-chain counters live in shared memory;
//  * every accept_reject step draws from its own Philox sub-stream, so warps that have run out of chains
//    evaluate FUTURE iterations of the chains still running on their SM speculatively ("speculative evaluation"
//    below): bit-identical results, the tail of a batch shrinks by a third.
#pragma once
#include "gbp_fdem.cuh"
#include "gbp_fdem_f2.cuh"
#include "gbp_tdem.cuh"
#include "gbp_tdem_f2.cuh"

namespace gbp {

constexpr int NPACK = GBP_MAXL * (GBP_MAXL + 1) / 2;
enum { ACT_BIRTH = 0, ACT_DEATH = 1, ACT_MOVE = 2, ACT_NONE = 3 };
// cold per-chain integers kept in shared memory
enum { CT_N_ACCEPT = 0, CT_N_FWD, CT_N_SENS, CT_ACT0, CT_ACT1, CT_ACT2, CT_ACT3, CT_BEST_K, CT_BEST_ITER, CT_BURN_ITER,
       CT_N_ZERO, CT_N_RESETS, CT_LIMITERS, CT_ACC_WIN, CT_TO_PLOT, CT_N_SPEC, CT_N = 16 };
enum { BV_POSTERIOR = 0, BV_REL, BV_ADD, BV_REL2, BV_ADD2, BV_N = 8 };

// KIND of datapoint a kernel instantiation inverts: frequency domain (FdemDataPoint, one system) or time
// domain (TdemDataPoint, one or two systems with per-system errors, TdemDataPoint.py:329-379)
// KIND_FDEM_Z: a frequency-domain datapoint whose sensor height is sampled too (solve_z: Point.perturb :614-622,
// Point.set_priors :959-961).  Its own instantiation, so that the fixed-height kernels are what they were.
// KIND_TDEM_Z: a time-domain datapoint whose TRANSMITTER height is sampled (the options file's solve_transmitter_z:
// TdemDataPoint.perturb :681-683 -> Loop_pair.perturb -> Point.perturb on the transmitter loop; prior added after the
// error priors, TdemDataPoint.probability :950-951; fixed receiver offset).
// KIND_TEMPEST: a fixed-wing Tempest_datapoint (classes/data/datapoint/Tempest_datapoint.py): X and Z components (two
// passes of tdem_eval, J1 / J0 Hankel weights), data = secondary + primary field (:107-127), one relative error and one
// additive-error multiplier per COMPONENT over fixed per-channel additive levels (:141-176) - the two-"system" error arrays
// of KIND_TDEM indexed by component, the per-channel scale table holding the additive levels - a prior that counts the
// relative errors only (:478-488), and a multiplier re-drawn around its INITIAL value every step (:339-341).
enum { KIND_FDEM = 0, KIND_TDEM = 1, KIND_FDEM_Z = 2, KIND_TDEM_Z = 3, KIND_TEMPEST = 4 };
__host__ __device__ constexpr bool is_td(int kind) { return kind == KIND_TDEM || kind == KIND_TDEM_Z || kind == KIND_TEMPEST; }
__host__ __device__ constexpr bool has_z(int kind) { return kind == KIND_FDEM_Z || kind == KIND_TDEM_Z; }
__host__ __device__ constexpr int ns_of(int kind) { return is_td(kind) ? GBP_TD_MAXSYS : 1; }
template <typename T, int KIND> struct SysOf {
    typedef SysShared<T> shared;
    typedef SysDev dev;
};
template <typename T> struct SysOf<T, KIND_TDEM> {
    typedef TdShared<T> shared;
    typedef TdDev dev;
};
template <typename T> struct SysOf<T, KIND_TDEM_Z> {
    typedef TdShared<T> shared;
    typedef TdDev dev;
};
template <typename T> struct SysOf<T, KIND_TEMPEST> {
    typedef TdShared<T> shared;
    typedef TdDev dev;
};
// per-chain scratch of the forward operator
template <typename T, int KIND> struct FwdExtra {};
template <typename T> struct FwdExtra<T, KIND_TDEM> {
    T lam[1][GBP_TD_MAXLAM], wgt[1][GBP_TD_MAXLAM];  // this sounding's Hankel abscissae and geometry weights
    T sbuf[TD_ROWS];
    T alt[1];
    int cur;
};
// sampled transmitter height: the abscissae / weights depend on it, so there are two sets - the current height's
// (index cur) and the proposed height's, swapped on acceptance
template <typename T> struct FwdExtra<T, KIND_TDEM_Z> {
    T lam[2][GBP_TD_MAXLAM], wgt[2][GBP_TD_MAXLAM];
    T sbuf[TD_ROWS];
    T alt[2];   // height each set was computed for
    int cur;
};
// X and Z components: the same abscissae under two weight sets (wgt[0]: J0, the z channels; wgt[1]: J1 dx / r, the x channels)
template <typename T> struct FwdExtra<T, KIND_TEMPEST> {
    T lam[1][GBP_TD_MAXLAM], wgt[2][GBP_TD_MAXLAM];
    T sbuf[TD_ROWS];
    T alt[1];
    int cur;
};
// relative / additive errors, one per system (registers)
template <typename R, int NS> struct Errs {
    R rel[NS], add[NS];
};
// outcome of a speculatively evaluated step that would be ACCEPTED: what the owner needs to adopt the proposed state
template <typename R, int NS> struct SpecOut {
    int kn, mp, vp, changed;
    R misfit, prior, likelihood;
    Errs<R, NS> err, ln_err;
    R alt;   // proposed sensor height (KIND_FDEM_Z)
};
enum { OP_HITMAP = 0, OP_EDGES, OP_NCELLS, OP_REL, OP_ADD, OP_MISFIT, OP_ACCEPT, OP_HEIGHT, OP_N = 8 };

template <typename R> struct MeshBuf {
    R edges[GBP_MAXL + 2];  // edges[0] = 0, edges[k] = inf
    R lnh[GBP_MAXL];        // ln(thickness_i)           (gradient prior, RectilinearMesh1D.py:713)
    R t2[GBP_MAXL];         // gradient-operator weights (RectilinearMesh1D.py:747-786) / g^2
};
template <typename R> struct ValBuf {
    R sig[GBP_MAXL];        // conductivity
    R ls[GBP_MAXL];         // ln(conductivity)
};

// A TEAM of warps (2 or 4, on the same SM sub-partition) shares its forward evaluations: every member brings at most
// one request per round, the requests are cut into work units of one frequency each (so the result does not depend on
// who computes what) and the n_freq x team_size units are claimed dynamically, the longest first.  All members therefore
// run the same phase of accept_reject at the same time - the sampler code between two rounds is fetched once per
// team instead of once per warp (the kernel is bound by instruction fetch: profiles/README.md, round 2) - and a
// member without a chain keeps evaluating forwards for its team mates.
struct TeamShared {
    volatile int alive;   // members that still own a chain
    int unit;             // next work unit of the current round
    int freq_units;
    void* member[16];     // WarpState of member m
};

template <typename R, typename T, int NC, int KIND> struct __align__(16) WarpState {
    R A[NPACK];             // packed lower triangle: Gauss-Newton matrix, then its Cholesky factor
    MeshBuf<R> mesh[2];
    ValBuf<R> val[2];
    R ls_r[GBP_MAXL];       // ln(sigma) of the remapped model
    R vec[GBP_MAXL + 2];
    R data[NC], ivar[NC];   // observed data (0 where inactive), 1/variance (0 where inactive)
    alignas(16) T J[NC * KS];  // working Jacobian; the current model's Jacobian is mirrored in HBM/L2 (ChainParams.jstore)
    T pred[2][NC];
    T msig[KS], mthk[KS];
    int sbin[GBP_MAXL + 2];
    int ctr[CT_N];          // cold counters
    R bestv[BV_N];          // best posterior / errors
    void* outp[OP_N];       // this chain's output rows
    SpecOut<R, ns_of(KIND)> sout;  // written by a speculative step that ends in an acceptance
    FwdExtra<T, KIND> fx;
    // team mode (nullptr: this warp evaluates its forwards alone) and this warp's request of the current round
    TeamShared* team;
    int tm_size, tm_bar;
    int req_active, req_kk, req_sens;
    T req_alt;
    T* req_pred;
    T* req_J;
};

// Option-derived constants, computed once per CTA in fp64 and shared by its warps.
template <typename R> struct Consts {
    R cum0, cum1, cum2;                       // cumulative event probabilities
    R ln_min_edge, ln_edge_span, min_edge, max_edge, min_width;
    R inv_s2, inv_g2, alpha;                  // 1/ln(1+factor)^2, 1/grad_std^2, covariance_scaling
    R lp_k;                                   // -ln(kmax - 1)
    R c_grad, c_val;                          // per-dimension constants of the gradient / value prior
    R half_log2pi;
    R rel_lnmin[2], rel_lnmax[2], rel_sd[2], rel_ln0[2], rel0[2], rel_dx[2];   // per system
    R add_lnmin[2], add_lnmax[2], add_sd[2], add_ln0[2], add0[2], add_dx[2];
    R err_lp;                                 // log prior of the (always in-bounds) errors, all systems, and height
    R z_sd, z_max, z_dx;                      // solve_z: proposal std, half-width of the Uniform prior, bin width
    R sig_halfspan, sig_dx, depth_step, depth_max;
    R ln_half, ln_3half;
    int kmax, n_depth, n_sig, n_err, C, solve_par, solve_grad, solve_rel, solve_add;
    int n_chains, upe, burn_min, reset_limit;
    int n_sys;
    R tsc[GBP_TD_MAXC];                       // additive-error scale of channel c (time domain), else unused
    unsigned char csys[GBP_TD_MAXC];          // system of channel c
};

struct ChainParams {
    gbp_options opt;
    int B, n_depth, C, n_warps_total;
    const double* data;      // [B][C]
    const double* altitude;  // [B]
    unsigned long long seed, first_index;
    long long max_iterations;
    gbp_chain_buffers out;
    double data_scale;       // observed data and additive errors are multiplied by this on the way in (fp32
                             // time-domain path: 2^40, so that squares of 1e-15 V/Am^4 stay normal numbers)
    int team_size;           // warps per team (1 = every warp evaluates its own forwards; 2 or 4: team_round)
    int team_spread;         // 1: the members of a team sit on different SM sub-partitions (warps t*T .. t*T + T - 1)
    int team_freq_units;     // 1: work units of one frequency; 0: two halves per request
    int spec_helpers;        // max warps that evaluate future iterations of one chain speculatively (0 = off)
    int spec_min_rejections; // a chain speculates once it has rejected this many steps in a row
    int* work_counter;
    void* jstore;            // [B][NC*KS] of T: Jacobian of each chain's current model
    unsigned long long* finish_ns;  // debug (GBP_DEBUG_TIMELINE): [B + 1] %globaltimer at the end of each chain, [B] = earliest start
};

template <typename R>
__device__ __noinline__ void make_consts(const gbp_options& o, int n_depth, int C, Consts<R>& c, const bool height_last, const bool add_prior = true)
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
    c.n_sys = o.n_systems > 1 ? 2 : 1;
    // Point.probability :160-196 comes first in DataPoint.probability: Uniform(z0 - dz, z0 + dz), always in bounds
    // (a time-domain datapoint adds the height prior AFTER the error priors, TdemDataPoint.probability :950-951: the
    // order only matters for the last bit of the fp64 sum, which the trajectory twin of the oracle shares)
    double lp = (o.solve_height && !height_last) ? -dlog_(2.0 * o.max_height_change) : 0.0;
    c.z_sd = (R)::sqrt(o.height_prop_var);
    c.z_max = (R)o.max_height_change;
    c.z_dx = (R)(2.0 * o.max_height_change / (double)o.n_err_bins);
    for (int i = 0; i < c.n_sys; ++i) {
        const double rmin = i ? o.rel_min2 : o.rel_min, rmax = i ? o.rel_max2 : o.rel_max;
        const double amin = i ? o.add_min2 : o.add_min, amax = i ? o.add_max2 : o.add_max;
        c.rel_lnmin[i] = (R)dlog_(rmin);
        c.rel_lnmax[i] = (R)dlog_(rmax);
        c.rel_sd[i] = (R)::sqrt(i ? o.rel_prop_var2 : o.rel_prop_var);
        c.rel_ln0[i] = (R)dlog_(i ? o.rel_init2 : o.rel_init);
        c.rel0[i] = (R)(i ? o.rel_init2 : o.rel_init);
        c.rel_dx[i] = (R)((dlog_(rmax) - dlog_(rmin)) / (double)o.n_err_bins);
        c.add_lnmin[i] = (R)dlog_(amin);
        c.add_lnmax[i] = (R)dlog_(amax);
        c.add_sd[i] = (R)::sqrt(i ? o.add_prop_var2 : o.add_prop_var);
        c.add_ln0[i] = (R)dlog_(i ? o.add_init2 : o.add_init);
        c.add0[i] = (R)(i ? o.add_init2 : o.add_init);
        c.add_dx[i] = (R)((dlog_(amax) - dlog_(amin)) / (double)o.n_err_bins);
        // Uniform(log=True) priors: the proposals are forced inside their bounds (or fall back to the current
        // values), so DataPoint.probability (:351-395) is this constant
        if (o.solve_relative_error) lp += -dlog_(dlog_(rmax) - dlog_(rmin));
        // (Tempest: the additive-error multiplier's prior only gives its histogram its bins, DataPoint.probability never sees it)
        if (o.solve_additive_error && add_prior) lp += -dlog_(dlog_(amax) - dlog_(amin));
    }
    if (c.n_sys == 1) {
        c.rel_lnmin[1] = c.rel_lnmin[0]; c.rel_lnmax[1] = c.rel_lnmax[0]; c.rel_sd[1] = c.rel_sd[0];
        c.rel_ln0[1] = c.rel_ln0[0]; c.rel0[1] = c.rel0[0]; c.rel_dx[1] = c.rel_dx[0];
        c.add_lnmin[1] = c.add_lnmin[0]; c.add_lnmax[1] = c.add_lnmax[0]; c.add_sd[1] = c.add_sd[0];
        c.add_ln0[1] = c.add_ln0[0]; c.add0[1] = c.add0[0]; c.add_dx[1] = c.add_dx[0];
    }
    if (o.solve_height && height_last) lp += -dlog_(2.0 * o.max_height_change);
    c.err_lp = (R)lp;
    for (int i = 0; i < GBP_TD_MAXC; ++i) {
        c.tsc[i] = R(1);
        c.csys[i] = 0;
    }
    c.sig_halfspan = (R)(o.sigma_bins_nstd * s);
    c.sig_dx = (R)(2.0 * o.sigma_bins_nstd * s / (double)o.n_sigma_bins);
    c.depth_step = (R)(0.5 * o.min_width);
    c.depth_max = (R)((double)n_depth * 0.5 * o.min_width);
    c.ln_half = (R)dlog_(0.5);
    c.ln_3half = (R)dlog_(1.5);
    c.kmax = o.max_layers;
    c.n_depth = n_depth;
    c.n_sig = o.n_sigma_bins;
    c.n_err = o.n_err_bins;
    c.C = C;
    c.solve_par = o.solve_parameter;
    c.solve_grad = o.solve_gradient;
    c.solve_rel = o.solve_relative_error;
    c.solve_add = o.solve_additive_error;
    c.n_chains = o.n_markov_chains;
    c.upe = o.update_plot_every;
    c.burn_min = o.burn_in_min_iter;
    c.reset_limit = o.reset_limit;
}

__device__ __forceinline__ int pk(int i, int j) { return i * (i + 1) / 2 + j; }
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

#define GBP_SHARED(p) __builtin_assume(__isShared(p))

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

// ================================================================ shared, non-inlined helpers
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One round of a team: every member arrives with its request (req_active = 0: none) in its WarpState.  Returns false
// when no member owns a chain any more (the team dissolves; nothing is computed in that round).
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ bool team_round(WarpState<R, T, NC, KIND>* w, const typename SysOf<T, KIND>::shared* S, const T* tab)
{
    if constexpr (!is_td(KIND) && sizeof(T) == 4) {
        GBP_SHARED(w);
        TeamShared* tm = w->team;
        GBP_SHARED(tm);
        const int lane = lane_id();
        const int tsz = w->tm_size, nthreads = 32 * tsz, bar = w->tm_bar, fu = tm->freq_units;
        const int tlog = 31 - __clz(tsz);   // team sizes are powers of two
        const int n_units = (fu ? S->n_freq : 2) << tlog;
        __syncwarp();
        named_bar_sync(bar, nthreads);   // requests (and `alive`) of all members are visible
        if (tm->alive <= 0) return false;
#pragma unroll 1
        for (;;) {
            int u = 0;
            if (lane == 0) u = atomicAdd(&tm->unit, 1);
            u = __shfl_sync(FULL, u, 0);
            if (u >= n_units) break;
            // unit u = (frequency rank u / T, request of member u % T): the frequencies with most chunks go first
            WarpState<R, T, NC, KIND>* wi = (WarpState<R, T, NC, KIND>*)tm->member[u & (tsz - 1)];
            GBP_SHARED(wi);
            if (!wi->req_active) continue;
            const int ur = u >> tlog;
            const int c0 = fu ? (int)S->unit_begin[ur] : S->half_begin[ur];
            const int c1 = fu ? c0 + (int)S->unit_count[ur] : S->half_begin[ur + 1];
            if (wi->req_sens) fdem_sens_f2(*S, tab, wi->req_alt, wi->req_kk, wi->msig, wi->mthk, wi->req_pred, wi->req_J, c0, c1);
            else fdem_fwd_f2(*S, tab, wi->req_alt, wi->req_kk, wi->msig, wi->mthk, wi->req_pred, c0, c1);
        }
        named_bar_sync(bar, nthreads);   // results are visible to their owners
        if (lane == 0 && tm->member[0] == (void*)w) tm->unit = 0;   // ordered before the next round's first barrier
        return true;
    } else {
        return false;
    }
}

// a round without a request (keeps the members of a team in the same phase: every accept_reject step is two rounds)
template <typename R, typename T, int NC, int KIND>
__device__ __forceinline__ void ch_round_pad(WarpState<R, T, NC, KIND>* w, const typename SysOf<T, KIND>::shared* S, const T* tab)
{
    if constexpr (!is_td(KIND) && sizeof(T) == 4) {
        if (w->team) {
            if (lane_id() == 0) w->req_active = 0;
            team_round<R, T, NC, KIND>(w, S, tab);
        }
    }
}

// J == nullptr: forward only.  Otherwise forward + Jacobian in one pass (FdemDataPoint.fm_dlogc :535)
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_forward(WarpState<R, T, NC, KIND>* w, const typename SysOf<T, KIND>::shared* S, const T* tab,
                                        T alt, int kk, const R* sig, const R* edges, T* pred, T* J)
{
    GBP_SHARED(w);
    GBP_SHARED(sig);
    GBP_SHARED(edges);
    const int lane = lane_id();
    if (lane < kk) {
        w->msig[lane] = (T)sig[lane];
        w->mthk[lane] = (T)(edges[lane + 1] - edges[lane]);
    }
    if (lane == 0) {
        w->ctr[CT_N_FWD]++;
        if (J) w->ctr[CT_N_SENS]++;
    }
    __syncwarp();
    if constexpr (is_td(KIND)) {
        // the geometry set of the height this evaluation is for (KIND_TDEM_Z: current or proposed)
        int g = 0;
        if constexpr (KIND == KIND_TDEM_Z) g = (alt == w->fx.alt[w->fx.cur]) ? w->fx.cur : (w->fx.cur ^ 1);
        if constexpr (KIND == KIND_TEMPEST) {
            // one pass per component (each writes the channels of its component), then the predicted primary field:
            // predictedData = predicted secondary + primary (Tempest_datapoint.py:120-127)
            tdem_eval<T>(*S, tab, w->fx.lam[0], w->fx.wgt[0], kk, w->msig, w->mthk, w->fx.sbuf, pred, J, J != nullptr, S->ccomp, 0);
            tdem_eval<T>(*S, tab, w->fx.lam[0], w->fx.wgt[1], kk, w->msig, w->mthk, w->fx.sbuf, pred, J, J != nullptr, S->ccomp, 1);
#pragma unroll 1
            for (int c = lane; c < S->C; c += 32) pred[c] += S->poff[c];
            __syncwarp();
        } else {
            tdem_eval<T>(*S, tab, w->fx.lam[g], w->fx.wgt[g], kk, w->msig, w->mthk, w->fx.sbuf, pred, J, J != nullptr);
        }
    } else {
        if constexpr (sizeof(T) == 4) {
            if (w->team) {   // the team evaluates it (and this warp its share of the team's other requests)
                if (lane == 0) {
                    w->req_active = 1;
                    w->req_kk = kk;
                    w->req_sens = J != nullptr;
                    w->req_alt = alt;
                    w->req_pred = pred;
                    w->req_J = J;
                }
                team_round<R, T, NC, KIND>(w, S, tab);
                return;
            }
        }
        fdem_run<T>(*S, tab, alt, kk, w->msig, w->mthk, pred, J, J != nullptr);
    }
}

// DataPoint.std :268-282 -> 1/variance per active channel (EmDataPoint.active :44-56)
// TdemDataPoint.std :329-379: per-system errors, additive error of channel c times (t_c / 1 ms)^-0.5
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_set_ivar(WarpState<R, T, NC, KIND>* w, const Consts<R>* K, int C, Errs<R, ns_of(KIND)> e)
{
    GBP_SHARED(w);
    GBP_SHARED(K);
    const int lane = lane_id();
#pragma unroll
    for (int p = 0; p < (NC + 31) / 32; ++p) {
        const int c = lane + 32 * p;
        if (c < C) {
            R d = w->data[c];
            R r = e.rel[0], a = e.add[0];
            if constexpr (is_td(KIND)) {
                if (K->csys[c]) {
                    r = e.rel[ns_of(KIND) - 1];
                    a = e.add[ns_of(KIND) - 1];
                }
                a *= K->tsc[c];
            }
            R s = r * d;
            w->ivar[c] = (d > R(0)) ? R(1) / (s * s + a * a) : R(0);
        }
    }
    __syncwarp();
}

// misfit (DataPoint.py:502-525) and Gaussian log-likelihood (MvNormalDistribution.py:209-216)
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ pair_t<R> ch_misfit_like(WarpState<R, T, NC, KIND>* w, int C, R n_active_half_log2pi, const T* pred)
{
    GBP_SHARED(w);
    GBP_SHARED(pred);
    const int lane = lane_id();
    R q = R(0), ld = R(0);
#pragma unroll
    for (int p = 0; p < (NC + 31) / 32; ++p) {
        const int c = lane + 32 * p;
        if (c < C) {
            R iv = w->ivar[c];
            if (iv > R(0)) {
                R r = (R)pred[c] - w->data[c];
                q += r * r * iv;
                ld += -rt<R>::log(iv);
            }
        }
    }
    q = warp_sum(q);
    ld = warp_sum(ld);
    return pair_t<R>{q, -n_active_half_log2pi - R(0.5) * ld - R(0.5) * q};
}

// Model.probability :533-575 (value_bounds = None); ls = ln sigma, lnh = ln thickness
template <typename R>
__device__ __noinline__ R ch_model_prob(const Consts<R>* K, int kk, const R* ls, const R* lnh, R ln_ref)
{
    GBP_SHARED(K);
    GBP_SHARED(ls);
    GBP_SHARED(lnh);
    const int lane = lane_id();
    R p = (kk >= 1 && kk <= K->kmax) ? K->lp_k : (R)-INFINITY;
    if (K->solve_par) {
        R q = R(0);
        if (lane < kk) {
            R d = ls[lane] - ln_ref;
            q = d * d * K->inv_s2;
        }
        p += (R)kk * K->c_val - R(0.5) * warp_sum(q);
    }
    if (K->solve_grad) {
        if (kk == 1) {
            p += K->c_grad;  // Model.py:230-232: a virtual 2-layer model with equal values
        } else {
            R q = R(0);
            if (lane < kk - 1) {
                R g = (ls[lane + 1] - ls[lane]) / lnh[lane];
                q = g * g * K->inv_g2;
            }
            p += (R)(kk - 1) * K->c_grad - R(0.5) * warp_sum(q);
        }
    }
    return p;
}

// lnh[i] = ln(thickness_i), t2[i] = 1/(g^2 (c2c_i (k-1))^2)  (RectilinearMesh1D.gradient_operator :747-786)
template <typename R> __device__ __noinline__ void ch_mesh_setup(const Consts<R>* K, int kk, MeshBuf<R>* m)
{
    GBP_SHARED(K);
    GBP_SHARED(m);
    const int lane = lane_id();
    if (kk >= 2 && lane < kk - 1) {
        const R* e = m->edges;
        R x0 = e[lane + 1] - e[lane];
        R x1;
        if (lane + 1 < kk - 1) x1 = e[lane + 2] - e[lane + 1];
        else x1 = (kk == 2) ? x0 : (e[kk - 1] - e[kk - 2]) + (e[kk - 1] - e[0]);
        R c2c = R(0.5) * (x0 + x1);
        R t = R(1) / (c2c * (R)(kk - 1));
        m->t2[lane] = t * t * K->inv_g2;
        m->lnh[lane] = rt<R>::log(x0);
    }
    __syncwarp();
}

template <typename R> __device__ __forceinline__ R prior_op(const Consts<R>* K, int kk, const R* t2, int i, int j)
{
    if (kk == 1) return K->inv_s2 + K->inv_g2;  // gradient_operator = ones((1,1))
    if (i == j) {
        R d = K->inv_s2;
        if (i > 0) d += t2[i - 1];
        if (i < kk - 1) d += t2[i];
        return d;
    }
    if (i == j + 1) return -t2[j];
    if (j == i + 1) return -t2[i];
    return R(0);
}

// gradient of lane i: Wm'Wm (ln s - ln ref) + J' Wd'Wd (pred - d)   (Model.local_gradient :347-357)
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ R ch_gradient(WarpState<R, T, NC, KIND>* w, const Consts<R>* K, int kk, const R* t2, const R* ls,
                                      const T* J, const T* pred, R ln_ref)
{
    GBP_SHARED(w);
    GBP_SHARED(K);
    GBP_SHARED(t2);
    GBP_SHARED(ls);
    GBP_SHARED(J);
    GBP_SHARED(pred);
    const int lane = lane_id();
    const int C = K->C;
    R g = R(0);
    if (lane < kk) {
        g = prior_op(K, kk, t2, lane, lane) * (ls[lane] - ln_ref);
        if (lane > 0) g += prior_op(K, kk, t2, lane, lane - 1) * (ls[lane - 1] - ln_ref);
        if (lane < kk - 1) g += prior_op(K, kk, t2, lane, lane + 1) * (ls[lane + 1] - ln_ref);
#pragma unroll 1
        for (int c = 0; c < C; ++c) g += (R)J[c * KS + lane] * (((R)pred[c] - w->data[c]) * w->ivar[c]);
    }
    return g;
}

// A = Wm'Wm + J' Wd'Wd J (Model.local_precision :250-272), packed lower triangle
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_assemble(WarpState<R, T, NC, KIND>* w, const Consts<R>* K, int kk, const R* t2, const T* J)
{
    GBP_SHARED(w);
    GBP_SHARED(K);
    GBP_SHARED(t2);
    GBP_SHARED(J);
    const int lane = lane_id();
    const int C = K->C;
#pragma unroll 1
    for (int i = 0; i < kk; ++i) {
        if (lane <= i) {
            R s = prior_op(K, kk, t2, i, lane);
#pragma unroll 1
            for (int c = 0; c < C; ++c) s += (R)J[c * KS + i] * w->ivar[c] * (R)J[c * KS + lane];
            w->A[pk(i, lane)] = s;
        }
    }
    __syncwarp();
}

// in-place packed Cholesky, lane = row.  Returns false if the matrix is not positive definite.
template <typename R> __device__ __noinline__ bool ch_cholesky(R* A, int kk)
{
    GBP_SHARED(A);
    const int lane = lane_id();
    bool ok = true;
#pragma unroll 1
    for (int j = 0; j < kk; ++j) {
        R s = R(0);
        if (lane >= j && lane < kk) {
            s = A[pk(lane, j)];
#pragma unroll 1
            for (int p = 0; p < j; ++p) s -= A[pk(lane, p)] * A[pk(j, p)];
        }
        const R d = __shfl_sync(FULL, s, j);
        if (!(d > R(0))) ok = false;
        const R dj = rt<R>::sqrt(d);
        if (lane == j) A[pk(j, j)] = dj;
        else if (lane > j && lane < kk) A[pk(lane, j)] = s / dj;
        __syncwarp();
    }
    return ok;
}
// lane i holds b_i; returns lane i of L^-1 b
template <typename R> __device__ __noinline__ R ch_solve_L(const R* A, int kk, R x)
{
    GBP_SHARED(A);
    const int lane = lane_id();
#pragma unroll 1
    for (int j = 0; j < kk; ++j) {
        const R yj = __shfl_sync(FULL, x, j) / A[pk(j, j)];
        if (lane == j) x = yj;
        else if (lane > j && lane < kk) x -= A[pk(lane, j)] * yj;
    }
    return x;
}
template <typename R> __device__ __noinline__ R ch_solve_LT(const R* A, int kk, R x)
{
    GBP_SHARED(A);
    const int lane = lane_id();
#pragma unroll 1
    for (int j = kk - 1; j >= 0; --j) {
        const R xj = __shfl_sync(FULL, x, j) / A[pk(j, j)];
        if (lane == j) x = xj;
        else if (lane < j) x -= A[pk(j, lane)] * xj;
    }
    return x;
}
// v' A v = |L' v|^2 with v_i held by lane i
template <typename R> __device__ __noinline__ R ch_quad(const R* A, R* vec, int kk, R v)
{
    GBP_SHARED(A);
    GBP_SHARED(vec);
    const int lane = lane_id();
    if (lane < kk) vec[lane] = v;
    __syncwarp();
    R s = R(0);
    if (lane < kk) {
#pragma unroll 1
        for (int i = lane; i < kk; ++i) s += A[pk(i, lane)] * vec[i];
    }
    __syncwarp();
    return warp_sum(s * s);
}

// StatArray.propose(imposePrior=True) for a 1-D log-normal random walk (StatArray.py:578-638), in ln space
template <typename R> struct prop_t {
    R x;
    uint32_t block;  // advanced Philox block counter
};
template <typename R> __device__ __noinline__ prop_t<R> ch_propose_ln_error(Rng g, R ln_cur, R sd, R lnmin, R lnmax)
{
    R x = ln_cur + sd * rng_normal<R>(g);
    int tries = 0;
#pragma unroll 1
    while (x < lnmin || x > lnmax) {
        x = ln_cur + sd * rng_normal<R>(g);
        tries++;
        if (tries == 10) {
            x = ln_cur;
            break;
        }
    }
    return prop_t<R>{x, g.block};
}

// Dual-moment datapoint: the 2-vector is proposed jointly (MvLogNormal with diagonal variance) and re-drawn, at
// most 10 times, while ANY component leaves its prior; then the whole vector falls back (StatArray.py:619-636).
// One Box-Muller pair per draw.
template <typename R> struct prop2_t {
    R x0, x1;
    uint32_t block;
};
template <typename R>
__device__ __noinline__ prop2_t<R> ch_propose_ln_error2(Rng g, R c0, R c1, const R* sd, const R* lnmin, const R* lnmax)
{
    GBP_SHARED(sd);
    GBP_SHARED(lnmin);
    GBP_SHARED(lnmax);
    pair_t<R> z = normal2_at<R>(g.block++, g.iter, g.snd_lo, g.snd_hi, g.seed_lo, g.seed_hi);
    R x0 = c0 + sd[0] * z.a, x1 = c1 + sd[1] * z.b;
    int tries = 0;
#pragma unroll 1
    while (x0 < lnmin[0] || x0 > lnmax[0] || x1 < lnmin[1] || x1 > lnmax[1]) {
        z = normal2_at<R>(g.block++, g.iter, g.snd_lo, g.snd_hi, g.seed_lo, g.seed_hi);
        x0 = c0 + sd[0] * z.a;
        x1 = c1 + sd[1] * z.b;
        tries++;
        if (tries == 10) {
            x0 = c0;
            x1 = c1;
            break;
        }
    }
    return prop2_t<R>{x0, x1, g.block};
}

// Point.perturb :614-622 = StatArray.propose :578-638 with imposePrior: Normal(z, var) random walk re-drawn while
// outside the Uniform prior [lo, hi] (closed support), at most 10 times, then the current height is kept
template <typename R> __device__ __noinline__ prop_t<R> ch_propose_height(Rng g, R cur, R sd, R lo, R hi)
{
    R x = cur + sd * rng_normal<R>(g);
    int tries = 0;
#pragma unroll 1
    while (!(x >= lo && x <= hi)) {
        x = cur + sd * rng_normal<R>(g);
        tries++;
        if (tries == 10) {
            x = cur;
            break;
        }
    }
    return prop_t<R>{x, g.block};
}

// add `count` visits of the current height to its histogram (Point.update_posteriors :1022-1025; bins relative to
// the height the prior is centred on, Point.set_z_posterior :1013-1020)
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_flush_height(WarpState<R, T, NC, KIND>* w, const Consts<R>* K, R dz, int count)
{
    GBP_SHARED(w);
    GBP_SHARED(K);
    if (count <= 0) return;
    int32_t* hh = (int32_t*)w->outp[OP_HEIGHT];
    if (hh && lane_id() == 0) hh[uniform_bin<R>(dz, -K->z_max, K->z_dx, K->n_err)] += count;
    __syncwarp();
}

// fire-and-forget integer add to GLOBAL memory (SASS REDG): the output rows come through void* slots, so a plain
// atomicAdd compiles to the generic-address ATOM
__device__ __forceinline__ void red_add_global(int32_t* p, int v)
{
    asm volatile("red.global.add.s32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}

// add `count` visits of the current model / errors to every histogram
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_flush(WarpState<R, T, NC, KIND>* w, const Consts<R>* K, int k, int mcur, int vcur,
                                      Errs<R, ns_of(KIND)> ln_err, R sig_lo, int count)
{
    GBP_SHARED(w);
    GBP_SHARED(K);
    if (count <= 0) return;
    const int lane = lane_id();
    const int nd = K->n_depth, nsb = K->n_sig, neb = K->n_err;
    const MeshBuf<R>& m = w->mesh[mcur];
    const ValBuf<R>& v = w->val[vcur];
    int32_t* hitmap = (int32_t*)w->outp[OP_HITMAP];
    int32_t* edges_hist = (int32_t*)w->outp[OP_EDGES];
    if (lane == 0) {
        int32_t* nc = (int32_t*)w->outp[OP_NCELLS];
        int32_t* rh = (int32_t*)w->outp[OP_REL];
        int32_t* ah = (int32_t*)w->outp[OP_ADD];
        if (nc) nc[k] += count;
#pragma unroll
        for (int s = 0; s < ns_of(KIND); ++s) {
            if (s < K->n_sys) {
                if (rh && K->solve_rel) rh[s * neb + uniform_bin<R>(ln_err.rel[s], K->rel_lnmin[s], K->rel_dx[s], neb)] += count;
                if (ah && K->solve_add) ah[s * neb + uniform_bin<R>(ln_err.add[s], K->add_lnmin[s], K->add_dx[s], neb)] += count;
            }
        }
    }
    // per-layer conductivity bin; interface histogram (RectilinearMesh1D.update_posteriors :1594-1610):
    // ratio sigma_i / sigma_{i-1} <= 0.5 or >= 1.5
    if (lane < k) w->sbin[lane] = uniform_bin<R>(v.ls[lane], sig_lo, K->sig_dx, nsb);
    if (lane >= 1 && lane < k && edges_hist) {
        const R dl = v.ls[lane] - v.ls[lane - 1];
        const R d = m.edges[lane];
        if ((dl <= K->ln_half || dl >= K->ln_3half) && d >= R(0) && d < K->depth_max)
            red_add_global(&edges_hist[uniform_bin<R>(d, R(0), K->depth_step, nd)], count);
    }
    __syncwarp();
    // hitmap (Model.update_parameter_posterior :819-847; staircase interp RectilinearMesh1D.py:1148-1158)
    if (hitmap) {
#pragma unroll 1
        for (int j = lane; j < nd; j += 32) {
            const R y = ((R)j + R(0.5)) * K->depth_step;
            int b = w->sbin[k - 1];
#pragma unroll 1
            for (int i = 1; i < k; ++i) {
                const R e = m.edges[i];
                if (y < e) {
                    b = w->sbin[i - 1];
                    break;
                }
                const R e2 = e * R(1.000001);
                if (y < e2) {
                    const R t = (y - e) / (e2 - e);
                    const R s = v.sig[i - 1] + t * (v.sig[i] - v.sig[i - 1]);
                    b = uniform_bin<R>(rt<R>::log(s), sig_lo, K->sig_dx, nsb);
                    break;
                }
            }
            red_add_global(&hitmap[(size_t)b * nd + j], count);  // REDG.ADD, coalesced along depth
        }
    }
    __syncwarp();
}

template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_zero_posteriors(WarpState<R, T, NC, KIND>* w, const Consts<R>* K)
{
    GBP_SHARED(w);
    GBP_SHARED(K);
    const int lane = lane_id();
    const int nd = K->n_depth;
    int32_t* p;
    if ((p = (int32_t*)w->outp[OP_HITMAP])) {
        const size_t n = (size_t)K->n_sig * nd;
#pragma unroll 1
        for (size_t i = lane; i < n; i += 32) p[i] = 0;
    }
    if ((p = (int32_t*)w->outp[OP_EDGES])) {
#pragma unroll 1
        for (int i = lane; i < nd; i += 32) p[i] = 0;
    }
    if ((p = (int32_t*)w->outp[OP_NCELLS])) {
#pragma unroll 1
        for (int i = lane; i <= K->kmax; i += 32) p[i] = 0;
    }
    if ((p = (int32_t*)w->outp[OP_REL])) {
#pragma unroll 1
        for (int i = lane; i < K->n_sys * K->n_err; i += 32) p[i] = 0;
    }
    if ((p = (int32_t*)w->outp[OP_ADD])) {
#pragma unroll 1
        for (int i = lane; i < K->n_sys * K->n_err; i += 32) p[i] = 0;
    }
    if ((p = (int32_t*)w->outp[OP_HEIGHT])) {
#pragma unroll 1
        for (int i = lane; i < K->n_err; i += 32) p[i] = 0;
    }
    __syncwarp();
}

template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_write_model(WarpState<R, T, NC, KIND>* w, int ml, int k, int mcur, int vcur, double* sig_out, double* edges_out)
{
    GBP_SHARED(w);
    const int lane = lane_id();
    const MeshBuf<R>& m = w->mesh[mcur];
    const ValBuf<R>& v = w->val[vcur];
    if (sig_out && lane < ml) sig_out[lane] = lane < k ? (double)v.sig[lane] : NAN;
    if (edges_out) {
#pragma unroll 1
        for (int i = lane; i <= ml; i += 32) edges_out[i] = i <= k ? (double)m.edges[i] : NAN;
    }
}

// 16-byte vectorised copy of nbytes (multiple of 16) by one warp
__device__ __noinline__ void ch_copy16(void* dst, const void* src, int nbytes)
{
    const int4* s4 = (const int4*)src;
    int4* d4 = (int4*)dst;
#pragma unroll 1
    for (int i = lane_id(); i < nbytes / 16; i += 32) d4[i] = s4[i];
    __syncwarp();
}

template <typename R> struct init_out {
    R ln_ref, misfit, likelihood, prior;
    double sigma_ref;
};

// Inference1D.initialize :353-464 / initialize_model :485-535: best half-space, first forward + Jacobian,
// initial misfit / likelihood / prior, cold counters.  `first == false` is reset() (:984-999).
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ init_out<R> ch_initialize(WarpState<R, T, NC, KIND>* w, const Consts<R>* K,
                                                  const typename SysOf<T, KIND>::shared* S, const T* tab, T alt, R nahl,
                                                  bool first, double* best_sig_out, double* best_edg_out)
{
    GBP_SHARED(w);
    GBP_SHARED(K);
    const int lane = lane_id();
    const int C = K->C;
    Errs<R, ns_of(KIND)> e0;
#pragma unroll
    for (int s = 0; s < ns_of(KIND); ++s) {
        e0.rel[s] = K->rel0[s];
        e0.add[s] = K->add0[s];
    }
    ch_set_ivar(w, K, C, e0);
    // EmDataPoint.find_best_halfspace :148-186: argmin misfit over logspace(-4, 4, 100)
    MeshBuf<R>& m = w->mesh[0];
    ValBuf<R>& v = w->val[0];
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
        ch_forward(w, S, tab, alt, 1, v.sig, m.edges, w->pred[0], (T*)nullptr);
        const pair_t<R> ml_ = ch_misfit_like(w, C, nahl, w->pred[0]);
        if (ml_.a < best) {
            best = ml_.a;
            best_c = c;
        }
    }
    init_out<R> io;
    io.sigma_ref = best_c;
    io.ln_ref = (R)dlog_(best_c);
    if (lane == 0) {
        v.sig[0] = (R)best_c;
        v.ls[0] = io.ln_ref;
    }
    __syncwarp();
    ch_forward(w, S, tab, alt, 1, v.sig, m.edges, w->pred[0], w->J);
    ch_round_pad(w, S, tab);   // 101 forwards: one more round, so that a team stays in phase (two rounds per step)
    if (!first) {  // reset(): posteriors and traces are re-created
        ch_zero_posteriors(w, K);
        const long long N2 = 2 * (long long)K->n_chains;
        double* mt = (double*)w->outp[OP_MISFIT];
        uint8_t* at = (uint8_t*)w->outp[OP_ACCEPT];
        if (mt) {
#pragma unroll 1
            for (long long i = lane; i < N2; i += 32) mt[i] = 0.0;
        }
        if (at) {
#pragma unroll 1
            for (long long i = lane; i < N2; i += 32) at[i] = 0;
        }
        __syncwarp();
    }
    const pair_t<R> ml0 = ch_misfit_like(w, C, nahl, w->pred[0]);
    io.misfit = ml0.a;
    io.likelihood = ml0.b;
    io.prior = ch_model_prob(K, 1, v.ls, m.lnh, io.ln_ref) + K->err_lp;
    if (lane == 0) {
        double* mt = (double*)w->outp[OP_MISFIT];
        if (mt) mt[0] = (double)io.misfit;
        w->ctr[CT_BURN_ITER] = 0;
        w->ctr[CT_N_ZERO] = 0;
        w->ctr[CT_ACC_WIN] = 0;
        w->ctr[CT_TO_PLOT] = K->upe;
        w->ctr[CT_BEST_K] = 1;
        w->ctr[CT_BEST_ITER] = 0;
        w->bestv[BV_POSTERIOR] = io.likelihood + io.prior;
        w->bestv[BV_REL] = K->rel0[0];
        w->bestv[BV_ADD] = K->add0[0];
        w->bestv[BV_REL2] = K->rel0[1];
        w->bestv[BV_ADD2] = K->add0[1];
    }
    ch_write_model(w, K->kmax, 1, 0, 0, best_sig_out, best_edg_out);
    __syncwarp();
    return io;
}

template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void ch_save_best(WarpState<R, T, NC, KIND>* w, int ml, int k, int mcur, int vcur, int iteration,
                                          R posterior, Errs<R, ns_of(KIND)> e, double* best_sig_out, double* best_edg_out)
{
    GBP_SHARED(w);
    ch_write_model(w, ml, k, mcur, vcur, best_sig_out, best_edg_out);
    if (lane_id() == 0) {
        w->ctr[CT_BEST_K] = k;
        w->ctr[CT_BEST_ITER] = iteration;
        w->bestv[BV_POSTERIOR] = posterior;
        w->bestv[BV_REL] = e.rel[0];
        w->bestv[BV_ADD] = e.add[0];
        w->bestv[BV_REL2] = e.rel[ns_of(KIND) - 1];
        w->bestv[BV_ADD2] = e.add[ns_of(KIND) - 1];
    }
    __syncwarp();
}

// ================================================================ one accept_reject step
// Hot state of a chain (registers of the warp that runs it).
template <typename R, int NS> struct Hot {
    Rng rng;
    int k, mcur, vcur, pcur;
    bool j_valid;            // the shared-memory Jacobian is the current model's
    Errs<R, NS> ln_err, err; // ln(relative / additive error) and the errors themselves, per system
    R ln_ref, sig_lo;
    R misfit, prior, likelihood;
    int dwell;
};

// Inference1D.accept_reject :537-631.  SPEC = false: the chain's own step (state updated on acceptance, outgoing
// model flushed to the posteriors).  SPEC = true: speculative evaluation of a FUTURE iteration by another warp on
// a private copy of the chain state: same arithmetic, same sub-stream of random numbers, but nothing is committed -
// the caller only learns whether the step would be rejected.
template <bool SPEC, typename R, typename T, int NC, int KIND>
__device__ __forceinline__ void ar_step(WarpState<R, T, NC, KIND>* w, const Consts<R>* K,
                                        const typename SysOf<T, KIND>::shared* S, const T* tab, T& alt, const T alt_ref,
                                        const R nahl, T* const jg, Hot<R, ns_of(KIND)>& h, bool& accepted, bool& chol_failed)
{
    constexpr int NS = ns_of(KIND);
    typedef Errs<R, NS> errs_t;
    constexpr int JBYTES = NC * KS * (int)sizeof(T);
    GBP_SHARED(w);
    GBP_SHARED(K);
    const int lane = lane_id();
    const int C = K->C;
    const int ml = K->kmax;
    Rng& rng = h.rng;
    int &k = h.k, &mcur = h.mcur, &vcur = h.vcur, &pcur = h.pcur, &dwell = h.dwell;
    bool& j_valid = h.j_valid;
    errs_t &ln_err = h.ln_err, &err = h.err;
    R &misfit = h.misfit, &prior = h.prior, &likelihood = h.likelihood;
    const R ln_ref = h.ln_ref, sig_lo = h.sig_lo;
    {
        // ---- RectilinearMesh1D.perturb :993-1120 (warp-uniform).  Proposed mesh -> mesh[mcur^1],
        //      remapped values -> val[vcur^1].sig and ls_r; for ACT_NONE only ls_r is filled.
        const MeshBuf<R>& m0 = w->mesh[mcur];
        const ValBuf<R>& v0 = w->val[vcur];
        MeshBuf<R>& m1 = w->mesh[mcur ^ 1];
        ValBuf<R>& v1 = w->val[vcur ^ 1];
        int action, kn;
#pragma unroll 1
        for (;;) {
            int event;
#pragma unroll 1
            for (;;) {  // Categorical.rng: searchsorted(cumsum(p), U), re-drawn while illegal (:1041-1049)
                const R u = rng_uniform<R>(rng);
                event = (u <= K->cum0) ? 0 : (u <= K->cum1) ? 1 : (u <= K->cum2) ? 2 : 3;
                if (k == 1 && (event == 1 || event == 2)) continue;
                if (k == ml && event == 0) continue;
                break;
            }
            action = event;
            kn = k;
            if (event == ACT_NONE) {
                if (lane < k) w->ls_r[lane] = v0.ls[lane];
                break;
            }
            if (event == ACT_BIRTH) {  // :1061-1081
                bool ok = false;
                int pos = 0;
                R e = R(0);
#pragma unroll 1
                for (int tries = 1; tries <= 10; ++tries) {
                    e = rt<R>::exp(K->ln_min_edge + K->ln_edge_span * rng_uniform<R>(rng));
                    pos = __popc(__ballot_sync(FULL, lane <= k && m0.edges[lane] < e));  // searchsorted (left)
                    R d = INFINITY;  // widths after insertion: cell pos-1 is split in two
                    if (lane < k)
                        d = (lane == pos - 1) ? fmin(e - m0.edges[lane], m0.edges[lane + 1] - e)
                                              : m0.edges[lane + 1] - m0.edges[lane];
                    const R h = warp_min(d);
                    if (tries == 10) break;  // the 10th try always restarts (:1078-1080)
                    if (h > K->min_width) {
                        ok = true;
                        break;
                    }
                }
                if (!ok) continue;
                if (lane <= k + 1) m1.edges[lane] = (lane < pos) ? m0.edges[lane] : (lane == pos ? e : m0.edges[lane - 1]);
                if (lane <= k) {  // values.insert(pos, values[pos-1]) (:835)
                    const int src = (lane < pos) ? lane : lane - 1;
                    v1.sig[lane] = v0.sig[src];
                    w->ls_r[lane] = v0.ls[src];
                }
                kn = k + 1;
                break;
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
                    w->ls_r[lane] = l;
                }
                kn = k - 1;
                break;
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
                    dz = sgn * K->min_width * rng_uniform<R>(rng);
                    R d = INFINITY;
                    if (lane < k)
                        d = (m0.edges[lane + 1] + (lane + 1 == i ? dz : R(0))) - (m0.edges[lane] + (lane == i ? dz : R(0)));
                    const R h = warp_min(d);
                    const R z1 = m0.edges[1] + (i == 1 ? dz : R(0));
                    const R zl = m0.edges[k - 1] + (i == k - 1 ? dz : R(0));
                    if (tries == 10) break;
                    if (h > K->min_width && z1 > K->min_edge && zl < K->max_edge) {
                        ok = true;
                        break;
                    }
                }
                if (!ok) continue;
                if (lane <= k) m1.edges[lane] = m0.edges[lane] + (lane == i ? dz : R(0));
                if (lane < k) {
                    v1.sig[lane] = v0.sig[lane];
                    w->ls_r[lane] = v0.ls[lane];
                }
                break;
            }
        }
        __syncwarp();
        if (lane == 0) w->ctr[CT_ACT0 + action]++;

        // ---- Model.stochastic_newton_perturbation :368-419
        const bool changed = action != ACT_NONE;
        const int mp = changed ? (mcur ^ 1) : mcur;  // proposed mesh buffer
        const int vp = vcur ^ 1;                     // proposed values buffer
        MeshBuf<R>& mesh_p = w->mesh[mp];
        ValBuf<R>& val_p = w->val[vp];
        const T* ph = w->pred[pcur];
        T* pred_t = w->pred[pcur ^ 1];
        T* const Jh = w->J;
        T* const J_t = w->J;
        if (changed) {  // observation.fm_dlogc(remapped_model): J and predicted data of the test datapoint
            ch_mesh_setup(K, kn, &mesh_p);
            ch_forward(w, S, tab, alt, kn, val_p.sig, mesh_p.edges, pred_t, J_t);
            j_valid = false;
            // REFERENCE QUIRK (kept, it shapes the proposal): FdemDataPoint.fm_dlogc stores the remapped model's
            // predicted data (FdemDataPoint.py:535-545); TdemDataPoint.fm_dlogc stores only the Jacobian, its
            // predicted-data update is commented out (TdemDataPoint.py:1031-1055), so a time-domain death / move
            // forms the Newton gradient with the CURRENT model's predicted data.
            if constexpr (!is_td(KIND)) ph = pred_t;
        } else {
            if (!SPEC) ch_round_pad(w, S, tab);   // team mode: the round of the forward this step does not need
            if (!j_valid) {  // bring the current model's (possibly stale, as in the reference) Jacobian back
                ch_copy16(w->J, jg, JBYTES);
                j_valid = true;
            }
        }
        ch_set_ivar(w, K, C, err);
        const R ln_r = (lane < kn) ? w->ls_r[lane] : R(0);
        const R g = ch_gradient(w, K, kn, mesh_p.t2, w->ls_r, Jh, ph, ln_ref);
        ch_assemble(w, K, kn, mesh_p.t2, Jh);
        if (!ch_cholesky(w->A, kn)) {
            chol_failed = true;
            if (!SPEC) ch_round_pad(w, S, tab);
        } else {
            const R stepv = ch_solve_LT(w->A, kn, ch_solve_L(w->A, kn, g));  // H * dfk
            const R mean = ln_r - K->alpha * stepv;  // ln sigma + alpha * pk, pk = -H dfk
            // sigma' ~ exp(N(mean, H)),  H = (L L')^-1  ->  mean + L^-T z
            pair_t<R> zz = {R(0), R(0)};
            const int npair = (kn + 1) / 2;
            if (lane < npair) zz = normal2_at<R>(rng.block + (uint32_t)lane, rng.iter, rng.snd_lo, rng.snd_hi, rng.seed_lo, rng.seed_hi);
            rng.block += (uint32_t)npair;
            const R za = __shfl_sync(FULL, zz.a, lane >> 1), zb = __shfl_sync(FULL, zz.b, lane >> 1);
            const R zi = (lane < kn) ? ((lane & 1) ? zb : za) : R(0);
            const R ln_t = mean + ch_solve_LT(w->A, kn, zi);
            if (lane < kn) {
                val_p.ls[lane] = ln_t;
                val_p.sig[lane] = rt<R>::exp(ln_t);
            }
            __syncwarp();

            // ---- test_datapoint.perturb() (DataPoint.py:531-573): height first (Point.perturb :614-622), then errors
            T alt_t = alt;
            if constexpr (KIND == KIND_FDEM_Z) {
                const prop_t<R> pz = ch_propose_height<R>(rng, (R)alt, K->z_sd, (R)alt_ref - K->z_max, (R)alt_ref + K->z_max);
                alt_t = (T)pz.x;
                rng.block = pz.block;
            }
            errs_t ln_t_err = ln_err, err_t = err;
            if (NS == 1 || K->n_sys == 1) {
                if (K->solve_rel) {
                    const prop_t<R> pr = ch_propose_ln_error<R>(rng, ln_err.rel[0], K->rel_sd[0], K->rel_lnmin[0], K->rel_lnmax[0]);
                    ln_t_err.rel[0] = pr.x;
                    rng.block = pr.block;
                }
                if (K->solve_add) {
                    const prop_t<R> pr = ch_propose_ln_error<R>(rng, ln_err.add[0], K->add_sd[0], K->add_lnmin[0], K->add_lnmax[0]);
                    ln_t_err.add[0] = pr.x;
                    rng.block = pr.block;
                }
            } else {
                if (K->solve_rel) {
                    const prop2_t<R> pr = ch_propose_ln_error2<R>(rng, ln_err.rel[0], ln_err.rel[NS - 1], K->rel_sd, K->rel_lnmin, K->rel_lnmax);
                    ln_t_err.rel[0] = pr.x0;
                    ln_t_err.rel[NS - 1] = pr.x1;
                    rng.block = pr.block;
                }
                if (K->solve_add) {
                    if constexpr (KIND == KIND_TEMPEST) {
                        // additive_error_multiplier.perturb() with the defaults (Tempest_datapoint.perturb :339-341): no prior
                        // imposed, and the proposal's mean is never moved - every step draws around the INITIAL multiplier
                        const pair_t<R> z = normal2_at<R>(rng.block++, rng.iter, rng.snd_lo, rng.snd_hi, rng.seed_lo, rng.seed_hi);
                        ln_t_err.add[0] = K->add_ln0[0] + K->add_sd[0] * z.a;
                        ln_t_err.add[NS - 1] = K->add_ln0[1] + K->add_sd[1] * z.b;
                    } else {
                        const prop2_t<R> pr = ch_propose_ln_error2<R>(rng, ln_err.add[0], ln_err.add[NS - 1], K->add_sd, K->add_lnmin, K->add_lnmax);
                        ln_t_err.add[0] = pr.x0;
                        ln_t_err.add[NS - 1] = pr.x1;
                        rng.block = pr.block;
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                err_t.rel[s] = (ln_t_err.rel[s] == ln_err.rel[s]) ? err.rel[s] : rt<R>::exp(ln_t_err.rel[s]);
                err_t.add[s] = (ln_t_err.add[s] == ln_err.add[s]) ? err.add[s] : rt<R>::exp(ln_t_err.add[s]);
            }
            // a time-domain datapoint perturbs its loops AFTER the errors (TdemDataPoint.perturb :681-683): the
            // transmitter height, and with it the Hankel abscissae / geometry weights of the candidate's forward
            if constexpr (KIND == KIND_TDEM_Z) {
                const prop_t<R> pz = ch_propose_height<R>(rng, (R)alt, K->z_sd, (R)alt_ref - K->z_max, (R)alt_ref + K->z_max);
                alt_t = (T)pz.x;
                rng.block = pz.block;
                if (alt_t != alt) {
                    const int g = w->fx.cur ^ 1;
                    if constexpr (sizeof(T) == 4) td_geometry_f32(*S, alt_t, w->fx.lam[g], w->fx.wgt[g]);
                    else td_geometry<T>(*S, (double)alt_t, w->fx.lam[g], w->fx.wgt[g]);
                    if (lane == 0) w->fx.alt[g] = alt_t;
                    __syncwarp();
                }
            }

            const bool jump = (action == ACT_BIRTH || action == ACT_DEATH);
            // forward at the candidate; for birth/death the Jacobian at the candidate is needed as well
            // (Model.proposal_probabilities :619) - fused into the same pass.
            ch_forward(w, S, tab, alt_t, kn, val_p.sig, mesh_p.edges, pred_t, jump ? J_t : (T*)nullptr);
            ch_set_ivar(w, K, C, err_t);
            const pair_t<R> tml = ch_misfit_like(w, C, nahl, pred_t);
            // error priors: the proposals above are forced inside their bounds (or fall back to the
            // current values), so DataPoint.probability is the constant err_lp
            R t_prior = K->err_lp;
            t_prior += ch_model_prob(K, kn, val_p.ls, mesh_p.lnh, ln_ref);
            if (t_prior != (R)-INFINITY) {  // early reject on -inf prior (:581, :589): no accept draw
                R proposal = R(1), proposal1 = R(1);
                if (jump) {
                    const R g2 = ch_gradient(w, K, kn, mesh_p.t2, val_p.ls, J_t, pred_t, ln_ref);
                    const R s2 = ch_solve_LT(w->A, kn, ch_solve_L(w->A, kn, g2));  // H dfk'
                    const R lv = ln_t + K->alpha * s2;  // Model.py:626 (sign as in the reference)
                    // mean = expReal(log_values) is inf above 11356 and underflows to 0 below the long-double
                    // denormal limit (base/utilities.py:827-856): Model.py:630-633 then returns -inf, -inf
                    const int bad = __any_sync(FULL, lane < kn && (lv > R(11356) || lv < R(-11399)));
                    const R q_r = ch_quad(w->A, w->vec, kn, (lane < kn) ? (ln_r - lv) : R(0));
                    const R q_f = ch_quad(w->A, w->vec, kn, (lane < kn) ? (ln_t - ln_r) : R(0));
                    const R logdetL = warp_sum((lane < kn) ? rt<R>::log(w->A[pk(lane, lane)]) : R(0));
                    if (bad) {
                        proposal = (R)-INFINITY;
                        proposal1 = (R)-INFINITY;
                    } else {
                        proposal = -(R)kn * K->half_log2pi + logdetL - R(0.5) * q_r;
                        proposal1 = -(R)kn * K->half_log2pi + logdetL - R(0.5) * q_f;
                    }
                }
                const R log_alpha = (t_prior - prior) + (tml.b - likelihood) + (proposal - proposal1);
                const R u = rng_uniform<R>(rng);
                accepted = rt<R>::exp(log_alpha) > u;
                if (accepted && !SPEC) {  // a speculative evaluation only reports the outcome
                    ch_flush(w, K, k, mcur, vcur, ln_err, sig_lo, dwell);  // the outgoing model's visits
                    if constexpr (has_z(KIND)) {
                        ch_flush_height(w, K, (R)(alt - alt_ref), dwell);
                        if constexpr (KIND == KIND_TDEM_Z) {
                            if (alt_t != alt && lane == 0) w->fx.cur ^= 1;   // the proposed height's geometry becomes current
                            __syncwarp();
                        }
                        alt = alt_t;
                    }
                    dwell = 0;
                    misfit = tml.a;
                    prior = t_prior;
                    likelihood = tml.b;
                    k = kn;
                    err = err_t;
                    ln_err = ln_t_err;
                    mcur = mp;
                    vcur = vp;
                    pcur ^= 1;
                    if (changed) {  // action none keeps the (stale) Jacobian, as the reference does
                        ch_copy16(jg, w->J, JBYTES);
                        j_valid = true;
                    }
                    if (lane == 0) w->ctr[CT_N_ACCEPT]++;
                } else if (accepted) {  // SPEC: leave the proposed state in this warp's buffers and describe it
                    if (lane == 0) {
                        w->sout.kn = kn;
                        w->sout.mp = mp;
                        w->sout.vp = vp;
                        w->sout.changed = changed ? 1 : 0;
                        w->sout.misfit = tml.a;
                        w->sout.prior = t_prior;
                        w->sout.likelihood = tml.b;
                        w->sout.err = err_t;
                        w->sout.ln_err = ln_t_err;
                        if constexpr (has_z(KIND)) w->sout.alt = (R)alt_t;
                    }
                    __syncwarp();
                }
            }
        }
    }
}

// ================================================================ speculative evaluation of future iterations
// Chains are sequential and their lengths differ by 6x (the reference's reset rule), so every batch ends with a
// tail in which most warps of an SM have no chain left; small batches never fill the machine at all.  An
// accept_reject step that is REJECTED leaves the chain state untouched, and (sub-streams, gbp_math.cuh) the random
// numbers of iteration t do not depend on iterations < t.  So idle warps of the CTA evaluate iterations t0, t0+1, ...
// of a running chain in parallel on private copies of its state, each assuming that all earlier ones are rejected.
// The owner commits the leading run of rejections (bookkeeping only) and, if the first step that is not a plain
// rejection is an acceptance, ADOPTS the proposed state from the warp that evaluated it.  Results are bitwise
// identical to the sequential chain.  Chains that are stuck (the ones the reset rule makes 3-6x longer) advance W
// iterations per iteration time; a chain with acceptance rate a advances about 1/a.
constexpr int SPEC_RES = 256;  // iterations one round can cover
// diagnostics of the speculation hand-off (gbp_debug_counters): [0] helper wake-ups, [1] cycles GO -> helper awake,
// [2] cycles copying the chain state, [3] speculative steps, [4] cycles in speculative steps, [5] cycles a stopped
// helper waits for release, [6] owner cycles waiting for `done`, [7] owner cycles copying the adopted proposal,
// [8] iterations committed from speculation, [9] rounds, [10] owner cycles spent in rounds.  These are the only
// outputs that depend on scheduling; they are kept out of the per-chain result arrays, which are bit-identical
// with speculation on or off.
__device__ unsigned long long g_diag[16];
enum { MB_BUSY = 0, MB_IDLE = 1, MB_CLAIMED = 2, MB_GO = 16 };  // mailbox states of a warp (MB_GO + owner warp)
enum { STOP_NONE = 0, STOP_ACCEPT = 1, STOP_OTHER = 2 };

template <typename R, typename T, int NS> struct SpecRound {
    Hot<R, NS> hot;            // chain state before iteration t0
    SpecOut<R, NS> win;        // outcome of the accepted step the owner adopts
    T alt, alt_ref;            // sensor height of the chain state; centre of its prior (KIND_FDEM_Z)
    R nahl;
    T* jg;
    const void* owner_ws;
    volatile int t0, t_end, W;
    volatile int first_stop;   // smallest iteration whose step was not a plain rejection (INT_MAX: none yet)
    volatile int done;         // helpers finished
    volatile unsigned seq, released;  // round number; helpers keep their state until released == seq
    long long t_go;
    int win_byte;
    unsigned char midx[32];    // member index + 1 of warp x in this round
    volatile unsigned char res[SPEC_RES];  // per iteration t0+i: 0x80 valid | 0x40 stop | nsens << 4 | nfwd << 2 | action
};
template <typename R, typename T, int NS> struct TailCtx {
    SpecRound<R, T, NS>* rounds;   // [n_warps]
    volatile int* mailbox;         // [n_warps]
    volatile int* n_alive;         // warps that still own a chain
    volatile int* n_idle;          // warps waiting in tail_service()
    int n_warps, max_helpers, warp;
};

// 4-byte-word copy of nbytes (multiple of 4) by one warp
__device__ __forceinline__ void ch_copy4(void* dst, const void* src, int nbytes)
{
    const int* s4 = (const int*)src;
    int* d4 = (int*)dst;
#pragma unroll 1
    for (int i = lane_id(); i < nbytes / 4; i += 32) d4[i] = s4[i];
}

// Evaluate iterations t0 + member, + W, ... < t_end speculatively on w (a private copy of the chain state).
// Returns true if this warp stopped on a step that was not a plain rejection (its buffers then hold the proposal).
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ bool spec_member_run(WarpState<R, T, NC, KIND>* w, SpecRound<R, T, ns_of(KIND)>* rd, const Consts<R>* K,
                                             const typename SysOf<T, KIND>::shared* S, const T* tab, const int member)
{
    GBP_SHARED(w);
    GBP_SHARED(rd);
    GBP_SHARED(K);
    const int lane = lane_id();
    Hot<R, ns_of(KIND)> h = rd->hot;
    h.j_valid = false;
    const int t0 = rd->t0, t_end = rd->t_end, W = rd->W;
    T alt = rd->alt;   // left alone by a speculative step; the proposed height goes back in SpecOut
    const T alt_ref = rd->alt_ref;
    const R nahl = rd->nahl;
    T* const jg = rd->jg;
    bool stopped = false;
#pragma unroll 1
    for (int t = t0 + member; t < t_end; t += W) {
        const int fs = __shfl_sync(FULL, (int)rd->first_stop, 0);
        if (t > fs) break;
        if (lane == 0) {
            w->ctr[CT_N_FWD] = 0;
            w->ctr[CT_N_SENS] = 0;
            w->ctr[CT_ACT0] = 0;
            w->ctr[CT_ACT1] = 0;
            w->ctr[CT_ACT2] = 0;
        }
        __syncwarp();
        h.rng.iter = (uint32_t)t + 1u;
        h.rng.block = 0u;
        bool accepted = false, chol_failed = false;
        const long long c_s = clock64();
        ar_step<true, R, T, NC, KIND>(w, K, S, tab, alt, alt_ref, nahl, jg, h, accepted, chol_failed);
        if (lane == 0) {
            atomicAdd(&g_diag[3], 1ull);
            atomicAdd(&g_diag[4], (unsigned long long)(clock64() - c_s));
        }
        const int stop = (accepted || chol_failed) ? 1 : 0;
        __syncwarp();
        if (lane == 0) {
            const int action = w->ctr[CT_ACT0] ? 0 : (w->ctr[CT_ACT1] ? 1 : (w->ctr[CT_ACT2] ? 2 : 3));
            const int byte = 0x80 | (stop << 6) | ((w->ctr[CT_N_SENS] & 3) << 4) | ((w->ctr[CT_N_FWD] & 3) << 2) | action;
            w->ctr[CT_N_ACCEPT] = (accepted && !chol_failed) ? 1 : 0;  // kind of stop, read by the owner
            if (stop) atomicMin((int*)&rd->first_stop, t);
            rd->res[t - t0] = (unsigned char)byte;
        }
        if (stop) {
            stopped = true;
            break;
        }
    }
    __threadfence_block();
    __syncwarp();
    return stopped;
}

// A warp without a chain: serve speculative rounds of the chains of this CTA until none is left.
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ void tail_service(WarpState<R, T, NC, KIND>* w, TailCtx<R, T, ns_of(KIND)> tc, const Consts<R>* K,
                                          const typename SysOf<T, KIND>::shared* S, const T* tab)
{
    const int lane = lane_id();
    if (lane == 0) {
        tc.mailbox[tc.warp] = MB_IDLE;
        __threadfence_block();
        atomicAdd((int*)tc.n_idle, 1);
    }
#pragma unroll 1
    for (;;) {
        int cmd = 0;
        if (lane == 0) {
            // poll with back-off: ncu showed 27 idle warps polling every ~50 ns (a short nanosleep returns almost at
            // once) issuing 95 % of all instructions of a tail-dominated run and halving the speed of the working warps
            unsigned ns = 1000u;
#pragma unroll 1
            for (;;) {
                cmd = tc.mailbox[tc.warp];
                if (cmd >= MB_GO) break;
                if (cmd == MB_IDLE && *tc.n_alive <= 0) {
                    // nobody left to help; leave unless an owner claimed this warp in the meantime
                    if (atomicCAS((int*)&tc.mailbox[tc.warp], MB_IDLE, MB_BUSY) == MB_IDLE) {
                        cmd = -1;
                        break;
                    }
                    continue;
                }
                __nanosleep(ns);
                if (ns < 8000u) ns += 1000u;
            }
        }
        cmd = __shfl_sync(FULL, cmd, 0);
        if (cmd < 0) break;
        SpecRound<R, T, ns_of(KIND)>* rd = tc.rounds + (cmd - MB_GO);
        __threadfence_block();
        const unsigned seq = __shfl_sync(FULL, (unsigned)rd->seq, 0);
        const long long c_wake = clock64();
        ch_copy16(w, rd->owner_ws, (int)sizeof(WarpState<R, T, NC, KIND>));
        if (lane == 0) w->team = nullptr;   // a private copy: its forwards are this warp's alone
        __syncwarp();
        const long long c_copied = clock64();
        if (lane == 0) {
            atomicAdd(&g_diag[0], 1ull);
            atomicAdd(&g_diag[1], (unsigned long long)(c_wake - rd->t_go));
            atomicAdd(&g_diag[2], (unsigned long long)(c_copied - c_wake));
        }
        const int member = (int)rd->midx[tc.warp] - 1;
        const bool stopped = spec_member_run<R, T, NC, KIND>(w, rd, K, S, tab, member);
        if (lane == 0) {
            atomicAdd((int*)&rd->done, 1);
            // a warp that stopped may hold the proposal the owner adopts: keep the buffers until the owner is done
            if (stopped) {
                const long long c_r = clock64();
#pragma unroll 1
                while (rd->released != seq) __nanosleep(1000);
                atomicAdd(&g_diag[5], (unsigned long long)(clock64() - c_r));
            }
            __threadfence_block();
            tc.mailbox[tc.warp] = MB_IDLE;
        }
        __syncwarp();
    }
    if (lane == 0) atomicSub((int*)tc.n_idle, 1);
}

// Owner: claim idle warps and start a round covering iterations total .. total+len-1.  Returns the number of helpers.
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ int spec_round_begin(WarpState<R, T, NC, KIND>* w, TailCtx<R, T, ns_of(KIND)> tc,
                                             const Hot<R, ns_of(KIND)> h, T alt, T alt_ref, R nahl, T* jg, int total, int mult,
                                             int want, int need)
{
    const int lane = lane_id();
    SpecRound<R, T, ns_of(KIND)>* rd = tc.rounds + tc.warp;
    GBP_SHARED(rd);
    int nh = 0;
    if (lane == 0) {
#pragma unroll 1
        for (int x = 0; x < tc.n_warps && nh < want; ++x)
            if (tc.mailbox[x] == MB_IDLE && atomicCAS((int*)&tc.mailbox[x], MB_IDLE, MB_CLAIMED) == MB_IDLE) rd->midx[x] = (unsigned char)(++nh);
        if (nh < need) {  // the owner waits during a round: with too few helpers a round is slower than the chain itself
#pragma unroll 1
            for (int x = 0; x < tc.n_warps; ++x)
                if (rd->midx[x] != 0) {
                    rd->midx[x] = 0;
                    tc.mailbox[x] = MB_IDLE;
                }
            nh = 0;
        }
    }
    nh = __shfl_sync(FULL, nh, 0);
    if (nh == 0) return 0;
    int len = nh * mult;
    if (len > SPEC_RES) len = SPEC_RES;
#pragma unroll 1
    for (int i = lane; i < len; i += 32) rd->res[i] = 0;
    if (lane == 0) {
        rd->hot = h;
        rd->alt = alt;
        rd->alt_ref = alt_ref;
        rd->nahl = nahl;
        rd->jg = jg;
        rd->owner_ws = w;
        rd->t0 = total;
        rd->t_end = total + len;
        rd->W = nh;
        rd->first_stop = 0x7fffffff;
        rd->done = 0;
        rd->seq = rd->seq + 1u;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        rd->t_go = clock64();
        __threadfence_block();
#pragma unroll 1
        for (int x = 0; x < tc.n_warps; ++x)
            if (tc.mailbox[x] == MB_CLAIMED && rd->midx[x] != 0) tc.mailbox[x] = MB_GO + tc.warp;
    }
    __syncwarp();
    return nh;
}

// Owner: wait for the helpers.  Returns the number of leading plain rejections from iteration t0 on (their results
// are in rd->res); *stop_kind says what follows them: nothing evaluated (STOP_NONE), an acceptance whose proposed
// state has been copied into the owner's scratch buffers and rd->win (STOP_ACCEPT), or something the owner has to
// execute itself (STOP_OTHER).
template <typename R, typename T, int NC, int KIND>
__device__ __noinline__ int spec_round_end(WarpState<R, T, NC, KIND>* w, TailCtx<R, T, ns_of(KIND)> tc, int nh, int* stop_kind)
{
    const int lane = lane_id();
    SpecRound<R, T, ns_of(KIND)>* rd = tc.rounds + tc.warp;
    GBP_SHARED(rd);
    GBP_SHARED(w);
    if (lane == 0) {
        const long long c_w = clock64();
#pragma unroll 1
        while (rd->done < nh) __nanosleep(1000);
        atomicAdd(&g_diag[6], (unsigned long long)(clock64() - c_w));
    }
    __threadfence_block();
    __syncwarp();
    const long long c_a = clock64();
    const int fs = rd->first_stop, t0 = rd->t0, te = rd->t_end, W = rd->W;
    int n = ((fs < te) ? fs : te) - t0;
    // every committed slot must hold a valid plain rejection (defensive: stop at the first that does not)
    int bad = n;
#pragma unroll 1
    for (int i = lane; i < n; i += 32)
        if ((rd->res[i] & 0xC0) != 0x80) bad = min(bad, i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad = min(bad, __shfl_xor_sync(FULL, bad, o));
    int kind = STOP_NONE;
    if (bad < n) {
        n = bad;
        kind = STOP_OTHER;
    } else if (fs < te) {
        kind = STOP_OTHER;
        // the warp that evaluated iteration fs: member (fs - t0) % W
        const int member = (fs - t0) % W;
        int x = -1;
        if (lane < tc.n_warps && (int)rd->midx[lane] == member + 1) x = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(FULL, x, o));
        if (x >= 0) {
            const WarpState<R, T, NC, KIND>* hw = w + (x - tc.warp);
            if (hw->ctr[CT_N_ACCEPT] == 1) {  // an acceptance: adopt the proposal
                kind = STOP_ACCEPT;
                const SpecOut<R, ns_of(KIND)> so = hw->sout;
                const int pc = rd->hot.pcur ^ 1;
                if (so.changed) ch_copy4(&w->mesh[so.mp], &hw->mesh[so.mp], (int)sizeof(MeshBuf<R>));
                ch_copy4(&w->val[so.vp], &hw->val[so.vp], (int)sizeof(ValBuf<R>));
                ch_copy4(w->pred[pc], hw->pred[pc], NC * (int)sizeof(T));
                if (so.changed) ch_copy16(w->J, hw->J, NC * KS * (int)sizeof(T));
                if constexpr (KIND == KIND_TDEM_Z) {   // the geometry set of the proposed transmitter height
                    const int g = w->fx.cur ^ 1;
                    ch_copy4(w->fx.lam[g], hw->fx.lam[g], GBP_TD_MAXLAM * (int)sizeof(T));
                    ch_copy4(w->fx.wgt[g], hw->fx.wgt[g], GBP_TD_MAXLAM * (int)sizeof(T));
                    if (lane == 0) w->fx.alt[g] = hw->fx.alt[g];
                }
                if (lane == 0) {
                    rd->win = so;
                    rd->win_byte = rd->res[fs - t0];
                }
            }
        }
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        for (int x = 0; x < tc.n_warps; ++x) rd->midx[x] = 0;
        rd->released = rd->seq;
        atomicAdd(&g_diag[7], (unsigned long long)(clock64() - c_a));
    }
    __syncwarp();
    *stop_kind = kind;
    return n;
}

// ================================================================ the chain (inlined into the kernel)
template <typename R, typename T, int NC, int KIND>
__device__ __forceinline__ void run_chain(WarpState<R, T, NC, KIND>* w, const Consts<R>* K,
                                          const typename SysOf<T, KIND>::shared* S, const typename SysOf<T, KIND>::dev& Sdev,
                                          const T* tab, const ChainParams& P, const int chain,
                                          const TailCtx<R, T, ns_of(KIND)>& tc)
{
    constexpr int NS = ns_of(KIND);
    typedef Errs<R, NS> errs_t;
    GBP_SHARED(w);
    GBP_SHARED(K);
    const int lane = lane_id();
    const int C = K->C;
    const int ml = K->kmax;
    const int N = K->n_chains;
#define N2 (2 * (long long)N)

    // ---- hot state (registers)
    Hot<R, NS> h;
    Rng& rng = h.rng;
    rng.seed_lo = (uint32_t)P.seed;
    rng.seed_hi = (uint32_t)(P.seed >> 32);
    {
        const unsigned long long snd = P.first_index + (unsigned long long)chain;
        rng.snd_lo = (uint32_t)snd;
        rng.snd_hi = (uint32_t)(snd >> 32);
    }
    rng.block = 0u;
    rng.iter = 0u;
    T alt = (T)P.altitude[chain];   // sensor height: constant unless KIND_FDEM_Z samples it
    T alt_ref = alt, best_alt = alt; // centre of the height prior; height of the best model
    int &k = h.k, &mcur = h.mcur, &vcur = h.vcur, &pcur = h.pcur, &dwell = h.dwell;
    bool& j_valid = h.j_valid;
    errs_t &ln_err = h.ln_err, &err = h.err;
    R &ln_ref = h.ln_ref, &sig_lo = h.sig_lo, &misfit = h.misfit, &prior = h.prior, &likelihood = h.likelihood;
    k = 1;
    mcur = vcur = pcur = 0;
    j_valid = true;
    ln_ref = sig_lo = misfit = prior = likelihood = R(0);
    dwell = 0;
    T* const jg = (T*)P.jstore + (size_t)chain * NC * KS;
    constexpr int JBYTES = NC * KS * (int)sizeof(T);
    double sigma_ref = 0.0;
    int iteration = 0, burned_in = 0;

    // ---- per-chain setup
    int act = 0;
#pragma unroll
    for (int p = 0; p < (NC + 31) / 32; ++p) {
        const int c = lane + 32 * p;
        if (c < C) {
            const double d = P.data[(size_t)chain * C + c];
            const int a = d > 0.0;   // EmDataPoint.active: observed > 0 and not NaN
            act += a;
            w->data[c] = a ? (R)(d * P.data_scale) : R(0);
        }
    }
    const int n_active = warp_sum_i(act);
    if constexpr (is_td(KIND)) {
        td_geometry<T>(*S, P.altitude[chain], w->fx.lam[0], w->fx.wgt[0]);
        if constexpr (KIND == KIND_TEMPEST) td_geometry<T>(*S, P.altitude[chain], w->fx.lam[0], w->fx.wgt[1], true);
        if (lane == 0) {
            w->fx.cur = 0;
            w->fx.alt[0] = alt;
        }
        __syncwarp();
    }
    const R nahl = (R)n_active * K->half_log2pi;
    if (lane == 0) {
        const gbp_chain_buffers& o = P.out;
        w->outp[OP_HITMAP] = o.hitmap ? o.hitmap + (size_t)chain * K->n_sig * K->n_depth : nullptr;
        w->outp[OP_EDGES] = o.edges_hist ? o.edges_hist + (size_t)chain * K->n_depth : nullptr;
        w->outp[OP_NCELLS] = o.ncells_hist ? o.ncells_hist + (size_t)chain * (ml + 1) : nullptr;
        w->outp[OP_REL] = o.rel_hist ? o.rel_hist + (size_t)chain * K->n_sys * K->n_err : nullptr;
        w->outp[OP_ADD] = o.add_hist ? o.add_hist + (size_t)chain * K->n_sys * K->n_err : nullptr;
        w->outp[OP_MISFIT] = o.misfit_trace ? o.misfit_trace + (size_t)chain * N2 : nullptr;
        w->outp[OP_ACCEPT] = o.accept_trace ? o.accept_trace + (size_t)chain * N2 : nullptr;
        w->outp[OP_HEIGHT] = (has_z(KIND) && o.height_hist) ? o.height_hist + (size_t)chain * K->n_err : nullptr;
        for (int i = 0; i < CT_N; ++i) w->ctr[i] = 0;
    }
    __syncwarp();
#define GBP_BEST_SIG (P.out.best_sigma ? P.out.best_sigma + (size_t)chain * ml : nullptr)
#define GBP_BEST_EDG (P.out.best_edges ? P.out.best_edges + (size_t)chain * (ml + 1) : nullptr)

    // Inference1D.initialize :353-464 (also used by reset() :984-999): cold, one shared copy
    auto initialize = [&](bool first) {
        // reset() re-initialises with the CURRENT datapoint (Inference1D.py:984-994): a sampled height is kept and
        // its prior, proposal and posterior bins are re-centred on it (Point.set_priors :959-961)
        alt_ref = alt;
        best_alt = alt;
        const init_out<R> io = ch_initialize<R, T, NC, KIND>(w, K, S, tab, alt, nahl, first, GBP_BEST_SIG, GBP_BEST_EDG);
        mcur = vcur = pcur = 0;
        ch_copy16(jg, w->J, JBYTES);
        j_valid = true;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            ln_err.rel[s] = K->rel_ln0[s];
            ln_err.add[s] = K->add_ln0[s];
            err.rel[s] = K->rel0[s];
            err.add[s] = K->add0[s];
        }
        sigma_ref = io.sigma_ref;
        ln_ref = io.ln_ref;
        sig_lo = ln_ref - K->sig_halfspan;  // Model.set_posteriors :666-684
        misfit = io.misfit;
        likelihood = io.likelihood;
        prior = io.prior;
        k = 1;
        burned_in = 0;
        iteration = 0;
        dwell = 0;
    };
    auto save_best = [&]() {
        best_alt = alt;
        ch_save_best<R, T, NC, KIND>(w, ml, k, mcur, vcur, iteration, likelihood + prior, err, GBP_BEST_SIG, GBP_BEST_EDG);
    };

    initialize(true);

    bool failed = (n_active == 0);
    bool go = !failed;
    int total = 0;
    int spec_left = 0, spec_pos = 0, spec_mult = 1, rej_run = 0, spec_stop = STOP_NONE;
    int spec_rounds = 0, spec_helpers_sum = 0;   // diagnostics
    long long spec_cycles = 0;
#pragma unroll 1
    while (go) {
        // ==================================================== Inference1D.accept_reject :537-631
        bool accepted = false;
        bool chol_failed = false;
        // only a chain that has just rejected a few steps speculates: long rejection runs make long rounds (the
        // hand-off is amortised), and a chain that accepts every other step gains nothing
        // (helpers are shared by the chains of a CTA: long rejection runs use them fully, a chain with acceptance rate a
        // commits about 1/a iterations per round whatever their number - lower thresholds were measured slower)
        // Team mode: a chain that gets stuck (a long rejection run) while it is the last one of its team and finds fewer
        // than 4 idle warps in the CTA DISSOLVES the team - its mates leave the round protocol and become speculation
        // helpers, the chain goes on alone.
        if constexpr (!is_td(KIND) && sizeof(T) == 4) {
            if (w->team != nullptr && tc.max_helpers > 0 && rej_run >= P.spec_min_rejections && w->team->alive <= 1 && *tc.n_idle < 4) {
                if (lane == 0) {
                    w->team->alive = 0;
                    w->req_active = 0;
                }
                team_round<R, T, NC, KIND>(w, S, tab);   // everybody sees alive == 0 and leaves
                if (lane == 0) w->team = nullptr;
                __syncwarp();
            }
        }
        // (a chain that still has its team speculates with the warps of OTHER teams that are done; its mates keep serving
        // its own steps)
        if (spec_left == 0 && spec_stop == STOP_NONE && tc.max_helpers > 0 && rej_run >= P.spec_min_rejections && *tc.n_idle > 0 &&
            (w->team == nullptr || w->team->alive <= 1)) {
            // helpers asked for grow with the rejection run (a short run usually ends within a few steps); the owner
            // waits during a round, so a round needs at least 2 (stuck chain) to 4 helpers to pay
            int want = 4 + (rej_run >> 2);
            if (want > tc.max_helpers) want = tc.max_helpers;
            const int need = (rej_run >= 24) ? 2 : 4;
            const long long c0 = clock64();
            const int nh = spec_round_begin<R, T, NC, KIND>(w, tc, h, alt, alt_ref, nahl, jg, total, spec_mult, want, need);
            if (nh > 0) {
                spec_left = spec_round_end<R, T, NC, KIND>(w, tc, nh, &spec_stop);
                spec_rounds++;
                spec_cycles += clock64() - c0;
                spec_helpers_sum += nh;
                spec_pos = 0;
                if (spec_stop == STOP_ACCEPT && tc.rounds[tc.warp].win.changed) j_valid = false;  // w->J now holds the proposal's
                // rounds grow while nothing but rejections comes back (a stuck chain), and start small again otherwise
                spec_mult = (spec_stop == STOP_NONE) ? min(spec_mult * 2, 32) : 1;
            }
        }
        if (spec_left > 0) {
            // this iteration was evaluated speculatively and is a plain rejection: only its bookkeeping remains
            const int byte = tc.rounds[tc.warp].res[spec_pos];
            spec_pos++;
            spec_left--;
            if (lane == 0) {
                w->ctr[CT_ACT0 + (byte & 3)]++;
                w->ctr[CT_N_FWD] += (byte >> 2) & 3;
                w->ctr[CT_N_SENS] += (byte >> 4) & 3;
                w->ctr[CT_N_SPEC]++;
            }
            __syncwarp();
        } else if (spec_stop == STOP_ACCEPT) {
            // this iteration was evaluated speculatively and is an ACCEPTANCE: adopt the proposed state (already
            // copied into this warp's scratch buffers), with the side effects of the chain's own accept path
            spec_stop = STOP_NONE;
            const SpecOut<R, NS> so = tc.rounds[tc.warp].win;
            const int byte = tc.rounds[tc.warp].win_byte;
            ch_flush(w, K, k, mcur, vcur, ln_err, sig_lo, dwell);  // the outgoing model's visits
            if constexpr (has_z(KIND)) {
                ch_flush_height(w, K, (R)(alt - alt_ref), dwell);
                if constexpr (KIND == KIND_TDEM_Z) {
                    // (spec_round_end copied the proposed height's geometry set into this warp's spare set)
                    if ((T)so.alt != alt && lane == 0) w->fx.cur ^= 1;
                    __syncwarp();
                }
                alt = (T)so.alt;
            }
            dwell = 0;
            misfit = so.misfit;
            prior = so.prior;
            likelihood = so.likelihood;
            k = so.kn;
            err = so.err;
            ln_err = so.ln_err;
            mcur = so.mp;
            vcur = so.vp;
            pcur ^= 1;
            if (so.changed) {
                ch_copy16(jg, w->J, JBYTES);
                j_valid = true;
            }
            if (lane == 0) {
                w->ctr[CT_N_ACCEPT]++;
                w->ctr[CT_ACT0 + (byte & 3)]++;
                w->ctr[CT_N_FWD] += (byte >> 2) & 3;
                w->ctr[CT_N_SENS] += (byte >> 4) & 3;
                w->ctr[CT_N_SPEC]++;
            }
            __syncwarp();
            accepted = true;
        } else {
            spec_stop = STOP_NONE;
            rng.iter = (uint32_t)total + 1u;  // sub-stream of this accept_reject step
            rng.block = 0u;
            ar_step<false, R, T, NC, KIND>(w, K, S, tab, alt, alt_ref, nahl, jg, h, accepted, chol_failed);
        }
        failed = chol_failed;
        rej_run = accepted ? 0 : rej_run + 1;

        // ==================================================== Inference1D.update :705-790
        bool do_reset = false;
        iteration++;
        {
            double* mt = (double*)w->outp[OP_MISFIT];
            if (mt && lane == 0 && iteration - 1 < N2) mt[iteration - 1] = (double)misfit;
        }
        if (!burned_in && iteration > K->burn_min && misfit < (R)n_active) {
            burned_in = 1;
            if (lane == 0) w->ctr[CT_BURN_ITER] = iteration;
            save_best();
            ch_zero_posteriors(w, K);
            dwell = 0;
        }
        if (likelihood + prior > w->bestv[BV_POSTERIOR]) save_best();
        {
            uint8_t* at = (uint8_t*)w->outp[OP_ACCEPT];
            if (at && lane == 0 && iteration < N2) at[iteration] = accepted ? 1 : 0;
        }
        {
            // acceptance over acceptance_v[it-upe : it] (Inference1D.py:125-131): excludes this iteration
            int to_plot = w->ctr[CT_TO_PLOT] - 1;
            int win = w->ctr[CT_ACC_WIN];
            __syncwarp();
            if (to_plot == 0) {
                to_plot = K->upe;
                if (K->upe > 1) {
                    if (!burned_in) {
                        if (win == 0) {
                            int nz = w->ctr[CT_N_ZERO] + 1;
                            if (nz == K->reset_limit) {
                                do_reset = true;
                                nz = 0;
                            }
                            __syncwarp();
                            if (lane == 0) w->ctr[CT_N_ZERO] = nz;
                        } else if (lane == 0) w->ctr[CT_N_ZERO] = 0;
                    } else if (win == 0 && lane == 0) w->ctr[CT_LIMITERS] = 0;
                }
                win = 0;
            }
            win += accepted ? 1 : 0;
            if (lane == 0) {
                w->ctr[CT_TO_PLOT] = to_plot;
                w->ctr[CT_ACC_WIN] = win;
            }
            __syncwarp();
        }
        if (!do_reset) dwell++;

        // ==================================================== Inference1D.infer :650-677 (loop control)
        if (P.finish_ns && (total & 2047) == 0 && lane == 0) {   // debug timeline: a stamp every 2048 iterations
            unsigned long long tnow;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
            P.finish_ns[(size_t)P.B + 1 + (size_t)chain * 32 + min(31, total >> 11)] = tnow;
        }
        total++;
        if (do_reset) {
            int nr = w->ctr[CT_N_RESETS] + 1;
            __syncwarp();
            if (lane == 0) w->ctr[CT_N_RESETS] = nr;
            initialize(false);
            spec_left = 0;  // speculated iterations assumed the old state
            spec_stop = STOP_NONE;
            dwell = 1;  // update() goes on to accumulate the re-initialised model
        }
        const int burn_iter = w->ctr[CT_BURN_ITER];
        go = !failed && (iteration <= N + burn_iter);
        if (!failed && !burned_in) {
            go = iteration < N;
            if (!go) failed = true;
        }
        if (w->ctr[CT_N_RESETS] == 3 && !burned_in) {
            const int lim = w->ctr[CT_LIMITERS];
            __syncwarp();
            if (!lim) {
                if (lane == 0) {
                    w->ctr[CT_LIMITERS] = 1;
                    w->ctr[CT_N_RESETS] = 1;
                }
                initialize(false);
            spec_left = 0;  // speculated iterations assumed the old state
            spec_stop = STOP_NONE;
            } else {
                go = false;
                failed = true;
            }
        }
        if (P.max_iterations > 0 && total >= P.max_iterations) go = false;
    }
    ch_flush(w, K, k, mcur, vcur, ln_err, sig_lo, dwell);
    if constexpr (has_z(KIND)) ch_flush_height(w, K, (R)(alt - alt_ref), dwell);

    ch_write_model(w, ml, k, mcur, vcur, P.out.cur_sigma ? P.out.cur_sigma + (size_t)chain * ml : nullptr,
                   P.out.cur_edges ? P.out.cur_edges + (size_t)chain * (ml + 1) : nullptr);
    if (lane == 0 && P.finish_ns) {
        unsigned long long tnow;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
        P.finish_ns[chain] = tnow;
    }
    if (lane == 0) {
        double* s = P.out.scalars + (size_t)chain * GBP_NSCALARS;
#pragma unroll 1
        for (int i = 0; i < GBP_NSCALARS; ++i) s[i] = 0.0;
        s[GBP_S_ITER] = (double)iteration;
        s[GBP_S_BURNED_IN] = burned_in;
        s[GBP_S_BURNED_IN_ITER] = (double)w->ctr[CT_BURN_ITER];
        s[GBP_S_BEST_ITER] = (double)w->ctr[CT_BEST_ITER];
        s[GBP_S_BEST_K] = w->ctr[CT_BEST_K];
        s[GBP_S_CUR_K] = k;
        s[GBP_S_HALFSPACE] = sigma_ref;
        s[GBP_S_FAILED] = failed ? 1.0 : 0.0;
        s[GBP_S_N_ACCEPT] = (double)w->ctr[CT_N_ACCEPT];
        s[GBP_S_N_FORWARD] = (double)w->ctr[CT_N_FWD];
        s[GBP_S_N_SENS] = (double)w->ctr[CT_N_SENS];
        // undo the data scaling: variances scale by data_scale^2, so the log-likelihood shifts by n ln(scale)
        // (Tempest: the additive unknowns are dimensionless multipliers - the scaling went into the additive levels)
        const double inv_sc = (KIND == KIND_TEMPEST) ? 1.0 : 1.0 / P.data_scale;
        const double lik_shift = (P.data_scale != 1.0) ? (double)n_active * dlog_(P.data_scale) : 0.0;
        s[GBP_S_BEST_POSTERIOR] = (double)w->bestv[BV_POSTERIOR] + lik_shift;
        s[GBP_S_CUR_REL] = (double)err.rel[0];
        s[GBP_S_CUR_ADD] = (double)err.add[0] * inv_sc;
        s[GBP_S_CUR_MISFIT] = (double)misfit;
        s[GBP_S_CUR_PRIOR] = (double)prior;
        s[GBP_S_CUR_LIKELIHOOD] = (double)likelihood + lik_shift;
        s[GBP_S_BEST_REL] = (double)w->bestv[BV_REL];
        s[GBP_S_BEST_ADD] = (double)w->bestv[BV_ADD] * inv_sc;
        if (NS > 1 && K->n_sys > 1) {
            s[GBP_S_CUR_REL2] = (double)err.rel[NS - 1];
            s[GBP_S_CUR_ADD2] = (double)err.add[NS - 1] * inv_sc;
            s[GBP_S_BEST_REL2] = (double)w->bestv[BV_REL2];
            s[GBP_S_BEST_ADD2] = (double)w->bestv[BV_ADD2] * inv_sc;
        }
        s[GBP_S_N_RESETS] = w->ctr[CT_N_RESETS];
        s[GBP_S_N_BIRTH] = (double)w->ctr[CT_ACT0];
        s[GBP_S_N_DEATH] = (double)w->ctr[CT_ACT1];
        s[GBP_S_N_MOVE] = (double)w->ctr[CT_ACT2];
        s[GBP_S_N_NONE] = (double)w->ctr[CT_ACT3];
        s[GBP_S_TOTAL_ITER] = (double)total;
        s[GBP_S_CUR_HEIGHT] = has_z(KIND) ? (double)alt : P.altitude[chain];
        s[GBP_S_BEST_HEIGHT] = has_z(KIND) ? (double)best_alt : P.altitude[chain];
        s[GBP_S_HEIGHT_REF] = has_z(KIND) ? (double)alt_ref : P.altitude[chain];
        if (spec_rounds > 0) {
            atomicAdd(&g_diag[8], (unsigned long long)w->ctr[CT_N_SPEC]);
            atomicAdd(&g_diag[9], (unsigned long long)spec_rounds);
            atomicAdd(&g_diag[10], (unsigned long long)spec_cycles);
        }
    }
    __syncwarp();
#undef N2
#undef GBP_BEST_SIG
#undef GBP_BEST_EDG
}

// ---------------------------------------------------------------- kernels
// R = sampler arithmetic, T = forward/Jacobian arithmetic, NC = channel capacity, WARPS = warps per CTA
// KIND_TDEM: S is the TdDev of the datapoint type and g_tab the window operator Mt (TD_ROWS x TD_CP)
template <typename R, typename T, int NC, int WARPS, int KIND>
__global__ void __launch_bounds__(WARPS * 32, 1)
    rjmcmc_kernel(const __grid_constant__ typename SysOf<T, KIND>::dev S, const T* __restrict__ g_tab,
                  const __grid_constant__ ChainParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ Consts<R> consts;
    __shared__ typename SysOf<T, KIND>::shared sys_s;
    __shared__ SpecRound<R, T, ns_of(KIND)> rounds[WARPS];
    __shared__ int mailbox[WARPS];
    __shared__ int n_alive, n_idle;
    __shared__ TeamShared teams[WARPS];
    T* tab = reinterpret_cast<T*>(smem);
    uint32_t tab_bytes;
    if constexpr (is_td(KIND)) tab_bytes = (uint32_t)(TD_ROWS * TD_CP * sizeof(T));
    else tab_bytes = fdem_table_bytes<T>(S);
    if (threadIdx.x == 0) {
        make_consts<R>(P.opt, P.n_depth, P.C, consts, is_td(KIND), KIND != KIND_TEMPEST);
        if constexpr (is_td(KIND)) {
            fill_td_shared<T>(S, sys_s);
            for (int i = 0; i < GBP_TD_MAXC; ++i) {
                consts.tsc[i] = (R)S.tsc[i];
                consts.csys[i] = (unsigned char)S.csys[i];
            }
        } else {
            fill_sys_shared<T>(S, sys_s);
        }
    }
    tma_stage(tab, g_tab, tab_bytes, &bar);
    __syncthreads();
    const uint32_t tab_pad = (tab_bytes + 127u) & ~127u;
    const int warp = threadIdx.x >> 5;
    WarpState<R, T, NC, KIND>* ws = reinterpret_cast<WarpState<R, T, NC, KIND>*>(smem + tab_pad) + warp;
    // persistent: every chain is claimed from a device-side counter.  Which WARPS take part in the first wave is static
    // (round-robin over the CTAs, one CTA per SM, so that a batch smaller than the machine still spreads evenly), which
    // chain each of them gets is not: when this launch overlaps the tail of the previous one on another stream, its CTAs
    // start one by one as SMs come free, and a CTA that starts late must not sit on chains the others could have run.
    const int lane = threadIdx.x & 31;
    int c = P.B;
#ifdef GBP_STATIC_FIRST_WAVE   // A/B build: the first wave dealt statically (chain = warp * grid + CTA), as in round 1
    c = warp * (int)gridDim.x + (int)blockIdx.x;
#else
    if (warp * (int)gridDim.x + (int)blockIdx.x < P.B) {
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(P.work_counter, 1);
        c = __shfl_sync(FULL, nxt, 0);
    }
#endif
    if (threadIdx.x == 0) {
        n_alive = 0;
        n_idle = 0;
    }
    // teams: T warps of the same SM sub-partition (warp ids congruent mod 4): team = (warp & 3) + 4 (warp / 4T),
    // member = (warp / 4) mod T; named barrier 1 + team
    int team_size = 1;
    if constexpr (!is_td(KIND) && sizeof(T) == 4) {
        const bool pow2 = P.team_size >= 2 && P.team_size <= 16 && (P.team_size & (P.team_size - 1)) == 0;
        if (pow2 && WARPS % (4 * P.team_size) == 0 && WARPS / P.team_size < 16) team_size = P.team_size;
        if (pow2 && P.team_spread && WARPS % P.team_size == 0 && WARPS / P.team_size < 16) team_size = P.team_size;
    }
    const int team_id = P.team_spread ? warp / team_size : (warp & 3) + 4 * (warp / (4 * team_size));
    const int team_member = P.team_spread ? warp % team_size : (warp >> 2) % team_size;
    if (lane == 0) {
        mailbox[warp] = MB_BUSY;
        for (int x = 0; x < 32; ++x) rounds[warp].midx[x] = 0;
        rounds[warp].seq = 0u;
        rounds[warp].released = 0u;
        ws->team = team_size > 1 ? &teams[team_id] : nullptr;
        ws->tm_size = team_size;
        ws->tm_bar = 1 + team_id;
        ws->req_active = 0;
        teams[warp].alive = 0;
        teams[warp].unit = 0;
        teams[warp].freq_units = P.team_freq_units;
    }
    __syncthreads();
    if (lane == 0) {
        if (c < P.B) atomicAdd(&n_alive, 1);
        if (team_size > 1) {
            teams[team_id].member[team_member] = (void*)ws;
            if (c < P.B) atomicAdd((int*)&teams[team_id].alive, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && P.finish_ns) {
        unsigned long long tnow;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
        atomicMin(P.finish_ns + P.B, tnow);
    }
    TailCtx<R, T, ns_of(KIND)> tc;
    tc.rounds = rounds;
    tc.mailbox = mailbox;
    tc.n_alive = &n_alive;
    tc.n_idle = &n_idle;
    tc.n_warps = WARPS;
    tc.max_helpers = P.spec_helpers;
    tc.warp = warp;
    const bool had_chain = c < P.B;
#pragma unroll 1
    while (c < P.B) {
        run_chain<R, T, NC, KIND>(ws, &consts, &sys_s, S, tab, P, c, tc);
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(P.work_counter, 1);
        c = __shfl_sync(FULL, nxt, 0);
    }
    if (team_size > 1) {
        // out of chains: keep evaluating forwards for the team mates until none of them owns a chain
        if (ws->team != nullptr) {   // (nullptr: this warp dissolved its team and finished alone)
            if (lane == 0) {
                if (had_chain) atomicSub((int*)&teams[team_id].alive, 1);
                ws->req_active = 0;
            }
            __syncwarp();
#pragma unroll 1
            while (team_round<R, T, NC, KIND>(ws, &sys_s, tab)) {
            }
            if (lane == 0) ws->team = nullptr;
            __syncwarp();
        }
    }
    if (P.spec_helpers > 0) {
        // out of chains: evaluate future iterations of the chains of this CTA that are still running
        if (lane == 0 && had_chain) atomicSub(&n_alive, 1);
        __syncwarp();
        tail_service<R, T, NC, KIND>(ws, tc, &consts, &sys_s, tab);
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
    const uint32_t tab_bytes = fdem_table_bytes<T>(S);
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
        fdem_run<T>(sys_s, tab, (T)altitude[b], L, msig, mthk, pred, SENS ? J : nullptr, SENS);
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
