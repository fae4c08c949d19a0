/* TEST INFRASTRUCTURE ONLY (oracle) - never linked or called by the product path.
 *
 * Plain-C fp64 restatement of the reference's per-sounding trans-dimensional
 * (reversible-jump) MCMC sampler for one sounding: an FDEM datapoint (one system) or a time-domain
 * datapoint with one or two systems (TdemDataPoint: per-system relative/additive errors, additive error
 * scaled by (t/1ms)^-0.5, TdemDataPoint.py:329-379).  One function call = one chain.
 *
 * PARITY PIN: the deterministic terms (Hessian, gradient, Newton mean, misfit, prior,
 * likelihood, forward/reverse proposal densities, posterior bin indices) are checked by
 * tests/test_oracle_golden.py against transition records captured from the live reference
 * in the build container (tests/golden/transitions.npz, made by tests/golden/make_golden.py),
 * and the stationary behaviour (posterior statistics, acceptance rate, layer-count
 * distribution) against reference chains (tests/golden/ref_chain_*.npz).  The reference's
 * NumPy random stream (PCG64DXSM + SVD-based multivariate_normal + data-dependent retry
 * loops) is not reproducible off-host, so the random stream here is our own
 * (Philox4x32-10, counter based) and is the bit-level twin of the CUDA kernel's.
 *
 * Reference followed (paths relative to geobipy/src/):
 *   inversion/Inference1D.py:353-464  initialize            -> chain_init()
 *   inversion/Inference1D.py:485-535  initialize_model      -> chain_init()
 *   classes/data/datapoint/EmDataPoint.py:148-186 find_best_halfspace -> best_halfspace()
 *   inversion/Inference1D.py:537-631  accept_reject         -> chain_step()
 *   inversion/Inference1D.py:633-688  infer (loop control)  -> gbo_run_chain()
 *   inversion/Inference1D.py:705-790  update                -> chain_update()
 *   classes/mesh/RectilinearMesh1D.py:993-1120 perturb      -> perturb_structure()
 *   classes/mesh/RectilinearMesh1D.py:643-689 delete_edge, :805-838 insert_edge
 *   classes/mesh/RectilinearMesh1D.py:691-714 gradient, :747-786 gradient_operator
 *   classes/model/Model.py:368-419 stochastic_newton_perturbation, :421-430 prior_derivative
 *   classes/model/Model.py:533-575 probability, :213-234 gradient_probability
 *   classes/model/Model.py:577-660 proposal_probabilities
 *   classes/model/Model.py:819-847 update_parameter_posterior
 *   classes/mesh/RectilinearMesh1D.py:1122-1160 piecewise_constant_interpolate
 *   classes/mesh/RectilinearMesh1D.py:1594-1610 update_posteriors
 *   classes/data/datapoint/DataPoint.py:268-282 std, :340-349 prior_derivative,
 *       :351-395 probability, :491-525 likelihood / data_misfit, :531-573 perturb
 *   classes/statistics/StatArray.py:578-638 propose (<= 10 prior-respecting retries)
 *   classes/statistics/MvNormalDistribution.py:183-216 log-pdf
 *   classes/statistics/CategoricalDistribution.py:67-82 rng
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define LOG2PI 1.8378770664093454835606594728112

/* ------------------------------------------------------------------ Philox4x32-10 */
void gbo_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Random stream of one chain: Philox4x32-10, key = seed, counter = (block, iteration, sounding lo, sounding hi).
 * Every accept_reject() call (iteration = 1, 2, ... counted over resets) starts its own sub-stream at block 0, so
 * that the numbers an iteration draws do not depend on how many the previous ones consumed (the CUDA sampler
 * evaluates future iterations of a chain speculatively on idle warps). */
typedef struct {
    uint64_t seed, sounding;
    uint32_t block, iteration;
} rng_t;

/* One Philox block -> two 53-bit uniforms in [0,1). */
static void rng_block(rng_t *g, double *ua, double *ub)
{
    uint32_t ctr[4] = {g->block, g->iteration, (uint32_t)g->sounding, (uint32_t)(g->sounding >> 32)};
    uint32_t key[2] = {(uint32_t)g->seed, (uint32_t)(g->seed >> 32)};
    uint32_t x[4];
    gbo_philox4x32_10(ctr, key, x);
    g->block++;
    *ua = ((double)(x[0] >> 5) * 67108864.0 + (double)(x[1] >> 6)) * (1.0 / 9007199254740992.0);
    *ub = ((double)(x[2] >> 5) * 67108864.0 + (double)(x[3] >> 6)) * (1.0 / 9007199254740992.0);
}
static double rng_uniform(rng_t *g)
{
    double a, b;
    rng_block(g, &a, &b);
    return a;
}
/* Box-Muller pair from one block. */
static void rng_normal2(rng_t *g, double *z0, double *z1)
{
    double a, b;
    rng_block(g, &a, &b);
    double r = sqrt(-2.0 * log(1.0 - a));
    double th = 2.0 * M_PI * b;
    *z0 = r * cos(th);
    *z1 = r * sin(th);
}
static double rng_normal(rng_t *g)
{
    double z0, z1;
    rng_normal2(g, &z0, &z1);
    return z0;
}

/* ------------------------------------------------------------------ small dense SPD algebra */
/* In-place lower Cholesky of row-major n x n (ld = n).  Returns 0 on success. */
static int cholesky(int n, double *A)
{
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int p = 0; p < j; ++p) d -= A[j * n + p] * A[j * n + p];
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int p = 0; p < j; ++p) s -= A[i * n + p] * A[j * n + p];
            A[i * n + j] = s / d;
        }
    }
    return 0;
}
/* Solve L y = b (forward) */
static void solve_L(int n, const double *Lm, const double *b, double *y)
{
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int p = 0; p < i; ++p) s -= Lm[i * n + p] * y[p];
        y[i] = s / Lm[i * n + i];
    }
}
/* Solve L^T x = y (backward) */
static void solve_LT(int n, const double *Lm, const double *y, double *x)
{
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int p = i + 1; p < n; ++p) s -= Lm[p * n + i] * x[p];
        x[i] = s / Lm[i * n + i];
    }
}
/* |L^T v|^2 = v' A v */
static double quad_L(int n, const double *Lm, const double *v)
{
    double q = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = j; i < n; ++i) s += Lm[i * n + j] * v[i];
        q += s * s;
    }
    return q;
}

/* ------------------------------------------------------------------ model / datapoint pieces */
typedef struct {
    int k;
    double edges[GBO_MAXL + 2]; /* edges[0] = 0, edges[k] = inf */
    double sigma[GBO_MAXL + 1];
} model_t;

typedef struct {
    double rel[GBO_MAXSYS], add[GBO_MAXSYS];
    double pred[GBO_MAXC];
    double J[GBO_MAXC * GBO_MAXL]; /* row-major [C][k] */
    int Jk;                        /* columns of J */
    double z;                      /* sensor height (Point.z); constant unless solve_height */
} dpoint_t;

static void model_thickness(const model_t *m, double *thk)
{
    for (int i = 0; i < m->k; ++i) thk[i] = m->edges[i + 1] - m->edges[i];
}

int gbo_n_depth(const gbo_options *o)
{
    /* numpy.arange(0, 1.1*max_edge, 0.5*min_width) has ceil(stop/step) entries = edges
     * (RectilinearMesh1D.py:1450) -> cells = entries - 1 */
    double stop = 1.1 * o->max_edge, step = 0.5 * o->min_width;
    int n_edges = (int)ceil(stop / step);
    return n_edges - 1;
}

/* what kind of datapoint the chain inverts */
typedef struct {
    const gbo_fdem_system *fdem;   /* frequency domain (one system) ... */
    const gbo_tdem_system *tdem;   /* ... or time domain (1-2 systems) */
    int C, n_sys;
    int sys_of[GBO_MAXC];          /* system of channel c */
    double log_t[GBO_MAXC];        /* ln(off_time) of channel c (TDEM) */
} survey_t;

static void survey_fdem(survey_t *v, const gbo_fdem_system *sys)
{
    memset(v, 0, sizeof(*v));
    v->fdem = sys;
    v->C = 2 * sys->n_freq;
    v->n_sys = 1;
}
static void survey_tdem(survey_t *v, const gbo_tdem_system *sys)
{
    memset(v, 0, sizeof(*v));
    v->tdem = sys;
    v->C = sys->C;
    v->n_sys = sys->n_sys;
    if (sys->tempest) {
        /* one system, errors per COMPONENT in channel order (x, then z): Tempest_datapoint.relative_error :129-139 */
        int has_x = 0;
        for (int c = 0; c < sys->C; ++c) has_x |= (sys->comp[c] == 1);
        v->n_sys = has_x ? 2 : 1;
        for (int c = 0; c < sys->C; ++c) {
            v->sys_of[c] = (sys->comp[c] == 1) ? 0 : has_x;
            v->log_t[c] = log(sys->t_centre[c]);
        }
        return;
    }
    int c = 0;
    for (int s = 0; s < sys->n_sys; ++s)
        for (int i = 0; i < sys->n_win[s]; ++i, ++c) {
            v->sys_of[c] = s;
            v->log_t[c] = log(sys->t_centre[c]);
        }
}

/* per-system option accessors (skytem_options gives lists, one entry per system) */
static double o_rel_init(const gbo_options *o, int s) { return s ? o->rel_init2 : o->rel_init; }
static double o_rel_min(const gbo_options *o, int s) { return s ? o->rel_min2 : o->rel_min; }
static double o_rel_max(const gbo_options *o, int s) { return s ? o->rel_max2 : o->rel_max; }
static double o_rel_var(const gbo_options *o, int s) { return s ? o->rel_prop_var2 : o->rel_prop_var; }
static double o_add_init(const gbo_options *o, int s) { return s ? o->add_init2 : o->add_init; }
static double o_add_min(const gbo_options *o, int s) { return s ? o->add_min2 : o->add_min; }
static double o_add_max(const gbo_options *o, int s) { return s ? o->add_max2 : o->add_max; }
static double o_add_var(const gbo_options *o, int s) { return s ? o->add_prop_var2 : o->add_prop_var; }

/* DataPoint.std (DataPoint.py:268-282): variance_i = (rel * d_i)^2 + add^2.
 * TdemDataPoint.std (TdemDataPoint.py:329-379): per system rel / add, and
 * add_i = exp(ln(add) - 0.5 (ln t_i - ln 1e-3)). */
static void data_variance(const survey_t *v, const double *data, const double *rel, const double *add, double *var)
{
    for (int i = 0; i < v->C; ++i) {
        if (v->tdem && v->tdem->tempest) {
            /* Tempest_datapoint.std :170-173: (rel_j d)^2 + (multiplier_j additive_c)^2 */
            const int s = v->sys_of[i];
            double a = rel[s] * data[i];
            double b = add[s] * v->tdem->add_level[i];
            var[i] = a * a + b * b;
        } else if (v->tdem) {
            const int s = v->sys_of[i];
            double a = rel[s] * data[i];
            double b = exp(log(add[s]) - 0.5 * (v->log_t[i] - log(1e-3)));
            var[i] = a * a + b * b;
        } else {
            double a = rel[0] * data[i];
            var[i] = a * a + add[0] * add[0];
        }
    }
}

/* EmDataPoint.active (EmDataPoint.py:44-56): observed > 0 and not NaN */
static int is_active(double d) { return d > 0.0; }

static double data_misfit(int C, const double *data, const double *pred, const double *var)
{
    double s = 0.0;
    for (int i = 0; i < C; ++i)
        if (is_active(data[i])) {
            double t = (1.0 / sqrt(var[i])) * (pred[i] - data[i]);
            s += t * t;
        }
    return s;
}

/* MvNormal log-pdf with diagonal covariance (MvNormalDistribution.py:209-216) */
static double data_likelihood(int C, const double *data, const double *pred, const double *var)
{
    double n = 0.0, logdet = 0.0, q = 0.0;
    for (int i = 0; i < C; ++i)
        if (is_active(data[i])) {
            n += 1.0;
            logdet += log(var[i]);
            double r = pred[i] - data[i];
            q += r * r / var[i];
        }
    return -(0.5 * n) * LOG2PI - 0.5 * logdet - 0.5 * q;
}

/* Uniform(log=True).probability(log=True) (UniformDistribution.py:109-121) */
static double log_uniform_logpdf(double x, double mn, double mx)
{
    double lx = log(x), a = log(mn), b = log(mx);
    if (lx < a || lx > b) return -INFINITY;
    return -log(b - a);
}

/* DataPoint.probability :351-395: multivariate Uniform(log=True) priors sum over the systems
 * (UniformDistribution.py:116) */
static double height_logprior(const gbo_options *o, double dz)
{
    return (dz >= -o->max_height_change && dz <= o->max_height_change) ? -log(2.0 * o->max_height_change) : -INFINITY;
}

/* height_last = 0 (frequency domain, solve_z): Point.probability :160-196 first (height prior, dz = z - z_ref), then the
 * errors (DataPoint.probability :352-389).  height_last = 1 (time domain, solve_transmitter_z): the sampled height is the
 * transmitter loop's and its prior is added after the errors (TdemDataPoint.probability :950-951 =
 * DataPoint.probability + Loop_pair.probability, Loop_pair.py:294-295). */
static double datapoint_probability(const gbo_options *o, const survey_t *v, const double *rel, const double *add, double dz)
{
    const int n_sys = v->n_sys, height_last = v->tdem != NULL;
    /* Tempest: the additive-error multiplier carries a prior only for its histogram bins; DataPoint.probability :385-389
     * never sees it (additive_error itself has none, Tempest_datapoint.set_priors :478-488) */
    const int add_prior = !(v->tdem && v->tdem->tempest);
    double p = 0.0;
    if (o->solve_height && !height_last) p += height_logprior(o, dz);
    for (int s = 0; s < n_sys; ++s) {
        if (o->solve_relative_error) p += log_uniform_logpdf(rel[s], o_rel_min(o, s), o_rel_max(o, s));
        if (o->solve_additive_error && add_prior) p += log_uniform_logpdf(add[s], o_add_min(o, s), o_add_max(o, s));
    }
    if (o->solve_height && height_last) p += height_logprior(o, dz);
    return p;
}

/* Model.probability (Model.py:533-575) with value_bounds = None */
static double model_probability(const gbo_options *o, const model_t *m, double sigma_ref)
{
    /* Uniform(1, kmax).probability(k) = scipy uniform.logpdf(k, 1, kmax - 1) */
    double p = (m->k >= 1 && m->k <= o->max_layers) ? -log((double)o->max_layers - 1.0) : -INFINITY;
    if (o->solve_parameter) {
        double s2 = log(1.0 + o->factor);
        s2 *= s2;
        double q = 0.0;
        for (int i = 0; i < m->k; ++i) {
            double d = log(m->sigma[i]) - log(sigma_ref);
            q += d * d / s2;
        }
        p += -(0.5 * m->k) * LOG2PI - 0.5 * m->k * log(s2) - 0.5 * q;
    }
    if (o->solve_gradient) {
        double g2 = o->gradient_std * o->gradient_std;
        if (m->k == 1) {
            /* Model.py:230-232: a virtual 2-layer model with equal values -> gradient 0 */
            p += -0.5 * LOG2PI - 0.5 * log(g2);
        } else {
            int n = m->k - 1;
            double q = 0.0;
            for (int i = 0; i < n; ++i) {
                /* RectilinearMesh1D.py:713: diff(ln sigma) / ln(width_i) */
                double g = (log(m->sigma[i + 1]) - log(m->sigma[i])) / log(m->edges[i + 1] - m->edges[i]);
                q += g * g / g2;
            }
            p += -(0.5 * n) * LOG2PI - 0.5 * n * log(g2) - 0.5 * q;
        }
    }
    return p;
}

/* Wm'Wm = values-prior precision + Wz' (1/g^2) Wz (Model.py:421-430, RectilinearMesh1D.py:747-786) */
static void prior_operator(const gbo_options *o, const model_t *m, double *A)
{
    const int k = m->k;
    double s2 = log(1.0 + o->factor);
    s2 *= s2;
    const double g2 = o->gradient_std * o->gradient_std;
    memset(A, 0, sizeof(double) * k * k);
    for (int i = 0; i < k; ++i) A[i * k + i] = 1.0 / s2;
    if (k == 1) {
        A[0] += 1.0 / g2; /* gradient_operator = ones((1,1)) */
        return;
    }
    double x[GBO_MAXL];
    model_thickness(m, x);
    if (k == 2) x[k - 1] = x[0];
    else x[k - 1] = x[k - 2] + (m->edges[k - 1] - m->edges[0]);
    for (int i = 0; i < k - 1; ++i) {
        double c2c = 0.5 * (x[i] + x[i + 1]);
        double t = 1.0 / (c2c * (double)(k - 1));
        double t2 = t * t / g2;
        A[i * k + i] += t2;
        A[(i + 1) * k + (i + 1)] += t2;
        A[i * k + (i + 1)] -= t2;
        A[(i + 1) * k + i] -= t2;
    }
}

/* A = Wm'Wm + J' Wd'Wd J ; g = Wm'Wm (ln sigma - ln sigma_ref) + J' Wd'Wd (pred - data)
 * (Model.py:250-272, :347-357; DataPoint.py:340-349) */
static void hessian_gradient(const gbo_options *o, const model_t *m, double sigma_ref, int C, const double *data,
                             const double *var, const double *J, const double *pred, double *A, double *g)
{
    const int k = m->k;
    double P[GBO_MAXL * GBO_MAXL];
    prior_operator(o, m, P);
    for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += P[i * k + j] * (log(m->sigma[j]) - log(sigma_ref));
        g[i] = s;
    }
    if (A) memcpy(A, P, sizeof(double) * k * k);
    for (int c = 0; c < C; ++c) {
        if (!is_active(data[c])) continue;
        double w = 1.0 / var[c];
        double r = (pred[c] - data[c]) * w;
        for (int i = 0; i < k; ++i) {
            g[i] += J[c * k + i] * r;
            if (A)
                for (int j = 0; j < k; ++j) A[i * k + j] += J[c * k + i] * w * J[c * k + j];
        }
    }
}

/* ------------------------------------------------------------------ structure proposal */
enum { ACT_BIRTH = 0, ACT_DEATH = 1, ACT_MOVE = 2, ACT_NONE = 3 };

static double min_width_of(int nedges, const double *z)
{
    double h = INFINITY;
    for (int i = 0; i + 1 < nedges; ++i) {
        double d = z[i + 1] - z[i];
        if (d < h) h = d;
    }
    return h;
}

/* RectilinearMesh1D.perturb (:993-1120).  Returns the action; writes the remapped model. */
static int perturb_structure(const gbo_options *o, rng_t *g, const model_t *cur, model_t *out)
{
    const double cum0 = o->p_birth, cum1 = cum0 + o->p_death, cum2 = cum1 + o->p_move, cum3 = cum2 + o->p_none;
    const int k = cur->k;
    for (;;) {
        int event;
        for (;;) {
            /* Categorical.rng: searchsorted(cumsum(p), U) (left) */
            double u = rng_uniform(g);
            event = (u <= cum0) ? 0 : (u <= cum1) ? 1 : (u <= cum2) ? 2 : 3;
            (void)cum3;
            if (k == 1 && (event == 1 || event == 2)) continue;
            if (k == o->max_layers && event == 0) continue;
            break;
        }
        if (event == ACT_NONE) {
            *out = *cur;
            return ACT_NONE;
        }
        if (event == ACT_BIRTH) {
            int ok = 0, pos = 0;
            double e = 0.0;
            double z[GBO_MAXL + 3];
            for (int tries = 1; tries <= 10; ++tries) {
                double lo = log(o->min_edge), hi = log(o->max_edge);
                e = exp(lo + (hi - lo) * rng_uniform(g));
                pos = 0; /* searchsorted(edges, e) (left) */
                while (pos <= k && cur->edges[pos] < e) ++pos;
                for (int i = 0; i < pos; ++i) z[i] = cur->edges[i];
                z[pos] = e;
                for (int i = pos; i <= k; ++i) z[i + 1] = cur->edges[i];
                double h = min_width_of(k + 2, z);
                if (tries == 10) break; /* 10th try always restarts (:1078-1080) */
                if (h > o->min_width) { ok = 1; break; }
            }
            if (!ok) continue;
            out->k = k + 1;
            memcpy(out->edges, z, sizeof(double) * (k + 2));
            /* values.insert(pos, values[pos-1]) (:835) */
            for (int i = 0; i < pos; ++i) out->sigma[i] = cur->sigma[i];
            out->sigma[pos] = cur->sigma[pos - 1];
            for (int i = pos; i < k; ++i) out->sigma[i + 1] = cur->sigma[i];
            return ACT_BIRTH;
        }
        if (event == ACT_DEATH) {
            int i = (int)(rng_uniform(g) * (double)(k - 1)) + 1; /* :1085 */
            out->k = k - 1;
            for (int j = 0; j < i; ++j) out->edges[j] = cur->edges[j];
            for (int j = i + 1; j <= k; ++j) out->edges[j - 1] = cur->edges[j];
            double val = 0.5 * (cur->sigma[i - 1] + cur->sigma[i]); /* :684-686 */
            for (int j = 0; j < i; ++j) out->sigma[j] = cur->sigma[j];
            for (int j = i + 1; j < k; ++j) out->sigma[j - 1] = cur->sigma[j];
            out->sigma[i - 1] = val;
            return ACT_DEATH;
        }
        /* ACT_MOVE (:1088-1118) */
        {
            int ok = 0;
            double z[GBO_MAXL + 2];
            for (int tries = 1; tries <= 10; ++tries) {
                memcpy(z, cur->edges, sizeof(double) * (k + 1));
                int i = (int)(1.0 + ((double)k - 1.0) * rng_uniform(g)); /* uniform(1, nEdges-1) */
                double zn = rng_normal(g);
                double sgn = (zn > 0.0) ? 1.0 : (zn < 0.0 ? -1.0 : 0.0);
                double dz = sgn * o->min_width * rng_uniform(g);
                z[i] += dz;
                double h = min_width_of(k + 1, z);
                if (tries == 10) break;
                if (h > o->min_width && z[1] > o->min_edge && z[k - 1] < o->max_edge) { ok = 1; break; }
            }
            if (!ok) continue;
            *out = *cur;
            memcpy(out->edges, z, sizeof(double) * (k + 1));
            return ACT_MOVE;
        }
    }
}

/* StatArray.propose with imposePrior (StatArray.py:578-638) for a 1-D MvLogNormal random walk */
static double propose_error(rng_t *g, double cur, double prop_var, double mn, double mx)
{
    double sd = sqrt(prop_var);
    double x = exp(log(cur) + sd * rng_normal(g));
    int tries = 0;
    while (log_uniform_logpdf(x, mn, mx) == -INFINITY) {
        x = exp(log(cur) + sd * rng_normal(g));
        tries++;
        if (tries == 10) return cur;
    }
    return x;
}

/* Point.perturb :614-622 = StatArray.propose with imposePrior for the Normal(z, var) height random walk under the
 * Uniform[z_ref - dz, z_ref + dz] prior (scipy uniform.logpdf: closed support) */
static double propose_height(rng_t *g, double cur, double prop_var, double lo, double hi)
{
    double sd = sqrt(prop_var);
    double x = cur + sd * rng_normal(g);
    int tries = 0;
    while (!(x >= lo && x <= hi)) {
        x = cur + sd * rng_normal(g);
        tries++;
        if (tries == 10) return cur;
    }
    return x;
}

/* Same for a dual-moment datapoint: the 2-vector is proposed jointly (MvLogNormal, diagonal variance) and
 * re-drawn, at most 10 times, while ANY component leaves its prior; then the whole vector falls back
 * (StatArray.py:619-636).  One Box-Muller pair per draw. */
static void propose_error2(rng_t *g, double *x, const double *var, const double *mn, const double *mx)
{
    const double c0 = x[0], c1 = x[1];
    double z0, z1;
    rng_normal2(g, &z0, &z1);
    x[0] = exp(log(c0) + sqrt(var[0]) * z0);
    x[1] = exp(log(c1) + sqrt(var[1]) * z1);
    int tries = 0;
    while (log_uniform_logpdf(x[0], mn[0], mx[0]) == -INFINITY || log_uniform_logpdf(x[1], mn[1], mx[1]) == -INFINITY) {
        rng_normal2(g, &z0, &z1);
        x[0] = exp(log(c0) + sqrt(var[0]) * z0);
        x[1] = exp(log(c1) + sqrt(var[1]) * z1);
        tries++;
        if (tries == 10) {
            x[0] = c0;
            x[1] = c1;
            return;
        }
    }
}

/* ------------------------------------------------------------------ posterior accumulators */
typedef struct {
    int n_depth, n_sig, n_err, kmax;
    double depth_step;
    double sig_lo, sig_dx;   /* ln(sigma) bins */
    double rel_lo[GBO_MAXSYS], rel_dx[GBO_MAXSYS], add_lo[GBO_MAXSYS], add_dx[GBO_MAXSYS]; /* ln(err) bins */
    int n_sys;
    double z_lo, z_dx;       /* height bins, relative to the height the prior was centred on */
} grids_t;

static void make_grids(const gbo_options *o, int n_sys, double sigma_ref, grids_t *G)
{
    G->n_sys = n_sys;
    G->n_depth = gbo_n_depth(o);
    G->depth_step = 0.5 * o->min_width;
    G->n_sig = o->n_sigma_bins;
    G->n_err = o->n_err_bins;
    G->kmax = o->max_layers;
    /* Model.set_posteriors :666-684 + MvLogNormal.bins: linspace(-nStd*s, nStd*s, n+1) + ln(sigma_ref) */
    double s = log(1.0 + o->factor);
    G->sig_lo = log(sigma_ref) - o->sigma_bins_nstd * s;
    G->sig_dx = 2.0 * o->sigma_bins_nstd * s / (double)G->n_sig;
    /* DataPoint.set_relative_error_posterior :668-695 + Uniform.bins: linspace(ln min, ln max, n+1) */
    for (int s = 0; s < n_sys; ++s) {
        G->rel_lo[s] = log(o_rel_min(o, s));
        G->rel_dx[s] = (log(o_rel_max(o, s)) - log(o_rel_min(o, s))) / (double)G->n_err;
        G->add_lo[s] = log(o_add_min(o, s));
        G->add_dx[s] = (log(o_add_max(o, s)) - log(o_add_min(o, s))) / (double)G->n_err;
    }
    /* Point.set_z_posterior :1013-1020: Uniform.bins = linspace(z0 - dz, z0 + dz, 100), mesh relative_to z0 */
    G->z_lo = -o->max_height_change;
    G->z_dx = 2.0 * o->max_height_change / (double)G->n_err;
}

/* searchsorted(edges, v, 'right') - 1 clipped, for uniform edges lo + i*dx */
static int uniform_bin(double v, double lo, double dx, int n)
{
    double f = floor((v - lo) / dx);
    int i = (f < 0.0) ? 0 : (f > (double)(n - 1) ? n - 1 : (int)f);
    /* guard against round-off of the division at an edge */
    while (i + 1 < n && v >= lo + (double)(i + 1) * dx) ++i;
    while (i > 0 && v < lo + (double)i * dx) --i;
    return i;
}

/* value at depth y of np.interp on the duplicated-edge staircase (RectilinearMesh1D.py:1148-1158) */
static double staircase_value(const gbo_options *o, const model_t *m, double y)
{
    const int k = m->k;
    if (k == 1) return m->sigma[0];
    /* xp = [e0*1.000001, e1, e1*1.000001, e2, ..., e_k := max_edge], fp = [v0,v0,v1,v1,...] */
    for (int i = 1; i < k; ++i) {
        double e = m->edges[i];
        double e2 = e * 1.000001;
        if (y < e) return m->sigma[i - 1];
        if (y < e2) { /* linear ramp between the two layers */
            double t = (y - e) / (e2 - e);
            return m->sigma[i - 1] + t * (m->sigma[i] - m->sigma[i - 1]);
        }
    }
    (void)o;
    return m->sigma[k - 1];
}

static void accumulate_posteriors(const gbo_options *o, const grids_t *G, const model_t *m, const double *rel,
                                  const double *add, double dz, gbo_chain_out *out)
{
    /* height histogram (Point.update_posteriors :1022-1025); dz = z - z_ref */
    if (o->solve_height && out->height_hist) out->height_hist[uniform_bin(dz, G->z_lo, G->z_dx, G->n_err)] += 1;
    /* nCells histogram (RectilinearMesh1D.py:1597) */
    out->ncells_hist[m->k] += 1;
    /* interface histogram (:1600-1610) with ratio = 0.5 */
    for (int i = 1; i < m->k; ++i) {
        double r = exp(log(m->sigma[i]) - log(m->sigma[i - 1]));
        if (r <= 0.5 || r >= 1.5) {
            double d = m->edges[i];
            if (d >= 0.0 && d < (double)G->n_depth * G->depth_step) {
                int j = uniform_bin(d, 0.0, G->depth_step, G->n_depth);
                out->edges_hist[j] += 1;
            }
        }
    }
    /* hitmap (Model.py:819-847) */
    for (int j = 0; j < G->n_depth; ++j) {
        double y = ((double)j + 0.5) * G->depth_step;
        double v = staircase_value(o, m, y);
        int b = uniform_bin(log(v), G->sig_lo, G->sig_dx, G->n_sig);
        out->hitmap[(size_t)b * G->n_depth + j] += 1;
    }
    /* error histograms (EmDataPoint.py:225-239) */
    for (int s = 0; s < G->n_sys; ++s) {
        if (o->solve_relative_error)
            out->rel_hist[s * G->n_err + uniform_bin(log(rel[s]), G->rel_lo[s], G->rel_dx[s], G->n_err)] += 1;
        if (o->solve_additive_error)
            out->add_hist[s * G->n_err + uniform_bin(log(add[s]), G->add_lo[s], G->add_dx[s], G->n_err)] += 1;
    }
}

static void reset_posteriors(const gbo_options *o, const grids_t *G, gbo_chain_out *out)
{
    memset(out->hitmap, 0, sizeof(int32_t) * (size_t)G->n_sig * G->n_depth);
    memset(out->edges_hist, 0, sizeof(int32_t) * G->n_depth);
    memset(out->ncells_hist, 0, sizeof(int32_t) * (o->max_layers + 1));
    memset(out->rel_hist, 0, sizeof(int32_t) * G->n_sys * G->n_err);
    memset(out->add_hist, 0, sizeof(int32_t) * G->n_sys * G->n_err);
    if (o->solve_height && out->height_hist) memset(out->height_hist, 0, sizeof(int32_t) * G->n_err);
}

/* ------------------------------------------------------------------ the chain */
typedef struct {
    survey_t sv;
    const gbo_options *o;
    int C;
    double data[GBO_MAXC];
    double altitude;
    double z_ref;      /* centre of the height prior: the datapoint's height when the priors were (re)set */
    double sigma_ref;
    grids_t G;
    rng_t rng;
    model_t model;
    dpoint_t dp;
    double misfit, prior, likelihood, posterior;
    int64_t iteration;
    int burned_in;
    int64_t burned_in_iter, best_iter;
    model_t best_model;
    double best_rel[GBO_MAXSYS], best_add[GBO_MAXSYS], best_z, best_posterior;
    int accepted;
    int n_zero_acc, n_resets, limiters;
    int64_t n_accept, n_forward, n_sens, n_act[4];
    int n_active;
} chain_t;

static void forward(chain_t *c, const model_t *m, double z, double *pred)
{
    double thk[GBO_MAXL];
    model_thickness(m, thk);
    if (c->sv.tdem) {
        gbo_tdem_forward(c->sv.tdem, z, m->k, m->sigma, thk, pred);
        if (c->sv.tdem->tempest)   /* predictedData = predicted secondary + predicted primary field (:120-127) */
            for (int i = 0; i < c->sv.C; ++i) pred[i] += c->sv.tdem->primary[i];
    } else gbo_fdem_forward(c->sv.fdem, z, m->k, m->sigma, thk, pred);
    c->n_forward++;
}
static void sensitivity(chain_t *c, const model_t *m, dpoint_t *dp)
{
    double thk[GBO_MAXL];
    model_thickness(m, thk);
    /* at the height the datapoint holds NOW: fm_dlogc(remapped) runs before datapoint.perturb() */
    if (c->sv.tdem) gbo_tdem_sensitivity(c->sv.tdem, dp->z, m->k, m->sigma, thk, dp->J);
    else gbo_fdem_sensitivity(c->sv.fdem, dp->z, m->k, m->sigma, thk, dp->J);
    dp->Jk = m->k;
    c->n_sens++;
}

/* EmDataPoint.find_best_halfspace: argmin of misfit over logspace(-4, 4, 100) */
static double best_halfspace(chain_t *c, const double *rel, const double *add, double z)
{
    double var[GBO_MAXC], pred[GBO_MAXC];
    data_variance(&c->sv, c->data, rel, add, var);
    model_t m;
    m.k = 1;
    m.edges[0] = 0.0;
    m.edges[1] = INFINITY;
    double best = INFINITY, best_c = 0.0;
    for (int i = 0; i < 100; ++i) {
        /* numpy.logspace(-4, 4, 100) = 10 ** linspace(-4, 4, 100) */
        double e = -4.0 + (double)i * (8.0 / 99.0);
        if (i == 99) e = 4.0;
        m.sigma[0] = pow(10.0, e);
        forward(c, &m, z, pred);
        double phi = data_misfit(c->C, c->data, pred, var);
        if (phi < best) { best = phi; best_c = m.sigma[0]; }
    }
    return best_c;
}

static void chain_init(chain_t *c, gbo_chain_out *out)
{
    const gbo_options *o = c->o;
    for (int s = 0; s < c->sv.n_sys; ++s) {
        c->dp.rel[s] = o_rel_init(o, s);
        c->dp.add[s] = o_add_init(o, s);
    }
    /* Inference1D.reset :984-994 re-initialises with the CURRENT datapoint: its height is kept and the height prior,
     * proposal and posterior bins are re-centred on it (Point.set_priors :959-961) */
    c->z_ref = c->dp.z;
    c->sigma_ref = best_halfspace(c, c->dp.rel, c->dp.add, c->dp.z);
    c->model.k = 1;
    c->model.edges[0] = 0.0;
    c->model.edges[1] = INFINITY;
    c->model.sigma[0] = c->sigma_ref;
    forward(c, &c->model, c->dp.z, c->dp.pred);
    sensitivity(c, &c->model, &c->dp);
    make_grids(o, c->sv.n_sys, c->sigma_ref, &c->G);
    reset_posteriors(o, &c->G, out);
    memset(out->misfit_trace, 0, sizeof(double) * 2 * (size_t)o->n_markov_chains);
    memset(out->accept_trace, 0, 2 * (size_t)o->n_markov_chains);
    double var[GBO_MAXC];
    data_variance(&c->sv, c->data, c->dp.rel, c->dp.add, var);
    c->misfit = data_misfit(c->C, c->data, c->dp.pred, var);
    c->prior = model_probability(o, &c->model, c->sigma_ref) + datapoint_probability(o, &c->sv, c->dp.rel, c->dp.add, 0.0);
    c->likelihood = data_likelihood(c->C, c->data, c->dp.pred, var);
    c->posterior = c->likelihood + c->prior;
    c->burned_in = 0;
    c->burned_in_iter = 0;
    c->iteration = 0;
    out->misfit_trace[0] = c->misfit;
    c->accepted = 0;
    c->best_model = c->model;
    memcpy(c->best_rel, c->dp.rel, sizeof(c->best_rel));
    memcpy(c->best_add, c->dp.add, sizeof(c->best_add));
    c->best_z = c->dp.z;
    c->best_posterior = c->posterior;
    c->best_iter = 0;
    c->n_zero_acc = 0;
}

/* Inference1D.accept_reject.  Returns 1 if the chain failed (singular Hessian). */
static int chain_step(chain_t *c)
{
    const gbo_options *o = c->o;
    const int C = c->C;
    model_t remap, test;
    dpoint_t tdp = c->dp; /* deepcopy(self.datapoint) */
    double var[GBO_MAXC], A[GBO_MAXL * GBO_MAXL], grad[GBO_MAXL], y[GBO_MAXL], step[GBO_MAXL], z[GBO_MAXL + 1];

    c->accepted = 0;
    c->rng.iteration++;
    c->rng.block = 0;
    int action = perturb_structure(o, &c->rng, &c->model, &remap);
    c->n_act[action]++;
    const int k = remap.k;

    if (action != ACT_NONE) { /* observation.fm_dlogc(remapped) */
        /* REFERENCE QUIRK (kept, it shapes the proposal): FdemDataPoint.fm_dlogc stores the predicted data of the
         * remapped model (FdemDataPoint.py:535-545), TdemDataPoint.fm_dlogc only stores the Jacobian - its
         * predicted-data update is commented out (TdemDataPoint.py:1031-1055) - so the Newton gradient of a
         * time-domain death / move uses the CURRENT model's predicted data with the remapped model's Jacobian. */
        double scratch[GBO_MAXC];
        forward(c, &remap, tdp.z, c->sv.tdem ? scratch : tdp.pred);
        sensitivity(c, &remap, &tdp);
    }
    data_variance(&c->sv, c->data, tdp.rel, tdp.add, var);
    hessian_gradient(o, &remap, c->sigma_ref, C, c->data, var, tdp.J, tdp.pred, A, grad);
    if (cholesky(k, A)) return 1;
    /* pk = -H grad ; mean = ln sigma + alpha pk */
    solve_L(k, A, grad, y);
    solve_LT(k, A, y, step);
    double mean[GBO_MAXL];
    for (int i = 0; i < k; ++i) mean[i] = log(remap.sigma[i]) - o->covariance_scaling * step[i];
    /* sigma' ~ exp(N(mean, H)) with H = A^-1 = L^-T L^-1  ->  mean + L^-T z */
    for (int j = 0; j < (k + 1) / 2; ++j) rng_normal2(&c->rng, &z[2 * j], &z[2 * j + 1]);
    double dx[GBO_MAXL];
    solve_LT(k, A, z, dx);
    test = remap;
    for (int i = 0; i < k; ++i) test.sigma[i] = exp(mean[i] + dx[i]);

    /* test_datapoint.perturb(): height first (Point.perturb :614-622), then the errors (DataPoint.perturb :561-573);
     * a time-domain datapoint perturbs its loops after the errors (TdemDataPoint.perturb :681-683 -> Loop_pair.perturb
     * Loop_pair.py:161-164 -> EmLoop.perturb -> Point.perturb on the transmitter: solve_transmitter_z) */
    if (o->solve_height && !c->sv.tdem)
        tdp.z = propose_height(&c->rng, tdp.z, o->height_prop_var, c->z_ref - o->max_height_change, c->z_ref + o->max_height_change);
    if (c->sv.n_sys == 1) {
        if (o->solve_relative_error) tdp.rel[0] = propose_error(&c->rng, tdp.rel[0], o->rel_prop_var, o->rel_min, o->rel_max);
        if (o->solve_additive_error) tdp.add[0] = propose_error(&c->rng, tdp.add[0], o->add_prop_var, o->add_min, o->add_max);
    } else {
        const double rv[2] = {o_rel_var(o, 0), o_rel_var(o, 1)}, rmn[2] = {o_rel_min(o, 0), o_rel_min(o, 1)}, rmx[2] = {o_rel_max(o, 0), o_rel_max(o, 1)};
        const double av[2] = {o_add_var(o, 0), o_add_var(o, 1)}, amn[2] = {o_add_min(o, 0), o_add_min(o, 1)}, amx[2] = {o_add_max(o, 0), o_add_max(o, 1)};
        if (o->solve_relative_error) propose_error2(&c->rng, tdp.rel, rv, rmn, rmx);
        if (o->solve_additive_error) {
            if (c->sv.tdem && c->sv.tdem->tempest) {
                /* Tempest_datapoint.perturb :339-341: additive_error_multiplier.perturb() with the defaults - no prior imposed,
                 * and the proposal's mean is never moved (DataPoint.perturb :561-573 moves only those of the two errors): every
                 * step draws exp(N(ln initial, var)) around the INITIAL multiplier */
                double z0, z1;
                rng_normal2(&c->rng, &z0, &z1);
                tdp.add[0] = exp(log(o_add_init(o, 0)) + sqrt(av[0]) * z0);
                tdp.add[1] = exp(log(o_add_init(o, 1)) + sqrt(av[1]) * z1);
            } else propose_error2(&c->rng, tdp.add, av, amn, amx);
        }
    }
    if (o->solve_height && c->sv.tdem)
        tdp.z = propose_height(&c->rng, tdp.z, o->height_prop_var, c->z_ref - o->max_height_change, c->z_ref + o->max_height_change);

    forward(c, &test, tdp.z, tdp.pred);
    data_variance(&c->sv, c->data, tdp.rel, tdp.add, var);
    double t_misfit = data_misfit(C, c->data, tdp.pred, var);
    double t_prior = datapoint_probability(o, &c->sv, tdp.rel, tdp.add, tdp.z - c->z_ref);
    if (t_prior == -INFINITY) return 0;
    t_prior += model_probability(o, &test, c->sigma_ref);
    if (t_prior == -INFINITY) return 0;
    double t_like = data_likelihood(C, c->data, tdp.pred, var);

    double proposal = 1.0, proposal1 = 1.0;
    if (action == ACT_BIRTH || action == ACT_DEATH) {
        sensitivity(c, &test, &tdp);
        double g2[GBO_MAXL], s2[GBO_MAXL], xr[GBO_MAXL], xf[GBO_MAXL];
        hessian_gradient(o, &test, c->sigma_ref, C, c->data, var, tdp.J, tdp.pred, NULL, g2);
        solve_L(k, A, g2, y);
        solve_LT(k, A, y, s2); /* H dfk */
        int bad = 0;
        double logdetL = 0.0;
        for (int i = 0; i < k; ++i) {
            /* log_values = ln(sigma') - alpha*pk with pk = -H dfk  (Model.py:626-628) */
            double lv = log(test.sigma[i]) + o->covariance_scaling * s2[i];
            /* mean = expReal(log_values) is inf above 11356 and underflows to 0 below the long-double
             * denormal limit (base/utilities.py:827-856, Model.py:630-633) */
            if (lv > 11356.0 || lv < -11399.0) bad = 1;
            xr[i] = log(remap.sigma[i]) - lv;
            xf[i] = log(test.sigma[i]) - log(remap.sigma[i]);
            logdetL += log(A[i * k + i]);
        }
        if (bad) {
            proposal = -INFINITY;
            proposal1 = -INFINITY;
        } else {
            proposal = -(0.5 * k) * LOG2PI + logdetL - 0.5 * quad_L(k, A, xr);
            proposal1 = -(0.5 * k) * LOG2PI + logdetL - 0.5 * quad_L(k, A, xf);
        }
    }
    double log_alpha = (t_prior - c->prior) + (t_like - c->likelihood) + (proposal - proposal1);
    double u = rng_uniform(&c->rng);
    c->accepted = exp(log_alpha) > u; /* NaN compares false */
    if (c->accepted) {
        c->misfit = t_misfit;
        c->prior = t_prior;
        c->likelihood = t_like;
        c->posterior = t_prior + t_like;
        c->model = test;
        c->dp = tdp;
        c->n_accept++;
    }
    return 0;
}

/* Inference1D.update.  Returns 1 if a reset was triggered. */
static int chain_update(chain_t *c, gbo_chain_out *out)
{
    const gbo_options *o = c->o;
    const int64_t N2 = 2 * (int64_t)o->n_markov_chains;
    int do_reset = 0;
    c->iteration++;
    if (c->iteration - 1 < N2) out->misfit_trace[c->iteration - 1] = c->misfit;
    if (!c->burned_in) {
        if (c->iteration > o->burn_in_min_iter && c->misfit < (double)c->n_active) {
            c->burned_in = 1;
            c->burned_in_iter = c->iteration;
            c->best_iter = c->iteration;
            c->best_model = c->model;
            memcpy(c->best_rel, c->dp.rel, sizeof(c->best_rel));
            memcpy(c->best_add, c->dp.add, sizeof(c->best_add));
            c->best_z = c->dp.z;
            c->best_posterior = c->posterior;
            reset_posteriors(o, &c->G, out);
        }
    }
    if (c->posterior > c->best_posterior) {
        c->best_iter = c->iteration;
        c->best_model = c->model;
        memcpy(c->best_rel, c->dp.rel, sizeof(c->best_rel));
        memcpy(c->best_add, c->dp.add, sizeof(c->best_add));
        c->best_z = c->dp.z;
        c->best_posterior = c->posterior;
    }
    if (c->iteration < N2) out->accept_trace[c->iteration] = (uint8_t)c->accepted;
    if (c->iteration % o->update_plot_every == 0) {
        /* acceptance_percent over acceptance_v[it-upe : it] (Inference1D.py:125-131) */
        int64_t lo = c->iteration > o->update_plot_every ? c->iteration - o->update_plot_every : 0;
        int64_t s = 0;
        for (int64_t i = lo; i < c->iteration && i < N2; ++i) s += out->accept_trace[i];
        if (o->update_plot_every > 1) {
            if (!c->burned_in) {
                if (s == 0) {
                    c->n_zero_acc++;
                    if (c->n_zero_acc == o->reset_limit) {
                        do_reset = 1;
                        c->n_zero_acc = 0;
                    }
                } else c->n_zero_acc = 0;
            } else if (s == 0) c->limiters = 0;
        }
    }
    if (do_reset) return 1; /* reset() re-initialises everything, then update continues below on the new state */
    accumulate_posteriors(o, &c->G, &c->model, c->dp.rel, c->dp.add, c->dp.z - c->z_ref, out);
    return 0;
}

static int run_chain_impl(const survey_t *sv, const gbo_options *opt, const double *data, double altitude,
                          uint64_t seed, uint64_t sounding_index, int64_t max_iterations, gbo_chain_out *out)
{
    chain_t *c = (chain_t *)calloc(1, sizeof(chain_t));
    if (!c) return -1;
    c->sv = *sv;
    c->o = opt;
    c->C = sv->C;
    memcpy(c->data, data, sizeof(double) * c->C);
    c->altitude = altitude;
    c->dp.z = altitude;
    c->rng.seed = seed;
    c->rng.sounding = sounding_index;
    c->rng.block = 0;
    c->rng.iteration = 0;
    c->n_active = 0;
    for (int i = 0; i < c->C; ++i) c->n_active += is_active(data[i]);
    chain_init(c, out);

    int failed = (c->n_active == 0);
    int go = !failed;
    int64_t total = 0;
    const int64_t N = opt->n_markov_chains;
    while (go) {
        failed = chain_step(c);
        int reset = chain_update(c, out);
        total++;
        if (reset) {
            /* Inference1D.reset(): re-initialise, continue the random stream */
            c->n_resets++;
            chain_init(c, out);
            accumulate_posteriors(opt, &c->G, &c->model, c->dp.rel, c->dp.add, c->dp.z - c->z_ref, out);
        }
        go = !failed && (c->iteration <= N + c->burned_in_iter);
        if (!failed && !c->burned_in) {
            go = c->iteration < N;
            if (!go) failed = 1;
        }
        if (c->n_resets == 3 && !c->burned_in) {
            if (!c->limiters) {
                c->limiters = 1;
                c->n_resets = 1;
                chain_init(c, out);
            } else {
                go = 0;
                failed = 1;
            }
        }
        if (max_iterations > 0 && total >= max_iterations) go = 0;
    }

    double *s = out->scalars;
    memset(s, 0, sizeof(double) * GBO_NSCALARS);
    s[GBO_S_ITER] = (double)c->iteration;
    s[GBO_S_BURNED_IN] = c->burned_in;
    s[GBO_S_BURNED_IN_ITER] = (double)c->burned_in_iter;
    s[GBO_S_BEST_ITER] = (double)c->best_iter;
    s[GBO_S_BEST_K] = c->best_model.k;
    s[GBO_S_CUR_K] = c->model.k;
    s[GBO_S_HALFSPACE] = c->sigma_ref;
    s[GBO_S_FAILED] = failed;
    s[GBO_S_N_ACCEPT] = (double)c->n_accept;
    s[GBO_S_N_FORWARD] = (double)c->n_forward;
    s[GBO_S_N_SENS] = (double)c->n_sens;
    s[GBO_S_BEST_POSTERIOR] = c->best_posterior;
    s[GBO_S_CUR_REL] = c->dp.rel[0];
    s[GBO_S_CUR_ADD] = c->dp.add[0];
    s[GBO_S_CUR_REL2] = c->dp.rel[1];
    s[GBO_S_CUR_ADD2] = c->dp.add[1];
    s[GBO_S_BEST_REL2] = c->best_rel[1];
    s[GBO_S_BEST_ADD2] = c->best_add[1];
    s[GBO_S_CUR_MISFIT] = c->misfit;
    s[GBO_S_CUR_PRIOR] = c->prior;
    s[GBO_S_CUR_LIKELIHOOD] = c->likelihood;
    s[GBO_S_BEST_REL] = c->best_rel[0];
    s[GBO_S_BEST_ADD] = c->best_add[0];
    s[GBO_S_N_RESETS] = c->n_resets;
    s[GBO_S_N_BIRTH] = (double)c->n_act[0];
    s[GBO_S_N_DEATH] = (double)c->n_act[1];
    s[GBO_S_N_MOVE] = (double)c->n_act[2];
    s[GBO_S_N_NONE] = (double)c->n_act[3];
    s[GBO_S_TOTAL_ITER] = (double)total;
    s[GBO_S_CUR_HEIGHT] = c->dp.z;
    s[GBO_S_BEST_HEIGHT] = c->best_z;
    s[GBO_S_HEIGHT_REF] = c->z_ref;
    for (int i = 0; i < opt->max_layers; ++i) {
        out->best_sigma[i] = i < c->best_model.k ? c->best_model.sigma[i] : NAN;
        out->cur_sigma[i] = i < c->model.k ? c->model.sigma[i] : NAN;
    }
    for (int i = 0; i <= opt->max_layers; ++i) {
        out->best_edges[i] = i <= c->best_model.k ? c->best_model.edges[i] : NAN;
        out->cur_edges[i] = i <= c->model.k ? c->model.edges[i] : NAN;
    }
    free(c);
    return 0;
}

int gbo_run_chain(const gbo_fdem_system *sys, const gbo_options *opt, const double *data, double altitude,
                  uint64_t seed, uint64_t sounding_index, int64_t max_iterations, gbo_chain_out *out)
{
    survey_t sv;
    survey_fdem(&sv, sys);
    return run_chain_impl(&sv, opt, data, altitude, seed, sounding_index, max_iterations, out);
}

int gbo_run_chain_tdem(const gbo_tdem_system *sys, const gbo_options *opt, const double *data, double altitude,
                       uint64_t seed, uint64_t sounding_index, int64_t max_iterations, gbo_chain_out *out)
{
    survey_t sv;
    survey_tdem(&sv, sys);
    if ((opt->n_systems > 1 ? opt->n_systems : 1) != sv.n_sys) return -2;
    /* solve_height of a time-domain datapoint = the options file's solve_transmitter_z (the transmitter loop's height,
     * receiver offset fixed: Loop_pair.Geometry Loop_pair.py:62-78); the other loop-geometry unknowns are not restated */
    return run_chain_impl(&sv, opt, data, altitude, seed, sounding_index, max_iterations, out);
}

/* ------------------------------------------------------------------ term-level pin */
static void sv_forward(const survey_t *sv, double alt, int k, const double *sig, const double *thk, double *pred)
{
    if (sv->tdem) {
        gbo_tdem_forward(sv->tdem, alt, k, sig, thk, pred);
        if (sv->tdem->tempest)
            for (int i = 0; i < sv->C; ++i) pred[i] += sv->tdem->primary[i];
    } else gbo_fdem_forward(sv->fdem, alt, k, sig, thk, pred);
}
static void sv_sensitivity(const survey_t *sv, double alt, int k, const double *sig, const double *thk, double *J)
{
    if (sv->tdem) gbo_tdem_sensitivity(sv->tdem, alt, k, sig, thk, J);
    else gbo_fdem_sensitivity(sv->fdem, alt, k, sig, thk, J);
}

static int eval_transition_impl(const survey_t *sv, const gbo_options *o, gbo_transition *t);
int gbo_eval_transition(const gbo_fdem_system *sys, const gbo_options *o, gbo_transition *t)
{
    survey_t sv;
    survey_fdem(&sv, sys);
    return eval_transition_impl(&sv, o, t);
}
int gbo_eval_transition_tdem(const gbo_tdem_system *sys, const gbo_options *o, gbo_transition *t)
{
    survey_t sv;
    survey_tdem(&sv, sys);
    return eval_transition_impl(&sv, o, t);
}

static int eval_transition_impl(const survey_t *sv, const gbo_options *o, gbo_transition *t)
{
    const int k = t->k, C = sv->C;
    model_t remap, test;
    remap.k = test.k = k;
    memcpy(remap.edges, t->edges, sizeof(double) * (k + 1));
    memcpy(test.edges, t->edges, sizeof(double) * (k + 1));
    memcpy(remap.sigma, t->sigma_remap, sizeof(double) * k);
    memcpy(test.sigma, t->sigma_test, sizeof(double) * k);
    double thk[GBO_MAXL], var[GBO_MAXC], J[GBO_MAXC * GBO_MAXL], pred[GBO_MAXC];
    double A[GBO_MAXL * GBO_MAXL], y[GBO_MAXL], step[GBO_MAXL];
    model_thickness(&remap, thk);
    if (t->action != ACT_NONE) {
        if (sv->tdem) memcpy(pred, t->pred_in, sizeof(double) * C); /* TdemDataPoint.fm_dlogc quirk, see chain_step */
        else sv_forward(sv, t->altitude, k, remap.sigma, thk, pred);
        sv_sensitivity(sv, t->altitude, k, remap.sigma, thk, J);
    } else {
        memcpy(pred, t->pred_in, sizeof(double) * C);
        memcpy(J, t->J_in, sizeof(double) * C * k);
    }
    data_variance(sv, t->data, t->rel_cur, t->add_cur, var);
    hessian_gradient(o, &remap, t->sigma_ref, C, t->data, var, J, pred, A, t->gradient);
    memcpy(t->hessian, A, sizeof(double) * k * k);
    if (cholesky(k, A)) return 1;
    solve_L(k, A, t->gradient, y);
    solve_LT(k, A, y, step);
    for (int i = 0; i < k; ++i) t->newton_mean[i] = exp(log(remap.sigma[i]) - o->covariance_scaling * step[i]);

    const double z_test = o->solve_height ? t->altitude_test : t->altitude;
    sv_forward(sv, z_test, k, test.sigma, thk, t->pred_test);
    data_variance(sv, t->data, t->rel_test, t->add_test, var);
    t->misfit_test = data_misfit(C, t->data, t->pred_test, var);
    t->prior_test = datapoint_probability(o, sv, t->rel_test, t->add_test, o->solve_height ? z_test - t->altitude_ref : 0.0) +
                    model_probability(o, &test, t->sigma_ref);
    t->likelihood_test = data_likelihood(C, t->data, t->pred_test, var);
    t->proposal = 1.0;
    t->proposal1 = 1.0;
    if (t->action == ACT_BIRTH || t->action == ACT_DEATH) {
        double g2[GBO_MAXL], s2[GBO_MAXL], xr[GBO_MAXL], xf[GBO_MAXL], logdetL = 0.0;
        sv_sensitivity(sv, z_test, k, test.sigma, thk, J);
        hessian_gradient(o, &test, t->sigma_ref, C, t->data, var, J, t->pred_test, NULL, g2);
        solve_L(k, A, g2, y);
        solve_LT(k, A, y, s2);
        int bad = 0;
        for (int i = 0; i < k; ++i) {
            double lv = log(test.sigma[i]) + o->covariance_scaling * s2[i];
            if (lv > 11356.0 || lv < -11399.0) bad = 1; /* Model.py:630-633, as in chain_step */
            xr[i] = log(remap.sigma[i]) - lv;
            xf[i] = log(test.sigma[i]) - log(remap.sigma[i]);
            logdetL += log(A[i * k + i]);
        }
        t->proposal = bad ? -INFINITY : -(0.5 * k) * LOG2PI + logdetL - 0.5 * quad_L(k, A, xr);
        t->proposal1 = bad ? -INFINITY : -(0.5 * k) * LOG2PI + logdetL - 0.5 * quad_L(k, A, xf);
    }
    return 0;
}
