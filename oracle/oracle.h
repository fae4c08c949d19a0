/* TEST INFRASTRUCTURE ONLY (oracle) - see fdem1d_oracle.c / rjmcmc_oracle.c headers. */
#ifndef GBO_ORACLE_H
#define GBO_ORACLE_H

#include <stdint.h>

#define GBO_MAXL 64  /* max layers the oracle accepts */
#define GBO_MAXF 16  /* max frequencies per system   */
#define GBO_MAXSYS 2      /* systems per datapoint (SkyTEM dual moment) */
#define GBO_TD_NFREQ 32   /* spline nodes of the TDEM frequency-domain response */
#define GBO_TD_MAXLAM 32  /* Hankel abscissae */
#define GBO_MAXC 64       /* max data channels (FDEM: 2 x frequencies; TDEM: windows of all systems) */

/* One FDEM acquisition system (one row per frequency of an .stm file,
 * geobipy/src/classes/system/FdemSystem.py:146-183).
 * tid = 1 + 3*rx_orientation + tx_orientation with x=0,y=1,z=2 (FdemSystem.py:199-203). */
typedef struct {
    int32_t n_freq;
    int32_t tid[GBO_MAXF];
    double freq[GBO_MAXF];
    double tmom[GBO_MAXF], tx[GBO_MAXF], ty[GBO_MAXF], tz[GBO_MAXF];
    double rmom[GBO_MAXF], rx[GBO_MAXF], ry[GBO_MAXF], rz[GBO_MAXF];
} gbo_fdem_system;

/* Sampler options (documentation_source/.../options_files/resolve_options,
 * geobipy/src/inversion/user_parameters.py:40-44, Inference1D.py:78-96). */
typedef struct {
    int32_t n_markov_chains;
    int32_t update_plot_every;
    int32_t max_layers;
    int32_t solve_parameter, solve_gradient, solve_relative_error, solve_additive_error;
    int32_t reset_limit;
    double min_edge, max_edge, min_width;
    double p_birth, p_death, p_move, p_none;
    double factor;                 /* value prior std = ln(1 + factor) */
    double gradient_std;
    double covariance_scaling;     /* alpha of the stochastic-Newton step */
    double rel_init, rel_min, rel_max, rel_prop_var;
    double add_init, add_min, add_max, add_prop_var;
    int32_t n_sigma_bins;          /* 250 */
    int32_t n_err_bins;            /* 99 */
    double sigma_bins_nstd;        /* 4.0 */
    int32_t burn_in_min_iter;      /* 5000 (Inference1D.py:726) */
    int32_t n_systems;             /* 0/1: one system; 2: the *_2 fields below describe system 1 (skytem_options lists) */
    double rel_init2, rel_min2, rel_max2, rel_prop_var2;
    double add_init2, add_min2, add_max2, add_prop_var2;
    /* solve_z (Point.set_priors :959-961, set_proposals :977-979): sensor height sampled with a Uniform prior
     * [z0 - max_height_change, z0 + max_height_change] and a Normal(z, height_prop_var) random walk.  For a time-domain
     * datapoint this is the options file's solve_transmitter_z (tempest_options :104-108): the transmitter loop's
     * height, receiver offset fixed (Loop_pair.set_priors Loop_pair.py:166-178, Loop_pair.Geometry :62-78), drawn
     * after the error proposals (TdemDataPoint.perturb :681-683) */
    int32_t solve_height;
    int32_t pad_h_;
    double max_height_change, height_prop_var;
} gbo_options;

/* One time-domain datapoint type = n_sys GA-AEM style systems sharing one transmitter/receiver geometry
 * (TdemDataPoint with system=[SkytemHM.stm, SkytemLM.stm]).  Everything model independent is prebuilt by
 * the caller (oracle_py.make_tdem_system): see tdem1d_oracle.c. */
typedef struct {
    int32_t n_sys, n_freq, n_lam, C;
    int32_t n_win[GBO_MAXSYS];
    int32_t pad_[2];
    double freq[GBO_TD_NFREQ];        /* spline nodes [Hz], log spaced, shared by the systems */
    double xi[GBO_TD_MAXLAM];         /* ln(lambda * ZH / 2), uniform */
    double rx_dx, rx_dy, rx_dz;       /* receiver offset from the transmitter [m] (z up) */
    double loop_radius;               /* ModellingLoopRadius */
    double MR[GBO_MAXC * GBO_TD_NFREQ], MI[GBO_MAXC * GBO_TD_NFREQ]; /* [C][n_freq] window operator */
    double t_centre[GBO_MAXC];        /* window centre times = off_time (TdemSystem_GAAEM.py:33) */
    int32_t comp[GBO_MAXC];           /* component of channel c: 0 = z, 1 = x (fixed-wing systems: Tempest measures X and Z;
                                         TdemDataPoint.forward :1008-1016 stacks SX then -SZ) */
    double rx_cx;                     /* dx / r: direction cosine of the receiver offset (x component of the horizontal field) */
    /* Tempest_datapoint (classes/data/datapoint/Tempest_datapoint.py): data = secondary + primary field per component
     * (:107-127); std_c = sqrt((rel[component] data_c)^2 + (multiplier[component] additive_c)^2) with a FIXED additive level
     * per channel (:141-176); the unknowns are the relative errors and the multipliers, one per component (:478-510) */
    int32_t tempest, pad_t;
    double add_level[GBO_MAXC];       /* additive error of channel c (the options file's initial_additive_error vector) */
    double primary[GBO_MAXC];         /* predicted primary field of channel c's component, added to the forward response */
} gbo_tdem_system;

/* Everything one chain produces (caller allocates; sizes from gbo_sizes()). */
typedef struct {
    int32_t *hitmap;        /* [n_sigma_bins][n_depth] */
    int32_t *edges_hist;    /* [n_depth] */
    int32_t *ncells_hist;   /* [max_layers + 1] */
    int32_t *rel_hist;      /* [n_systems][n_err_bins] */
    int32_t *add_hist;      /* [n_systems][n_err_bins] */
    double *misfit_trace;   /* [2 * n_markov_chains] */
    uint8_t *accept_trace;  /* [2 * n_markov_chains] */
    double *best_sigma;     /* [max_layers] */
    double *best_edges;     /* [max_layers + 1] */
    double *cur_sigma;      /* [max_layers] */
    double *cur_edges;      /* [max_layers + 1] */
    double *scalars;        /* [GBO_NSCALARS], see below */
    int32_t *height_hist;   /* [n_err_bins] (Point.set_z_posterior :1013-1020), may be NULL unless solve_height */
} gbo_chain_out;

enum {
    GBO_S_ITER = 0, GBO_S_BURNED_IN, GBO_S_BURNED_IN_ITER, GBO_S_BEST_ITER, GBO_S_BEST_K, GBO_S_CUR_K,
    GBO_S_HALFSPACE, GBO_S_FAILED, GBO_S_N_ACCEPT, GBO_S_N_FORWARD, GBO_S_N_SENS, GBO_S_BEST_POSTERIOR,
    GBO_S_CUR_REL, GBO_S_CUR_ADD, GBO_S_CUR_MISFIT, GBO_S_CUR_PRIOR, GBO_S_CUR_LIKELIHOOD,
    GBO_S_BEST_REL, GBO_S_BEST_ADD, GBO_S_N_RESETS, GBO_S_N_BIRTH, GBO_S_N_DEATH, GBO_S_N_MOVE, GBO_S_N_NONE, GBO_S_TOTAL_ITER,
    GBO_S_CUR_REL2, GBO_S_CUR_ADD2, GBO_S_BEST_REL2, GBO_S_BEST_ADD2,   /* system 1 of a dual-moment datapoint */
    GBO_S_CUR_HEIGHT, GBO_S_BEST_HEIGHT,                                /* sensor height (solve_height) */
    GBO_S_HEIGHT_REF,                                                   /* centre of the height prior (re-centred by reset()) */
    GBO_NSCALARS = 32
};

int gbo_n_depth(const gbo_options *o);

void gbo_fdem_geometry(const gbo_fdem_system *sys, double altitude, double *tHeight, double *rHeight,
                       double *scale, double *xsep, double *sep);
int gbo_fdem_forward(const gbo_fdem_system *sys, double altitude, int L, const double *sigma,
                     const double *thickness, double *out);
int gbo_fdem_sensitivity(const gbo_fdem_system *sys, double altitude, int L, const double *sigma,
                         const double *thickness, double *J);

int gbo_tdem_forward(const gbo_tdem_system *sys, double altitude, int L, const double *sigma,
                     const double *thickness, double *out);
int gbo_tdem_sensitivity(const gbo_tdem_system *sys, double altitude, int L, const double *sigma,
                         const double *thickness, double *J);
int gbo_tdem_frequency_response(const gbo_tdem_system *sys, double altitude, int L, const double *sigma,
                                const double *thickness, double *S_re, double *S_im);

/* Philox4x32-10 block (Salmon et al. 2011). */
void gbo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Deterministic term evaluation for one proposed transition (used to pin the restatement
 * against records captured from the live reference).  See rjmcmc_oracle.c. */
typedef struct {
    /* inputs */
    int32_t k;                    /* layers of remapped/test model */
    int32_t action;               /* 0 birth 1 death 2 move 3 none */
    double altitude;
    double sigma_ref;             /* half-space conductivity (value prior mean) */
    double edges[GBO_MAXL + 1];   /* edges of remapped == test model, edges[k] = inf */
    double sigma_remap[GBO_MAXL];
    double sigma_test[GBO_MAXL];
    double rel_cur[GBO_MAXSYS], add_cur[GBO_MAXSYS];    /* errors of the current datapoint (used for the Hessian) */
    double rel_test[GBO_MAXSYS], add_test[GBO_MAXSYS];  /* errors of the proposed datapoint */
    double data[GBO_MAXC];        /* observed */
    double J_in[GBO_MAXC * GBO_MAXL];   /* stored Jacobian (row-major [C][k]) used when action == none */
    double pred_in[GBO_MAXC];           /* stored predicted data used when action == none */
    /* outputs */
    double hessian[GBO_MAXL * GBO_MAXL];  /* precision A = Wm'Wm + J'Wd'WdJ, row-major [k][k] */
    double gradient[GBO_MAXL];
    double newton_mean[GBO_MAXL];         /* exp(ln sigma_remap - alpha * A^-1 gradient) */
    double pred_test[GBO_MAXC];
    double misfit_test, prior_test, likelihood_test, proposal, proposal1;
    /* input, read only when opt->solve_height: height of the proposed datapoint (altitude = the current one's),
     * and the centre of the height prior */
    double altitude_test, altitude_ref;
} gbo_transition;

int gbo_eval_transition(const gbo_fdem_system *sys, const gbo_options *opt, gbo_transition *t);
int gbo_eval_transition_tdem(const gbo_tdem_system *sys, const gbo_options *opt, gbo_transition *t);

/* Full chain for one sounding.  Random stream: Philox4x32-10, key = (seed lo, seed hi),
 * counter = (block, iteration, sounding lo, sounding hi): one sub-stream per accept_reject() call. */
int gbo_run_chain(const gbo_fdem_system *sys, const gbo_options *opt, const double *data, double altitude,
                  uint64_t seed, uint64_t sounding_index, int64_t max_iterations, gbo_chain_out *out);
/* Same sampler for a time-domain (dual moment) datapoint: data [C] = windows of system 0 then system 1. */
int gbo_run_chain_tdem(const gbo_tdem_system *sys, const gbo_options *opt, const double *data, double altitude,
                       uint64_t seed, uint64_t sounding_index, int64_t max_iterations, gbo_chain_out *out);

#endif
