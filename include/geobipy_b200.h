/* geobipy_b200 - C-ABI of the B200-native per-sounding rjMCMC EM inversion path.
 *
 * The reference (DOI-USGS/geobipy) is pure Python; it has no FFI seam, its seams are the
 * Python call signatures listed in SURVEY.md section 8(b).  This header is the drop-in
 * boundary those seams bind to (ctypes stub: INTEGRATION.md):
 *
 *   gbp_fdem_forward*      replaces nbFdem1dfwd / fdem1dfwd / FdemDataPoint.forward
 *                          (geobipy/src/classes/forwardmodelling/Electromagnetic/FD/fdem1d_numba.py:25,
 *                           fdem1d.py:10, classes/data/datapoint/FdemDataPoint.py:524-545)
 *   gbp_fdem_sensitivity*  replaces nbFdem1dsen / fdem1dsen / FdemDataPoint.sensitivity
 *                          (fdem1d_numba.py:72, fdem1d.py:87, FdemDataPoint.py:548-557)
 *   gbp_rjmcmc_run*        replaces Inference1D.initialize + Inference1D.infer for a batch of
 *                          soundings (geobipy/src/inversion/Inference1D.py:353, :633), i.e. the
 *                          loop Inference3D.infer_serial / _infer_mpi_worker_task drive
 *                          (geobipy/src/inversion/Inference3D.py:458-492, :587-635).
 *
 *   gbp_tdem_forward*      replaces TdemDataPoint.forward and the external gatdaem1d forward model it calls
 *                          (classes/data/datapoint/TdemDataPoint.py:997-1022,
 *                           classes/forwardmodelling/Electromagnetic/TD/tdem1d.py:89-96)
 *   gbp_tdem_sensitivity*  replaces TdemDataPoint.sensitivity / fm_dlogc (TdemDataPoint.py:1024-1055, tdem1d.py:98-154)
 *   gbp_tdem_rjmcmc_run*   the same sampler for time-domain (single or dual moment) datapoints
 *
 * Plain pointers and sizes only.  "_host" entry points take HOST buffers and do the
 * host<->device copies themselves; the others take DEVICE pointers and a cudaStream_t
 * (passed as void*) and are stream ordered.  All functions return 0 on success, non-zero
 * on error (message: gbp_last_error()).  There is no CPU fallback: without a CUDA device
 * every compute entry point fails.
 *
 * Threading contract.  Every entry point may be called concurrently from several host threads, on several devices
 * and on several streams of one device: per-call scratch (the work counter and the per-chain Jacobian mirror of the
 * samplers) is allocated stream-ordered per call, launch sequences are serialised per device by a mutex (they are
 * asynchronous and short), the "_host" samplers serialise per device on their cached device buffers, and
 * gbp_last_error() is thread local.  The kernel timings (gbp_last_kernel_ms / gbp_kernel_ms_stats) and
 * gbp_debug_counters are per DEVICE, not per stream: with launches in flight on several streams of one device they
 * describe whichever launches came last.  The device entry points use the CURRENT device of the calling thread, which
 * must be the device the pointers and the stream belong to.
 */
#ifndef GEOBIPY_B200_H
#define GEOBIPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBP_MAXF 16              /* max frequencies of one FDEM system */
#define GBP_MAXC (2 * GBP_MAXF)  /* max data channels (in-phase + quadrature) */
#define GBP_MAXL 30              /* max layers (resolve_options: maximum_number_of_layers) */

#define GBP_TD_MAXSYS 2          /* systems of one time-domain datapoint (SkyTEM: high + low moment) */
#define GBP_TD_NFREQ 32          /* spline nodes of the frequency-domain response (one per lane) */
#define GBP_TD_MAXLAM 32         /* Hankel abscissae */
#define GBP_TD_MAXWIN 32         /* receiver windows of one system */
#define GBP_TD_MAXC 64           /* data channels of one time-domain datapoint (forward / Jacobian operators) */
#define GBP_TD_SAMPLER_MAXC 48   /* data channels the time-domain SAMPLER holds per chain (gbp_tdem_rjmcmc_run*
                                    reject datapoint types with more; SkyTEM dual moment has 45) */
#define GBP_TD_MAXWAVE 64        /* vertices of the current waveform */
#define GBP_TD_MAXFILT 4         /* receiver low-pass filters */

#define GBP_PRECISION_F32 32     /* forward / Jacobian arithmetic in fp32 (fast path) */
#define GBP_PRECISION_F64 64     /* forward / Jacobian arithmetic in fp64 (validation path) */

/* One FDEM acquisition system = the rows of an .stm file (FdemSystem.read, FdemSystem.py:146-183).
 * tid = 1 + 3*rx_orientation + tx_orientation, x=0 y=1 z=2 (FdemSystem.py:199-203): zz=9, xx=1, 3, 7. */
typedef struct {
    int32_t n_freq;
    int32_t tid[GBP_MAXF];
    double freq[GBP_MAXF];
    double tmom[GBP_MAXF], tx[GBP_MAXF], ty[GBP_MAXF], tz[GBP_MAXF];
    double rmom[GBP_MAXF], rx[GBP_MAXF], ry[GBP_MAXF], rz[GBP_MAXF];
} gbp_fdem_system;

/* Sampler options = the keys of a reference options file (resolve_options) with the defaults of
 * user_parameters.py:40-44 and Inference1D.__init__ (Inference1D.py:78-96). */
typedef struct {
    int32_t n_markov_chains;
    int32_t update_plot_every;
    int32_t max_layers;                 /* maximum_number_of_layers, <= GBP_MAXL */
    int32_t solve_parameter, solve_gradient, solve_relative_error, solve_additive_error;
    int32_t reset_limit;
    double min_edge, max_edge, min_width;  /* minimum_depth, maximum_depth, minimum_thickness */
    double p_birth, p_death, p_move, p_none;
    double factor;                      /* value prior std = ln(1 + factor) */
    double gradient_std;                /* gradient_standard_deviation */
    double covariance_scaling;
    double rel_init, rel_min, rel_max, rel_prop_var;
    double add_init, add_min, add_max, add_prop_var;
    int32_t n_sigma_bins;               /* 250 (Model.set_posteriors, Model.py:675) */
    int32_t n_err_bins;                 /* 99  (Uniform.bins default) */
    double sigma_bins_nstd;             /* 4.0 */
    int32_t burn_in_min_iter;           /* 5000 (Inference1D.py:726) */
    int32_t n_systems;                  /* 0 or 1: one system.  2: dual-moment datapoint, the *2 fields below are the
                                           second entries of the options file's lists (skytem_options) */
    double rel_init2, rel_min2, rel_max2, rel_prop_var2;
    double add_init2, add_min2, add_max2, add_prop_var2;
    /* Sensor height as an unknown (the options file's solve_z / maximum_z_change / z_proposal_variance:
     * Point.set_priors pointcloud/Point.py:959-961, set_proposals :977-979, perturb :614-622).  Frequency-domain
     * datapoints with <= 6 frequencies.  For the time-domain entry points it is the TRANSMITTER height
     * (solve_transmitter_z / maximum_transmitter_z_change / transmitter_z_proposal_variance: TdemDataPoint.perturb
     * :681-683, prior after the error priors :950-951; the receiver offset stays fixed). */
    int32_t solve_height;
    int32_t pad_h_;
    double max_height_change;           /* prior Uniform[z0 - max_height_change, z0 + max_height_change] */
    double height_prop_var;             /* variance of the Normal random-walk proposal */
} gbp_options;

/* One time-domain acquisition system = the contents of a GA-AEM .stm file as gatdaem1d's TDAEMSystem reads
 * it (classes/system/TdemSystem_GAAEM.py:26-40; e.g. documentation_source/.../data/SkytemHM.stm, tempest.stm).
 * Vertical-axis transmitter (finite loop or point dipole), zero attitudes, no secondary-field normalisation; receiver
 * components Z and / or X; OutputType dB/dt (the receiver voltage, -dB/dt) or B.  Channel order of a system, as
 * TdemDataPoint.forward stacks them (TdemDataPoint.py:1008-1016): the X windows (if measured), then the Z windows. */
typedef struct {
    int32_t n_wave, n_windows, n_filters;
    int32_t n_abscissae;                /* NumberOfAbsiccaInHankelTransformEvaluation */
    double base_frequency;              /* BaseFrequency [Hz] */
    double digitising_frequency;        /* WaveformDigitisingFrequency [Hz] (harmonics up to its Nyquist limit) */
    double loop_radius;                 /* ModellingLoopRadius [m], 0 = point dipole */
    double wave_time[GBP_TD_MAXWAVE], wave_current[GBP_TD_MAXWAVE];  /* WaveFormCurrent: half a period */
    double window_start[GBP_TD_MAXWIN], window_end[GBP_TD_MAXWIN];   /* WindowTimes [s] */
    double filter_cutoff[GBP_TD_MAXFILT];                            /* LowPassFilter CutOffFrequency [Hz] */
    int32_t filter_order[GBP_TD_MAXFILT];                            /* LowPassFilter Order */
    /* (appended in round 2; an all-zero tail means the SkyTEM-type defaults: dB/dt, 1 A, Z only with scaling 1) */
    int32_t output_type;                /* OutputType: 0 = dB/dt, 1 = B */
    int32_t pad2_;
    double peak_current;                /* PeakCurrent [A]: scales the normalised waveform; 0 is read as 1 */
    double x_scaling, z_scaling;        /* X / ZOutputScaling: a component is measured where its scaling is non-zero (Tempest:
                                           1e15 = fT); (0, 0) is read as Z only, scaling 1.  The waveform may cover half a period
                                           (the second half is its negative) or the whole period. */
} gbp_tdem_system;

/* A time-domain datapoint type: its systems and the transmitter->receiver offset (Loop_pair.Geometry,
 * classes/system/Loop_pair.py:62-78; z up, zero pitch/roll/yaw).
 * error_model (sampler only): 0 = TdemDataPoint (TdemDataPoint.std :329-379: relative and additive error per SYSTEM, the
 * additive error scaled by (t / 1 ms)^-1/2); 1 = Tempest_datapoint (Tempest_datapoint.py:107-176, :478-510): ONE system,
 * the data handed to the sampler are secondary + primary field, the errors are per COMPONENT in channel order (x, z):
 * gbp_options.rel_* / rel_*2 the relative errors, gbp_options.add_* / add_*2 the additive-error MULTIPLIERS over the fixed
 * additive_level of every channel (the options file's initial_additive_error vector). */
typedef struct {
    int32_t n_systems, error_model;
    gbp_tdem_system sys[GBP_TD_MAXSYS];
    double rx_dx, rx_dy, rx_dz;
    double additive_level[GBP_TD_MAXC];   /* error_model 1: additive error of channel c, in the output scaling's unit */
} gbp_tdem_survey;

/* Per-chain scalar slots of gbp_chain_buffers.scalars ([B][GBP_NSCALARS] doubles). */
enum {
    GBP_S_ITER = 0, GBP_S_BURNED_IN, GBP_S_BURNED_IN_ITER, GBP_S_BEST_ITER, GBP_S_BEST_K, GBP_S_CUR_K,
    GBP_S_HALFSPACE, GBP_S_FAILED, GBP_S_N_ACCEPT, GBP_S_N_FORWARD, GBP_S_N_SENS, GBP_S_BEST_POSTERIOR,
    GBP_S_CUR_REL, GBP_S_CUR_ADD, GBP_S_CUR_MISFIT, GBP_S_CUR_PRIOR, GBP_S_CUR_LIKELIHOOD,
    GBP_S_BEST_REL, GBP_S_BEST_ADD, GBP_S_N_RESETS, GBP_S_N_BIRTH, GBP_S_N_DEATH, GBP_S_N_MOVE, GBP_S_N_NONE,
    GBP_S_TOTAL_ITER,   /* accept_reject+update pairs executed, including those before a reset() */
    GBP_S_CUR_REL2, GBP_S_CUR_ADD2, GBP_S_BEST_REL2, GBP_S_BEST_ADD2,  /* system 1 of a dual-moment datapoint */
    GBP_S_CUR_HEIGHT, GBP_S_BEST_HEIGHT,   /* datapoint.z / best_datapoint.z (the input height unless solve_height) */
    GBP_S_HEIGHT_REF,   /* centre of the height prior and of the height_hist bins: the input height, re-centred on the
                           sampled height by every reset() (Inference1D.py:984-994, Point.set_priors :959-961) */
    GBP_NSCALARS = 32
};

/* Posterior / result arrays of a batch of B chains - exactly what Inference1D.writeHdf
 * serialises per sounding (Inference1D.py:1050-1090).  Any pointer may be NULL to skip that
 * output, except scalars.  Row-major, leading dimension B. */
typedef struct {
    int32_t *hitmap;        /* [B][n_sigma_bins][n_depth]   model.values.posterior.counts   */
    int32_t *edges_hist;    /* [B][n_depth]                 model.mesh.edges.posterior      */
    int32_t *ncells_hist;   /* [B][max_layers + 1]          model.mesh.nCells.posterior     */
    int32_t *rel_hist;      /* [B][n_systems][n_err_bins]   datapoint.relative_error.posterior */
    int32_t *add_hist;      /* [B][n_systems][n_err_bins]   datapoint.additive_error.posterior */
    double *misfit_trace;   /* [B][2 * n_markov_chains]     data_misfit_v                   */
    uint8_t *accept_trace;  /* [B][2 * n_markov_chains]     acceptance_v                    */
    double *best_sigma;     /* [B][max_layers]   NaN padded best_model.values               */
    double *best_edges;     /* [B][max_layers + 1]          best_model.mesh.edges           */
    double *cur_sigma;      /* [B][max_layers]              model.values                    */
    double *cur_edges;      /* [B][max_layers + 1]          model.mesh.edges                */
    double *scalars;        /* [B][GBP_NSCALARS]                                            */
    int32_t *height_hist;   /* [B][n_err_bins]  datapoint.z.posterior (Point.set_z_posterior :1013-1020): bins of
                               z - z0 over [-max_height_change, max_height_change]; written when solve_height */
} gbp_chain_buffers;

/* ---- library ---------------------------------------------------------------------------- */
const char *gbp_version(void);
const char *gbp_last_error(void);
int gbp_device_count(void);
/* depth cells of the posterior grids: len(arange(0, 1.1*max_edge, 0.5*min_width)) - 1
 * (RectilinearMesh1D.set_posteriors, RectilinearMesh1D.py:1450) */
int gbp_n_depth(const gbp_options *opt);
/* algorithmic flop / special-function count of one forward (SURVEY.md section 8(d) convention) */
double gbp_flops_per_forward(const gbp_fdem_system *sys, int n_layers);
/* number of (frequency, abscissa) filter points one forward evaluates (RESOLVE: 860) */
int gbp_filter_points(const gbp_fdem_system *sys);
/* special-function-unit (MUFU: rcp, sqrt, ex2, sin, cos) operations of one forward, same convention */
double gbp_mufu_per_forward(const gbp_fdem_system *sys, int n_layers);
/* measured peaks of the two pipes that bound this path on the current device: dependent-free FFMA chains
 * (TFLOP/s, 2 flops per FMA) and MUFU operations (Gop/s).  Two tiny kernels, best of 3 after warm-up. */
int gbp_measure_peaks(double *fp32_tflops, double *mufu_gops);
/* diagnostics of speculative evaluation, summed over all launches on the current device since the last reset
 * (16 counters, see gbp_chain.cuh g_diag: [8] iterations committed from speculation, [9] rounds, ...).  The only
 * outputs that depend on scheduling; synchronises the device */
int gbp_debug_counters(unsigned long long *out16, int reset);
/* with the environment variable GBP_DEBUG_TIMELINE set, the sampler records when each chain of the last launch
 * on the current device finished; out_ms[i] = milliseconds after the start of the kernel (synchronises) */
int gbp_debug_finish_times(double *out_ms, int n);
/* the same run's progress stamps: out_ms[i][j] = milliseconds after the start of the kernel at which chain i had done
 * 2048 j iterations (j = 0 is its start; -1 where it never got there; j = 31 absorbs everything beyond) */
int gbp_debug_progress_times(double *out_ms, int n);
/* kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t gbp_launch_count(void);
/* mean duration [ms] and launch count of the last gbp_rjmcmc_run / forward kernel, measured with CUDA
 * events on the launching stream (valid after the stream has been synchronised) */
int gbp_last_kernel_ms(float *ms);
/* mean duration [ms] of the last `last_n` (<= 32) kernels this library launched on the current device (*counted =
 * how many were averaged); synchronises on their events.  bench.py's roofline.kernel_ms */
int gbp_kernel_ms_stats(int last_n, float *mean_ms, int *counted);

/* ---- FDEM forward / Jacobian, DEVICE pointers --------------------------------------------- */
/* sigma, thickness: [B][l_stride] (thickness of the last layer is ignored = infinite half-space);
 * out: [B][2F] (real parts then imaginary parts, ppm); J: [B][2F][l_stride] = d out / d ln(sigma). */
int gbp_fdem_forward(const gbp_fdem_system *sys, int B, int l_stride, const int32_t *d_nlayers,
                     const double *d_sigma, const double *d_thickness, const double *d_altitude,
                     double *d_out, int precision, void *stream);
int gbp_fdem_sensitivity(const gbp_fdem_system *sys, int B, int l_stride, const int32_t *d_nlayers,
                         const double *d_sigma, const double *d_thickness, const double *d_altitude,
                         double *d_out, double *d_J, int precision, void *stream);

/* ---- FDEM forward / Jacobian, HOST pointers (copies inside) -------------------------------- */
int gbp_fdem_forward_host(const gbp_fdem_system *sys, int B, int l_stride, const int32_t *nlayers,
                          const double *sigma, const double *thickness, const double *altitude,
                          double *out, int precision, int device);
int gbp_fdem_sensitivity_host(const gbp_fdem_system *sys, int B, int l_stride, const int32_t *nlayers,
                              const double *sigma, const double *thickness, const double *altitude,
                              double *out, double *J, int precision, int device);

/* ---- rjMCMC ------------------------------------------------------------------------------- */
/* Runs B independent chains (one warp each, persistent kernel).  data: [B][2F] observed ppm
 * (<= 0 or NaN = inactive channel, EmDataPoint.active); altitude: [B].
 * Random stream of chain b: Philox4x32-10, key = seed, counter = (block, first_index + b).
 * max_iterations > 0 caps the number of accept_reject+update pairs per chain (tests);
 * 0 = run to the reference's own termination rule (Inference1D.infer :650-677).
 * d_buf: struct in HOST memory whose members are DEVICE pointers. */
int gbp_rjmcmc_run(const gbp_fdem_system *sys, const gbp_options *opt, int B, const double *d_data,
                   const double *d_altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                   const gbp_chain_buffers *d_buf, int precision, void *stream);
/* Same with HOST buffers: uploads data/altitude, allocates + zeroes device results, runs,
 * downloads every non-NULL member of h_buf. */
int gbp_rjmcmc_run_host(const gbp_fdem_system *sys, const gbp_options *opt, int B, const double *data,
                        const double *altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                        const gbp_chain_buffers *h_buf, int precision, int device);
/* Posterior summaries of hitmaps on the device (replaces Mesh._mean / Mesh._percentile per sounding,
 * geobipy/src/classes/mesh/Mesh.py:80, :173-217, as used by Inference2D's mean / percentile sections):
 * d_hitmap [B][n_sig][n_depth]; sigma bin s of sounding b spans ln(sigma) in d_sig_lo[b] + [s, s+1) * dx.
 * d_mean [B][n_depth]; d_pct [n_pct][B][n_depth] = centre of the first bin whose cumulative count reaches p %
 * of the column total, in ln(sigma).  DEVICE pointers, stream ordered; at most 8 percentiles. */
int gbp_summarise_hitmap(const int32_t *d_hitmap, int B, int n_sig, int n_depth, const double *d_sig_lo, double dx,
                         const double *percentiles, int n_pct, double *d_mean, double *d_pct, void *stream);
/* The same plus the remaining per-cell summaries of the reference: d_mode [B][n_depth] = centre of the first fullest
 * bin (Mesh._mode, Mesh.py:138-165) and d_range_bins [B][n_depth] (int32) = the credible range of `credible_percent`
 * (Mesh._credible_range, Mesh.py:58-78) in BINS: |log10 hi - log10 lo| = bins * dx / ln(10).  d_mode / d_range_bins may be
 * NULL; a credible range uses two of the 8 percentile slots.  Percentile rule: first bin whose cumulative fraction
 * (fp64) reaches percent * 0.01 (fp64) - the reference's rule including its round-off at exact ties. */
int gbp_summarise_posterior(const int32_t *d_hitmap, int B, int n_sig, int n_depth, const double *d_sig_lo, double dx,
                            const double *percentiles, int n_pct, double credible_percent, double *d_mean, double *d_pct,
                            double *d_mode, int32_t *d_range_bins, void *stream);
/* Opacity, depth of investigation and opacity level from credible ranges in bins (Histogram.transparency / opacity
 * Histogram.py:330-354, :509-541; Inference2D.compute_doi Inference2D.py:493-532; Histogram.opacity_level :356-367).
 * d_group [B] (int32, or NULL = every sounding its own group) names the set over which the range is normalised to
 * [0, 1]: the sounding (Histogram.opacity of one hitmap) or its flight line (Inference2D.compute_opacity :1011-1023).
 * d_group_minmax: 2 x n_groups int32 of scratch (n_groups = B when d_group is NULL).  d_opacity [B][n_depth];
 * d_doi_cell [B] = depth cell of the depth of investigation (deepest cell, from the bottom up, whose opacity reaches
 * doi_percent / 100, never below index 0); d_level_cell [B] = depth cell of Histogram.opacity_level(level_percent).
 * Any output may be NULL.  DEVICE pointers, stream ordered. */
int gbp_opacity_doi(const int32_t *d_range_bins, int B, int n_depth, const int32_t *d_group, int n_groups, double doi_percent,
                    double level_percent, int32_t *d_group_minmax, double *d_opacity, int32_t *d_doi_cell,
                    int32_t *d_level_cell, void *stream);
/* The *_rjmcmc_run_host entry points keep their device buffers between calls (per device); this frees them. */
int gbp_release_host_buffers(void);

/* ---- time domain (SkyTEM-type systems) ---------------------------------------------------------- */
/* total number of data channels = (measured components) x windows of every system, system 0 first (TdemDataPoint channel order) */
int gbp_tdem_n_channels(const gbp_tdem_survey *sv);
/* The model-independent tables the kernels use (tests / inspection): freq [GBP_TD_NFREQ] spline-node
 * frequencies; MR, MI [C][GBP_TD_NFREQ] window operator d_c = sum_i MR[c][i] Re S_i + MI[c][i] Im S_i;
 * t_centre [C] window centres.  Any pointer may be NULL.  Host only, no device needed. */
int gbp_tdem_window_operator(const gbp_tdem_survey *sv, double *freq, double *MR, double *MI, double *t_centre);
/* algorithmic flop count of one forward: n_freq x n_abscissae admittance recursions (75 L + 39 each, the
 * SURVEY.md 8(d) convention) + the 2 C x 2 n_freq window products */
double gbp_tdem_flops_per_forward(const gbp_tdem_survey *sv, int n_layers);
double gbp_tdem_mufu_per_forward(const gbp_tdem_survey *sv, int n_layers);

/* out: [B][C] window averages of the SECONDARY field (dB/dt systems: V/(A m^4) for unit moment; B systems: the output
 * scaling's unit), J: [B][C][l_stride] = d out / d ln(sigma).  altitude = transmitter height above ground [B].  DEVICE
 * pointers, stream ordered.  Pinned on the reference's SkyTEM (Z, dB/dt, finite loop) and Tempest (X + Z, B, point dipole)
 * known-answer vectors. */
int gbp_tdem_forward(const gbp_tdem_survey *sv, int B, int l_stride, const int32_t *d_nlayers,
                     const double *d_sigma, const double *d_thickness, const double *d_altitude,
                     double *d_out, int precision, void *stream);
int gbp_tdem_sensitivity(const gbp_tdem_survey *sv, int B, int l_stride, const int32_t *d_nlayers,
                         const double *d_sigma, const double *d_thickness, const double *d_altitude,
                         double *d_out, double *d_J, int precision, void *stream);
/* HOST pointers (copies inside) */
int gbp_tdem_forward_host(const gbp_tdem_survey *sv, int B, int l_stride, const int32_t *nlayers,
                          const double *sigma, const double *thickness, const double *altitude,
                          double *out, int precision, int device);
int gbp_tdem_sensitivity_host(const gbp_tdem_survey *sv, int B, int l_stride, const int32_t *nlayers,
                              const double *sigma, const double *thickness, const double *altitude,
                              double *out, double *J, int precision, int device);
/* primary field of the transmitter dipole at the receiver during the windows, per system and measured component in channel
 * order ([n_systems x components]: Tempest PX, PZ; TdemDataPoint.forward :1008-1016 stacks PX, -PZ), in the output
 * scaling's unit.  Host only.  Returns the number of values written. */
int gbp_tdem_primary_field(const gbp_tdem_survey *sv, double *out);
/* rjMCMC for time-domain datapoints: data [B][C].  error_model 0: errors per system (gbp_options.n_systems must equal
 * sv->n_systems), additive error of channel c scaled by (t_c / 1 ms)^-0.5 (TdemDataPoint.std :329-379), Z-component dB/dt
 * systems.  error_model 1 (Tempest): data = secondary + primary field, gbp_options.n_systems = the number of measured
 * components; the transmitter height is not sampled for this datapoint type. */
int gbp_tdem_rjmcmc_run(const gbp_tdem_survey *sv, const gbp_options *opt, int B, const double *d_data,
                        const double *d_altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                        const gbp_chain_buffers *d_buf, int precision, void *stream);
int gbp_tdem_rjmcmc_run_host(const gbp_tdem_survey *sv, const gbp_options *opt, int B, const double *data,
                             const double *altitude, uint64_t seed, uint64_t first_index, int64_t max_iterations,
                             const gbp_chain_buffers *h_buf, int precision, int device);

#ifdef __cplusplus
}
#endif
#endif
