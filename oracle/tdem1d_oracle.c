/* TEST INFRASTRUCTURE ONLY (oracle) - never linked or called by the product path.
 *
 * Plain-C fp64 restatement of the airborne time-domain EM forward model the reference calls for
 * SkyTEM-type systems:
 *
 *   TdemDataPoint.forward / .sensitivity / .fm_dlogc   classes/data/datapoint/TdemDataPoint.py:997-1055
 *   gaTdem1dfwd / ga_fm_dlogc / gaTdem1dsen             classes/forwardmodelling/Electromagnetic/TD/tdem1d.py:89-154
 *   Loop_pair.Geometry                                  classes/system/Loop_pair.py:62-78
 *   Model.Earth                                         classes/model/Model.py:152-159
 *
 * The arithmetic itself lives in a THIRD-PARTY dependency that is absent from /root/reference:
 * module `gatdaem1d` of GeoscienceAustralia/ga-aem (C++/FFTW, no version pinned anywhere in the
 * reference: not in pyproject.toml / requirements.txt, `.SUBMODULES.json` lists none; provenance is
 * documentation_source/source/content/getting_started/installation.rst:93-192).  What is restated here
 * is its published algorithm (Brodie, GA-AEM: layered-earth frequency-domain response of a magnetic
 * dipole/loop source at log-spaced frequencies -> cubic spline in log-frequency -> transmitter waveform
 * spectrum and receiver low-pass filters -> time domain -> window averages), with our own
 * discretisation of each step:
 *
 *   1. secondary B_z of a horizontal circular loop (radius a, unit moment) at height h over an L-layer
 *      earth, receiver at horizontal offset r and height h + dz:
 *          S(w) = mu0/(4 pi) * Int_0^inf rTE(lam, w) lam^2 exp(-lam (2h + dz)) J0(lam r) [2 J1(lam a)/(lam a)] dlam
 *      evaluated by the trapezoid rule on n_lam log-spaced abscissae (the .stm key
 *      NumberOfAbsiccaInHankelTransformEvaluation), lam_j = (2/ZH) exp(xi_j), ZH = 2h + dz;
 *   2. rTE from the admittance recursion written on the DIFFERENCE D_k = Y_k - lam (no cancellation at
 *      low induction numbers):  D_L = a_L,  D_k = a_k + 2 e u_k (D_{k+1} - a_k) / ((1+e) u_k + (1-e)(lam + D_{k+1})),
 *      a_k = i w mu0 sigma_k / (u_k + lam),  u_k = sqrt(lam^2 + i w mu0 sigma_k),  e = exp(-2 u_k t_k),
 *      rTE = -D_1 / (2 lam + D_1);
 *   3. everything after S(w_i) at the n_freq spline nodes is LINEAR in S and is folded into one real
 *      matrix per system, built by the caller (oracle_py.tdem_window_operator: not-a-knot cubic spline in
 *      log10 f, Fourier series of the piecewise-linear bipolar current waveform over the odd harmonics
 *      up to the digitising Nyquist frequency, low-pass filters as cascaded first-order stages, exact
 *      window averages, the reference's sign flip of the z component TdemDataPoint.py:1015-1016):
 *          d_c = sum_i MR[c][i] Re S_i + MI[c][i] Im S_i.
 *
 * PARITY PIN: the only known-answer vectors are the reference's
 * tests/data_checks/skytem_*_clean.csv (6 models x 79 soundings x 45 windows, tests/test_synthetic_data.py:32-48),
 * committed as tests/golden/skytem_clean.npz.  tests/test_oracle_golden.py checks this restatement against
 * them: median |rel. error| ~1e-3, tolerance stated there.  Two facts about gatdaem1d were inferred from
 * those vectors and are part of the restatement: (i) the transmitter is a finite loop of the .stm's
 * ModellingLoopRadius, not a point dipole; (ii) a LowPassFilter of order n is n cascaded first-order
 * stages 1/(1 + i f/fc)^n, not an n-th order Butterworth.  There are NO golden vectors for the TDEM
 * Jacobian: TDEM sensitivities are "parity unpinned" against the reference and are checked against
 * finite differences of this forward instead.
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <string.h>

#include "oracle.h"

#define MU0 (4e-7 * M_PI)

typedef double complex cplx;

/* per-sounding abscissae and geometry weights */
static void td_geometry(const gbo_tdem_system *s, double altitude, double *lam, double *wgt, double *wgtx)
{
    const double ZH = 2.0 * altitude + s->rx_dz;
    const double r = hypot(s->rx_dx, s->rx_dy);
    const double dxi = s->xi[1] - s->xi[0];
    for (int j = 0; j < s->n_lam; ++j) {
        const double l = (2.0 / ZH) * exp(s->xi[j]);
        const double t = dxi * ((j == 0 || j == s->n_lam - 1) ? 0.5 : 1.0);
        double w = t, wx = t;
        w *= l * l * l * exp(-l * ZH) * j0(l * r);
        /* horizontal (x) secondary field of the vertical dipole: the same integrand with J1(lam r) dx / r */
        wx *= l * l * l * exp(-l * ZH) * (j1(l * r) * s->rx_cx);
        if (s->loop_radius > 0.0) {
            const double x = l * s->loop_radius;
            w *= 2.0 * j1(x) / x;
            wx *= 2.0 * j1(x) / x;
        }
        lam[j] = l;
        wgt[j] = w * MU0 / (4.0 * M_PI);
        wgtx[j] = wx * MU0 / (4.0 * M_PI);
    }
}

/* S_i and (optionally) dS_i / d ln sigma_k for every spline node */
static void td_frequency_response(const gbo_tdem_system *s, double altitude, int L, const double *sigma,
                                  const double *thick, cplx *S, cplx *dS /* [n_freq][L] or NULL */, int xcomp)
{
    double lam[GBO_TD_MAXLAM], wgt[GBO_TD_MAXLAM], wgtx[GBO_TD_MAXLAM];
    td_geometry(s, altitude, lam, xcomp ? wgtx : wgt, xcomp ? wgt : wgtx);   /* wgt = the weights of the asked component */
    for (int i = 0; i < s->n_freq; ++i) {
        const double omu = 2.0 * M_PI * s->freq[i] * MU0;
        cplx acc = 0.0;
        if (dS)
            for (int k = 0; k < L; ++k) dS[i * L + k] = 0.0;
        for (int j = 0; j < s->n_lam; ++j) {
            const double l = lam[j];
            cplx u[GBO_MAXL], dDdD[GBO_MAXL], dDdu[GBO_MAXL];
            /* basement */
            u[L - 1] = csqrt(l * l + I * omu * sigma[L - 1]);
            cplx D = I * omu * sigma[L - 1] / (u[L - 1] + l);
            dDdu[L - 1] = 1.0;
            for (int k = L - 2; k >= 0; --k) {
                const cplx uk = csqrt(l * l + I * omu * sigma[k]);
                u[k] = uk;
                const cplx a = I * omu * sigma[k] / (uk + l);
                const double h = thick[k];
                cplx e = (2.0 * h * creal(uk) > 700.0) ? 0.0 : cexp(-2.0 * uk * h);
                const cplx E = D - a;
                const cplx Y = l + D;
                const cplx q = (1.0 + e) * uk + (1.0 - e) * Y;
                if (dS) {
                    dDdD[k] = 4.0 * e * uk * uk / (q * q);
                    dDdu[k] = 1.0 + 2.0 * e * (-2.0 * h * uk * E + E - uk) / q
                              - 2.0 * e * uk * E * ((1.0 + e) + 2.0 * h * e * E) / (q * q);
                }
                D = a + 2.0 * e * uk * E / q;
            }
            const cplx den = 2.0 * l + D;
            acc += wgt[j] * (-D / den);
            if (dS) {
                cplx P = wgt[j] * (-2.0 * l / (den * den)); /* w * d rTE / d D_1 */
                for (int k = 0; k < L; ++k) {
                    /* d D_k / d ln sigma_k = dD_k/du_k * i w mu0 sigma_k / (2 u_k) */
                    dS[i * L + k] += P * dDdu[k] * (I * omu * sigma[k] / (2.0 * u[k]));
                    if (k < L - 1) P *= dDdD[k];
                }
            }
        }
        S[i] = acc;
    }
}

int gbo_tdem_forward(const gbo_tdem_system *s, double altitude, int L, const double *sigma,
                     const double *thickness, double *out)
{
    if (L < 1 || L > GBO_MAXL || altitude <= 0.0) return 1;
    cplx S[2][GBO_TD_NFREQ];
    int any_x = 0;
    for (int c = 0; c < s->C; ++c) any_x |= (s->comp[c] == 1);
    td_frequency_response(s, altitude, L, sigma, thickness, S[0], NULL, 0);
    if (any_x) td_frequency_response(s, altitude, L, sigma, thickness, S[1], NULL, 1);
    for (int c = 0; c < s->C; ++c) {
        const cplx *Sc = S[s->comp[c] == 1];
        double d = 0.0;
        for (int i = 0; i < s->n_freq; ++i)
            d += s->MR[c * GBO_TD_NFREQ + i] * creal(Sc[i]) + s->MI[c * GBO_TD_NFREQ + i] * cimag(Sc[i]);
        out[c] = d;
    }
    return 0;
}

/* J[c][k] = d out_c / d ln sigma_k, row-major [C][L] (the reference's sigma-scaled Jacobian, tdem1d.py:152) */
int gbo_tdem_sensitivity(const gbo_tdem_system *s, double altitude, int L, const double *sigma,
                         const double *thickness, double *J)
{
    if (L < 1 || L > GBO_MAXL || altitude <= 0.0) return 1;
    cplx S[GBO_TD_NFREQ];
    static __thread cplx dS2[2][GBO_TD_NFREQ * GBO_MAXL];
    int any_x = 0;
    for (int c = 0; c < s->C; ++c) any_x |= (s->comp[c] == 1);
    td_frequency_response(s, altitude, L, sigma, thickness, S, dS2[0], 0);
    if (any_x) td_frequency_response(s, altitude, L, sigma, thickness, S, dS2[1], 1);
    for (int c = 0; c < s->C; ++c) {
        const cplx *dS = dS2[s->comp[c] == 1];
        for (int k = 0; k < L; ++k) {
            double d = 0.0;
            for (int i = 0; i < s->n_freq; ++i)
                d += s->MR[c * GBO_TD_NFREQ + i] * creal(dS[i * L + k]) + s->MI[c * GBO_TD_NFREQ + i] * cimag(dS[i * L + k]);
            J[c * L + k] = d;
        }
    }
    return 0;
}

/* the frequency-domain response itself (tests: spline-node values against dense quadrature) */
int gbo_tdem_frequency_response(const gbo_tdem_system *s, double altitude, int L, const double *sigma,
                                const double *thickness, double *S_re, double *S_im)
{
    if (L < 1 || L > GBO_MAXL || altitude <= 0.0) return 1;
    cplx S[GBO_TD_NFREQ];
    td_frequency_response(s, altitude, L, sigma, thickness, S, NULL, 0);
    for (int i = 0; i < s->n_freq; ++i) {
        S_re[i] = creal(S[i]);
        S_im[i] = cimag(S[i]);
    }
    return 0;
}
