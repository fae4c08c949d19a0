/* TEST INFRASTRUCTURE ONLY (oracle) - never linked or called by the product path.
 *
 * Plain-C, fp64 / complex128 restatement of the reference's 1D frequency-domain EM
 * forward operator and its analytic Jacobian.  Parity is PINNED: tests/test_oracle_golden.py
 * checks this file against (a) the reference's own golden CSVs
 * (tests/data_checks/resolve_*_clean.csv, committed as tests/golden/resolve_clean.npz) and
 * (b) outputs of the reference's Numba kernels run in the build container
 * (tests/golden/fdem_random_models.npz, made by tests/golden/make_golden.py).
 *
 * Reference followed (paths relative to geobipy/src/classes/forwardmodelling/Electromagnetic/FD/):
 *   fdem1d.py:29-52         geometry preparation           -> gbo_fdem_geometry()
 *   fdem1d_numba.py:158-191 initCoefficients               -> layer_coefficients()
 *   fdem1d_numba.py:195-219 M1_0 (admittance recursion)    -> reflection()
 *   fdem1d_numba.py:223-301 M1_1 (recursion + d/dln(sigma))-> reflection_sens()
 *   fdem1d_numba.py:307-438 Hxx/Hxz/Hzx/Hzz Hankel sums    -> hankel()
 *   fdem1d_numba.py:442-448 cTanh
 *   fdem1d_numba.py:25-68   nbFdem1dfwd                    -> gbo_fdem_forward()
 *   fdem1d_numba.py:72-121  nbFdem1dsen                    -> gbo_fdem_sensitivity()
 * The reference evaluates both filters for every frequency; this restatement evaluates only
 * the (frequency, filter) pairs each tensor id consumes - results are identical.
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gbp_filter_tables.h"
#include "oracle.h"

typedef double complex cplx;

static const double MU0 = 4.e-7 * M_PI;
static const double C_LIGHT = 299792458.0;

/* fdem1d_numba.py:442-448 */
static cplx c_tanh(cplx z)
{
    if (creal(z) > 0.0) {
        cplx t = cexp(-2.0 * z);
        return (1.0 - t) / (1.0 + t);
    } else {
        cplx t = cexp(2.0 * z);
        return (t - 1.0) / (t + 1.0);
    }
}

/* Per-abscissa state for an (L+1)-layer stack (layer 0 = air). */
typedef struct {
    cplx un[GBO_MAXL + 1];
    cplx Yn[GBO_MAXL + 1];
    cplx Y[GBO_MAXL + 2];
} stack_t;

/* fdem1d_numba.py:158-191. par[0] = 0 (air), par[1..L] = sigma. kappa = perm = 0. */
static void layer_coefficients(int L, double omega, double lam2, const double *par, stack_t *s)
{
    const double eps0 = 1.0 / (MU0 * (C_LIGHT * C_LIGHT));
    for (int k = 0; k <= L; ++k) {
        cplx yn = par[k] + (omega * eps0) * I;
        cplx zn = (omega * MU0) * I;
        cplx ynzn = yn * zn;
        cplx zn1 = 1.0 / zn;
        cplx tmp = csqrt(ynzn + lam2);
        s->un[k] = tmp;
        s->Yn[k] = tmp * zn1;
    }
    s->Y[L] = s->Yn[L];
}

/* fdem1d_numba.py:195-219: returns rTE, writes u0. thk[0] = 0 (air), thk[1..L]. */
static cplx reflection(int L, const double *thk, stack_t *s, cplx *u0)
{
    for (int k = L - 1; k >= 1; --k) {
        cplx Yn_ = s->Yn[k];
        cplx Y_ = s->Y[k + 1];
        cplx a0 = c_tanh(s->un[k] * thk[k]);
        s->Y[k] = Yn_ * (Y_ + (Yn_ * a0)) / (Yn_ + (Y_ * a0));
    }
    *u0 = s->un[0];
    return (s->Yn[0] - s->Y[1]) / (s->Yn[0] + s->Y[1]);
}

/* fdem1d_numba.py:131-154 + :223-301. sens[k], k = 0..L-1 is d rTE / d ln(sigma_k). */
static void reflection_sens(int L, double omega, const double *par, const double *thk, stack_t *s,
                            cplx *u0, cplx *sens)
{
    cplx accumulate[GBO_MAXL];
    if (L == 1) {
        sens[0] = par[1] / (2.0 * s->un[1]);
        *u0 = s->un[0];
        cplx a0 = s->Yn[0];
        cplx a1 = s->Y[1];
        cplx a2 = 1.0 / (a0 + a1);
        sens[0] = -2.0 * a0 * sens[0] * (a2 * a2);
        return;
    }
    for (int k = L - 1; k >= 1; --k) {
        int k2 = k - 1;
        double p = par[k];
        double t = thk[k];
        cplx oTmp = (omega * MU0 * t) * I;
        cplx Yn_ = s->Yn[k];
        cplx Yn_2 = Yn_ * Yn_;
        cplx Yn_3 = Yn_2 * Yn_;
        cplx Y_ = s->Y[k + 1];
        cplx Y_2 = Y_ * Y_;
        cplx un_ = s->un[k];
        cplx tanuh = c_tanh(un_ * t);
        cplx tanuh2 = tanuh * tanuh;
        cplx num = Y_ + (Yn_ * tanuh);
        cplx den = Yn_ + (Y_ * tanuh);
        s->Y[k] = Yn_ * num / den;
        accumulate[k2] = (Yn_2 * (1.0 - tanuh2)) / (den * den);
        cplx kappaFactor = oTmp * ((Y_2 * Yn_) - Yn_3);
        sens[k2] = (p / (2.0 * un_ * (den * den))) *
                   ((2.0 * Yn_ * Y_ * tanuh2) + (kappaFactor * tanuh2 - kappaFactor) +
                    ((Y_2 - Yn_2) * tanuh) + (2.0 * Yn_2));
    }
    sens[L - 1] = par[L] / (2.0 * s->un[L]);
    for (int k = 1; k < L - 1; ++k) accumulate[k] = accumulate[k] * accumulate[k - 1];
    *u0 = s->un[0];
    cplx a0 = s->Yn[0];
    cplx a1 = s->Y[1];
    cplx a2 = 1.0 / (a0 + a1);
    cplx top = -2.0 * a0 * (a2 * a2);
    sens[0] *= top;
    for (int k = 1; k < L; ++k) sens[k] *= top * accumulate[k - 1];
}

/* fdem1d.py:31-34 */
void gbo_fdem_geometry(const gbo_fdem_system *sys, double altitude, double *tHeight, double *rHeight,
                       double *scale, double *xsep, double *sep)
{
    for (int i = 0; i < sys->n_freq; ++i) {
        tHeight[i] = altitude + sys->tz[i];
        rHeight[i] = -tHeight[i] + sys->rz[i];
        scale[i] = sys->tmom[i] * sys->rmom[i];
        double dx = sys->rx[i] - sys->tx[i];
        double dy = sys->ry[i] - sys->ty[i];
        double dz = sys->rz[i] - sys->tz[i];
        xsep[i] = dx;
        sep[i] = sqrt(dx * dx + dy * dy + dz * dz);
    }
}

/* One frequency.  nout = 1 (forward; v = rTE) or L (sensitivity; v = d rTE/d ln sigma_k).
 * Writes H[k] and H0 following Hxx/Hxz/Hzx/Hzz term by term (fdem1d_numba.py:307-438). */
static void one_frequency(int tid, double freq, double tHeight, double rHeight, double moment, double rx,
                          double separation, int L, const double *par, const double *thk, int want_sens,
                          cplx *H, cplx *H0)
{
    const double pi4 = 4.0 * M_PI;
    const double omega = 2.0 * M_PI * freq;
    const int nout = want_sens ? L : 1;
    const double hSum = rHeight + tHeight;
    const double hDiff = rHeight - tHeight;
    const double r = 1.0 / separation;
    stack_t st;
    cplx v[GBO_MAXL], u0;

    for (int k = 0; k < nout; ++k) H[k] = 0.0;
    *H0 = 0.0;

    const int useJ0 = (tid == 1 || tid == 9);
    const int useJ1 = (tid == 1 || tid == 3 || tid == 7);

    if (useJ0) {
        for (int jc = 0; jc < GBP_NJ0; ++jc) {
            double lam = pow(10.0, ((double)jc * GBP_J0_S) + GBP_J0_A) * r;
            double lam2 = lam * lam;
            layer_coefficients(L, omega, lam2, par, &st);
            if (want_sens) reflection_sens(L, omega, par, thk, &st, &u0, v);
            else v[0] = reflection(L, thk, &st, &u0);
            if (tid == 9) { /* Hzz :411-438 */
                double a2 = moment / (pi4 * separation);
                double w0_ = a2 * GBP_W0[jc];
                cplx a0 = cexp(-u0 * hSum);
                cplx a1 = (lam * lam * lam) / u0;
                for (int k = 0; k < nout; ++k) H[k] += ((a0 + (v[k] * cexp(u0 * hDiff))) * a1) * w0_;
                *H0 += (a0 * a1) * w0_;
            } else { /* Hxx, J0 part :322-335 */
                double c0 = -(moment / pi4) * r;
                double d0 = c0 * ((rx * r) * (rx * r));
                double w0_ = d0 * GBP_W0[jc];
                double a0 = exp(-lam * hSum);
                for (int k = 0; k < nout; ++k) H[k] += ((a0 - (v[k] * exp(lam * hDiff))) * lam2) * w0_;
                *H0 += (a0 * lam2) * w0_;
            }
        }
    }
    if (useJ1) {
        for (int jc = 0; jc < GBP_NJ1; ++jc) {
            double lam = pow(10.0, ((double)jc * GBP_J1_S) + GBP_J1_A) * r;
            double lam2 = lam * lam;
            layer_coefficients(L, omega, lam2, par, &st);
            if (want_sens) reflection_sens(L, omega, par, thk, &st, &u0, v);
            else v[0] = reflection(L, thk, &st, &u0);
            if (tid == 1) { /* Hxx, J1 part :337-353 */
                double c0 = -(moment / pi4) * r;
                double d1 = c0 * (r - ((2.0 * rx * rx) * (r * r * r)));
                double w1_ = d1 * GBP_W1[jc];
                double b0 = exp(-lam * hSum);
                for (int k = 0; k < nout; ++k) H[k] += ((b0 - (v[k] * exp(lam * hDiff))) * lam) * w1_;
                *H0 += (b0 * lam) * w1_;
            } else if (tid == 3) { /* Hxz :359-381 */
                double d1 = (rx * moment) / (pi4 * separation);
                double w1_ = d1 * GBP_W1[jc];
                double b0 = exp(-lam * hSum);
                for (int k = 0; k < nout; ++k) H[k] += ((b0 - (v[k] * exp(lam * hDiff))) * lam2) * w1_;
                *H0 += (b0 * lam2) * w1_;
            } else { /* Hzx :385-408 */
                double d1 = (rx * moment) / (pi4 * separation);
                double w1_ = d1 * GBP_W1[jc];
                cplx b0 = cexp(-u0 * hSum);
                for (int k = 0; k < nout; ++k) H[k] += ((b0 - (v[k] * cexp(u0 * hDiff))) * lam2) * w1_;
                *H0 += (b0 * lam2) * w1_;
            }
        }
    }
}

static int pad_model(int L, const double *sigma, const double *thickness, double *par, double *thk)
{
    if (L < 1 || L > GBO_MAXL) return -1;
    par[0] = 0.0;
    thk[0] = 0.0;
    for (int k = 0; k < L; ++k) {
        par[k + 1] = sigma[k];
        thk[k + 1] = thickness[k];
    }
    return 0;
}

/* nbFdem1dfwd.  out[0..F) = real, out[F..2F) = imag  (FdemDataPoint.py:543-545). */
int gbo_fdem_forward(const gbo_fdem_system *sys, double altitude, int L, const double *sigma,
                     const double *thickness, double *out)
{
    double par[GBO_MAXL + 1], thk[GBO_MAXL + 1];
    double tH[GBO_MAXF], rH[GBO_MAXF], scl[GBO_MAXF], xs[GBO_MAXF], sp[GBO_MAXF];
    if (pad_model(L, sigma, thickness, par, thk)) return -1;
    const int F = sys->n_freq;
    gbo_fdem_geometry(sys, altitude, tH, rH, scl, xs, sp);
    for (int i = 0; i < F; ++i) {
        cplx H, H0;
        one_frequency(sys->tid[i], sys->freq[i], tH[i], rH[i], sys->tmom[i], xs[i], sp[i], L, par, thk, 0, &H, &H0);
        cplx d = 1.e6 * scl[i] * ((H - H0) / H0);
        out[i] = creal(d);
        out[F + i] = cimag(d);
    }
    return 0;
}

/* nbFdem1dsen.  J is row-major [2F][L]: rows 0..F-1 real, F..2F-1 imag (FdemDataPoint.py:552-557). */
int gbo_fdem_sensitivity(const gbo_fdem_system *sys, double altitude, int L, const double *sigma,
                         const double *thickness, double *J)
{
    double par[GBO_MAXL + 1], thk[GBO_MAXL + 1];
    double tH[GBO_MAXF], rH[GBO_MAXF], scl[GBO_MAXF], xs[GBO_MAXF], sp[GBO_MAXF];
    if (pad_model(L, sigma, thickness, par, thk)) return -1;
    const int F = sys->n_freq;
    gbo_fdem_geometry(sys, altitude, tH, rH, scl, xs, sp);
    for (int i = 0; i < F; ++i) {
        cplx dH[GBO_MAXL], dH0;
        one_frequency(sys->tid[i], sys->freq[i], tH[i], rH[i], sys->tmom[i], xs[i], sp[i], L, par, thk, 1, dH, &dH0);
        for (int k = 0; k < L; ++k) {
            cplx d = 1.e6 * scl[i] * (dH[k] - dH0) / dH0;
            J[(size_t)i * L + k] = creal(d);
            J[(size_t)(F + i) * L + k] = cimag(d);
        }
    }
    return 0;
}
