// Host-side (fp64) construction of everything model independent in the time-domain forward:
//   * the spline-node frequencies shared by the systems of a datapoint type,
//   * the Hankel abscissae (as xi = ln(lambda ZH / 2)) and trapezoid weights,
//   * the window operator Mt: window averages of dBz/dt from the node values of the frequency-domain
//     secondary field S(f):   d_c = sum_i Mt[i][c] Re S_i + Mt[32 + i][c] Im S_i.
//
// Built from a gbp_tdem_survey = the contents of the .stm files gatdaem1d's TDAEMSystem reads
// (classes/system/TdemSystem_GAAEM.py:26-40): base frequency, piecewise-linear current waveform over half a
// period (the second half is its negative), digitising frequency (-> Nyquist limit of the harmonic sum),
// receiver windows, low-pass filters (order n = n cascaded first-order stages), ModellingLoopRadius.
//
//   dBz/dt(t) = sum over odd harmonics n of  2 Re[ dI_n S(f_n) F(f_n) exp(i w_n t) ],
//   dI_n = (2/T) sum_seg slope_seg (exp(-i w_n t_a) - exp(-i w_n t_b)) / (i w_n)      (Fourier series of dI/dt)
//   window average over [ta, tb]: factor (exp(i w tb) - exp(i w ta)) / (i w (tb - ta))
//   S(f_n) = not-a-knot cubic spline in log10 f through the 32 node values (linear in them).
// Sign: the reference negates gatdaem1d's z component (TdemDataPoint.py:1015-1016); with z up and
// e^{+iwt} that is the minus sign folded into Mt here.
#pragma once
#include <cmath>
#include <complex>
#include <cstring>
#include <string>
#include <vector>

#include "gbp_tdem.cuh"

namespace gbp {

constexpr double TD_XI_LO = -4.6, TD_XI_HI = 2.3;  // ln(lambda ZH / 2) range of the Hankel trapezoid rule

struct TdHost {
    TdDev dev;
    double freq[TD_NF];
    std::vector<double> Mt;  // TD_ROWS x TD_CP, fp64, unscaled
    std::string error;
};

// dense solve A X = B (n x n, m right-hand sides), partial pivoting
inline bool td_solve(int n, int m, std::vector<double>& A, std::vector<double>& B)
{
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
        if (A[p * n + c] == 0.0) return false;
        if (p != c) {
            for (int k = 0; k < n; ++k) std::swap(A[p * n + k], A[c * n + k]);
            for (int k = 0; k < m; ++k) std::swap(B[p * m + k], B[c * m + k]);
        }
        for (int r = c + 1; r < n; ++r) {
            const double f = A[r * n + c] / A[c * n + c];
            if (f == 0.0) continue;
            for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
            for (int k = 0; k < m; ++k) B[r * m + k] -= f * B[c * m + k];
        }
    }
    for (int c = n - 1; c >= 0; --c) {
        for (int k = 0; k < m; ++k) {
            double s = B[c * m + k];
            for (int j = c + 1; j < n; ++j) s -= A[c * n + j] * B[j * m + k];
            B[c * m + k] = s / A[c * n + c];
        }
    }
    return true;
}

// what a system measures: components (x, z) with their output scalings, PeakCurrent, B or dB/dt.  An all-zero tail of the
// struct (a caller built against the round-1 header) means Z only, scaling 1, 1 A, dB/dt.
struct TdOutput {
    int n_comp;
    int comp[2];        // 1 = x, 0 = z, in channel order (x first: TdemDataPoint.forward :1008-1016)
    double scale[2];    // sign x PeakCurrent x Output Scaling
    bool b_field;
    double peak;
};
inline TdOutput td_output(const gbp_tdem_system& y)
{
    TdOutput o;
    o.b_field = y.output_type == 1;
    o.peak = y.peak_current != 0.0 ? y.peak_current : 1.0;
    double sx = y.x_scaling, sz = y.z_scaling;
    if (sx == 0.0 && sz == 0.0) sz = 1.0;
    // signs, pinned by the reference's known-answer vectors: a dB/dt system reports the receiver voltage -dB/dt (SkyTEM,
    // z), a B system the field itself (Tempest, z); the x component has the opposite sign of z (Tempest; unpinned for dB/dt)
    const double sgn_z = o.b_field ? 1.0 : -1.0;
    o.n_comp = 0;
    if (sx != 0.0) {
        o.comp[o.n_comp] = 1;
        o.scale[o.n_comp++] = -sgn_z * o.peak * sx;
    }
    if (sz != 0.0) {
        o.comp[o.n_comp] = 0;
        o.scale[o.n_comp++] = sgn_z * o.peak * sz;
    }
    return o;
}
inline int td_n_channels(const gbp_tdem_survey& sv)
{
    int c = 0;
    for (int s = 0; s < sv.n_systems && s < GBP_TD_MAXSYS; ++s) c += sv.sys[s].n_windows * td_output(sv.sys[s]).n_comp;
    return c;
}

inline bool build_tdem_tables(const gbp_tdem_survey& sv, TdHost& out)
{
    typedef std::complex<double> cd;
    const double PI = 3.14159265358979323846, MU0 = 4.e-7 * PI;
    const cd I(0.0, 1.0);
    std::memset(&out.dev, 0, sizeof(TdDev));
    if (sv.n_systems < 1 || sv.n_systems > GBP_TD_MAXSYS) {
        out.error = "n_systems must be 1 or 2";
        return false;
    }
    double flo = 1e300, fhi = 0.0;
    int nab = 0, C = 0;
    for (int s = 0; s < sv.n_systems; ++s) {
        const gbp_tdem_system& y = sv.sys[s];
        if (y.n_wave < 2 || y.n_wave > GBP_TD_MAXWAVE || y.n_windows < 1 || y.n_windows > GBP_TD_MAXWIN ||
            y.n_filters < 0 || y.n_filters > GBP_TD_MAXFILT || !(y.base_frequency > 0.0) ||
            !(y.digitising_frequency > 2.0 * y.base_frequency)) {
            out.error = "invalid time-domain system description";
            return false;
        }
        const double span = y.wave_time[y.n_wave - 1] - y.wave_time[0];
        if (std::fabs(span * 2.0 * y.base_frequency - 1.0) > 1e-3 && std::fabs(span * y.base_frequency - 1.0) > 1e-3) {
            out.error = "the current waveform must span half a period or one period of the base frequency";
            return false;
        }
        for (int i = 0; i < y.n_windows; ++i)
            if (!(y.window_end[i] > y.window_start[i])) {
                out.error = "receiver window with non-positive width";
                return false;
            }
        flo = std::fmin(flo, y.base_frequency);
        fhi = std::fmax(fhi, 0.5 * y.digitising_frequency);
        if (y.n_abscissae > nab) nab = y.n_abscissae;
        if (y.loop_radius != sv.sys[0].loop_radius) {
            out.error = "systems of one datapoint type must share ModellingLoopRadius";
            return false;
        }
        C += y.n_windows * td_output(y).n_comp;
    }
    if (C > GBP_TD_MAXC) {
        out.error = "too many windows";
        return false;
    }
    int n_lam = 2 * ((nab + 1) / 2);  // NumberOfAbsiccaInHankelTransformEvaluation, rounded up to even
    if (n_lam < 8) n_lam = 8;
    if (n_lam > GBP_TD_MAXLAM) n_lam = GBP_TD_MAXLAM;
    TdDev& d = out.dev;
    d.n_sys = sv.n_systems;
    d.n_lam = n_lam;
    d.C = C;
    d.rx_r = std::hypot(sv.rx_dx, sv.rx_dy);
    d.rx_cx = d.rx_r > 0.0 ? sv.rx_dx / d.rx_r : 0.0;
    d.rx_dz = sv.rx_dz;
    d.loop_radius = sv.sys[0].loop_radius;
    const double dxi = (TD_XI_HI - TD_XI_LO) / (double)(n_lam - 1);
    for (int j = 0; j < n_lam; ++j) {
        d.xi[j] = (j == n_lam - 1) ? TD_XI_HI : TD_XI_LO + dxi * (double)j;
        d.tw[j] = dxi * ((j == 0 || j == n_lam - 1) ? 0.5 : 1.0) * MU0 / (4.0 * PI);
    }
    std::vector<double> lf(TD_NF);
    for (int i = 0; i < TD_NF; ++i) {
        out.freq[i] = flo * std::pow(fhi / flo, (double)i / (double)(TD_NF - 1));
        lf[i] = std::log10(out.freq[i]);
        d.omu[i] = 2.0 * PI * out.freq[i] * MU0;
    }
    // not-a-knot cubic spline: second derivatives m = K y, K = A^-1 R (both 32 x 32)
    const int n = TD_NF;
    std::vector<double> A(n * n, 0.0), K(n * n, 0.0);
    std::vector<double> h(n - 1);
    for (int i = 0; i < n - 1; ++i) h[i] = lf[i + 1] - lf[i];
    for (int i = 1; i < n - 1; ++i) {
        A[i * n + i - 1] = h[i - 1];
        A[i * n + i] = 2.0 * (h[i - 1] + h[i]);
        A[i * n + i + 1] = h[i];
        K[i * n + i - 1] = 6.0 / h[i - 1];
        K[i * n + i] = -6.0 / h[i - 1] - 6.0 / h[i];
        K[i * n + i + 1] = 6.0 / h[i];
    }
    A[0] = h[1];
    A[1] = -(h[0] + h[1]);
    A[2] = h[0];
    A[(n - 1) * n + n - 3] = h[n - 2];
    A[(n - 1) * n + n - 2] = -(h[n - 3] + h[n - 2]);
    A[(n - 1) * n + n - 1] = h[n - 3];
    if (!td_solve(n, n, A, K)) {
        out.error = "spline system is singular";
        return false;
    }
    out.Mt.assign((size_t)TD_ROWS * TD_CP, 0.0);
    int c0 = 0;
    for (int s = 0; s < sv.n_systems; ++s) {
        const gbp_tdem_system& y = sv.sys[s];
        d.n_win[s] = y.n_windows;
        const double T = 1.0 / y.base_frequency;
        const TdOutput o = td_output(y);
        const int nharm = (int)(0.5 * y.digitising_frequency / y.base_frequency);
        std::vector<double> g(n);
        for (int hn = 1; hn <= nharm; hn += 2) {
            const double f = (double)hn * y.base_frequency, w = 2.0 * PI * f;
            cd dn = 0.0;
            for (int j = 0; j + 1 < y.n_wave; ++j) {
                const double dt = y.wave_time[j + 1] - y.wave_time[j];
                const double slope = (y.wave_current[j + 1] - y.wave_current[j]) / dt;
                if (slope != 0.0) dn += slope * (std::exp(-I * (w * y.wave_time[j])) - std::exp(-I * (w * y.wave_time[j + 1]))) / (I * w);
            }
            const double span = y.wave_time[y.n_wave - 1] - y.wave_time[0];
            dn *= (std::fabs(span * y.base_frequency - 1.0) <= 1e-3 ? 1.0 : 2.0) / T;   // whole period given, or half of it
            cd F = 1.0;
            for (int k = 0; k < y.n_filters; ++k) {
                const cd stage = 1.0 / (1.0 + I * (f / y.filter_cutoff[k]));
                for (int o = 0; o < y.filter_order[k]; ++o) F *= stage;
            }
            // spline basis values at log10 f: g[k] = d S(f) / d y_k
            const double x = std::log10(f);
            int iv = 0;
            while (iv < n - 2 && x > lf[iv + 1]) ++iv;
            const double hh = h[iv], a = lf[iv + 1] - x, b = x - lf[iv];
            const double ca = a * a * a / (6.0 * hh) - hh * a / 6.0, cb = b * b * b / (6.0 * hh) - hh * b / 6.0;
            for (int k = 0; k < n; ++k) g[k] = ca * K[iv * n + k] + cb * K[(iv + 1) * n + k];
            g[iv] += a / hh;
            g[iv + 1] += b / hh;
            for (int i = 0; i < y.n_windows; ++i) {
                const double ta = y.window_start[i], tb = y.window_end[i];
                cd Aw = dn * F * (std::exp(I * (w * tb)) - std::exp(I * (w * ta))) / (I * w * (tb - ta));
                if (o.b_field) Aw /= I * w;   // B = the integral of dB/dt
                // the time signal of harmonic coefficient A is 2 Re(A S e^{iwt}): (2 Re A) Re S - (2 Im A) Im S
                for (int q = 0; q < o.n_comp; ++q) {
                    const int c = c0 + q * y.n_windows + i;
                    for (int k = 0; k < n; ++k) {
                        out.Mt[(size_t)k * TD_CP + c] += o.scale[q] * 2.0 * Aw.real() * g[k];
                        out.Mt[(size_t)(TD_NF + k) * TD_CP + c] += -o.scale[q] * 2.0 * Aw.imag() * g[k];
                    }
                }
            }
        }
        for (int q = 0; q < o.n_comp; ++q)
            for (int i = 0; i < y.n_windows; ++i) {
                const int c = c0 + q * y.n_windows + i;
                const double tc = 0.5 * (y.window_start[i] + y.window_end[i]);  // off_time = window centre
                d.tsc[c] = std::exp(-0.5 * (std::log(tc) - std::log(1e-3)));
                d.csys[c] = s;
                d.ccomp[c] = o.comp[q];
                if (o.comp[q] == 1) d.has_x = 1;
            }
        c0 += y.n_windows * o.n_comp;
    }
    if (sv.error_model == 1) {
        // Tempest_datapoint: errors per COMPONENT in channel order (x, then z), a fixed additive level per channel in place
        // of the t^-1/2 scale, and the predicted primary field of the channel's component added to the response
        if (sv.n_systems != 1) {
            out.error = "the Tempest error model takes one system";
            return false;
        }
        d.tempest = 1;
        const TdOutput o = td_output(sv.sys[0]);
        const double x = sv.rx_dx, yy = sv.rx_dy, z = sv.rx_dz, R = std::sqrt(x * x + yy * yy + z * z);
        const double sx = sv.sys[0].x_scaling, sz = (sx == 0.0 && sv.sys[0].z_scaling == 0.0) ? 1.0 : sv.sys[0].z_scaling;
        for (int c = 0; c < C; ++c) {
            if (!(sv.additive_level[c] > 0.0)) {
                out.error = "the Tempest error model needs a positive additive level for every channel";
                return false;
            }
            d.tsc[c] = sv.additive_level[c];
            d.csys[c] = (d.ccomp[c] == 1) ? 0 : (d.has_x ? 1 : 0);
            d.poff[c] = d.ccomp[c] == 1 ? 1e-7 * o.peak * 3.0 * x * z / std::pow(R, 5) * sx
                                        : -1e-7 * o.peak * (3.0 * z * z / std::pow(R, 5) - 1.0 / std::pow(R, 3)) * sz;
        }
    }
    return true;
}

}  // namespace gbp
