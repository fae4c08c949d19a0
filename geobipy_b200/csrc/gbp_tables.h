// Host-side (fp64) construction of the per-system filter-point table consumed by fdem_eval().
//
// Everything here is independent of the earth model and of the sounding, so it is computed once per
// system on the host: abscissae (FdemSystem.lamda0/lamda1, FdemSystem.py:67-101), air-layer wavenumber
// u0 = un[0] (initCoefficients, fdem1d_numba.py:172-185 with sigma = 0), geometry factors and filter
// weights of Hxx/Hxz/Hzx/Hzz (fdem1d_numba.py:307-438), the primary field H0 and the ppm
// normalisation 1e6*scale/H0 (fdem1d_numba.py:68).
#pragma once
#include <cmath>
#include <complex>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gbp_filter_tables.h"
#include "gbp_fdem.cuh"

namespace gbp {

struct SysHost {
    SysDev dev;
    std::vector<double> tab;  // TAB_ROWS x tab_stride, fp64
    std::vector<float> tab_f32;  // packed chunk layout of the fp32 path (gbp_fdem_f2.cuh): n_chunks x CHUNK_FLOATS
    std::string error;
};

inline bool build_system_tables(const gbp_fdem_system& sys, SysHost& out)
{
    typedef std::complex<double> cd;
    const double PI = 3.14159265358979323846;
    const double MU0 = 4.e-7 * PI, C0 = 299792458.0;
    const double EPS0 = 1.0 / (MU0 * (C0 * C0));
    const int F = sys.n_freq;
    if (F < 1 || F > GBP_MAXF) {
        out.error = "n_freq out of range";
        return false;
    }
    std::memset(&out.dev, 0, sizeof(SysDev));
    std::vector<double> lam, u0r, u0i, er, ei, cr, ci;
    int nseg = 0;
    for (int f = 0; f < F; ++f) {
        const int tid = sys.tid[f];
        if (!(tid == 1 || tid == 3 || tid == 7 || tid == 9)) {
            // fdem1d_numba.py:57-66 has no branch for y-oriented coils either
            out.error = "unsupported tensor id (only xx=1, 3, 7, zz=9 are handled by the reference)";
            return false;
        }
        const double omega = 2.0 * PI * sys.freq[f];
        const double dx = sys.rx[f] - sys.tx[f], dy = sys.ry[f] - sys.ty[f], dz = sys.rz[f] - sys.tz[f];
        const double sep = std::sqrt(dx * dx + dy * dy + dz * dz);
        const double r = 1.0 / sep, rx = dx, mom = sys.tmom[f], pi4 = 4.0 * PI;
        const double scale = sys.tmom[f] * sys.rmom[f];
        const double hSum = sys.rz[f];                      // rHeight + tHeight
        out.dev.hd0[f] = sys.rz[f] - 2.0 * sys.tz[f];       // rHeight - tHeight + 2*altitude
        out.dev.omu[f] = omega * MU0;
        out.dev.k2re[f] = -(omega * EPS0) * (omega * MU0);
        const cd ynzn_air = cd(0.0, omega * EPS0) * cd(0.0, omega * MU0);
        const size_t first = lam.size();
        cd H0 = 0.0;
        auto push = [&](double l, cd u0, cd e, cd c) {
            lam.push_back(l);
            u0r.push_back(u0.real());
            u0i.push_back(u0.imag());
            er.push_back(e.real());
            ei.push_back(e.imag());
            cr.push_back(c.real());
            ci.push_back(c.imag());
        };
        if (tid == 1 || tid == 9) {
            out.dev.seg[nseg++] = Seg{(int)lam.size(), GBP_NJ0, f, 0};
            for (int j = 0; j < GBP_NJ0; ++j) {
                const double l = std::pow(10.0, ((double)j * GBP_J0_S) + GBP_J0_A) * r;
                const cd u0 = std::sqrt(ynzn_air + l * l);
                if (tid == 9) {
                    const double a2 = mom / (pi4 * sep);
                    const cd c = (l * l * l) / u0 * (a2 * GBP_W0[j]);
                    H0 += std::exp(-u0 * hSum) * c;
                    push(l, u0, u0, c);
                } else {
                    const double c0 = -(mom / pi4) * r, d0 = c0 * ((rx * r) * (rx * r));
                    const double c = (l * l) * (d0 * GBP_W0[j]);
                    H0 += std::exp(-l * hSum) * c;
                    push(l, u0, cd(l, 0.0), cd(-c, 0.0));
                }
            }
        }
        if (tid == 1 || tid == 3 || tid == 7) {
            out.dev.seg[nseg++] = Seg{(int)lam.size(), GBP_NJ1, f, 0};
            for (int j = 0; j < GBP_NJ1; ++j) {
                const double l = std::pow(10.0, ((double)j * GBP_J1_S) + GBP_J1_A) * r;
                const cd u0 = std::sqrt(ynzn_air + l * l);
                if (tid == 1) {
                    const double c0 = -(mom / pi4) * r, d1 = c0 * (r - ((2.0 * rx * rx) * (r * r * r)));
                    const double c = l * (d1 * GBP_W1[j]);
                    H0 += std::exp(-l * hSum) * c;
                    push(l, u0, cd(l, 0.0), cd(-c, 0.0));
                } else {
                    const double d1 = (rx * mom) / (pi4 * sep);
                    const double c = (l * l) * (d1 * GBP_W1[j]);
                    if (tid == 3) {
                        H0 += std::exp(-l * hSum) * c;
                        push(l, u0, cd(l, 0.0), cd(-c, 0.0));
                    } else {
                        H0 += std::exp(-u0 * hSum) * c;
                        push(l, u0, u0, cd(-c, 0.0));
                    }
                }
            }
        }
        const cd coef = 1.e6 * scale / H0;
        for (size_t i = first; i < lam.size(); ++i) {
            const cd c = cd(cr[i], ci[i]) * coef;
            cr[i] = c.real();
            ci[i] = c.imag();
        }
    }
    const int n = (int)lam.size();
    const int stride = (n + 3) & ~3;  // rows stay 16-byte aligned for fp32 and fp64
    out.dev.n_freq = F;
    out.dev.n_seg = nseg;
    out.dev.n_items = n;
    out.dev.tab_stride = stride;
    out.tab.assign((size_t)TAB_ROWS * stride, 0.0);
    const std::vector<double>* rows[TAB_ROWS] = {&lam, &u0r, &u0i, &er, &ei, &cr, &ci};
    for (int q = 0; q < TAB_ROWS; ++q)
        for (int i = 0; i < n; ++i) out.tab[(size_t)q * stride + i] = (*rows[q])[i];
    // padded entries: lam = 1, u0 = 1, c = 0 -> contribute nothing if ever touched
    for (int i = n; i < stride; ++i) {
        out.tab[i] = 1.0;
        out.tab[(size_t)stride + i] = 1.0;
    }
    // fp32 path: chunks of 64 abscissae of one (frequency, filter) segment; lane l of chunk c owns abscissae
    // 64 c' + l and 64 c' + 32 + l of its segment (a missing partner: lam = u0 = 1, weight 0).  Four float4 per lane,
    // stored quad-major so that the 32 lanes of one LDS.128 are contiguous (conflict free):
    //   quad 0 = (lam_a, lam_b, u0r_a, u0r_b)  quad 1 = (u0i_a, u0i_b, er_a, er_b)
    //   quad 2 = (ei_a, ei_b, cr_a, cr_b)      quad 3 = (ci_a, ci_b, 0, 0)
    // Chunk ORDER: the frequencies are dealt into two halves of (nearly) equal chunk count (longest first), the chunks
    // of half 0 come first: a team of warps splits one forward at half_begin[1] (gbp_chain.cuh, team_round) and every
    // frequency is still summed by one warp in the same order, so results do not depend on the split.
    int nchunk = 0;
    out.tab_f32.clear();
    int cnt[GBP_MAXF] = {0}, half_of[GBP_MAXF] = {0}, load[2] = {0, 0};
    for (int sgi = 0; sgi < nseg; ++sgi) cnt[out.dev.seg[sgi].freq] += (out.dev.seg[sgi].count + CHUNK - 1) / CHUNK;
    {
        bool done[GBP_MAXF] = {false};
        for (int it = 0; it < F; ++it) {
            int best = -1;
            for (int f = 0; f < F; ++f)
                if (!done[f] && (best < 0 || cnt[f] > cnt[best])) best = f;
            done[best] = true;
            const int h = load[1] < load[0] ? 1 : 0;
            half_of[best] = h;
            load[h] += cnt[best];
        }
    }
    for (int h = 0; h < 2; ++h) {
        out.dev.half_begin[h] = nchunk;
        for (int sgi = 0; sgi < nseg; ++sgi) {
            const Seg& sg = out.dev.seg[sgi];
            if (half_of[sg.freq] != h) continue;
            for (int c0 = 0; c0 < sg.count; c0 += CHUNK) {
                if (nchunk >= MAX_CHUNK) {
                    out.error = "too many filter chunks";
                    return false;
                }
                out.dev.chunk_freq[nchunk] = (unsigned char)sg.freq;
                const size_t base = out.tab_f32.size();
                out.tab_f32.resize(base + CHUNK_FLOATS, 0.f);
                for (int l = 0; l < 32; ++l) {
                    const int ja = c0 + l, jb = c0 + 32 + l;
                    const bool va = ja < sg.count, vb = jb < sg.count;
                    const int ia = sg.start + (va ? ja : 0), ib = sg.start + (vb ? jb : 0);
                    auto get = [&](const std::vector<double>& r, bool v, int i, double dflt) { return (float)(v ? r[i] : dflt); };
                    float* q0 = &out.tab_f32[base + (size_t)(0 * 32 + l) * 4];
                    float* q1 = &out.tab_f32[base + (size_t)(1 * 32 + l) * 4];
                    float* q2 = &out.tab_f32[base + (size_t)(2 * 32 + l) * 4];
                    float* q3 = &out.tab_f32[base + (size_t)(3 * 32 + l) * 4];
                    q0[0] = get(lam, va, ia, 1.0); q0[1] = get(lam, vb, ib, 1.0);
                    q0[2] = get(u0r, va, ia, 1.0); q0[3] = get(u0r, vb, ib, 1.0);
                    q1[0] = get(u0i, va, ia, 0.0); q1[1] = get(u0i, vb, ib, 0.0);
                    q1[2] = get(er, va, ia, 0.0);  q1[3] = get(er, vb, ib, 0.0);
                    q2[0] = get(ei, va, ia, 0.0);  q2[1] = get(ei, vb, ib, 0.0);
                    q2[2] = get(cr, va, ia, 0.0);  q2[3] = get(cr, vb, ib, 0.0);
                    q3[0] = get(ci, va, ia, 0.0);  q3[1] = get(ci, vb, ib, 0.0);
                }
                ++nchunk;
            }
        }
    }
    out.dev.half_begin[2] = nchunk;
    {   // per-frequency work units, most chunks first
        int fb[GBP_MAXF], fc[GBP_MAXF] = {0};
        for (int c = nchunk - 1; c >= 0; --c) {
            fb[out.dev.chunk_freq[c]] = c;
            fc[out.dev.chunk_freq[c]]++;
        }
        bool done[GBP_MAXF] = {false};
        for (int r = 0; r < F; ++r) {
            int best = -1;
            for (int f = 0; f < F; ++f)
                if (!done[f] && (best < 0 || fc[f] > fc[best])) best = f;
            done[best] = true;
            out.dev.unit_begin[r] = (unsigned char)fb[best];
            out.dev.unit_count[r] = (unsigned char)fc[best];
        }
    }
    out.dev.n_chunks = nchunk;
    return true;
}

}  // namespace gbp
