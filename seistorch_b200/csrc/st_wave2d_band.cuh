// Absorbing-frame ("band") cells of the ISO|HABC equation as a precomputed-tap gather.
//
// The one-way blend of acoustic_habc.py:79-204 is linear in (h1, h2) with time-invariant
// coefficients, so the whole frame update of a cell q is
//     Y(q) = sum_o F1[o](q) h1(q+o) + sum_o F2[o](q) h2(q+o)
// over 9 / 5 fixed offsets o in {0, +-z, +-x, +-2z, +-2x}.  st_wave2d_prepare() evaluates the
// taps ONCE per call from the generic definitions in st_wave2d_math.cuh (side ownership,
// corner diagonals, blend weights) together with their r-derivatives H1/H2; the per-step
// kernels then only gather:
//     forward : Y(p)      = sum_o F1[o](p)     h1(p+o) + F2[o](p) h2(p+o)
//     adjoint : Lam_i(p)  = sum_o F1[-o](p+o) L1(p+o) + F2[-o](p+o) L2(p+o)
//     gradient: g_r(p)   += L1(p) * ( sum_o H1[o](p) S1(p+o) + H2[o](p) S2(p+o) )
//               g_ciso(p)+= pre(p) L1(p) lap(S1)(p)
// The only non-local entries of the reference operator -- the wrap-around neighbour of the
// (bw+1)-deep strip, non-zero for at most 4 cells -- are applied as explicit fix-ups.
#pragma once
#include "st_wave2d.cuh"

// plane layout: F1[ST_NTAP1], F2[ST_NTAP2], H1[ST_NTAP1], H2[ST_NTAP2].  The first ST_NTAP1C (= 9) taps of h1 form the cross
// {0, +-z, +-x, +-2z, +-2x}; taps 9..12 are the four diagonal neighbours of the mixed derivative (TTI: ST_F_XZ), zero planes
// for the other equations (whose kernels loop over the first 9 only).
constexpr int ST_NTAP1 = 13, ST_NTAP1C = 9, ST_NTAP2 = 5;
constexpr int ST_TAP_PLANES = 2 * (ST_NTAP1 + ST_NTAP2);      // F1, F2, H1, H2
// tap offsets (dz, dx): 0:(0,0) 1:(-1,0) 2:(+1,0) 3:(0,-1) 4:(0,+1) 5:(-2,0) 6:(+2,0) 7:(0,-2) 8:(0,+2)
//                       9:(-1,-1) 10:(-1,+1) 11:(+1,-1) 12:(+1,+1)
ST_HD int st_tap_dz(int o) { return o == 1 ? -1 : o == 2 ? 1 : o == 5 ? -2 : o == 6 ? 2 : (o == 9 || o == 10) ? -1 : (o == 11 || o == 12) ? 1 : 0; }
ST_HD int st_tap_dx(int o) { return o == 3 ? -1 : o == 4 ? 1 : o == 7 ? -2 : o == 8 ? 2 : (o == 9 || o == 11) ? -1 : (o == 10 || o == 12) ? 1 : 0; }
ST_HD int st_tap_neg(int o) { return o == 0 ? 0 : o >= 9 ? 21 - o : ((o - 1) ^ 1) + 1; }      // index of -offset
// sign of the mixed-derivative stencil  cxz ((SE - SW) - (NE - NW))  at diagonal tap o (9..12)
ST_HD float st_tap_xz_sign(int o) { return (o == 9 || o == 12) ? 1.f : -1.f; }
// tap index of k steps along the inward normal of side s (0 top, 1 bottom, 2 left, 3 right)
ST_HD int st_tap_normal(int s, int k) { return (k == 1 ? 0 : 4) + (s == 0 ? 2 : s == 1 ? 1 : s == 2 ? 4 : 3); }

// compact enumeration of the cells closer than `bd` to an absorbing edge
struct BandCells {
    int bd, nx, nz, top_rows, n_top, n_bot, total;
};
ST_HD BandCells st_band_cells(const W2Geom& g, int bd) {
    BandCells c;
    c.bd = bd; c.nx = g.nx; c.nz = g.nz;
    c.top_rows = g.multiple ? 0 : bd;
    c.n_top = c.top_rows * g.nx;
    c.n_bot = bd * g.nx;
    c.total = c.n_top + c.n_bot + (g.nz - c.top_rows - bd) * 2 * bd;
    return c;
}
ST_HD bool st_band_ok(const W2Geom& g, int bd) { return g.nx >= 2 * bd && g.nz >= (g.multiple ? 1 : 2) * bd; }
ST_HD void st_band_decode(const BandCells& c, int i, int& z, int& x) {
    if (i < c.n_top) { z = i / c.nx; x = i - z * c.nx; return; }
    i -= c.n_top;
    if (i < c.n_bot) { const int r = i / c.nx; z = c.nz - c.bd + r; x = i - r * c.nx; return; }
    i -= c.n_bot;
    const int w = 2 * c.bd, r = i / w, col = i - r * w;
    z = c.top_rows + r;
    x = col < c.bd ? col : c.nx - w + col;
}
ST_HD int st_band_encode(const BandCells& c, int z, int x) {
    if (z < c.top_rows) return z * c.nx + x;
    if (z >= c.nz - c.bd) return c.n_top + (z - (c.nz - c.bd)) * c.nx + x;
    const int w = 2 * c.bd;
    if (x < c.bd) return c.n_top + c.n_bot + (z - c.top_rows) * w + x;
    if (x >= c.nx - c.bd) return c.n_top + c.n_bot + (z - c.top_rows) * w + (x - (c.nx - w));
    return -1;
}

// the (at most 4) cells whose one-way extrapolation wraps around the strip: cell q at depth
// bw-1 of side s reads depth 0 as its "j+2" neighbour.  k = 0..3 enumerates the candidates
// that can carry a non-zero weight (the ends of the TR / BL corner-block diagonals).
ST_HD void st_wrap_candidate(const W2Geom& g, int k, int& z, int& x, int& s, int& zw, int& xw) {
    const int w = g.bw;
    if (k == 0) { z = w - 1; x = g.nx - 1; s = 0; zw = 0; xw = x; }                 // TR diagonal end, top side
    else if (k == 1) { z = 0; x = g.nx - w; s = 3; zw = z; xw = g.nx - 1; }         // TR diagonal start, right side
    else if (k == 2) { z = g.nz - w; x = 0; s = 1; zw = g.nz - 1; xw = x; }         // BL diagonal start, bottom side
    else { z = g.nz - 1; x = w - 1; s = 2; zw = z; xw = 0; }                        // BL diagonal end, left side
}

// flag sets whose frame runs on the taps: the single-field equations (9 taps; 13 with the mixed derivative of tti_habc), and
// the Born pairs (acoustic_lsrtm_habc, acoustic_{vti,tti}_lsrtm_habc) -- both of its fields take the SAME
// taps (the one-way blend acts on each field by itself, acoustic_vti_lsrtm_habc.py:60-66), the scattered field adds
// the coupling term  pre m A[p1]  evaluated from the coefficient planes
ST_HD constexpr bool st_flags_tapped(int fl) {
    return fl == (ST_F_ISO | ST_F_HABC) || fl == ST_F_HABC || fl == (ST_F_ISO | ST_F_HABC | ST_F_G1) ||
           fl == (ST_F_HABC | ST_F_G1) || fl == (ST_F_HABC | ST_F_BORN) || fl == (ST_F_HABC | ST_F_XZ) ||
           fl == (ST_F_HABC | ST_F_XZ | ST_F_BORN);
}
// ... of which the straight top / bottom strips run on the vectorised strip blocks (single-field equations without mixed
// derivative only; the band threads serve every frame cell of the others)
ST_HD constexpr bool st_flags_stripped(int fl) { return st_flags_tapped(fl) && !(fl & (ST_F_BORN | ST_F_XZ)); }
// the forward strip blocks serve every tapped flag set (the mixed derivative and the Born pairs included)
#ifndef ST_STRIP_FWD_ALL
#define ST_STRIP_FWD_ALL 1
#endif
ST_HD constexpr bool st_flags_stripped_fwd(int fl) { return ST_STRIP_FWD_ALL ? st_flags_tapped(fl) : st_flags_stripped(fl); }
// the adjoint strip blocks also serve the mixed derivative and the Born pairs without it (all but the TTI Born pair)
#ifndef ST_STRIP_ADJ_ALL
#define ST_STRIP_ADJ_ALL 1
#endif
ST_HD constexpr bool st_flags_stripped_adj(int fl) {
    return ST_STRIP_ADJ_ALL ? st_flags_tapped(fl) && !((fl & ST_F_BORN) && (fl & ST_F_XZ)) : st_flags_stripped(fl);
}

#ifdef __CUDACC__
int st_wave2d_launch_prepare(int flags, const W2Args& a, cudaStream_t st);
#endif
