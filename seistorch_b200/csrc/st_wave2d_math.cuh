// Per-cell arithmetic of the 2D second-order wave-equation family.
//
// One set of __host__ __device__ functions serves the sm_100a kernels
// (st_wave2d.cu) and the host-side emulation used by the CPU unit tests
// (tests/hostcheck), so the adjoint algebra is checked on CPU against the
// oracle before it ever runs on a GPU.
//
// Reference behaviour restated here (file:line relative to the reference tree):
//   equations2d/acoustic.py:73-86              damped (PML) update
//   equations2d/acoustic_habc.py:79-101,147-221 one-way blend + side assignment
//   seistorch/habc.py:4-40                     side masks
//   equations2d/vti_habc2.py:25-60, tti_habc.py:25-62   anisotropic Laplacian
//   equations2d/acoustic_fwim_habc.py:31-67    first-derivative terms
//   equations2d/acoustic_{vti,tti}_lsrtm_habc.py  Born pair
//
// Formulation (see DESIGN.md "numerics"): fields are advanced in increment form
//   y = h1 + alpha*(h1 - h2) + A[h1]
// which is algebraically identical to the reference's expression but carries
// ~15x less fp32 rounding noise.  Coefficients are dimensionless and
// precomputed once per call from the model parameters (seistorch_b200/coefficients.py):
//   r = vp*dt/h (one-way blend: lam = 2r, mu = r^2), b = blend weight / damping,
//   cxx, czz, cxz, ax, az = spatial-operator coefficients, m = reflectivity.
// ISO flag sets share one Laplacian coefficient: cxx holds  ciso = r^2  (HABC) or
// r^2/(1+b dt) (PML), and for PML czz holds  alpha = (1-b dt)/(1+b dt).
#pragma once
#include "st_common.cuh"

struct W2Geom {
    int nz, nx, ld;     // grid and row pitch (floats)
    int bw;             // absorbing frame width (50)
    int multiple;       // free surface on top (no top frame)
};

struct W2Coef {         // per-cell coefficient values
    float r, b, cxx, czz, cxz, ax, az, m;
};

// ----------------------------------------------------------------- HABC geometry
ST_HD bool w2_in_frame(int z, int x, const W2Geom& g) {
    return (x < g.bw) || (x >= g.nx - g.bw) || (z >= g.nz - g.bw) || (!g.multiple && z < g.bw);
}

// weight of each side's one-way blend (0:top 1:bottom 2:left 3:right) in the final
// value of cell (z,x): closed form of the masks (habc.py:4-40) and of the assignment
// order top,bottom,left,right followed by the four corner-block main diagonals
// (acoustic_habc.py:165-202).
ST_HD void w2_side_weights(int z, int x, const W2Geom& g, float f[4]) {
    const int w = g.bw, nz = g.nz, nx = g.nx;
    f[0] = f[1] = f[2] = f[3] = 0.f;
    const int zb = nz - 1 - z, xr = nx - 1 - x;
    if (!g.multiple && z < w && z <= x && x <= nx - 1 - z) f[0] = 1.f;
    if (zb < w && zb <= x && x <= nx - 1 - zb) { f[0] = 0.f; f[1] = 1.f; }
    const bool tri = (g.multiple && z < w);
    if (x < w && ((x <= z && x <= nz - 1 - z) || tri)) { f[0] = f[1] = 0.f; f[2] = 1.f; }
    if (xr < w && ((xr <= z && xr <= nz - 1 - z) || tri)) { f[0] = f[1] = f[2] = 0.f; f[3] = 1.f; }
    if (!g.multiple && z < w) {
        if (x == z) { f[0] = 0.5f; f[2] = 0.5f; f[1] = f[3] = 0.f; }
        if (x == nx - w + z) { f[0] = 0.5f; f[3] = 0.5f; f[1] = f[2] = 0.f; }
    }
    const int i = z - (nz - w);
    if (i >= 0) {
        if (x == i) { f[1] = 0.5f; f[2] = 0.5f; f[0] = f[3] = 0.f; }
        if (x == nx - w + i) { f[1] = 0.5f; f[3] = 0.5f; f[0] = f[2] = 0.f; }
    }
}

// Side (0 top, 1 bottom, 2 left, 3 right) whose STRAIGHT part holds every frame cell that can interact with
// (z,x) through a one-way blend, or -1: (z,x) is at most bw deep on that side and at least bw+2 away from the
// two adjacent sides, so the cells q = p - k n (k = 0..2 along the inward normal n) that read it are owned by
// that side alone with weight 1 (no corner ownership, no corner diagonal), and no other side reaches it.
ST_HD int w2_straight_side(int z, int x, const W2Geom& g) {
    const int w = g.bw, m = g.bw + 2;
    const bool xmid = x >= m && x < g.nx - m, zmid = z >= m && z < g.nz - m;
    if (xmid) {
        if (!g.multiple && z <= w) return (g.nz - 1 - z > w) ? 0 : -1;
        if (g.nz - 1 - z <= w) return (g.multiple || z > w) ? 1 : -1;
    }
    if (zmid) {
        if (x <= w) return (g.nx - 1 - x > w) ? 2 : -1;
        if (g.nx - 1 - x <= w) return (x > w) ? 3 : -1;
    }
    return -1;
}

ST_HD int w2_depth(int s, int z, int x, const W2Geom& g) {
    return s == 0 ? z : (s == 1 ? g.nz - 1 - z : (s == 2 ? x : g.nx - 1 - x));
}

// coordinates of the cell at depth j from side s, same lateral position as (z,x)
ST_HD void w2_at_depth(int s, int j, int z, int x, const W2Geom& g, int& zz, int& xx) {
    zz = z; xx = x;
    if (s == 0) zz = j;
    else if (s == 1) zz = g.nz - 1 - j;
    else if (s == 2) xx = j;
    else xx = g.nx - 1 - j;
}

// one-way (Higdon) extrapolation differences of side s at (z,x):
//   one = 2h1_j - h2_j + lam*Dlam + mu*Dmu,  lam = 2r, mu = r^2   (acoustic_habc.py:89-98)
// neighbours j+1, j+2 live inside the (bw+1)-deep strip with wrap-around (torch.roll on
// the cut strip); only depth bw-1 wraps (to depth 0).
template <class F1, class F2>
ST_HD void w2_oneway_diffs(int s, int z, int x, const W2Geom& g, F1 h1, F2 h2,
                           float& base, float& dlam, float& dmu) {
    const int j = w2_depth(s, z, x, g);
    int z1, x1, z2, x2;
    int j1 = j + 1, j2 = j + 2;              // wrap inside the (bw+1)-deep strip
    if (j1 > g.bw) j1 -= g.bw + 1;
    if (j2 > g.bw) j2 -= g.bw + 1;
    w2_at_depth(s, j1, z, x, g, z1, x1);
    w2_at_depth(s, j2, z, x, g, z2, x2);
    const float a0 = h1(z, x), a1 = h1(z1, x1), a2 = h1(z2, x2);
    const float p0 = h2(z, x), p1 = h2(z1, x1);
    base = a0 + (a0 - p0);
    dlam = (a1 - a0) - (p1 - p0);
    dmu = (a1 - a0) - (a2 - a1);
}

template <class F1, class F2>
ST_HD float w2_habc_blend(float y, int z, int x, const W2Geom& g, float r, float b, F1 h1, F2 h2) {
    float f[4];
    w2_side_weights(z, x, g, f);
    const float lam = 2.f * r, mu = r * r;
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (f[s] != 0.f) {
            float base, dlam, dmu;
            w2_oneway_diffs(s, z, x, g, h1, h2, base, dlam, dmu);
            const float one = base + lam * dlam + mu * dmu;
            acc += f[s] * (one - y);
        }
    }
    return y + b * acc;
}

// ----------------------------------------------------------------- spatial operator
template <int FL, class F1>
ST_HD float w2_stencil(int z, int x, const W2Coef& c, float ciso, F1 u) {
    const float C = u(z, x), N = u(z - 1, x), S = u(z + 1, x), W = u(z, x - 1), E = u(z, x + 1);
    float A;
    if (FL & ST_F_ISO) A = ciso * (((N - C) + (S - C)) + ((E - C) + (W - C)));
    else A = c.cxx * ((E - C) + (W - C)) + c.czz * ((N - C) + (S - C));
    if (FL & ST_F_XZ) A += c.cxz * ((u(z + 1, x + 1) - u(z + 1, x - 1)) - (u(z - 1, x + 1) - u(z - 1, x - 1)));
    if (FL & ST_F_G1) A += c.ax * (E - W) + c.az * (S - N);
    return A;
}

// Forward update of all field channels of one cell.
//   H1(f,z,x), H2(f,z,x): current / previous field, zero outside the domain.
//   out[f]: new value (before source injection).
template <int FL, class FH1, class FH2>
ST_HD void w2_forward_cell(int z, int x, const W2Geom& g, const W2Coef& c, float dt,
                           FH1 H1, FH2 H2, float out[2]) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    const float alpha = (FL & ST_F_PML) ? c.czz : 1.f;
    const float ciso = c.cxx;
    float A0 = 0.f;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        auto h1 = [&](int zz, int xx) { return H1(f, zz, xx); };
        auto h2 = [&](int zz, int xx) { return H2(f, zz, xx); };
        const float u1 = h1(z, x), u2 = h2(z, x);
        float A = w2_stencil<FL>(z, x, c, ciso, h1);
        if (f == 0) A0 = A;
        else A += c.m * A0;
        float y = u1 + alpha * (u1 - u2) + A;
        if ((FL & ST_F_HABC) && w2_in_frame(z, x, g)) y = w2_habc_blend(y, z, x, g, c.r, c.b, h1, h2);
        out[f] = y;
    }
}

// ----------------------------------------------------------------- adjoint
// Computes, for cell p=(z,x), the cotangent Lam_i(f,p) of the field state S_i given
//   L1 = Lam_{i+1}, L2 = Lam_{i+2}   (zero outside the domain)
// and the contribution of forward step i+1 (which maps S_i, S_{i-1} -> S_{i+1}) to the
// coefficient gradients at p:   S1 = S_i, S2 = S_{i-1}.
//   CF(z,x) returns the W2Coef of an in-domain cell (used for the centre cell only);
//   CK(k,z,x) returns ONE coefficient plane value (k: 0 r, 1 b, 2 cxx, 3 czz, 4 cxz, 5 ax, 6 az, 7 m)
//   so neighbour cells load just what the transposed stencil needs.
// grad[] layout: 0:r (one-way blend only) 1:cxx (= ciso for ISO) 2:czz 3:cxz 4:ax 5:az 6:m   (accumulated, +=)
//
// Transpose algebra (DESIGN.md "adjoint"):  forward  Y = (1-b*M) y + b*sum_s f_s one_s,
//   y = h1 + alpha (h1-h2) + A[h1]  =>
//   Lam_i = (1+alpha) l1' + A^T[l1'] + Ha^T[L1] - alpha l2' + Hb^T[L2],   l' = (1-b*M) L.
template <int FL, class FL1, class FL2, class FS1, class FS2, class FC, class FK>
ST_HD void w2_adjoint_cell(int z, int x, const W2Geom& g, float dt,
                           FL1 L1, FL2 L2, FS1 S1, FS2 S2, FC CF, FK CK,
                           float out[2], float grad[7], bool want_grad) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    const bool habc = (FL & ST_F_HABC) != 0;
    auto inside = [&](int zz, int xx) { return zz >= 0 && zz < g.nz && xx >= 0 && xx < g.nx; };
    // pre-blend cotangent factor (1 - b*M) at q
    // Note: inside the frame sum_s f_s == 1 (habc.py masks tile the frame; checked by
    // tests/test_hostcheck.py), so M == in_frame.
    auto preq = [&](int zz, int xx) {
        if (!habc || !w2_in_frame(zz, xx, g)) return 1.f;
        return 1.f - CK(1, zz, xx);
    };
    // effective cotangent that multiplies the spatial operator of field f at q
    auto leffq = [&](int f, int zz, int xx) {
        const float pq = preq(zz, xx);
        float v = pq * L1(f, zz, xx);
        if ((FL & ST_F_BORN) && f == 0) v += CK(7, zz, xx) * (pq * L1(1, zz, xx));
        return v;
    };
    const W2Coef cp = CF(z, x);
    const float alpha = (FL & ST_F_PML) ? cp.czz : 1.f;
    const float prep = (habc && w2_in_frame(z, x, g)) ? 1.f - cp.b : 1.f;

#pragma unroll
    for (int f = 0; f < NF; ++f) {
        // ---- pointwise terms
        float acc = (1.f + alpha) * (prep * L1(f, z, x)) - alpha * (prep * L2(f, z, x));
        // ---- transposed spatial operator: stencil applied to the products C(q)*leff(q)
        auto wk = [&](int kind, int zz, int xx) -> float {
            if (!inside(zz, xx)) return 0.f;
            // kind 0: cxx (= ciso for ISO), 1: czz, 2: cxz, 3: ax, 4: az  -> coefficient planes 2..6
            return CK(2 + kind, zz, xx) * leffq(f, zz, xx);
        };
        if (FL & ST_F_ISO) {
            const float wc = wk(0, z, x);
            acc += ((wk(0, z - 1, x) - wc) + (wk(0, z + 1, x) - wc)) + ((wk(0, z, x - 1) - wc) + (wk(0, z, x + 1) - wc));
        } else {
            const float wx = wk(0, z, x), wz = wk(1, z, x);
            acc += ((wk(0, z, x - 1) - wx) + (wk(0, z, x + 1) - wx)) + ((wk(1, z - 1, x) - wz) + (wk(1, z + 1, x) - wz));
        }
        if (FL & ST_F_XZ)
            acc += (wk(2, z - 1, x - 1) - wk(2, z - 1, x + 1)) - (wk(2, z + 1, x - 1) - wk(2, z + 1, x + 1));
        if (FL & ST_F_G1)
            acc += (wk(3, z, x - 1) - wk(3, z, x + 1)) + (wk(4, z - 1, x) - wk(4, z + 1, x));
        // ---- transposed one-way blend: gather over the cells q whose extrapolation reads p
        const int straight = habc ? w2_straight_side(z, x, g) : -1;
        if (habc && straight >= 0) {
            // straight side: the three cells outward of p (and the wrap partner of depth 0) with weight b(q)
            const int jp = w2_depth(straight, z, x, g);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int jq, kk;
                if (k < 3) { jq = jp - k; kk = k; if (jq < 0 || jq > g.bw - 1 || jq + kk > g.bw) continue; }
                else { if (jp != 0) continue; jq = g.bw - 1; kk = 2; }
                int zq, xq;
                w2_at_depth(straight, jq, z, x, g, zq, xq);
                const float wgt = CK(1, zq, xq);
                if (wgt == 0.f) continue;
                const float rq = CK(0, zq, xq);
                const float lam = 2.f * rq, mu = rq * rq;
                const float a = kk == 0 ? (2.f - lam - mu) : (kk == 1 ? (lam + 2.f * mu) : -mu);
                acc += wgt * a * L1(f, zq, xq);
                if (kk <= 1) {
                    const float bt = kk == 0 ? (lam - 1.f) : -lam;
                    acc += wgt * bt * L2(f, zq, xq);
                }
            }
        } else if (habc) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int jp = w2_depth(s, z, x, g);
                if (jp > g.bw) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k) {       // k==3 encodes the wrap (depth bw-1 reads depth 0 as j+2)
                    int jq, kk;
                    if (k < 3) { jq = jp - k; kk = k; if (jq < 0 || jq > g.bw - 1 || jq + kk > g.bw) continue; }
                    else { if (jp != 0) continue; jq = g.bw - 1; kk = 2; }
                    int zq, xq;
                    w2_at_depth(s, jq, z, x, g, zq, xq);
                    if (!inside(zq, xq)) continue;
                    float fq[4];
                    w2_side_weights(zq, xq, g, fq);
                    if (fq[s] == 0.f) continue;
                    const float rq = CK(0, zq, xq);
                    const float lam = 2.f * rq, mu = rq * rq;
                    const float wgt = CK(1, zq, xq) * fq[s];
                    const float a = kk == 0 ? (2.f - lam - mu) : (kk == 1 ? (lam + 2.f * mu) : -mu);
                    acc += wgt * a * L1(f, zq, xq);
                    if (kk <= 1) {
                        const float bt = kk == 0 ? (lam - 1.f) : -lam;
                        acc += wgt * bt * L2(f, zq, xq);
                    }
                }
            }
        }
        out[f] = acc;
    }

    if (!want_grad) return;
    // ---- coefficient gradients of forward step i+1 at p
    {
        const float ci = cp.cxx;
        float A0 = 0.f;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            auto s1 = [&](int zz, int xx) { return S1(f, zz, xx); };
            auto s2 = [&](int zz, int xx) { return S2(f, zz, xx); };
            const float lraw = L1(f, z, x);
            const float lpre = prep * lraw;
            float le = lpre;
            if ((FL & ST_F_BORN) && f == 0) le += cp.m * (prep * L1(1, z, x));
            const float C = s1(z, x), N = s1(z - 1, x), S = s1(z + 1, x), W = s1(z, x - 1), E = s1(z, x + 1);
            if (FL & ST_F_ISO) {
                grad[1] += le * (((N - C) + (S - C)) + ((E - C) + (W - C)));       // d/d ciso
            } else {
                grad[1] += le * ((E - C) + (W - C));
                grad[2] += le * ((N - C) + (S - C));
            }
            float cross = 0.f;
            if (FL & ST_F_XZ) {
                cross = (s1(z + 1, x + 1) - s1(z + 1, x - 1)) - (s1(z - 1, x + 1) - s1(z - 1, x - 1));
                grad[3] += le * cross;
            }
            if (FL & ST_F_G1) {
                grad[4] += le * (E - W);
                grad[5] += le * (S - N);
            }
            if (FL & ST_F_BORN) {
                if (f == 0) A0 = w2_stencil<FL>(z, x, cp, ci, s1);
                else grad[6] += lpre * A0;
            }
            if (habc && w2_in_frame(z, x, g)) {
                float fq[4];
                w2_side_weights(z, x, g, fq);
                float dsum = 0.f;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (fq[s] != 0.f) {
                        float base, dlam, dmu;
                        w2_oneway_diffs(s, z, x, g, s1, s2, base, dlam, dmu);
                        dsum += fq[s] * (2.f * dlam + 2.f * cp.r * dmu);
                    }
                }
                grad[0] += lraw * cp.b * dsum;
            }
        }
    }
}
