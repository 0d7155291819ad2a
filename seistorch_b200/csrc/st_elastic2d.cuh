// 2D elastic velocity-stress: launch arguments and per-cell arithmetic shared by the
// sm_100a kernels (st_elastic2d.cu) and the CPU host-check build.
//
// Reference restated: equations2d/elastic.py:7-37 with equations2d/utils.py:3-48
//   forward_diff  D+ : x[i]-x[i-1], 0 at i=0        backward_diff D- : x[i+1]-x[i], 0 at last
//   vx_x = D+x vx   vz_z = D-z vz   vx_z = D+z vx   vz_x = D-x vz          (elastic.py:15-18)
//   txx_x = D-x txx'  txz_z = D-z txz'  tzz_z = D+z tzz'  txz_x = D+x txz'  (elastic.py:27-30)
// Coefficients (precomputed per call, see include/seistorch_b200.h):
//   ca=(1-c)/(1+c)  cl2m=(lambda+2mu)dt/h/(1+c)  cl=lambda dt/h/(1+c)  cm=mu dt/h/(1+c)  cb=dt/(rho h)/(1+c)
#pragma once
#include "st_common.cuh"

struct E2Args {
    int nz, nx, ld, B;
    long long fs, cs;           // plane / channel strides (floats)
    const float* coef[5];       // ca, cl2m, cl, cm, cb
    const float* cur;           // forward: S_{i-1}  [5][B][nz][ld]   (vx,vz,txx,tzz,txz)
    float* next;                // forward: S_i
    const float* lam1;          // adjoint: Lam_{i+1} (nullptr == 0)
    float* lam0;                // adjoint: Lam_i
    const float* s0;            // adjoint: S_i
    const float* s1;            // adjoint: S_{i+1}
    float* gacc;                // [nchunk][4][nz*ld]
    int bchunk;
    int ns; const int* src_b; const int* src_z; const int* src_x;
    const float* amp; float* gamp; int src_fmask;
    const int* row_start; const int* rec_x; const int* rec_orig;
    int R; int nchan; int chan_f[4];
    float* rec_out; const float* rec_adj;
};

struct E2Coef { float ca, cl2m, cl, cm, cb; };

// new stresses at q=(z,x) from the old state.  V(f,z,x): f=0 vx, 1 vz (in-domain reads only)
template <class FV>
ST_HD void e2_stress_cell(int z, int x, int nz, int nx, const E2Coef& c, FV V,
                          float txx, float tzz, float txz, float out[3]) {
    const float vx_x = x > 0 ? V(0, z, x) - V(0, z, x - 1) : 0.f;
    const float vz_z = z < nz - 1 ? V(1, z + 1, x) - V(1, z, x) : 0.f;
    const float vx_z = z > 0 ? V(0, z, x) - V(0, z - 1, x) : 0.f;
    const float vz_x = x < nx - 1 ? V(1, z, x + 1) - V(1, z, x) : 0.f;
    out[0] = c.ca * txx + (c.cl2m * vx_x + c.cl * vz_z);
    out[1] = c.ca * tzz + (c.cl2m * vz_z + c.cl * vx_x);
    out[2] = c.ca * txz + c.cm * (vz_x + vx_z);
}

// divergence of the NEW stresses at p.  T(f,z,x): f=0 txx', 1 tzz', 2 txz'
template <class FT>
ST_HD void e2_stress_div(int z, int x, int nz, int nx, FT T, float& fx, float& fz) {
    const float txx_x = x < nx - 1 ? T(0, z, x + 1) - T(0, z, x) : 0.f;
    const float txz_z = z < nz - 1 ? T(2, z + 1, x) - T(2, z, x) : 0.f;
    const float tzz_z = z > 0 ? T(1, z, x) - T(1, z - 1, x) : 0.f;
    const float txz_x = x > 0 ? T(2, z, x) - T(2, z, x - 1) : 0.f;
    fx = txx_x + txz_z;
    fz = txz_x + tzz_z;
}

// ---- adjoint, stage A: total cotangent of the new stresses at q.
//   W(f,z,x) = cb*lam_v (f=0: vx, 1: vz), in-domain reads only.
template <class FW>
ST_HD void e2_adj_stress_tot(int z, int x, int nz, int nx, FW W, float ltxx, float ltzz, float ltxz, float out[3]) {
    float a = ltxx, b = ltzz, e = ltxz;
    if (x >= 1) a += W(0, z, x - 1);          // vx'(q-x) reads txx'(q)  (x-1 < nx-1 always)
    if (x < nx - 1) a -= W(0, z, x);
    if (z >= 1) e += W(0, z - 1, x);          // vx'(q-z) reads txz'(q) via txz_z
    if (z < nz - 1) e -= W(0, z, x);
    if (x > 0) e += W(1, z, x);               // vz'(q) reads txz'(q) via txz_x
    if (x + 1 <= nx - 1) e -= W(1, z, x + 1);
    if (z > 0) b += W(1, z, x);               // vz'(q) reads tzz'(q) via tzz_z
    if (z + 1 <= nz - 1) b -= W(1, z + 1, x);
    out[0] = a; out[1] = b; out[2] = e;
}

// ---- adjoint, stage B: cotangent of the old velocities at p.
//   G(k,z,x): k=0 a=cl2m*Ltxx+cl*Ltzz, k=1 b=cl*Ltxx+cl2m*Ltzz, k=2 e=cm*Ltxz  (L = stage-A totals)
template <class FG>
ST_HD void e2_adj_velocity(int z, int x, int nz, int nx, FG G, float ca, float lvx, float lvz, float& ovx, float& ovz) {
    float ax = ca * lvx, az = ca * lvz;
    if (x > 0) ax += G(0, z, x);
    if (x + 1 <= nx - 1) ax -= G(0, z, x + 1);
    if (z > 0) ax += G(2, z, x);
    if (z + 1 <= nz - 1) ax -= G(2, z + 1, x);
    if (z - 1 >= 0) az += G(1, z - 1, x);
    if (z < nz - 1) az -= G(1, z, x);
    if (x - 1 >= 0) az += G(2, z, x - 1);
    if (x < nx - 1) az -= G(2, z, x);
    ovx = ax; ovz = az;
}

#ifdef __CUDACC__
int st_elastic2d_launch_forward(const E2Args& a, cudaStream_t st);
int st_elastic2d_launch_adjoint(const E2Args& a, cudaStream_t st);
#endif
