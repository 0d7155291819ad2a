// TEST INFRASTRUCTURE ONLY.  Host (CPU) emulation of the per-cell kernel arithmetic in
// seistorch_b200/csrc/*_math.cuh / st_elastic2d.cuh, compiled with g++ for the
// `-m "not gpu"` tests: it runs the very same __host__ __device__ functions the sm_100a
// kernels call, cell by cell, so the forward/adjoint algebra is checked against the
// oracle on CPU.  It is never linked into the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "st_wave2d_math.cuh"
#include "st_elastic2d.cuh"

struct HcW2 {
    int flags, B, nz, nx, ld, bw, multiple;
    float dt;
    const float* coef[8];
};

static inline W2Coef hc_coef(const HcW2& p, long long idx) {
    W2Coef c;
    c.r = p.coef[0][idx]; c.b = p.coef[1][idx];
    c.cxx = p.coef[2] ? p.coef[2][idx] : 0.f; c.czz = p.coef[3] ? p.coef[3][idx] : 0.f;
    c.cxz = p.coef[4] ? p.coef[4][idx] : 0.f; c.ax = p.coef[5] ? p.coef[5][idx] : 0.f;
    c.az = p.coef[6] ? p.coef[6][idx] : 0.f; c.m = p.coef[7] ? p.coef[7][idx] : 0.f;
    return c;
}

template <int FL>
static void w2_fwd(const HcW2& p, const float* prev, const float* cur, float* next) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    W2Geom g{p.nz, p.nx, p.ld, p.bw, p.multiple};
    const long long fs = (long long)p.nz * p.ld, cs = fs * p.B;
    for (int b = 0; b < p.B; ++b)
        for (int z = 0; z < p.nz; ++z)
            for (int x = 0; x < p.nx; ++x) {
                auto inb = [&](int zz, int xx) { return zz >= 0 && zz < p.nz && xx >= 0 && xx < p.nx; };
                auto H1 = [&](int f, int zz, int xx) -> float { return inb(zz, xx) ? cur[f * cs + b * fs + (long long)zz * p.ld + xx] : 0.f; };
                auto H2 = [&](int f, int zz, int xx) -> float { return inb(zz, xx) ? prev[f * cs + b * fs + (long long)zz * p.ld + xx] : 0.f; };
                float out[2];
                w2_forward_cell<FL>(z, x, g, hc_coef(p, (long long)z * p.ld + x), p.dt, H1, H2, out);
                for (int f = 0; f < NF; ++f) next[f * cs + b * fs + (long long)z * p.ld + x] = out[f];
            }
}

template <int FL>
static void w2_adj(const HcW2& p, const float* l1, const float* l2, const float* s1, const float* s2, float* l0, float* gacc) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    W2Geom g{p.nz, p.nx, p.ld, p.bw, p.multiple};
    const long long fs = (long long)p.nz * p.ld, cs = fs * p.B;
    for (int b = 0; b < p.B; ++b)
        for (int z = 0; z < p.nz; ++z)
            for (int x = 0; x < p.nx; ++x) {
                auto inb = [&](int zz, int xx) { return zz >= 0 && zz < p.nz && xx >= 0 && xx < p.nx; };
                auto mk = [&](const float* a) {
                    return [=](int f, int zz, int xx) -> float { return inb(zz, xx) ? a[f * cs + b * fs + (long long)zz * p.ld + xx] : 0.f; };
                };
                auto CF = [&](int zz, int xx) { return hc_coef(p, (long long)zz * p.ld + xx); };
                auto CK = [&](int k, int zz, int xx) -> float { return p.coef[k] ? p.coef[k][(long long)zz * p.ld + xx] : 0.f; };
                float out[2], gr[7] = {0, 0, 0, 0, 0, 0, 0};
                w2_adjoint_cell<FL>(z, x, g, p.dt, mk(l1), mk(l2), mk(s1), mk(s2), CF, CK, out, gr, gacc != nullptr);
                for (int f = 0; f < NF; ++f) l0[f * cs + b * fs + (long long)z * p.ld + x] = out[f];
                if (gacc)
                    for (int q = 0; q < 7; ++q) gacc[q * fs + (long long)z * p.ld + x] += gr[q];
            }
}

#define HC_DISPATCH(FN, ...)                                                                   \
    switch (p->flags) {                                                                         \
        case ST_F_ISO | ST_F_PML: FN<ST_F_ISO | ST_F_PML>(__VA_ARGS__); break;                  \
        case ST_F_ISO | ST_F_HABC: FN<ST_F_ISO | ST_F_HABC>(__VA_ARGS__); break;                \
        case ST_F_HABC: FN<ST_F_HABC>(__VA_ARGS__); break;                                      \
        case ST_F_HABC | ST_F_XZ: FN<ST_F_HABC | ST_F_XZ>(__VA_ARGS__); break;                  \
        case ST_F_HABC | ST_F_G1: FN<ST_F_HABC | ST_F_G1>(__VA_ARGS__); break;                  \
        case ST_F_ISO | ST_F_HABC | ST_F_G1: FN<ST_F_ISO | ST_F_HABC | ST_F_G1>(__VA_ARGS__); break; \
        case ST_F_HABC | ST_F_BORN: FN<ST_F_HABC | ST_F_BORN>(__VA_ARGS__); break;              \
        case ST_F_HABC | ST_F_XZ | ST_F_BORN: FN<ST_F_HABC | ST_F_XZ | ST_F_BORN>(__VA_ARGS__); break; \
        default: return -2;                                                                     \
    }

extern "C" int hc_wave2d_forward(const HcW2* p, const float* prev, const float* cur, float* next) {
    HC_DISPATCH(w2_fwd, *p, prev, cur, next)
    return 0;
}
extern "C" int hc_wave2d_adjoint(const HcW2* p, const float* l1, const float* l2, const float* s1, const float* s2,
                                 float* l0, float* gacc) {
    HC_DISPATCH(w2_adj, *p, l1, l2, s1, s2, l0, gacc)
    return 0;
}
extern "C" void hc_side_weights(int nz, int nx, int bw, int multiple, float* f /*[4][nz][nx]*/, uint8_t* frame) {
    W2Geom g{nz, nx, nx, bw, multiple};
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) {
            float w[4];
            w2_side_weights(z, x, g, w);
            for (int s = 0; s < 4; ++s) f[((long long)s * nz + z) * nx + x] = w[s];
            frame[(long long)z * nx + x] = w2_in_frame(z, x, g) ? 1 : 0;
        }
}

// ---------------------------------------------------------------- elastic
struct HcE2 { int B, nz, nx, ld; const float* coef[5]; };

extern "C" int hc_elastic2d_forward(const HcE2* p, const float* cur, float* next) {
    const int nz = p->nz, nx = p->nx, ld = p->ld;
    const long long fs = (long long)nz * ld, cs = fs * p->B;
    for (int b = 0; b < p->B; ++b) {
        const float* c0 = cur + b * fs;
        float* n0 = next + b * fs;
        auto V = [&](int f, int z, int x) -> float { return c0[f * cs + (long long)z * ld + x]; };
        for (int z = 0; z < nz; ++z)
            for (int x = 0; x < nx; ++x) {
                const long long idx = (long long)z * ld + x;
                E2Coef c{p->coef[0][idx], p->coef[1][idx], p->coef[2][idx], p->coef[3][idx], p->coef[4][idx]};
                float t[3];
                e2_stress_cell(z, x, nz, nx, c, V, c0[2 * cs + idx], c0[3 * cs + idx], c0[4 * cs + idx], t);
                n0[2 * cs + idx] = t[0]; n0[3 * cs + idx] = t[1]; n0[4 * cs + idx] = t[2];
            }
        auto T = [&](int f, int z, int x) -> float { return n0[(2 + f) * cs + (long long)z * ld + x]; };
        for (int z = 0; z < nz; ++z)
            for (int x = 0; x < nx; ++x) {
                const long long idx = (long long)z * ld + x;
                float fx, fz;
                e2_stress_div(z, x, nz, nx, T, fx, fz);
                n0[idx] = p->coef[0][idx] * c0[idx] + p->coef[4][idx] * fx;
                n0[cs + idx] = p->coef[0][idx] * c0[cs + idx] + p->coef[4][idx] * fz;
            }
    }
    return 0;
}

// Lam_i from Lam_{i+1}; s0 = S_i, s1pre = new state of step i+1 BEFORE the source add.
extern "C" int hc_elastic2d_adjoint(const HcE2* p, const float* l1, const float* s0, const float* s1pre, float* l0, float* gacc) {
    const int nz = p->nz, nx = p->nx, ld = p->ld;
    const long long fs = (long long)nz * ld, cs = fs * p->B;
    std::vector<float> Lt(3 * fs), G(3 * fs);
    for (int b = 0; b < p->B; ++b) {
        const float* L = l1 + b * fs;
        float* O = l0 + b * fs;
        auto W = [&](int f, int z, int x) -> float { const long long idx = (long long)z * ld + x; return p->coef[4][idx] * L[f * cs + idx]; };
        for (int z = 0; z < nz; ++z)
            for (int x = 0; x < nx; ++x) {
                const long long idx = (long long)z * ld + x;
                float t[3];
                e2_adj_stress_tot(z, x, nz, nx, W, L[2 * cs + idx], L[3 * cs + idx], L[4 * cs + idx], t);
                for (int k = 0; k < 3; ++k) Lt[k * fs + idx] = t[k];
                G[0 * fs + idx] = p->coef[1][idx] * t[0] + p->coef[2][idx] * t[1];
                G[1 * fs + idx] = p->coef[2][idx] * t[0] + p->coef[1][idx] * t[1];
                G[2 * fs + idx] = p->coef[3][idx] * t[2];
                for (int k = 0; k < 3; ++k) O[(2 + k) * cs + idx] = p->coef[0][idx] * t[k];
            }
        auto Gf = [&](int k, int z, int x) -> float { return G[k * fs + (long long)z * ld + x]; };
        for (int z = 0; z < nz; ++z)
            for (int x = 0; x < nx; ++x) {
                const long long idx = (long long)z * ld + x;
                float ovx, ovz;
                e2_adj_velocity(z, x, nz, nx, Gf, p->coef[0][idx], L[idx], L[cs + idx], ovx, ovz);
                O[idx] = ovx; O[cs + idx] = ovz;
                if (gacc) {
                    auto V = [&](int f, int zz, int xx) -> float { return s0[b * fs + f * cs + (long long)zz * ld + xx]; };
                    const float vx_x = x > 0 ? V(0, z, x) - V(0, z, x - 1) : 0.f;
                    const float vz_z = z < nz - 1 ? V(1, z + 1, x) - V(1, z, x) : 0.f;
                    const float vx_z = z > 0 ? V(0, z, x) - V(0, z - 1, x) : 0.f;
                    const float vz_x = x < nx - 1 ? V(1, z, x + 1) - V(1, z, x) : 0.f;
                    gacc[0 * fs + idx] += Lt[idx] * vx_x + Lt[fs + idx] * vz_z;
                    gacc[1 * fs + idx] += Lt[idx] * vz_z + Lt[fs + idx] * vx_x;
                    gacc[2 * fs + idx] += Lt[2 * fs + idx] * (vz_x + vx_z);
                    auto Tn = [&](int f, int zz, int xx) -> float { return s1pre[b * fs + (2 + f) * cs + (long long)zz * ld + xx]; };
                    float fx, fz;
                    e2_stress_div(z, x, nz, nx, Tn, fx, fz);
                    gacc[3 * fs + idx] += L[idx] * fx + L[cs + idx] * fz;
                }
            }
    }
    return 0;
}
