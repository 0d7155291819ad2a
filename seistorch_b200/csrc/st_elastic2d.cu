// sm_100a kernels for the 2D elastic velocity-stress equation (one launch per time step,
// stresses then velocities fused through shared memory, source add + receiver gather
// fused in).  Reference: equations2d/elastic.py:7-37 (~100 ATen launches per step).
//
// Layout: state [5][B][nz][ld] = (vx, vz, txx, tzz, txz); coefficient planes [nz][ld].
#include "st_elastic2d.cuh"

namespace {

constexpr int TX = 64, TZ = 32;
constexpr int NTX = 64, NTY = 4, NT = NTX * NTY, RPT = TZ / NTY;
constexpr int VW = TX + 4, VH = TZ + 4;     // velocity tiles, halo 2
constexpr int SW = TX + 2, SH = TZ + 2;     // stress tiles, halo 1

__device__ __forceinline__ E2Coef load_ecoef(const E2Args& a, long long idx) {
    E2Coef c;
    c.ca = __ldg(a.coef[0] + idx); c.cl2m = __ldg(a.coef[1] + idx); c.cl = __ldg(a.coef[2] + idx);
    c.cm = __ldg(a.coef[3] + idx); c.cb = __ldg(a.coef[4] + idx);
    return c;
}

// source add + receiver gather of the cells [z0,zn) x [x0,xn) of shot b (after the block stored them)
__device__ __forceinline__ void elastic_forward_tail(const E2Args& a, int b, int z0, int zn, int x0, int xn, int tid) {
    const int nz = a.nz, ld = a.ld;
    float* nxt = a.next + (long long)b * a.fs;
    __syncthreads();
    for (int s = tid; s < a.ns; s += NT) {
        if (a.src_b[s] != b) continue;
        const int sz = a.src_z[s], sx = a.src_x[s];
        if (sz >= z0 && sz < zn && sx >= x0 && sx < xn) {
            const float v = a.amp[s];
#pragma unroll
            for (int f = 0; f < 5; ++f)
                if (a.src_fmask >> f & 1) atomicAdd(nxt + f * a.cs + (long long)sz * ld + sx, v);
        }
    }
    if (!a.rec_out) return;
    __shared__ int s_cnt, s_rows[64];
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (tid < zn - z0 && z0 + tid < nz) {
        const int row = b * nz + z0 + tid;
        if (a.row_start[row + 1] > a.row_start[row]) s_rows[atomicAdd(&s_cnt, 1)] = z0 + tid;
    }
    __syncthreads();
    const int cnt = s_cnt;
    for (int i = 0; i < cnt; ++i) {
        const int z = s_rows[i];
        const int lo = a.row_start[b * nz + z], hi = a.row_start[b * nz + z + 1];
        for (int r = lo + tid; r < hi; r += NT) {
            const int rx = a.rec_x[r];
            if (rx >= x0 && rx < xn) {
                const long long o = (long long)a.rec_orig[r] * a.nchan;
                for (int ch = 0; ch < a.nchan; ++ch)
                    a.rec_out[o + ch] = nxt[a.chan_f[ch] * a.cs + (long long)z * ld + rx];
            }
        }
    }
}

// ---- border tiles: shared-memory version (handles the zeroed differences at the domain edges)
__device__ __forceinline__ void elastic_forward_tile(const E2Args& a, int tx, int tz, int b,
                                                     float (*sv)[VH][VW], float (*st)[SH][SW]) {
    const int nz = a.nz, nx = a.nx, ld = a.ld;
    const int tid = threadIdx.x;
    const int tidx = tid & (NTX - 1), tidy = tid / NTX;
    const int x0 = tx * TX, z0 = tz * TZ;
    const long long boff = (long long)b * a.fs;
    const float* cur = a.cur + boff;
    float* nxt = a.next + boff;

    // tile completely inside the domain (incl. its 2-cell halo): no bounds predicates needed
    const bool inner = z0 >= 2 && z0 + TZ + 2 <= nz && x0 >= 2 && x0 + TX + 2 <= nx;
    for (int lz = tidy; lz < VH; lz += NTY) {
        const int z = z0 - 2 + lz;
        for (int lx = tidx; lx < VW; lx += NTX) {
            const int x = x0 - 2 + lx;
            float v0 = 0.f, v1 = 0.f;
            if (inner || (z >= 0 && z < nz && x >= 0 && x < nx)) {
                const int idx = z * ld + x;
                v0 = __ldg(cur + idx);
                v1 = __ldg(cur + a.cs + idx);
            }
            sv[0][lz][lx] = v0;
            sv[1][lz][lx] = v1;
        }
    }
    __syncthreads();
    auto V = [&](int f, int zz, int xx) -> float { return sv[f][zz - z0 + 2][xx - x0 + 2]; };
    for (int lz = tidy; lz < SH; lz += NTY) {
        const int z = z0 - 1 + lz;
        for (int lx = tidx; lx < SW; lx += NTX) {
            const int x = x0 - 1 + lx;
            float t[3] = {0.f, 0.f, 0.f};
            if (inner || (z >= 0 && z < nz && x >= 0 && x < nx)) {
                const int idx = z * ld + x;
                const E2Coef c = load_ecoef(a, idx);
                e2_stress_cell(z, x, nz, nx, c, V, __ldg(cur + 2 * a.cs + idx), __ldg(cur + 3 * a.cs + idx),
                               __ldg(cur + 4 * a.cs + idx), t);
                if (lz >= 1 && lz <= TZ && lx >= 1 && lx <= TX) {
                    nxt[2 * a.cs + idx] = t[0];
                    nxt[3 * a.cs + idx] = t[1];
                    nxt[4 * a.cs + idx] = t[2];
                }
            }
            st[0][lz][lx] = t[0]; st[1][lz][lx] = t[1]; st[2][lz][lx] = t[2];
        }
    }
    __syncthreads();
    auto T = [&](int f, int zz, int xx) -> float { return st[f][zz - z0 + 1][xx - x0 + 1]; };
    const int x = x0 + tidx;
    if (x < nx) {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int z = z0 + tidy + k * NTY;
            if (z >= nz) break;
            const long long idx = (long long)z * ld + x;
            float fx, fz;
            e2_stress_div(z, x, nz, nx, T, fx, fz);
            const float ca = __ldg(a.coef[0] + idx), cb = __ldg(a.coef[4] + idx);
            nxt[idx] = ca * V(0, z, x) + cb * fx;
            nxt[a.cs + idx] = ca * V(1, z, x) + cb * fz;
        }
    }
    elastic_forward_tail(a, b, z0, z0 + TZ, x0, x0 + TX, tid);
}

// ---- interior tiles: register/shuffle version.  A warp owns 128 columns x FRZ rows, every lane 4
// consecutive cells; the new stresses are computed one row ahead of the velocities (which need
// txz'(z+1) and tzz'(z-1)), everything stays in registers; x-neighbours come from warp shuffles,
// the stress halos of the warp's edge cells (txx' right of lane 31, txz' left of lane 0) are
// recomputed by those lanes from a few scalar loads.
#ifndef ST_EL_FRZ
#define ST_EL_FRZ 8
#endif
constexpr int FW = 128, FRZ = ST_EL_FRZ, FH = FRZ * (NT / 32);       // 128 x 64 fast tile = 2 x 2 border tiles

__device__ __forceinline__ float f4g(const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4s(float4& v, int e, float s) { if (e == 0) v.x = s; else if (e == 1) v.y = s; else if (e == 2) v.z = s; else v.w = s; }

struct StressRow { float4 xx, zz, xz; float xx_r, xz_l; };

__device__ __forceinline__ void elastic_forward_fast(const E2Args& a, int fx, int fz, int b, int tid) {
    const int ld = a.ld;
    const int warp = tid >> 5, lane = tid & 31;
    const int x0 = fx * FW, zb0 = fz * FH, z0 = zb0 + warp * FRZ;
    const int x = x0 + 4 * lane;
    const long long boff = (long long)b * a.fs, cs = a.cs;
    const float* VX = a.cur + boff;
    const float* VZ = VX + cs;
    const float* TXX = VX + 2 * cs;
    const float* TZZ = VX + 3 * cs;
    const float* TXZ = VX + 4 * cs;
    float* nxt = a.next + boff;
    const float* CA = a.coef[0];
    const float* C2 = a.coef[1];
    const float* CL = a.coef[2];
    const float* CM = a.coef[3];
    const float* CB = a.coef[4];
    const bool e0 = lane == 0, e31 = lane == 31;
    auto L4 = [&](const float* p, int r) { return __ldg(reinterpret_cast<const float4*>(p + (r * ld + x))); };
    auto L1 = [&](const float* p, int r, int xx) { return __ldg(p + (r * ld + xx)); };

    // new stresses of row r from vx(r-1), vx(r), vz(r), vz(r+1)
    auto stress = [&](int r, const float4& vxu, const float4& vxr, const float4& vzr, const float4& vzd) {
        StressRow s;
        const float4 ca = L4(CA, r), c2 = L4(C2, r), cl = L4(CL, r), cm = L4(CM, r);
        const float4 oxx = L4(TXX, r), ozz = L4(TZZ, r), oxz = L4(TXZ, r);
        float vx_l = __shfl_up_sync(0xffffffffu, vxr.w, 1);
        float vz_r = __shfl_down_sync(0xffffffffu, vzr.x, 1);
        s.xx_r = 0.f; s.xz_l = 0.f;
        if (e0) {
            const int xl = x0 - 1;
            vx_l = L1(VX, r, xl);
            // txz'(r, x0-1): vz_x = vz(r,x0) - vz(r,x0-1), vx_z = vx(r,x0-1) - vx(r-1,x0-1)
            const float vz_x = vzr.x - L1(VZ, r, xl), vx_z = vx_l - L1(VX, r - 1, xl);
            s.xz_l = L1(CA, r, xl) * L1(TXZ, r, xl) + L1(CM, r, xl) * (vz_x + vx_z);
        }
        if (e31) {
            const int xr = x0 + FW;
            vz_r = L1(VZ, r, xr);
            // txx'(r, x0+128): vx_x = vx(r,x0+128) - vx(r,x0+127), vz_z = vz(r+1,x0+128) - vz(r,x0+128)
            const float vx_x = L1(VX, r, xr) - vxr.w, vz_z = L1(VZ, r + 1, xr) - vz_r;
            s.xx_r = L1(CA, r, xr) * L1(TXX, r, xr) + (L1(C2, r, xr) * vx_x + L1(CL, r, xr) * vz_z);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float vx_x = f4g(vxr, e) - (e == 0 ? vx_l : f4g(vxr, e - 1));
            const float vz_z = f4g(vzd, e) - f4g(vzr, e);
            const float vx_z = f4g(vxr, e) - f4g(vxu, e);
            const float vz_x = (e == 3 ? vz_r : f4g(vzr, e + 1)) - f4g(vzr, e);
            f4s(s.xx, e, f4g(ca, e) * f4g(oxx, e) + (f4g(c2, e) * vx_x + f4g(cl, e) * vz_z));
            f4s(s.zz, e, f4g(ca, e) * f4g(ozz, e) + (f4g(c2, e) * vz_z + f4g(cl, e) * vx_x));
            f4s(s.xz, e, f4g(ca, e) * f4g(oxz, e) + f4g(cm, e) * (vz_x + vx_z));
        }
        return s;
    };

    float4 vxm = L4(VX, z0 - 2), vx0 = L4(VX, z0 - 1), vz0 = L4(VZ, z0 - 1), vz1 = L4(VZ, z0);
    float4 tzz_prev = stress(z0 - 1, vxm, vx0, vz0, vz1).zz;
    vxm = vx0; vx0 = L4(VX, z0); vz0 = vz1; vz1 = L4(VZ, z0 + 1);
    StressRow sc = stress(z0, vxm, vx0, vz0, vz1);
#pragma unroll 2
    for (int k = 0; k < FRZ; ++k) {
        const int z = z0 + k, ro = z * ld + x;
        const float4 vx_old = vx0, vz_old = vz0;
        vxm = vx0; vx0 = L4(VX, z + 1); vz0 = vz1; vz1 = L4(VZ, z + 2);
        const StressRow sn = stress(z + 1, vxm, vx0, vz0, vz1);
        *reinterpret_cast<float4*>(nxt + 2 * cs + ro) = sc.xx;
        *reinterpret_cast<float4*>(nxt + 3 * cs + ro) = sc.zz;
        *reinterpret_cast<float4*>(nxt + 4 * cs + ro) = sc.xz;
        float txx_r = __shfl_down_sync(0xffffffffu, sc.xx.x, 1);
        float txz_l = __shfl_up_sync(0xffffffffu, sc.xz.w, 1);
        txx_r = e31 ? sc.xx_r : txx_r;
        txz_l = e0 ? sc.xz_l : txz_l;
        const float4 ca = L4(CA, z), cb = L4(CB, z);
        float4 nvx, nvz;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float txx_x = (e == 3 ? txx_r : f4g(sc.xx, e + 1)) - f4g(sc.xx, e);
            const float txz_z = f4g(sn.xz, e) - f4g(sc.xz, e);
            const float tzz_z = f4g(sc.zz, e) - f4g(tzz_prev, e);
            const float txz_x = f4g(sc.xz, e) - (e == 0 ? txz_l : f4g(sc.xz, e - 1));
            f4s(nvx, e, f4g(ca, e) * f4g(vx_old, e) + f4g(cb, e) * (txx_x + txz_z));
            f4s(nvz, e, f4g(ca, e) * f4g(vz_old, e) + f4g(cb, e) * (txz_x + tzz_z));
        }
        *reinterpret_cast<float4*>(nxt + ro) = nvx;
        *reinterpret_cast<float4*>(nxt + cs + ro) = nvz;
        tzz_prev = sc.zz;
        sc = sn;
    }
    elastic_forward_tail(a, b, zb0, zb0 + FH, x0, x0 + FW, tid);
}

// range of fast tiles whose cells and halos are all strictly inside the domain
struct FastRange { int fx_lo, fx_hi, fz_lo, fz_hi; };      // inclusive; empty if hi < lo
__host__ __device__ inline FastRange elastic_fast_range(int nz, int nx) {
    FastRange r;
    r.fx_lo = 1; r.fz_lo = 1;
    r.fx_hi = (nx - 2 - FW) / FW;            // x0 + FW <= nx - 2
    r.fz_hi = (nz - 2 - FH) / FH;            // z0 + FH + 1 <= nz - 1
    if (nx - 2 - FW < 0) r.fx_hi = -1;
    if (nz - 2 - FH < 0) r.fz_hi = -1;
    return r;
}

#ifndef ST_EL_MINB
#define ST_EL_MINB 3
#endif
#ifndef ST_EL_FRZ
#define ST_EL_FRZ 8
#endif
__global__ void __launch_bounds__(NT, ST_EL_MINB) elastic2d_forward_kernel(const E2Args a, int nxt_t, int nborder) {
    __shared__ float sv[2][VH][VW];
    __shared__ float st[3][SH][SW];
    const int tid = threadIdx.x, b = blockIdx.y;
    const FastRange fr = elastic_fast_range(a.nz, a.nx);
    if ((int)blockIdx.x < nborder) {
        const int tz = blockIdx.x / nxt_t, tx = blockIdx.x - tz * nxt_t;
        const int fx = tx / 2, fz = tz / 2;                  // fast tile = 2 x 2 border tiles
        if (fx >= fr.fx_lo && fx <= fr.fx_hi && fz >= fr.fz_lo && fz <= fr.fz_hi) return;   // owned by a fast block
        elastic_forward_tile(a, tx, tz, b, sv, st);
    } else {
        const int q = blockIdx.x - nborder, nfx = fr.fx_hi - fr.fx_lo + 1;
        const int fz = fr.fz_lo + q / nfx, fx = fr.fx_lo + q % nfx;
        elastic_forward_fast(a, fx, fz, b, tid);
    }
}

constexpr int MAXSRC = 32;   // sources of one shot cached per tile for the stress correction

__global__ void __launch_bounds__(NT) elastic2d_adjoint_kernel(const E2Args a) {
    __shared__ float sw[2][VH][VW];      // cb * Lam_v
    __shared__ float sg[3][SH][SW];      // a, b, e
    __shared__ int s_src[MAXSRC];
    __shared__ int s_nsrc;
    const int nz = a.nz, nx = a.nx, ld = a.ld;
    const int tid = threadIdx.y * NTX + threadIdx.x;
    const int x0 = blockIdx.x * TX, z0 = blockIdx.y * TZ;
    const int x = x0 + threadIdx.x;
    const bool live = a.lam1 != nullptr;            // Lam_{i+1} == 0 otherwise
    const bool want_grad = live && a.gacc != nullptr;
    const bool fix_src = want_grad && a.amp && (a.src_fmask & 0x1c);

    float gsum[RPT][4];
#pragma unroll
    for (int k = 0; k < RPT; ++k) gsum[k][0] = gsum[k][1] = gsum[k][2] = gsum[k][3] = 0.f;

    const int b_lo = blockIdx.z * a.bchunk, b_hi = min(b_lo + a.bchunk, a.B);
    for (int b = b_lo; b < b_hi; ++b) {
        const long long boff = (long long)b * a.fs;
        float* l0 = a.lam0 + boff;
        if (live) {
            const float* l1 = a.lam1 + boff;
            const float* S0 = a.s0 + boff;
            const float* S1 = a.s1 + boff;
            __syncthreads();
            if (tid == 0) s_nsrc = 0;
            for (int i = tid; i < VH * VW; i += NT) {
                const int lz = i / VW, lx = i - lz * VW;
                const int z = z0 - 2 + lz, xx = x0 - 2 + lx;
                float w0 = 0.f, w1 = 0.f;
                if (z >= 0 && z < nz && xx >= 0 && xx < nx) {
                    const long long idx = (long long)z * ld + xx;
                    const float cb = __ldg(a.coef[4] + idx);
                    w0 = cb * __ldg(l1 + idx);
                    w1 = cb * __ldg(l1 + a.cs + idx);
                }
                sw[0][lz][lx] = w0;
                sw[1][lz][lx] = w1;
            }
            __syncthreads();
            if (fix_src) {
                for (int s = tid; s < a.ns; s += NT) {
                    if (a.src_b[s] != b) continue;
                    const int sz = a.src_z[s], sx = a.src_x[s];
                    if (sz >= z0 - 1 && sz <= z0 + TZ && sx >= x0 - 1 && sx <= x0 + TX) {
                        const int slot = atomicAdd(&s_nsrc, 1);
                        if (slot < MAXSRC) s_src[slot] = s;
                    }
                }
            }
            auto W = [&](int f, int zz, int xx) -> float { return sw[f][zz - z0 + 2][xx - x0 + 2]; };
            // ---- stage A on the tile + 1-cell ring
            for (int i = tid; i < SH * SW; i += NT) {
                const int lz = i / SW, lx = i - lz * SW;
                const int z = z0 - 1 + lz, xx = x0 - 1 + lx;
                float ga = 0.f, gb = 0.f, ge = 0.f;
                if (z >= 0 && z < nz && xx >= 0 && xx < nx) {
                    const long long idx = (long long)z * ld + xx;
                    const E2Coef c = load_ecoef(a, idx);
                    float Lt[3];
                    e2_adj_stress_tot(z, xx, nz, nx, W, __ldg(l1 + 2 * a.cs + idx), __ldg(l1 + 3 * a.cs + idx),
                                      __ldg(l1 + 4 * a.cs + idx), Lt);
                    ga = c.cl2m * Lt[0] + c.cl * Lt[1];
                    gb = c.cl * Lt[0] + c.cl2m * Lt[1];
                    ge = c.cm * Lt[2];
                    const bool owned = lz >= 1 && lz <= TZ && lx >= 1 && lx <= TX;
                    if (owned) {
                        l0[2 * a.cs + idx] = c.ca * Lt[0];
                        l0[3 * a.cs + idx] = c.ca * Lt[1];
                        l0[4 * a.cs + idx] = c.ca * Lt[2];
                    }
                }
                sg[0][lz][lx] = ga; sg[1][lz][lx] = gb; sg[2][lz][lx] = ge;
            }
            __syncthreads();
            auto G = [&](int k, int zz, int xx) -> float { return sg[k][zz - z0 + 1][xx - x0 + 1]; };
            const int nsrc_tile = fix_src ? min(s_nsrc, MAXSRC) : 0;
            const bool src_overflow = fix_src && s_nsrc > MAXSRC;
            if (x < nx) {
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int z = z0 + threadIdx.y + k * NTY;
                    if (z >= nz) break;
                    const long long idx = (long long)z * ld + x;
                    const E2Coef c = load_ecoef(a, idx);
                    const float lvx = __ldg(l1 + idx), lvz = __ldg(l1 + a.cs + idx);
                    float ovx, ovz;
                    e2_adj_velocity(z, x, nz, nx, G, c.ca, lvx, lvz, ovx, ovz);
                    l0[idx] = ovx;
                    l0[a.cs + idx] = ovz;
                    if (want_grad) {
                        // stage-A totals of this cell, recomputed (cheap) for the gradient
                        float Lt[3];
                        e2_adj_stress_tot(z, x, nz, nx, W, __ldg(l1 + 2 * a.cs + idx), __ldg(l1 + 3 * a.cs + idx),
                                          __ldg(l1 + 4 * a.cs + idx), Lt);
                        auto V = [&](int f, int zz, int xx) -> float { return __ldg(S0 + f * a.cs + (long long)zz * ld + xx); };
                        const float vx_x = x > 0 ? V(0, z, x) - V(0, z, x - 1) : 0.f;
                        const float vz_z = z < nz - 1 ? V(1, z + 1, x) - V(1, z, x) : 0.f;
                        const float vx_z = z > 0 ? V(0, z, x) - V(0, z - 1, x) : 0.f;
                        const float vz_x = x < nx - 1 ? V(1, z, x + 1) - V(1, z, x) : 0.f;
                        gsum[k][0] += Lt[0] * vx_x + Lt[1] * vz_z;
                        gsum[k][1] += Lt[0] * vz_z + Lt[1] * vx_x;
                        gsum[k][2] += Lt[2] * (vz_x + vx_z);
                        // new stresses of step i+1 BEFORE the source add: S_{i+1} minus the injected sample
                        auto Tn = [&](int f, int zz, int xx) -> float {
                            float v = __ldg(S1 + (2 + f) * a.cs + (long long)zz * ld + xx);
                            if (fix_src && (a.src_fmask >> (2 + f) & 1)) {
                                if (!src_overflow) {
                                    for (int q = 0; q < nsrc_tile; ++q) {
                                        const int s = s_src[q];
                                        if (a.src_z[s] == zz && a.src_x[s] == xx) v -= a.amp[s];
                                    }
                                } else {
                                    for (int s = 0; s < a.ns; ++s)
                                        if (a.src_b[s] == b && a.src_z[s] == zz && a.src_x[s] == xx) v -= a.amp[s];
                                }
                            }
                            return v;
                        };
                        float fx, fz;
                        e2_stress_div(z, x, nz, nx, Tn, fx, fz);
                        gsum[k][3] += lvx * fx + lvz * fz;
                    }
                }
            }
        } else {
            // Lam_{i+1} == 0: Lam_i is the receiver term only
            if (x < nx) {
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int z = z0 + threadIdx.y + k * NTY;
                    if (z >= nz) break;
                    const long long idx = (long long)z * ld + x;
#pragma unroll
                    for (int f = 0; f < 5; ++f) l0[f * a.cs + idx] = 0.f;
                }
            }
        }
        __syncthreads();
        if (a.rec_adj) {
            __shared__ int s_cnt, s_rows[TZ];
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            if (tid < TZ && z0 + tid < nz) {
                const int row = b * nz + z0 + tid;
                if (a.row_start[row + 1] > a.row_start[row]) s_rows[atomicAdd(&s_cnt, 1)] = z0 + tid;
            }
            __syncthreads();
            const int cnt = s_cnt;
            for (int i = 0; i < cnt; ++i) {
                const int z = s_rows[i];
                const int lo = a.row_start[b * nz + z], hi = a.row_start[b * nz + z + 1];
                for (int r = lo + tid; r < hi; r += NT) {
                    const int rx = a.rec_x[r];
                    if (rx >= x0 && rx < x0 + TX) {
                        const long long o = (long long)a.rec_orig[r] * a.nchan;
                        for (int ch = 0; ch < a.nchan; ++ch)
                            atomicAdd(l0 + a.chan_f[ch] * a.cs + (long long)z * ld + rx, a.rec_adj[o + ch]);
                    }
                }
            }
        }
        if (a.gamp) {
            __syncthreads();
            for (int s = tid; s < a.ns; s += NT) {
                if (a.src_b[s] != b) continue;
                const int sz = a.src_z[s], sx = a.src_x[s];
                if (sz >= z0 && sz < z0 + TZ && sx >= x0 && sx < x0 + TX) {
                    float v = 0.f;
#pragma unroll
                    for (int f = 0; f < 5; ++f)
                        if (a.src_fmask >> f & 1) v += l0[f * a.cs + (long long)sz * ld + sx];
                    a.gamp[s] = v;
                }
            }
        }
    }
    if (want_grad && x < nx) {
        const long long plane = (long long)nz * ld;
        float* gb = a.gacc + (long long)blockIdx.z * 4 * plane;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int z = z0 + threadIdx.y + k * NTY;
            if (z >= nz) break;
            const long long idx = (long long)z * ld + x;
#pragma unroll
            for (int q = 0; q < 4; ++q) gb[q * plane + idx] += gsum[k][q];
        }
    }
}

}  // namespace

int st_elastic2d_launch_forward(const E2Args& a, cudaStream_t st) {
    const int nxt_t = (a.nx + TX - 1) / TX, nzt = (a.nz + TZ - 1) / TZ;
    const FastRange fr = elastic_fast_range(a.nz, a.nx);
    const int nfast = (fr.fx_hi >= fr.fx_lo && fr.fz_hi >= fr.fz_lo) ? (fr.fx_hi - fr.fx_lo + 1) * (fr.fz_hi - fr.fz_lo + 1) : 0;
    dim3 grid(nxt_t * nzt + nfast, a.B);
    elastic2d_forward_kernel<<<grid, NT, 0, st>>>(a, nxt_t, nxt_t * nzt);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

int st_elastic2d_launch_adjoint(const E2Args& a, cudaStream_t st) {
    const int nchunk = (a.B + a.bchunk - 1) / a.bchunk;
    dim3 grid((a.nx + TX - 1) / TX, (a.nz + TZ - 1) / TZ, nchunk), block(NTX, NTY);
    elastic2d_adjoint_kernel<<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}
