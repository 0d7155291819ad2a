// sm_100a kernels for the 2D elastic velocity-stress equation (one launch per time step,
// stresses then velocities fused through shared memory, source add + receiver gather
// fused in).  Reference: equations2d/elastic.py:7-37 (~100 ATen launches per step).
//
// Layout: state [5][B][nz][ld] = (vx, vz, txx, tzz, txz); coefficient planes [nz][ld].
#include <cstdlib>

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

// EDGE = true: the tile touches the domain boundary (or hangs over it).  Loads outside the domain return zero and
// the four one-sided differences that the reference zeroes at the edges (equations2d/utils.py:3-48: D+ is 0 at index
// 0, D- is 0 at the last index) are multiplied by 0/1 masks; nothing else differs, so the interior tiles (EDGE =
// false) skip all of it.
#ifndef ST_EL_STORE
#define ST_EL_STORE 0                       // tuning: 1 = streaming (evict-first) stores of the new state, 2 = write-through
#endif
__device__ __forceinline__ void st4(float* p, const float4& v) {
#if ST_EL_STORE == 1
    __stcs(reinterpret_cast<float4*>(p), v);
#elif ST_EL_STORE == 2
    __stwt(reinterpret_cast<float4*>(p), v);
#else
    *reinterpret_cast<float4*>(p) = v;
#endif
}

template <bool EDGE>
__device__ __forceinline__ void elastic_forward_fast(const E2Args& a, int fx, int fz, int b, int tid) {
    const int ld = a.ld, nz = a.nz, nx = a.nx;
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
    auto L4 = [&](const float* p, int r) {
        if (EDGE && (r < 0 || r >= nz || x >= ld)) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(p + (r * ld + x)));
    };
    auto L1 = [&](const float* p, int r, int xx) {
        if (EDGE && (r < 0 || r >= nz || xx < 0 || xx >= nx)) return 0.f;
        return __ldg(p + (r * ld + xx));
    };
    // 0/1 masks of the one-sided differences (EDGE only): mlo[e] = column x+e > 0, mhi[e] = column x+e < nx-1
    float mlo[4], mhi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { mlo[e] = (!EDGE || x + e > 0) ? 1.f : 0.f; mhi[e] = (!EDGE || x + e < nx - 1) ? 1.f : 0.f; }

    // new stresses of row r from vx(r-1), vx(r), vz(r), vz(r+1)
    auto stress = [&](int r, const float4& vxu, const float4& vxr, const float4& vzr, const float4& vzd) {
        StressRow s;
        const float4 ca = L4(CA, r), c2 = L4(C2, r), cl = L4(CL, r), cm = L4(CM, r);
        const float4 oxx = L4(TXX, r), ozz = L4(TZZ, r), oxz = L4(TXZ, r);
        float vx_l = __shfl_up_sync(0xffffffffu, vxr.w, 1);
        float vz_r = __shfl_down_sync(0xffffffffu, vzr.x, 1);
        s.xx_r = 0.f; s.xz_l = 0.f;
        const float mz0 = (!EDGE || r > 0) ? 1.f : 0.f, mz1 = (!EDGE || r < nz - 1) ? 1.f : 0.f;   // row masks of D+z / D-z
        {
            // halo cells of the two edge lanes, evaluated by every lane without a branch (broadcast loads: one sector
            // each) so that they are issued together with the vector loads instead of as dependent load -> use chains
            const int xl = x0 - 1, xr = x0 + FW;
            const float hvx_l = L1(VX, r, xl), hvz_l = L1(VZ, r, xl), hvx_ul = L1(VX, r - 1, xl);
            const float hca_l = L1(CA, r, xl), hxz_l = L1(TXZ, r, xl), hcm_l = L1(CM, r, xl);
            const float hvz_r = L1(VZ, r, xr), hvx_r = L1(VX, r, xr), hvz_dr = L1(VZ, r + 1, xr);
            const float hca_r = L1(CA, r, xr), hxx_r = L1(TXX, r, xr), hc2_r = L1(C2, r, xr), hcl_r = L1(CL, r, xr);
            // txz'(r, x0-1): vz_x = vz(r,x0) - vz(r,x0-1), vx_z = vx(r,x0-1) - vx(r-1,x0-1)   (x0-1 < nx-1 always)
            float vz_x = vzr.x - hvz_l, vx_z = hvx_l - hvx_ul;
            if (EDGE) vx_z *= mz0;
            s.xz_l = hca_l * hxz_l + hcm_l * (vz_x + vx_z);
            // txx'(r, x0+128): vx_x = vx(r,x0+128) - vx(r,x0+127), vz_z = vz(r+1,x0+128) - vz(r,x0+128)   (x0+FW > 0 always)
            float vx_x = hvx_r - vxr.w, vz_z = hvz_dr - hvz_r;
            if (EDGE) vz_z *= mz1;
            s.xx_r = hca_r * hxx_r + (hc2_r * vx_x + hcl_r * vz_z);
            vx_l = e0 ? hvx_l : vx_l;
            vz_r = e31 ? hvz_r : vz_r;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float vx_x = f4g(vxr, e) - (e == 0 ? vx_l : f4g(vxr, e - 1));
            float vz_z = f4g(vzd, e) - f4g(vzr, e);
            float vx_z = f4g(vxr, e) - f4g(vxu, e);
            float vz_x = (e == 3 ? vz_r : f4g(vzr, e + 1)) - f4g(vzr, e);
            if (EDGE) { vx_x *= mlo[e]; vz_z *= mz1; vx_z *= mz0; vz_x *= mhi[e]; }
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
        const bool st_ok = !EDGE || (z < nz && x < ld);       // pad columns [nx, ld) come out as exact zeros (zero coefficients)
        if (st_ok) {
            st4(nxt + 2 * cs + ro, sc.xx);
            st4(nxt + 3 * cs + ro, sc.zz);
            st4(nxt + 4 * cs + ro, sc.xz);
        }
        const float mz0 = (!EDGE || z > 0) ? 1.f : 0.f, mz1 = (!EDGE || z < nz - 1) ? 1.f : 0.f;
        float txx_r = __shfl_down_sync(0xffffffffu, sc.xx.x, 1);
        float txz_l = __shfl_up_sync(0xffffffffu, sc.xz.w, 1);
        txx_r = e31 ? sc.xx_r : txx_r;
        txz_l = e0 ? sc.xz_l : txz_l;
        const float4 ca = L4(CA, z), cb = L4(CB, z);
        float4 nvx, nvz;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float txx_x = (e == 3 ? txx_r : f4g(sc.xx, e + 1)) - f4g(sc.xx, e);
            float txz_z = f4g(sn.xz, e) - f4g(sc.xz, e);
            float tzz_z = f4g(sc.zz, e) - f4g(tzz_prev, e);
            float txz_x = f4g(sc.xz, e) - (e == 0 ? txz_l : f4g(sc.xz, e - 1));
            if (EDGE) { txx_x *= mhi[e]; txz_z *= mz1; tzz_z *= mz0; txz_x *= mlo[e]; }
            f4s(nvx, e, f4g(ca, e) * f4g(vx_old, e) + f4g(cb, e) * (txx_x + txz_z));
            f4s(nvz, e, f4g(ca, e) * f4g(vz_old, e) + f4g(cb, e) * (txz_x + tzz_z));
        }
        if (st_ok) {
            st4(nxt + ro, nvx);
            st4(nxt + cs + ro, nvz);
        }
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
#define ST_EL_MINB 2
#endif
#ifndef ST_EL_FRZ
#define ST_EL_FRZ 8
#endif
// grid = (fast tiles, shots): every 128 x 64 tile runs the register / shuffle path; tiles that touch the domain
// boundary take the masked variant.
__global__ void __launch_bounds__(NT, ST_EL_MINB) elastic2d_forward_kernel(const E2Args a, int nfx) {
    st_pdl_launch_dependents();                             // (no-ops unless launched with programmatic stream serialization)
    st_pdl_wait();
    const int tid = threadIdx.x, b = blockIdx.y;
    const FastRange fr = elastic_fast_range(a.nz, a.nx);
    const int fz = blockIdx.x / nfx, fx = blockIdx.x - fz * nfx;
    if (fx >= fr.fx_lo && fx <= fr.fx_hi && fz >= fr.fz_lo && fz <= fr.fz_hi) elastic_forward_fast<false>(a, fx, fz, b, tid);
    else elastic_forward_fast<true>(a, fx, fz, b, tid);
}

// receiver terms (transpose of the gather: Lam_i += d loss / d sample) and d loss / d wavelet sample of the cells
// [z0,zn) x [x0,xn) of shot b, after the block stored Lam_i there
template <int NTHREADS>
__device__ __forceinline__ void elastic_adjoint_tail_t(const E2Args& a, int b, int z0, int zn, int x0, int xn, int tid) {
    const int nz = a.nz, ld = a.ld;
    float* l0 = a.lam0 + (long long)b * a.fs;
    __syncthreads();
    if (a.rec_adj) {
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
            for (int r = lo + tid; r < hi; r += NTHREADS) {
                const int rx = a.rec_x[r];
                if (rx >= x0 && rx < xn) {
                    const long long o = (long long)a.rec_orig[r] * a.nchan;
                    for (int ch = 0; ch < a.nchan; ++ch)
                        atomicAdd(l0 + a.chan_f[ch] * a.cs + (long long)z * ld + rx, a.rec_adj[o + ch]);
                }
            }
        }
    }
    if (a.gamp) {
        __syncthreads();
        for (int s = tid; s < a.ns; s += NTHREADS) {
            if (a.src_b[s] != b) continue;
            const int sz = a.src_z[s], sx = a.src_x[s];
            if (sz >= z0 && sz < zn && sx >= x0 && sx < xn) {
                float v = 0.f;
#pragma unroll
                for (int f = 0; f < 5; ++f)
                    if (a.src_fmask >> f & 1) v += l0[f * a.cs + (long long)sz * ld + sx];
                a.gamp[s] = v;
            }
        }
    }
}
__device__ __forceinline__ void elastic_adjoint_tail(const E2Args& a, int b, int z0, int zn, int x0, int xn, int tid) {
    elastic_adjoint_tail_t<NT>(a, b, z0, zn, x0, xn, tid);
}

// ------------------------------------------------------------------------------------------------ adjoint, fast path
// Exact transpose of the two-stage update (DESIGN.md "adjoint"), vectorised like the forward fast path: a warp owns
// 128 columns x AFRZ rows, a lane 4 consecutive cells, rows are marched top-down with every operand loaded once as a
// 128-bit vector; x-neighbours by warp shuffle, the two halo columns of the warp recomputed by the edge lanes from scalar
// loads.  With  W = cb * Lam_v  (v = vx, vz) and the 0/1 masks of the reference's one-sided differences
//   stage A (row r):  Ltxx = ltxx + mA Wx(x-1) - mB Wx          Ltzz = ltzz + m0 Wz - m1 Wz(r+1)
//                     Ltxz = ltxz + m0 Wx(r-1) - m1 Wx + mA Wz - mB Wz(x+1)
//                     a = c2 Ltxx + cl Ltzz,  b = cl Ltxx + c2 Ltzz,  e = cm Ltxz;   Lam_i(stress) = ca L
//   stage B (row z):  Lam_i(vx) = ca lvx + mA a - mB a(x+1) + m0 e - m1 e(z+1)
//                     Lam_i(vz) = ca lvz + m0 b(z-1) - m1 b + mA e(x-1) - mB e
//   gradient:         g_c2 += Ltxx vx_x + Ltzz vz_z,  g_cl += Ltxx vz_z + Ltzz vx_x,  g_cm += Ltxz (vz_x + vx_z)   (S_i)
//                     g_cb += lvx (txx_x + txz_z) + lvz (txz_x + tzz_z)                                            (S_{i+1})
// (mA: x > 0, mB: x < nx-1, m0: row > 0, m1: row < nz-1).  Stage A runs one row ahead of stage B.  Each (tile, shot)
// block read-modify-writes its own cells of the shot's gradient planes (bchunk = 1: one plane set per shot), so the
// result does not depend on block scheduling.
#ifndef ST_EL_AFRZ
#define ST_EL_AFRZ 4
#endif
#ifndef ST_EL_PF
#define ST_EL_PF 0                          // tuning: 1 = L2 prefetch of the warp's rows before the marching loop (adjoint)
#endif
#ifndef ST_EL_AMINB
#define ST_EL_AMINB 2
#endif
constexpr int AFRZ = ST_EL_AFRZ, AFH = AFRZ * (NT / 32);
constexpr int AGP = (NT / 32) * AFRZ * 32;          // float4s per shared-memory gradient plane of an adjoint block
constexpr int ADJ_FAST_SMEM = 4 * AGP * 16;         // c2, cl, cm, cb

__device__ __forceinline__ float4 f4mul(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4z() { return make_float4(0.f, 0.f, 0.f, 0.f); }

struct AdjRow { float4 a, b, e; float a_r, e_l; };       // stage-A products of one row (+ halo columns x0+FW / x0-1)

template <bool EDGE>
__device__ __forceinline__ void elastic_adjoint_fast(const E2Args& a, int fx, int fz, int b, int tid, bool grad_on, float4* __restrict__ gs) {
    // gs: this lane's float4 of the block's shared-memory gradient planes (row stride 32, plane stride AGP float4s); the
    // partial sums of the block's shots stay there and are flushed once per chunk (elastic2d_adjoint_fast_kernel)
    const int ld = a.ld, nz = a.nz, nx = a.nx;
    const int warp = tid >> 5, lane = tid & 31;
    const int x0 = fx * FW, zb0 = fz * AFH, z0 = zb0 + warp * AFRZ;
    const int x = x0 + 4 * lane;
    const long long boff = (long long)b * a.fs, cs = a.cs;
    const float* LVX = a.lam1 + boff;
    const float* LVZ = LVX + cs;
    const float* LXX = LVX + 2 * cs;
    const float* LZZ = LVX + 3 * cs;
    const float* LXZ = LVX + 4 * cs;
    float* out = a.lam0 + boff;
    const float* SVX = a.s0 + boff;
    const float* SVZ = SVX + cs;
    const float* TXX = a.s1 + boff + 2 * cs;
    const float* TZZ = TXX + cs;
    const float* TXZ = TXX + 2 * cs;
    const float* CA = a.coef[0];
    const float* C2 = a.coef[1];
    const float* CL = a.coef[2];
    const float* CM = a.coef[3];
    const float* CB = a.coef[4];
    const bool e0 = lane == 0, e31 = lane == 31;
    const bool grad = grad_on;
    auto L4 = [&](const float* p, int r) {
        if (EDGE && (r < 0 || r >= nz || x >= ld)) return f4z();
        return __ldg(reinterpret_cast<const float4*>(p + (r * ld + x)));
    };
    auto L1 = [&](const float* p, int r, int xx) {
        if (EDGE && (r < 0 || r >= nz || xx < 0 || xx >= nx)) return 0.f;
        return __ldg(p + (r * ld + xx));
    };
    float mA[4], mB[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { mA[e] = (!EDGE || x + e > 0) ? 1.f : 0.f; mB[e] = (!EDGE || x + e < nx - 1) ? 1.f : 0.f; }
    const int xl = x0 - 1, xr = x0 + FW;                        // halo columns
    const int xh = lane < 16 ? xl : xr;                         // the one this lane would need if it were an edge lane
    const float mA_l = (!EDGE || xl > 0) ? 1.f : 0.f;           // masks at the halo cells
    const float mB_r = (!EDGE || xr < nx - 1) ? 1.f : 0.f;

    // W = cb * Lam_v of row r: vectors + the halo scalars of the edge lanes (lane 0: column x0-1, lane 31: column x0+FW)
    struct WRow { float4 wx, wz; float hx, hz; };
    auto load_w = [&](int r) {
        WRow w;
        const float4 cb = L4(CB, r);
        w.wx = f4mul(cb, L4(LVX, r));
        w.wz = f4mul(cb, L4(LVZ, r));
        // halo scalars without a branch: the lower half-warp looks left, the upper half right (only lanes 0 / 31 use them)
        const float cbh = L1(CB, r, xh);
        w.hx = cbh * L1(LVX, r, xh);
        w.hz = cbh * L1(LVZ, r, xh);
        return w;
    };

    // stage A of row r.  up / cur / dn = W rows r-1, r, r+1.  own: the row belongs to this warp (store + gradient).
    float4 svx_up = f4z();                                      // S_i vx of the row above (gradient)
    float4 svz_cur = f4z();                                     // S_i vz of the current row (gradient)
    auto stage_a = [&](int r, const WRow& up, const WRow& cur, const WRow& dn, bool own) {
        AdjRow o;
        const float m0 = (!EDGE || r > 0) ? 1.f : 0.f, m1 = (!EDGE || r < nz - 1) ? 1.f : 0.f;
        const float4 lxx = L4(LXX, r), lzz = L4(LZZ, r), lxz = L4(LXZ, r);
        const float4 ca = L4(CA, r), c2 = L4(C2, r), cl = L4(CL, r), cm = L4(CM, r);
        float wx_l = __shfl_up_sync(0xffffffffu, cur.wx.w, 1);
        float wz_r = __shfl_down_sync(0xffffffffu, cur.wz.x, 1);
        wx_l = e0 ? cur.hx : wx_l;
        wz_r = e31 ? cur.hz : wz_r;
        float4 Lxx, Lzz, Lxz;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float wxc = f4g(cur.wx, e), wzc = f4g(cur.wz, e);
            const float wxw = e == 0 ? wx_l : f4g(cur.wx, e - 1);
            const float wze = e == 3 ? wz_r : f4g(cur.wz, e + 1);
            float lx = f4g(lxx, e), lz = f4g(lzz, e), le = f4g(lxz, e);
            if (EDGE) {
                lx += mA[e] * wxw - mB[e] * wxc;
                lz += m0 * wzc - m1 * f4g(dn.wz, e);
                le += (m0 * f4g(up.wx, e) - m1 * wxc) + (mA[e] * wzc - mB[e] * wze);
            } else {
                lx += wxw - wxc;
                lz += wzc - f4g(dn.wz, e);
                le += (f4g(up.wx, e) - wxc) + (wzc - wze);
            }
            f4s(Lxx, e, lx); f4s(Lzz, e, lz); f4s(Lxz, e, le);
            f4s(o.a, e, f4g(c2, e) * lx + f4g(cl, e) * lz);
            f4s(o.b, e, f4g(cl, e) * lx + f4g(c2, e) * lz);
            f4s(o.e, e, f4g(cm, e) * le);
        }
        // halo cells of the edge lanes: e(r, x0-1) for lane 0, a(r, x0+FW) for lane 31
        {
            // (every lane evaluates both, branch-free; only lanes 0 / 31 use the result)
            // Ltxz(q), q = (r, x0-1): Wx(r-1,q) / Wx(r,q) / Wz(r,q) are the halo scalars, Wz(r, x0) is this lane's own .x
            const float le = L1(LXZ, r, xl) + (m0 * up.hx - m1 * cur.hx) + (mA_l * cur.hz - cur.wz.x);
            o.e_l = L1(CM, r, xl) * le;
            // Ltxx(q), Ltzz(q), q = (r, x0+FW): Wx(r, x0+FW-1) is this lane's own .w
            const float lx = L1(LXX, r, xr) + (cur.wx.w - mB_r * cur.hx);
            const float lz = L1(LZZ, r, xr) + (m0 * cur.hz - m1 * dn.hz);
            o.a_r = L1(C2, r, xr) * lx + L1(CL, r, xr) * lz;
        }
        // gradient operands: S_i velocities of rows r-1 (vx), r, r+1 (vz)
        float4 svx = f4z(), svz_dn = f4z();
        if (grad) { svx = L4(SVX, r); svz_dn = L4(SVZ, r + 1); }
        if (own) {
            const int ro = r * ld + x;
            const bool st_ok = !EDGE || (r < nz && x < ld);
            if (st_ok) {
                *reinterpret_cast<float4*>(out + 2 * cs + ro) = f4mul(ca, Lxx);
                *reinterpret_cast<float4*>(out + 3 * cs + ro) = f4mul(ca, Lzz);
                *reinterpret_cast<float4*>(out + 4 * cs + ro) = f4mul(ca, Lxz);
            }
            if (grad) {
                float vx_l = __shfl_up_sync(0xffffffffu, svx.w, 1);
                float vz_r = __shfl_down_sync(0xffffffffu, svz_cur.x, 1);
                const float hvx = L1(SVX, r, xl), hvz = L1(SVZ, r, xr);
                vx_l = e0 ? hvx : vx_l;
                vz_r = e31 ? hvz : vz_r;
                if (st_ok) {
                    float4* gp = gs + (r - z0) * 32;
                    float4 g2 = gp[0], gl = gp[AGP], gm = gp[2 * AGP];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float vx_x = f4g(svx, e) - (e == 0 ? vx_l : f4g(svx, e - 1));
                        float vz_z = f4g(svz_dn, e) - f4g(svz_cur, e);
                        float vx_z = f4g(svx, e) - f4g(svx_up, e);
                        float vz_x = (e == 3 ? vz_r : f4g(svz_cur, e + 1)) - f4g(svz_cur, e);
                        if (EDGE) { vx_x *= mA[e]; vz_z *= m1; vx_z *= m0; vz_x *= mB[e]; }
                        const float lx = f4g(Lxx, e), lz = f4g(Lzz, e), le = f4g(Lxz, e);
                        f4s(g2, e, f4g(g2, e) + (lx * vx_x + lz * vz_z));
                        f4s(gl, e, f4g(gl, e) + (lx * vz_z + lz * vx_x));
                        f4s(gm, e, f4g(gm, e) + le * (vz_x + vx_z));
                    }
                    gp[0] = g2; gp[AGP] = gl; gp[2 * AGP] = gm;
                }
            }
        }
        svx_up = svx;
        svz_cur = svz_dn;
        return o;
    };

#if ST_EL_PF
    // pull every row this warp will stream into L2 up front (one prefetch per 128-byte line): the marching loop below is
    // latency-bound at 16 warps per SM, an L2 hit costs a third of a DRAM miss
    {
        auto pf = [&](const float* base, int r) {
            if ((lane & 7) == 0 && r >= 0 && r < nz && x < ld) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (r * ld + x)));
        };
        for (int r = z0 - 2; r <= z0 + AFRZ + 1; ++r) { pf(LVX, r); pf(LVZ, r); }
        for (int r = z0 - 1; r <= z0 + AFRZ; ++r) { pf(LXX, r); pf(LZZ, r); pf(LXZ, r); }
        if (grad) {
            for (int r = z0 - 2; r <= z0 + AFRZ + 1; ++r) { pf(SVX, r); pf(SVZ, r); }
            for (int r = z0 - 1; r <= z0 + AFRZ; ++r) { pf(TXX, r); pf(TZZ, r); pf(TXZ, r); }
        }
    }
#endif
    // ---- prologue: W rows z0-2 .. z0, stage A of rows z0-1 (b only is used) and z0
    WRow w_up = load_w(z0 - 2), w_cur = load_w(z0 - 1), w_dn = load_w(z0);
    if (grad) { svx_up = L4(SVX, z0 - 2); svz_cur = L4(SVZ, z0 - 1); }
    const AdjRow am = stage_a(z0 - 1, w_up, w_cur, w_dn, false);
    float4 b_prev = am.b;
    w_up = w_cur; w_cur = w_dn; w_dn = load_w(z0 + 1);
    AdjRow ac = stage_a(z0, w_up, w_cur, w_dn, true);
    float4 tzz_prev = grad ? L4(TZZ, z0 - 1) : f4z();
    float4 txz_cur = grad ? L4(TXZ, z0) : f4z();
#pragma unroll 1
    for (int k = 0; k < AFRZ; ++k) {
        const int z = z0 + k, ro = z * ld + x;
        // stage A one row ahead (row z+1; owned unless it is the first row of the next warp)
        w_up = w_cur; w_cur = w_dn; w_dn = load_w(z + 2);
        const AdjRow an = stage_a(z + 1, w_up, w_cur, w_dn, k + 1 < AFRZ);
        // stage B of row z
        const float m0 = (!EDGE || z > 0) ? 1.f : 0.f, m1 = (!EDGE || z < nz - 1) ? 1.f : 0.f;
        const float4 ca = L4(CA, z), lvx = L4(LVX, z), lvz = L4(LVZ, z);
        float a_r = __shfl_down_sync(0xffffffffu, ac.a.x, 1);
        float e_l = __shfl_up_sync(0xffffffffu, ac.e.w, 1);
        a_r = e31 ? ac.a_r : a_r;
        e_l = e0 ? ac.e_l : e_l;
        float4 ovx, ovz;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float ae = e == 3 ? a_r : f4g(ac.a, e + 1);
            const float ew = e == 0 ? e_l : f4g(ac.e, e - 1);
            float vx = f4g(ca, e) * f4g(lvx, e), vz = f4g(ca, e) * f4g(lvz, e);
            if (EDGE) {
                vx += (mA[e] * f4g(ac.a, e) - mB[e] * ae) + (m0 * f4g(ac.e, e) - m1 * f4g(an.e, e));
                vz += (m0 * f4g(b_prev, e) - m1 * f4g(ac.b, e)) + (mA[e] * ew - mB[e] * f4g(ac.e, e));
            } else {
                vx += (f4g(ac.a, e) - ae) + (f4g(ac.e, e) - f4g(an.e, e));
                vz += (f4g(b_prev, e) - f4g(ac.b, e)) + (ew - f4g(ac.e, e));
            }
            f4s(ovx, e, vx); f4s(ovz, e, vz);
        }
        const bool st_ok = !EDGE || (z < nz && x < ld);
        if (st_ok) {
            *reinterpret_cast<float4*>(out + ro) = ovx;
            *reinterpret_cast<float4*>(out + cs + ro) = ovz;
        }
        if (grad) {
            // g_cb: Lam_v . divergence of the new stresses of step i+1 (S_{i+1}; velocity-type sources only, see launch)
            const float4 txx = L4(TXX, z), tzz = L4(TZZ, z), txz_dn = L4(TXZ, z + 1);
            float txx_r = __shfl_down_sync(0xffffffffu, txx.x, 1);
            float txz_l = __shfl_up_sync(0xffffffffu, txz_cur.w, 1);
            const float hxx = L1(TXX, z, xr), hxz = L1(TXZ, z, xl);
            txx_r = e31 ? hxx : txx_r;
            txz_l = e0 ? hxz : txz_l;
            if (st_ok) {
                float4 gb = gs[3 * AGP + k * 32];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float txx_x = (e == 3 ? txx_r : f4g(txx, e + 1)) - f4g(txx, e);
                    float txz_z = f4g(txz_dn, e) - f4g(txz_cur, e);
                    float tzz_z = f4g(tzz, e) - f4g(tzz_prev, e);
                    float txz_x = f4g(txz_cur, e) - (e == 0 ? txz_l : f4g(txz_cur, e - 1));
                    if (EDGE) { txx_x *= mB[e]; txz_z *= m1; tzz_z *= m0; txz_x *= mA[e]; }
                    f4s(gb, e, f4g(gb, e) + (f4g(lvx, e) * (txx_x + txz_z) + f4g(lvz, e) * (txz_x + tzz_z)));
                }
                gs[3 * AGP + k * 32] = gb;
            }
            tzz_prev = tzz;
            txz_cur = txz_dn;
        }
        b_prev = ac.b;
        ac = an;
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
        elastic_adjoint_tail(a, b, z0, z0 + TZ, x0, x0 + TX, tid);
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

// grid = (tiles of 128 x AFH cells, shot chunks); tiles touching the domain boundary take the masked variant.  A block walks
// the shots of its chunk with the gradient partial sums of its cells in shared memory and read-modify-writes the chunk's
// gradient planes once (every cell has one owner: deterministic, no atomics).
__global__ void __launch_bounds__(NT, ST_EL_AMINB) elastic2d_adjoint_fast_kernel(const E2Args a, int nfx) {
    st_pdl_launch_dependents();
    st_pdl_wait();
    extern __shared__ __align__(16) float4 el_gsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, chunk = blockIdx.y;
    const int fz = blockIdx.x / nfx, fx = blockIdx.x - fz * nfx;
    const int x0 = fx * FW, z0 = fz * AFH;
    const bool grad = a.gacc != nullptr;
    float4* gs = el_gsm + warp * AFRZ * 32 + lane;
    if (grad) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < AFRZ; ++k) gs[q * AGP + k * 32] = f4z();
    }
    const bool inner = z0 >= 3 && z0 + AFH + 3 <= a.nz && x0 >= 2 && x0 + FW + 2 <= a.nx;
    const int b_lo = chunk * a.bchunk, b_hi = min(b_lo + a.bchunk, a.B);
    for (int b = b_lo; b < b_hi; ++b) {
        if (inner) elastic_adjoint_fast<false>(a, fx, fz, b, tid, grad, gs);
        else elastic_adjoint_fast<true>(a, fx, fz, b, tid, grad, gs);
        elastic_adjoint_tail(a, b, z0, z0 + AFH, x0, x0 + FW, tid);
    }
    if (grad) {
        const long long plane = (long long)a.nz * a.ld;
        float* gpl = a.gacc + (long long)chunk * 4 * plane;       // one set of 4 planes per chunk
        const int x = x0 + 4 * lane;
#pragma unroll
        for (int k = 0; k < AFRZ; ++k) {
            const int z = z0 + warp * AFRZ + k;
            if (z < a.nz && x < a.ld) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4* p4 = reinterpret_cast<float4*>(gpl + q * plane + (z * a.ld + x));
                    const float4 v = *p4, w = gs[q * AGP + k * 32];
                    *p4 = make_float4(v.x + w.x, v.y + w.y, v.z + w.z, v.w + w.w);
                }
            }
        }
    }
}

}  // namespace

int st_elastic2d_launch_forward(const E2Args& a, cudaStream_t st) {
    const int nfx = (a.nx + FW - 1) / FW, nfz = (a.nz + FH - 1) / FH;
    dim3 grid(nfx * nfz, a.B);
    return st_pdl_launch(elastic2d_forward_kernel, grid, dim3(NT), 0, st, a, nfx) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

int st_elastic2d_launch_adjoint(const E2Args& a, cudaStream_t st) {
    // fast path: Lam_{i+1} live, one gradient plane set per shot, no stress-type sources (their injected samples would
    // have to be taken out of S_{i+1} for the cb gradient: the generic kernel does that)
    static const bool fast_on = !(getenv("SEISTORCH_B200_EL_ADJ_FAST") && atoi(getenv("SEISTORCH_B200_EL_ADJ_FAST")) == 0);
    if (fast_on && a.lam1 != nullptr && !(a.gacc && a.amp && (a.src_fmask & 0x1c))) {
        const int nfx = (a.nx + FW - 1) / FW, nfz = (a.nz + AFH - 1) / AFH;
        dim3 grid(nfx * nfz, (a.B + a.bchunk - 1) / a.bchunk);
        if (st_set_max_smem<elastic2d_adjoint_fast_kernel>(ADJ_FAST_SMEM) != cudaSuccess) return ST_ERR_CUDA;
        return st_pdl_launch(elastic2d_adjoint_fast_kernel, grid, dim3(NT), ADJ_FAST_SMEM, st, a, nfx) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
    }
    const int nchunk = (a.B + a.bchunk - 1) / a.bchunk;
    dim3 grid((a.nx + TX - 1) / TX, (a.nz + TZ - 1) / TZ, nchunk), block(NTX, NTY);
    elastic2d_adjoint_kernel<<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}
