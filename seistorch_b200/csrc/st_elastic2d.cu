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

__global__ void __launch_bounds__(NT) elastic2d_forward_kernel(const E2Args a) {
    __shared__ float sv[2][VH][VW];
    __shared__ float st[3][SH][SW];
    const int nz = a.nz, nx = a.nx, ld = a.ld;
    const int tid = threadIdx.y * NTX + threadIdx.x;
    const int x0 = blockIdx.x * TX, z0 = blockIdx.y * TZ, b = blockIdx.z;
    const long long boff = (long long)b * a.fs;
    const float* cur = a.cur + boff;
    float* nxt = a.next + boff;

    // tile completely inside the domain (incl. its 2-cell halo): no bounds predicates needed
    const bool inner = z0 >= 2 && z0 + TZ + 2 <= nz && x0 >= 2 && x0 + TX + 2 <= nx;
    for (int lz = threadIdx.y; lz < VH; lz += NTY) {
        const int z = z0 - 2 + lz;
        for (int lx = threadIdx.x; lx < VW; lx += NTX) {
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
    for (int lz = threadIdx.y; lz < SH; lz += NTY) {
        const int z = z0 - 1 + lz;
        for (int lx = threadIdx.x; lx < SW; lx += NTX) {
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
    const int x = x0 + threadIdx.x;
    if (x < nx) {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int z = z0 + threadIdx.y + k * NTY;
            if (z >= nz) break;
            const long long idx = (long long)z * ld + x;
            float fx, fz;
            e2_stress_div(z, x, nz, nx, T, fx, fz);
            const float ca = __ldg(a.coef[0] + idx), cb = __ldg(a.coef[4] + idx);
            nxt[idx] = ca * V(0, z, x) + cb * fx;
            nxt[a.cs + idx] = ca * V(1, z, x) + cb * fz;
        }
    }
    __syncthreads();
    for (int s = tid; s < a.ns; s += NT) {
        if (a.src_b[s] != b) continue;
        const int sz = a.src_z[s], sx = a.src_x[s];
        if (sz >= z0 && sz < z0 + TZ && sx >= x0 && sx < x0 + TX) {
            const float v = a.amp[s];
#pragma unroll
            for (int f = 0; f < 5; ++f)
                if (a.src_fmask >> f & 1) atomicAdd(nxt + f * a.cs + (long long)sz * ld + sx, v);
        }
    }
    if (a.rec_out) {
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
                        a.rec_out[o + ch] = nxt[a.chan_f[ch] * a.cs + (long long)z * ld + rx];
                }
            }
        }
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
    dim3 grid((a.nx + TX - 1) / TX, (a.nz + TZ - 1) / TZ, a.B), block(NTX, NTY);
    elastic2d_forward_kernel<<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

int st_elastic2d_launch_adjoint(const E2Args& a, cudaStream_t st) {
    const int nchunk = (a.B + a.bchunk - 1) / a.bchunk;
    dim3 grid((a.nx + TX - 1) / TX, (a.nz + TZ - 1) / TZ, nchunk), block(NTX, NTY);
    elastic2d_adjoint_kernel<<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}
