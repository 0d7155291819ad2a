// sm_100a kernels for the 2D second-order wave-equation family:
//   acoustic (PML), acoustic_habc, vti_habc2, tti_habc, acoustic_fwim_habc,
//   acoustic_{vti,tti}_lsrtm_habc   -- one template, seven flag sets.
//
// One launch = one time step of every shot in the batch, with the source add and the
// receiver gather fused in (reference: ~30-190 ATen launches per step).
//
// Data layout in HBM: fields [NF][B][nz][ld] fp32, row pitch ld a multiple of 4 so
// that every row starts 16-byte aligned; coefficient planes [nz][ld] shared by all
// shots (they stay L2-resident: <= 8 MB each at the BASELINE sizes vs 126 MB of L2).
#include "st_wave2d.cuh"

namespace {

constexpr int TX = 64;          // tile width  (x, fastest)
constexpr int TZ = 32;          // tile height (z)
constexpr int HALO = 2;         // the one-way blend reads j+1, j+2 along the normal
constexpr int SW = TX + 2 * HALO;
constexpr int SH = TZ + 2 * HALO;
constexpr int NTX = 64, NTY = 4;            // 256 threads; each owns TZ/NTY = 8 rows of one column
constexpr int RPT = TZ / NTY;

__device__ __forceinline__ W2Coef load_coef(const W2Args& a, long long idx) {
    W2Coef c;
    c.r = a.coef[0] ? __ldg(a.coef[0] + idx) : 0.f;
    c.b = a.coef[1] ? __ldg(a.coef[1] + idx) : 0.f;
    c.cxx = a.coef[2] ? __ldg(a.coef[2] + idx) : 0.f;
    c.czz = a.coef[3] ? __ldg(a.coef[3] + idx) : 0.f;
    c.cxz = a.coef[4] ? __ldg(a.coef[4] + idx) : 0.f;
    c.ax = a.coef[5] ? __ldg(a.coef[5] + idx) : 0.f;
    c.az = a.coef[6] ? __ldg(a.coef[6] + idx) : 0.f;
    c.m = a.coef[7] ? __ldg(a.coef[7] + idx) : 0.f;
    return c;
}

template <int FL>
__device__ __forceinline__ W2Coef load_coef_fl(const W2Args& a, long long idx) {
    W2Coef c;
    c.r = __ldg(a.coef[0] + idx);
    c.b = __ldg(a.coef[1] + idx);
    c.cxx = c.czz = c.cxz = c.ax = c.az = c.m = 0.f;
    if (!(FL & ST_F_ISO)) { c.cxx = __ldg(a.coef[2] + idx); c.czz = __ldg(a.coef[3] + idx); }
    if (FL & ST_F_XZ) c.cxz = __ldg(a.coef[4] + idx);
    if (FL & ST_F_G1) { c.ax = __ldg(a.coef[5] + idx); c.az = __ldg(a.coef[6] + idx); }
    if (FL & ST_F_BORN) c.m = __ldg(a.coef[7] + idx);
    return c;
}

// cooperative load of a (TZ+4)x(TX+4) tile (zero outside the domain)
__device__ __forceinline__ void load_tile(float (*s)[SW], const float* __restrict__ src,
                                          int z0, int x0, const W2Geom& g, int tid) {
    for (int i = tid; i < SH * SW; i += NTX * NTY) {
        const int lz = i / SW, lx = i - lz * SW;
        const int z = z0 - HALO + lz, x = x0 - HALO + lx;
        float v = 0.f;
        if (z >= 0 && z < g.nz && x >= 0 && x < g.nx) v = __ldg(src + (long long)z * g.ld + x);
        s[lz][lx] = v;
    }
}

// ------------------------------------------------------------------------------ forward
template <int FL>
__global__ void __launch_bounds__(NTX * NTY) wave2d_forward_kernel(const W2Args a) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    __shared__ float s1[NF][SH][SW];
    const W2Geom g = a.g;
    const int tid = threadIdx.y * NTX + threadIdx.x;
    const int x0 = blockIdx.x * TX, z0 = blockIdx.y * TZ, b = blockIdx.z;
    const long long boff = (long long)b * a.fs;

#pragma unroll
    for (int f = 0; f < NF; ++f) load_tile(s1[f], a.cur + f * a.cs + boff, z0, x0, g, tid);
    __syncthreads();

    const int x = x0 + threadIdx.x;
    if (x < g.nx) {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int z = z0 + threadIdx.y + k * NTY;
            if (z >= g.nz) break;
            const long long idx = (long long)z * g.ld + x;
            const W2Coef c = load_coef_fl<FL>(a, idx);
            // current field: smem tile with global fallback (only the wrapped one-way
            // neighbour of depth bw-1 ever leaves the tile)
            auto H1 = [&](int f, int zz, int xx) -> float {
                const int lz = zz - z0 + HALO, lx = xx - x0 + HALO;
                if (lz >= 0 && lz < SH && lx >= 0 && lx < SW) return s1[f][lz][lx];
                if (zz < 0 || zz >= g.nz || xx < 0 || xx >= g.nx) return 0.f;
                return __ldg(a.cur + f * a.cs + boff + (long long)zz * g.ld + xx);
            };
            auto H2 = [&](int f, int zz, int xx) -> float {
                if (zz < 0 || zz >= g.nz || xx < 0 || xx >= g.nx) return 0.f;
                return __ldg(a.prev + f * a.cs + boff + (long long)zz * g.ld + xx);
            };
            float out[2];
            w2_forward_cell<FL>(z, x, g, c, a.dt, H1, H2, out);
#pragma unroll
            for (int f = 0; f < NF; ++f) a.next[f * a.cs + boff + idx] = out[f];
        }
    }
    __syncthreads();
    // ---- fused source add (source.py:47-57: added to the new field after the step)
    for (int s = tid; s < a.ns; s += NTX * NTY) {
        if (a.src_b[s] != b) continue;
        const int sz = a.src_z[s], sx = a.src_x[s];
        if (sz >= z0 && sz < z0 + TZ && sx >= x0 && sx < x0 + TX) {
            const float v = a.amp[s];
#pragma unroll
            for (int f = 0; f < NF; ++f)
                if (a.src_fmask >> f & 1) atomicAdd(a.next + f * a.cs + boff + (long long)sz * g.ld + sx, v);
        }
    }
    __syncthreads();
    // ---- fused receiver gather (probe.py:42-44: sampled after the source add)
    if (a.rec_out) {
        const int zend = min(z0 + TZ, g.nz);
        for (int z = z0; z < zend; ++z) {
            const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
            for (int r = lo + tid; r < hi; r += NTX * NTY) {
                const int rx = a.rec_x[r];
                if (rx >= x0 && rx < x0 + TX) {
                    const long long o = (long long)a.rec_orig[r] * a.nchan;
                    for (int ch = 0; ch < a.nchan; ++ch)
                        a.rec_out[o + ch] = a.next[a.chan_f[ch] * a.cs + boff + (long long)z * g.ld + rx];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ adjoint
// which of the 7 gradient accumulators (r,cxx,czz,cxz,ax,az,m) a flag set touches
template <int FL>
__host__ __device__ constexpr bool grad_used(int q) {
    return q == 0 ? ((FL & ST_F_ISO) || (FL & ST_F_HABC))
         : (q == 1 || q == 2) ? !(FL & ST_F_ISO)
         : q == 3 ? (FL & ST_F_XZ) != 0
         : (q == 4 || q == 5) ? (FL & ST_F_G1) != 0
         : (FL & ST_F_BORN) != 0;
}

template <int FL>
__global__ void __launch_bounds__(NTX * NTY) wave2d_adjoint_kernel(const W2Args a) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    __shared__ float sl[NF][SH][SW];     // Lam_{i+1}
    __shared__ float ss[NF][SH][SW];     // S_i
    const W2Geom g = a.g;
    const int tid = threadIdx.y * NTX + threadIdx.x;
    const int x0 = blockIdx.x * TX, z0 = blockIdx.y * TZ;
    const int x = x0 + threadIdx.x;
    const bool want_grad = a.gacc != nullptr;

    float gsum[RPT][7];
#pragma unroll
    for (int k = 0; k < RPT; ++k)
#pragma unroll
        for (int q = 0; q < 7; ++q) gsum[k][q] = 0.f;

    const int b_lo = blockIdx.z * a.bchunk;
    const int b_hi = min(b_lo + a.bchunk, a.B);
    for (int b = b_lo; b < b_hi; ++b) {
        const long long boff = (long long)b * a.fs;
        __syncthreads();
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            load_tile(sl[f], a.lam1 + f * a.cs + boff, z0, x0, g, tid);
            load_tile(ss[f], a.s1 + f * a.cs + boff, z0, x0, g, tid);
        }
        __syncthreads();
        if (x < g.nx) {
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int z = z0 + threadIdx.y + k * NTY;
                if (z >= g.nz) break;
                const long long idx = (long long)z * g.ld + x;
                auto inb = [&](int zz, int xx) { return zz >= 0 && zz < g.nz && xx >= 0 && xx < g.nx; };
                auto L1 = [&](int f, int zz, int xx) -> float {
                    const int lz = zz - z0 + HALO, lx = xx - x0 + HALO;
                    if (lz >= 0 && lz < SH && lx >= 0 && lx < SW) return sl[f][lz][lx];
                    if (!inb(zz, xx)) return 0.f;
                    return __ldg(a.lam1 + f * a.cs + boff + (long long)zz * g.ld + xx);
                };
                auto S1 = [&](int f, int zz, int xx) -> float {
                    const int lz = zz - z0 + HALO, lx = xx - x0 + HALO;
                    if (lz >= 0 && lz < SH && lx >= 0 && lx < SW) return ss[f][lz][lx];
                    if (!inb(zz, xx)) return 0.f;
                    return __ldg(a.s1 + f * a.cs + boff + (long long)zz * g.ld + xx);
                };
                auto L2 = [&](int f, int zz, int xx) -> float {
                    if (!inb(zz, xx)) return 0.f;
                    return __ldg(a.lam2 + f * a.cs + boff + (long long)zz * g.ld + xx);
                };
                auto S2 = [&](int f, int zz, int xx) -> float {
                    if (!inb(zz, xx)) return 0.f;
                    return __ldg(a.s2 + f * a.cs + boff + (long long)zz * g.ld + xx);
                };
                auto CF = [&](int zz, int xx) -> W2Coef { return load_coef_fl<FL>(a, (long long)zz * g.ld + xx); };
                float out[2];
                w2_adjoint_cell<FL>(z, x, g, a.dt, L1, L2, S1, S2, CF, out, gsum[k], want_grad);
#pragma unroll
                for (int f = 0; f < NF; ++f) a.lam0[f * a.cs + boff + idx] = out[f];
            }
        }
        __syncthreads();
        // ---- adjoint of the receiver gather: scatter-add d loss / d sample into Lam_i
        if (a.rec_adj) {
            const int zend = min(z0 + TZ, g.nz);
            for (int z = z0; z < zend; ++z) {
                const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
                for (int r = lo + tid; r < hi; r += NTX * NTY) {
                    const int rx = a.rec_x[r];
                    if (rx >= x0 && rx < x0 + TX) {
                        const long long o = (long long)a.rec_orig[r] * a.nchan;
                        for (int ch = 0; ch < a.nchan; ++ch)
                            atomicAdd(a.lam0 + a.chan_f[ch] * a.cs + boff + (long long)z * g.ld + rx, a.rec_adj[o + ch]);
                    }
                }
            }
        }
        // ---- adjoint of the source add: d loss / d amplitude = Lam_i at the source cell
        if (a.gamp) {
            __syncthreads();
            for (int s = tid; s < a.ns; s += NTX * NTY) {
                if (a.src_b[s] != b) continue;
                const int sz = a.src_z[s], sx = a.src_x[s];
                if (sz >= z0 && sz < z0 + TZ && sx >= x0 && sx < x0 + TX) {
                    float v = 0.f;
#pragma unroll
                    for (int f = 0; f < NF; ++f)
                        if (a.src_fmask >> f & 1) v += a.lam0[f * a.cs + boff + (long long)sz * g.ld + sx];
                    a.gamp[s] = v;
                }
            }
        }
    }
    if (want_grad && x < g.nx) {
        const long long plane = (long long)g.nz * g.ld;
        float* gb = a.gacc + (long long)blockIdx.z * 7 * plane;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int z = z0 + threadIdx.y + k * NTY;
            if (z >= g.nz) break;
            const long long idx = (long long)z * g.ld + x;
#pragma unroll
            for (int q = 0; q < 7; ++q)
                if (grad_used<FL>(q)) gb[q * plane + idx] += gsum[k][q];
        }
    }
}

}  // namespace

template <int FL>
int st_w2_launch_fwd(const W2Args& a, cudaStream_t st);
template <int FL>
int st_w2_launch_adj(const W2Args& a, cudaStream_t st);

#ifndef ST_W2_DISPATCH_ONLY
template <int FL>
int st_w2_launch_fwd(const W2Args& a, cudaStream_t st) {
    dim3 grid((a.g.nx + TX - 1) / TX, (a.g.nz + TZ - 1) / TZ, a.B), block(NTX, NTY);
    wave2d_forward_kernel<FL><<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}
template <int FL>
int st_w2_launch_adj(const W2Args& a, cudaStream_t st) {
    const int nchunk = (a.B + a.bchunk - 1) / a.bchunk;
    dim3 grid((a.g.nx + TX - 1) / TX, (a.g.nz + TZ - 1) / TZ, nchunk), block(NTX, NTY);
    wave2d_adjoint_kernel<FL><<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

#ifdef ST_W2_INSTANCE
template int st_w2_launch_fwd<ST_W2_INSTANCE>(const W2Args&, cudaStream_t);
template int st_w2_launch_adj<ST_W2_INSTANCE>(const W2Args&, cudaStream_t);
#endif
#endif  // !ST_W2_DISPATCH_ONLY

#if defined(ST_W2_DISPATCH_ONLY) || !defined(ST_W2_INSTANCE)
#define ST_W2_DISPATCH(FN)                                                                      \
    switch (flags) {                                                                            \
        case ST_F_ISO | ST_F_PML: return FN<ST_F_ISO | ST_F_PML>(a, st);                        \
        case ST_F_ISO | ST_F_HABC: return FN<ST_F_ISO | ST_F_HABC>(a, st);                      \
        case ST_F_HABC: return FN<ST_F_HABC>(a, st);                                            \
        case ST_F_HABC | ST_F_XZ: return FN<ST_F_HABC | ST_F_XZ>(a, st);                        \
        case ST_F_ISO | ST_F_HABC | ST_F_G1: return FN<ST_F_ISO | ST_F_HABC | ST_F_G1>(a, st);  \
        case ST_F_HABC | ST_F_BORN: return FN<ST_F_HABC | ST_F_BORN>(a, st);                    \
        case ST_F_HABC | ST_F_XZ | ST_F_BORN: return FN<ST_F_HABC | ST_F_XZ | ST_F_BORN>(a, st);\
        default: st_set_error("wave2d: unsupported flag set %d", flags); return ST_ERR_UNSUPPORTED; \
    }

int st_wave2d_launch_forward(int flags, const W2Args& a, cudaStream_t st) { ST_W2_DISPATCH(st_w2_launch_fwd) }
int st_wave2d_launch_adjoint(int flags, const W2Args& a, cudaStream_t st) { ST_W2_DISPATCH(st_w2_launch_adj) }
#endif
