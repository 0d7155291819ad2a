// sm_100a kernels for the 3D acoustic equation with PML damping.
// Reference: equations3d/acoustic.py:65-85 (conv3d 7-point Laplacian + ~12 elementwise
// passes per step, each over a 300 MB field at the BASELINE size).
//
//   forward:  y = h1 + alpha (h1 - h2) + ciso * lap7(h1)            ciso = (vp dt/h)^2/(1+b dt)
//   adjoint:  the damped acoustic operator is self-adjoint up to the diagonal scaling ciso:
//             with the scaled cotangent  w = ciso * Lam  the exact discrete adjoint
//                 Lam_i = (1+alpha) Lam_{i+1} + lap7(ciso Lam_{i+1}) - alpha Lam_{i+2}
//             becomes  w_i = w1 + alpha (w1 - w2) + ciso * lap7(w1)  -- the forward kernel --
//             plus the imaging condition  g_ciso += (w1/ciso) * lap7(S_i)  and receiver terms
//             scaled by ciso.  One template serves both.  The kernel accumulates  w1 * lap7(S_i)  only: 1/ciso does
//             not depend on time, the caller divides the accumulated plane once (include/seistorch_b200.h).
//
// Same streaming structure as the 2D fast path: a warp owns 128 columns (n2, fastest) x RZ
// rows (n1) of one n0-plane; rows are 128-bit vector loads marched through a 3-row register
// pipeline, n2 neighbours by warp shuffle, the n0 neighbours are two more vector loads that
// hit L2 (planes are visited in order, a plane is 0.6 MB at the BASELINE size).
#include "st_acoustic3d.cuh"

namespace {

constexpr int NT = 256, NWARP = NT / 32;
constexpr int FW = 128;         // columns per warp
#ifndef ST_A3_RZ
#define ST_A3_RZ 4
#endif
constexpr int RZ = ST_A3_RZ;    // rows per warp
constexpr int FH = RZ * NWARP;  // rows per block

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4get(const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4set(float4& v, int e, float s) { if (e == 0) v.x = s; else if (e == 1) v.y = s; else if (e == 2) v.z = s; else v.w = s; }

struct G3 { int n0, n1, n2, ld; long long ps; };

// row (i0, i1) of a field as one float4 per lane, zero outside the domain / pitch
__device__ __forceinline__ float4 ldrow3(const float* __restrict__ base, int i0, int i1, int x, const G3& g) {
    if (i0 < 0 || i0 >= g.n0 || i1 < 0 || i1 >= g.n1 || x >= g.ld) return f4zero();
    return __ldg(reinterpret_cast<const float4*>(base + (i0 * g.ps + (long long)i1 * g.ld + x)));
}
__device__ __forceinline__ void halo3(const float4& c, const float* __restrict__ base, int i0, int i1, int x0, int lane,
                                      const G3& g, float& left, float& right) {
    left = __shfl_up_sync(0xffffffffu, c.w, 1);
    right = __shfl_down_sync(0xffffffffu, c.x, 1);
    // halo column of the two edge lanes, loaded by every lane (lower half-warp looks left, upper half right) so the
    // load carries no divergent branch and is issued together with the vector loads
    const int xx = lane < 16 ? x0 - 1 : x0 + FW;
    float v = 0.f;
    if (i1 >= 0 && i1 < g.n1 && xx >= 0 && xx < g.n2) v = __ldg(base + (i0 * g.ps + (long long)i1 * g.ld + xx));
    left = lane == 0 ? v : left;
    right = lane == 31 ? v : right;
}
__device__ __forceinline__ float lap7(const float4& C, const float4& U, const float4& D, const float4& F, const float4& Bk,
                                      float l, float r, int e) {
    const float c = f4get(C, e);
    const float w = e == 0 ? l : f4get(C, e - 1), ea = e == 3 ? r : f4get(C, e + 1);
    return (((f4get(U, e) - c) + (f4get(D, e) - c)) + ((ea - c) + (w - c))) + ((f4get(F, e) - c) + (f4get(Bk, e) - c));
}

template <bool ADJ>
__global__ void __launch_bounds__(NT, ADJ ? 3 : 4) acoustic3d_kernel(const A3Args a, int nfx, int nfz) {
    st_pdl_launch_dependents();                             // (no-ops unless launched with programmatic stream serialization)
    st_pdl_wait();
    __shared__ __align__(16) float gsm[ADJ ? NWARP * RZ * FW : 4];
    __shared__ int s_cnt, s_rows[FH];
    const G3 g{a.n0, a.n1, a.n2, a.ld, a.ps};
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // blockIdx.x enumerates (i0, tile) with i0 slowest so neighbouring planes are processed close in time
    const int tiles = nfx * nfz;
    const int i0 = blockIdx.x / tiles, t = blockIdx.x - i0 * tiles;
    const int fz = t / nfx, fx = t - fz * nfx;
    const int x0 = fx * FW, zb0 = fz * FH, z0 = zb0 + warp * RZ;
    const int x = x0 + 4 * lane;
    const bool rows = z0 < g.n1;
    const int zn = min(z0 + RZ, g.n1);
    const bool full = x0 + FW <= g.n2;                       // all four cells of every lane are inside the domain
    const bool want_grad = ADJ && a.gacc != nullptr;
    // bchunk == 1 and every lane's 4 cells inside the domain: the pad columns [n2, ld) of a `full` tile do not exist
    const bool direct = want_grad && a.bchunk == 1 && full;
    float4* gsl = reinterpret_cast<float4*>(gsm + (warp * RZ) * FW + 4 * lane);
    if (want_grad && !direct) {
#pragma unroll
        for (int k = 0; k < RZ; ++k) gsl[k * (FW / 4)] = f4zero();
    }
    const int b_lo = ADJ ? blockIdx.y * a.bchunk : blockIdx.y;
    const int b_hi = ADJ ? min(b_lo + a.bchunk, a.B) : b_lo + 1;
    for (int b = b_lo; b < b_hi; ++b) {
        const long long boff = (long long)b * a.fs;
        const float* cur = (ADJ ? a.lam1 : a.cur) + boff;
        const float* prv = (ADJ ? a.lam2 : a.prev) + boff;
        float* out = (ADJ ? a.lam0 : a.next) + boff;
        const float* S = ADJ ? a.s1 + boff : nullptr;
        if (rows) {
            float4 U = ldrow3(cur, i0, z0 - 1, x, g), C = ldrow3(cur, i0, z0, x, g), D;
            float4 sU = f4zero(), sC = f4zero(), sD = f4zero();
            if (want_grad) { sU = ldrow3(S, i0, z0 - 1, x, g); sC = ldrow3(S, i0, z0, x, g); }
#pragma unroll
            for (int k = 0; k < RZ; ++k) {
                const int z = z0 + k;
                if (z < zn) {
                    D = ldrow3(cur, i0, z + 1, x, g);
                    const float4 F = ldrow3(cur, i0 + 1, z, x, g), Bk = ldrow3(cur, i0 - 1, z, x, g);
                    const float4 P = ldrow3(prv, i0, z, x, g);
                    const float4 CI = ldrow3(a.r, i0, z, x, g), AL = ldrow3(a.b, i0, z, x, g);
                    float l, r;
                    halo3(C, cur, i0, z, x0, lane, g, l, r);
                    float4 Y;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float c = f4get(C, e);
                        f4set(Y, e, c + f4get(AL, e) * (c - f4get(P, e)) + f4get(CI, e) * lap7(C, U, D, F, Bk, l, r, e));
                    }
                    float* o = out + (i0 * g.ps + (long long)z * g.ld + x);
                    if (full) {
                        *reinterpret_cast<float4*>(o) = Y;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (x + e < g.n2) o[e] = f4get(Y, e);
                            else if (x + e < g.ld) o[e] = 0.f;
                        }
                    }
                    if (want_grad) {
                        sD = ldrow3(S, i0, z + 1, x, g);
                        const float4 sF = ldrow3(S, i0 + 1, z, x, g), sB = ldrow3(S, i0 - 1, z, x, g);
                        float sl, sr;
                        halo3(sC, S, i0, z, x0, lane, g, sl, sr);
                        // w1 * lap7(S_i); the time-invariant factor 1/ciso is applied once by the caller
                        if (direct) {
                            // one shot per block: straight read-modify-write of the gradient plane (128-bit), no staging
                            float* go = a.gacc + (long long)blockIdx.y * a.fs + (i0 * g.ps + (long long)z * g.ld + x);
                            float4 acc = *reinterpret_cast<const float4*>(go);
#pragma unroll
                            for (int e = 0; e < 4; ++e) f4set(acc, e, f4get(acc, e) + f4get(C, e) * lap7(sC, sU, sD, sF, sB, sl, sr, e));
                            *reinterpret_cast<float4*>(go) = acc;
                        } else {
                            float4 acc = gsl[k * (FW / 4)];
#pragma unroll
                            for (int e = 0; e < 4; ++e) f4set(acc, e, f4get(acc, e) + f4get(C, e) * lap7(sC, sU, sD, sF, sB, sl, sr, e));
                            gsl[k * (FW / 4)] = acc;
                        }
                        sU = sC; sC = sD;
                    }
                    U = C; C = D;
                }
            }
        }
        __syncthreads();
        // ---- source term: forward adds the wavelet sample; adjoint reads d loss / d amplitude = Lam_i(src)
        for (int s = tid; s < a.ns; s += NT) {
            if (a.src_b[s] != b) continue;
            const int s0 = a.src_i0[s], s1 = a.src_i1[s], s2 = a.src_i2[s];
            if (s0 == i0 && s1 >= zb0 && s1 < zb0 + FH && s2 >= x0 && s2 < x0 + FW) {
                const long long q = s0 * g.ps + (long long)s1 * g.ld + s2;
                if (!ADJ) atomicAdd(out + q, a.amp[s]);
            }
        }
        // ---- receivers of this tile: rows found by one thread each, then served by the block
        const float* rsrc = ADJ ? a.rec_adj : a.rec_out;
        if (rsrc) {
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            if (tid < FH && zb0 + tid < g.n1) {
                const long long row = ((long long)b * g.n0 + i0) * g.n1 + zb0 + tid;
                if (a.row_start[row + 1] > a.row_start[row]) s_rows[atomicAdd(&s_cnt, 1)] = zb0 + tid;
            }
            __syncthreads();
            const int cnt = s_cnt;
            for (int k = 0; k < cnt; ++k) {
                const int z = s_rows[k];
                const long long row = ((long long)b * g.n0 + i0) * g.n1 + z;
                const int lo = a.row_start[row], hi = a.row_start[row + 1];
                for (int r = lo + tid; r < hi; r += NT) {
                    const int rc = a.rec_col[r];
                    if (rc >= x0 && rc < x0 + FW) {
                        const long long q = i0 * g.ps + (long long)z * g.ld + rc;
                        if (ADJ) atomicAdd(out + q, __ldg(a.r + q) * a.rec_adj[a.rec_orig[r]]);     // w += ciso * dL/drec
                        else a.rec_out[a.rec_orig[r]] = out[q];
                    }
                }
            }
        }
        if (ADJ && a.gamp) {
            __syncthreads();
            for (int s = tid; s < a.ns; s += NT) {
                if (a.src_b[s] != b) continue;
                const int s0 = a.src_i0[s], s1 = a.src_i1[s], s2 = a.src_i2[s];
                if (s0 == i0 && s1 >= zb0 && s1 < zb0 + FH && s2 >= x0 && s2 < x0 + FW) {
                    const long long q = s0 * g.ps + (long long)s1 * g.ld + s2;
                    const float ci = __ldg(a.r + q);
                    a.gamp[s] = ci != 0.f ? out[q] / ci : 0.f;
                }
            }
        }
        __syncthreads();
    }
    if (want_grad && !direct && rows && x < g.ld) {
        float* gb = a.gacc + (long long)blockIdx.y * a.fs;
#pragma unroll
        for (int k = 0; k < RZ; ++k) {
            const int z = z0 + k;
            if (z < zn) {
                float* o = gb + (i0 * g.ps + (long long)z * g.ld + x);
                const float4 acc = gsl[k * (FW / 4)];
                if (full) {
                    float4 v = *reinterpret_cast<float4*>(o);
                    v.x += acc.x; v.y += acc.y; v.z += acc.z; v.w += acc.w;
                    *reinterpret_cast<float4*>(o) = v;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (x + e < g.n2) o[e] += f4get(acc, e);
                }
            }
        }
    }
}

}  // namespace

int st_acoustic3d_launch_forward(const A3Args& a, cudaStream_t st) {
    const int nfx = (a.n2 + FW - 1) / FW, nfz = (a.n1 + FH - 1) / FH;
    dim3 grid(a.n0 * nfx * nfz, a.B);
    return st_pdl_launch(acoustic3d_kernel<false>, grid, dim3(NT), 0, st, a, nfx, nfz) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

int st_acoustic3d_launch_adjoint(const A3Args& a, cudaStream_t st) {
    const int nfx = (a.n2 + FW - 1) / FW, nfz = (a.n1 + FH - 1) / FH;
    const int nchunk = (a.B + a.bchunk - 1) / a.bchunk;
    dim3 grid(a.n0 * nfx * nfz, nchunk);
    return st_pdl_launch(acoustic3d_kernel<true>, grid, dim3(NT), 0, st, a, nfx, nfz) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}
