// Persistent multi-timestep kernels for small grids (acoustic PML: equations2d/acoustic.py:73-86, the time loop of
// rnn.py:178-205 with source.py:47-57 and probe.py:42-44 inside) -- ONE launch advances `nsteps` time steps.
//
// The per-step kernels are launch-latency bound on a cfg1-class grid (250 x 400 cells: 7 us per step for 2 us of
// work).  Here one thread-block CLUSTER owns one shot for the whole time loop:
//   * the padded domain is cut into row strips, one per CTA of the cluster (up to 16 CTAs, non-portable size);
//   * every thread owns 4 consecutive columns of RPW rows and keeps their current / previous field values and
//     their two coefficients in REGISTERS for all time steps; the field never goes back to HBM unless a wavefield
//     history is requested (gradient runs) -- then it leaves as fire-and-forget 128-bit stores;
//   * the current field is published to a triple-buffered shared-memory copy once per step: x-neighbours come
//     from warp shuffles (edge lanes read one scalar of the published copy), z-neighbours inside a thread's rows
//     from registers, the two halo rows from the published copy -- of the neighbouring CTA through distributed
//     shared memory (ld.shared::cluster) at the strip boundaries;
//   * one barrier.cluster arrive/wait pair per time step is the only synchronisation; the history store and the
//     receiver gather of the previous step sit between arrive and wait;
//   * source add by the thread that owns the source cell (before publishing), receiver gather from the published
//     copy by the first threads of the CTA (cached (cell, record) pairs).
// The per-cell arithmetic is the same expression as in wave2d_forward_kernel<ISO|PML> (st_wave2d.cu,
// forward_fast_rows), so the records agree bit for bit with the per-step kernels.
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>

#include "st_wave2d.cuh"
#include "st_wave2d_persist.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int PW = 128;                 // columns per warp (32 lanes x float4)
constexpr int XPAD = 4;                 // zero columns left of column 0 in the published copy (keeps rows 16-byte aligned)
constexpr int RECCAP = 2048;            // cached (cell, record) pairs per CTA
constexpr int NCOPY = 3;                // published copies of the field: S_i is written to copy i % 3 while S_{i-1} is being read by
                                        // the stencil and S_{i-2} by the receiver gather that runs in the barrier shadow
constexpr int PERSIST_SMEM_MAX = 200 * 1024 + 2 * RECCAP * 4;
// dynamic shared memory: NCOPY published copies + the zero row (+ the receiver staging rows of the adjoint) + the receiver cache
__host__ __device__ constexpr int persist_smem(int rpc, int ldp, bool adjoint) {
    return ((NCOPY + (adjoint ? 1 : 0)) * rpc + 1) * ldp * (int)sizeof(float) + 2 * RECCAP * (int)sizeof(int);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned map_rank(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_cluster4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// loads the compiler may not move: issued where they are written (between the barrier's arrive and wait), so the data
// travels while the barrier completes instead of being sunk to the first use in the next step
__device__ __forceinline__ float4 ldg4_pinned(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ldg1_pinned(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__device__ __forceinline__ float f4e(const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4s(float4& v, int e, float s) { if (e == 0) v.x = s; else if (e == 1) v.y = s; else if (e == 2) v.z = s; else v.w = s; }

// y = c + alpha (c - p) + ciso (((n - c) + (s - c)) + ((e - c) + (w - c)))   -- forward_fast_rows<ISO|PML>
__device__ __forceinline__ float pml_update(float c, float p, float n, float s, float w, float e, float alpha, float ciso) {
    const float A = ciso * (((n - c) + (s - c)) + ((e - c) + (w - c)));
    return c + alpha * (c - p) + A;
}

template <int NW, int RPW>
__global__ void __launch_bounds__(NW * 32, 1) wave2d_persist_forward_kernel(const W2Args a, const W2Persist pp) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cluster_rank();
    const int b = blockIdx.x / pp.cs;                       // shot of this cluster
    const W2Geom& g = a.g;
    const int ldp = pp.ldp, rpc = pp.rpc;
    float* pub = reinterpret_cast<float*>(dsm);             // [2][rpc][ldp] published copies
    float* zrow = pub + NCOPY * rpc * ldp;                   // [ldp] zeros: the halo row of a strip at the domain edge
    int* rec_cell = reinterpret_cast<int*>(zrow + ldp);     // [RECCAP] float offset inside one published copy
    int* rec_dst = rec_cell + RECCAP;                       // [RECCAP] record index * nchan
    __shared__ int s_nrec, s_rec_lo, s_rec_hi;

    const int strip = warp % pp.nstrips, rg = warp / pp.nstrips;
    const bool active = rg < pp.nrg;
    const int zc0 = (int)crank * rpc;                       // first row of this CTA
    const int lr0 = rg * RPW;                               // first local row of this thread
    const int x = strip * PW + 4 * lane;
    const long long boff = (long long)b * a.fs;
    const float* u = pp.u;

    // ---- zero the published copies (pads stay zero for the whole run)
    for (int i = tid; i < (NCOPY * rpc + 1) * ldp; i += NW * 32) pub[i] = 0.f;

    // ---- receivers of this CTA's rows -> cached (cell, record) pairs
    if (tid == 0) {
        const int zlo = min(zc0, g.nz), zhi = min(zc0 + rpc, g.nz);
        const bool any = a.rec_out != nullptr && a.R > 0;
        s_rec_lo = any ? a.row_start[b * g.nz + zlo] : 0;
        s_rec_hi = any ? a.row_start[b * g.nz + zhi] : 0;
        s_nrec = min(s_rec_hi - s_rec_lo, RECCAP);
    }
    __syncthreads();
    if (a.rec_out != nullptr && s_rec_hi > s_rec_lo) {
        const int zhi = min(zc0 + rpc, g.nz);
        for (int z = zc0 + warp; z < zhi; z += NW) {
            const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
            for (int r = lo + lane; r < hi; r += 32) {
                const int k = r - s_rec_lo;
                if (k < RECCAP) {
                    rec_cell[k] = (z - zc0) * ldp + XPAD + a.rec_x[r];
                    rec_dst[k] = a.rec_orig[r] * a.nchan;
                }
            }
        }
    }

    // ---- registers: coefficients and the two field states of the owned cells
    float4 cur[RPW], prv[RPW], al[RPW], ci[RPW];
    const float* c_ciso = a.coef[2];
    const float* c_alpha = a.coef[3];
    const long long slotf = pp.slot;                        // floats per state slot
    const float* s_prev = u + slotf * (pp.slot0 % pp.nslots) + boff;
    const float* s_cur = u + slotf * ((pp.slot0 + 1) % pp.nslots) + boff;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int z = zc0 + lr0 + r;
        const bool in = active && z < g.nz && x < g.ld;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        cur[r] = prv[r] = al[r] = ci[r] = zero;
        if (in) {
            const long long o = (long long)z * g.ld + x;
            cur[r] = __ldg(reinterpret_cast<const float4*>(s_cur + o));
            prv[r] = __ldg(reinterpret_cast<const float4*>(s_prev + o));
            al[r] = __ldg(reinterpret_cast<const float4*>(c_alpha + o));
            ci[r] = __ldg(reinterpret_cast<const float4*>(c_ciso + o));
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (x + e >= g.nx) { f4s(al[r], e, 0.f); f4s(ci[r], e, 0.f); f4s(cur[r], e, 0.f); f4s(prv[r], e, 0.f); }
        }
    }
    __syncthreads();                                        // zero fill done before the first publication
    if (active) {                                           // copy 0 holds S_{i0-1}
#pragma unroll
        for (int r = 0; r < RPW; ++r)
            *reinterpret_cast<float4*>(pub + (lr0 + r) * ldp + XPAD + x) = cur[r];
    }
    // sources owned by this thread: found once (a thread owns a source cell or not for the whole run)
    int my_src[2] = {-1, -1}, my_src_rc[2] = {0, 0};
    int nmine = 0;
    bool src_overflow = false;
    if (active) {
        for (int s = 0; s < a.ns; ++s) {
            if (a.src_b[s] != b) continue;
            const int sz = a.src_z[s] - (zc0 + lr0), sx = a.src_x[s] - x;
            if (sz >= 0 && sz < RPW && sx >= 0 && sx < 4 && (a.src_fmask & 1)) {
                if (nmine < 2) { my_src[nmine] = s; my_src_rc[nmine] = sz * 4 + sx; }
                else src_overflow = true;
                ++nmine;
            }
        }
    }
    cluster_arrive();
    cluster_wait();

    const unsigned pub_addr = smem_u32(pub);
    const bool up_remote = lr0 == 0, dn_remote = lr0 + RPW == rpc;
    const bool has_up = !(up_remote && crank == 0), has_dn = !(dn_remote && (int)crank == pp.cs - 1);
    const unsigned up_rank = up_remote ? crank - 1 : crank, dn_rank = dn_remote ? crank + 1 : crank;
    const int up_row = up_remote ? rpc - 1 : lr0 - 1, dn_row = dn_remote ? 0 : lr0 + RPW;
    const bool edge_l = lane == 0, edge_r = lane == 31;
    // addresses of the two halo rows in either published copy (DSMEM window of the neighbour at the strip boundaries)
    // (branch-free: a strip at the domain edge reads the zero row instead, with a zero copy stride)
    const unsigned copy_bytes = (unsigned)(rpc * ldp) * 4u;
    const unsigned zrow_addr = pub_addr + (unsigned)(NCOPY * rpc * ldp + XPAD + x) * 4u;
    const unsigned up_addr = has_up ? map_rank(pub_addr + (unsigned)(up_row * ldp + XPAD + x) * 4u, up_rank) : zrow_addr;
    const unsigned dn_addr = has_dn ? map_rank(pub_addr + (unsigned)(dn_row * ldp + XPAD + x) * 4u, dn_rank) : zrow_addr;
    const unsigned up_stride = has_up ? copy_bytes : 0u, dn_stride = has_dn ? copy_bytes : 0u;
    const int own_off = lr0 * ldp + XPAD + x;               // first owned row inside a published copy
    int slot_w = (pp.slot0 + 2) % pp.nslots;                // slot the next state is stored to
    float amp_next[2] = {0.f, 0.f};                         // wavelet samples of the coming step (owner thread only)
    if (nmine > 0 && !src_overflow && pp.nsteps > 0) {
        for (int q = 0; q < 2; ++q)
            if (q < nmine) amp_next[q] = a.amp[my_src[q]];
    }

    // Receiver gather of step k from the copy that holds S_k (probe.py:42-44).  It runs in the shadow of the NEXT
    // step's barrier (that copy is not overwritten before the barrier after that one).
    auto gather = [&](int k, int copy) {
        if (a.rec_out == nullptr) return;
        float* out = a.rec_out + (long long)k * a.R * a.nchan;
        const float* P = pub + copy * rpc * ldp;
        for (int r = tid; r < s_nrec; r += NW * 32) {
            const float v = P[rec_cell[r]];
            for (int ch = 0; ch < a.nchan; ++ch) out[rec_dst[r] + ch] = v;
        }
        if (s_rec_hi - s_rec_lo > RECCAP) {                   // more receivers than the cache holds: walk the CSR
            const int zhi = min(zc0 + rpc, g.nz);
            for (int z = zc0 + warp; z < zhi; z += NW) {
                const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
                for (int r = max(lo, s_rec_lo + RECCAP) + lane; r < hi; r += 32) {
                    const float v = P[(z - zc0) * ldp + XPAD + a.rec_x[r]];
                    for (int ch = 0; ch < a.nchan; ++ch) out[(long long)a.rec_orig[r] * a.nchan + ch] = v;
                }
            }
        }
    };

    // One time step: C = S_{i-1} (published in copy `pc`), P = S_{i-2}; S_i overwrites P in place (a cell's previous
    // value is needed by that cell only) and is published to the other copy.
    auto step = [&](float4 (&C)[RPW], float4 (&P)[RPW], int k, int pc) {
        const int pn = pc + 1 == NCOPY ? 0 : pc + 1;        // copy S_k is published to
        if (active) {
            const float4 up = ld_cluster4(up_addr + pc * up_stride);
            const float4 dn = ld_cluster4(dn_addr + pc * dn_stride);
            const float* Arow = pub + pc * rpc * ldp + own_off;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const float4 c = C[r];
                const float4 n = r == 0 ? up : C[r - 1];
                const float4 s = r == RPW - 1 ? dn : C[r + 1];
                // x-neighbours: shuffle; the two edge lanes take the published value instead (loaded by every lane:
                // no divergent branch, the whole patch stays one basic block for the scheduler)
                const float lh = Arow[r * ldp - 1], rh = Arow[r * ldp + 4];
                float lc = __shfl_up_sync(0xffffffffu, c.w, 1);
                float rc = __shfl_down_sync(0xffffffffu, c.x, 1);
                lc = edge_l ? lh : lc;
                rc = edge_r ? rh : rc;
                float4 y;
                y.x = pml_update(c.x, P[r].x, n.x, s.x, lc, c.y, al[r].x, ci[r].x);
                y.y = pml_update(c.y, P[r].y, n.y, s.y, c.x, c.z, al[r].y, ci[r].y);
                y.z = pml_update(c.z, P[r].z, n.z, s.z, c.y, c.w, al[r].z, ci[r].z);
                y.w = pml_update(c.w, P[r].w, n.w, s.w, c.z, rc, al[r].w, ci[r].w);
                P[r] = y;
            }
            // source add (source.py:47-57: field += wavelet sample, after the update, before sampling)
            if (nmine > 0) {
                if (!src_overflow) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (q < nmine) {
                            const int rr = my_src_rc[q] >> 2, e = my_src_rc[q] & 3;
#pragma unroll
                            for (int r = 0; r < RPW; ++r)
                                if (r == rr) f4s(P[r], e, f4e(P[r], e) + amp_next[q]);
                        }
                    }
                } else {
                    const float* amp = a.amp + (long long)k * a.ns;
                    for (int s = 0; s < a.ns; ++s) {
                        if (a.src_b[s] != b) continue;
                        const int sz = a.src_z[s] - (zc0 + lr0), sx = a.src_x[s] - x;
                        if (sz >= 0 && sz < RPW && sx >= 0 && sx < 4) {
#pragma unroll
                            for (int r = 0; r < RPW; ++r)
                                if (r == sz) f4s(P[r], sx, f4e(P[r], sx) + amp[s]);
                        }
                    }
                }
            }
            float* Brow = pub + pn * rpc * ldp + own_off;
#pragma unroll
            for (int r = 0; r < RPW; ++r) *reinterpret_cast<float4*>(Brow + r * ldp) = P[r];
        }
        cluster_arrive();
        // ---- in the shadow of the barrier: history store, next wavelet sample, receiver gather of the previous step
        if (active) {
            if (pp.history || k >= pp.nsteps - 2) {         // rolling 3-slot state: only the last two states are kept
                float* dst = pp.u + slotf * slot_w + boff + (long long)(zc0 + lr0) * g.ld + x;
#pragma unroll
                for (int r = 0; r < RPW; ++r)
                    if (zc0 + lr0 + r < g.nz && x < g.ld) *reinterpret_cast<float4*>(dst + (long long)r * g.ld) = P[r];
            }
            if (nmine > 0 && !src_overflow && k + 1 < pp.nsteps) {
                const float* amp = a.amp + (long long)(k + 1) * a.ns;
                for (int q = 0; q < 2; ++q)
                    if (q < nmine) amp_next[q] = amp[my_src[q]];
            }
        }
        slot_w = slot_w + 1 == pp.nslots ? 0 : slot_w + 1;
        if (k > 0) gather(k - 1, pc);
        cluster_wait();
    };

    int k = 0, pc = 0;                                      // copy `pc` holds S_{k-1}
    for (; k + 1 < pp.nsteps; k += 2) {
        step(cur, prv, k, pc);                              // S_k -> prv registers
        pc = pc + 1 == NCOPY ? 0 : pc + 1;
        step(prv, cur, k + 1, pc);                          // S_{k+1} -> cur registers
        pc = pc + 1 == NCOPY ? 0 : pc + 1;
    }
    if (k < pp.nsteps) { step(cur, prv, k, pc); pc = pc + 1 == NCOPY ? 0 : pc + 1; }
    if (pp.nsteps > 0) gather(pp.nsteps - 1, pc);
    cluster_arrive();                                       // nobody leaves while a neighbour may still read its copy
    cluster_wait();
}

// ------------------------------------------------------------------------------------------------ adjoint twin
// The exact adjoint of the acoustic PML step (wave2d_adjoint_kernel<ISO|PML>, adjoint_fast_rows + adjoint_tail):
//     Lam_i = (1 + alpha) L1 + lap(ciso L1) - alpha L2  (+ d loss / d record_i at the receiver cells),   L1 = Lam_{i+1}, L2 = Lam_{i+2}
//     g_ciso += L1 lap(S_i),      d loss / d wavelet_i(s) = Lam_i(source cell s)
// for `nsteps` time steps (descending i) in ONE launch, same decomposition as the forward kernel: the two cotangents, the
// coefficients AND the gradient accumulator of a thread's cells live in registers for the whole loop (the gradient plane
// is read-modify-written once, at the end); the products w = ciso L1 are what neighbours need and what is published to the
// triple-buffered shared-memory copy; S_i (own rows + the two halo rows + the edge columns) comes straight from the
// wavefield history in HBM, loaded one time step ahead of its use; the receiver cotangents of the next step are
// prefetched by the threads that will inject them (shared-memory staging row per CTA, so duplicates add up).
template <int NW, int RPW>
__global__ void __launch_bounds__(NW * 32, 1) wave2d_persist_adjoint_kernel(const W2Args a, const W2Persist pp) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned crank = cluster_rank();
    const int b = blockIdx.x / pp.cs;
    const W2Geom& g = a.g;
    const int ldp = pp.ldp, rpc = pp.rpc;
    float* pub = reinterpret_cast<float*>(dsm);             // [NCOPY][rpc][ldp] published w = ciso * Lam
    float* zrow = pub + NCOPY * rpc * ldp;                   // [ldp] zeros
    float* inj = zrow + ldp;                                 // [rpc][ldp] receiver cotangents of the current step (zero otherwise)
    int* rec_cell = reinterpret_cast<int*>(inj + rpc * ldp);
    int* rec_src = rec_cell + RECCAP;                       // [RECCAP] index into rec_adj of one step (record * nchan)
    __shared__ int s_nrec, s_rec_lo, s_rec_hi;

    const int strip = warp % pp.nstrips, rg = warp / pp.nstrips;
    const bool active = rg < pp.nrg;
    const int zc0 = (int)crank * rpc, lr0 = rg * RPW;
    const int x = strip * PW + 4 * lane;
    const long long boff = (long long)b * a.fs;
    const long long slotf = pp.slot;

    for (int i = tid; i < ((NCOPY + 1) * rpc + 1) * ldp; i += NW * 32) pub[i] = 0.f;
    if (tid == 0) {
        const int zlo = min(zc0, g.nz), zhi = min(zc0 + rpc, g.nz);
        const bool any = a.rec_adj != nullptr && a.R > 0;
        s_rec_lo = any ? a.row_start[b * g.nz + zlo] : 0;
        s_rec_hi = any ? a.row_start[b * g.nz + zhi] : 0;
        s_nrec = s_rec_hi - s_rec_lo;
    }
    __syncthreads();
    const int nrec = s_nrec;                                // (the plan caps it at RECCAP: more -> per-step kernels)
    if (nrec > 0) {
        const int zhi = min(zc0 + rpc, g.nz);
        for (int z = zc0 + warp; z < zhi; z += NW) {
            const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
            for (int r = lo + lane; r < hi; r += 32) {
                const int k = r - s_rec_lo;
                if (k < RECCAP) {
                    rec_cell[k] = (z - zc0) * ldp + XPAD + a.rec_x[r];
                    rec_src[k] = a.rec_orig[r] * a.nchan;
                }
            }
        }
    }

    // ---- registers: cotangents, coefficients, gradient accumulator
    float4 L1[RPW], L2[RPW], al[RPW], ci[RPW], gacc[RPW];
    const float* lam = pp.lam;                              // [3][B][nz][ld] ring: Lam_i lives in slot i mod 3
    const int i_hi = pp.i0;                                  // first (highest) step of this call
    const float* p1 = lam + slotf * ((i_hi + 1) % 3) + boff;
    const float* p2 = lam + slotf * ((i_hi + 2) % 3) + boff;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    bool in[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int z = zc0 + lr0 + r;
        in[r] = active && z < g.nz && x < g.ld;
        L1[r] = L2[r] = al[r] = ci[r] = gacc[r] = zero;
        if (in[r]) {
            const long long o = (long long)z * g.ld + x;
            // the accumulator starts from the plane's running value and is written back at the end: the sum over time
            // steps is then associated the same way however the loop is cut into launches (checkpoint segments)
            if (a.gacc != nullptr) gacc[r] = *reinterpret_cast<const float4*>(a.gacc + ((long long)b * 7 + 1) * ((long long)g.nz * g.ld) + o);
            L1[r] = __ldg(reinterpret_cast<const float4*>(p1 + o));
            L2[r] = __ldg(reinterpret_cast<const float4*>(p2 + o));
            al[r] = __ldg(reinterpret_cast<const float4*>(a.coef[3] + o));
            ci[r] = __ldg(reinterpret_cast<const float4*>(a.coef[2] + o));
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (x + e >= g.nx) { f4s(al[r], e, 0.f); f4s(ci[r], e, 0.f); f4s(L1[r], e, 0.f); f4s(L2[r], e, 0.f); }
        }
    }
    __syncthreads();
    if (active) {                                           // copy 0 holds w(Lam_{i_hi+1})
#pragma unroll
        for (int r = 0; r < RPW; ++r)
            *reinterpret_cast<float4*>(pub + (lr0 + r) * ldp + XPAD + x) =
                make_float4(ci[r].x * L1[r].x, ci[r].y * L1[r].y, ci[r].z * L1[r].z, ci[r].w * L1[r].w);
    }
    // sources owned by this thread (for d loss / d wavelet)
    int my_src[2] = {-1, -1}, my_src_rc[2] = {0, 0};
    int nmine = 0;
    bool src_overflow = false;
    if (active && a.gamp != nullptr) {
        for (int s = 0; s < a.ns; ++s) {
            if (a.src_b[s] != b) continue;
            const int sz = a.src_z[s] - (zc0 + lr0), sx = a.src_x[s] - x;
            if (sz >= 0 && sz < RPW && sx >= 0 && sx < 4 && (a.src_fmask & 1)) {
                if (nmine < 2) { my_src[nmine] = s; my_src_rc[nmine] = sz * 4 + sx; }
                else src_overflow = true;
                ++nmine;
            }
        }
    }
    cluster_arrive();
    cluster_wait();

    const unsigned pub_addr = smem_u32(pub);
    const bool up_remote = lr0 == 0, dn_remote = lr0 + RPW == rpc;
    const bool has_up = !(up_remote && crank == 0), has_dn = !(dn_remote && (int)crank == pp.cs - 1);
    const unsigned up_rank = up_remote ? crank - 1 : crank, dn_rank = dn_remote ? crank + 1 : crank;
    const int up_row = up_remote ? rpc - 1 : lr0 - 1, dn_row = dn_remote ? 0 : lr0 + RPW;
    const bool edge_l = lane == 0, edge_r = lane == 31;
    const unsigned copy_bytes = (unsigned)(rpc * ldp) * 4u;
    const unsigned zrow_addr = pub_addr + (unsigned)(NCOPY * rpc * ldp + XPAD + x) * 4u;
    const unsigned up_addr = has_up ? map_rank(pub_addr + (unsigned)(up_row * ldp + XPAD + x) * 4u, up_rank) : zrow_addr;
    const unsigned dn_addr = has_dn ? map_rank(pub_addr + (unsigned)(dn_row * ldp + XPAD + x) * 4u, dn_rank) : zrow_addr;
    const unsigned up_stride = has_up ? copy_bytes : 0u, dn_stride = has_dn ? copy_bytes : 0u;
    const int own_off = lr0 * ldp + XPAD + x;

    // S_i of one step: own rows, the row above / below, the two edge columns (every lane loads; lanes 0 / 31 use them)
    struct SRows { float4 c[RPW], up, dn; float l[RPW], r[RPW]; };
    const int zt = zc0 + lr0;                               // first owned row (global)
    const int xl = strip * PW - 1, xr = strip * PW + PW;
    // (offsets and validity of the ten loads are fixed for the whole loop: computed once)
    const long long o_own = (long long)zt * g.ld + x;
    const bool v_up = active && zt - 1 >= 0 && zt - 1 < g.nz && x < g.ld, v_dn = active && zt + RPW < g.nz && x < g.ld;
    const bool v_l = active && xl >= 0, v_r = active && xr < g.nx;
    int hslot = ((pp.slot0 % pp.nslots) + pp.nslots) % pp.nslots;      // history slot of the NEXT S to load (walks down, wraps)
    auto load_s = [&](int) {
        SRows q;
#if defined(ST_PADJ_DBG) && (ST_PADJ_DBG & 1)
        q.up = q.dn = zero;
        for (int r = 0; r < RPW; ++r) { q.c[r] = zero; q.l[r] = q.r[r] = 0.f; }
        return q;
#endif
        const float* S = pp.u + slotf * hslot + boff + o_own;
        hslot = hslot == 0 ? pp.nslots - 1 : hslot - 1;
        q.up = q.dn = zero;
        if (v_up) q.up = ldg4_pinned(S - g.ld);
        if (v_dn) q.dn = ldg4_pinned(S + RPW * g.ld);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            q.c[r] = zero; q.l[r] = 0.f; q.r[r] = 0.f;
            if (in[r]) q.c[r] = ldg4_pinned(S + r * g.ld);
            if (v_l && zt + r < g.nz) q.l[r] = ldg1_pinned(S + r * g.ld + (xl - x));
            if (v_r && zt + r < g.nz) q.r[r] = ldg1_pinned(S + r * g.ld + (xr - x));
        }
        return q;
    };
    const int ncached = min(nrec, RECCAP);
    auto load_rec = [&](int k) {                            // the first cached record of this thread, one step ahead
        float v = 0.f;
        if (tid < ncached && k < pp.nsteps) {
            const float* ra = a.rec_adj - (long long)k * a.R * a.nchan;     // a.rec_adj points at step i_hi
            for (int ch = 0; ch < a.nchan; ++ch) v += ra[rec_src[tid] + ch];
        }
        return v;
    };
    // records beyond one per thread / beyond the cache: loaded on the spot (rare: dense receiver carpets)
    auto stage_rest = [&](int k) {
        const float* ra = a.rec_adj - (long long)k * a.R * a.nchan;
        for (int t = tid + NW * 32; t < ncached; t += NW * 32) {
            float v = 0.f;
            for (int ch = 0; ch < a.nchan; ++ch) v += ra[rec_src[t] + ch];
            atomicAdd(inj + rec_cell[t], v);
        }
        if (nrec > RECCAP) {
            const int zhi = min(zc0 + rpc, g.nz);
            for (int z = zc0 + warp; z < zhi; z += NW) {
                const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
                for (int r = max(lo, s_rec_lo + RECCAP) + lane; r < hi; r += 32) {
                    float v = 0.f;
                    for (int ch = 0; ch < a.nchan; ++ch) v += ra[(long long)a.rec_orig[r] * a.nchan + ch];
                    atomicAdd(inj + (z - zc0) * ldp + XPAD + a.rec_x[r], v);
                }
            }
        }
    };
    // (An L2 prefetch of the history rows 6 steps ahead was measured and changed nothing: the twin is not waiting for
    // HBM but for its own dependent arithmetic -- two Laplacians per cell at 2-4 warps per scheduler.)
    SRows Sn = load_s(0);
    float rec_next = load_rec(0);
    int slot_w = ((i_hi % 3) + 3) % 3;                      // ring slot Lam_i is stored to
    auto step = [&](float4 (&L1)[RPW], float4 (&L2)[RPW], int k, int pc) {
        const int pn = pc + 1 == NCOPY ? 0 : pc + 1;
        const SRows Sc = Sn;
        // ---- stage the receiver cotangents of this step (CTA-uniform: only CTAs that hold receivers)
        if (nrec > 0) {
            if (tid < ncached) atomicAdd(inj + rec_cell[tid], rec_next);
            if (ncached > NW * 32 || nrec > RECCAP) stage_rest(k);
            __syncthreads();
        }
        if (active) {
            const float4 up = ld_cluster4(up_addr + pc * up_stride);
            const float4 dn = ld_cluster4(dn_addr + pc * dn_stride);
            const float* Arow = pub + pc * rpc * ldp + own_off;
            float4 w[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r) w[r] = make_float4(ci[r].x * L1[r].x, ci[r].y * L1[r].y, ci[r].z * L1[r].z, ci[r].w * L1[r].w);
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const float4 c = w[r];
                const float4 n = r == 0 ? up : w[r > 0 ? r - 1 : 0];
                const float4 s = r == RPW - 1 ? dn : w[r < RPW - 1 ? r + 1 : r];
                const float lh = Arow[r * ldp - 1], rh = Arow[r * ldp + 4];
                float wl = __shfl_up_sync(0xffffffffu, c.w, 1);
                float wr = __shfl_down_sync(0xffffffffu, c.x, 1);
                wl = edge_l ? lh : wl;
                wr = edge_r ? rh : wr;
                // gradient operand: lap(S_i)
                const float4 sc = Sc.c[r];
                const float4 sn = r == 0 ? Sc.up : Sc.c[r - 1];
                const float4 ss = r == RPW - 1 ? Sc.dn : Sc.c[r + 1];
                float sl = __shfl_up_sync(0xffffffffu, sc.w, 1);
                float sr = __shfl_down_sync(0xffffffffu, sc.x, 1);
                sl = edge_l ? Sc.l[r] : sl;
                sr = edge_r ? Sc.r[r] : sr;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float cc = f4e(c, e);
                    const float we = e == 0 ? wl : f4e(c, e - 1), ea = e == 3 ? wr : f4e(c, e + 1);
                    const float lapw = ((f4e(n, e) - cc) + (f4e(s, e) - cc)) + ((ea - cc) + (we - cc));
                    const float alpha = f4e(al[r], e), l1c = f4e(L1[r], e);
                    f4s(L2[r], e, (1.f + alpha) * l1c + lapw - alpha * f4e(L2[r], e));      // Lam_i overwrites Lam_{i+2} in place
                    const float scc = f4e(sc, e);
                    const float swv = e == 0 ? sl : f4e(sc, e - 1), sev = e == 3 ? sr : f4e(sc, e + 1);
                    const float laps = ((f4e(sn, e) - scc) + (f4e(ss, e) - scc)) + ((sev - scc) + (swv - scc));
#if !(defined(ST_PADJ_DBG) && (ST_PADJ_DBG & 2))
                    f4s(gacc[r], e, f4e(gacc[r], e) + l1c * laps);
#endif
                }
            }
            if (nrec > 0) {                                   // + d loss / d record_i, then clear the staging rows
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    float4* ip = reinterpret_cast<float4*>(inj + own_off + r * ldp);
                    const float4 v = *ip;
                    L2[r].x += v.x; L2[r].y += v.y; L2[r].z += v.z; L2[r].w += v.w;
                    *ip = zero;
                }
            }
#pragma unroll
            for (int r = 0; r < RPW; ++r) {                 // cells past nx stay zero (zero coefficients, zero injection)
                *reinterpret_cast<float4*>(pub + pn * rpc * ldp + own_off + r * ldp) =
                    make_float4(ci[r].x * L2[r].x, ci[r].y * L2[r].y, ci[r].z * L2[r].z, ci[r].w * L2[r].w);
            }
        }
        cluster_arrive();
        // ---- barrier shadow: d loss / d wavelet, the last two cotangents to the ring, prefetch of the next step's operands
        if (active) {
            if (nmine > 0) {
                float* gamp = a.gamp - (long long)k * a.ns;   // a.gamp points at step i_hi
                if (!src_overflow) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (q < nmine) {
                            const int rr = my_src_rc[q] >> 2, e = my_src_rc[q] & 3;
                            float v = 0.f;
#pragma unroll
                            for (int r = 0; r < RPW; ++r)
                                if (r == rr) v = f4e(L2[r], e);
                            gamp[my_src[q]] = v;
                        }
                    }
                } else {
                    for (int s = 0; s < a.ns; ++s) {
                        if (a.src_b[s] != b) continue;
                        const int sz = a.src_z[s] - (zc0 + lr0), sx = a.src_x[s] - x;
                        if (sz >= 0 && sz < RPW && sx >= 0 && sx < 4) {
                            float v = 0.f;
#pragma unroll
                            for (int r = 0; r < RPW; ++r)
                                if (r == sz) v = f4e(L2[r], sx);
                            gamp[s] = v;
                        }
                    }
                }
            }
            if (k >= pp.nsteps - 2) {
                float* dst = pp.lam + slotf * slot_w + boff + (long long)zt * g.ld + x;
#pragma unroll
                for (int r = 0; r < RPW; ++r)
                    if (in[r]) *reinterpret_cast<float4*>(dst + (long long)r * g.ld) = L2[r];
            }
        }
        slot_w = slot_w == 0 ? 2 : slot_w - 1;
        if (k + 1 < pp.nsteps) { Sn = load_s(k + 1); rec_next = load_rec(k + 1); }
        cluster_wait();
    };
    {
        int k = 0, pc = 0;
        for (; k + 1 < pp.nsteps; k += 2) {
            step(L1, L2, k, pc);                             // Lam_i -> L2 registers
            pc = pc + 1 == NCOPY ? 0 : pc + 1;
            step(L2, L1, k + 1, pc);                         // Lam_{i-1} -> L1 registers
            pc = pc + 1 == NCOPY ? 0 : pc + 1;
        }
        if (k < pp.nsteps) step(L1, L2, k, pc);
    }
    // ---- the gradient plane of this shot: read once at the start, written once here
    if (a.gacc != nullptr) {
        float* gb = a.gacc + ((long long)b * 7 + 1) * ((long long)g.nz * g.ld) + (long long)zt * g.ld + x;     // slot 1: d/d ciso
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            if (in[r]) *reinterpret_cast<float4*>(gb + (long long)r * g.ld) = gacc[r];      // (cells past nx: zero + zero)
        }
    }
    cluster_arrive();
    cluster_wait();
}

template <bool ADJ, int NW, int RPW>
int launch_persist(const W2Args& a, W2Persist pp, cudaStream_t st) {
    auto kern = ADJ ? wave2d_persist_adjoint_kernel<NW, RPW> : wave2d_persist_forward_kernel<NW, RPW>;
    const int smem = persist_smem(pp.rpc, pp.ldp, ADJ);
    // the limit is raised once per device to the largest size any plan may ask for (the plan caps it)
    cudaError_t e = ADJ ? st_set_max_smem<wave2d_persist_adjoint_kernel<NW, RPW>>(PERSIST_SMEM_MAX)
                        : st_set_max_smem<wave2d_persist_forward_kernel<NW, RPW>>(PERSIST_SMEM_MAX);
    if (e != cudaSuccess) return ST_ERR_CUDA;
    if (pp.cs > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return ST_ERR_CUDA;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(a.B * pp.cs));
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pp.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (pp.probe) {                                         // can one cluster of this shape be resident at all?
        int ncl = 0;
        const cudaError_t eo = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
        if (eo != cudaSuccess || ncl < 1) {
            if (getenv("SEISTORCH_B200_PERSIST_DEBUG"))
                fprintf(stderr, "[st_wave2d_persist] no resident cluster: cs=%d smem=%d threads=%d -> %s, %d clusters\n", pp.cs, smem,
                        NW * 32, cudaGetErrorString(eo), ncl);
            cudaGetLastError();
            return ST_PERSIST_NA;
        }
        return ST_OK;
    }
    return cudaLaunchKernelEx(&cfg, kern, a, pp) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

}  // namespace

// Plans the decomposition; ST_PERSIST_NA when the problem is not of the class this kernel serves.
int st_wave2d_persist_plan(int flags, const W2Args& a, bool adjoint, W2Persist& pp) {
    memset(&pp, 0, sizeof(pp));
    if (flags != (ST_F_ISO | ST_F_PML)) return ST_PERSIST_NA;
    if (a.nchan > 4 || (a.src_fmask & ~1)) return ST_PERSIST_NA;
    const W2Geom& g = a.g;
    // thread shape: 0 = 16 warps x 4 rows per thread (128 registers), 1 = 32 warps x 2 rows (64 registers).  The forward
    // kernel is 2-4 % faster with 1; the adjoint keeps five register arrays per row: 16 x 4 (a few spills) measured 5 % faster
    // than 8 warps x 8 rows (255 registers, no spills).
    const char* ev = getenv("SEISTORCH_B200_PERSIST_VARIANT");
    const char* eva = getenv("SEISTORCH_B200_PERSIST_ADJ_VARIANT");     // adjoint: 0 = 16 x 4 (default), 2 = 8 warps x 8 rows
    pp.variant = adjoint ? (eva && *eva ? atoi(eva) : 0) : (ev && *ev ? atoi(ev) : 1);
    const int NW = pp.variant == 1 ? 32 : pp.variant == 2 ? 8 : 16, RPW = pp.variant == 1 ? 2 : pp.variant == 2 ? 8 : 4;
    pp.nstrips = (g.ld + PW - 1) / PW;
    if (pp.nstrips < 1 || pp.nstrips > NW) return ST_PERSIST_NA;
    pp.nrg = NW / pp.nstrips;
    pp.rpc = pp.nrg * RPW;
    pp.cs = 0;
    for (int cs = 1; cs <= 16; cs *= 2)
        if (cs * pp.rpc >= g.nz) { pp.cs = cs; break; }
    if (pp.cs == 0) return ST_PERSIST_NA;
    pp.ldp = pp.nstrips * PW + 2 * XPAD;
    if (persist_smem(pp.rpc, pp.ldp, adjoint) > PERSIST_SMEM_MAX) return ST_PERSIST_NA;
    return ST_OK;
}

int st_wave2d_persist_forward(const W2Args& a, const W2Persist& pp, cudaStream_t st) {
    return pp.variant == 1 ? launch_persist<false, 32, 2>(a, pp, st) : launch_persist<false, 16, 4>(a, pp, st);
}

int st_wave2d_persist_adjoint(const W2Args& a, const W2Persist& pp, cudaStream_t st) {
    return pp.variant == 2 ? launch_persist<true, 8, 8>(a, pp, st) : launch_persist<true, 16, 4>(a, pp, st);
}
