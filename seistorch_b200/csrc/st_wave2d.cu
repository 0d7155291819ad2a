// sm_100a kernels for the 2D second-order wave-equation family:
//   acoustic (PML), acoustic_habc, vti_habc2, tti_habc, acoustic_fwim_habc, acoustic_rho_habc,
//   acoustic_lsrtm_habc, acoustic_{vti,tti}_lsrtm_habc   -- one template, eight flag sets.
//
// One launch = one time step of every shot in the batch, with the source add and the
// receiver gather fused in (reference: ~30-190 ATen launches per step).
//
// Two kernel families:
//
// (1) TMA kernels (wave2d_forward_tma_kernel / wave2d_adjoint_tma_kernel; acoustic and acoustic_habc):
//     16 x 128 tiles, (tile, shot) items pulled through an mbarrier ring of cp.async.bulk.tensor boxes
//     (halos and the grid edge come from the box: out-of-range elements are zero-filled), consumed as
//     128-bit shared-memory rows with warp-shuffle x-neighbours; the straight sides of the absorbing
//     frame in closed form (tile kinds), the four corners as generic tiles; programmatic dependent
//     launch between time steps.  See DESIGN.md section 5.
//
// (2) Register kernels (wave2d_forward_kernel / wave2d_adjoint_kernel; every flag set), two kinds of
//     thread blocks per launch:
//   * FAST blocks stream the non-frame cells: a warp owns a 128-column x 4-row tile, every
//     lane 4 consecutive cells; rows are loaded once as 128-bit vectors and marched through
//     a 3-row register pipeline, left/right neighbours come from warp shuffles (plus one
//     predicated halo load per edge lane), no shared memory.
//   * FRAME blocks evaluate the absorbing frame: a precomputed-tap gather (st_wave2d_band.cuh) for
//     the single-field equations without mixed derivative, else cell by cell from a shared-memory
//     tile with a 2-cell halo (one-way blend, side ownership, wrap-around strip neighbours).
// Block kinds write disjoint cells; whoever stores a cell also applies the source
// add / receiver gather for it, so no inter-block ordering is needed.
//
// Data layout in HBM: fields [NF][B][nz][ld] fp32, row pitch ld a multiple of 4 so that
// every row starts 16-byte aligned and columns [nx, ld) stay zero; coefficient planes
// [nz][ld] shared by all shots (L2-resident: <= 8 MB each at the BASELINE sizes).
#include <cstdlib>
#include <cstring>

#include "st_wave2d.cuh"
#include "st_wave2d_band.cuh"

namespace {

constexpr int NT = 256;                     // threads per block (both block kinds)
// ---- frame (general) tiles
constexpr int TX = 64, TZ = 32, HALO = 2;
constexpr int SW = TX + 2 * HALO, SH = TZ + 2 * HALO;
constexpr int NTX = 64, NTY = 4, RPT = TZ / NTY;
// ---- fast tiles
constexpr int FW = 128;                     // columns per warp (32 lanes x float4)
#ifndef ST_FRZ
#define ST_FRZ 2                            // rows per warp of the fast (register / shuffle) tiles.  2 beats 4 and 8 on every adjoint
#endif                                      // (more, smaller blocks: the kernels are latency-bound; measured)
#ifndef ST_ADJ_MINB
#define ST_ADJ_MINB 3
#endif
#ifndef ST_ADJ_MINB_XZ
#define ST_ADJ_MINB_XZ 3                    // blocks / SM of the tti_habc adjoint
#endif
#ifndef ST_ADJ_MINB_BORN
#define ST_ADJ_MINB_BORN 2                  // blocks / SM of the Born-pair adjoint (3 was measured: see DESIGN.md)
#endif
constexpr int FRZ = ST_FRZ;                 // rows per warp
constexpr int NWARP = NT / 32;
constexpr int FH = FRZ * NWARP;             // rows per fast block
#ifndef ST_BAND_SHOTS
#define ST_BAND_SHOTS 12
#endif
constexpr int BSH = ST_BAND_SHOTS;          // most shots a band thread walks with its taps in registers (B200, 12 shots 600x1300,
                                            // adjoint of the VTI Born pair: groups of 4 / 6 / 8+4 / 12 shots: 282 / 267 / 268 / 251 us)
// the shots of a launch are split into equal groups of at most BSH
__host__ __device__ __forceinline__ int band_groups(int B) { return (B + BSH - 1) / BSH; }
__host__ __device__ __forceinline__ int band_group_shots(int B) { const int n = band_groups(B); return (B + n - 1) / n; }
#ifndef ST_DBG_SKIP
#define ST_DBG_SKIP 0                       // tuning only: bit 0 / 1 = frame / fast blocks of the register adjoint kernel return at once, bit 3 = TMA blocks return at once, bit 4 = all tiles as kind 0, bit 5 = corner blocks return
#endif
// ---- TMA-staged tiles (st_wave2d.cuh: W2Tma)
constexpr int TC = ST_TMA_TC, TR = ST_TMA_TR, HC = ST_TMA_HC, H1R = ST_TMA_H1, H2R = ST_TMA_H2, XO = ST_TMA_XO;
static_assert(TC == FW && TR == 2 * NWARP && FH % TR == 0, "TMA tile = one float4 per lane, two rows per warp");
constexpr int TMA_H1_BYTES = (H1R * HC * 4 + 127) / 128 * 128;       // 9856  (box: 9792)
constexpr int TMA_H2_BYTES = (H2R * HC * 4 + 127) / 128 * 128;       // 10880
constexpr int TMA_CORE_BYTES = TR * TC * 4;                          // 8192
#ifndef ST_CORNER_UNROLL
#define ST_CORNER_UNROLL 1                   // cells in flight per thread in the generic corner tiles of the TMA kernels
#endif
#ifndef ST_TMA_FWD_STAGES
#define ST_TMA_FWD_STAGES 3
#endif
#ifndef ST_TMA_ADJ_STAGES
#define ST_TMA_ADJ_STAGES 2
#endif
// stage layouts (sized for the frame tiles; interior tiles use smaller boxes at the same offsets)
//   forward: [cur: 2-deep halo][prev: 1-deep halo]          interior: [cur: 1-deep halo][prev: core]
//   adjoint: [Lam1: 2-deep][S_i: 2-deep][Lam2: 1-deep][S_{i-1}: 1-deep]   interior: [Lam1: 1-deep][S_i: 1-deep][Lam2: core]
// (PML has frame-free tiles only: the regions shrink to the interior boxes and more blocks fit an SM)
template <int FL> __host__ __device__ constexpr int tma_r0() { return (FL & ST_F_HABC) ? TMA_H2_BYTES : TMA_H1_BYTES; }   // first region
template <int FL> __host__ __device__ constexpr int tma_fwd_stage() { return (FL & ST_F_HABC) ? TMA_H2_BYTES + TMA_H1_BYTES : TMA_H1_BYTES + TMA_CORE_BYTES; }
template <int FL> __host__ __device__ constexpr int tma_adj_stage() { return (FL & ST_F_HABC) ? 2 * TMA_H2_BYTES + 2 * TMA_H1_BYTES : 2 * TMA_H1_BYTES + TMA_CORE_BYTES; }
template <int FL> __host__ __device__ constexpr int tma_fwd_smem() { return ST_TMA_FWD_STAGES * tma_fwd_stage<FL>(); }
// adjoint ring: ST_TMA_ADJ_STAGES frame-tile stages or one more of the (smaller) frame-free stages in the same memory
constexpr int TMA_ADJ_STAGE0 = 2 * TMA_H1_BYTES + TMA_CORE_BYTES;   // frame-free tile: Lam1 + S_i (1-deep halo) + Lam2 (core)
template <int FL> __host__ __device__ constexpr int tma_adj_smem() {
    constexpr int fr = ST_TMA_ADJ_STAGES * tma_adj_stage<FL>(), in = (ST_TMA_ADJ_STAGES + 1) * TMA_ADJ_STAGE0;
    return (FL & ST_F_HABC) ? (fr > in ? fr : in) : ST_TMA_ADJ_STAGES * tma_adj_stage<FL>();
}
// resident blocks per SM the TMA kernels are compiled for
#ifndef ST_TMA_FWD_MINB_HABC
#define ST_TMA_FWD_MINB_HABC 3
#endif
template <int FL> __host__ __device__ constexpr int tma_fwd_minb() { return (FL & ST_F_HABC) ? ST_TMA_FWD_MINB_HABC : 4; }
#ifndef ST_TMA_ADJ_MINB_HABC
#define ST_TMA_ADJ_MINB_HABC 2
#endif
template <int FL> __host__ __device__ constexpr int tma_adj_minb() { return (FL & ST_F_HABC) ? ST_TMA_ADJ_MINB_HABC : 3; }
// flag sets with a TMA path
template <int FL>
__host__ __device__ constexpr bool tma_ok() { return FL == (ST_F_ISO | ST_F_PML) || FL == (ST_F_ISO | ST_F_HABC); }

// A TMA block streams a CHUNK of up to `tpb` consecutive tiles of one kind (along x inside a tile row of the
// column band, along z inside a side column) x `tsh` shots through its ring.
__host__ __device__ inline int tma_side_chunks(const W2Tma& tm) { return tm.sr1 > tm.sr0 ? tm.ntr : 0; }     // one tile each, every tile row
__host__ __device__ inline int tma_row_chunks(const W2Tma& tm) { return (tm.tx1 - tm.tx0 + tm.tpb - 1) / tm.tpb; }
__host__ __device__ inline int tma_chunks(const W2Tma& tm) { return 2 * tma_side_chunks(tm) + tm.ntr * tma_row_chunks(tm); }
// blocks: [acquisition-row tiles x shots (one tile, one shot each)] ++ [chunks x shot groups]
__host__ __device__ inline int tma_acq_blocks(const W2Tma& tm, int B) { return tm.ar1 > tm.ar0 ? (tm.ar1 - tm.ar0) * (tm.tx1 - tm.tx0 + 2) * B : 0; }
__host__ __device__ inline int tma_blocks(const W2Tma& tm, int B) {
    return tm.enabled ? tma_acq_blocks(tm, B) + tma_chunks(tm) * ((B + tm.tsh - 1) / tm.tsh) : 0;
}
struct TmaChunk { int z0, x0, kind, ntile, dz, dx, b_lo, nsh, plane, mx0, mx1; };    // ntile == 0: nothing to do; columns [mx0, mx1) are stored
// c-th chunk -> first tile origin, kind (0 frame-free, +1/-1 top/bottom frame, +2/-2 left/right frame), tile
// count and step.  Heaviest first: side chunks, bottom-frame rows, top-frame rows, then the frame-free rows.
// first column of the right-hand generic corner tiles
__host__ __device__ inline int corner_xr(const W2Tma& tm, const W2Geom& g) { return g.nx - TX > tm.tx1 * FW ? g.nx - TX : tm.tx1 * FW; }
__device__ __forceinline__ int tma_row_kind(const W2Tma& tm, const W2Geom& g, bool habc, int z0) {
    if (!habc || (ST_DBG_SKIP & 16)) return 0;
    if (!g.multiple && z0 < tm.band) return 1;
    if (z0 + TR > g.nz - tm.band) return -1;
    return 0;
}
// side tile of tile row `tr` (side 0 left, 1 right): rows [sr0, sr1) are whole tiles of the straight left / right
// frame; the corner rows keep only the columns next to the generic corner tile (straight top / bottom frame or
// frame-free cells), selected by a column mask
__device__ __forceinline__ void tma_side_tile(const W2Tma& tm, const W2Geom& g, bool habc, int nfx, int tr, int side, TmaChunk& q) {
    q.z0 = tr * TR;
    q.x0 = side ? (nfx - 1) * FW : 0;
    q.ntile = 1; q.dz = q.dx = 0;
    if (tr >= tm.sr0 && tr < tm.sr1) {
        q.kind = side ? -2 : 2;
    } else {
        q.kind = tma_row_kind(tm, g, habc, q.z0);
        if (side) { q.mx0 = q.x0; q.mx1 = corner_xr(tm, g); } else { q.mx0 = TX; q.mx1 = FW; }
        if (q.mx1 <= q.mx0) q.ntile = 0;
    }
    if (ST_DBG_SKIP & 16) q.kind = 0;
}
__device__ __forceinline__ TmaChunk tma_block_decode(const W2Tma& tm, const W2Geom& g, bool habc, int nfx, int B, int bid) {
    TmaChunk q;
    q.mx0 = 0; q.mx1 = 1 << 30;
    const int ntx = tm.tx1 - tm.tx0;
    const int nacq = tma_acq_blocks(tm, B);
    if (bid < nacq) {                                       // (shot, acquisition tile row, tile): one tile, one shot
        const int per = ntx + 2, nar = tm.ar1 - tm.ar0;
        const int b = bid / (nar * per), rem = bid - b * nar * per;
        const int tr = tm.ar0 + rem / per, i = rem % per;
        q.b_lo = b; q.nsh = 1; q.plane = b;
        if (i < ntx) {
            q.z0 = tr * TR;
            q.ntile = 1; q.dz = q.dx = 0;
            q.x0 = (tm.tx0 + i) * FW;
            q.kind = tma_row_kind(tm, g, habc, q.z0);
        } else if (tm.sr1 > tm.sr0) {
            tma_side_tile(tm, g, habc, nfx, tr, i - ntx, q);
        } else {
            q.ntile = 0;
        }
        return q;
    }
    bid -= nacq;
    const int nchunk = tma_chunks(tm);
    const int grp = bid / nchunk;
    int c = bid - grp * nchunk;
    q.b_lo = grp * tm.tsh;
    q.nsh = min(tm.tsh, B - q.b_lo);
    q.plane = grp;
    const int nsc = tma_side_chunks(tm);
    int tr;
    if (c < 2 * nsc) {
        const int side = c / nsc;
        tr = c - side * nsc;
        tma_side_tile(tm, g, habc, nfx, tr, side, q);
    } else {
        c -= 2 * nsc;
        const int nrc = tma_row_chunks(tm);
        tr = c / nrc;
        const int i = c - tr * nrc;
        tr = tr < tm.nbot ? tm.ntr - tm.nbot + tr : tr - tm.nbot;
        const int c0 = tm.tx0 + i * ntx / nrc;              // balanced split of the tile row into nrc chunks
        q.z0 = tr * TR;
        q.x0 = c0 * FW;
        q.ntile = tm.tx0 + (i + 1) * ntx / nrc - c0;
        q.dz = 0; q.dx = FW;
        q.kind = tma_row_kind(tm, g, habc, q.z0);
    }
    // tiles of the acquisition rows belong to the per-shot blocks above
    if (tm.ar1 > tm.ar0 && tr >= tm.ar0 && tr < tm.ar1) q.ntile = 0;
    return q;
}
template <int FL>
__device__ __forceinline__ W2Coef load_coef_fl(const W2Args& a, long long idx) {
    W2Coef c;
    c.r = __ldg(a.coef[0] + idx);
    c.b = __ldg(a.coef[1] + idx);
    c.cxx = c.czz = c.cxz = c.ax = c.az = c.m = 0.f;
    c.cxx = __ldg(a.coef[2] + idx);
    if (!(FL & ST_F_ISO) || (FL & ST_F_PML)) c.czz = __ldg(a.coef[3] + idx);
    if (FL & ST_F_XZ) c.cxz = __ldg(a.coef[4] + idx);
    if (FL & ST_F_G1) { c.ax = __ldg(a.coef[5] + idx); c.az = __ldg(a.coef[6] + idx); }
    if (FL & ST_F_BORN) c.m = __ldg(a.coef[7] + idx);
    return c;
}

__device__ __forceinline__ void load_tile(float (*s)[SW], const float* __restrict__ src,
                                          int z0, int x0, const W2Geom& g, int tid) {
    for (int i = tid; i < SH * SW; i += NT) {
        const int lz = i / SW, lx = i - lz * SW;
        const int z = z0 - HALO + lz, x = x0 - HALO + lx;
        float v = 0.f;
        if (z >= 0 && z < g.nz && x >= 0 && x < g.nx) v = __ldg(src + (long long)z * g.ld + x);
        s[lz][lx] = v;
    }
}

// distance to the nearest absorbing edge (the free surface of `multiple` is not one)
__device__ __forceinline__ int edge_depth(int z, int x, const W2Geom& g) {
    int d = min(min(x, g.nx - 1 - x), g.nz - 1 - z);
    if (!g.multiple) d = min(d, z);
    return d;
}

// ---- enumeration of the TX x TZ tiles that touch the band of `band` cells along the
//      absorbing edges (band = bw for the forward frame, bw+1 for the adjoint)
struct BandTiles {
    int nxt, nzt, rows_top, rows_bot, cols_l, cols_r, count;
};
__host__ __device__ inline BandTiles band_tiles(const W2Geom& g, int band) {
    BandTiles t;
    t.nxt = (g.nx + TX - 1) / TX;
    t.nzt = (g.nz + TZ - 1) / TZ;
    const int n_top = g.multiple ? 0 : (band + TZ - 1) / TZ;
    const int n_bot = t.nzt - max(g.nz - band, 0) / TZ;
    const int n_l = (band + TX - 1) / TX;
    const int n_r = t.nxt - max(g.nx - band, 0) / TX;
    t.rows_top = min(n_top, t.nzt);
    t.rows_bot = min(n_bot, t.nzt - t.rows_top);
    t.cols_l = min(n_l, t.nxt);
    t.cols_r = min(n_r, t.nxt - t.cols_l);
    const int mid = t.nzt - t.rows_top - t.rows_bot;
    t.count = (t.rows_top + t.rows_bot) * t.nxt + mid * (t.cols_l + t.cols_r);
    return t;
}
__device__ __forceinline__ void band_tile_decode(const BandTiles& t, int id, int& tz, int& tx) {
    const int nb = (t.rows_top + t.rows_bot) * t.nxt;
    if (id < nb) {
        const int r = id / t.nxt;
        tx = id - r * t.nxt;
        tz = r < t.rows_top ? r : t.nzt - t.rows_bot + (r - t.rows_top);
    } else {
        const int j = id - nb, w = t.cols_l + t.cols_r;
        const int r = j / w, c = j - r * w;
        tz = t.rows_top + r;
        tx = c < t.cols_l ? c : t.nxt - t.cols_r + (c - t.cols_l);
    }
}

// ---- small float4 helpers
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4get(const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4set(float4& v, int e, float s) { if (e == 0) v.x = s; else if (e == 1) v.y = s; else if (e == 2) v.z = s; else v.w = s; }

// one row of a field as a float4 per lane (zero outside the domain / pitch)
__device__ __forceinline__ float4 ldrow(const float* __restrict__ base, int z, int x, const W2Geom& g) {
    if (z < 0 || z >= g.nz || x >= g.ld) return f4zero();
    return __ldg(reinterpret_cast<const float4*>(base + (z * g.ld + x)));
}
// SAFE: the caller knows the row is inside the domain (tiles away from the domain edge): no bounds predicate
template <bool SAFE>
__device__ __forceinline__ float4 ldrow_s(const float* __restrict__ base, int z, int x, const W2Geom& g) {
    if (SAFE) return __ldg(reinterpret_cast<const float4*>(base + (z * g.ld + x)));
    return ldrow(base, z, x, g);
}
// left / right neighbours of the lane's 4 cells: shuffles + predicated halo loads at the warp edges
__device__ __forceinline__ void row_halo(const float4& c, const float* __restrict__ base, int z, int x0, int lane,
                                         const W2Geom& g, float& left, float& right) {
    left = __shfl_up_sync(0xffffffffu, c.w, 1);
    right = __shfl_down_sync(0xffffffffu, c.x, 1);
    const bool zin = z >= 0 && z < g.nz;
    if (lane == 0) left = (zin && x0 > 0) ? __ldg(base + (z * g.ld + x0 - 1)) : 0.f;
    if (lane == 31) right = (zin && x0 + FW < g.nx) ? __ldg(base + (z * g.ld + x0 + FW)) : 0.f;
}

// source add + receiver gather for the cells this block stored
template <int NF, class Own>
__device__ __forceinline__ void forward_tail(const W2Args& a, int b, int z0, int zn, int x0, int xn, int tid, Own owns) {
    if (zn <= a.row_lo || z0 > a.row_hi) return;          // no source / receiver in these rows (block-uniform)
    const W2Geom& g = a.g;
    const long long boff = (long long)b * a.fs;
    __syncthreads();
    for (int s = tid; s < a.ns; s += NT) {
        if (a.src_b[s] != b) continue;
        const int sz = a.src_z[s], sx = a.src_x[s];
        if (sz >= z0 && sz < zn && sx >= x0 && sx < xn && owns(sz, sx)) {
            const float v = a.amp[s];
#pragma unroll
            for (int f = 0; f < NF; ++f)
                if (a.src_fmask >> f & 1) atomicAdd(a.next + f * a.cs + boff + (long long)sz * g.ld + sx, v);
        }
    }
    if (!a.rec_out) return;
    // rows of this tile that hold receivers (usually none or one): found by one thread per
    // row, then served by the whole block
    __shared__ int s_cnt, s_rows[FH];
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const int nrow = min(zn, g.nz) - z0;
    if (tid < nrow) {
        const int row = b * g.nz + z0 + tid;
        if (a.row_start[row + 1] > a.row_start[row]) s_rows[atomicAdd(&s_cnt, 1)] = z0 + tid;
    }
    __syncthreads();
    const int cnt = s_cnt;
    for (int i = 0; i < cnt; ++i) {
        const int z = s_rows[i];
        const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
        for (int r = lo + tid; r < hi; r += NT) {
            const int rx = a.rec_x[r];
            if (rx >= x0 && rx < xn && owns(z, rx)) {
                const long long o = (long long)a.rec_orig[r] * a.nchan;
                for (int ch = 0; ch < a.nchan; ++ch)
                    a.rec_out[o + ch] = a.next[a.chan_f[ch] * a.cs + boff + (long long)z * g.ld + rx];
            }
        }
    }
}


// ------------------------------------------------------------------------------ band: straight strips
// The straight part of the top / bottom band (columns [52, nx-52), 70 % of the band cells at the
// BASELINE size) has only z-direction far taps, so it vectorises like the interior: a warp owns one
// band row x 128 columns, each lane 4 cells (128-bit loads of fields AND tap planes), x-neighbours
// by shuffle; every lane walks all shots of its group with the transposed taps in registers.
struct StripGeom { int xs, xe, ncol, toprows, nrows, bd; };
__host__ __device__ inline StripGeom strip_geom(const W2Geom& g, int bd) {
    StripGeom t;
    t.bd = bd;
    t.xs = (g.bw + 2 + 3) / 4 * 4;
    t.xe = (g.nx - g.bw - 2) / 4 * 4;
    if (t.xe < t.xs) t.xe = t.xs;
    t.ncol = (t.xe - t.xs + FW - 1) / FW;
    t.toprows = g.multiple ? 0 : bd;
    t.nrows = t.toprows + bd;
    return t;
}
__host__ __device__ inline int strip_blocks(const StripGeom& t) { return (t.nrows * t.ncol + NWARP - 1) / NWARP; }
__device__ __forceinline__ bool in_strip(const StripGeom& t, const W2Geom& g, int z, int x) {
    return x >= t.xs && x < t.xe && ((!g.multiple && z < t.bd) || z >= g.nz - t.bd);
}
// aligned float4 of plane/field row z at column x (zero outside the domain rows / past the grid)
__device__ __forceinline__ float4 strip_row(const float* __restrict__ p, int z, int x, const W2Geom& g) {
    if (z < 0 || z >= g.nz || x + 3 >= g.nx) return f4zero();
    return __ldg(reinterpret_cast<const float4*>(p + (z * g.ld + x)));
}
// values at x-1 / x+4 of the lane's float4 (shuffle + scalar loads at the warp edges)
__device__ __forceinline__ void strip_lr(const float4& c, const float* __restrict__ p, int z, int x0c, int lane,
                                         const W2Geom& g, float& left, float& right) {
    left = __shfl_up_sync(0xffffffffu, c.w, 1);
    right = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lane == 0) left = __ldg(p + (z * g.ld + x0c - 1));
    if (lane == 31) right = (x0c + FW < g.nx) ? __ldg(p + (z * g.ld + x0c + FW)) : 0.f;
}
__device__ __forceinline__ float4 f4fma(const float4& a, const float4& b, const float4& c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4sub(const float4& a, const float4& b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4add(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4mul(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
// shifted views of a row: cells x-1..x+2 and x+1..x+4
__device__ __forceinline__ float4 f4shl(const float4& c, float left) { return make_float4(left, c.x, c.y, c.z); }
__device__ __forceinline__ float4 f4shr(const float4& c, float right) { return make_float4(c.y, c.z, c.w, right); }

// FL only selects the extras: the four diagonal taps of the mixed derivative (XZ) and the second field + coupling
// pre m A[p1] of the Born pairs; the single-field cross is the round-1 code.
template <int FL>
__device__ __forceinline__ void forward_strip_block(const W2Args& a, int blk, int b_lo, int b_hi, int tid) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool XZ = (FL & ST_F_XZ) != 0;
    const W2Geom g = a.g;
    const StripGeom t = strip_geom(g, g.bw);
    const int warp = tid >> 5, lane = tid & 31;
    const int item = blk * NWARP + warp;
    if (item >= t.nrows * t.ncol) return;                       // whole warp
    const int row = item / t.ncol, chunk = item - row * t.ncol;
    const bool top = row < t.toprows;
    const int z = top ? row : g.nz - t.bd + (row - t.toprows);
    const int n = top ? 1 : -1;
    const int x0c = t.xs + chunk * FW, x = x0c + 4 * lane;
    const bool active = x < t.xe;
    const long long plane = (long long)g.nz * g.ld;
    const int on1 = top ? 2 : 1, on2 = top ? 6 : 5;
    // x-1 / x+4 neighbours of a row that may lie outside the domain in z
    auto lr = [&](const float4& c, const float* p, int zq, float& left, float& right) {
        left = __shfl_up_sync(0xffffffffu, c.w, 1);
        right = __shfl_down_sync(0xffffffffu, c.x, 1);
        const bool zin = zq >= 0 && zq < g.nz;
        if (lane == 0) left = zin ? __ldg(p + (zq * g.ld + x0c - 1)) : 0.f;
        if (lane == 31) right = (zin && x0c + FW < g.nx) ? __ldg(p + (zq * g.ld + x0c + FW)) : 0.f;
    };
    // forward taps of this cell (increment form, st_wave2d_band.cuh)
    const float4 Tm = strip_row(a.taps + 1 * plane, z, x, g), Tp = strip_row(a.taps + 2 * plane, z, x, g);
    const float4 Tl = strip_row(a.taps + 3 * plane, z, x, g), Tr = strip_row(a.taps + 4 * plane, z, x, g);
    const float4 Tf = strip_row(a.taps + on2 * plane, z, x, g);
    const float4 T2 = strip_row(a.taps + (ST_NTAP1 + on1) * plane, z, x, g);
    float4 Tnw = f4zero(), Tne = f4zero(), Tsw = f4zero(), Tse = f4zero();
    if (XZ) {
        Tnw = strip_row(a.taps + 9 * plane, z, x, g); Tne = strip_row(a.taps + 10 * plane, z, x, g);
        Tsw = strip_row(a.taps + 11 * plane, z, x, g); Tse = strip_row(a.taps + 12 * plane, z, x, g);
    }
    // Born coupling  pre m A[p1]:  kz (N + S - 2C) + kx (W + E - 2C) + kxz ((SE - SW) - (NE - NW))  (every strip row is a frame row)
    float4 kx = f4zero(), kz = f4zero(), kxz = f4zero();
    if (NF == 2) {
        const float4 bb = strip_row(a.coef[1], z, x, g), mm = strip_row(a.coef[7], z, x, g);
        const float4 pm = make_float4((1.f - bb.x) * mm.x, (1.f - bb.y) * mm.y, (1.f - bb.z) * mm.z, (1.f - bb.w) * mm.w);
        kx = f4mul(pm, strip_row(a.coef[2], z, x, g));
        kz = f4mul(pm, strip_row(a.coef[3], z, x, g));
        if (XZ) kxz = f4mul(pm, strip_row(a.coef[4], z, x, g));
    }
    for (int b = b_lo; b < b_hi; ++b) {
        float4 cpl = f4zero();
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const long long boff = f * a.cs + (long long)b * a.fs;
            const float* cur = a.cur + boff;
            const float* prv = a.prev + boff;
            const float4 C = strip_row(cur, z, x, g), Um = strip_row(cur, z - 1, x, g), Up = strip_row(cur, z + 1, x, g);
            const float4 Uf = strip_row(cur, z + 2 * n, x, g);
            const float4 P = strip_row(prv, z, x, g), Pn = strip_row(prv, z + n, x, g);
            float l, r;
            strip_lr(C, cur, z, x0c, lane, g, l, r);
            const float4 dN = f4sub(Um, C), dS = f4sub(Up, C), dW = f4sub(f4shl(C, l), C), dE = f4sub(f4shr(C, r), C);
            float4 acc = f4mul(Tm, dN);
            acc = f4fma(Tp, dS, acc);
            acc = f4fma(Tl, dW, acc);
            acc = f4fma(Tr, dE, acc);
            acc = f4fma(Tf, f4sub(Uf, C), acc);
            acc = f4fma(T2, f4sub(Pn, P), acc);
            float4 dNW = f4zero(), dNE = f4zero(), dSW = f4zero(), dSE = f4zero();
            if (XZ) {
                float ml, mr, pl, pr;
                lr(Um, cur, z - 1, ml, mr);
                lr(Up, cur, z + 1, pl, pr);
                dNW = f4sub(f4shl(Um, ml), C); dNE = f4sub(f4shr(Um, mr), C);
                dSW = f4sub(f4shl(Up, pl), C); dSE = f4sub(f4shr(Up, pr), C);
                acc = f4fma(Tnw, dNW, acc);
                acc = f4fma(Tne, dNE, acc);
                acc = f4fma(Tsw, dSW, acc);
                acc = f4fma(Tse, dSE, acc);
            }
            if (NF == 2 && f == 0) {
                cpl = f4mul(kz, dN);
                cpl = f4fma(kz, dS, cpl);
                cpl = f4fma(kx, dW, cpl);
                cpl = f4fma(kx, dE, cpl);
                if (XZ) cpl = f4fma(kxz, f4sub(f4add(dNW, dSE), f4add(dNE, dSW)), cpl);
            }
            if (NF == 2 && f == 1) acc = f4add(acc, cpl);
            if (active) *reinterpret_cast<float4*>(a.next + boff + (z * g.ld + x)) = f4add(C, f4add(f4sub(C, P), acc));
        }
    }
    // sources / receivers inside this warp's cells (ordering only needs the warp's own stores)
    if (z < a.row_lo || z > a.row_hi) return;
    __syncwarp();
    const int xhi = min(x0c + FW, t.xe);
    for (int s = lane; s < a.ns; s += 32) {
        const int sb = a.src_b[s], sx = a.src_x[s];
        if (a.src_z[s] == z && sx >= x0c && sx < xhi && sb >= b_lo && sb < b_hi) {
#pragma unroll
            for (int f = 0; f < NF; ++f)
                if (a.src_fmask >> f & 1) atomicAdd(a.next + f * a.cs + (long long)sb * a.fs + (z * g.ld + sx), a.amp[s]);
        }
    }
    if (!a.rec_out) return;
    __syncwarp();
    for (int b = b_lo; b < b_hi; ++b) {
        const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
        for (int r = lo + lane; r < hi; r += 32) {
            const int rx = a.rec_x[r];
            if (rx >= x0c && rx < xhi) {
                const long long o = (long long)a.rec_orig[r] * a.nchan;
                for (int ch = 0; ch < a.nchan; ++ch)
                    a.rec_out[o + ch] = a.next[a.chan_f[ch] * a.cs + (long long)b * a.fs + (z * g.ld + rx)];
            }
        }
    }
}

// Adjoint of the straight strips.  Besides the single-field cross (round 1) it serves the mixed derivative of tti_habc
// (the four diagonal transposed taps) and the Born pairs without mixed derivative (both fields through the same taps; the
// scattered cotangent reaches the background one through  K = pre m T,  T = czz / cxx of the z / x neighbours, held as
// rows in registers like the taps).  The TTI Born pair would need ~120 registers of row constants and stays on the band threads.
template <int FL>
__device__ __forceinline__ void adjoint_strip_block(const W2Args& a, int blk, int b_lo, int b_hi, int gplane, int tid) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool XZ = (FL & ST_F_XZ) != 0;
    static_assert(!(XZ && NF == 2), "strip blocks: no TTI Born pair");
    const W2Geom g = a.g;
    const StripGeom t = strip_geom(g, g.bw + 1);
    const int warp = tid >> 5, lane = tid & 31;
    const int item = blk * NWARP + warp;
    if (item >= t.nrows * t.ncol) return;
    const int row = item / t.ncol, chunk = item - row * t.ncol;
    const bool top = row < t.toprows;
    const int z = top ? row : g.nz - t.bd + (row - t.toprows);
    const int n = top ? 1 : -1;
    const int x0c = t.xs + chunk * FW, x = x0c + 4 * lane;
    const bool active = x < t.xe;
    const long long plane = (long long)g.nz * g.ld;
    const bool want_grad = a.gacc != nullptr;
    auto frame_row = [&](int zq) { return (!g.multiple && zq < g.bw) || zq >= g.nz - g.bw; };
    const bool frame = frame_row(z);                            // the deepest band row is not a frame row
    const float* F1 = a.taps;
    const float* F2 = a.taps + ST_NTAP1 * plane;
    // x-1 / x+4 neighbours of a row that may lie outside the domain in z
    auto lr = [&](const float4& c, const float* p, int zq, float& left, float& right) {
        left = __shfl_up_sync(0xffffffffu, c.w, 1);
        right = __shfl_down_sync(0xffffffffu, c.x, 1);
        const bool zin = zq >= 0 && zq < g.nz;
        if (lane == 0) left = zin ? __ldg(p + (zq * g.ld + x0c - 1)) : 0.f;
        if (lane == 31) right = (zin && x0c + FW < g.nx) ? __ldg(p + (zq * g.ld + x0c + FW)) : 0.f;
    };
    // transposed taps: coefficient of L(p+o) is tap -o of cell p+o
    const float4 G0 = strip_row(F1, z, x, g);
    const float4 Gm = strip_row(F1 + 2 * plane, z - 1, x, g), Gp = strip_row(F1 + 1 * plane, z + 1, x, g);
    const float4 Gmm = strip_row(F1 + 6 * plane, z - 2, x, g), Gpp = strip_row(F1 + 5 * plane, z + 2, x, g);
    const float4 A4 = strip_row(F1 + 4 * plane, z, x, g), B4 = strip_row(F1 + 3 * plane, z, x, g);
    float al, ar, bl, br;
    strip_lr(A4, F1 + 4 * plane, z, x0c, lane, g, al, ar);
    strip_lr(B4, F1 + 3 * plane, z, x0c, lane, g, bl, br);
    const float4 Gl = f4shl(A4, al);                  // F1[+x](z, x-1): the left neighbour reads us through its +x tap
    const float4 Gr = f4shr(B4, br);                  // F1[-x](z, x+1)
    const float4 K0 = strip_row(F2, z, x, g), Km = strip_row(F2 + 2 * plane, z - 1, x, g), Kp = strip_row(F2 + 1 * plane, z + 1, x, g);
    // mixed derivative: the diagonal neighbours read us through their opposite diagonal tap (st_tap_neg: 9 <-> 12, 10 <-> 11)
    float4 Gnw = f4zero(), Gne = f4zero(), Gsw = f4zero(), Gse = f4zero();
    if (XZ) {
        float l_, r_;
        const float4 t12 = strip_row(F1 + 12 * plane, z - 1, x, g);
        lr(t12, F1 + 12 * plane, z - 1, l_, r_); Gnw = f4shl(t12, l_);
        const float4 t11 = strip_row(F1 + 11 * plane, z - 1, x, g);
        lr(t11, F1 + 11 * plane, z - 1, l_, r_); Gne = f4shr(t11, r_);
        const float4 t10 = strip_row(F1 + 10 * plane, z + 1, x, g);
        lr(t10, F1 + 10 * plane, z + 1, l_, r_); Gsw = f4shl(t10, l_);
        const float4 t9 = strip_row(F1 + 9 * plane, z + 1, x, g);
        lr(t9, F1 + 9 * plane, z + 1, l_, r_); Gse = f4shr(t9, r_);
    }
    float4 pre = make_float4(1.f, 1.f, 1.f, 1.f);
    if (frame) { const float4 bb = strip_row(a.coef[1], z, x, g); pre = make_float4(1.f - bb.x, 1.f - bb.y, 1.f - bb.z, 1.f - bb.w); }
    // Born coupling rows  K[-o](p+o) = (pre m T)(p+o)  and  K[0](p) = -(pre m)(p) (2 cxx + 2 czz)(p)
    float4 Czm = f4zero(), Czp = f4zero(), Cxl = f4zero(), Cxr = f4zero(), C0 = f4zero(), pm = f4zero(), cxr = f4zero(), czr = f4zero();
    if (NF == 2) {
        auto pm_row = [&](int zq) {                   // (pre m)(zq, x..x+3)
            float4 v = strip_row(a.coef[7], zq, x, g);
            if (frame_row(zq)) {
                const float4 bb = strip_row(a.coef[1], zq, x, g);
                v = make_float4(v.x * (1.f - bb.x), v.y * (1.f - bb.y), v.z * (1.f - bb.z), v.w * (1.f - bb.w));
            }
            return v;
        };
        pm = pm_row(z);
        cxr = strip_row(a.coef[2], z, x, g); czr = strip_row(a.coef[3], z, x, g);
        Czm = f4mul(pm_row(z - 1), strip_row(a.coef[3], z - 1, x, g));
        Czp = f4mul(pm_row(z + 1), strip_row(a.coef[3], z + 1, x, g));
        const float4 px = f4mul(pm, cxr);
        float pl_ = __shfl_up_sync(0xffffffffu, px.w, 1), pr_ = __shfl_down_sync(0xffffffffu, px.x, 1);
        if (lane == 0 || lane == 31) {
            const int xq = lane == 0 ? x0c - 1 : x0c + FW;
            float v = 0.f;
            if (xq < g.nx) {
                const int o = z * g.ld + xq;
                v = __ldg(a.coef[7] + o) * __ldg(a.coef[2] + o) * (frame ? 1.f - __ldg(a.coef[1] + o) : 1.f);
            }
            if (lane == 0) pl_ = v; else pr_ = v;
        }
        Cxl = f4shl(px, pl_); Cxr = f4shr(px, pr_);
        C0 = make_float4(-pm.x * (2.f * cxr.x + 2.f * czr.x), -pm.y * (2.f * cxr.y + 2.f * czr.y),
                         -pm.z * (2.f * cxr.z + 2.f * czr.z), -pm.w * (2.f * cxr.w + 2.f * czr.w));
    }
    const int on1 = top ? 2 : 1, on2 = top ? 6 : 5;
    const float* H1 = a.taps + (ST_NTAP1 + ST_NTAP2) * plane;
    const float* H2 = a.taps + (2 * ST_NTAP1 + ST_NTAP2) * plane;
    float4 gc = f4zero(), gr = f4zero(), gz = f4zero(), gax = f4zero(), gaz = f4zero(), gxz = f4zero(), gm = f4zero();
    for (int b = b_lo; b < b_hi; ++b) {
        float4 Lc[NF], sxx0 = f4zero(), szz0 = f4zero();
#pragma unroll
        for (int f = NF - 1; f >= 0; --f) {             // the scattered field first: its cotangent rows feed the background one
            const long long boff = f * a.cs + (long long)b * a.fs;
            const float* l1 = a.lam1 + boff;
            const float* l2 = a.lam2 + boff;
            const float4 L0 = strip_row(l1, z, x, g), Lm = strip_row(l1, z - 1, x, g), Lp = strip_row(l1, z + 1, x, g);
            Lc[f] = L0;
            float ll, lr_;
            strip_lr(L0, l1, z, x0c, lane, g, ll, lr_);
            float4 acc = f4mul(G0, L0);
            acc = f4fma(Gm, Lm, acc);
            acc = f4fma(Gp, Lp, acc);
            acc = f4fma(Gmm, strip_row(l1, z - 2, x, g), acc);
            acc = f4fma(Gpp, strip_row(l1, z + 2, x, g), acc);
            acc = f4fma(Gl, f4shl(L0, ll), acc);
            acc = f4fma(Gr, f4shr(L0, lr_), acc);
            acc = f4fma(K0, strip_row(l2, z, x, g), acc);
            acc = f4fma(Km, strip_row(l2, z - 1, x, g), acc);
            acc = f4fma(Kp, strip_row(l2, z + 1, x, g), acc);
            if (XZ) {
                float ml, mr, pl2, pr2;
                lr(Lm, l1, z - 1, ml, mr);
                lr(Lp, l1, z + 1, pl2, pr2);
                acc = f4fma(Gnw, f4shl(Lm, ml), acc);
                acc = f4fma(Gne, f4shr(Lm, mr), acc);
                acc = f4fma(Gsw, f4shl(Lp, pl2), acc);
                acc = f4fma(Gse, f4shr(Lp, pr2), acc);
            }
            if (NF == 2 && f == 1) {
                // coupling into the background cotangent (stored with it below): held in Cacc until f == 0
                float4 c = f4mul(C0, L0);
                c = f4fma(Czm, Lm, c);
                c = f4fma(Czp, Lp, c);
                c = f4fma(Cxl, f4shl(L0, ll), c);
                c = f4fma(Cxr, f4shr(L0, lr_), c);
                sxx0 = c;                               // (register reuse: sxx0 is overwritten when the gradients are formed)
            }
            if (NF == 2 && f == 0) acc = f4add(acc, sxx0);
            if (active) *reinterpret_cast<float4*>(a.lam0 + boff + (z * g.ld + x)) = acc;
        }
        if (want_grad) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const long long boff = f * a.cs + (long long)b * a.fs;
                const float* S1 = a.s1 + boff;
                const float4 s0 = strip_row(S1, z, x, g), sm = strip_row(S1, z - 1, x, g), sp = strip_row(S1, z + 1, x, g);
                float sl, sr;
                strip_lr(s0, S1, z, x0c, lane, g, sl, sr);
                const float4 sw4 = f4shl(s0, sl), se4 = f4shr(s0, sr);
                const float4 szz = f4add(f4sub(sm, s0), f4sub(sp, s0)), sxx = f4add(f4sub(sw4, s0), f4sub(se4, s0));
                float4 pl = f4mul(pre, Lc[f]);                                   // effective cotangent of this field
                if (NF == 2 && f == 0) { pl = f4fma(pm, Lc[NF - 1], pl); sxx0 = sxx; szz0 = szz; }
                if (FL & ST_F_ISO) gc = f4fma(pl, f4add(szz, sxx), gc);
                else { gc = f4fma(pl, sxx, gc); gz = f4fma(pl, szz, gz); }
                if (FL & ST_F_G1) { gax = f4fma(pl, f4sub(se4, sw4), gax); gaz = f4fma(pl, f4sub(sp, sm), gaz); }
                if (XZ) {                                                        // (SE - SW) - (NE - NW)
                    float ml, mr, pl2, pr2;
                    lr(sm, S1, z - 1, ml, mr);
                    lr(sp, S1, z + 1, pl2, pr2);
                    const float4 cross = f4sub(f4sub(f4shr(sp, pr2), f4shl(sp, pl2)), f4sub(f4shr(sm, mr), f4shl(sm, ml)));
                    gxz = f4fma(pl, cross, gxz);
                }
                if (NF == 2 && f == 1)                                           // g_m += (pre L1_s) A[S_background]
                    gm = f4fma(f4mul(pre, Lc[NF - 1]), f4fma(cxr, sxx0, f4mul(czr, szz0)), gm);
                if (frame) {
                    const float* S2 = a.s2 + boff;
                    const float4 sn = top ? sp : sm;
                    float4 tt = f4mul(strip_row(H1, z, x, g), s0);
                    tt = f4fma(strip_row(H1 + on1 * plane, z, x, g), sn, tt);
                    tt = f4fma(strip_row(H1 + on2 * plane, z, x, g), strip_row(S1, z + 2 * n, x, g), tt);
                    tt = f4fma(strip_row(H2, z, x, g), strip_row(S2, z, x, g), tt);
                    tt = f4fma(strip_row(H2 + on1 * plane, z, x, g), strip_row(S2, z + n, x, g), tt);
                    gr = f4fma(Lc[f], tt, gr);
                }
            }
        }
    }
    if (want_grad && active) {
        float* gb = a.gacc + (long long)gplane * 7 * plane + (z * g.ld + x);
        auto rmw = [&](int slot, const float4& v) {
            float4* p4 = reinterpret_cast<float4*>(gb + slot * plane);
            *p4 = f4add(*p4, v);
        };
        rmw(1, gc);                                       // slot 1: d/d ciso (ISO) or d/d cxx
        if (!(FL & ST_F_ISO)) rmw(2, gz);                 // slot 2: d/d czz
        if (XZ) rmw(3, gxz);                              // slot 3: d/d cxz
        if (FL & ST_F_G1) { rmw(4, gax); rmw(5, gaz); }
        if (NF == 2) rmw(6, gm);                          // slot 6: d/d m
        if (frame) rmw(0, gr);                            // slot 0: d/d r
    }
    if (z < a.row_lo || z > a.row_hi) return;
    __syncwarp();
    const int xhi = min(x0c + FW, t.xe);
    if (a.rec_adj) {
        for (int b = b_lo; b < b_hi; ++b) {
            const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
            for (int r = lo + lane; r < hi; r += 32) {
                const int rx = a.rec_x[r];
                if (rx >= x0c && rx < xhi) {
                    const long long o = (long long)a.rec_orig[r] * a.nchan;
                    for (int ch = 0; ch < a.nchan; ++ch)
                        atomicAdd(a.lam0 + a.chan_f[ch] * a.cs + (long long)b * a.fs + (z * g.ld + rx), a.rec_adj[o + ch]);
                }
            }
        }
    }
    if (a.gamp) {
        __syncwarp();
        for (int s = lane; s < a.ns; s += 32) {
            const int sb = a.src_b[s], sx = a.src_x[s];
            if (a.src_z[s] == z && sx >= x0c && sx < xhi && sb >= b_lo && sb < b_hi) {
                float v = 0.f;
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    if (a.src_fmask >> f & 1) v += a.lam0[f * a.cs + (long long)sb * a.fs + (z * g.ld + sx)];
                if (NF == 2 || (a.src_fmask & 1)) a.gamp[s] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------ band (tap gather)
// rows touched by the 256 consecutive band cells of a block -> shared list (<= 16 rows)
// ---- cell maps of the tap-gather blocks: which cell a thread owns, which cells the block owns
#ifndef ST_TAP_INFLIGHT
#define ST_TAP_INFLIGHT 2                  // shots a band thread of the adjoint has in flight
#endif
#ifndef ST_BAND_DENSE
#define ST_BAND_DENSE 0
#endif
#ifndef ST_BAND_DENSE_BORN
#define ST_BAND_DENSE_BORN 1
#endif
// BandMap: the compact enumeration of the absorbing band (st_band_cells), minus the vectorised strips, NT cells per block
struct BandMap {
    static constexpr bool dense = ST_BAND_DENSE != 0;    // false: zero taps (most band cells use one side only) are skipped
    W2Geom g; BandCells bc; StripGeom sg; bool strips, frame_only; int i0;
    __device__ __forceinline__ bool cell(int t, int& z, int& x) const {
        const int i = i0 + t;
        if (t < 0 || i >= bc.total) return false;
        st_band_decode(bc, i, z, x);
        return !(strips && in_strip(sg, g, z, x));
    }
    __device__ __forceinline__ bool mine(int z, int x) const {
        if (z < 0 || z >= g.nz || x < 0 || x >= g.nx || (frame_only && !w2_in_frame(z, x, g)) || (strips && in_strip(sg, g, z, x))) return false;
        const int e = st_band_encode(bc, z, x);
        return e >= i0 && e < i0 + NT;
    }
    // true when none of the block's cells lies in the acquisition row range (block-uniform)
    __device__ __forceinline__ bool outside_rows(const W2Args& a) const {
        int zf, xf, zl, xl;
        st_band_decode(bc, i0, zf, xf);
        st_band_decode(bc, min(i0 + NT, bc.total) - 1, zl, xl);
        const int lo = min(zf, zl), hi = max(zf, zl);
        // rows of a block are contiguous inside one rectangle; straddling blocks are never skipped
        const bool same = (i0 + NT <= bc.n_top) || (i0 >= bc.n_top && i0 + NT <= bc.n_top + bc.n_bot) || (i0 >= bc.n_top + bc.n_bot);
        return same && (hi < a.row_lo || lo > a.row_hi);
    }
};
// RectMap: NT/w rows x w columns of a rectangle (the corner tiles of the TMA launches: every cell, frame or not)
struct RectMap {
    // corner cells carry taps of two sides: every tap is loaded unconditionally, so the loads of one shot are issued
    // together instead of as a chain of branch-guarded load -> use round trips (measured on the TMA forward: the chained
    // version made the corner blocks the critical path of the launch)
    static constexpr bool dense = true;
    W2Geom g; int z0, x0, w;            // first row / column of the block's cells, columns per row
    __device__ __forceinline__ bool cell(int t, int& z, int& x) const {
        if (t < 0) return false;
        z = z0 + t / w; x = x0 + t % w;
        return z < g.nz && x < g.nx;
    }
    __device__ __forceinline__ bool mine(int z, int x) const {
        return z >= z0 && z < z0 + NT / w && z < g.nz && x >= x0 && x < x0 + w && x < g.nx;
    }
    __device__ __forceinline__ bool outside_rows(const W2Args& a) const { return z0 + NT / w - 1 < a.row_lo || z0 > a.row_hi; }
};

// distinct rows of the block's cells (at most 16 per block), for the receiver epilogue
template <class Map>
__device__ __forceinline__ int tap_block_rows(const Map& map, int tid, int* s_rows, int* s_cnt) {
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
    int z, x, zp = -1, xp;
    if (map.cell(tid, z, x)) {
        if (!map.cell(tid - 1, zp, xp)) zp = -1;
        if (z != zp) {
            const int k = atomicAdd(s_cnt, 1);
            if (k < 16) s_rows[k] = z;
        }
    }
    __syncthreads();
    return min(*s_cnt, 16);
}

// Every thread owns ONE cell and walks all shots with its taps held in registers, so the
// tap planes are read once per step, not once per shot.
template <int FL, class Map>
__device__ __forceinline__ void forward_tap_block(const W2Args& a, const Map& map, int b_lo, int b_hi, int tid) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool XZ = (FL & ST_F_XZ) != 0;
    constexpr int NT1 = XZ ? ST_NTAP1 : ST_NTAP1C;           // taps of h1 this equation uses (13 with the mixed derivative)
    constexpr int NK = XZ ? ST_NTAP1 : ST_NTAP2;             // taps of the Born coupling stencil (cross, + diagonals for XZ)
    const W2Geom g = a.g;
    const long long plane = (long long)g.nz * g.ld;
    auto mine = [&](int z, int x) { return map.mine(z, x); };
    int zc = -1, xc = -1;
    if (map.cell(tid, zc, xc)) {
        const int z = zc, x = xc;
        const int idx = z * g.ld + x;
        // increment form: Y = h1 + (h1 - h2) + sum_{o != 0} F1[o] (h1(p+o) - h1(p)) + F2[o] (h2(p+o) - h2(p))
        // (the taps of h1 sum to 2 and those of h2 to -1 exactly, DESIGN.md "numerics")
        // taps that fall outside the domain read zero: t (0 - c) is folded into a self coefficient so
        // every load below is unconditional (index clamped to the cell itself)
        float t1[NT1 - 1], t2[ST_NTAP2 - 1], t1self = 0.f, t2self = 0.f;
        int q1[NT1 - 1];
#pragma unroll
        for (int o = 1; o < NT1; ++o) {
            const int zz = z + st_tap_dz(o), xx = x + st_tap_dx(o);
            const bool in = zz >= 0 && zz < g.nz && xx >= 0 && xx < g.nx;
            q1[o - 1] = in ? zz * g.ld + xx : idx;
            const float t = __ldg(a.taps + o * plane + idx);
            t1[o - 1] = in ? t : 0.f;
            t1self -= in ? 0.f : t;
            if (o < ST_NTAP2) {
                const float u = __ldg(a.taps + (ST_NTAP1 + o) * plane + idx);
                t2[o - 1] = in ? u : 0.f;
                t2self -= in ? 0.f : u;
            }
        }
        // Born pair: coupling  pre m A[p1]  into the scattered field, A = czz (N + S - 2C) + cxx (W + E - 2C) with zero
        // padding (w2_forward_cell); kc[o-1] multiplies (p1(q+o) - p1(q)), kself the centre value for absent neighbours
        float kc[NK - 1], kself = 0.f;
#pragma unroll
        for (int o = 1; o < NK; ++o) kc[o - 1] = 0.f;
        if (NF == 2) {
            const float pm = (1.f - __ldg(a.coef[1] + idx)) * __ldg(a.coef[7] + idx);
            const float cx = __ldg(a.coef[2] + idx), cz = __ldg(a.coef[3] + idx);
            const float cxz = XZ ? __ldg(a.coef[4] + idx) : 0.f;
#pragma unroll
            for (int o = 1; o < NK; ++o) {
                if (o >= ST_NTAP2 && o < ST_NTAP1C) continue;            // the far taps are not part of the stencil
                const int zz = z + st_tap_dz(o), xx = x + st_tap_dx(o);
                const bool in = zz >= 0 && zz < g.nz && xx >= 0 && xx < g.nx;
                const float t = pm * (o <= 2 ? cz : o <= 4 ? cx : cxz * st_tap_xz_sign(o));
                kc[o - 1] = in ? t : 0.f;
                kself -= in ? 0.f : t;
            }
        }
        for (int b = b_lo; b < b_hi; b += 2) {
            // two shots per iteration: two independent load chains in flight
            const bool two = b + 1 < b_hi;
            const long long boff0 = (long long)b * a.fs, boff1 = two ? boff0 + a.fs : boff0;
            float cpl0 = 0.f, cpl1 = 0.f;                       // coupling term (computed with field 0, used by field 1)
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const float* cur0 = a.cur + f * a.cs + boff0;
                const float* prv0 = a.prev + f * a.cs + boff0;
                const float* cur1 = a.cur + f * a.cs + boff1;
                const float* prv1 = a.prev + f * a.cs + boff1;
                const float c0 = __ldg(cur0 + idx), p0 = __ldg(prv0 + idx);
                const float c1 = __ldg(cur1 + idx), p1 = __ldg(prv1 + idx);
                float acc0 = t1self * c0 + t2self * p0, acc1 = t1self * c1 + t2self * p1;
                if (NF == 2 && f == 0) { cpl0 = kself * c0; cpl1 = kself * c1; }
#pragma unroll
                for (int o = 1; o < NT1; ++o) {
                    // most cells use one side only: 3 of the 4 far taps and 3 of the 4 h2 taps are zero
                    const bool kcpl = NF == 2 && f == 0 && o < NK && kc[o < NK ? o - 1 : 0] != 0.f;
                    const bool need = Map::dense || t1[o - 1] != 0.f || kcpl;
                    if (need) {
                        const float d0 = __ldg(cur0 + q1[o - 1]) - c0, d1 = __ldg(cur1 + q1[o - 1]) - c1;
                        acc0 += t1[o - 1] * d0;
                        acc1 += t1[o - 1] * d1;
                        if (NF == 2 && f == 0 && o < NK) { cpl0 += kc[o < NK ? o - 1 : 0] * d0; cpl1 += kc[o < NK ? o - 1 : 0] * d1; }
                    }
                    if (o < ST_NTAP2 && (Map::dense || t2[o < ST_NTAP2 ? o - 1 : 0] != 0.f)) {
                        acc0 += t2[o - 1] * (__ldg(prv0 + q1[o - 1]) - p0);
                        acc1 += t2[o - 1] * (__ldg(prv1 + q1[o - 1]) - p1);
                    }
                }
                if (NF == 2 && f == 1) { acc0 += cpl0; acc1 += cpl1; }
                a.next[f * a.cs + boff0 + idx] = c0 + ((c0 - p0) + acc0);
                if (two) a.next[f * a.cs + boff1 + idx] = c1 + ((c1 - p1) + acc1);
            }
        }
    }
    __syncthreads();
    // wrap-around neighbour of the strip (at most 4 cells of the whole grid)
    if (tid < 4) {
        int z, x, s, zw, xw;
        st_wrap_candidate(g, tid, z, x, s, zw, xw);
        if (mine(z, x)) {
            float f[4];
            w2_side_weights(z, x, g, f);
            const int idx = z * g.ld + x;
            const float r = __ldg(a.coef[0] + idx), w = __ldg(a.coef[1] + idx) * f[s];
            // true operator: -mu*w*h1(wrap); the increment form above assumed taps summing to 2, i.e. it
            // implicitly carries -mu*w*h1(q) for the missing tap: replace it
            if (w != 0.f)
                for (int b = b_lo; b < b_hi; ++b) {
#pragma unroll
                    for (int fl = 0; fl < NF; ++fl) {
                        const long long boff = fl * a.cs + (long long)b * a.fs;
                        a.next[boff + idx] += w * (-r * r) * (__ldg(a.cur + boff + (zw * g.ld + xw)) - __ldg(a.cur + boff + idx));
                    }
                }
        }
    }
    if (map.outside_rows(a)) return;
    __syncthreads();
    for (int s = tid; s < a.ns; s += NT) {
        const int sz = a.src_z[s], sx = a.src_x[s], sb = a.src_b[s];
        if (sb >= b_lo && sb < b_hi && mine(sz, sx)) {
#pragma unroll
            for (int fl = 0; fl < NF; ++fl)
                if (a.src_fmask >> fl & 1) atomicAdd(a.next + fl * a.cs + (long long)sb * a.fs + (sz * g.ld + sx), a.amp[s]);
        }
    }
    if (!a.rec_out) return;
    __shared__ int s_cnt, s_rows[16];
    const int cnt = tap_block_rows(map, tid, s_rows, &s_cnt);
    for (int k = 0; k < cnt; ++k) {
        const int z = s_rows[k];
        for (int b = b_lo; b < b_hi; ++b) {
            const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
            for (int r = lo + tid; r < hi; r += NT) {
                const int rx = a.rec_x[r];
                if (mine(z, rx)) {
                    const long long o = (long long)a.rec_orig[r] * a.nchan;
                    for (int ch = 0; ch < a.nchan; ++ch)
                        a.rec_out[o + ch] = a.next[a.chan_f[ch] * a.cs + (long long)b * a.fs + (z * g.ld + rx)];
                }
            }
        }
    }
}

template <int FL>
__device__ __forceinline__ void forward_band_block(const W2Args& a, int blk, int b_lo, int b_hi, int tid) {
    const BandMap map{a.g, st_band_cells(a.g, a.g.bw), strip_geom(a.g, a.g.bw), st_flags_stripped_fwd(FL), true, blk * NT};
    forward_tap_block<FL>(a, map, b_lo, b_hi, tid);
}

template <int FL, class Map>
__device__ __forceinline__ void adjoint_tap_block(const W2Args& a, const Map& map, int b_lo, int b_hi, int gplane, int tid) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool XZ = (FL & ST_F_XZ) != 0;
    constexpr int NT1 = XZ ? ST_NTAP1 : ST_NTAP1C;           // taps of h1 this equation uses (13 with the mixed derivative)
    constexpr int NK = XZ ? ST_NTAP1 : ST_NTAP2;             // taps of the spatial stencil (cross, + diagonals for XZ)
    // zero taps: skipped by the single-field equations (saves loads); the Born pairs load every tap unconditionally -- their
    // guard needs the tap AND the coupling weight, and the guarded version issues 3x the instructions (B200: 268 -> 224 us)
    constexpr bool DENSE = Map::dense || (ST_BAND_DENSE_BORN != 0 && NF == 2);
    const W2Geom g = a.g;
    const long long plane = (long long)g.nz * g.ld;
    const bool want_grad = a.gacc != nullptr;
    auto inb = [&](int z, int x) { return z >= 0 && z < g.nz && x >= 0 && x < g.nx; };
    auto mine = [&](int z, int x) { return map.mine(z, x); };
    int zc = -1, xc = -1;
    if (map.cell(tid, zc, xc)) {
        const int z = zc, x = xc;
        const int idx = z * g.ld + x;
        // transposed taps: coefficient of L(p+o) is the tap -o of cell p+o
        float g1[NT1], g2[ST_NTAP2], h1[ST_NTAP1C], h2[ST_NTAP2], m[NT1];
        int q[NT1];
        const bool frame = w2_in_frame(z, x, g);
        const float pre = frame ? 1.f - __ldg(a.coef[1] + idx) : 1.f;
#pragma unroll
        for (int o = 0; o < NT1; ++o) {
            const int zz = z + st_tap_dz(o), xx = x + st_tap_dx(o);
            const bool in = inb(zz, xx);
            q[o] = in ? zz * g.ld + xx : idx;                 // clamped: every load below is unconditional
            g1[o] = in ? __ldg(a.taps + st_tap_neg(o) * plane + q[o]) : 0.f;
            m[o] = in ? 1.f : 0.f;
            if (o < ST_NTAP1C) h1[o] = (want_grad && frame && in) ? __ldg(a.taps + (ST_NTAP1 + ST_NTAP2 + o) * plane + idx) : 0.f;
            if (o < ST_NTAP2) {
                g2[o] = in ? __ldg(a.taps + (ST_NTAP1 + st_tap_neg(o)) * plane + q[o]) : 0.f;
                h2[o] = (want_grad && frame && in) ? __ldg(a.taps + (2 * ST_NTAP1 + ST_NTAP2 + o) * plane + idx) : 0.f;
            }
        }
        // Born pair: the scattered field's cotangent reaches the background field through the coupling
        //   sp' += pre m A[p1]  =>  Lam_p(p) += sum_o K[-o](p+o) L1_s(p+o),   K[o](q) = pre(q) m(q) T[o](q), K[0] = -sum_o K[o]
        // (T = czz for the z-neighbours, cxx for the x-neighbours, +-cxz for the diagonal ones; w2_adjoint_cell: leffq)
        float gk[NK];
#pragma unroll
        for (int o = 0; o < NK; ++o) gk[o] = 0.f;
        float pm_p = 0.f, cx_p = 0.f, cz_p = 0.f, cxz_p = 0.f;
        if (NF == 2 || XZ) {
            cx_p = __ldg(a.coef[2] + idx); cz_p = __ldg(a.coef[3] + idx);
            if (XZ) cxz_p = __ldg(a.coef[4] + idx);
        }
        if (NF == 2) {
            pm_p = pre * __ldg(a.coef[7] + idx);
            gk[0] = -pm_p * (2.f * cx_p + 2.f * cz_p);      // (the four diagonal taps sum to zero)
#pragma unroll
            for (int o = 1; o < NK; ++o) {
                if (o >= ST_NTAP2 && o < ST_NTAP1C) continue;            // far taps: not part of the stencil
                if (m[o] != 0.f) {
                    const int zz = z + st_tap_dz(o), xx = x + st_tap_dx(o);
                    const float preq = w2_in_frame(zz, xx, g) ? 1.f - __ldg(a.coef[1] + q[o]) : 1.f;
                    const float tq = o <= 2 ? __ldg(a.coef[3] + q[o]) : o <= 4 ? __ldg(a.coef[2] + q[o])
                                                                                 : st_tap_xz_sign(o) * __ldg(a.coef[4] + q[o]);
                    gk[o] = preq * __ldg(a.coef[7] + q[o]) * tq;
                }
            }
        }
        float gr = 0.f, gc = 0.f, gz = 0.f, gax = 0.f, gaz = 0.f, gm = 0.f, gxz = 0.f;
        // one shot: Lam_i(p) of every field and the gradient contributions; written as a lambda so two shots can be
        // issued back to back (two independent load chains in flight)
        auto one_shot = [&](int b, float (&accOut)[NF], float& gcOut, float& grOut, float& gzOut, float& gaxOut, float& gazOut,
                            float& gmOut, float& gxzOut) {
            const long long boff = (long long)b * a.fs;
            float lc[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) accOut[f] = 0.f;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const float* l1 = a.lam1 + f * a.cs + boff;
                const float* l2 = a.lam2 + f * a.cs + boff;
                float acc = 0.f;
                lc[f] = 0.f;
#pragma unroll
                for (int o = 0; o < NT1; ++o) {
                    // zero taps (most far taps of single-side cells) are skipped; the centre value is
                    // always needed for the imaging condition
                    const bool cpl = NF == 2 && f == 1 && o < NK && gk[o < NK ? o : 0] != 0.f;
                    if (o == 0 || DENSE || g1[o] != 0.f || cpl) {
                        const float v = __ldg(l1 + q[o]);
                        if (o == 0) lc[f] = v;
                        acc += g1[o] * v;
                        if (NF == 2 && f == 1 && o < NK) accOut[0] += gk[o < NK ? o : 0] * v;
                    }
                    if (o < ST_NTAP2 && (DENSE || g2[o < ST_NTAP2 ? o : 0] != 0.f)) acc += g2[o < ST_NTAP2 ? o : 0] * __ldg(l2 + q[o]);
                }
                accOut[f] += acc;
            }
            if (want_grad) {
                float A0 = 0.f;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const float* S1 = a.s1 + f * a.cs + boff;
                    const float* S2 = a.s2 + f * a.cs + boff;
                    float s[ST_NTAP2], t = 0.f;
#pragma unroll
                    for (int o = 0; o < ST_NTAP1C; ++o) {
                        if (o < ST_NTAP2) {
                            s[o] = m[o] * __ldg(S1 + q[o]);
                            t += h1[o] * s[o];
                            if (DENSE || h2[o] != 0.f) t += h2[o] * __ldg(S2 + q[o]);
                        } else if (DENSE || h1[o] != 0.f) {
                            t += h1[o] * __ldg(S1 + q[o]);
                        }
                    }
                    const float szz = (s[1] - s[0]) + (s[2] - s[0]), sxx = (s[3] - s[0]) + (s[4] - s[0]);
                    float cross = 0.f;
                    if (XZ) {       // (SE - SW) - (NE - NW), zero outside the domain
                        const float nw = m[XZ ? 9 : 0] * __ldg(S1 + q[XZ ? 9 : 0]), ne = m[XZ ? 10 : 0] * __ldg(S1 + q[XZ ? 10 : 0]);
                        const float sw = m[XZ ? 11 : 0] * __ldg(S1 + q[XZ ? 11 : 0]), se = m[XZ ? 12 : 0] * __ldg(S1 + q[XZ ? 12 : 0]);
                        cross = (se - sw) - (ne - nw);
                    }
                    float le = pre * lc[f];
                    if (NF == 2 && f == 0) le += pm_p * lc[1];
                    if (FL & ST_F_ISO) gcOut += le * (szz + sxx);
                    else { gcOut += le * sxx; gzOut += le * szz; }
                    if (XZ) gxzOut += le * cross;
                    if (FL & ST_F_G1) { gaxOut += le * (s[4] - s[3]); gazOut += le * (s[2] - s[1]); }
                    if (NF == 2) {
                        if (f == 0) A0 = cx_p * sxx + cz_p * szz + cxz_p * cross;
                        else gmOut += (pre * lc[1]) * A0;
                    }
                    grOut += lc[f] * t;
                }
            }
        };
        // NIF shots are issued back to back (independent load chains in flight); their gradient sums are kept apart and
        // added in shot order
        constexpr int NIF = ST_TAP_INFLIGHT;
        for (int b = b_lo; b < b_hi; b += NIF) {
            float accs[NIF][NF], gs[NIF][7];
#pragma unroll
            for (int i = 0; i < NIF; ++i) {
#pragma unroll
                for (int q = 0; q < 7; ++q) gs[i][q] = 0.f;
                if (i == 0 || b + i < b_hi) one_shot(b + i, accs[i], gs[i][0], gs[i][1], gs[i][2], gs[i][3], gs[i][4], gs[i][5], gs[i][6]);
            }
#pragma unroll
            for (int i = 0; i < NIF; ++i) {
                if (i == 0 || b + i < b_hi) {
#pragma unroll
                    for (int f = 0; f < NF; ++f) a.lam0[f * a.cs + (long long)(b + i) * a.fs + idx] = accs[i][f];
                }
                gc += gs[i][0]; gr += gs[i][1]; gz += gs[i][2]; gax += gs[i][3]; gaz += gs[i][4]; gm += gs[i][5]; gxz += gs[i][6];
            }
        }
        if (want_grad) {
            float* gb = a.gacc + (long long)gplane * 7 * plane;
            gb[plane + idx] += gc;              // slot 1: d/d ciso (ISO) or d/d cxx
            if (!(FL & ST_F_ISO)) gb[2 * plane + idx] += gz;                   // slot 2: d/d czz
            if (XZ) gb[3 * plane + idx] += gxz;                                 // slot 3: d/d cxz
            if (FL & ST_F_G1) { gb[4 * plane + idx] += gax; gb[5 * plane + idx] += gaz; }
            if (FL & ST_F_BORN) gb[6 * plane + idx] += gm;                      // slot 6: d/d m
            if (frame) gb[idx] += gr;           // slot 0: d/d r
        }
    }
    __syncthreads();
    if (tid < 4) {
        int z, x, s, zw, xw;
        st_wrap_candidate(g, tid, z, x, s, zw, xw);
        float f[4];
        w2_side_weights(z, x, g, f);
        const int qq = z * g.ld + x, qw = zw * g.ld + xw;
        const float r = __ldg(a.coef[0] + qq), w = __ldg(a.coef[1] + qq) * f[s];
        if (w != 0.f) {
            for (int b = b_lo; b < b_hi; ++b) {
#pragma unroll
                for (int fl = 0; fl < NF; ++fl) {
                    const long long boff = fl * a.cs + (long long)b * a.fs;
                    const float lq = __ldg(a.lam1 + boff + qq);
                    if (mine(zw, xw)) atomicAdd(a.lam0 + boff + qw, w * (-r * r) * lq);
                    if (want_grad && mine(z, x)) atomicAdd(a.gacc + (long long)gplane * 7 * plane + qq, lq * w * (-2.f * r) * __ldg(a.s1 + boff + qw));
                }
            }
        }
    }
    if (map.outside_rows(a)) return;
    __syncthreads();
    if (a.rec_adj) {
        __shared__ int s_cnt, s_rows[16];
        const int cnt = tap_block_rows(map, tid, s_rows, &s_cnt);
        for (int k = 0; k < cnt; ++k) {
            const int z = s_rows[k];
            for (int b = b_lo; b < b_hi; ++b) {
                const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
                for (int r = lo + tid; r < hi; r += NT) {
                    const int rx = a.rec_x[r];
                    if (mine(z, rx)) {
                        const long long o = (long long)a.rec_orig[r] * a.nchan;
                        for (int ch = 0; ch < a.nchan; ++ch)
                            atomicAdd(a.lam0 + a.chan_f[ch] * a.cs + (long long)b * a.fs + (z * g.ld + rx), a.rec_adj[o + ch]);
                    }
                }
            }
        }
    }
    if (a.gamp) {
        __syncthreads();
        for (int s = tid; s < a.ns; s += NT) {
            const int sz = a.src_z[s], sx = a.src_x[s], sb = a.src_b[s];
            if (sb >= b_lo && sb < b_hi && mine(sz, sx)) {
                float v = 0.f;
#pragma unroll
                for (int fl = 0; fl < NF; ++fl)
                    if (a.src_fmask >> fl & 1) v += a.lam0[fl * a.cs + (long long)sb * a.fs + (sz * g.ld + sx)];
                a.gamp[s] = v;
            }
        }
    }
}


template <int FL>
__device__ __forceinline__ void adjoint_band_block(const W2Args& a, int blk, int b_lo, int b_hi, int gplane, int tid) {
    const BandMap map{a.g, st_band_cells(a.g, a.g.bw + 1), strip_geom(a.g, a.g.bw + 1), st_flags_stripped_adj(FL), false, blk * NT};
    adjoint_tap_block<FL>(a, map, b_lo, b_hi, gplane, tid);
}

// ------------------------------------------------------------------------------ forward
// rows of one warp tile.  SAFE: every load of the tile (rows z0-1..z0+FRZ, columns x0-1..x0+FW)
// is inside the domain, so no bounds predicate is needed and pointers are simply advanced by
// the row pitch; otherwise the checked loaders are used (domain-edge tiles only).
template <int FL, bool SAFE>
__device__ __forceinline__ void forward_fast_rows(const W2Args& a, const W2Geom& g, int b, int x0, int z0, int zn,
                                                  int lane, bool clean) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool HABC = (FL & ST_F_HABC) != 0;
    const int x = x0 + 4 * lane;
    const int ld = g.ld;
    const long long boff = (long long)b * a.fs;
    const float* cur[NF];
    const float* prv[NF];
    float* nxt[NF];
    float4 U[NF], Cc[NF], D[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        cur[f] = a.cur + f * a.cs + boff;
        prv[f] = a.prev + f * a.cs + boff;
        nxt[f] = a.next + f * a.cs + boff;
        if (SAFE) {
            U[f] = __ldg(reinterpret_cast<const float4*>(cur[f] + ((z0 - 1) * ld + x)));
            Cc[f] = __ldg(reinterpret_cast<const float4*>(cur[f] + (z0 * ld + x)));
        } else {
            U[f] = ldrow(cur[f], z0 - 1, x, g);
            Cc[f] = ldrow(cur[f], z0, x, g);
        }
    }
    const float* c2 = a.coef[2];
    const float* c3 = a.coef[3];
    const float* c4 = a.coef[4];
    const float* c5 = a.coef[5];
    const float* c6 = a.coef[6];
    const float* c7 = a.coef[7];
    const bool edge = lane == 0 || lane == 31;
    const int xh = lane == 0 ? x0 - 1 : x0 + FW;          // halo column of the edge lanes
    int ro = z0 * ld + x;                                  // offset of row z, this lane
#pragma unroll
    for (int k = 0; k < FRZ; ++k, ro += ld) {
        const int z = z0 + k;
        if (SAFE || z < zn) {
            float4 P[NF];
            float lc[NF], rc[NF], lu[NF], ru[NF], ldn[NF], rdn[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                if (SAFE) {
                    D[f] = __ldg(reinterpret_cast<const float4*>(cur[f] + (ro + ld)));
                    P[f] = __ldg(reinterpret_cast<const float4*>(prv[f] + ro));
                    float hv = 0.f, hu = 0.f, hd = 0.f;
                    if (edge) {
                        const int ho = ro - x + xh;
                        hv = __ldg(cur[f] + ho);
                        if (FL & ST_F_XZ) { hu = __ldg(cur[f] + (ho - ld)); hd = __ldg(cur[f] + (ho + ld)); }
                    }
                    lc[f] = __shfl_up_sync(0xffffffffu, Cc[f].w, 1);
                    rc[f] = __shfl_down_sync(0xffffffffu, Cc[f].x, 1);
                    lc[f] = lane == 0 ? hv : lc[f];
                    rc[f] = lane == 31 ? hv : rc[f];
                    if (FL & ST_F_XZ) {
                        lu[f] = __shfl_up_sync(0xffffffffu, U[f].w, 1);
                        ru[f] = __shfl_down_sync(0xffffffffu, U[f].x, 1);
                        ldn[f] = __shfl_up_sync(0xffffffffu, D[f].w, 1);
                        rdn[f] = __shfl_down_sync(0xffffffffu, D[f].x, 1);
                        lu[f] = lane == 0 ? hu : lu[f];
                        ru[f] = lane == 31 ? hu : ru[f];
                        ldn[f] = lane == 0 ? hd : ldn[f];
                        rdn[f] = lane == 31 ? hd : rdn[f];
                    }
                } else {
                    D[f] = ldrow(cur[f], z + 1, x, g);
                    P[f] = ldrow(prv[f], z, x, g);
                    row_halo(Cc[f], cur[f], z, x0, lane, g, lc[f], rc[f]);
                    if (FL & ST_F_XZ) {
                        row_halo(U[f], cur[f], z - 1, x0, lane, g, lu[f], ru[f]);
                        row_halo(D[f], cur[f], z + 1, x0, lane, g, ldn[f], rdn[f]);
                    }
                }
            }
            float4 CXX = f4zero(), CZZ = f4zero(), CXZ = f4zero(), AX = f4zero(), AZ = f4zero(), M = f4zero();
            if (SAFE || x < ld) {
                CXX = __ldg(reinterpret_cast<const float4*>(c2 + ro));                              // ciso for ISO
                if (!(FL & ST_F_ISO) || (FL & ST_F_PML)) CZZ = __ldg(reinterpret_cast<const float4*>(c3 + ro));   // alpha for ISO|PML
                if (FL & ST_F_XZ) CXZ = __ldg(reinterpret_cast<const float4*>(c4 + ro));
                if (FL & ST_F_G1) {
                    AX = __ldg(reinterpret_cast<const float4*>(c5 + ro));
                    AZ = __ldg(reinterpret_cast<const float4*>(c6 + ro));
                }
                if (FL & ST_F_BORN) M = __ldg(reinterpret_cast<const float4*>(c7 + ro));
            }
            float4 Y[NF];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float alpha = (FL & ST_F_PML) ? f4get(CZZ, e) : 1.f;
                float A0 = 0.f;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const float c = f4get(Cc[f], e), n = f4get(U[f], e), s_ = f4get(D[f], e);
                    const float w = e == 0 ? lc[f] : f4get(Cc[f], e - 1);
                    const float ea = e == 3 ? rc[f] : f4get(Cc[f], e + 1);
                    float A;
                    if (FL & ST_F_ISO) A = f4get(CXX, e) * (((n - c) + (s_ - c)) + ((ea - c) + (w - c)));
                    else A = f4get(CXX, e) * ((ea - c) + (w - c)) + f4get(CZZ, e) * ((n - c) + (s_ - c));
                    if (FL & ST_F_XZ) {
                        const float nw = e == 0 ? lu[f] : f4get(U[f], e - 1), ne = e == 3 ? ru[f] : f4get(U[f], e + 1);
                        const float sw = e == 0 ? ldn[f] : f4get(D[f], e - 1), se = e == 3 ? rdn[f] : f4get(D[f], e + 1);
                        A += f4get(CXZ, e) * ((se - sw) - (ne - nw));
                    }
                    if (FL & ST_F_G1) A += f4get(AX, e) * (ea - w) + f4get(AZ, e) * (s_ - n);
                    if (f == 0) A0 = A;
                    else A += f4get(M, e) * A0;
                    f4set(Y[f], e, c + alpha * (c - f4get(P[f], e)) + A);
                }
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float* o = nxt[f] + ro;
                if (clean) {
                    *reinterpret_cast<float4*>(o) = Y[f];
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (x + e < g.nx) {
                            if (!HABC || !w2_in_frame(z, x + e, g)) o[e] = f4get(Y[f], e);
                        } else if (x + e < ld) {
                            o[e] = 0.f;          // keep the pitch padding zero (history slots are not pre-cleared)
                        }
                    }
                }
                U[f] = Cc[f];
                Cc[f] = D[f];
            }
        }
    }
}

template <int FL>
__device__ __forceinline__ void forward_fast_block(const W2Args& a, int bid, int nfx, int b, int tid) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool HABC = (FL & ST_F_HABC) != 0;
    const W2Geom g = a.g;
    const int warp = tid >> 5, lane = tid & 31;
    const int fz = bid / nfx, fx = bid - fz * nfx;
    const int x0 = fx * FW, zb0 = fz * FH;
    const int z0 = zb0 + warp * FRZ;
    if (z0 < g.nz) {
        const int zn = min(z0 + FRZ, g.nz);
        // warp tile completely inside the frame-free region -> unpredicated vector stores
        bool clean = x0 + FW <= g.nx;
        if (HABC) clean = clean && edge_depth(z0, x0, g) >= g.bw && edge_depth(zn - 1, x0 + FW - 1, g) >= g.bw &&
                          edge_depth(z0, x0 + FW - 1, g) >= g.bw && edge_depth(zn - 1, x0, g) >= g.bw;
        // a tile lying entirely inside the frame has nothing to store
        const bool dead = HABC && (zn <= (g.multiple ? 0 : g.bw) || z0 >= g.nz - g.bw || x0 + FW <= g.bw || x0 >= g.nx - g.bw);
        const bool safe = z0 >= 1 && z0 + FRZ + 1 <= g.nz && x0 >= 1 && x0 + FW + 1 <= g.nx;
        if (!dead) {
            if (safe) forward_fast_rows<FL, true>(a, g, b, x0, z0, zn, lane, clean);
            else forward_fast_rows<FL, false>(a, g, b, x0, z0, zn, lane, clean);
        }
    }
    forward_tail<NF>(a, b, zb0, zb0 + FH, x0, x0 + FW, tid,
                     [&](int z, int xx) { return !HABC || !w2_in_frame(z, xx, g); });
}

// ALL: every cell of the tile (the corner tiles of the TMA kernels), else only the frame cells
template <int FL, bool ALL = false, int UNR = 1>
__device__ __forceinline__ void forward_frame_block(const W2Args& a, int tz, int tx, int b, int tid, float (*s1)[SH][SW], int xoff = 0) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    const W2Geom g = a.g;
    const int x0 = tx * TX + xoff, z0 = tz * TZ;
    const long long boff = (long long)b * a.fs;
#pragma unroll
    for (int f = 0; f < NF; ++f) load_tile(s1[f], a.cur + f * a.cs + boff, z0, x0, g, tid);
    __syncthreads();
    const int x = x0 + (tid & (NTX - 1)), ty = tid / NTX;
    if (x < g.nx) {
#pragma unroll UNR
        for (int k = 0; k < RPT; ++k) {
            const int z = z0 + ty + k * NTY;
            if (z >= g.nz) break;
            if (!ALL && !w2_in_frame(z, x, g)) continue;
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
    if (ALL && x >= g.nx && x < g.ld) {                    // keep the pitch padding zero (history slots are not pre-cleared)
        for (int k = 0; k < RPT; ++k) {
            const int z = z0 + ty + k * NTY;
            if (z >= g.nz) break;
#pragma unroll
            for (int f = 0; f < NF; ++f) a.next[f * a.cs + boff + (long long)z * g.ld + x] = 0.f;
        }
    }
    forward_tail<NF>(a, b, z0, z0 + TZ, x0, x0 + TX, tid, [&](int z, int xx) { return ALL || w2_in_frame(z, xx, g); });
}


// The cells the TMA tiles do not cover (HABC: the four corners, where side ownership, corner diagonals and the
// wrap-around neighbour live) as generic TX-wide tiles: rows [0, sr0 TR) and [sr1 TR, nz) of the columns [0, TX) and
// [corner_xr, nx).  The rest of those rows inside the side tile columns belongs to masked TMA tiles.
struct CornerTiles { int rt, rb, count; };
__host__ __device__ inline CornerTiles corner_tiles(const W2Tma& tm, const W2Geom& g) {
    CornerTiles c{0, 0, 0};
    if (tm.sr1 <= tm.sr0) return c;
    c.rt = tm.sr0 * TR / TZ;
    c.rb = (g.nz - tm.sr1 * TR + TZ - 1) / TZ;
    c.count = (c.rt + c.rb) * 2;
    return c;
}
// blocks that serve the corner tiles of one launch.  With the precomputed frame taps (acoustic_habc: always, see
// st_wave2d_prepare) a corner tile is cut into TX*TZ/NT blocks of NT cells, one cell per thread, each walking a group of
// up to BSH shots with its taps in registers (forward_tap_block / adjoint_tap_block on a RectMap); without taps one
// generic per-cell block per (tile, shot).
// The ADJOINT TMA kernel runs its corner tiles on tap blocks (74.6 -> 69.9 us per 8-shot launch at the BASELINE size).  The
// FORWARD kernel keeps the generic per-cell corner tiles, one shot per block: a tap block walks all 8 shots of its group,
// and the one that holds the acquisition row then scans 8 x 1151 receivers in its epilogue -- longer than the whole 38 us
// forward launch (measured: 46 us with tap corners), while it hides inside the 70 us adjoint launch.
#ifndef ST_CORNER_TAP_FWD
#define ST_CORNER_TAP_FWD 0
#endif
constexpr int CORNER_SUB = TX * TZ / NT;                    // tap blocks per corner tile (4 rows x 64 columns each)
__host__ __device__ inline int corner_block_count(const CornerTiles& c, int B, bool tapped) {
    return tapped ? c.count * CORNER_SUB * band_groups(B) : c.count * B;
}
__device__ __forceinline__ void corner_tile_decode(const CornerTiles& c, const W2Tma& tm, const W2Geom& g, int i, int& tz, int& xoff) {
    const int ri = i >> 1;
    tz = ri < c.rt ? ri : tm.sr1 * TR / TZ + (ri - c.rt);
    xoff = (i & 1) ? corner_xr(tm, g) : 0;
}

// TMA block: one TR x TC tile, `tsh` shots pulled through a ring of bulk tensor loads.
// Tile kinds (block-uniform): 0 = frame-free rows (same arithmetic as forward_fast_rows<ISO>); +1 / -1 / +2 / -2 =
// the tile touches the straight top / bottom / left / right absorbing frame: every cell gets y + b (one - y)
// with the one-way extrapolation along the inward normal +z / -z / +x / -x (w2_habc_blend on a straight
// side; b == 0 on the frame-free cells of the tile).  The coefficient rows stay in registers across the shots.
template <int FL, int KIND>
__device__ __forceinline__ void forward_tma_tile(const W2Args& a, const W2Tma& tm, const TmaChunk& q, int tid, unsigned char* dsm,
                                                 uint64_t* bars) {
    constexpr bool PML = (FL & ST_F_PML) != 0, HABC = (FL & ST_F_HABC) != 0;
    constexpr int NS = ST_TMA_FWD_STAGES, STAGE = tma_fwd_stage<FL>(), R0 = tma_r0<FL>();
    const W2Geom& g = a.g;
    const int ld = g.ld;
    constexpr int kind = KIND;                              // block-uniform tile kind, compile-time here
    const int b_lo = q.b_lo, nsh = q.nsh;
    const int nitem = q.ntile * nsh;                        // item j = (tile j / nsh, shot j % nsh)
    const int warp = tid >> 5, lane = tid & 31;
    constexpr bool zdir = KIND == 1 || KIND == -1;
    const bool masked = q.mx0 > 0 || q.mx1 < (1 << 30);
    const bool xrag = q.x0 + q.ntile * q.dx + FW > g.nx;    // the tile column that crosses nx
    constexpr int hoff = zdir ? 2 : 1;                      // rows above z0 in the `cur` box
    const CUtensorMap* mcur = zdir ? &tm.u_h2 : &tm.u_h1;
    const CUtensorMap* mprev = kind ? &tm.u_h1 : &tm.u_core;
    auto issue = [&](int j) {
        const int stg = j % NS, ti = j / nsh, sh = j - ti * nsh;
        const int z0 = q.z0 + ti * q.dz, x0 = q.x0 + ti * q.dx;
        unsigned char* dst = dsm + stg * STAGE;
        st_mbar_expect_tx(&bars[stg], (zdir ? H2R : H1R) * HC * 4 + (kind ? H1R * HC * 4 : TMA_CORE_BYTES));
        st_tma_load_3d(dst, mcur, &bars[stg], x0 - XO, z0 - hoff, tm.pl_cur + b_lo + sh);
        if (kind) st_tma_load_3d(dst + R0, mprev, &bars[stg], x0 - XO, z0 - 1, tm.pl_prev + b_lo + sh);
        else st_tma_load_3d(dst + R0, mprev, &bars[stg], x0, z0, tm.pl_prev + b_lo + sh);
    };
    if (tid == 0)
        for (int j = 0; j < NS && j < nitem; ++j) issue(j);
    // `prev` box geometry: core box (frame-free tiles) or 1-deep halo box (frame tiles)
    constexpr int ppitch = KIND ? HC : TC, poff = KIND ? HC + XO : 0;   // offset of (z0, x0)
    float4 ci[2], al[2], bb[2], rr[2];
    bool zok[2];
    int z0 = q.z0, x0 = q.x0, zr = 0, x = 0;
    for (int j = 0, sh = 0; j < nitem; ++j) {
        if (sh == 0) {                                      // new tile: its coefficient rows
            zr = z0 + 2 * warp;
            x = x0 + 4 * lane;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                zok[k] = zr + k < g.nz && x < ld;
                const int o = (zr + k) * ld + x;
                ci[k] = zok[k] ? __ldg(reinterpret_cast<const float4*>(a.coef[2] + o)) : f4zero();
                al[k] = (PML && zok[k]) ? __ldg(reinterpret_cast<const float4*>(a.coef[3] + o)) : f4zero();
                bb[k] = (HABC && kind && zok[k]) ? __ldg(reinterpret_cast<const float4*>(a.coef[1] + o)) : f4zero();
                rr[k] = (HABC && kind && zok[k]) ? __ldg(reinterpret_cast<const float4*>(a.coef[0] + o)) : f4zero();
            }
        }
        const int stg = j % NS, b = b_lo + sh;
        st_mbar_wait(&bars[stg], (j / NS) & 1);
        const float* h1 = reinterpret_cast<const float*>(dsm + stg * STAGE);
        const float* h2 = reinterpret_cast<const float*>(dsm + stg * STAGE + R0) + poff;
        float* out = a.next + (long long)b * a.fs + (zr * ld + x);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float* rowc = h1 + (2 * warp + k + hoff) * HC;                                    // box row of z = zr + k
            const float* rowp = h2 + (2 * warp + k) * ppitch;
            const float4 C = *reinterpret_cast<const float4*>(rowc + XO + 4 * lane);
            const float4 U = *reinterpret_cast<const float4*>(rowc - HC + XO + 4 * lane);
            const float4 D = *reinterpret_cast<const float4*>(rowc + HC + XO + 4 * lane);
            const float4 P = *reinterpret_cast<const float4*>(rowp + 4 * lane);
            float lc = __shfl_up_sync(0xffffffffu, C.w, 1), rc = __shfl_down_sync(0xffffffffu, C.x, 1);
            {   // halo columns: broadcast shared loads + selects (no divergent branch in the hot loop)
                const float hl = rowc[XO - 1], hr = rowc[XO + TC];
                lc = lane == 0 ? hl : lc;
                rc = lane == 31 ? hr : rc;
            }
            float4 Y;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float c = f4get(C, e), n = f4get(U, e), s_ = f4get(D, e);
                const float w = e == 0 ? lc : f4get(C, e - 1);
                const float ea = e == 3 ? rc : f4get(C, e + 1);
                const float A = f4get(ci[k], e) * (((n - c) + (s_ - c)) + ((ea - c) + (w - c)));
                const float alpha = PML ? f4get(al[k], e) : 1.f;
                f4set(Y, e, (!PML || x + e < g.nx) ? c + alpha * (c - f4get(P, e)) + A : 0.f);   // PML: the tile column past nx
            }
            if (HABC && kind) {
                float4 A1, A2, P1;                          // h1 one / two cells inward, h2 one cell inward
                if (zdir) {
                    A1 = kind > 0 ? D : U;
                    A2 = *reinterpret_cast<const float4*>(rowc + 2 * kind * HC + XO + 4 * lane);
                    P1 = *reinterpret_cast<const float4*>(rowp + kind * ppitch + 4 * lane);
                } else if (kind > 0) {
                    float rc2 = __shfl_down_sync(0xffffffffu, C.y, 1), prc = __shfl_down_sync(0xffffffffu, P.x, 1);
                    if (lane == 31) { rc2 = rowc[XO + 1 + TC]; prc = rowp[TC]; }
                    A1 = f4shr(C, rc);
                    A2 = make_float4(C.z, C.w, rc, rc2);
                    P1 = f4shr(P, prc);
                } else {
                    float lc2 = __shfl_up_sync(0xffffffffu, C.z, 1), plc = __shfl_up_sync(0xffffffffu, P.w, 1);
                    if (lane == 0) { lc2 = rowc[XO - 2]; plc = rowp[-1]; }
                    A1 = f4shl(C, lc);
                    A2 = make_float4(lc2, lc, C.x, C.y);
                    P1 = f4shl(P, plc);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a0 = f4get(C, e), a1 = f4get(A1, e), a2 = f4get(A2, e);
                    const float p0 = f4get(P, e), p1 = f4get(P1, e);
                    const float r = f4get(rr[k], e), lam = 2.f * r, mu = r * r;
                    const float base = a0 + (a0 - p0);
                    const float dlam = (a1 - a0) - (p1 - p0);
                    const float dmu = (a1 - a0) - (a2 - a1);
                    const float one = base + lam * dlam + mu * dmu;
                    const float y = f4get(Y, e);
                    f4set(Y, e, (!xrag || x + e < g.nx) ? y + f4get(bb[k], e) * (one - y) : 0.f);   // pitch padding stays zero
                }
            }
            if (zok[k]) {
                if (!masked) {
                    *reinterpret_cast<float4*>(out + k * ld) = Y;
                } else {                                    // corner rows: only the columns this tile owns (+ zero padding)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (x + e >= q.mx0 && x + e < q.mx1) out[k * ld + e] = f4get(Y, e);
                        else if (x + e >= g.nx && x + e < ld) out[k * ld + e] = 0.f;
                    }
                }
            }
        }
        forward_tail<1>(a, b, z0, z0 + TR, x0, x0 + FW, tid, [&](int, int xx) { return xx >= q.mx0 && xx < q.mx1; });
        __syncthreads();                                   // every warp is done with this stage
        if (tid == 0 && j + NS < nitem) issue(j + NS);
        if (++sh == nsh) { sh = 0; z0 += q.dz; x0 += q.dx; }
    }
}

template <int FL>
__device__ __forceinline__ void forward_tma_block(const W2Args& a, const W2Tma& tm, int nfx, int bid, int tid, unsigned char* dsm) {
    constexpr int NS = ST_TMA_FWD_STAGES;
    __shared__ __align__(8) uint64_t bars[NS];
    const TmaChunk q = tma_block_decode(tm, a.g, (FL & ST_F_HABC) != 0, nfx, a.B, bid);
    if (q.ntile == 0) return;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) st_mbar_init(&bars[s], 1);
        st_mbar_init_fence();
    }
    __syncthreads();
    st_pdl_wait();                                          // the previous step's fields are complete from here on
    if constexpr ((FL & ST_F_HABC) != 0) {
        if (q.kind == 1) { forward_tma_tile<FL, 1>(a, tm, q, tid, dsm, bars); return; }
        if (q.kind == -1) { forward_tma_tile<FL, -1>(a, tm, q, tid, dsm, bars); return; }
        if (q.kind == 2) { forward_tma_tile<FL, 2>(a, tm, q, tid, dsm, bars); return; }
        if (q.kind == -2) { forward_tma_tile<FL, -2>(a, tm, q, tid, dsm, bars); return; }
    }
    forward_tma_tile<FL, 0>(a, tm, q, tid, dsm, bars);
}

#ifndef ST_FWD_MINB
#define ST_FWD_MINB 4
#endif
template <int FL>
__global__ void __launch_bounds__(NT, ST_FWD_MINB) wave2d_forward_kernel(const W2Args a, int nfx, int nfast, BandTiles bt) {
    st_pdl_launch_dependents();                             // (no-ops unless launched with programmatic stream serialization)
    st_pdl_wait();
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool HABC = (FL & ST_F_HABC) != 0;
    __shared__ float s1[HABC ? NF : 1][HABC ? SH : 1][SW];
    const int bid = blockIdx.x, tid = threadIdx.x;
    // grid.x = [frame blocks] ++ [fast blocks x shots]; the (slower) frame blocks get the low ids so
    // they are scheduled first.  Tapped frame blocks walk all shots themselves.
    const bool tapped = st_flags_tapped(FL) && a.taps != nullptr;
    const int ngrp = band_groups(a.B), gsh = band_group_shots(a.B);
    const int nstrip = (tapped && st_flags_stripped_fwd(FL)) ? strip_blocks(strip_geom(a.g, a.g.bw)) : 0;
    const int nframe = HABC ? (bt.count + nstrip) * (tapped ? ngrp : a.B) : 0;
    if (bid >= nframe) {
        const int q = bid - nframe;
        forward_fast_block<FL>(a, q % nfast, nfx, q / nfast, tid);
    } else if (tapped) {
        const int per = bt.count + nstrip, grp = bid / per, k = bid - grp * per;
        const int b_lo = grp * gsh, b_hi = min(b_lo + gsh, a.B);
        if (k < nstrip) forward_strip_block<FL>(a, k, b_lo, b_hi, tid);
        else forward_band_block<FL>(a, k - nstrip, b_lo, b_hi, tid);
    } else {
        if constexpr (HABC) {
            int tz, tx;
            const int b = bid / bt.count;
            band_tile_decode(bt, bid - b * bt.count, tz, tx);
            forward_frame_block<FL>(a, tz, tx, b, tid, reinterpret_cast<float (*)[SH][SW]>(s1));
        }
    }
}

#ifdef ST_DBG_TIMELINE
__device__ unsigned long long g_dbg_tl[4 * 8192];
__device__ __forceinline__ unsigned long long dbg_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned dbg_smid() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
#endif
// TMA forward kernel: grid.x = [corner tiles x shots (generic per-cell code, one shot each: short blocks that
// start first)] ++ [TMA chunk blocks].
template <int FL>
__global__ void __launch_bounds__(NT, tma_fwd_minb<FL>()) wave2d_forward_tma_kernel(const W2Args a, int nfx, const __grid_constant__ W2Tma tm) {
    extern __shared__ __align__(128) unsigned char dsm[];
    const int bid = blockIdx.x, tid = threadIdx.x;
    const CornerTiles ct = corner_tiles(tm, a.g);
    const bool ctap = ST_CORNER_TAP_FWD && st_flags_tapped(FL) && a.taps != nullptr;
    const int ncorner = corner_block_count(ct, a.B, ctap);
    st_pdl_launch_dependents();
    if (bid < ncorner) {
        st_pdl_wait();
        if (ST_DBG_SKIP & 32) return;
        if constexpr ((FL & ST_F_HABC) != 0) {
            int tz, tx;
            if (ctap) {
                const int per = ct.count * CORNER_SUB, grp = bid / per, r = bid - grp * per;
                corner_tile_decode(ct, tm, a.g, r / CORNER_SUB, tz, tx);
                const RectMap map{a.g, tz * TZ + (r % CORNER_SUB) * (NT / TX), tx, TX};
                forward_tap_block<FL>(a, map, grp * band_group_shots(a.B), min((grp + 1) * band_group_shots(a.B), a.B), tid);
            } else {
                const int b = bid / ct.count;
                corner_tile_decode(ct, tm, a.g, bid - b * ct.count, tz, tx);
                forward_frame_block<FL, true, ST_CORNER_UNROLL>(a, tz, 0, b, tid, reinterpret_cast<float (*)[SH][SW]>(dsm), tx);
            }
        }
        return;
    }
    if (ST_DBG_SKIP & 8) return;
#ifdef ST_DBG_TIMELINE
    unsigned long long t0 = dbg_now();
#endif
    forward_tma_block<FL>(a, tm, nfx, bid - ncorner, tid, dsm);
#ifdef ST_DBG_TIMELINE
    if (tid == 0 && bid < 8192) { g_dbg_tl[4 * bid] = t0; g_dbg_tl[4 * bid + 1] = dbg_now(); g_dbg_tl[4 * bid + 2] = dbg_smid(); g_dbg_tl[4 * bid + 3] = 1; }
#endif
}

// ------------------------------------------------------------------------------ adjoint
// which of the 7 gradient accumulators (r,cxx,czz,cxz,ax,az,m) a flag set touches
template <int FL>
__host__ __device__ constexpr bool grad_used(int q) {
    return q == 0 ? (FL & ST_F_HABC) != 0
         : q == 1 ? true
         : q == 2 ? !(FL & ST_F_ISO)
         : q == 3 ? (FL & ST_F_XZ) != 0
         : (q == 4 || q == 5) ? (FL & ST_F_G1) != 0
         : (FL & ST_F_BORN) != 0;
}
// flag sets with a vectorised adjoint fast path (all of them; the generic whole-grid path below is kept for reference)
template <int FL>
__host__ __device__ constexpr bool adj_fast() { return true; }
// ... of which the two acoustic ones keep their gradient partial sums in shared memory
template <int FL>
__host__ __device__ constexpr bool adj_iso_only() { return FL == (ST_F_ISO | ST_F_PML) || FL == (ST_F_ISO | ST_F_HABC); }

// receiver-adjoint scatter + source-amplitude gradient for the cells this block stored
template <int NF, class Own>
__device__ __forceinline__ void adjoint_tail(const W2Args& a, int b, int z0, int zn, int x0, int xn, int tid, Own owns) {
    if (zn <= a.row_lo || z0 > a.row_hi) return;          // no source / receiver in these rows (block-uniform)
    const W2Geom& g = a.g;
    const long long boff = (long long)b * a.fs;
    __shared__ int s_cnt, s_rows[FH];
    __syncthreads();
    if (a.rec_adj) {
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        const int nrow = min(zn, g.nz) - z0;
        if (tid < nrow) {
            const int row = b * g.nz + z0 + tid;
            if (a.row_start[row + 1] > a.row_start[row]) s_rows[atomicAdd(&s_cnt, 1)] = z0 + tid;
        }
        __syncthreads();
        const int cnt = s_cnt;
        for (int i = 0; i < cnt; ++i) {
            const int z = s_rows[i];
            const int lo = a.row_start[b * g.nz + z], hi = a.row_start[b * g.nz + z + 1];
            for (int r = lo + tid; r < hi; r += NT) {
                const int rx = a.rec_x[r];
                if (rx >= x0 && rx < xn && owns(z, rx)) {
                    const long long o = (long long)a.rec_orig[r] * a.nchan;
                    for (int ch = 0; ch < a.nchan; ++ch)
                        atomicAdd(a.lam0 + a.chan_f[ch] * a.cs + boff + (long long)z * g.ld + rx, a.rec_adj[o + ch]);
                }
            }
        }
    }
    if (a.gamp) {
        __syncthreads();
        for (int s = tid; s < a.ns; s += NT) {
            if (a.src_b[s] != b) continue;
            const int sz = a.src_z[s], sx = a.src_x[s];
            if (sz >= z0 && sz < zn && sx >= x0 && sx < xn && owns(sz, sx)) {
                float v = 0.f;
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    if (a.src_fmask >> f & 1) v += a.lam0[f * a.cs + boff + (long long)sz * g.ld + sx];
                a.gamp[s] = v;
            }
        }
    }
}

// ISO fast path: Lam_i = (1+alpha) L1 + lap(ciso L1) - alpha L2 ; g_ciso += L1 * lap(S_i).
// Gradient partial sums of the block's shots live in shared memory (one float4 per lane and
// row, conflict-free) so the register budget stays small and HBM sees one read-modify-write
// of the gradient plane per `bchunk` shots.
template <int FL, bool SAFE, class Own>
__device__ __forceinline__ void adjoint_fast_rows(const W2Args& a, const W2Geom& g, int b, int x0, int z0, int zn,
                                                  int lane, bool clean, bool want_grad, float4* gsl, Own owns) {
    constexpr bool PML = (FL & ST_F_PML) != 0;
    const int x = x0 + 4 * lane;
    const int ld = g.ld;
    const long long boff = (long long)b * a.fs;
    const float* l1 = a.lam1 + boff;
    const float* l2 = a.lam2 + boff;
    const float* S = a.s1 + boff;
    float* l0 = a.lam0 + boff;
    const float* ciso = a.coef[2];
    const float* alp = a.coef[3];
    const bool edge = lane == 0 || lane == 31;
    const int xh = lane == 0 ? x0 - 1 : x0 + FW;
    auto ld4 = [&](const float* p, int z, int off) -> float4 {
        if (SAFE) return __ldg(reinterpret_cast<const float4*>(p + off));
        return ldrow(p, z, x, g);
    };
    auto ld1 = [&](const float* p, int z, int off) -> float {        // halo column of an edge lane
        if (SAFE) return __ldg(p + off);
        return (z >= 0 && z < g.nz && xh >= 0 && xh < g.nx) ? __ldg(p + off) : 0.f;
    };
    auto mul4 = [](const float4& u, const float4& v) { return make_float4(u.x * v.x, u.y * v.y, u.z * v.z, u.w * v.w); };
    int ro = z0 * ld + x;
    float4 lC = ld4(l1, z0, ro), lD;
    float4 wU = mul4(ld4(ciso, z0 - 1, ro - ld), ld4(l1, z0 - 1, ro - ld));
    float4 wC = mul4(ld4(ciso, z0, ro), lC);
    float4 sU = ld4(S, z0 - 1, ro - ld), sC = ld4(S, z0, ro);
#pragma unroll
    for (int k = 0; k < FRZ; ++k, ro += ld) {
        const int z = z0 + k;
        if (SAFE || z < zn) {
            lD = ld4(l1, z + 1, ro + ld);
            const float4 wD = mul4(ld4(ciso, z + 1, ro + ld), lD);
            const float4 sD = ld4(S, z + 1, ro + ld);
            const float4 p2 = ld4(l2, z, ro);
            float4 al = f4zero();
            if (PML) al = ld4(alp, z, ro);
            float hw = 0.f, hs = 0.f;
            if (edge) {
                const int ho = ro - x + xh;
                hw = ld1(ciso, z, ho) * ld1(l1, z, ho);
                hs = ld1(S, z, ho);
            }
            float wl = __shfl_up_sync(0xffffffffu, wC.w, 1), wr = __shfl_down_sync(0xffffffffu, wC.x, 1);
            float sl = __shfl_up_sync(0xffffffffu, sC.w, 1), sr = __shfl_down_sync(0xffffffffu, sC.x, 1);
            wl = lane == 0 ? hw : wl;
            wr = lane == 31 ? hw : wr;
            sl = lane == 0 ? hs : sl;
            sr = lane == 31 ? hs : sr;
            float4 out, gq;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float c = f4get(wC, e);
                const float w = e == 0 ? wl : f4get(wC, e - 1), ea = e == 3 ? wr : f4get(wC, e + 1);
                const float lapw = ((f4get(wU, e) - c) + (f4get(wD, e) - c)) + ((ea - c) + (w - c));
                const float alpha = PML ? f4get(al, e) : 1.f;
                const float l1c = f4get(lC, e);
                f4set(out, e, (1.f + alpha) * l1c + lapw - alpha * f4get(p2, e));
                const float sc = f4get(sC, e);
                const float sw_ = e == 0 ? sl : f4get(sC, e - 1), se_ = e == 3 ? sr : f4get(sC, e + 1);
                const float laps = ((f4get(sU, e) - sc) + (f4get(sD, e) - sc)) + ((se_ - sc) + (sw_ - sc));
                f4set(gq, e, l1c * laps);
            }
            float* o = l0 + ro;
            if (clean) {
                *reinterpret_cast<float4*>(o) = out;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (x + e < g.nx) {
                        if (owns(z, x + e)) o[e] = f4get(out, e);
                    } else if (x + e < ld) {
                        o[e] = 0.f;
                    }
                }
            }
            if (want_grad) {
                float4 acc = gsl[k * (FW / 4)];
                acc.x += gq.x; acc.y += gq.y; acc.z += gq.z; acc.w += gq.w;
                gsl[k * (FW / 4)] = acc;
            }
            wU = wC; wC = wD; sU = sC; sC = sD; lC = lD;
        }
    }
}


#ifndef ST_GEN_SMEM_GRAD
#define ST_GEN_SMEM_GRAD 1                  // 0: the single-field non-ISO equations add their gradients straight to the global plane
#endif
constexpr int GPL4 = NWARP * FRZ * FW / 4;      // float4s per shared-memory gradient plane of a fast block
// position of gradient slot `slot` among the slots the flag set uses (= its shared-memory plane), and their number
template <int FL>
__host__ __device__ constexpr int adj_gslot_index(int slot) {
    int n = 0;
    for (int q = 1; q < slot; ++q) if (grad_used<FL>(q)) ++n;
    return n;
}
template <int FL>
__host__ __device__ constexpr int adj_smem_planes() { return adj_gslot_index<FL>(7); }
// Single-field, non-Born flag sets (vti_habc2, tti_habc, acoustic_fwim_habc): interior cells
//   Lam_i = 2 L1 - L2 + dxx(cxx L1) + dzz(czz L1) + dxz^T(cxz L1) - dx(ax L1) - dz(az L1)
// evaluated on rows of the coefficient-times-cotangent products, which are formed once per row
// when it is loaded and marched through 3-row register pipelines; coefficient gradients
// (imaging condition) are added straight to the block's gradient plane.
template <int FL, bool SAFE, class Own>
__device__ __forceinline__ void adjoint_fast_rows_gen(const W2Args& a, const W2Geom& g, int b, int chunk, int x0, int z0,
                                                      int zn, int lane, bool clean, bool want_grad, Own owns,
                                                      float4* gsl = nullptr) {       // gsl: shared-memory gradient planes
    constexpr bool ISO = (FL & ST_F_ISO) != 0, XZ = (FL & ST_F_XZ) != 0, G1 = (FL & ST_F_G1) != 0;
    const int x = x0 + 4 * lane;
    const int ld = g.ld;
    const long long boff = (long long)b * a.fs, plane = (long long)g.nz * ld;
    const float* l1 = a.lam1 + boff;
    const float* l2 = a.lam2 + boff;
    const float* S = a.s1 + boff;
    float* l0 = a.lam0 + boff;
    float* gb = want_grad ? a.gacc + (long long)chunk * 7 * plane : nullptr;
    const bool edge = lane == 0 || lane == 31;
    const int xh = lane == 0 ? x0 - 1 : x0 + FW;
    // product rows of row z:  PA = cxx*L (ciso*L for ISO), PB = czz*L, PC = cxz*L / az*L, PD = ax*L
    struct Prod { float4 l, a, b, c, d; float al, ar, cl, cr, dl, dr; };
    auto load_prod = [&](int z) {
        Prod p;
        p.l = ldrow_s<SAFE>(l1, z, x, g);
        const float4 ca = ldrow_s<SAFE>(a.coef[2], z, x, g);
        p.a = f4mul(ca, p.l);
        p.b = ISO ? p.a : f4mul(ldrow_s<SAFE>(a.coef[3], z, x, g), p.l);
        p.c = XZ ? f4mul(ldrow_s<SAFE>(a.coef[4], z, x, g), p.l) : (G1 ? f4mul(ldrow_s<SAFE>(a.coef[6], z, x, g), p.l) : f4zero());
        p.d = G1 ? f4mul(ldrow_s<SAFE>(a.coef[5], z, x, g), p.l) : f4zero();
        // x-neighbours of the products (shuffles; the warp's edge lanes read coefficient and cotangent)
        p.al = __shfl_up_sync(0xffffffffu, p.a.w, 1); p.ar = __shfl_down_sync(0xffffffffu, p.a.x, 1);
        p.cl = p.cr = p.dl = p.dr = 0.f;
        if (XZ) { p.cl = __shfl_up_sync(0xffffffffu, p.c.w, 1); p.cr = __shfl_down_sync(0xffffffffu, p.c.x, 1); }
        if (G1) { p.dl = __shfl_up_sync(0xffffffffu, p.d.w, 1); p.dr = __shfl_down_sync(0xffffffffu, p.d.x, 1); }
        if (edge) {
            float va = 0.f, vc = 0.f, vd = 0.f;
            if (SAFE || (z >= 0 && z < g.nz && xh >= 0 && xh < g.nx)) {
                const int o = z * ld + xh;
                const float lv = __ldg(l1 + o);
                va = __ldg(a.coef[2] + o) * lv;
                if (XZ) vc = __ldg(a.coef[4] + o) * lv;
                if (G1) vd = __ldg(a.coef[5] + o) * lv;
            }
            if (lane == 0) { p.al = va; p.cl = vc; p.dl = vd; } else { p.ar = va; p.cr = vc; p.dr = vd; }
        }
        return p;
    };
    struct SRow { float4 s; float l, r; };
    auto load_s = [&](int z) {
        SRow q;
        q.s = ldrow_s<SAFE>(S, z, x, g);
        row_halo(q.s, S, z, x0, lane, g, q.l, q.r);
        return q;
    };
    Prod U = load_prod(z0 - 1), C = load_prod(z0), D;
    SRow sU, sC, sD;
    if (want_grad) { sU = load_s(z0 - 1); sC = load_s(z0); }
#pragma unroll
    for (int k = 0; k < FRZ; ++k) {
        const int z = z0 + k;
        if (z < zn) {
            D = load_prod(z + 1);
            const float4 p2 = ldrow_s<SAFE>(l2, z, x, g);
            if (want_grad) sD = load_s(z + 1);
            float4 out, g1v = f4zero(), g2v = f4zero(), g3v = f4zero(), g4v = f4zero(), g5v = f4zero();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lc = f4get(C.l, e);
                const float ac = f4get(C.a, e);
                const float aw = e == 0 ? C.al : f4get(C.a, e - 1), ae = e == 3 ? C.ar : f4get(C.a, e + 1);
                float acc = 2.f * lc - f4get(p2, e);
                acc += ((ae - ac) + (aw - ac));                                              // dxx of cxx*L (or ciso*L)
                acc += ((f4get(U.b, e) - f4get(C.b, e)) + (f4get(D.b, e) - f4get(C.b, e)));  // dzz of czz*L (or ciso*L)
                if (XZ) {
                    const float uw = e == 0 ? U.cl : f4get(U.c, e - 1), ue = e == 3 ? U.cr : f4get(U.c, e + 1);
                    const float dw = e == 0 ? D.cl : f4get(D.c, e - 1), de = e == 3 ? D.cr : f4get(D.c, e + 1);
                    acc += (uw - ue) - (dw - de);
                }
                if (G1) {
                    const float dwv = e == 0 ? C.dl : f4get(C.d, e - 1), dev = e == 3 ? C.dr : f4get(C.d, e + 1);
                    acc += (dwv - dev) + (f4get(U.c, e) - f4get(D.c, e));                    // -dx(ax L) - dz(az L)
                }
                f4set(out, e, acc);
                if (want_grad) {
                    const float sc = f4get(sC.s, e);
                    const float sw_ = e == 0 ? sC.l : f4get(sC.s, e - 1), se_ = e == 3 ? sC.r : f4get(sC.s, e + 1);
                    const float sn = f4get(sU.s, e), ss = f4get(sD.s, e);
                    const float sxx = (se_ - sc) + (sw_ - sc), szz = (sn - sc) + (ss - sc);
                    if (ISO) f4set(g1v, e, lc * (szz + sxx));
                    else { f4set(g1v, e, lc * sxx); f4set(g2v, e, lc * szz); }
                    if (XZ) {
                        const float nw = e == 0 ? sU.l : f4get(sU.s, e - 1), ne = e == 3 ? sU.r : f4get(sU.s, e + 1);
                        const float sw2 = e == 0 ? sD.l : f4get(sD.s, e - 1), se2 = e == 3 ? sD.r : f4get(sD.s, e + 1);
                        f4set(g3v, e, lc * ((se2 - sw2) - (ne - nw)));
                    }
                    if (G1) { f4set(g4v, e, lc * (se_ - sw_)); f4set(g5v, e, lc * (ss - sn)); }
                }
            }
            const int ro = z * ld + x;
            if (gsl != nullptr) {
                // gradient partial sums of the block's shots in shared memory; cells the block does not own are never flushed
                if (want_grad) {
                    float4* p4 = gsl + k * (FW / 4);
                    auto add = [&](int slot_index, const float4& v) { p4[slot_index * GPL4] = f4add(p4[slot_index * GPL4], v); };
                    add(adj_gslot_index<FL>(1), g1v);
                    if (!ISO) add(adj_gslot_index<FL>(2), g2v);
                    if (XZ) add(adj_gslot_index<FL>(3), g3v);
                    if (G1) { add(adj_gslot_index<FL>(4), g4v); add(adj_gslot_index<FL>(5), g5v); }
                }
                if (clean) {
                    *reinterpret_cast<float4*>(l0 + ro) = out;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (x + e < g.nx && owns(z, x + e)) l0[ro + e] = f4get(out, e);
                        else if (x + e >= g.nx && x + e < ld) l0[ro + e] = 0.f;
                    }
                }
            } else if (clean) {
                // the whole tile is frame-free and inside the domain: 128-bit stores and gradient read-modify-writes
                *reinterpret_cast<float4*>(l0 + ro) = out;
                if (want_grad) {
                    auto rmw = [&](int slot, const float4& v) {
                        float4* p4 = reinterpret_cast<float4*>(gb + slot * plane + ro);
                        *p4 = f4add(*p4, v);
                    };
                    rmw(1, g1v);
                    if (!ISO) rmw(2, g2v);
                    if (XZ) rmw(3, g3v);
                    if (G1) { rmw(4, g4v); rmw(5, g5v); }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool mine = x + e < g.nx && owns(z, x + e);
                    if (mine) {
                        l0[ro + e] = f4get(out, e);
                        if (want_grad) {
                            gb[plane + ro + e] += f4get(g1v, e);
                            if (!ISO) gb[2 * plane + ro + e] += f4get(g2v, e);
                            if (XZ) gb[3 * plane + ro + e] += f4get(g3v, e);
                            if (G1) { gb[4 * plane + ro + e] += f4get(g4v, e); gb[5 * plane + ro + e] += f4get(g5v, e); }
                        }
                    } else if (x + e >= g.nx && x + e < ld) {
                        l0[ro + e] = 0.f;
                    }
                }
            }
            U = C; C = D;
            if (want_grad) { sU = sC; sC = sD; }
        }
    }
}

// Born pairs (acoustic_vti_lsrtm_habc, acoustic_tti_lsrtm_habc): interior cells of both fields
//   Lam_i(f) = 2 L1(f) - L2(f) + dxx(cxx le_f) + dzz(czz le_f) + dxz^T(cxz le_f),   le_1 = L1(1),  le_0 = L1(0) + m L1(1)
// (the background field also receives the scattered field's cotangent through m A[p]); gradients
//   g_cxx += le_f dxx S_f,  g_czz += le_f dzz S_f,  g_cxz += le_f dxz S_f  (both fields),   g_m += L1(1) A[S_0].
// Same row-marching register pipeline as adjoint_fast_rows_gen, one field after the other.
template <int FL, class Own>
__device__ __forceinline__ void adjoint_fast_rows_born(const W2Args& a, const W2Geom& g, int b, int chunk, int x0, int z0,
                                                       int zn, int lane, bool clean, bool want_grad, Own owns,
                                                       float4* gsl = nullptr) {      // gsl: shared-memory gradient planes
    constexpr bool XZ = (FL & ST_F_XZ) != 0;
    const int x = x0 + 4 * lane;
    const int ld = g.ld;
    const long long boff = (long long)b * a.fs, plane = (long long)g.nz * ld;
    float* gb = want_grad ? a.gacc + (long long)chunk * 7 * plane : nullptr;
    const bool edge = lane == 0 || lane == 31;
    const int xh = lane == 0 ? x0 - 1 : x0 + FW;
    const float* l1s = a.lam1 + a.cs + boff;                // cotangent of the scattered field
#pragma unroll 1
    for (int f = 1; f >= 0; --f) {
        const float* l1 = a.lam1 + f * a.cs + boff;
        const float* l2 = a.lam2 + f * a.cs + boff;
        const float* S = a.s1 + f * a.cs + boff;
        float* l0 = a.lam0 + f * a.cs + boff;
        // product rows of row z:  PA = cxx*le, PB = czz*le, PC = cxz*le  (+ x-neighbours of PA, PC)
        struct Prod { float4 l, a, b, c; float al, ar, cl, cr; };
        auto load_prod = [&](int z) {
            Prod p;
            p.l = ldrow(l1, z, x, g);
            if (f == 0) p.l = f4fma(ldrow(a.coef[7], z, x, g), ldrow(l1s, z, x, g), p.l);
            p.a = f4mul(ldrow(a.coef[2], z, x, g), p.l);
            p.b = f4mul(ldrow(a.coef[3], z, x, g), p.l);
            p.c = XZ ? f4mul(ldrow(a.coef[4], z, x, g), p.l) : f4zero();
            p.al = __shfl_up_sync(0xffffffffu, p.a.w, 1); p.ar = __shfl_down_sync(0xffffffffu, p.a.x, 1);
            p.cl = p.cr = 0.f;
            if (XZ) { p.cl = __shfl_up_sync(0xffffffffu, p.c.w, 1); p.cr = __shfl_down_sync(0xffffffffu, p.c.x, 1); }
            if (edge) {
                float va = 0.f, vc = 0.f;
                if (z >= 0 && z < g.nz && xh >= 0 && xh < g.nx) {
                    const int o = z * ld + xh;
                    float lv = __ldg(l1 + o);
                    if (f == 0) lv += __ldg(a.coef[7] + o) * __ldg(l1s + o);
                    va = __ldg(a.coef[2] + o) * lv;
                    if (XZ) vc = __ldg(a.coef[4] + o) * lv;
                }
                if (lane == 0) { p.al = va; p.cl = vc; } else { p.ar = va; p.cr = vc; }
            }
            return p;
        };
        struct SRow { float4 s; float l, r; };
        auto load_s = [&](int z) {
            SRow q;
            q.s = ldrow(S, z, x, g);
            row_halo(q.s, S, z, x0, lane, g, q.l, q.r);
            return q;
        };
        Prod U = load_prod(z0 - 1), C = load_prod(z0), D;
        SRow sU, sC, sD;
        if (want_grad) { sU = load_s(z0 - 1); sC = load_s(z0); }
#pragma unroll
        for (int k = 0; k < FRZ; ++k) {
            const int z = z0 + k;
            if (z < zn) {
                D = load_prod(z + 1);
                const float4 p2 = ldrow(l2, z, x, g);
                const float4 lraw = f == 0 ? ldrow(l1, z, x, g) : C.l;     // 2 L1(f): the field's own cotangent
                if (want_grad) sD = load_s(z + 1);
                float4 out, g1v = f4zero(), g2v = f4zero(), g3v = f4zero(), g6v = f4zero();
                float4 ca = f4zero(), cb = f4zero(), cc = f4zero(), ls = f4zero();
                if (want_grad && f == 0) {                                   // g_m needs A[S_0] and L1(1) at the cell
                    ca = ldrow(a.coef[2], z, x, g); cb = ldrow(a.coef[3], z, x, g);
                    if (XZ) cc = ldrow(a.coef[4], z, x, g);
                    ls = ldrow(l1s, z, x, g);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float ac = f4get(C.a, e);
                    const float aw = e == 0 ? C.al : f4get(C.a, e - 1), ae = e == 3 ? C.ar : f4get(C.a, e + 1);
                    float acc = 2.f * f4get(lraw, e) - f4get(p2, e);
                    acc += ((ae - ac) + (aw - ac));                                              // dxx of cxx*le
                    acc += ((f4get(U.b, e) - f4get(C.b, e)) + (f4get(D.b, e) - f4get(C.b, e)));  // dzz of czz*le
                    if (XZ) {
                        const float uw = e == 0 ? U.cl : f4get(U.c, e - 1), ue = e == 3 ? U.cr : f4get(U.c, e + 1);
                        const float dw = e == 0 ? D.cl : f4get(D.c, e - 1), de = e == 3 ? D.cr : f4get(D.c, e + 1);
                        acc += (uw - ue) - (dw - de);
                    }
                    f4set(out, e, acc);
                    if (want_grad) {
                        const float le = f4get(C.l, e);
                        const float sc = f4get(sC.s, e);
                        const float sw_ = e == 0 ? sC.l : f4get(sC.s, e - 1), se_ = e == 3 ? sC.r : f4get(sC.s, e + 1);
                        const float sn = f4get(sU.s, e), ss = f4get(sD.s, e);
                        const float sxx = (se_ - sc) + (sw_ - sc), szz = (sn - sc) + (ss - sc);
                        f4set(g1v, e, le * sxx);
                        f4set(g2v, e, le * szz);
                        float cross = 0.f;
                        if (XZ) {
                            const float nw = e == 0 ? sU.l : f4get(sU.s, e - 1), ne = e == 3 ? sU.r : f4get(sU.s, e + 1);
                            const float sw2 = e == 0 ? sD.l : f4get(sD.s, e - 1), se2 = e == 3 ? sD.r : f4get(sD.s, e + 1);
                            cross = (se2 - sw2) - (ne - nw);
                            f4set(g3v, e, le * cross);
                        }
                        if (f == 0) f4set(g6v, e, f4get(ls, e) * (f4get(ca, e) * sxx + f4get(cb, e) * szz + f4get(cc, e) * cross));
                    }
                }
                const int ro = z * ld + x;
                if (gsl != nullptr) {
                    // gradient partial sums of the block's shots in shared memory; cells the block does not own are never flushed
                    if (want_grad) {
                        float4* p4 = gsl + k * (FW / 4);
                        p4[0] = f4add(p4[0], g1v);
                        p4[GPL4] = f4add(p4[GPL4], g2v);
                        if (XZ) p4[2 * GPL4] = f4add(p4[2 * GPL4], g3v);
                        if (f == 0) p4[(XZ ? 3 : 2) * GPL4] = f4add(p4[(XZ ? 3 : 2) * GPL4], g6v);
                    }
                    if (clean) {
                        *reinterpret_cast<float4*>(l0 + ro) = out;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (x + e < g.nx && owns(z, x + e)) l0[ro + e] = f4get(out, e);
                            else if (x + e >= g.nx && x + e < ld) l0[ro + e] = 0.f;
                        }
                    }
                } else if (clean) {
                    // the whole tile is frame-free and inside the domain: 128-bit stores and gradient read-modify-writes
                    *reinterpret_cast<float4*>(l0 + ro) = out;
                    if (want_grad) {
                        auto rmw = [&](int slot, const float4& v) {
                            float4* p4 = reinterpret_cast<float4*>(gb + slot * plane + ro);
                            *p4 = f4add(*p4, v);
                        };
                        rmw(1, g1v);
                        rmw(2, g2v);
                        if (XZ) rmw(3, g3v);
                        if (f == 0) rmw(6, g6v);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool mine = x + e < g.nx && owns(z, x + e);
                        if (mine) {
                            l0[ro + e] = f4get(out, e);
                            if (want_grad) {
                                gb[plane + ro + e] += f4get(g1v, e);
                                gb[2 * plane + ro + e] += f4get(g2v, e);
                                if (XZ) gb[3 * plane + ro + e] += f4get(g3v, e);
                                if (f == 0) gb[6 * plane + ro + e] += f4get(g6v, e);
                            }
                        } else if (x + e >= g.nx && x + e < ld) {
                            l0[ro + e] = 0.f;
                        }
                    }
                }
                U = C; C = D;
                if (want_grad) { sU = sC; sC = sD; }
            }
        }
    }
}

// Born pairs, BOTH fields in one pass over the rows (default; ST_BORN_FUSED=0 keeps the field-after-field version above for
// A/B timing): one load of cxx, czz (cxz), m and of the scattered cotangent serves the two fields, the loads of a row are
// all independent (one memory round trip per row instead of one per field and gradient plane), and the gradient partial
// sums of the block's shots live in shared memory (planes cxx, czz, [cxz,] m; flushed once per `bchunk` shots).
#ifndef ST_BORN_FUSED
#define ST_BORN_FUSED 2                     // 0: field after field, global gradient RMW (round-2 first version); 1: fused rows for
#endif                                      // the VTI pair, field after field + shared-memory gradients for the TTI pair; 2: fused
                                            // rows for both (B200, 12 shots 600x1300: VTI 393 / 275 / 276 us, TTI 496 / 439 / 386 us)
template <int FL, bool SAFE, class Own>
__device__ __forceinline__ void adjoint_fast_rows_born2(const W2Args& a, const W2Geom& g, int b, int x0, int z0, int zn,
                                                        int lane, bool clean, bool want_grad, float4* gsl, Own owns) {
    constexpr bool XZ = (FL & ST_F_XZ) != 0;
    const int x = x0 + 4 * lane;
    const int ld = g.ld;
    const long long boff = (long long)b * a.fs;
    const float* __restrict__ l1b = a.lam1 + boff;             // cotangent of the background field
    const float* __restrict__ l1s = a.lam1 + a.cs + boff;      // ... of the scattered field
    const float* __restrict__ S0 = a.s1 + boff;
    const float* __restrict__ S1 = a.s1 + a.cs + boff;
    const float* __restrict__ cxx = a.coef[2];
    const float* __restrict__ czz = a.coef[3];
    const float* __restrict__ cxz = a.coef[4];
    const float* __restrict__ mm = a.coef[7];
    const int xh = lane == 0 ? x0 - 1 : x0 + FW;               // halo column (every lane issues the same predicated loads;
    const bool xhin = xh >= 0 && xh < g.nx;                    //  lanes 1..30 read lane 31's column and drop it)
    struct Row {
        float4 le0, le1, r0;                   // effective cotangents le0 = L1(0) + m L1(1), le1 = L1(1); raw L1(0)
        float4 a0, a1, b0, b1, c0, c1;         // cxx*le, czz*le, cxz*le
        float4 s0, s1;                         // S_i rows of the two fields
        float a0l, a0r, a1l, a1r, c0l, c0r, c1l, c1r, s0l, s0r, s1l, s1r;
    };
    auto edges = [&](const float4& v, float h, float& l, float& r) {
        l = __shfl_up_sync(0xffffffffu, v.w, 1);
        r = __shfl_down_sync(0xffffffffu, v.x, 1);
        l = lane == 0 ? h : l;
        r = lane == 31 ? h : r;
    };
    auto load_row = [&](int z) {
        Row q;
        q.r0 = ldrow_s<SAFE>(l1b, z, x, g);
        q.le1 = ldrow_s<SAFE>(l1s, z, x, g);
        q.le0 = f4fma(ldrow_s<SAFE>(mm, z, x, g), q.le1, q.r0);
        const float4 ca = ldrow_s<SAFE>(cxx, z, x, g);
        const float4 cb = ldrow_s<SAFE>(czz, z, x, g);
        const float4 cc = XZ ? ldrow_s<SAFE>(cxz, z, x, g) : f4zero();
        q.a0 = f4mul(ca, q.le0); q.a1 = f4mul(ca, q.le1);
        q.b0 = f4mul(cb, q.le0); q.b1 = f4mul(cb, q.le1);
        q.c0 = f4mul(cc, q.le0); q.c1 = f4mul(cc, q.le1);
        q.s0 = want_grad ? ldrow_s<SAFE>(S0, z, x, g) : f4zero();
        q.s1 = want_grad ? ldrow_s<SAFE>(S1, z, x, g) : f4zero();
        // halo column
        const bool in = SAFE || (xhin && z >= 0 && z < g.nz);
        const int o = z * ld + xh;
        const float h1 = in ? __ldg(l1s + o) : 0.f;
        const float h0 = in ? fmaf(__ldg(mm + o), h1, __ldg(l1b + o)) : 0.f;
        const float hca = in ? __ldg(cxx + o) : 0.f;
        const float hcc = (XZ && in) ? __ldg(cxz + o) : 0.f;
        const float hs0 = (want_grad && in) ? __ldg(S0 + o) : 0.f;
        const float hs1 = (want_grad && in) ? __ldg(S1 + o) : 0.f;
        edges(q.a0, hca * h0, q.a0l, q.a0r);
        edges(q.a1, hca * h1, q.a1l, q.a1r);
        edges(q.c0, hcc * h0, q.c0l, q.c0r);
        edges(q.c1, hcc * h1, q.c1l, q.c1r);
        edges(q.s0, hs0, q.s0l, q.s0r);
        edges(q.s1, hs1, q.s1l, q.s1r);
        return q;
    };
    Row U = load_row(z0 - 1), C = load_row(z0), D;
#pragma unroll
    for (int k = 0; k < FRZ; ++k) {
        const int z = z0 + k;
        if (z < zn) {
            D = load_row(z + 1);
            const float4 p20 = ldrow_s<SAFE>(a.lam2 + boff, z, x, g);
            const float4 p21 = ldrow_s<SAFE>(a.lam2 + a.cs + boff, z, x, g);
            float4 out0, out1, g1v = f4zero(), g2v = f4zero(), g3v = f4zero(), g6v = f4zero();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                auto field = [&](const float4& raw, const float4& p2, const float4& ca_, float al, float ar,
                                 const float4& ub, const float4& cb_, const float4& db,
                                 const float4& uc, float ucl, float ucr, const float4& dc, float dcl, float dcr) {
                    const float ac = f4get(ca_, e);
                    const float aw = e == 0 ? al : f4get(ca_, e - 1), ae = e == 3 ? ar : f4get(ca_, e + 1);
                    float acc = 2.f * f4get(raw, e) - f4get(p2, e);
                    acc += ((ae - ac) + (aw - ac));                                              // dxx of cxx*le
                    acc += ((f4get(ub, e) - f4get(cb_, e)) + (f4get(db, e) - f4get(cb_, e)));    // dzz of czz*le
                    if (XZ) {
                        const float uw = e == 0 ? ucl : f4get(uc, e - 1), ue = e == 3 ? ucr : f4get(uc, e + 1);
                        const float dw = e == 0 ? dcl : f4get(dc, e - 1), de = e == 3 ? dcr : f4get(dc, e + 1);
                        acc += (uw - ue) - (dw - de);
                    }
                    return acc;
                };
                f4set(out0, e, field(C.r0, p20, C.a0, C.a0l, C.a0r, U.b0, C.b0, D.b0, U.c0, U.c0l, U.c0r, D.c0, D.c0l, D.c0r));
                f4set(out1, e, field(C.le1, p21, C.a1, C.a1l, C.a1r, U.b1, C.b1, D.b1, U.c1, U.c1l, U.c1r, D.c1, D.c1l, D.c1r));
                if (want_grad) {
                    auto second = [&](const float4& su, const float4& sc_, const float4& sd, float scl, float scr,
                                      float sul, float sur, float sdl, float sdr, float& sxx, float& szz, float& cross) {
                        const float sc = f4get(sc_, e);
                        const float sw_ = e == 0 ? scl : f4get(sc_, e - 1), se_ = e == 3 ? scr : f4get(sc_, e + 1);
                        sxx = (se_ - sc) + (sw_ - sc);
                        szz = (f4get(su, e) - sc) + (f4get(sd, e) - sc);
                        cross = 0.f;
                        if (XZ) {
                            const float nw = e == 0 ? sul : f4get(su, e - 1), ne = e == 3 ? sur : f4get(su, e + 1);
                            const float sw2 = e == 0 ? sdl : f4get(sd, e - 1), se2 = e == 3 ? sdr : f4get(sd, e + 1);
                            cross = (se2 - sw2) - (ne - nw);
                        }
                    };
                    float sxx0, szz0, cr0, sxx1, szz1, cr1;
                    second(U.s0, C.s0, D.s0, C.s0l, C.s0r, U.s0l, U.s0r, D.s0l, D.s0r, sxx0, szz0, cr0);
                    second(U.s1, C.s1, D.s1, C.s1l, C.s1r, U.s1l, U.s1r, D.s1l, D.s1r, sxx1, szz1, cr1);
                    const float le0 = f4get(C.le0, e), le1 = f4get(C.le1, e);
                    f4set(g1v, e, fmaf(le0, sxx0, le1 * sxx1));
                    f4set(g2v, e, fmaf(le0, szz0, le1 * szz1));
                    if (XZ) f4set(g3v, e, fmaf(le0, cr0, le1 * cr1));
                    // g_m += L1(1) A[S_0] = (cxx L1(1)) dxx S_0 + (czz L1(1)) dzz S_0 + (cxz L1(1)) dxz S_0: the products are at hand
                    f4set(g6v, e, f4get(C.a1, e) * sxx0 + f4get(C.b1, e) * szz0 + (XZ ? f4get(C.c1, e) * cr0 : 0.f));
                }
            }
            const int ro = z * ld + x;
            float* o0 = a.lam0 + boff + ro;
            float* o1 = a.lam0 + a.cs + boff + ro;
            if (clean) {
                *reinterpret_cast<float4*>(o0) = out0;
                *reinterpret_cast<float4*>(o1) = out1;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (x + e < g.nx && owns(z, x + e)) { o0[e] = f4get(out0, e); o1[e] = f4get(out1, e); }
                    else if (x + e >= g.nx && x + e < ld) { o0[e] = 0.f; o1[e] = 0.f; }
                }
            }
            if (want_grad) {                   // cells this block does not own are never flushed (adjoint_fast_block)
                float4* p4 = gsl + k * (FW / 4);
                p4[0] = f4add(p4[0], g1v);
                p4[GPL4] = f4add(p4[GPL4], g2v);
                if (XZ) p4[2 * GPL4] = f4add(p4[2 * GPL4], g3v);
                p4[(XZ ? 3 : 2) * GPL4] = f4add(p4[(XZ ? 3 : 2) * GPL4], g6v);
            }
            U = C; C = D;
        }
    }
}

// The fast blocks keep the gradient partial sums of their shots in shared memory: one plane per coefficient gradient the
// flag set has (slots 1..6 of the gradient accumulator; slot 0, d/d r, belongs to the frame blocks), flushed once per chunk.
template <int FL>
__host__ __device__ constexpr bool adj_smem_grad() {
    return adj_iso_only<FL>() || ((FL & ST_F_BORN) ? ST_BORN_FUSED != 0 : ST_GEN_SMEM_GRAD != 0);
}
template <int FL>
__host__ __device__ constexpr int adj_smem_slot(int i) {       // i-th gradient slot in use
    int n = 0;
    for (int q = 1; q <= 6; ++q)
        if (grad_used<FL>(q)) { if (n == i) return q; ++n; }
    return 1;
}

template <int FL>
__device__ __forceinline__ void adjoint_fast_block(const W2Args& a, int bid, int nfx, int chunk, int tid,
                                                   float (*gsm)[FRZ][FW]) {
    constexpr bool HABC = (FL & ST_F_HABC) != 0;
    constexpr int NGS = adj_smem_planes<FL>();
    const W2Geom g = a.g;
    const int warp = tid >> 5, lane = tid & 31;
    const int fz = bid / nfx, fx = bid - fz * nfx;
    const int x0 = fx * FW, zb0 = fz * FH;
    const int z0 = zb0 + warp * FRZ;
    const int x = x0 + 4 * lane;
    const int band = g.bw + 1;               // cells this deep or deeper are untouched by the frame
    const bool want_grad = a.gacc != nullptr;
    const int zn = min(z0 + FRZ, g.nz);
    bool rows = z0 < g.nz;
    bool clean = x0 + FW <= g.nx;
    if (HABC && rows) {
        clean = clean && edge_depth(z0, x0, g) >= band && edge_depth(zn - 1, x0 + FW - 1, g) >= band &&
                edge_depth(z0, x0 + FW - 1, g) >= band && edge_depth(zn - 1, x0, g) >= band;
        // tile entirely inside the band: nothing to do here (the band blocks own it)
        if (zn <= (g.multiple ? 0 : band) || z0 >= g.nz - band || x0 + FW <= band || x0 >= g.nx - band) rows = false;
    }
    const bool safe = z0 >= 1 && z0 + FRZ + 1 <= g.nz && x0 >= 1 && x0 + FW + 1 <= g.nx;
    auto owns = [&](int z, int xx) { return !HABC || edge_depth(z, xx, g) >= band; };
    float4* gsl = reinterpret_cast<float4*>(&gsm[warp][0][4 * lane]);      // stride FW/4 float4 per row, GPL4 per plane
    if (want_grad && adj_smem_grad<FL>()) {
#pragma unroll
        for (int q = 0; q < NGS; ++q)
#pragma unroll
            for (int k = 0; k < FRZ; ++k) gsl[q * GPL4 + k * (FW / 4)] = f4zero();
    }
    const int b_lo = chunk * a.bchunk, b_hi = min(b_lo + a.bchunk, a.B);
    for (int b = b_lo; b < b_hi; ++b) {
        if (rows) {
            if constexpr (adj_iso_only<FL>()) {
                if (safe) adjoint_fast_rows<FL, true>(a, g, b, x0, z0, zn, lane, clean, want_grad, gsl, owns);
                else adjoint_fast_rows<FL, false>(a, g, b, x0, z0, zn, lane, clean, want_grad, gsl, owns);
            } else if constexpr ((FL & ST_F_BORN) != 0) {
                if constexpr (ST_BORN_FUSED == 2 || (ST_BORN_FUSED == 1 && !(FL & ST_F_XZ)))
                {
                    // frame-free tiles of an HABC grid lie at least bw + 1 cells inside the domain: unchecked loads
                    if (clean && safe) adjoint_fast_rows_born2<FL, true>(a, g, b, x0, z0, zn, lane, clean, want_grad, gsl, owns);
                    else adjoint_fast_rows_born2<FL, false>(a, g, b, x0, z0, zn, lane, clean, want_grad, gsl, owns);
                }
                else adjoint_fast_rows_born<FL>(a, g, b, chunk, x0, z0, zn, lane, clean, want_grad, owns, ST_BORN_FUSED ? gsl : nullptr);
            } else {
                // (an unchecked-load instantiation for the frame-free tiles was measured: two copies of the rows at 80 registers
                //  spill more and run 5-8 % slower; the Born pairs, at 128 registers, gain 2-3 % from theirs)
                adjoint_fast_rows_gen<FL, false>(a, g, b, chunk, x0, z0, zn, lane, clean, want_grad, owns, adj_smem_grad<FL>() ? gsl : nullptr);
            }
        }
        adjoint_tail<(FL & ST_F_BORN) ? 2 : 1>(a, b, zb0, zb0 + FH, x0, x0 + FW, tid, owns);
    }
    if (want_grad && adj_smem_grad<FL>() && rows && x < g.ld) {
        const long long plane = (long long)g.nz * g.ld;
#pragma unroll
        for (int q = 0; q < NGS; ++q) {
            float* gb = a.gacc + ((long long)chunk * 7 + adj_smem_slot<FL>(q)) * plane;
#pragma unroll
            for (int k = 0; k < FRZ; ++k) {
                const int z = z0 + k;
                if (z < zn) {
                    float* o = gb + (z * g.ld + x);
                    const float4 acc = gsl[q * GPL4 + k * (FW / 4)];
                    if (clean) {
                        float4 v = *reinterpret_cast<float4*>(o);
                        v.x += acc.x; v.y += acc.y; v.z += acc.z; v.w += acc.w;
                        *reinterpret_cast<float4*>(o) = v;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (x + e < g.nx && owns(z, x + e)) o[e] += f4get(acc, e);
                    }
                }
            }
        }
    }
}

// general cell-by-cell adjoint of one TX x TZ tile; `band` < 0: every cell, else only the
// cells closer than `band` to an absorbing edge
// Shots b_lo..b_hi-1 are processed in turn; gradient contributions go to plane `gplane`.
template <int FL, int UNR = 1>
__device__ __forceinline__ void adjoint_general_block(const W2Args& a, int tz, int tx, int b_lo, int b_hi, int gplane,
                                                      int tid, int band, float (*sl)[SH][SW], float (*ss)[SH][SW], int xoff = 0) {
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    const W2Geom g = a.g;
    const int x0 = tx * TX + xoff, z0 = tz * TZ;
    const int x = x0 + (tid & (NTX - 1)), ty = tid / NTX;
    const bool want_grad = a.gacc != nullptr;
    const long long plane = (long long)g.nz * g.ld;
    auto owns = [&](int z, int xx) { return band < 0 || edge_depth(z, xx, g) < band; };
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
#pragma unroll UNR
            for (int k = 0; k < RPT; ++k) {
                const int z = z0 + ty + k * NTY;
                if (z >= g.nz) break;
                if (!owns(z, x)) continue;
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
                auto CK = [&](int k, int zz, int xx) -> float { return __ldg(a.coef[k] + (zz * g.ld + xx)); };
                float out[2], gr[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                w2_adjoint_cell<FL>(z, x, g, a.dt, L1, L2, S1, S2, CF, CK, out, gr, want_grad);
#pragma unroll
                for (int f = 0; f < NF; ++f) a.lam0[f * a.cs + boff + idx] = out[f];
                if (want_grad) {
                    float* gb = a.gacc + (long long)gplane * 7 * plane + idx;
#pragma unroll
                    for (int q = 0; q < 7; ++q)
                        if (grad_used<FL>(q)) gb[q * plane] += gr[q];
                }
            }
        }
        adjoint_tail<NF>(a, b, z0, z0 + TZ, x0, x0 + TX, tid, owns);
    }
}

// resident blocks per SM the register adjoint is compiled for (the Born pairs carry two fields: 128 registers)
template <int FL>
__host__ __device__ constexpr int adj_minb() { return (FL & ST_F_BORN) ? ST_ADJ_MINB_BORN : (FL & ST_F_XZ) ? ST_ADJ_MINB_XZ : ST_ADJ_MINB; }
#ifndef ST_ADJ_SPLIT
#define ST_ADJ_SPLIT 0                      // 1: frame blocks and fast blocks of the register adjoint in two launches (tuning;
                                            // measured on B200 with 2 / 3 / 4 resident fast blocks per SM: no gain -- the fast
                                            // rows of the Born / TTI equations need ~100 registers themselves)
#endif
#ifndef ST_ADJ_MINB_FAST
#define ST_ADJ_MINB_FAST 3                  // resident blocks per SM the fast-only launch is compiled for
#endif
template <int FL>
__host__ __device__ constexpr bool adj_split() { return ST_ADJ_SPLIT != 0 && (FL & (ST_F_BORN | ST_F_XZ | ST_F_G1)) != 0; }

// PART: 0 = frame blocks and fast blocks in one launch; 1 = frame blocks only, 2 = fast blocks only (two launches: the
// register allocation of a kernel is the maximum over its code paths, and the tap / generic frame code of the Born and
// TTI equations needs far more registers than their fast rows -- see st_w2_launch_adj)
template <int FL, int PART = 0>
__global__ void __launch_bounds__(NT, PART == 2 ? ST_ADJ_MINB_FAST : adj_minb<FL>()) wave2d_adjoint_kernel(const W2Args a, int nfx, int nfast, BandTiles bt) {
    st_pdl_launch_dependents();
    st_pdl_wait();
    constexpr int NF = (FL & ST_F_BORN) ? 2 : 1;
    constexpr bool NEED_GEN = PART != 2 && (!adj_fast<FL>() || (FL & ST_F_HABC));
    constexpr int GEN_FLOATS = NEED_GEN ? 2 * NF * SH * SW : 1;
    constexpr int FAST_FLOATS = adj_smem_grad<FL>() ? adj_smem_planes<FL>() * NWARP * FRZ * FW : 1;
    __shared__ __align__(16) float smem[GEN_FLOATS > FAST_FLOATS ? GEN_FLOATS : FAST_FLOATS];
    const int tid = threadIdx.x;
    // grid.x = [band blocks: one per (tile, shot)] ++ [fast blocks: one per (fast tile, shot chunk)]
    // for the fast-path equations, else one general block per (tile, shot chunk).
    if constexpr (adj_fast<FL>()) {
        const bool tapped = st_flags_tapped(FL) && a.taps != nullptr;
        const int ngrp = band_groups(a.B), gsh = band_group_shots(a.B);
        const int nstrip = (tapped && st_flags_stripped_adj(FL)) ? strip_blocks(strip_geom(a.g, a.g.bw + 1)) : 0;
        const int nband = (FL & ST_F_HABC) ? (bt.count + nstrip) * (tapped ? ngrp : a.B) : 0;
        const int bid = PART == 2 ? blockIdx.x + nband : blockIdx.x;     // a fast-only launch enumerates the fast blocks from 0
        if ((ST_DBG_SKIP & 1) && bid < nband) return;          // tuning builds only: frame blocks off
        if ((ST_DBG_SKIP & 2) && bid >= nband) return;         //                     fast blocks off
        if (bid >= nband) {
            if constexpr (PART != 1) {
                const int q = bid - nband;
                adjoint_fast_block<FL>(a, q % nfast, nfx, q / nfast, tid, reinterpret_cast<float (*)[FRZ][FW]>(smem));
            }
        } else if constexpr (PART == 2) {
            return;
        } else if (tapped) {
            const int per = bt.count + nstrip, grp = bid / per, k = bid - grp * per;     // gradient plane = group id (< B planes exist)
            const int b_lo = grp * gsh, b_hi = min(b_lo + gsh, a.B);
            if constexpr (st_flags_stripped_adj(FL)) {
                if (k < nstrip) adjoint_strip_block<FL>(a, k, b_lo, b_hi, grp, tid);
                else adjoint_band_block<FL>(a, k - nstrip, b_lo, b_hi, grp, tid);
            } else {
                adjoint_band_block<FL>(a, k, b_lo, b_hi, grp, tid);
            }
        } else {
            if constexpr (NEED_GEN) {
                int tz, tx;
                const int b = bid / bt.count;
                band_tile_decode(bt, bid - b * bt.count, tz, tx);
                // band gradients of shot b accumulate in gradient plane b (see nchunk in the launcher)
                adjoint_general_block<FL>(a, tz, tx, b, b + 1, b, tid, a.g.bw + 1, reinterpret_cast<float (*)[SH][SW]>(smem),
                                          reinterpret_cast<float (*)[SH][SW]>(smem + NF * SH * SW));
            }
        }
    } else {
        const int bid = blockIdx.x;
        const int ntile = bt.nxt * bt.nzt;
        const int chunk = bid / ntile, t = bid - chunk * ntile;
        const int tz = t / bt.nxt, tx = t - tz * bt.nxt;
        const int b_lo = chunk * a.bchunk, b_hi = min(b_lo + a.bchunk, a.B);
        adjoint_general_block<FL>(a, tz, tx, b_lo, b_hi, chunk, tid, -1, reinterpret_cast<float (*)[SH][SW]>(smem),
                                  reinterpret_cast<float (*)[SH][SW]>(smem + NF * SH * SW));
    }
}

// TMA block of the adjoint (tile kinds as in forward_tma_block).  Interior tiles: the arithmetic of
// adjoint_fast_rows.  Frame tiles (straight top / bottom side, inward normal n = +z / -z), per cell p:
//   Lam_i(p) = lap(c pre L1)(p) + [2 pre + b (2 - lam - mu)](p) L1(p) + [b (lam + 2 mu)](p-n) L1(p-n) - [b mu](p-2n) L1(p-2n)
//              + [-pre + b (lam - 1)](p) L2(p) - [b lam](p-n) L2(p-n),      pre = 1 - b, lam = 2 r, mu = r^2
//   g_ciso(p) += pre L1 lap(S_i)(p)
//   g_r(p)    += b L1(p) [(-2 - 2r) S_i(p) + (2 + 4r) S_i(p+n) - 2r S_i(p+2n) + 2 S_{i-1}(p) - 2 S_{i-1}(p+n)]
// (the transposed st_wave2d_band.cu taps of a straight side; b == 0 beyond the frame makes the rows at
// depth bw, bw+1 and the frame-free rows of the tile come out right with the same expression).
// Coefficient rows and gradient partial sums stay in registers across the block's shots.
template <int FL, int KIND>
__device__ __forceinline__ void adjoint_tma_tile(const W2Args& a, const W2Tma& tm, const TmaChunk& q, int tid,
                                                 unsigned char* dsm, uint64_t* bars) {
    constexpr bool PML = (FL & ST_F_PML) != 0;
    constexpr bool ZDIR = KIND == 1 || KIND == -1, XDIR = KIND == 2 || KIND == -2;
    constexpr int N = KIND > 0 ? 1 : -1;                    // inward normal along z (ZDIR) or x (XDIR)
    constexpr bool DEEP = KIND == 0 && (FL & ST_F_HABC) != 0;          // frame-free tiles of HABC: smaller stages, one more of them
    constexpr int NS = ST_TMA_ADJ_STAGES + (DEEP ? 1 : 0), STAGE = DEEP ? TMA_ADJ_STAGE0 : tma_adj_stage<FL>();
    constexpr int R0 = DEEP ? TMA_H1_BYTES : tma_r0<FL>();
    constexpr int HOFF = ZDIR ? 2 : 1;                      // rows above z0 in the Lam1 / S_i boxes
    const W2Geom& g = a.g;
    const int ld = g.ld;
    const int b_lo = q.b_lo, nsh = q.nsh;
    const int nitem = q.ntile * nsh;                        // item j = (tile j / nsh, shot j % nsh)
    const int warp = tid >> 5, lane = tid & 31;
    const bool want_grad = a.gacc != nullptr;
    const bool masked = q.mx0 > 0 || q.mx1 < (1 << 30);
    const bool xrag = q.x0 + q.ntile * q.dx + FW > g.nx;    // the tile column that crosses nx
    auto issue = [&](int j) {
        const int stg = j % NS, ti = j / nsh, s = j - ti * nsh;
        const int z0 = q.z0 + ti * q.dz, x0 = q.x0 + ti * q.dx;
        unsigned char* dst = dsm + stg * STAGE;
        if (KIND == 0) {
            st_mbar_expect_tx(&bars[stg], 2 * H1R * HC * 4 + TMA_CORE_BYTES);
            st_tma_load_3d(dst, &tm.l_h1, &bars[stg], x0 - XO, z0 - 1, tm.pl_l1 + b_lo + s);
            st_tma_load_3d(dst + R0, &tm.u_h1, &bars[stg], x0 - XO, z0 - 1, tm.pl_s1 + b_lo + s);
            st_tma_load_3d(dst + 2 * R0, &tm.l_core, &bars[stg], x0, z0, tm.pl_l2 + b_lo + s);
        } else {
            st_mbar_expect_tx(&bars[stg], 2 * (ZDIR ? H2R : H1R) * HC * 4 + 2 * H1R * HC * 4);
            st_tma_load_3d(dst, ZDIR ? &tm.l_h2 : &tm.l_h1, &bars[stg], x0 - XO, z0 - HOFF, tm.pl_l1 + b_lo + s);
            st_tma_load_3d(dst + R0, ZDIR ? &tm.u_h2 : &tm.u_h1, &bars[stg], x0 - XO, z0 - HOFF, tm.pl_s1 + b_lo + s);
            st_tma_load_3d(dst + 2 * R0, &tm.l_h1, &bars[stg], x0 - XO, z0 - 1, tm.pl_l2 + b_lo + s);
            st_tma_load_3d(dst + 2 * R0 + TMA_H1_BYTES, &tm.u_h1, &bars[stg], x0 - XO, z0 - 1, tm.pl_s2 + b_lo + s);
        }
    };
    if (tid == 0)
        for (int j = 0; j < NS && j < nitem; ++j) issue(j);
    int z0 = q.z0, x0 = q.x0, zr = 0, x = 0;
    auto inz = [&](int z) { return z >= 0 && z < g.nz; };
    auto ldc4 = [&](int q, int z) { return (inz(z) && x < ld) ? __ldg(reinterpret_cast<const float4*>(a.coef[q] + (z * ld + x))) : f4zero(); };
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
    // cp = ciso (1 - b) on rows zr-1 .. zr+2 and on the halo column of the edge lanes (rows zr, zr+1)
    float4 cp[4], al[2], gc[2], gr[2];
    float ch[2] = {0.f, 0.f};
    float4 bb[4], rr[4];
    constexpr int OWN0 = (ZDIR && N > 0) ? 2 : 0;
    for (int j = 0, sh = 0; j < nitem; ++j) {
      if (sh == 0) {                                        // new tile: its coefficient rows, cleared gradient sums
        zr = z0 + 2 * warp;
        x = x0 + 4 * lane;
        const int xh = lane == 0 ? x0 - 1 : x0 + FW;
        ch[0] = ch[1] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cp[k] = ldc4(2, zr - 1 + k);
            if (KIND) cp[k] = f4mul(cp[k], f4sub(one4, ldc4(1, zr - 1 + k)));
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            al[k] = PML ? ldc4(3, zr + k) : f4zero();
            gc[k] = gr[k] = f4zero();
            if ((lane == 0 || lane == 31) && xh >= 0 && xh < g.nx && inz(zr + k)) {
                ch[k] = __ldg(a.coef[2] + ((zr + k) * ld + xh));
                if (KIND) ch[k] *= 1.f - __ldg(a.coef[1] + ((zr + k) * ld + xh));
            }
        }
    // frame tiles: b and r.  ZDIR: the four rows {own rows, the two rows outward of them}
    //   (N = +1: rows zr-2 .. zr+1, own rows at index 2, 3;  N = -1: rows zr .. zr+3, own rows at index 0, 1);
    //   XDIR: the two own rows (index 0, 1).
        if (KIND) {
#pragma unroll
            for (int i = 0; i < (ZDIR ? 4 : 2); ++i) {
                const int z = zr - OWN0 + i;
                bb[i] = ldc4(1, z);
                rr[i] = ldc4(0, z);
            }
        }
      }
      {
        const int stg = j % NS, b = b_lo + sh;
        st_mbar_wait(&bars[stg], (j / NS) & 1);
        const float* l1 = reinterpret_cast<const float*>(dsm + stg * STAGE);
        const float* S = reinterpret_cast<const float*>(dsm + stg * STAGE + R0);
        // Lam2 / S_{i-1}: pointer to (z0, x0), row pitch
        const float* l2 = reinterpret_cast<const float*>(dsm + stg * STAGE + 2 * R0) + (KIND ? HC + XO : 0);
        const float* S2 = reinterpret_cast<const float*>(dsm + stg * STAGE + 2 * R0 + TMA_H1_BYTES) + HC + XO;
        constexpr int P2 = KIND ? HC : TC;
        float* out = a.lam0 + (long long)b * a.fs + (zr * ld + x);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int ro = (2 * warp + k + HOFF) * HC + XO + 4 * lane;                  // box offset of row z = zr + k, this lane
            const int hrow = (2 * warp + k + HOFF) * HC;
            const float* l2r = l2 + (2 * warp + k) * P2;
            const int z = zr + k;
            constexpr int IO = OWN0;                                                   // index of own row k is IO + k
            // ---- phase A: the cotangent
            {
                const float4 lC = *reinterpret_cast<const float4*>(l1 + ro);
                const float4 lU = *reinterpret_cast<const float4*>(l1 + ro - HC), lD = *reinterpret_cast<const float4*>(l1 + ro + HC);
                const float4 p2 = *reinterpret_cast<const float4*>(l2r + 4 * lane);
                const float4 wU = f4mul(cp[k], lU), wC = f4mul(cp[k + 1], lC), wD = f4mul(cp[k + 2], lD);
                float wl = __shfl_up_sync(0xffffffffu, wC.w, 1), wr = __shfl_down_sync(0xffffffffu, wC.x, 1);
                {   // halo columns: broadcast shared loads + selects
                    const float hl = l1[hrow + XO - 1], hr = l1[hrow + XO + TC];
                    wl = lane == 0 ? ch[k] * hl : wl;
                    wr = lane == 31 ? ch[k] * hr : wr;
                }
                float4 o4;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float c = f4get(wC, e);
                    const float w = e == 0 ? wl : f4get(wC, e - 1), ea = e == 3 ? wr : f4get(wC, e + 1);
                    const float lapw = ((f4get(wU, e) - c) + (f4get(wD, e) - c)) + ((ea - c) + (w - c));
                    if (KIND == 0) {
                        const float alpha = PML ? f4get(al[k], e) : 1.f;
                        f4set(o4, e, (!PML || x + e < g.nx) ? (1.f + alpha) * f4get(lC, e) + lapw - alpha * f4get(p2, e) : 0.f);
                    } else {
                        f4set(o4, e, lapw);
                    }
                }
                if (KIND) {
                    const float4 bz = bb[IO + k], rz = rr[IO + k];
                    // one-way terms of the cells outward of p:  X1 = [b (lam + 2 mu) L1](p-n),  X2 = [b mu L1](p-2n),
                    // X3 = [b lam L2](p-n); X2 only where the j+2 tap stays inside the strip (depth(p) <= bw)
                    float4 X1, X2, X3;
                    if (ZDIR) {
                        const float4 b1 = bb[IO + k - N], r1 = rr[IO + k - N];
                        const float4 b2 = bb[(IO + k - 2 * N) & 3], r2 = rr[(IO + k - 2 * N) & 3];
                        const int depth = N > 0 ? z : g.nz - 1 - z;
                        const float4 lO1 = N > 0 ? lU : lD;
                        const float4 lO2 = *reinterpret_cast<const float4*>(l1 + ro - 2 * N * HC);
                        const float4 p2O = *reinterpret_cast<const float4*>(l2r - N * P2 + 4 * lane);
                        const float far = depth <= g.bw ? 1.f : 0.f;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float b1e = f4get(b1, e), r1e = f4get(r1, e), r2e = f4get(r2, e);
                            f4set(X1, e, b1e * (2.f * r1e + 2.f * r1e * r1e) * f4get(lO1, e));
                            f4set(X2, e, far * f4get(b2, e) * (r2e * r2e) * f4get(lO2, e));
                            f4set(X3, e, b1e * (2.f * r1e) * f4get(p2O, e));
                        }
                    } else {
                        float4 T1, T2, T3;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float be = f4get(bz, e), re = f4get(rz, e);
                            f4set(T1, e, be * (2.f * re + 2.f * re * re) * f4get(lC, e));
                            f4set(T2, e, be * (re * re) * f4get(lC, e));
                            f4set(T3, e, be * (2.f * re) * f4get(p2, e));
                        }
                        // the products of the cells beyond the tile's outward edge are zero (outside the grid)
                        if (N > 0) {
                            float t1 = __shfl_up_sync(0xffffffffu, T1.w, 1), t3 = __shfl_up_sync(0xffffffffu, T3.w, 1);
                            float t2a = __shfl_up_sync(0xffffffffu, T2.z, 1), t2b = __shfl_up_sync(0xffffffffu, T2.w, 1);
                            if (lane == 0) t1 = t3 = t2a = t2b = 0.f;
                            X1 = f4shl(T1, t1);
                            X3 = f4shl(T3, t3);
                            X2 = make_float4(t2a, t2b, T2.x, T2.y);
                        } else {
                            float t1 = __shfl_down_sync(0xffffffffu, T1.x, 1), t3 = __shfl_down_sync(0xffffffffu, T3.x, 1);
                            float t2a = __shfl_down_sync(0xffffffffu, T2.x, 1), t2b = __shfl_down_sync(0xffffffffu, T2.y, 1);
                            if (lane == 31) t1 = t3 = t2a = t2b = 0.f;
                            X1 = f4shr(T1, t1);
                            X3 = f4shr(T3, t3);
                            X2 = make_float4(T2.z, T2.w, t2a, t2b);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int depth = N > 0 ? x + e : g.nx - 1 - (x + e);
                            if (depth > g.bw) f4set(X2, e, 0.f);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float bze = f4get(bz, e), rze = f4get(rz, e), pre = 1.f - bze;
                        float v = f4get(o4, e);
                        v += (2.f * pre + bze * (2.f - 2.f * rze - rze * rze)) * f4get(lC, e);
                        v += f4get(X1, e);
                        v -= f4get(X2, e);
                        v += (bze * (2.f * rze - 1.f) - pre) * f4get(p2, e);
                        v -= f4get(X3, e);
                        f4set(o4, e, (!xrag || x + e < g.nx) ? v : 0.f);                 // pitch padding stays zero
                    }
                }
                if (z < g.nz && x < ld) {
                    if (!masked) {
                        *reinterpret_cast<float4*>(out + k * ld) = o4;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (x + e >= q.mx0 && x + e < q.mx1) out[k * ld + e] = f4get(o4, e);
                    }
                }
            }
            // ---- phase B: imaging condition
            if (want_grad) {
                const float4 lC = *reinterpret_cast<const float4*>(l1 + ro);
                const float4 sC = *reinterpret_cast<const float4*>(S + ro);
                const float4 sU = *reinterpret_cast<const float4*>(S + ro - HC), sD = *reinterpret_cast<const float4*>(S + ro + HC);
                float sl = __shfl_up_sync(0xffffffffu, sC.w, 1), sr = __shfl_down_sync(0xffffffffu, sC.x, 1);
                {
                    const float hl = S[hrow + XO - 1], hr = S[hrow + XO + TC];
                    sl = lane == 0 ? hl : sl;
                    sr = lane == 31 ? hr : sr;
                }
                float4 gq;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float sc = f4get(sC, e);
                    const float sw_ = e == 0 ? sl : f4get(sC, e - 1), se_ = e == 3 ? sr : f4get(sC, e + 1);
                    const float laps = ((f4get(sU, e) - sc) + (f4get(sD, e) - sc)) + ((se_ - sc) + (sw_ - sc));
                    float pl = f4get(lC, e);
                    if (KIND) pl *= 1.f - f4get(bb[IO + k], e);
                    f4set(gq, e, pl * laps);
                }
                gc[k] = f4add(gc[k], gq);
                if (KIND) {
                    const float4 bz = bb[IO + k], rz = rr[IO + k];
                    const float* s2r = S2 + (2 * warp + k) * HC;
                    const float4 q0 = *reinterpret_cast<const float4*>(s2r + 4 * lane);
                    float4 sI1, sI2, q1, far;                 // S_i one / two cells inward, S_{i-1} one cell inward, j+2 <= bw
                    if (ZDIR) {
                        const int depth = N > 0 ? z : g.nz - 1 - z;
                        const float f = depth + 2 <= g.bw ? 1.f : 0.f;
                        far = make_float4(f, f, f, f);
                        sI1 = N > 0 ? sD : sU;
                        sI2 = *reinterpret_cast<const float4*>(S + ro + 2 * N * HC);
                        q1 = *reinterpret_cast<const float4*>(s2r + N * HC + XO * lane);
                    } else if (N > 0) {
                        float sr2 = __shfl_down_sync(0xffffffffu, sC.y, 1), qr = __shfl_down_sync(0xffffffffu, q0.x, 1);
                        if (lane == 31) { sr2 = S[hrow + XO + 1 + TC]; qr = s2r[TC]; }
                        sI1 = f4shr(sC, sr);
                        sI2 = make_float4(sC.z, sC.w, sr, sr2);
                        q1 = f4shr(q0, qr);
                    } else {
                        float sl2 = __shfl_up_sync(0xffffffffu, sC.z, 1), ql = __shfl_up_sync(0xffffffffu, q0.w, 1);
                        if (lane == 0) { sl2 = S[hrow + XO - 2]; ql = s2r[-1]; }
                        sI1 = f4shl(sC, sl);
                        sI2 = make_float4(sl2, sl, sC.x, sC.y);
                        q1 = f4shl(q0, ql);
                    }
                    if (XDIR) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int depth = N > 0 ? x + e : g.nx - 1 - (x + e);
                            f4set(far, e, depth + 2 <= g.bw ? 1.f : 0.f);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float rze = f4get(rz, e);
                        const float tt = (-2.f - 2.f * rze) * f4get(sC, e) + (2.f + 4.f * rze) * f4get(sI1, e)
                                       - f4get(far, e) * 2.f * rze * f4get(sI2, e) + 2.f * f4get(q0, e) - 2.f * f4get(q1, e);
                        f4set(gr[k], e, f4get(gr[k], e) + f4get(bz, e) * f4get(lC, e) * tt);
                    }
                }
            }
        }
        adjoint_tail<1>(a, b, z0, z0 + TR, x0, x0 + FW, tid, [&](int, int xx) { return xx >= q.mx0 && xx < q.mx1; });
        __syncthreads();                                   // every warp is done with this stage
        if (tid == 0 && j + NS < nitem) issue(j + NS);
      }
      if (++sh < nsh) continue;
      // tile done: flush its gradient sums, move on
      sh = 0;
      if (want_grad && x < ld) {
        float* gb = a.gacc + (long long)q.plane * 7 * ((long long)g.nz * ld) + (zr * ld + x);
        const long long plane = (long long)g.nz * ld;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (zr + k >= g.nz) continue;
            if (masked) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (x + e < q.mx0 || x + e >= q.mx1) continue;
                    gb[plane + k * ld + e] += f4get(gc[k], e);
                    if (KIND) gb[k * ld + e] += f4get(gr[k], e);
                }
                continue;
            }
            float4 v = *reinterpret_cast<float4*>(gb + plane + k * ld);
            *reinterpret_cast<float4*>(gb + plane + k * ld) = f4add(v, gc[k]);            // slot 1: d/d ciso
            if (KIND) {
                float4 u = *reinterpret_cast<float4*>(gb + k * ld);
                *reinterpret_cast<float4*>(gb + k * ld) = f4add(u, gr[k]);                // slot 0: d/d r
            }
        }
      }
      z0 += q.dz;
      x0 += q.dx;
    }
}

template <int FL>
__device__ __forceinline__ void adjoint_tma_block(const W2Args& a, const W2Tma& tm, int nfx, int bid, int tid, unsigned char* dsm) {
    constexpr int NS = ST_TMA_ADJ_STAGES + 1;
    __shared__ __align__(8) uint64_t bars[NS];
    const TmaChunk q = tma_block_decode(tm, a.g, (FL & ST_F_HABC) != 0, nfx, a.B, bid);
    if (q.ntile == 0) return;
    const int kind = q.kind;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) st_mbar_init(&bars[s], 1);
        st_mbar_init_fence();
    }
    __syncthreads();
    st_pdl_wait();                                          // the previous step's cotangents are complete from here on
    if constexpr ((FL & ST_F_HABC) != 0) {
        if (kind == 1) { adjoint_tma_tile<FL, 1>(a, tm, q, tid, dsm, bars); return; }
        if (kind == -1) { adjoint_tma_tile<FL, -1>(a, tm, q, tid, dsm, bars); return; }
        if (kind == 2) { adjoint_tma_tile<FL, 2>(a, tm, q, tid, dsm, bars); return; }
        if (kind == -2) { adjoint_tma_tile<FL, -2>(a, tm, q, tid, dsm, bars); return; }
    }
    adjoint_tma_tile<FL, 0>(a, tm, q, tid, dsm, bars);
}

// TMA adjoint kernel: grid.x = [corner tiles x shots (generic per-cell code; gradient plane = shot)] ++ [TMA chunk blocks]
template <int FL>
__global__ void __launch_bounds__(NT, tma_adj_minb<FL>()) wave2d_adjoint_tma_kernel(const W2Args a, int nfx, const __grid_constant__ W2Tma tm) {
    extern __shared__ __align__(128) unsigned char dsm[];
    const int bid = blockIdx.x, tid = threadIdx.x;
    const CornerTiles ct = corner_tiles(tm, a.g);
    const bool ctap = st_flags_tapped(FL) && a.taps != nullptr;
    const int ncorner = corner_block_count(ct, a.B, ctap);
#ifdef ST_DBG_TIMELINE
    const unsigned long long t0 = dbg_now();
#endif
    st_pdl_launch_dependents();
    if (bid < ncorner) {
        st_pdl_wait();
        if (ST_DBG_SKIP & 32) return;
        if constexpr ((FL & ST_F_HABC) != 0) {
            int tz, tx;
            if (ctap) {
                // gradient plane = shot-group id (fewer than B planes; the TMA blocks own other cells of it)
                const int per = ct.count * CORNER_SUB, grp = bid / per, r = bid - grp * per;
                corner_tile_decode(ct, tm, a.g, r / CORNER_SUB, tz, tx);
                const RectMap map{a.g, tz * TZ + (r % CORNER_SUB) * (NT / TX), tx, TX};
                adjoint_tap_block<FL>(a, map, grp * band_group_shots(a.B), min((grp + 1) * band_group_shots(a.B), a.B), grp, tid);
            } else {
                const int b = bid / ct.count;
                corner_tile_decode(ct, tm, a.g, bid - b * ct.count, tz, tx);
                float* smem = reinterpret_cast<float*>(dsm);
                adjoint_general_block<FL, ST_CORNER_UNROLL>(a, tz, 0, b, b + 1, b, tid, -1, reinterpret_cast<float (*)[SH][SW]>(smem),
                                          reinterpret_cast<float (*)[SH][SW]>(smem + SH * SW), tx);
            }
        }
    } else {
        if (ST_DBG_SKIP & 8) return;
        adjoint_tma_block<FL>(a, tm, nfx, bid - ncorner, tid, dsm);
    }
#ifdef ST_DBG_TIMELINE
    if (tid == 0 && bid < 8192) { g_dbg_tl[4 * bid] = t0; g_dbg_tl[4 * bid + 1] = dbg_now(); g_dbg_tl[4 * bid + 2] = dbg_smid(); g_dbg_tl[4 * bid + 3] = bid < ncorner ? 2 : 1; }
#endif
}

}  // namespace

template <int FL>
int st_w2_launch_fwd(const W2Args& a, const W2Tma& tm, cudaStream_t st);
template <int FL>
int st_w2_launch_adj(const W2Args& a, const W2Tma& tm, cudaStream_t st);

#ifndef ST_W2_DISPATCH_ONLY
// TMA kernels are launched with programmatic stream serialization (st_tma.cuh: st_pdl_*), so the prologue of step
// i+1 overlaps the tail of step i; SEISTORCH_B200_PDL=0 falls back to ordinary launches.
template <class K>
static int tma_launch(K kernel, dim3 grid, int smem, cudaStream_t st, const W2Args& a, int nfx, const W2Tma& tm) {
    static const bool pdl = [] { const char* e = getenv("SEISTORCH_B200_PDL"); return !(e && atoi(e) == 0); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, a, nfx, tm) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

// register kernels: launched with programmatic stream serialization too (st_common.cuh: st_pdl_launch).  Every block
// executes launch_dependents + wait first thing, so the launch processing and block scheduling of step i+1 overlap step i
// -- what matters on small grids, where a step is a few microseconds and the GPU is mostly empty (measured on the
// 250 x 400 acoustic grid: 6.5 -> 4.7 us per forward step, 7.6 -> 6.2 us per adjoint step).
template <class K, class... Args>
static int pdl_launch(K kernel, dim3 grid, cudaStream_t st, Args... args) {
    return st_pdl_launch(kernel, grid, dim3(NT), 0, st, args...) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

template <int FL>
int st_w2_launch_fwd(const W2Args& a, const W2Tma& tm, cudaStream_t st) {
    const int nfx = (a.g.nx + FW - 1) / FW, nfz = (a.g.nz + FH - 1) / FH;
    if constexpr (tma_ok<FL>()) {
        if (tm.enabled) {
            if (st_set_max_smem<wave2d_forward_tma_kernel<FL>>(tma_fwd_smem<FL>()) != cudaSuccess) return ST_ERR_CUDA;
            const bool ctap = ST_CORNER_TAP_FWD && st_flags_tapped(FL) && a.taps != nullptr;
            dim3 grid((unsigned)((long long)corner_block_count(corner_tiles(tm, a.g), a.B, ctap) + tma_blocks(tm, a.B)));
            return tma_launch(wave2d_forward_tma_kernel<FL>, grid, tma_fwd_smem<FL>(), st, a, nfx, tm);
        }
    }
    const int nfast = nfx * nfz;
    BandTiles bt = band_tiles(a.g, a.g.bw);
    if (!(FL & ST_F_HABC)) bt.count = 0;
    const bool tapped = st_flags_tapped(FL) && a.taps != nullptr;
    if (tapped) bt.count = (st_band_cells(a.g, a.g.bw).total + NT - 1) / NT;
    const int nstrip = (tapped && st_flags_stripped_fwd(FL)) ? strip_blocks(strip_geom(a.g, a.g.bw)) : 0;
    dim3 grid((unsigned)((long long)nfast * a.B + (long long)(bt.count + nstrip) * (tapped ? band_groups(a.B) : a.B)));
    return pdl_launch(wave2d_forward_kernel<FL>, grid, st, a, nfx, nfast, bt);
}
template <int FL>
int st_w2_launch_adj(const W2Args& a, const W2Tma& tm, cudaStream_t st) {
    const int nchunk = (a.B + a.bchunk - 1) / a.bchunk;
    const int nfx = (a.g.nx + FW - 1) / FW, nfz = (a.g.nz + FH - 1) / FH;
    if constexpr (tma_ok<FL>()) {
        if (tm.enabled) {
            if (st_set_max_smem<wave2d_adjoint_tma_kernel<FL>>(tma_adj_smem<FL>()) != cudaSuccess) return ST_ERR_CUDA;
            const bool ctap = st_flags_tapped(FL) && a.taps != nullptr;
            dim3 grid((unsigned)((long long)corner_block_count(corner_tiles(tm, a.g), a.B, ctap) + tma_blocks(tm, a.B)));
            return tma_launch(wave2d_adjoint_tma_kernel<FL>, grid, tma_adj_smem<FL>(), st, a, nfx, tm);
        }
    }
    const int nfast = nfx * nfz;
    BandTiles bt = band_tiles(a.g, a.g.bw + 1);
    const bool tapped = st_flags_tapped(FL) && a.taps != nullptr;
    if (tapped) bt.count = (st_band_cells(a.g, a.g.bw + 1).total + NT - 1) / NT;
    long long nblocks;
    const int nstrip = (tapped && st_flags_stripped_adj(FL)) ? strip_blocks(strip_geom(a.g, a.g.bw + 1)) : 0;
    if (adj_fast<FL>()) nblocks = (long long)nfast * nchunk + ((FL & ST_F_HABC) ? (long long)(bt.count + nstrip) * (tapped ? band_groups(a.B) : a.B) : 0);
    else nblocks = (long long)bt.nxt * bt.nzt * nchunk;
    if constexpr (adj_fast<FL>() && (FL & ST_F_HABC) && adj_split<FL>()) {
        // two launches: the frame blocks (tap gather, needs the registers) and the fast blocks (compiled for more resident
        // blocks per SM); they write disjoint cells and read the same inputs
        const long long nband = nblocks - (long long)nfast * nchunk;
        if (nband > 0 && pdl_launch(wave2d_adjoint_kernel<FL, 1>, dim3((unsigned)nband), st, a, nfx, nfast, bt) != ST_OK) return ST_ERR_CUDA;
        return pdl_launch(wave2d_adjoint_kernel<FL, 2>, dim3((unsigned)((long long)nfast * nchunk)), st, a, nfx, nfast, bt);
    }
    dim3 grid((unsigned)nblocks);
    return pdl_launch(wave2d_adjoint_kernel<FL, 0>, grid, st, a, nfx, nfast, bt);
}
#if defined(ST_DBG_TIMELINE) && defined(ST_W2_INSTANCE)
#if ST_W2_INSTANCE == 5
extern "C" int st_debug_timeline(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_dbg_tl, sizeof(unsigned long long) * n);
}
#endif
#endif
#ifdef ST_W2_INSTANCE
template int st_w2_launch_fwd<ST_W2_INSTANCE>(const W2Args&, const W2Tma&, cudaStream_t);
template int st_w2_launch_adj<ST_W2_INSTANCE>(const W2Args&, const W2Tma&, cudaStream_t);
#endif
#endif  // !ST_W2_DISPATCH_ONLY

#if defined(ST_W2_DISPATCH_ONLY) || !defined(ST_W2_INSTANCE)
#define ST_W2_DISPATCH(FN)                                                                          \
    switch (flags) {                                                                                \
        case ST_F_ISO | ST_F_PML: return FN<ST_F_ISO | ST_F_PML>(a, tm, st);                        \
        case ST_F_ISO | ST_F_HABC: return FN<ST_F_ISO | ST_F_HABC>(a, tm, st);                      \
        case ST_F_HABC: return FN<ST_F_HABC>(a, tm, st);                                            \
        case ST_F_HABC | ST_F_XZ: return FN<ST_F_HABC | ST_F_XZ>(a, tm, st);                        \
        case ST_F_HABC | ST_F_G1: return FN<ST_F_HABC | ST_F_G1>(a, tm, st);                        \
        case ST_F_ISO | ST_F_HABC | ST_F_G1: return FN<ST_F_ISO | ST_F_HABC | ST_F_G1>(a, tm, st);  \
        case ST_F_HABC | ST_F_BORN: return FN<ST_F_HABC | ST_F_BORN>(a, tm, st);                    \
        case ST_F_HABC | ST_F_XZ | ST_F_BORN: return FN<ST_F_HABC | ST_F_XZ | ST_F_BORN>(a, tm, st);\
        default: st_set_error("wave2d: unsupported flag set %d", flags); return ST_ERR_UNSUPPORTED; \
    }

int st_wave2d_launch_forward(int flags, const W2Args& a, const W2Tma& tm, cudaStream_t st) { ST_W2_DISPATCH(st_w2_launch_fwd) }

int st_wave2d_launch_adjoint(int flags, const W2Args& a, const W2Tma& tm, cudaStream_t st) { ST_W2_DISPATCH(st_w2_launch_adj) }

// Whether the TMA kernels apply, and their tiling.  They take over the WHOLE launch or nothing:
//   acoustic (PML): every tile is a frame-free tile (the tile column past nx is masked);
//   acoustic_habc : a band of whole tile columns [tx0, tx1) whose cells and x-neighbours are deeper than bw+1 from
//                   the left / right edge (straight top / bottom frame only), plus rows [sr0, sr1) of the one tile
//                   column on either side (straight left / right frame only); the four corners run the generic
//                   per-cell code.  Needs exactly one tile column per side and room between the corners.
int st_wave2d_tma_setup(int flags, const W2Args& a, const float* u, long long u_planes, const float* lam, long long lam_planes,
                        bool adjoint, int mode, W2Tma& tm) {
    memset(&tm, 0, sizeof(tm));
    if (mode == 0) return ST_OK;
    if (flags != (ST_F_ISO | ST_F_PML) && flags != (ST_F_ISO | ST_F_HABC)) return ST_OK;
    const W2Geom& g = a.g;
    const int nfx = (g.nx + FW - 1) / FW;
    int tx0 = 0, tx1 = nfx;
    tm.ntr = (g.nz + TR - 1) / TR;
    if (flags & ST_F_HABC) {
        tx0 = (g.bw + 2 + FW - 1) / FW;
        tx1 = (g.nx - g.bw - 2) / FW;
        const int fz0 = (g.bw + 2 + FH - 1) / FH, fz1 = (g.nz - g.bw - 2) / FH;
        if (tx0 != 1 || tx1 != nfx - 1 || tx1 <= tx0 || fz1 <= fz0 || !st_band_ok(g, g.bw + 1) ||
            g.nz < 2 * (g.bw + 2) + 2 * TR) return ST_OK;
        tm.sr0 = fz0 * (FH / TR);
        tm.sr1 = fz1 * (FH / TR);
        tm.band = adjoint ? g.bw + 2 : g.bw;
        tm.nbot = tm.ntr - (g.nz - tm.band) / TR;            // tile rows with z0 + TR > nz - band
    }
    const long long work = (long long)g.nz * g.nx * a.B;
    if (mode < 0 && (a.B < 2 || work < (1LL << 21))) return ST_OK;       // too little work to fill the rings
    tm.tx0 = tx0; tm.tx1 = tx1;
    // shots per block: all (up to 8) in the forward pass; in the adjoint at least bchunk (the caller sized the
    // gradient planes for ceil(B / bchunk) groups)
    tm.tsh = a.B < 8 ? a.B : 8;
    if (adjoint && tm.tsh < a.bchunk) tm.tsh = a.bchunk;
    if (const char* e = getenv("SEISTORCH_B200_TSH")) {
        const int v = atoi(e);
        if (v > 0 && (!adjoint || v >= a.bchunk)) tm.tsh = v < a.B ? v : a.B;
    }
    const long long fs = a.fs;
    int rc = st_tma_encode_planes(&tm.u_h1, u, g.nx, g.nz, u_planes, g.ld, fs, HC, H1R);
    if (!rc) rc = st_tma_encode_planes(&tm.u_h2, u, g.nx, g.nz, u_planes, g.ld, fs, HC, H2R);
    if (!rc && !adjoint) rc = st_tma_encode_planes(&tm.u_core, u, g.nx, g.nz, u_planes, g.ld, fs, TC, TR);
    if (!rc && adjoint) rc = st_tma_encode_planes(&tm.l_h1, lam, g.nx, g.nz, lam_planes, g.ld, fs, HC, H1R);
    if (!rc && adjoint) rc = st_tma_encode_planes(&tm.l_h2, lam, g.nx, g.nz, lam_planes, g.ld, fs, HC, H2R);
    if (!rc && adjoint) rc = st_tma_encode_planes(&tm.l_core, lam, g.nx, g.nz, lam_planes, g.ld, fs, TC, TR);
    if (rc) { st_set_error("wave2d: cuTensorMapEncodeTiled failed (%d)", rc); return ST_ERR_CUDA; }
    tm.tpb = 1;
    if (const char* e = getenv("SEISTORCH_B200_TPB")) { if (atoi(e) > 0) tm.tpb = atoi(e); }
    // acquisition rows: one shot per block (measured: the source / receiver epilogue costs 3-4 us per item, a chain of
    // dependent loads, against 0.8 us for the item itself; eight of them in a row make those blocks finish 3x later
    // than all the others).  The adjoint needs one gradient plane per shot for that.
    const bool planes_ok = true;                            // gacc has max(nchunk, B) planes (include/seistorch_b200.h)
    const bool any_acq = a.ns > 0 || a.R > 0;
    if (tm.tpb == 1 && tm.tsh > 1 && planes_ok && any_acq && a.row_hi >= a.row_lo && a.row_hi - a.row_lo < 8 * TR &&
        getenv("SEISTORCH_B200_TMA_NOACQ") == nullptr) {
        tm.ar0 = a.row_lo / TR;
        tm.ar1 = a.row_hi / TR + 1;
    }
    tm.enabled = 1;
    return ST_OK;
}
#endif
