// Launch-argument block shared by the 2D second-order kernels and their host drivers.
#pragma once
#include "st_wave2d_math.cuh"
#include "st_tma.cuh"

struct W2Args {
    W2Geom g;
    int B;                  // shots in this call
    float dt;
    long long fs;           // floats per (shot, field) plane  = nz*ld
    long long cs;           // floats per field channel        = B*fs
    const float* coef[8];   // r,b,cxx,czz,cxz,ax,az,m  each [nz][ld]; nullptr if unused
    float* taps;            // ISO|HABC: ST_TAP_PLANES precomputed frame-tap planes (st_wave2d_band.cuh) or nullptr
    // ---- forward: field states  [NF][B][nz][ld]
    const float* prev;      // S_{i-2}
    const float* cur;       // S_{i-1}
    float* next;            // S_i  (output)
    // ---- adjoint: cotangents [NF][B][nz][ld]
    const float* lam1;      // Lam_{i+1}
    const float* lam2;      // Lam_{i+2}
    float* lam0;            // Lam_i (output)
    const float* s1;        // S_i
    const float* s2;        // S_{i-1}
    float* gacc;            // [nchunk][7][nz*ld] coefficient-gradient accumulators (or nullptr)
    int bchunk;             // shots per block in the adjoint kernel
    // ---- sources (one entry per point source)
    int ns;
    const int* src_b; const int* src_z; const int* src_x;
    const float* amp;       // [ns] amplitudes of this step (forward)
    float* gamp;            // [ns] d loss / d amplitude of this step (adjoint; nullptr to skip)
    int src_fmask;          // bit f set: inject into field channel f
    // ---- receivers, sorted by (shot, z, x); CSR over rows (shot*nz + z)
    const int* row_start; const int* rec_x; const int* rec_orig;
    int row_lo, row_hi;     // rows that hold sources/receivers (epilogue skipped elsewhere)
    int R;                  // total receivers (all shots)
    int nchan; int chan_f[4];
    float* rec_out;         // forward:  [R][nchan] sample of this step
    const float* rec_adj;   // adjoint:  [R][nchan] d loss / d sample of this step
};

// TMA-staged interior path (acoustic / acoustic_habc): a rectangle of whole fast tiles whose cells are
// all frame-free is processed by blocks that pull (tile, shot) boxes through a shared-memory ring
// of bulk tensor loads.  Built once per C-ABI call (st_wave2d_tma_setup), passed as a
// __grid_constant__ kernel parameter.
constexpr int ST_TMA_TC = 128;              // tile columns  (one float4 per lane)
constexpr int ST_TMA_TR = 16;               // tile rows     (2 per warp)
#ifndef ST_TMA_XHALO
#define ST_TMA_XHALO 4
#endif
constexpr int ST_TMA_XO = ST_TMA_XHALO;        // halo columns fetched on either side of the core (keeps it 16-byte aligned)
constexpr int ST_TMA_HC = ST_TMA_TC + 2 * ST_TMA_XO;   // halo box columns: [x0-XO, x0+TC+XO)
constexpr int ST_TMA_H1 = ST_TMA_TR + 2;    // 1-deep halo box rows: [z0-1, z0+TR+1)
constexpr int ST_TMA_H2 = ST_TMA_TR + 4;    // 2-deep halo box rows: [z0-2, z0+TR+2)  (one-way blend of the top / bottom frame)
struct alignas(64) W2Tma {
    CUtensorMap u_h1, u_h2, u_core;         // boxes over the field/history buffer `u`
    CUtensorMap l_h1, l_h2, l_core;         // boxes over the adjoint ring `lam`
    int enabled;                            // 0: no TMA blocks in this launch
    int tx0, tx1;                           // the column band, in fast-tile units (FW columns); all rows
    int ntr;                                // tile rows = ceil(nz / TR)
    int nbot;                               // tile rows touching the bottom frame (enumerated first: heaviest tiles)
    int sr0, sr1;                           // HABC: tile rows [sr0, sr1) of the two side columns (tile column 0 and
                                            // nfx-1) are TMA tiles too (straight left / right frame); sr1 <= sr0: none
    int band;                               // HABC: rows closer than this to the top / bottom edge make a frame tile
    int tsh;                                // shots per TMA block
    int ar0, ar1;                           // tile rows [ar0, ar1) hold the sources / receivers: their tiles run one shot per
                                            // block (the source / receiver epilogue is a dependent-load chain that must
                                            // not be repeated tsh times inside one block); ar1 <= ar0: no such rows
    int tpb;                                // consecutive tiles (same kind: along x in the band, along z in a side
                                            // column) one TMA block streams through its ring
    // plane index (within u / lam) of shot 0 of the slots used by this step
    int pl_prev, pl_cur, pl_l1, pl_l2, pl_s1, pl_s2;
};

#ifdef __CUDACC__
int st_wave2d_launch_forward(int flags, const W2Args& a, const W2Tma& tm, cudaStream_t st);
int st_wave2d_launch_adjoint(int flags, const W2Args& a, const W2Tma& tm, cudaStream_t st);
// decides whether/where the TMA path applies (mode: 0 off, 1 forced, -1 auto) and encodes the maps
int st_wave2d_tma_setup(int flags, const W2Args& a, const float* u, long long u_planes, const float* lam, long long lam_planes,
                        bool adjoint, int mode, W2Tma& tm);
#endif
