// Launch-argument block shared by the 2D second-order kernels and their host drivers.
#pragma once
#include "st_wave2d_math.cuh"

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

#ifdef __CUDACC__
int st_wave2d_launch_forward(int flags, const W2Args& a, cudaStream_t st);
int st_wave2d_launch_adjoint(int flags, const W2Args& a, cudaStream_t st);
#endif
