// 3D acoustic (PML): launch arguments.  Reference: equations3d/acoustic.py:65-85.
// Tensor layout (B, n0, n1, n2) = the reference's (B, x, z, y); n2 fastest, pitch ld.
#pragma once
#include "st_common.cuh"

struct A3Args {
    int n0, n1, n2, ld, B;
    float dt;
    long long ps, fs;           // plane stride n1*ld, shot stride n0*ps (floats)
    const float* r; const float* b;          // [n0][n1][ld]
    const float* prev; const float* cur; float* next;       // forward  [B][n0][n1][ld]
    const float* lam1; const float* lam2; float* lam0;       // adjoint
    const float* s1;                                          // S_i
    float* gacc;                                              // [nchunk][n0][n1][ld]
    int bchunk;
    int ns; const int* src_b; const int* src_i0; const int* src_i1; const int* src_i2;
    const float* amp; float* gamp;
    const int* row_start; const int* rec_col; const int* rec_orig; int R;
    float* rec_out; const float* rec_adj;
};

#ifdef __CUDACC__
int st_acoustic3d_launch_forward(const A3Args& a, cudaStream_t st);
int st_acoustic3d_launch_adjoint(const A3Args& a, cudaStream_t st);
#endif
