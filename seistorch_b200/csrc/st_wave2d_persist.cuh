// Launch plan of the persistent multi-timestep kernels (st_wave2d_persist.cu).
#pragma once
#include "st_wave2d.cuh"

#define ST_PERSIST_NA 1                 // "not applicable": the caller falls back to the per-step kernels

struct W2Persist {
    float* u;                           // field / history buffer [nslots][B][nz][ld]
    float* lam;                         // adjoint: cotangent ring [3][B][nz][ld] (Lam_i in slot i mod 3), else nullptr
    long long slot;                     // floats per state slot (= B * nz * ld)
    int nslots, slot0;                  // S_{i0-2} sits in slot0, S_{i0-1} in slot0 + 1 (mod nslots)
    int i0, nsteps;                     // forward: first step, ascending; adjoint: highest step i_hi, descending
                                        // (adjoint: slot0 = history slot of S_{i_hi})
    int history;                        // 1: every state is stored (slot0 + k + 2); 0: only the last two (rolling state)
    int cs;                             // CTAs per cluster (= per shot)
    int nstrips, nrg;                   // 128-column strips per row, row groups per CTA
    int rpc, ldp;                       // rows per CTA, row pitch of the published shared-memory copy
    int variant;                        // thread shape: 0 = 16 warps x 4 rows per thread, 1 = 32 warps x 2 rows
    int probe;                          // 1: only check that a cluster of this shape can be resident
};

#ifdef __CUDACC__
int st_wave2d_persist_plan(int flags, const W2Args& a, bool adjoint, W2Persist& pp);
int st_wave2d_persist_forward(const W2Args& a, const W2Persist& pp, cudaStream_t st);
int st_wave2d_persist_adjoint(const W2Args& a, const W2Persist& pp, cudaStream_t st);
#endif
