// Common definitions for the seistorch_b200 CUDA kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cstdio>

#ifdef __CUDACC__
#define ST_HD __host__ __device__ __forceinline__
#else
#define ST_HD inline
#endif

// ---- equation-variant flags of the 2D second-order family ---------------------------
// (which reference plugin maps to which flag set is listed in include/seistorch_b200.h)
enum : int {
    ST_F_ISO  = 1,    // Cxx == Czz == r^2 (acoustic-type Laplacian); coefficient derived from r in-kernel
    ST_F_PML  = 2,    // damped update  y = h1 + a2 (h1-h2) + a3 lap(h1)      (equations2d/acoustic.py:73-86)
    ST_F_HABC = 4,    // undamped update followed by the Higdon one-way blend   (equations2d/acoustic_habc.py:147-221)
    ST_F_XZ   = 8,    // mixed-derivative term (TTI)                            (equations2d/tti_habc.py:40-57)
    ST_F_G1   = 16,   // first-derivative terms (joint FWI-LSRTM "FWIM")        (equations2d/acoustic_fwim_habc.py:38-60)
    ST_F_BORN = 32,   // background + scattered pair coupled through m          (equations2d/acoustic_*_lsrtm_habc.py)
};

// error codes returned by every C-ABI entry point
#define ST_OK 0
#ifndef ST_ERR_BADARG
#define ST_ERR_BADARG (-1)
#define ST_ERR_UNSUPPORTED (-2)
#define ST_ERR_CUDA (-3)
#endif

void st_set_error(const char* fmt, ...);

#ifdef __CUDACC__
#include <atomic>
#include <cstdlib>
// Programmatic dependent launch (the step kernels of one time loop are launched back to back on one stream):
// st_pdl_launch_dependents() lets the next grid start launching once every block of this one has begun,
// st_pdl_wait() blocks until the previous grid has completed and its writes are visible.  Everything a block
// does before st_pdl_wait() (tile decode, mbarrier init, descriptor prefetch) overlaps the previous kernel's tail; on small
// grids the launch processing and block scheduling of step i+1 overlap step i.  No-ops in a plain launch.
__device__ __forceinline__ void st_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void st_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// launch with programmatic stream serialization (SEISTORCH_B200_PDL=0 turns it off)
template <class K, class... Args>
inline cudaError_t st_pdl_launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static const bool pdl = [] { const char* e = getenv("SEISTORCH_B200_PDL"); return !(e && atoi(e) == 0); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
// Raise the dynamic shared-memory limit of one kernel.  The attribute is per DEVICE, so it is set once per
// (kernel, device) -- a process that drives several GPUs gets it on each of them -- and a failure is not cached.
template <auto Kernel>
inline cudaError_t st_set_max_smem(int bytes) {
    static std::atomic<unsigned long long> done{0};      // bit d = set on device d (d < 64)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (dev < 64 && (done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
    e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) done.fetch_or(bit, std::memory_order_release);
    return e;
}
#endif
