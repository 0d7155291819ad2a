// Gradient post-processing on the device (include/seistorch_b200.h: st_gaussian_smooth2d).
//
// Reference: PostProcess.smooth_gradient (seistorch/process.py:66-112) copies every parameter gradient to the host,
// reflect-pads it with numpy, convolves it with a truncated Gaussian along z then x through conv2d on the CPU
// (signal.py:247-319) `counts` times, and copies it back.  One pass of that filter is
//     y(i) = sum_{j=-p..p} w[j+p] x(mirror(i+j)),   mirror = numpy 'reflect' (no edge repeat), p = radius,
// with the normalised weights w (kernel size 2 radius + 1; the reference's even kernel size for odd radii changes the
// array shape and cannot be assigned back to the parameter: not supported, the Python wrapper raises).
// One thread per output element; the gradient plane (<= 8 MB at the BASELINE sizes) lives in L2.
#include <cuda_runtime.h>

#include "../../include/seistorch_b200.h"
#include "st_common.cuh"

namespace {

__device__ __forceinline__ int mirror(int i, int n) {
    // numpy.pad(mode='reflect') index map, valid for |offset| < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(256) gaussian_smooth2d_kernel(const float* __restrict__ in, float* __restrict__ out, int nz, int nx,
                                                                 const float* __restrict__ w, int radius, int axis) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx || z >= nz) return;
    float acc = 0.f;
    if (axis == 0) {
        for (int j = -radius; j <= radius; ++j) acc += __ldg(w + j + radius) * __ldg(in + (long long)mirror(z + j, nz) * nx + x);
    } else {
        const float* row = in + (long long)z * nx;
        for (int j = -radius; j <= radius; ++j) acc += __ldg(w + j + radius) * __ldg(row + mirror(x + j, nx));
    }
    out[(long long)z * nx + x] = acc;
}

// source illumination (rnn.py:127-128,204-205): out(cell) += sum over time steps and shots of field^2, read back from
// the wavefield history the gradient run keeps anyway.  One thread per 4 consecutive cells (128-bit loads).
__global__ void __launch_bounds__(256) illumination_kernel(const float* __restrict__ u, long long slot_stride, int nslots, int slot_first,
                                                            int count, long long chan_offset, int B, long long fs, long long n4,
                                                            float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 acc = reinterpret_cast<float4*>(out)[i];
    int slot = slot_first % nslots;
    for (int k = 0; k < count; ++k) {
        const float* base = u + slot_stride * slot + chan_offset;
        for (int b = 0; b < B; ++b) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(base + b * fs) + i);      // streaming: read once
            acc.x += v.x * v.x; acc.y += v.y * v.y; acc.z += v.z * v.z; acc.w += v.w * v.w;
        }
        slot = slot + 1 == nslots ? 0 : slot + 1;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
}

}  // namespace

extern "C" int st_illumination(const float* u, int64_t slot_stride, int32_t nslots, int32_t slot_first, int32_t count,
                               int64_t chan_offset, int32_t B, int64_t fs, float* out, void* stream) {
    if (!u || !out || nslots <= 0 || count < 0 || B <= 0 || fs <= 0 || fs % 4 != 0 || slot_stride % 4 != 0 || chan_offset % 4 != 0) {
        st_set_error("illumination: bad arguments (planes must be multiples of 4 floats)");
        return ST_ERR_BADARG;
    }
    if (count == 0) return ST_OK;
    const long long n4 = fs / 4;
    illumination_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(u, slot_stride, nslots, slot_first, count,
                                                                                         chan_offset, B, fs, n4, out);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("illumination: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_gaussian_smooth2d(const float* in, float* out, int32_t nz, int32_t nx, const float* weights, int32_t radius,
                                    int32_t axis, void* stream) {
    if (!in || !out || !weights || nz <= 0 || nx <= 0 || radius < 0 || (axis != 0 && axis != 1) || in == out) {
        st_set_error("gaussian_smooth2d: bad arguments");
        return ST_ERR_BADARG;
    }
    if (radius >= (axis == 0 ? nz : nx)) {
        st_set_error("gaussian_smooth2d: radius %d does not fit the axis (reflect padding needs radius < n)", radius);
        return ST_ERR_UNSUPPORTED;
    }
    dim3 grid((nx + 255) / 256, nz);
    gaussian_smooth2d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, nz, nx, weights, radius, axis);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("gaussian_smooth2d: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}
