// One-off (per call) evaluation of the frame taps of the 9-tap HABC equations (acoustic_habc,
// vti_habc2, acoustic_fwim_habc); see
// st_wave2d_band.cuh.  Reads the coefficient planes, writes ST_TAP_PLANES planes [nz][ld]:
//   F1[0..8], F2[0..4], H1[0..8] = dF1/dr, H2[0..4] = dF2/dr   (ciso held fixed).
#include "st_wave2d_band.cuh"

namespace {

__global__ void __launch_bounds__(256) wave2d_prepare_kernel(const W2Args a, int flags) {
    const W2Geom g = a.g;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= g.nx || z >= g.nz) return;
    const long long plane = (long long)g.nz * g.ld;
    const long long idx = (long long)z * g.ld + x;
    const float r = a.coef[0][idx], b = a.coef[1][idx];
    const float cx = a.coef[2][idx];
    const float cz = (flags & ST_F_ISO) ? cx : a.coef[3][idx];
    const float ax = (flags & ST_F_G1) ? a.coef[5][idx] : 0.f, az = (flags & ST_F_G1) ? a.coef[6][idx] : 0.f;
    const float cxz = (flags & ST_F_XZ) ? a.coef[4][idx] : 0.f;
    float F1[ST_NTAP1], F2[ST_NTAP2], H1[ST_NTAP1], H2[ST_NTAP2];
#pragma unroll
    for (int o = 0; o < ST_NTAP1; ++o) F1[o] = H1[o] = 0.f;
#pragma unroll
    for (int o = 0; o < ST_NTAP2; ++o) F2[o] = H2[o] = 0.f;
    const bool frame = w2_in_frame(z, x, g);
    const float pre = frame ? 1.f - b : 1.f;
    // y = 2 h1 - h2 + cx (E + W - 2C) + cz (N + S - 2C) + ax (E - W) + az (S - N)
    F1[0] = pre * (2.f - 2.f * cx - 2.f * cz);
    F1[1] = pre * (cz - az);      // N = (z-1, x)
    F1[2] = pre * (cz + az);      // S
    F1[3] = pre * (cx - ax);      // W = (z, x-1)
    F1[4] = pre * (cx + ax);      // E
    F2[0] = -pre;
    // + cxz ((SE - SW) - (NE - NW))   (tti_habc.py:40-57; zero planes for the other equations)
#pragma unroll
    for (int o = ST_NTAP1C; o < ST_NTAP1; ++o) F1[o] = pre * cxz * st_tap_xz_sign(o);
    if (frame) {
        float f[4];
        w2_side_weights(z, x, g, f);
        const float lam = 2.f * r, mu = r * r;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (f[s] == 0.f) continue;
            const float w = b * f[s];
            const int j = w2_depth(s, z, x, g);
            const int o1 = st_tap_normal(s, 1), o2 = st_tap_normal(s, 2);
            // one = (2 - lam - mu) h1_j + (lam + 2 mu) h1_{j+1} - mu h1_{j+2} + (lam - 1) h2_j - lam h2_{j+1}
            F1[0] += w * (2.f - lam - mu);  H1[0] += w * (-2.f - 2.f * r);
            F1[o1] += w * (lam + 2.f * mu); H1[o1] += w * (2.f + 4.f * r);
            if (j + 2 <= g.bw) { F1[o2] += w * (-mu); H1[o2] += w * (-2.f * r); }     // else: wrap fix-up
            F2[0] += w * (lam - 1.f);       H2[0] += w * 2.f;
            F2[o1] += w * (-lam);           H2[o1] += w * (-2.f);
        }
    }
    float* t = a.taps + idx;
#pragma unroll
    for (int o = 0; o < ST_NTAP1; ++o) t[o * plane] = F1[o];
#pragma unroll
    for (int o = 0; o < ST_NTAP2; ++o) t[(ST_NTAP1 + o) * plane] = F2[o];
#pragma unroll
    for (int o = 0; o < ST_NTAP1; ++o) t[(ST_NTAP1 + ST_NTAP2 + o) * plane] = H1[o];
#pragma unroll
    for (int o = 0; o < ST_NTAP2; ++o) t[(2 * ST_NTAP1 + ST_NTAP2 + o) * plane] = H2[o];
}

}  // namespace

int st_wave2d_launch_prepare(int flags, const W2Args& a, cudaStream_t st) {
    dim3 grid((a.g.nx + 255) / 256, a.g.nz);
    wave2d_prepare_kernel<<<grid, 256, 0, st>>>(a, flags);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}
