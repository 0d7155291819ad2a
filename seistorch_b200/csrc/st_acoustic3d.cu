// sm_100a kernels for the 3D acoustic equation with PML damping.
// Reference: equations3d/acoustic.py:65-85 (conv3d 7-point Laplacian + ~12 elementwise
// passes per step, each over a 300 MB field at the BASELINE size).
//
// 2.5-D marching: a block owns an (n1,n2) tile and walks CH planes along n0; the
// in-plane neighbours come from a shared-memory plane tile, the out-of-plane ones from a
// three-register pipeline, so every field value is read from HBM/L2 once per brick.
#include "st_acoustic3d.cuh"

namespace {

constexpr int TX = 64, TY = 8, NT = TX * TY;
constexpr int CH = 16;      // planes per brick

__device__ __forceinline__ float ld0(const float* p, long long plane_off, int i1, int i2, int n1, int n2, int ld) {
    if (i1 < 0 || i1 >= n1 || i2 < 0 || i2 >= n2) return 0.f;
    return __ldg(p + plane_off + (long long)i1 * ld + i2);
}

__global__ void __launch_bounds__(NT) acoustic3d_forward_kernel(const A3Args a) {
    __shared__ float sp[TY + 2][TX + 2];
    const int n0 = a.n0, n1 = a.n1, n2 = a.n2, ld = a.ld;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int nch = (n0 + CH - 1) / CH;
    const int b = blockIdx.z / nch, c0 = (blockIdx.z % nch) * CH, c1 = min(c0 + CH, n0);
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int i2 = x0 + tx, i1 = y0 + ty;
    const bool active = i2 < n2 && i1 < n1;
    const long long col = (long long)i1 * ld + i2;
    const float* cur = a.cur + (long long)b * a.fs;
    const float* prv = a.prev + (long long)b * a.fs;
    float* nxt = a.next + (long long)b * a.fs;

    float behind = (active && c0 > 0) ? __ldg(cur + (long long)(c0 - 1) * a.ps + col) : 0.f;
    float center = active ? __ldg(cur + (long long)c0 * a.ps + col) : 0.f;
    for (int i0 = c0; i0 < c1; ++i0) {
        const long long po = (long long)i0 * a.ps;
        const float ahead = (active && i0 + 1 < n0) ? __ldg(cur + po + a.ps + col) : 0.f;
        __syncthreads();
        sp[ty + 1][tx + 1] = center;
        if (ty == 0) sp[0][tx + 1] = ld0(cur, po, i1 - 1, i2, n1, n2, ld);
        if (ty == TY - 1) sp[TY + 1][tx + 1] = ld0(cur, po, i1 + 1, i2, n1, n2, ld);
        if (tx == 0) sp[ty + 1][0] = ld0(cur, po, i1, i2 - 1, n1, n2, ld);
        if (tx == TX - 1) sp[ty + 1][TX + 1] = ld0(cur, po, i1, i2 + 1, n1, n2, ld);
        __syncthreads();
        if (active) {
            const float lap = ((sp[ty][tx + 1] - center) + (sp[ty + 2][tx + 1] - center))
                            + ((sp[ty + 1][tx] - center) + (sp[ty + 1][tx + 2] - center))
                            + ((behind - center) + (ahead - center));
            const float r = __ldg(a.r + po + col), bd = __ldg(a.b + po + col) * a.dt;
            const float inv = 1.f / (1.f + bd);
            const float h2 = __ldg(prv + po + col);
            nxt[po + col] = center + ((1.f - bd) * inv) * (center - h2) + (r * r * inv) * lap;
        }
        behind = center;
        center = ahead;
    }
    __syncthreads();
    // ---- fused source add (source.py:59-70)
    for (int s = tid; s < a.ns; s += NT) {
        if (a.src_b[s] != b) continue;
        const int s0 = a.src_i0[s], s1 = a.src_i1[s], s2 = a.src_i2[s];
        if (s0 >= c0 && s0 < c1 && s1 >= y0 && s1 < y0 + TY && s2 >= x0 && s2 < x0 + TX)
            atomicAdd(nxt + (long long)s0 * a.ps + (long long)s1 * ld + s2, a.amp[s]);
    }
    __syncthreads();
    // ---- fused receiver gather (probe.py:46-48)
    if (a.rec_out) {
        const int nrow = (c1 - c0) * TY;
        for (int q = tid; q < nrow; q += NT) {
            const int i0 = c0 + q / TY, r1 = y0 + q % TY;
            if (r1 >= n1) continue;
            const long long row = ((long long)b * n0 + i0) * n1 + r1;
            const int lo = a.row_start[row], hi = a.row_start[row + 1];
            for (int r = lo; r < hi; ++r) {
                const int rc = a.rec_col[r];
                if (rc >= x0 && rc < x0 + TX)
                    a.rec_out[a.rec_orig[r]] = nxt[(long long)i0 * a.ps + (long long)r1 * ld + rc];
            }
        }
    }
}

__global__ void __launch_bounds__(NT) acoustic3d_adjoint_kernel(const A3Args a) {
    __shared__ float sw[TY + 2][TX + 2];     // a3 * Lam_{i+1}
    __shared__ float ss[TY + 2][TX + 2];     // S_i
    const int n0 = a.n0, n1 = a.n1, n2 = a.n2, ld = a.ld;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int nch = (n0 + CH - 1) / CH;
    const int chunk = blockIdx.z / nch, c0 = (blockIdx.z % nch) * CH, c1 = min(c0 + CH, n0);
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int i2 = x0 + tx, i1 = y0 + ty;
    const bool active = i2 < n2 && i1 < n1;
    const long long col = (long long)i1 * ld + i2;
    const bool want_grad = a.gacc != nullptr;
    const float dt = a.dt;

    auto wval = [&](const float* l1, long long po, int j1, int j2) -> float {
        if (j1 < 0 || j1 >= n1 || j2 < 0 || j2 >= n2) return 0.f;
        const long long o = po + (long long)j1 * ld + j2;
        const float r = __ldg(a.r + o), bd = __ldg(a.b + o) * dt;
        return r * r / (1.f + bd) * __ldg(l1 + o);
    };

    const int b_lo = chunk * a.bchunk, b_hi = min(b_lo + a.bchunk, a.B);
    for (int b = b_lo; b < b_hi; ++b) {
        const float* l1 = a.lam1 + (long long)b * a.fs;
        const float* l2 = a.lam2 + (long long)b * a.fs;
        const float* S = a.s1 + (long long)b * a.fs;
        float* l0 = a.lam0 + (long long)b * a.fs;
        float* gb = want_grad ? a.gacc + (long long)chunk * a.fs : nullptr;

        float wb = 0.f, sb = 0.f, wc = 0.f, sc = 0.f;
        if (active) {
            if (c0 > 0) { wb = wval(l1, (long long)(c0 - 1) * a.ps, i1, i2); sb = __ldg(S + (long long)(c0 - 1) * a.ps + col); }
            wc = wval(l1, (long long)c0 * a.ps, i1, i2);
            sc = __ldg(S + (long long)c0 * a.ps + col);
        }
        for (int i0 = c0; i0 < c1; ++i0) {
            const long long po = (long long)i0 * a.ps;
            float wa = 0.f, sa = 0.f;
            if (active && i0 + 1 < n0) { wa = wval(l1, po + a.ps, i1, i2); sa = __ldg(S + po + a.ps + col); }
            __syncthreads();
            sw[ty + 1][tx + 1] = wc;
            ss[ty + 1][tx + 1] = sc;
            if (ty == 0) { sw[0][tx + 1] = wval(l1, po, i1 - 1, i2); ss[0][tx + 1] = ld0(S, po, i1 - 1, i2, n1, n2, ld); }
            if (ty == TY - 1) { sw[TY + 1][tx + 1] = wval(l1, po, i1 + 1, i2); ss[TY + 1][tx + 1] = ld0(S, po, i1 + 1, i2, n1, n2, ld); }
            if (tx == 0) { sw[ty + 1][0] = wval(l1, po, i1, i2 - 1); ss[ty + 1][0] = ld0(S, po, i1, i2 - 1, n1, n2, ld); }
            if (tx == TX - 1) { sw[ty + 1][TX + 1] = wval(l1, po, i1, i2 + 1); ss[ty + 1][TX + 1] = ld0(S, po, i1, i2 + 1, n1, n2, ld); }
            __syncthreads();
            if (active) {
                const float lapw = ((sw[ty][tx + 1] - wc) + (sw[ty + 2][tx + 1] - wc))
                                 + ((sw[ty + 1][tx] - wc) + (sw[ty + 1][tx + 2] - wc)) + ((wb - wc) + (wa - wc));
                const float r = __ldg(a.r + po + col), bd = __ldg(a.b + po + col) * dt;
                const float inv = 1.f / (1.f + bd), a2 = (1.f - bd) * inv;
                const float l1c = __ldg(l1 + po + col), l2c = __ldg(l2 + po + col);
                l0[po + col] = (1.f + a2) * l1c + lapw - a2 * l2c;
                if (want_grad) {
                    const float laps = ((ss[ty][tx + 1] - sc) + (ss[ty + 2][tx + 1] - sc))
                                     + ((ss[ty + 1][tx] - sc) + (ss[ty + 1][tx + 2] - sc)) + ((sb - sc) + (sa - sc));
                    gb[po + col] += l1c * (2.f * r * inv) * laps;
                }
            }
            wb = wc; wc = wa; sb = sc; sc = sa;
        }
        __syncthreads();
        if (a.rec_adj) {
            const int nrow = (c1 - c0) * TY;
            for (int q = tid; q < nrow; q += NT) {
                const int i0 = c0 + q / TY, r1 = y0 + q % TY;
                if (r1 >= n1) continue;
                const long long row = ((long long)b * n0 + i0) * n1 + r1;
                const int lo = a.row_start[row], hi = a.row_start[row + 1];
                for (int r = lo; r < hi; ++r) {
                    const int rc = a.rec_col[r];
                    if (rc >= x0 && rc < x0 + TX)
                        atomicAdd(l0 + (long long)i0 * a.ps + (long long)r1 * ld + rc, a.rec_adj[a.rec_orig[r]]);
                }
            }
        }
        if (a.gamp) {
            __syncthreads();
            for (int s = tid; s < a.ns; s += NT) {
                if (a.src_b[s] != b) continue;
                const int s0 = a.src_i0[s], s1 = a.src_i1[s], s2 = a.src_i2[s];
                if (s0 >= c0 && s0 < c1 && s1 >= y0 && s1 < y0 + TY && s2 >= x0 && s2 < x0 + TX)
                    a.gamp[s] = l0[(long long)s0 * a.ps + (long long)s1 * ld + s2];
            }
        }
    }
}

}  // namespace

int st_acoustic3d_launch_forward(const A3Args& a, cudaStream_t st) {
    const int nch = (a.n0 + CH - 1) / CH;
    dim3 grid((a.n2 + TX - 1) / TX, (a.n1 + TY - 1) / TY, a.B * nch), block(TX, TY);
    acoustic3d_forward_kernel<<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

int st_acoustic3d_launch_adjoint(const A3Args& a, cudaStream_t st) {
    const int nch = (a.n0 + CH - 1) / CH;
    const int nchunk = (a.B + a.bchunk - 1) / a.bchunk;
    dim3 grid((a.n2 + TX - 1) / TX, (a.n1 + TY - 1) / TY, nchunk * nch), block(TX, TY);
    acoustic3d_adjoint_kernel<<<grid, block, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}
