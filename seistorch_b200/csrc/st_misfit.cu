// Misfit + adjoint-source kernels (include/seistorch_b200.h: st_misfit_*).
//   L2        : seistorch/loss.py:409-421   sum (syn-obs)^2           adj = 2 (syn-obs)
//   Envelope  : seistorch/loss.py:178-216 (method 'square') with the Hilbert transform of
//               seistorch/transform.py:27-66 (nfft = nt) restated as a circular
//               convolution with the fixed kernel hker = Im ifft(one-sided filter):
//                 E^2 = x^2 + (hker (*) x)^2,  loss = 0.5 sum (Es^2 - Eo^2)^2,
//                 adj = 2 r x + H^T (2 r Hx),  r = Es^2 - Eo^2.
//   L1        : seistorch/loss.py:381-393   sum |syn-obs|             adj = sign(syn-obs)
//   SML1      : seistorch/loss.py:395-407   SmoothL1Loss(sum, beta = 1e-3): 0.5 d^2/beta for |d| < beta else |d| - 0.5 beta
//   CC        : seistorch/loss.py:126-176   minus the zero-lag cross-correlation: - sum syn*obs          adj = -obs
//   INTEGRATION: seistorch/loss.py:366-379  per shot MSELoss(mean) of the time integrals (transform.integrate = cumsum)
//   CS        : seistorch/loss.py:52-85     per shot mean over traces of 1 - cos(syn_tr, obs_tr) along time
//               (F.cosine_similarity, eps = 1e-10):  sim = <x,y> / (max(|x|,eps) max(|y|,eps)),
//               adj = -(1/ntraces_of_the_shot) ( y/(|x||y|) - sim x/|x|^2 ).
//   NIM       : seistorch/loss.py:463-501 (criterion 'l2', method 'square'; Donno et al.): per trace
//               X = x^2 / sum_t x^2,  C = cumsum_t X  (same for obs),  loss = sum (Cx - Cy)^2;
//               adj_u = (2 x_u / Sx) (R_u - sum_s R_s X_s),  R_s = sum_{t >= s} 2 (Cx_t - Cy_t).
//   W1D       : seistorch/loss.py:900-955 (method 'linear'): the same with X = (x - c) / (sum_t (x - c) + 1e-18),
//               c = 1.1 min(min x, min y, 0) over the shot (a constant for the gradient);  adj_u = (R_u - G) / S.
//   filtfilt  : seistorch/signal.py:49-101 (backend 'torch': torchaudio filtfilt in double, clamp=False, zero initial
//               state, no padding): y = flip(lfilter(flip(lfilter(x)))).  As a matrix A^T A with A the causal IIR
//               operator, hence self-adjoint: the cotangent of the input is filtfilt(cotangent of the output).
//   TRAVELTIME: seistorch/loss.py:674-728 + signal.py:203-208: per trace, both records normalised by their max |.|
//               (+1e-16), full cross-correlation over the 2nt-1 lags, softmax over the lags (beta 1), expected lag
//               index minus (nt-1) = traveltime difference tau; loss = mean over traces of tau^2.
// Seismograms are [nt][ntraces] (ntraces = receivers x channels, fastest).
#include <cuda_runtime.h>

#include "../../include/seistorch_b200.h"
#include "st_common.cuh"

namespace {

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) l2_kernel(const float* __restrict__ syn, const float* __restrict__ obs,
                                                 long long n, float scale, double* loss, float* adj) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = syn[i] - obs[i];
        acc += (double)d * (double)d;
        if (adj) adj[i] = 2.f * scale * d;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * (double)scale);
}

// KIND 0: l1, 1: smooth l1 (beta = par), 2: minus zero-lag cross-correlation
template <int KIND>
__global__ void __launch_bounds__(256) l1_kernel(const float* __restrict__ syn, const float* __restrict__ obs,
                                                 long long n, float par, float scale, double* loss, float* adj) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = syn[i] - obs[i];
        if (KIND == 0) {
            acc += (double)fabsf(d);
            if (adj) adj[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
        } else if (KIND == 1) {
            const float ad = fabsf(d);
            acc += ad < par ? 0.5 * (double)d * (double)d / (double)par : (double)ad - 0.5 * (double)par;
            if (adj) adj[i] = scale * (ad < par ? d / par : (d > 0.f ? 1.f : -1.f));
        } else {
            acc -= (double)syn[i] * (double)obs[i];
            if (adj) adj[i] = -scale * obs[i];
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * (double)scale);
}

// one thread per trace (coalesced across traces at every time sample): three sums over time, then the adjoint source
__global__ void __launch_bounds__(256) cs_kernel(const float* __restrict__ syn, const float* __restrict__ obs, int nt, int ntr,
                                                 float inv_mean, float scale, double* loss, float* adj) {
    const int tr = blockIdx.x * blockDim.x + threadIdx.x;
    double term = 0.0;
    if (tr < ntr) {
        double sxy = 0.0, sxx = 0.0, syy = 0.0;
        for (int t = 0; t < nt; ++t) {
            const double x = syn[(long long)t * ntr + tr], y = obs[(long long)t * ntr + tr];
            sxy += x * y; sxx += x * x; syy += y * y;
        }
        const double eps = 1e-10;
        const double nx = sqrt(sxx), ny = sqrt(syy);
        const double cx = nx > eps ? nx : eps, cy = ny > eps ? ny : eps;
        const double sim = sxy / (cx * cy);
        term = (1.0 - sim) * (double)inv_mean;
        if (adj) {
            // d sim / d x = y / (cx cy) - [|x| > eps] sim x / |x|^2
            const double a = 1.0 / (cx * cy), bq = nx > eps ? sim / sxx : 0.0;
            const double w = -(double)scale * (double)inv_mean;
            for (int t = 0; t < nt; ++t) {
                const long long i = (long long)t * ntr + tr;
                adj[i] = (float)(w * (a * (double)obs[i] - bq * (double)syn[i]));
            }
        }
    }
    term = block_sum(term);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, term * (double)scale);
}

// one thread per trace; four sweeps over time (sums, loss + total residual, weighted residual sum, adjoint source)
// MODE 0: nim (samples squared);  MODE 1: w1d (samples shifted by -c, c read from `shift`, 1e-18 added to the sums);
// MODE 2: integration (plain cumsum, no normalisation: the caller folds the 1/N of the mean into `scale`)
template <int MODE>
__global__ void __launch_bounds__(256) nim_kernel(const float* __restrict__ syn, const float* __restrict__ obs, int nt, int ntr,
                                                  const float* __restrict__ shift, float scale, double* loss, float* adj) {
    const int tr = blockIdx.x * blockDim.x + threadIdx.x;
    double term = 0.0;
    const double c = MODE == 1 ? (double)__ldg(shift) : 0.0;
    auto tf = [&](double v) { return MODE == 0 ? v * v : v - c; };          // the non-negative transform (identity for MODE 2)
    if (tr < ntr) {
        double sx = 0.0, sy = 0.0;
        for (int t = 0; t < nt; ++t) {
            const double x = syn[(long long)t * ntr + tr], y = obs[(long long)t * ntr + tr];
            sx += tf(x); sy += tf(y);
        }
        if (MODE == 1) { sx += 1e-18; sy += 1e-18; }
        if (MODE == 2) sx = sy = 1.0;
        double cx = 0.0, cy = 0.0, T = 0.0;
        for (int t = 0; t < nt; ++t) {
            const double x = syn[(long long)t * ntr + tr], y = obs[(long long)t * ntr + tr];
            cx += tf(x) / sx; cy += tf(y) / sy;
            const double r = cx - cy;
            term += r * r;
            T += 2.0 * r;
        }
        if (adj) {
            double P = 0.0, G = 0.0;          // P: residual sum strictly before s;  R_s = T - P
            cx = cy = 0.0;
            for (int t = 0; t < nt; ++t) {
                const double x = syn[(long long)t * ntr + tr], y = obs[(long long)t * ntr + tr];
                G += (T - P) * (tf(x) / sx);
                cx += tf(x) / sx; cy += tf(y) / sy;
                P += 2.0 * (cx - cy);
            }
            P = 0.0; cx = cy = 0.0;
            for (int t = 0; t < nt; ++t) {
                const long long i = (long long)t * ntr + tr;
                const double x = syn[i], y = obs[i];
                adj[i] = (float)((double)scale * ((MODE == 0 ? 2.0 * x : 1.0) / sx) * ((T - P) - (MODE == 2 ? 0.0 : G)));
                cx += tf(x) / sx; cy += tf(y) / sy;
                P += 2.0 * (cx - cy);
            }
        }
    }
    term = block_sum(term);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, term * (double)scale);
}

constexpr int FILT_MAXC = 12;             // coefficients per polynomial (Butterworth band-pass of order 5 has 11)
struct FiltCoef { double b[FILT_MAXC], a[FILT_MAXC]; int n; };
// one thread per trace: causal pass into `work` (double), anti-causal pass into `y`; transposed direct form II
__global__ void __launch_bounds__(128) filtfilt_kernel(const float* __restrict__ x, float* __restrict__ y, double* __restrict__ work,
                                                       int nt, int ntr, const FiltCoef c) {
    const int tr = blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= ntr) return;
    double z[FILT_MAXC];
#pragma unroll
    for (int k = 0; k < FILT_MAXC; ++k) z[k] = 0.0;
    for (int t = 0; t < nt; ++t) {
        const double v = x[(long long)t * ntr + tr];
        const double o = c.b[0] * v + z[0];
#pragma unroll
        for (int k = 1; k < FILT_MAXC; ++k)
            if (k < c.n) z[k - 1] = c.b[k] * v + (k + 1 < c.n ? z[k] : 0.0) - c.a[k] * o;
        work[(long long)t * ntr + tr] = o;
    }
#pragma unroll
    for (int k = 0; k < FILT_MAXC; ++k) z[k] = 0.0;
    for (int t = nt - 1; t >= 0; --t) {
        const double v = work[(long long)t * ntr + tr];
        const double o = c.b[0] * v + z[0];
#pragma unroll
        for (int k = 1; k < FILT_MAXC; ++k)
            if (k < c.n) z[k - 1] = c.b[k] * v + (k + 1 < c.n ? z[k] : 0.0) - c.a[k] * o;
        y[(long long)t * ntr + tr] = (float)o;
    }
}

// one block per trace; dynamic shared memory: cc[2nt-1] (double), xn[nt], yn[nt] (float).  lagidx[k] = (n-1)*linspace(0,1,n)[k]
// exactly as the reference builds it (fp32 linspace), n = 2nt-1.
__global__ void __launch_bounds__(256) traveltime_kernel(const float* __restrict__ syn, const float* __restrict__ obs, int nt, int ntr,
                                                         const float* __restrict__ lagidx, float inv_mean, float scale,
                                                         double* loss, float* adj) {
    extern __shared__ double tsm[];
    double* cc = tsm;
    float* xn = reinterpret_cast<float*>(tsm + (2 * nt - 1));
    float* yn = xn + nt;
    __shared__ double red[33];
    __shared__ int s_arg;
    const int tr = blockIdx.x, tid = threadIdx.x, nl = 2 * nt - 1;
    auto bsum = [&](double v) {                      // block-wide sum, result to every thread
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w]; red[32] = t; }
        __syncthreads();
        return red[32];
    };
    auto bmax = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) { double t = red[0]; for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, red[w]); red[32] = t; }
        __syncthreads();
        return red[32];
    };
    // normalisation by max |.| (first index attaining it, like torch.max)
    double mx = 0.0, my = 0.0;
    for (int t = tid; t < nt; t += blockDim.x) {
        mx = fmax(mx, fabs((double)syn[(long long)t * ntr + tr]));
        my = fmax(my, fabs((double)obs[(long long)t * ntr + tr]));
    }
    mx = bmax(mx);
    my = bmax(my);
    if (tid == 0) s_arg = nt;
    __syncthreads();
    for (int t = tid; t < nt; t += blockDim.x)
        if (fabs((double)syn[(long long)t * ntr + tr]) == mx) atomicMin(&s_arg, t);
    const double dx = mx + 1e-16, dy = my + 1e-16;
    for (int t = tid; t < nt; t += blockDim.x) {
        xn[t] = (float)((double)syn[(long long)t * ntr + tr] / dx);
        yn[t] = (float)((double)obs[(long long)t * ntr + tr] / dy);
    }
    __syncthreads();
    // cc[k] = sum_j xn[k + j - (nt-1)] yn[j]
    double cmax = -1e300;
    for (int k = tid; k < nl; k += blockDim.x) {
        const int j0 = max(0, nt - 1 - k), j1 = min(nt, 2 * nt - 1 - k);
        double acc = 0.0;
        for (int j = j0; j < j1; ++j) acc += (double)xn[k + j - (nt - 1)] * (double)yn[j];
        cc[k] = acc;
        cmax = fmax(cmax, acc);
    }
    cmax = bmax(cmax);
    double se = 0.0, sk = 0.0;
    for (int k = tid; k < nl; k += blockDim.x) {
        const double e = exp(cc[k] - cmax);
        se += e;
        sk += e * (double)lagidx[k];
    }
    se = bsum(se);
    sk = bsum(sk);
    const double E = sk / se, tau = E - (double)(nt - 1);
    if (tid == 0 && loss) atomicAdd(loss, tau * tau * (double)inv_mean * (double)scale);
    if (!adj) return;
    // d loss / d cc[k] = (2 tau / N) p_k (idx_k - E)   (stored over cc)
    __syncthreads();
    for (int k = tid; k < nl; k += blockDim.x) {
        const double p = exp(cc[k] - cmax) / se;
        cc[k] = 2.0 * tau * (double)inv_mean * p * ((double)lagidx[k] - E);
    }
    __syncthreads();
    // d / d xn[s] = sum_j gcc[s - j + nt - 1] yn[j];   through xn = x / (max|x| + eps)
    double dot = 0.0;
    for (int s = tid; s < nt; s += blockDim.x) {
        double acc = 0.0;
        for (int j = 0; j < nt; ++j) acc += cc[s - j + nt - 1] * (double)yn[j];
        adj[(long long)s * ntr + tr] = (float)((double)scale * acc / dx);
        dot += acc * (double)syn[(long long)s * ntr + tr];
    }
    dot = bsum(dot);
    if (tid == 0 && s_arg < nt) {
        const long long i = (long long)s_arg * ntr + tr;
        const double sg = syn[i] > 0.f ? 1.0 : (syn[i] < 0.f ? -1.0 : 0.0);
        adj[i] -= (float)((double)scale * sg * dot / (dx * dx));
    }
}

constexpr int CT = 32;      // tile: 32 output samples x 32 traces, 32 taps per stage

// out[n][tr] = sum_m k[(n-m) mod nt] x[m][tr]      (transpose == false)
// out[n][tr] = sum_m k[(m-n) mod nt] x[m][tr]      (transpose == true)
__global__ void __launch_bounds__(CT * CT) circ_conv_kernel(const float* __restrict__ x, const float* __restrict__ ker,
                                                            float* __restrict__ out, int nt, int ntr, bool transpose) {
    __shared__ float xs[CT][CT + 1];
    __shared__ float ks[2 * CT];
    const int tt = threadIdx.x, tn = threadIdx.y;
    const int tr = blockIdx.x * CT + tt, n = blockIdx.y * CT + tn;
    const int n0 = blockIdx.y * CT;
    float acc = 0.f;
    for (int m0 = 0; m0 < nt; m0 += CT) {
        __syncthreads();
        const int m = m0 + tn;
        xs[tn][tt] = (m < nt && tr < ntr) ? __ldg(x + (long long)m * ntr + tr) : 0.f;
        // taps needed: d = (n - m) for n in [n0,n0+CT), m in [m0,m0+CT)  ->  d - (n0-m0) in (-CT, CT)
        const int tid = tn * CT + tt;
        if (tid < 2 * CT) {
            int d = (n0 - m0) + (tid - (CT - 1));
            if (transpose) d = -d;
            d %= nt;
            if (d < 0) d += nt;
            ks[tid] = __ldg(ker + d);
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < CT; ++mm) acc += ks[tn - mm + (CT - 1)] * xs[mm][tt];
    }
    if (n < nt && tr < ntr) out[(long long)n * ntr + tr] = acc;
}

// r = Es^2 - Eo^2; loss += 0.5 r^2; v = 2 r hs (in place over hs); base = 2 r xs -> adj
__global__ void __launch_bounds__(256) env_residual_kernel(const float* __restrict__ syn, const float* __restrict__ obs,
                                                           float* hs, const float* __restrict__ ho, long long n,
                                                           float scale, double* loss, float* adj) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float xs = syn[i], xo = obs[i], a = hs[i], b = ho[i];
        const float r = (xs * xs + a * a) - (xo * xo + b * b);
        acc += 0.5 * (double)r * (double)r;
        if (adj) adj[i] = 2.f * r * xs;
        hs[i] = 2.f * r * a;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * (double)scale);
}

__global__ void __launch_bounds__(256) env_combine_kernel(float* adj, const float* __restrict__ t, long long n, float scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        adj[i] = scale * (adj[i] + t[i]);
}

inline int nblocks(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace

extern "C" int st_misfit_l2(const float* syn, const float* obs, int64_t n, float scale, double* loss, float* adj, void* stream) {
    if (!syn || !obs || n < 0) { st_set_error("misfit_l2: bad arguments"); return ST_ERR_BADARG; }
    if (n == 0) return ST_OK;
    l2_kernel<<<nblocks(n), 256, 0, (cudaStream_t)stream>>>(syn, obs, n, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_l2: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_l1(const float* syn, const float* obs, int64_t n, float scale, double* loss, float* adj, void* stream) {
    if (!syn || !obs || n < 0) { st_set_error("misfit_l1: bad arguments"); return ST_ERR_BADARG; }
    if (n == 0) return ST_OK;
    l1_kernel<0><<<nblocks(n), 256, 0, (cudaStream_t)stream>>>(syn, obs, n, 0.f, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_l1: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_sml1(const float* syn, const float* obs, int64_t n, float beta, float scale, double* loss, float* adj,
                              void* stream) {
    if (!syn || !obs || n < 0 || !(beta > 0.f)) { st_set_error("misfit_sml1: bad arguments"); return ST_ERR_BADARG; }
    if (n == 0) return ST_OK;
    l1_kernel<1><<<nblocks(n), 256, 0, (cudaStream_t)stream>>>(syn, obs, n, beta, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_sml1: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_cc(const float* syn, const float* obs, int64_t n, float scale, double* loss, float* adj, void* stream) {
    if (!syn || !obs || n < 0) { st_set_error("misfit_cc: bad arguments"); return ST_ERR_BADARG; }
    if (n == 0) return ST_OK;
    l1_kernel<2><<<nblocks(n), 256, 0, (cudaStream_t)stream>>>(syn, obs, n, 0.f, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_cc: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_cs(const float* syn, const float* obs, int32_t nt, int32_t ntraces, int32_t mean_over, float scale,
                            double* loss, float* adj, void* stream) {
    if (!syn || !obs || nt <= 0 || ntraces < 0 || mean_over <= 0) { st_set_error("misfit_cs: bad arguments"); return ST_ERR_BADARG; }
    if (ntraces == 0) return ST_OK;
    cs_kernel<<<(ntraces + 255) / 256, 256, 0, (cudaStream_t)stream>>>(syn, obs, nt, ntraces, 1.f / (float)mean_over, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_cs: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_nim(const float* syn, const float* obs, int32_t nt, int32_t ntraces, float scale, double* loss, float* adj,
                             void* stream) {
    if (!syn || !obs || nt <= 0 || ntraces < 0) { st_set_error("misfit_nim: bad arguments"); return ST_ERR_BADARG; }
    if (ntraces == 0) return ST_OK;
    nim_kernel<0><<<(ntraces + 255) / 256, 256, 0, (cudaStream_t)stream>>>(syn, obs, nt, ntraces, nullptr, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_nim: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_integration(const float* syn, const float* obs, int32_t nt, int32_t ntraces, int64_t mean_over, float scale,
                                     double* loss, float* adj, void* stream) {
    if (!syn || !obs || nt <= 0 || ntraces < 0 || mean_over <= 0) { st_set_error("misfit_integration: bad arguments"); return ST_ERR_BADARG; }
    if (ntraces == 0) return ST_OK;
    nim_kernel<2><<<(ntraces + 255) / 256, 256, 0, (cudaStream_t)stream>>>(syn, obs, nt, ntraces, nullptr, scale / (float)mean_over, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_integration: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_traveltime(const float* syn, const float* obs, int32_t nt, int32_t ntraces, const float* lagidx,
                                    int32_t mean_over, float scale, double* loss, float* adj, void* stream) {
    if (!syn || !obs || !lagidx || nt <= 0 || ntraces < 0 || mean_over <= 0) { st_set_error("misfit_traveltime: bad arguments"); return ST_ERR_BADARG; }
    if (ntraces == 0) return ST_OK;
    const size_t smem = (size_t)(2 * nt - 1) * sizeof(double) + (size_t)2 * nt * sizeof(float);
    if (smem > 200 * 1024) { st_set_error("misfit_traveltime: nt = %d needs %zu bytes of shared memory (max 200 KB)", nt, smem); return ST_ERR_UNSUPPORTED; }
    if (st_set_max_smem<traveltime_kernel>(200 * 1024) != cudaSuccess) { st_set_error("misfit_traveltime: cannot raise the shared-memory limit"); return ST_ERR_CUDA; }
    traveltime_kernel<<<ntraces, 256, smem, (cudaStream_t)stream>>>(syn, obs, nt, ntraces, lagidx, 1.f / (float)mean_over, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_traveltime: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_misfit_w1d(const float* syn, const float* obs, int32_t nt, int32_t ntraces, const float* shift, float scale,
                             double* loss, float* adj, void* stream) {
    if (!syn || !obs || !shift || nt <= 0 || ntraces < 0) { st_set_error("misfit_w1d: bad arguments"); return ST_ERR_BADARG; }
    if (ntraces == 0) return ST_OK;
    nim_kernel<1><<<(ntraces + 255) / 256, 256, 0, (cudaStream_t)stream>>>(syn, obs, nt, ntraces, shift, scale, loss, adj);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_w1d: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int st_filtfilt(const float* x, float* y, double* work, int32_t nt, int32_t ntraces, const double* b, const double* a,
                           int32_t ncoef, void* stream) {
    if (!x || !y || !work || !b || !a || nt <= 0 || ntraces < 0) { st_set_error("filtfilt: bad arguments"); return ST_ERR_BADARG; }
    if (ncoef < 1 || ncoef > FILT_MAXC || a[0] == 0.0) { st_set_error("filtfilt: 1..%d coefficients with a[0] != 0 (got %d)", FILT_MAXC, ncoef); return ST_ERR_BADARG; }
    if (ntraces == 0) return ST_OK;
    FiltCoef c;
    for (int k = 0; k < FILT_MAXC; ++k) {
        c.b[k] = k < ncoef ? b[k] / a[0] : 0.0;
        c.a[k] = k < ncoef ? a[k] / a[0] : 0.0;
    }
    c.n = ncoef;
    filtfilt_kernel<<<(ntraces + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, y, work, nt, ntraces, c);
    if (cudaGetLastError() != cudaSuccess) { st_set_error("filtfilt: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}

extern "C" int64_t st_misfit_envelope_workspace(int32_t nt, int32_t ntraces) { return 3LL * nt * ntraces; }

extern "C" int st_misfit_envelope(const float* syn, const float* obs, int32_t nt, int32_t ntraces, const float* hker,
                                  float scale, double* loss, float* adj, float* workspace, void* stream) {
    if (!syn || !obs || !hker || !workspace || nt <= 0 || ntraces < 0) { st_set_error("misfit_envelope: bad arguments"); return ST_ERR_BADARG; }
    if (ntraces == 0) return ST_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)nt * ntraces;
    float* hs = workspace;
    float* ho = workspace + n;
    float* tmp = workspace + 2 * n;
    dim3 grid((ntraces + CT - 1) / CT, (nt + CT - 1) / CT), block(CT, CT);
    circ_conv_kernel<<<grid, block, 0, st>>>(syn, hker, hs, nt, ntraces, false);
    circ_conv_kernel<<<grid, block, 0, st>>>(obs, hker, ho, nt, ntraces, false);
    env_residual_kernel<<<nblocks(n), 256, 0, st>>>(syn, obs, hs, ho, n, scale, loss, adj);
    if (adj) {
        circ_conv_kernel<<<grid, block, 0, st>>>(hs, hker, tmp, nt, ntraces, true);
        env_combine_kernel<<<nblocks(n), 256, 0, st>>>(adj, tmp, n, scale);
    }
    if (cudaGetLastError() != cudaSuccess) { st_set_error("misfit_envelope: launch failed"); return ST_ERR_CUDA; }
    return ST_OK;
}
