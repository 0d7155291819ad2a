/* seistorch_b200 -- C ABI of the B200-native wave-propagation hot path.
 *
 * The reference (GeophyAI/seistorch) has no FFI: its plug-in surface is a Python
 * naming convention.  This header is the boundary underneath our
 * Python mirror of that surface; every entry point cites the reference code whose
 * work it replaces (file:line relative to the reference tree).  INTEGRATION.md shows
 * the ctypes binding (seistorch_b200/_lib.py) a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers owned by the
 *     caller (torch allocations); the library never allocates or frees;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), does
 *     not synchronise, keeps no mutable global state and is re-entrant;
 *   - return value 0 = ok, <0 = error (ST_ERR_*); st_last_error() gives the
 *     thread-local message;
 *   - fields are fp32, [channel][shot][row][pitch] with pitch (`ld`) a multiple of 4;
 *     index arrays are int32 (the Python layer converts the reference's int64).
 */
#ifndef SEISTORCH_B200_H
#define SEISTORCH_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_ABI_VERSION 1

/* equation-variant flags of the 2D second-order family (st_wave2d_*) */
#define ST_EQ_ISO   1   /* Cxx == Czz == r^2, taken from r                                  */
#define ST_EQ_PML   2   /* damped update, equations2d/acoustic.py:73-86                      */
#define ST_EQ_HABC  4   /* one-way blend,  equations2d/acoustic_habc.py:79-101,147-221       */
#define ST_EQ_XZ    8   /* mixed derivative, equations2d/tti_habc.py:40-57                   */
#define ST_EQ_G1   16   /* first-derivative terms, equations2d/acoustic_fwim_habc.py:38-60; with cxx != czz
                           (ST_EQ_HABC|ST_EQ_G1) the variable-density stencil of acoustic_rho_habc.py:32-57 */
#define ST_EQ_BORN 32   /* background+scattered pair, equations2d/acoustic_*_lsrtm_habc.py   */
/*   acoustic                 = ISO|PML          acoustic_habc            = ISO|HABC
 *   vti_habc2                = HABC             tti_habc                 = HABC|XZ
 *   acoustic_fwim_habc       = ISO|HABC|G1
 *   acoustic_vti_lsrtm_habc  = HABC|BORN        acoustic_tti_lsrtm_habc  = HABC|XZ|BORN   */

#define ST_ERR_BADARG      (-1)
#define ST_ERR_UNSUPPORTED (-2)
#define ST_ERR_CUDA        (-3)

int st_version(void);
const char* st_last_error(void);
/* The time loops of st_*_forward / st_*_adjoint (one launch per step, as the reference's python loop rnn.py:118-190 issues
 * its ~100 torch launches per step) are replayed as ONE CUDA graph when the identical call -- same problem struct bytes,
 * same step range, same stream -- is made again (csrc/st_graph.cuh; opt-in: SEISTORCH_B200_GRAPH=1).
 * out3 = {calls run as plain launch loops, calls captured into a graph, calls replayed from a graph} since process start. */
void st_graph_counters(int64_t* out3);
const char* st_graph_last_failure(void);      /* why the last capture attempt fell back to the plain loop ("" if none did) */

/* ------------------------------------------------------------------------------------
 * Acquisition shared by all propagators.
 *   sources  : replaces WaveSource.forward2d/3d (seistorch/source.py:47-70) and the one-hot
 *              mask built per call in rnn.py:160-166.  One entry per point source.
 *   receivers: replaces WaveProbe.forward2d/3d (seistorch/probe.py:42-48).  Receivers are
 *              sorted by (shot, row) and indexed by a CSR over rows; rec_orig maps back
 *              to the reference's concatenated order (rnn.py:51-73).
 * ---------------------------------------------------------------------------------- */
typedef struct st_acquisition {
    int32_t ns;                 /* number of point sources (all shots)                 */
    const int32_t* src_b;       /* [ns] shot index                                     */
    const int32_t* src_i0;      /* [ns] 3D: first tensor dim (x); 2D: unused           */
    const int32_t* src_i1;      /* [ns] row   (2D: z "y" index; 3D: z)                 */
    const int32_t* src_i2;      /* [ns] column (2D: x; 3D: y)                          */
    const float* amp;           /* [nt][ns] amplitude added at step i (wavelet sample) */
    float* gamp;                /* [nt][ns] out: d loss / d amp, or NULL               */
    int32_t src_fmask;          /* bit f: inject into field channel f                  */
    int32_t R;                  /* number of receivers (all shots)                     */
    const int32_t* row_start;   /* [nrows+1] CSR over rows: 2D row = b*nz+z; 3D row = (b*n0+i0)*n1+i1 */
    const int32_t* rec_col;     /* [R] column index (sorted order)                     */
    const int32_t* rec_orig;    /* [R] position in the reference's receiver order      */
    int32_t nchan;              /* receiver channels (len(receiver_type))              */
    int32_t chan_f[4];          /* field channel sampled by each receiver channel      */
    float* rec_out;             /* [nt][R][nchan] seismograms (forward), or NULL       */
    const float* rec_adj;       /* [nt][R][nchan] d loss / d seismogram (adjoint)      */
    int32_t row_lo, row_hi;     /* hint: every source and receiver has row index (2D: z, 3D: first dim) in
                                   [row_lo, row_hi]; blocks outside skip the source/receiver epilogue.
                                   row_lo > row_hi means "no hint" (all blocks run it)                   */
} st_acquisition;

/* ------------------------------------------------------------------------------------
 * 2D second-order family.  Replaces, for `nsteps` time steps per call,
 *   seistorch/rnn.py:178-205 (time loop), seistorch/cell.py:50-76 (step dispatch),
 *   the `_time_step` of equations2d/{acoustic,acoustic_habc,vti_habc2,tti_habc,
 *   acoustic_fwim_habc,acoustic_vti_lsrtm_habc,acoustic_tti_lsrtm_habc}.py,
 *   source.py:47-57 and probe.py:42-44;
 * st_wave2d_adjoint replaces torch autograd-through-time / checkpoint_new.py:146-217 with
 * the exact discrete adjoint (transposed stencil + imaging-condition accumulation).
 *
 * State S_j = field after step j (source added).  `u` holds `nslots` time slots of
 * [NF][B][nz][ld]; slot indices are taken modulo nslots (3 slots = rolling buffer for
 * pure forward modelling, K+2 slots = stored history of a K-step segment).
 * ---------------------------------------------------------------------------------- */
typedef struct st_wave2d_problem {
    int32_t flags;              /* ST_EQ_* */
    int32_t B, nz, nx, ld;      /* shots, padded grid, row pitch */
    int32_t bw, multiple;       /* absorbing width (50), free-surface flag */
    int32_t nt;
    float dt;                   /* only used by ST_EQ_PML (b*dt) */
    const float* coef[8];       /* r=vp*dt/h, b(damping d), cxx, czz, cxz, ax, az, m : [nz][ld] or NULL.
                                   ISO equations: cxx = r^2 (HABC) or r^2/(1+b dt) (PML), czz = (1-b dt)/(1+b dt) (PML) */
    float* taps;                /* optional workspace of st_wave2d_taps_floats() floats, filled by st_wave2d_prepare():
                                   precomputed absorbing-frame taps (ST_EQ_ISO|ST_EQ_HABC only); NULL = evaluate
                                   the one-way blend cell by cell */
    float* u;                   /* [nslots][NF][B][nz][ld] */
    int32_t nslots;
    float* lam;                 /* [3][NF][B][nz][ld] adjoint state, slot = i mod 3 (zero before the first adjoint call) */
    float* gacc;                /* [max(nchunk,B)][7][nz][ld] += coefficient gradients (r [one-way blend terms only],cxx,czz,
                                   cxz,ax,az,m), or NULL; the caller sums the planes */
    int32_t bchunk;             /* shots per block in the adjoint kernel; nchunk = ceil(B/bchunk) */
    st_acquisition acq;
} st_wave2d_problem;

/* size (floats) of the `taps` workspace, 0 if the flag set / grid does not use one */
int64_t st_wave2d_taps_floats(const st_wave2d_problem* p);
/* fill p->taps from the coefficient planes (once per call, before forward/adjoint) */
int st_wave2d_prepare(const st_wave2d_problem* p, void* stream);

/* 1 if st_wave2d_forward (adjoint = 0) / st_wave2d_adjoint (adjoint = 1) will run this problem on the TMA-staged
 * kernels (acoustic / acoustic_habc on grids with room for whole tile columns, enough work per launch; environment
 * SEISTORCH_B200_TMA=0/1 forces the register kernels / the TMA kernels wherever they apply), else 0.  Both kernel
 * families implement the same equations (reference: equations2d/acoustic.py:73-86, acoustic_habc.py:147-221);
 * results agree to fp32 rounding. */
int st_wave2d_uses_tma(const st_wave2d_problem* p, int32_t adjoint);

/* 1 if st_wave2d_forward(p, i0, nsteps, ...) will advance all `nsteps` time steps in ONE launch of the persistent
 * multi-timestep kernel (replaces the Python time loop rnn.py:178-205 for small grids: the acoustic PML equation on a
 * padded grid of at most 512 columns whose rows fit one thread-block cluster -- BASELINE configs[0] is the model case;
 * one cluster per shot keeps the wavefield on chip, halos travel through distributed shared memory, one cluster barrier
 * per time step), else 0 (one launch per time step).  Environment SEISTORCH_B200_PERSIST=0 turns it off.  Same
 * arithmetic as the per-step kernels: records agree bit for bit. */
int st_wave2d_uses_persist(const st_wave2d_problem* p, int32_t nsteps);
/* the same question for st_wave2d_adjoint(p, i_hi, nsteps, ...): the adjoint twin keeps the two cotangents and the
 * gradient accumulator of every cell in registers for the whole loop (checkpoint_new.py:146-217 / autograd through
 * rnn.py:178-205) and reads S_i from the wavefield history one step ahead of its use; gacc needs one plane set per shot. */
int st_wave2d_adjoint_uses_persist(const st_wave2d_problem* p, int32_t nsteps);

/* advance steps i0 .. i0+nsteps-1; S_{i0-2} lives in slot `slot0`, S_{i0-1} in slot0+1,
 * step i writes slot0+2+(i-i0).                                                         */
int st_wave2d_forward(const st_wave2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream);
/* compute Lam_i for i = i_hi .. i_hi-nsteps+1 (descending); S_{i_hi} lives in slot
 * `slot_hi`, S_{i} in slot_hi-(i_hi-i).  Accumulates the gradient of steps i+1.          */
int st_wave2d_adjoint(const st_wave2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream);

/* per-equation aliases (same argument meaning; they check p->flags) */
int st_acoustic2d_forward(const st_wave2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream);
int st_acoustic2d_adjoint(const st_wave2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream);
int st_acoustic2d_habc_forward(const st_wave2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream);
int st_acoustic2d_habc_adjoint(const st_wave2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream);
int st_qp2d_forward(const st_wave2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream);
int st_qp2d_adjoint(const st_wave2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream);
int st_fwim2d_forward(const st_wave2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream);
int st_fwim2d_adjoint(const st_wave2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream);

/* ------------------------------------------------------------------------------------
 * 2D elastic velocity-stress (Virieux).  Replaces equations2d/elastic.py:7-37 +
 * equations2d/utils.py:3-48 inside the same time loop; the adjoint replaces
 * checkpoint.py:146-230 / autograd.  State S_j = (vx,vz,txx,tzz,txz) after step j.
 * `u` holds nslots slots of [5][B][nz][ld].  Coefficient planes (all [nz][ld]):
 *   0: ca  = (1-c)/(1+c), c = 0.5*dt*d      1: cl2m = (lambda+2mu)*dt/h/(1+c)
 *   2: cl  = lambda*dt/h/(1+c)              3: cm   = mu*dt/h/(1+c)
 *   4: cb  = dt/(rho*h)/(1+c)
 * gacc: [nchunk][4][nz][ld] gradients w.r.t. (cl2m, cl, cm, cb).
 * ---------------------------------------------------------------------------------- */
typedef struct st_elastic2d_problem {
    int32_t B, nz, nx, ld, nt;
    const float* coef[5];
    float* u;                   /* [nslots][5][B][nz][ld] */
    int32_t nslots;
    float* lam;                 /* [2][5][B][nz][ld] adjoint state, slot = i mod 2 */
    float* gacc;                /* [nchunk][4][nz][ld] or NULL */
    int32_t bchunk;
    st_acquisition acq;
} st_elastic2d_problem;

/* advance steps i0..i0+nsteps-1; S_{i0-1} in slot0, step i writes slot0+1+(i-i0). */
int st_elastic2d_forward(const st_elastic2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream);
/* compute Lam_i for i = i_hi .. i_hi-nsteps+1; S_{i_hi+1} lives in slot `slot_hi1`
 * (S_i in slot_hi1-1-(i_hi-i)); accumulates the gradient of steps i+1.
 * Lam_{nt-1} (pure receiver term) is produced by calling with i_hi = nt-1 first: the
 * library treats Lam_nt as zero.                                                        */
int st_elastic2d_adjoint(const st_elastic2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi1, void* stream);

/* ------------------------------------------------------------------------------------
 * 3D acoustic (PML).  Replaces equations3d/acoustic.py:65-85, source.py:59-70,
 * probe.py:46-48 inside the time loop; adjoint replaces checkpoint_new.py + the
 * host-staged face copies of equations3d/utils.py:57-121.  Tensor layout (B, n0, n1, n2)
 * = the reference's (B, x, z, y), n2 fastest with pitch ld.
 * coef[0] = ciso = (vp*dt/h)^2/(1+b*dt), coef[1] = alpha = (1-b*dt)/(1+b*dt), each [n0][n1][ld].
 * The adjoint state `lam` holds the SCALED cotangent w = ciso * dL/dS (the damped acoustic
 * operator is self-adjoint up to that diagonal scaling); gacc accumulates ciso * dL/d ciso (the sum over time of
 * w_{i+1} * lap7(S_i)): 1/ciso does not depend on time, the caller divides the plane by ciso once after the last step.
 * ---------------------------------------------------------------------------------- */
typedef struct st_acoustic3d_problem {
    int32_t B, n0, n1, n2, ld, nt;
    float dt;
    const float* coef[2];
    float* u;                   /* [nslots][B][n0][n1][ld] */
    int32_t nslots;
    float* lam;                 /* [3][B][n0][n1][ld] */
    float* gacc;                /* [nchunk][n0][n1][ld] += ciso * d loss / d ciso, or NULL */
    int32_t bchunk;
    st_acquisition acq;
} st_acoustic3d_problem;

int st_acoustic3d_forward(const st_acoustic3d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream);
int st_acoustic3d_adjoint(const st_acoustic3d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream);

/* ------------------------------------------------------------------------------------
 * Misfits and their adjoint sources, on seismograms laid out [nt][R][nchan] (all shots
 * concatenated along R).  Replace seistorch/loss.py:409-421 (L2) and :178-216 (Envelope,
 * method 'square') with transform.py:24-66 (Hilbert transform, nfft = nt).
 *   loss   : [1] (double) += scale * sum over all samples (caller zeroes it)
 *   adj    : [nt][R][nchan] d loss / d syn (scaled by `scale`), or NULL
 * st_misfit_envelope needs `hker` [nt]: imaginary part of the analytic-signal impulse
 * response (ifft of the one-sided filter), computed once on the host, and a workspace of
 * 3*nt*ntraces floats.
 * ---------------------------------------------------------------------------------- */
int st_misfit_l2(const float* syn, const float* obs, int64_t n, float scale,
                 double* loss, float* adj, void* stream);
/* seistorch/loss.py:381-393 (L1Loss(reduction='sum') summed over shots); adj = scale * sign(syn - obs) */
int st_misfit_l1(const float* syn, const float* obs, int64_t n, float scale,
                 double* loss, float* adj, void* stream);
/* seistorch/loss.py:52-85 ("cs"): loss += scale * (1/mean_over) * sum over traces of 1 - cosine similarity along
 * time (eps 1e-10); mean_over = traces of one shot (the reference averages per shot, then sums the shots) */
int st_misfit_cs(const float* syn, const float* obs, int32_t nt, int32_t ntraces, int32_t mean_over, float scale,
                 double* loss, float* adj, void* stream);
/* seistorch/loss.py:463-501 ("nim", criterion 'l2', method 'square'): per trace the squared samples are
 * normalised by their sum over time and integrated (cumsum); loss += scale * sum of squared differences */
int st_misfit_nim(const float* syn, const float* obs, int32_t nt, int32_t ntraces, float scale,
                  double* loss, float* adj, void* stream);
/* seistorch/signal.py:49-101 (backend 'torch'): zero-phase IIR filter along time of records [nt][ntraces]:
 * causal pass then anti-causal pass, zero initial state, double precision inside (torchaudio filtfilt,
 * clamp=False).  b, a: HOST arrays of `ncoef` coefficients (scipy.signal.butter), work: nt*ntraces doubles.
 * The operator is self-adjoint: apply it to the cotangent for the backward pass. */
int st_filtfilt(const float* x, float* y, double* work, int32_t nt, int32_t ntraces, const double* b, const double* a,
                int32_t ncoef, void* stream);
/* seistorch/loss.py:900-955 ("w1d", method 'linear'): as nim with the samples shifted by -c instead of squared;
 * shift: DEVICE pointer to c = 1.1 * min(min syn, min obs, 0) of the shot (computed by the caller, no host sync) */
int st_misfit_w1d(const float* syn, const float* obs, int32_t nt, int32_t ntraces, const float* shift, float scale,
                  double* loss, float* adj, void* stream);
/* seistorch/loss.py:395-407 ("sml1": SmoothL1Loss(reduction='sum', beta)), :126-176 ("cc": minus the zero-lag
 * cross-correlation, - sum syn*obs), :366-379 ("integration": MSELoss(mean) of the cumulative sums along time;
 * mean_over = samples of one shot, nt * traces of the shot) */
int st_misfit_sml1(const float* syn, const float* obs, int64_t n, float beta, float scale, double* loss, float* adj, void* stream);
int st_misfit_cc(const float* syn, const float* obs, int64_t n, float scale, double* loss, float* adj, void* stream);
int st_misfit_integration(const float* syn, const float* obs, int32_t nt, int32_t ntraces, int64_t mean_over, float scale,
                          double* loss, float* adj, void* stream);
/* seistorch/loss.py:674-728 ("traveltime") with signal.py:203-208: loss += scale * (1/mean_over) * sum over traces
 * of tau^2, tau = soft-argmax lag of the cross-correlation of the max-normalised traces minus (nt-1).
 * lagidx: DEVICE array [2nt-1] = (2nt-2) * linspace(0, 1, 2nt-1) in fp32 (the reference's lag axis);
 * mean_over = traces of all shots (the reference takes one mean over shots x receivers x channels);
 * nt <= 8500 (24 nt bytes of shared memory per block). */
int st_misfit_traveltime(const float* syn, const float* obs, int32_t nt, int32_t ntraces, const float* lagidx,
                         int32_t mean_over, float scale, double* loss, float* adj, void* stream);
int st_misfit_envelope(const float* syn, const float* obs, int32_t nt, int32_t ntraces,
                       const float* hker, float scale, double* loss, float* adj,
                       float* workspace, void* stream);
int64_t st_misfit_envelope_workspace(int32_t nt, int32_t ntraces);

/* ---- gradient post-processing ------------------------------------------------------------------------------------
 * One pass of the reference's gradient smoothing (process.py:66-112 -> signal.py:247-319): truncated Gaussian of
 * `2*radius+1` normalised `weights` along `axis` (0 = z / rows, 1 = x / columns) of a contiguous [nz][nx] plane with
 * numpy 'reflect' boundaries.  in != out; radius < n along the axis. */
int st_gaussian_smooth2d(const float* in, float* out, int32_t nz, int32_t nx, const float* weights, int32_t radius,
                         int32_t axis, void* stream);

/* Source illumination (rnn.py:127-128,204-205: `precondition += sum_shots field[source_type]^2` every time step), read back
 * from the wavefield history of a gradient run: out[fs] += sum over `count` consecutive slots starting at `slot_first`
 * (mod nslots) and over the B shot planes of  u[slot * slot_stride + chan_offset + b * fs + cell]^2.
 * All strides in floats and multiples of 4; out is one pitched plane ([nz][ld] / [n0][n1][ld]). */
int st_illumination(const float* u, int64_t slot_stride, int32_t nslots, int32_t slot_first, int32_t count,
                    int64_t chan_offset, int32_t B, int64_t fs, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEISTORCH_B200_H */
