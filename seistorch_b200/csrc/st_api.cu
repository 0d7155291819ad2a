// C-ABI entry points (include/seistorch_b200.h): argument checks + host time loops.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/seistorch_b200.h"
#include "st_common.cuh"
#include "st_wave2d.cuh"
#include "st_wave2d_band.cuh"
#include "st_wave2d_persist.cuh"
#include "st_elastic2d.cuh"
#include "st_acoustic3d.cuh"
#include "st_graph.cuh"

static thread_local char g_err[512] = "";

void st_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define ST_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            st_set_error(__VA_ARGS__);   \
            return ST_ERR_BADARG;        \
        }                                \
    } while (0)

static inline int pmod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

extern "C" int st_version(void) { return ST_ABI_VERSION; }
extern "C" const char* st_last_error(void) { return g_err; }
// time loops run as plain launches / captured into a graph / replayed from one (st_graph.cuh), since process start
extern "C" void st_graph_counters(int64_t* out3) {
    std::lock_guard<std::mutex> lock(st_graph_mutex());
    out3[0] = st_graph_stats().plain; out3[1] = st_graph_stats().captured; out3[2] = st_graph_stats().replayed;
}
extern "C" const char* st_graph_last_failure(void) { return st_graph_failure(); }

static int check_acq(const st_acquisition& q, int nfields) {
    ST_REQUIRE(q.ns >= 0 && q.R >= 0, "acquisition: negative counts");
    if (q.ns > 0) ST_REQUIRE(q.src_b && q.src_i1 && q.src_i2, "acquisition: null source arrays");
    if (q.R > 0) {
        ST_REQUIRE(q.row_start && q.rec_col && q.rec_orig, "acquisition: null receiver arrays");
        ST_REQUIRE(q.nchan >= 1 && q.nchan <= 4, "acquisition: nchan must be 1..4 (got %d)", q.nchan);
        for (int c = 0; c < q.nchan; ++c)
            ST_REQUIRE(q.chan_f[c] >= 0 && q.chan_f[c] < nfields, "acquisition: receiver channel %d samples field %d of %d", c, q.chan_f[c], nfields);
    }
    return ST_OK;
}

// ===================================================================== wave2d
static int w2_check(const st_wave2d_problem* p) {
    ST_REQUIRE(p != nullptr, "wave2d: null problem");
    ST_REQUIRE(p->B > 0 && p->nz > 0 && p->nx > 0, "wave2d: bad shape B=%d nz=%d nx=%d", p->B, p->nz, p->nx);
    ST_REQUIRE(p->ld >= p->nx && p->ld % 4 == 0, "wave2d: row pitch %d must be >= nx and a multiple of 4", p->ld);
    ST_REQUIRE(p->coef[0] && p->coef[1], "wave2d: coef r and b are required");
    if (p->flags & ST_EQ_HABC) {
        ST_REQUIRE(p->bw > 0 && p->nx > 2 * p->bw && p->nz > (p->multiple ? 1 : 2) * p->bw,
                   "wave2d: HABC needs nx > 2*bw and nz > 2*bw (nz > bw with a free surface)");
    }
    ST_REQUIRE((long long)p->nz * p->ld < (1LL << 31), "wave2d: plane too large for 32-bit offsets");
    ST_REQUIRE(p->coef[2] != nullptr, "wave2d: cxx (ciso for ISO equations) required");
    if (!(p->flags & ST_EQ_ISO) || (p->flags & ST_EQ_PML)) ST_REQUIRE(p->coef[3] != nullptr, "wave2d: czz (alpha for PML) required");
    if (p->flags & ST_EQ_XZ) ST_REQUIRE(p->coef[4] != nullptr, "wave2d: cxz required");
    if (p->flags & ST_EQ_G1) ST_REQUIRE(p->coef[5] && p->coef[6], "wave2d: ax/az required");
    if (p->flags & ST_EQ_BORN) ST_REQUIRE(p->coef[7] != nullptr, "wave2d: m required");
    ST_REQUIRE(p->u && p->nslots >= 3, "wave2d: field buffer needs >= 3 slots");
    return check_acq(p->acq, (p->flags & ST_EQ_BORN) ? 2 : 1);
}

// precomputed frame taps: 9-tap HABC equations, and only when the band rectangles do not overlap
static bool w2_uses_taps(const st_wave2d_problem* p) {
    if (!st_flags_tapped(p->flags)) return false;
    W2Geom g{p->nz, p->nx, p->ld, p->bw, p->multiple};
    return st_band_ok(g, p->bw + 1);
}

static void w2_fill(const st_wave2d_problem* p, W2Args& a) {
    memset(&a, 0, sizeof(a));
    a.g.nz = p->nz; a.g.nx = p->nx; a.g.ld = p->ld; a.g.bw = p->bw; a.g.multiple = p->multiple;
    a.B = p->B; a.dt = p->dt;
    a.fs = (long long)p->nz * p->ld;
    a.cs = a.fs * p->B;
    for (int k = 0; k < 8; ++k) a.coef[k] = p->coef[k];
    a.taps = w2_uses_taps(p) ? p->taps : nullptr;
    a.row_lo = p->acq.row_lo; a.row_hi = p->acq.row_hi;
    if (a.row_lo > a.row_hi) { a.row_lo = 0; a.row_hi = p->nz - 1; }
    a.ns = p->acq.ns; a.src_b = p->acq.src_b; a.src_z = p->acq.src_i1; a.src_x = p->acq.src_i2;
    a.src_fmask = p->acq.src_fmask;
    a.row_start = p->acq.row_start; a.rec_x = p->acq.rec_col; a.rec_orig = p->acq.rec_orig;
    a.R = p->acq.R; a.nchan = p->acq.nchan;
    for (int c = 0; c < 4; ++c) a.chan_f[c] = p->acq.chan_f[c];
    a.bchunk = p->bchunk > 0 ? p->bchunk : 1;
}

// SEISTORCH_B200_TMA: "0" = never use the TMA interior path, "1" = whenever a rectangle exists, unset = auto
static int w2_tma_mode() {
    const char* e = getenv("SEISTORCH_B200_TMA");
    if (!e || !*e) return -1;
    return atoi(e) != 0 ? 1 : 0;
}

extern "C" int64_t st_wave2d_taps_floats(const st_wave2d_problem* p) {
    if (!p || !w2_uses_taps(p)) return 0;
    return (int64_t)ST_TAP_PLANES * p->nz * p->ld;
}

extern "C" int st_wave2d_prepare(const st_wave2d_problem* p, void* stream) {
    int rc = w2_check(p);
    if (rc) return rc;
    if (!w2_uses_taps(p) || p->taps == nullptr) return ST_OK;
    W2Args a;
    w2_fill(p, a);
    rc = st_wave2d_launch_prepare(p->flags, a, (cudaStream_t)stream);
    if (rc) st_set_error("wave2d_prepare: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

extern "C" int st_wave2d_uses_tma(const st_wave2d_problem* p, int32_t adjoint) {
    if (w2_check(p) != ST_OK) return 0;
    W2Args a;
    w2_fill(p, a);
    a.gacc = p->gacc;
    W2Tma tm;
    const int nf = (p->flags & ST_EQ_BORN) ? 2 : 1;
    if (adjoint && p->lam == nullptr) return 0;
    if (st_wave2d_tma_setup(p->flags, a, p->u, (long long)nf * p->B * p->nslots, p->lam, 3LL * nf * p->B, adjoint != 0, w2_tma_mode(), tm)) return 0;
    return tm.enabled;
}

// SEISTORCH_B200_PERSIST: "0" = never use the persistent multi-timestep kernel, unset / "1" = whenever it applies
static bool w2_persist_enabled() {
    const char* e = getenv("SEISTORCH_B200_PERSIST");
    return !(e && *e && atoi(e) == 0);
}

// Plans the persistent launch for steps [i0, i0+nsteps); false when the problem is outside its class.
static bool w2_persist_plan(const st_wave2d_problem* p, const W2Args& a, int i0, int nsteps, int slot0, W2Persist& pp,
                            bool adjoint = false) {
    if (!w2_persist_enabled() || nsteps < 4) return false;
    if (adjoint && p->lam == nullptr) return false;
    if (st_wave2d_persist_plan(p->flags, a, adjoint, pp) != ST_OK) return false;
    pp.u = p->u;
    pp.lam = adjoint ? p->lam : nullptr;
    pp.slot = a.cs;
    pp.nslots = p->nslots;
    pp.slot0 = pmod(slot0, p->nslots);
    pp.i0 = i0;
    pp.nsteps = nsteps;
    pp.history = p->nslots > 3;
    pp.probe = 1;
    if ((adjoint ? st_wave2d_persist_adjoint(a, pp, nullptr) : st_wave2d_persist_forward(a, pp, nullptr)) != ST_OK)
        return false;                                                           // no resident cluster of this shape
    pp.probe = 0;
    return true;
}

extern "C" int st_wave2d_uses_persist(const st_wave2d_problem* p, int32_t nsteps) {
    if (w2_check(p) != ST_OK) return 0;
    W2Args a;
    w2_fill(p, a);
    W2Persist pp;
    return w2_persist_plan(p, a, 0, nsteps, 0, pp) ? 1 : 0;
}

extern "C" int st_wave2d_adjoint_uses_persist(const st_wave2d_problem* p, int32_t nsteps) {
    if (w2_check(p) != ST_OK) return 0;
    W2Args a;
    w2_fill(p, a);
    a.gacc = p->gacc;
    W2Persist pp;
    return w2_persist_plan(p, a, 0, nsteps, 0, pp, true) ? 1 : 0;
}

extern "C" int st_wave2d_forward(const st_wave2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream) {
    int rc = w2_check(p);
    if (rc) return rc;
    ST_REQUIRE(i0 >= 0 && nsteps >= 0 && i0 + nsteps <= p->nt, "wave2d_forward: steps [%d,%d) outside [0,%d)", i0, i0 + nsteps, p->nt);
    if (p->acq.ns > 0) ST_REQUIRE(p->acq.amp != nullptr, "wave2d_forward: null amp");
    W2Args a;
    w2_fill(p, a);
    const int nf = (p->flags & ST_EQ_BORN) ? 2 : 1;
    const long long slot = a.cs * nf;
    cudaStream_t st = (cudaStream_t)stream;
    {
        W2Persist pp;
        if (w2_persist_plan(p, a, i0, nsteps, slot0, pp)) {
            a.amp = p->acq.amp ? p->acq.amp + (long long)i0 * p->acq.ns : nullptr;
            a.rec_out = (p->acq.rec_out && p->acq.R > 0) ? p->acq.rec_out + (long long)i0 * p->acq.R * p->acq.nchan : nullptr;
            rc = st_wave2d_persist_forward(a, pp, st);
            if (rc) { st_set_error("wave2d_forward: persistent launch failed: %s", cudaGetErrorString(cudaGetLastError())); return ST_ERR_CUDA; }
            return ST_OK;
        }
    }
    W2Tma tm;
    const int planes = nf * p->B;                // field planes per slot
    rc = st_wave2d_tma_setup(p->flags, a, p->u, (long long)planes * p->nslots, nullptr, 0, false, nsteps > 0 ? w2_tma_mode() : 0, tm);
    if (rc) return rc;
    auto loop = [&](cudaStream_t st) -> int {
    for (int k = 0; k < nsteps; ++k) {
        const int i = i0 + k;
        tm.pl_prev = planes * pmod(slot0 + k, p->nslots);
        tm.pl_cur = planes * pmod(slot0 + k + 1, p->nslots);
        a.prev = p->u + slot * pmod(slot0 + k, p->nslots);
        a.cur = p->u + slot * pmod(slot0 + k + 1, p->nslots);
        a.next = p->u + slot * pmod(slot0 + k + 2, p->nslots);
        a.amp = p->acq.amp ? p->acq.amp + (long long)i * p->acq.ns : nullptr;
        a.rec_out = (p->acq.rec_out && p->acq.R > 0) ? p->acq.rec_out + (long long)i * p->acq.R * p->acq.nchan : nullptr;
        const int rc = st_wave2d_launch_forward(p->flags, a, tm, st);
        if (rc) { if (rc == ST_ERR_CUDA) st_set_error("wave2d_forward: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    }
    return ST_OK;
    };
    return st_run_steps(p, sizeof(*p), 0 + 16 * (w2_tma_mode() + 1), i0, nsteps, slot0, st, loop);       // plain loop, or its CUDA-graph replay (st_graph.cuh)
}

extern "C" int st_wave2d_adjoint(const st_wave2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream) {
    int rc = w2_check(p);
    if (rc) return rc;
    ST_REQUIRE(p->lam != nullptr, "wave2d_adjoint: null adjoint state");
    ST_REQUIRE(nsteps >= 0 && i_hi < p->nt && i_hi - nsteps + 1 >= 0, "wave2d_adjoint: steps (%d..%d] outside [0,%d)", i_hi - nsteps, i_hi, p->nt);
    W2Args a;
    w2_fill(p, a);
    const int nf = (p->flags & ST_EQ_BORN) ? 2 : 1;
    const long long slot = a.cs * nf;
    cudaStream_t st = (cudaStream_t)stream;
    a.gacc = p->gacc;
    {
        W2Persist pp;
        if (w2_persist_plan(p, a, i_hi, nsteps, slot_hi, pp, true)) {
            a.rec_adj = (p->acq.rec_adj && p->acq.R > 0) ? p->acq.rec_adj + (long long)i_hi * p->acq.R * p->acq.nchan : nullptr;
            a.gamp = p->acq.gamp ? p->acq.gamp + (long long)i_hi * p->acq.ns : nullptr;
            rc = st_wave2d_persist_adjoint(a, pp, st);
            if (rc) { st_set_error("wave2d_adjoint: persistent launch failed: %s", cudaGetErrorString(cudaGetLastError())); return ST_ERR_CUDA; }
            return ST_OK;
        }
    }
    W2Tma tm;
    const int planes = nf * p->B;
    rc = st_wave2d_tma_setup(p->flags, a, p->u, (long long)planes * p->nslots, p->lam, 3LL * planes, true, nsteps > 0 ? w2_tma_mode() : 0, tm);
    if (rc) return rc;
    auto loop = [&](cudaStream_t st) -> int {
    for (int k = 0; k < nsteps; ++k) {
        const int i = i_hi - k;
        tm.pl_l1 = planes * pmod(i + 1, 3);
        tm.pl_l2 = planes * pmod(i + 2, 3);
        tm.pl_s1 = planes * pmod(slot_hi - k, p->nslots);
        tm.pl_s2 = planes * pmod(slot_hi - k - 1, p->nslots);
        a.lam0 = p->lam + slot * pmod(i, 3);
        a.lam1 = p->lam + slot * pmod(i + 1, 3);
        a.lam2 = p->lam + slot * pmod(i + 2, 3);
        a.s1 = p->u + slot * pmod(slot_hi - k, p->nslots);
        a.s2 = p->u + slot * pmod(slot_hi - k - 1, p->nslots);
        a.rec_adj = (p->acq.rec_adj && p->acq.R > 0) ? p->acq.rec_adj + (long long)i * p->acq.R * p->acq.nchan : nullptr;
        a.gamp = p->acq.gamp ? p->acq.gamp + (long long)i * p->acq.ns : nullptr;
        const int rc = st_wave2d_launch_adjoint(p->flags, a, tm, st);
        if (rc) { if (rc == ST_ERR_CUDA) st_set_error("wave2d_adjoint: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    }
    return ST_OK;
    };
    return st_run_steps(p, sizeof(*p), 1 + 16 * (w2_tma_mode() + 1), i_hi, nsteps, slot_hi, st, loop);   // (the kernel family is part of the key)
}

#define ST_ALIAS(NAME, COND, WHAT)                                                                                    \
    extern "C" int NAME##_forward(const st_wave2d_problem* p, int32_t i0, int32_t n, int32_t s, void* st) {         \
        ST_REQUIRE(p && (COND), #NAME ": flags do not describe " WHAT);                                               \
        return st_wave2d_forward(p, i0, n, s, st);                                                                    \
    }                                                                                                                 \
    extern "C" int NAME##_adjoint(const st_wave2d_problem* p, int32_t ih, int32_t n, int32_t s, void* st) {         \
        ST_REQUIRE(p && (COND), #NAME ": flags do not describe " WHAT);                                               \
        return st_wave2d_adjoint(p, ih, n, s, st);                                                                    \
    }
ST_ALIAS(st_acoustic2d, p->flags == (ST_EQ_ISO | ST_EQ_PML), "the PML acoustic equation")
ST_ALIAS(st_acoustic2d_habc, p->flags == (ST_EQ_ISO | ST_EQ_HABC), "the HABC acoustic equation")
ST_ALIAS(st_qp2d, (p->flags & ST_EQ_HABC) && !(p->flags & (ST_EQ_ISO | ST_EQ_G1 | ST_EQ_PML)), "a VTI/TTI qP equation")
ST_ALIAS(st_fwim2d, p->flags == (ST_EQ_ISO | ST_EQ_HABC | ST_EQ_G1), "the joint FWI-LSRTM equation")

// ===================================================================== elastic2d
static int e2_check(const st_elastic2d_problem* p) {
    ST_REQUIRE(p != nullptr, "elastic2d: null problem");
    ST_REQUIRE(p->B > 0 && p->nz > 0 && p->nx > 0, "elastic2d: bad shape");
    ST_REQUIRE(p->ld >= p->nx && p->ld % 4 == 0, "elastic2d: row pitch %d must be >= nx and a multiple of 4", p->ld);
    for (int k = 0; k < 5; ++k) ST_REQUIRE(p->coef[k] != nullptr, "elastic2d: coef %d is null", k);
    ST_REQUIRE(p->u && p->nslots >= 2, "elastic2d: field buffer needs >= 2 slots");
    return check_acq(p->acq, 5);
}

static void e2_fill(const st_elastic2d_problem* p, E2Args& a) {
    memset(&a, 0, sizeof(a));
    a.nz = p->nz; a.nx = p->nx; a.ld = p->ld; a.B = p->B;
    a.fs = (long long)p->nz * p->ld;
    a.cs = a.fs * p->B;
    for (int k = 0; k < 5; ++k) a.coef[k] = p->coef[k];
    a.ns = p->acq.ns; a.src_b = p->acq.src_b; a.src_z = p->acq.src_i1; a.src_x = p->acq.src_i2;
    a.src_fmask = p->acq.src_fmask;
    a.row_start = p->acq.row_start; a.rec_x = p->acq.rec_col; a.rec_orig = p->acq.rec_orig;
    a.R = p->acq.R; a.nchan = p->acq.nchan;
    for (int c = 0; c < 4; ++c) a.chan_f[c] = p->acq.chan_f[c];
    a.bchunk = p->bchunk > 0 ? p->bchunk : 1;
}

extern "C" int st_elastic2d_forward(const st_elastic2d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream) {
    int rc = e2_check(p);
    if (rc) return rc;
    ST_REQUIRE(i0 >= 0 && nsteps >= 0 && i0 + nsteps <= p->nt, "elastic2d_forward: steps outside [0,nt)");
    if (p->acq.ns > 0) ST_REQUIRE(p->acq.amp != nullptr, "elastic2d_forward: null amp");
    E2Args a;
    e2_fill(p, a);
    const long long slot = a.cs * 5;
    cudaStream_t st = (cudaStream_t)stream;
    auto loop = [&](cudaStream_t st) -> int {
    for (int k = 0; k < nsteps; ++k) {
        const int i = i0 + k;
        a.cur = p->u + slot * pmod(slot0 + k, p->nslots);
        a.next = p->u + slot * pmod(slot0 + k + 1, p->nslots);
        a.amp = p->acq.amp ? p->acq.amp + (long long)i * p->acq.ns : nullptr;
        a.rec_out = (p->acq.rec_out && p->acq.R > 0) ? p->acq.rec_out + (long long)i * p->acq.R * p->acq.nchan : nullptr;
        const int rc = st_elastic2d_launch_forward(a, st);
        if (rc) { st_set_error("elastic2d_forward: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    }
    return ST_OK;
    };
    return st_run_steps(p, sizeof(*p), 2, i0, nsteps, slot0, st, loop);
}

extern "C" int st_elastic2d_adjoint(const st_elastic2d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi1, void* stream) {
    int rc = e2_check(p);
    if (rc) return rc;
    ST_REQUIRE(p->lam != nullptr, "elastic2d_adjoint: null adjoint state");
    ST_REQUIRE(nsteps >= 0 && i_hi < p->nt && i_hi - nsteps + 1 >= -1, "elastic2d_adjoint: steps outside range");
    E2Args a;
    e2_fill(p, a);
    const long long slot = a.cs * 5;
    cudaStream_t st = (cudaStream_t)stream;
    a.gacc = p->gacc;
    auto loop = [&](cudaStream_t st) -> int {
    for (int k = 0; k < nsteps; ++k) {
        const int i = i_hi - k;
        a.lam0 = p->lam + slot * pmod(i, 2);
        a.lam1 = (i + 1 < p->nt) ? p->lam + slot * pmod(i + 1, 2) : nullptr;   // Lam_nt == 0
        a.s1 = p->u + slot * pmod(slot_hi1 - k, p->nslots);          // S_{i+1}
        a.s0 = p->u + slot * pmod(slot_hi1 - k - 1, p->nslots);      // S_i
        a.rec_adj = (p->acq.rec_adj && p->acq.R > 0 && i >= 0) ? p->acq.rec_adj + (long long)i * p->acq.R * p->acq.nchan : nullptr;
        a.amp = (p->acq.amp && i + 1 < p->nt) ? p->acq.amp + (long long)(i + 1) * p->acq.ns : nullptr;   // injected into S_{i+1}
        a.gamp = (p->acq.gamp && i >= 0) ? p->acq.gamp + (long long)i * p->acq.ns : nullptr;
        const int rc = st_elastic2d_launch_adjoint(a, st);
        if (rc) { st_set_error("elastic2d_adjoint: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    }
    return ST_OK;
    };
    return st_run_steps(p, sizeof(*p), 3, i_hi, nsteps, slot_hi1, st, loop);
}

// ===================================================================== acoustic3d
static int a3_check(const st_acoustic3d_problem* p) {
    ST_REQUIRE(p != nullptr, "acoustic3d: null problem");
    ST_REQUIRE(p->B > 0 && p->n0 > 0 && p->n1 > 0 && p->n2 > 0, "acoustic3d: bad shape");
    ST_REQUIRE(p->ld >= p->n2 && p->ld % 4 == 0, "acoustic3d: row pitch %d must be >= n2 and a multiple of 4", p->ld);
    ST_REQUIRE(p->coef[0] && p->coef[1], "acoustic3d: coef r and b are required");
    ST_REQUIRE(p->u && p->nslots >= 3, "acoustic3d: field buffer needs >= 3 slots");
    if (p->acq.ns > 0) ST_REQUIRE(p->acq.src_i0 != nullptr, "acoustic3d: null src_i0");
    return check_acq(p->acq, 1);
}

static void a3_fill(const st_acoustic3d_problem* p, A3Args& a) {
    memset(&a, 0, sizeof(a));
    a.n0 = p->n0; a.n1 = p->n1; a.n2 = p->n2; a.ld = p->ld; a.B = p->B; a.dt = p->dt;
    a.ps = (long long)p->n1 * p->ld;
    a.fs = a.ps * p->n0;
    a.r = p->coef[0]; a.b = p->coef[1];
    a.ns = p->acq.ns; a.src_b = p->acq.src_b; a.src_i0 = p->acq.src_i0; a.src_i1 = p->acq.src_i1; a.src_i2 = p->acq.src_i2;
    a.row_start = p->acq.row_start; a.rec_col = p->acq.rec_col; a.rec_orig = p->acq.rec_orig;
    a.R = p->acq.R;
    a.bchunk = p->bchunk > 0 ? p->bchunk : 1;
}

extern "C" int st_acoustic3d_forward(const st_acoustic3d_problem* p, int32_t i0, int32_t nsteps, int32_t slot0, void* stream) {
    int rc = a3_check(p);
    if (rc) return rc;
    ST_REQUIRE(i0 >= 0 && nsteps >= 0 && i0 + nsteps <= p->nt, "acoustic3d_forward: steps outside [0,nt)");
    if (p->acq.ns > 0) ST_REQUIRE(p->acq.amp != nullptr, "acoustic3d_forward: null amp");
    ST_REQUIRE(p->acq.R == 0 || p->acq.nchan == 1, "acoustic3d: one receiver channel");
    A3Args a;
    a3_fill(p, a);
    const long long slot = a.fs * p->B;
    cudaStream_t st = (cudaStream_t)stream;
    for (int k = 0; k < nsteps; ++k) {
        const int i = i0 + k;
        a.prev = p->u + slot * pmod(slot0 + k, p->nslots);
        a.cur = p->u + slot * pmod(slot0 + k + 1, p->nslots);
        a.next = p->u + slot * pmod(slot0 + k + 2, p->nslots);
        a.amp = p->acq.amp ? p->acq.amp + (long long)i * p->acq.ns : nullptr;
        a.rec_out = (p->acq.rec_out && p->acq.R > 0) ? p->acq.rec_out + (long long)i * p->acq.R : nullptr;
        rc = st_acoustic3d_launch_forward(a, st);
        if (rc) { st_set_error("acoustic3d_forward: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    }
    return ST_OK;
}

extern "C" int st_acoustic3d_adjoint(const st_acoustic3d_problem* p, int32_t i_hi, int32_t nsteps, int32_t slot_hi, void* stream) {
    int rc = a3_check(p);
    if (rc) return rc;
    ST_REQUIRE(p->lam != nullptr, "acoustic3d_adjoint: null adjoint state");
    ST_REQUIRE(nsteps >= 0 && i_hi < p->nt && i_hi - nsteps + 1 >= 0, "acoustic3d_adjoint: steps outside [0,nt)");
    A3Args a;
    a3_fill(p, a);
    const long long slot = a.fs * p->B;
    cudaStream_t st = (cudaStream_t)stream;
    a.gacc = p->gacc;
    for (int k = 0; k < nsteps; ++k) {
        const int i = i_hi - k;
        a.lam0 = p->lam + slot * pmod(i, 3);
        a.lam1 = p->lam + slot * pmod(i + 1, 3);
        a.lam2 = p->lam + slot * pmod(i + 2, 3);
        a.s1 = p->u + slot * pmod(slot_hi - k, p->nslots);
        a.rec_adj = (p->acq.rec_adj && p->acq.R > 0) ? p->acq.rec_adj + (long long)i * p->acq.R : nullptr;
        a.gamp = p->acq.gamp ? p->acq.gamp + (long long)i * p->acq.ns : nullptr;
        rc = st_acoustic3d_launch_adjoint(a, st);
        if (rc) { st_set_error("acoustic3d_adjoint: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    }
    return ST_OK;
}
