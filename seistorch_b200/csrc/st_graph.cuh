// Host time loops as CUDA graphs.
//
// One C-ABI call (st_*_forward / st_*_adjoint) launches one kernel per time step: 2000 launches for a BASELINE shot batch,
// and the launching thread stays ~1000 launches ahead of the GPU at best (the depth of the launch queue: 40-70 ms of work).
// Any longer hiccup of that thread -- measured on B200 boxes: sporadic 20-150 ms, with or without nvidia-smi polling --
// drains the queue and the GPU idles.  An inversion repeats the SAME call every iteration (same buffers, same step range),
// so the loop is captured once (stream capture of the very launches the loop issues, programmatic-dependent-launch edges
// included) and replayed with a single cudaGraphLaunch: the whole phase is queued at once and the host is out of the way.
//
// Rules: a call is identified by the bytes of its problem struct (every pointer, size and flag the kernels see) plus the
// step range; the first sighting runs the plain loop (one-off calls pay nothing), the second captures + instantiates, later
// ones replay.  The graph bakes pointers, never data, so equal keys mean equal launches.  At most ST_GRAPH_ENTRIES executable
// graphs are kept per process (least recently used out).  A stream that is already being captured by the caller is left
// alone (the launches simply join the caller's graph).
//
// Measured (B200, cfg2: 2 x 8 shots, 2000 steps, 30 gradient steps per run): a replayed step takes 427-432 ms against 433.5 ms
// with the plain loops (the host is out of the way, the launch gaps of the plain loop disappear), results bit-identical.  But
// (1) the sporadic long steps remain -- they also hit steps that were replayed, so they are not (only) launch-side -- and
// (2) torch hands out new addresses for the record / cotangent / gradient buffers often enough that 15 of 120 calls were
// re-captured at ~40 ms each.  Until those buffers are engine-owned like the history, the mechanism is therefore OPT-IN:
// SEISTORCH_B200_GRAPH=1 turns it on.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include <cuda_runtime.h>

#include "st_common.cuh"

#ifndef ST_GRAPH_ENTRIES
#define ST_GRAPH_ENTRIES 24
#endif
#ifndef ST_GRAPH_MIN_STEPS
#define ST_GRAPH_MIN_STEPS 64
#endif

struct StGraphEntry {
    std::vector<unsigned char> key;
    cudaGraphExec_t exec = nullptr;
    bool failed = false;             // capture or instantiation failed once: keep to the plain loop
    unsigned long long tick = 0;
};

struct StGraphStats { long long plain = 0, captured = 0, replayed = 0; };

inline std::mutex& st_graph_mutex() { static std::mutex m; return m; }
inline std::vector<StGraphEntry>& st_graph_entries() { static std::vector<StGraphEntry> v; return v; }
inline StGraphStats& st_graph_stats() { static StGraphStats s; return s; }

inline bool st_graph_enabled() {              // opt-in: SEISTORCH_B200_GRAPH=1 (see "Measured" above)
    const char* e = getenv("SEISTORCH_B200_GRAPH");
    return e && *e && atoi(e) != 0;
}
inline char* st_graph_failure() { static char msg[256] = ""; return msg; }     // why the last capture was abandoned
inline void st_graph_fail(StGraphEntry* e, const char* what, cudaError_t ce) {
    snprintf(st_graph_failure(), 256, "%s: %s", what, cudaGetErrorString(ce));
    (void)cudaGetLastError();
    e->failed = true;
    ++st_graph_stats().plain;
}
// torch's default stream is the legacy stream, which cannot be captured: the loop is captured on a private stream of the
// calling thread and the executable graph is launched on the caller's stream
inline cudaStream_t st_graph_capture_stream() {
    thread_local cudaStream_t s = nullptr;
    thread_local int dev_of = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (s == nullptr || dev_of != dev) {
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); s = nullptr; }
        dev_of = dev;
    }
    return s;
}

// runs `loop(stream)` (which issues the launches of the steps on that stream and returns ST_OK or an error code) on `st`, as
// a graph replay when this exact call has been seen before
template <class Loop>
int st_run_steps(const void* prob, size_t prob_len, int which, int a0, int nsteps, int a2, cudaStream_t st, Loop&& loop) {
    if (!st_graph_enabled() || nsteps < ST_GRAPH_MIN_STEPS) return loop(st);
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return loop(st);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<unsigned char> key(prob_len + 5 * sizeof(int) + sizeof(void*));
    {
        unsigned char* k = key.data();
        memcpy(k, prob, prob_len); k += prob_len;
        const int tail[5] = {which, a0, nsteps, a2, dev};
        memcpy(k, tail, sizeof(tail)); k += sizeof(tail);
        memcpy(k, &st, sizeof(void*));
    }
    std::lock_guard<std::mutex> lock(st_graph_mutex());      // (calls of one process are serialised here: they share the GPU anyway)
    static unsigned long long tick = 0;
    auto& entries = st_graph_entries();
    StGraphEntry* e = nullptr;
    for (auto& c : entries)
        if (c.key == key) { e = &c; break; }
    if (e == nullptr) {
        // first sighting: remember the call, run the plain loop
        if ((int)entries.size() >= ST_GRAPH_ENTRIES) {
            size_t victim = 0;
            for (size_t i = 1; i < entries.size(); ++i)
                if (entries[i].tick < entries[victim].tick) victim = i;
            if (entries[victim].exec) cudaGraphExecDestroy(entries[victim].exec);
            entries.erase(entries.begin() + victim);
        }
        StGraphEntry n;
        n.key = std::move(key);
        n.tick = ++tick;
        entries.push_back(std::move(n));
        ++st_graph_stats().plain;
        return loop(st);
    }
    e->tick = ++tick;
    if (e->failed) { ++st_graph_stats().plain; return loop(st); }
    if (e->exec == nullptr) {
        // second sighting: capture the loop's own launches
        cudaStream_t cap = st_graph_capture_stream();
        cudaError_t ce = cap ? cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) : cudaErrorUnknown;
        if (ce != cudaSuccess) { st_graph_fail(e, "begin capture", ce); return loop(st); }
        const int rc = loop(cap);
        cudaGraph_t g = nullptr;
        ce = cudaStreamEndCapture(cap, &g);
        if (rc != ST_OK || ce != cudaSuccess || g == nullptr) {
            if (g) cudaGraphDestroy(g);
            st_graph_fail(e, rc != ST_OK ? "launch inside capture" : "end capture", ce);
            if (rc != ST_OK) { --st_graph_stats().plain; return rc; }     // the loop's own error: nothing was launched
            return loop(st);
        }
        cudaGraphExec_t exec = nullptr;
        ce = cudaGraphInstantiate(&exec, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess || exec == nullptr) { st_graph_fail(e, "instantiate", ce); return loop(st); }
        e->exec = exec;
        ++st_graph_stats().captured;
    } else {
        ++st_graph_stats().replayed;
    }
    if (cudaGraphLaunch(e->exec, st) != cudaSuccess) {
        st_set_error("graph replay failed: %s", cudaGetErrorString(cudaGetLastError()));
        return ST_ERR_CUDA;
    }
    return ST_OK;
}
