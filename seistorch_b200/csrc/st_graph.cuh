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
// graphs are kept per process (least recently used out).  SEISTORCH_B200_GRAPH=0 turns the mechanism off; a stream that is
// already being captured by the caller is left alone (the launches simply join the caller's graph).
#pragma once
#include <cstdint>
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

inline bool st_graph_enabled() {
    static const bool on = !(getenv("SEISTORCH_B200_GRAPH") && atoi(getenv("SEISTORCH_B200_GRAPH")) == 0);
    return on;
}

// runs `loop` (which issues the launches of steps on stream `st` and returns ST_OK or an error code), as a graph replay
// when this exact call has been seen before
template <class Loop>
int st_run_steps(const void* prob, size_t prob_len, int which, int a0, int nsteps, int a2, cudaStream_t st, Loop&& loop) {
    if (!st_graph_enabled() || nsteps < ST_GRAPH_MIN_STEPS) return loop();
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return loop();
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
        return loop();
    }
    e->tick = ++tick;
    if (e->failed) { ++st_graph_stats().plain; return loop(); }
    if (e->exec == nullptr) {
        // second sighting: capture the loop's own launches
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            (void)cudaGetLastError();
            e->failed = true;
            ++st_graph_stats().plain;
            return loop();
        }
        const int rc = loop();
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(st, &g);
        if (rc != ST_OK || ce != cudaSuccess || g == nullptr) {
            (void)cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            e->failed = true;
            if (rc != ST_OK) return rc;                 // the loop's own error (bad argument ...): nothing was launched
            ++st_graph_stats().plain;
            return loop();
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ci = cudaGraphInstantiate(&exec, g, 0);
        cudaGraphDestroy(g);
        if (ci != cudaSuccess || exec == nullptr) {
            (void)cudaGetLastError();
            e->failed = true;
            ++st_graph_stats().plain;
            return loop();
        }
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
