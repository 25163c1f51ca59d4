// common.cuh -- error plumbing and small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/g16b200.h"

namespace g16 {

// Thread-local last error, surfaced through g16_last_error() (include/g16b200.h).
void set_error(const std::string& msg);
const char* get_error();
// G16_TRACE=1: host-side timestamps (us since the first call, thread id) of pipeline stages on stderr
void trace(const char* tag, long a = 0);


#define G16_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _buf[512];                                                                  \
            snprintf(_buf, sizeof _buf, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                     cudaGetErrorString(_e));                                                \
            ::g16::set_error(_buf);                                                          \
            return ::G16_E_CUDA;                                                        \
        }                                                                                    \
    } while (0)

#define G16_TRY(expr)                  \
    do {                               \
        int _rc = (expr);              \
        if (_rc != ::G16_OK) return _rc; \
    } while (0)

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline line).
// Events are pooled and only read back in read(): nothing synchronises inside the timed region.
struct KernelProfiler {
    struct Rec {
        cudaEvent_t a, b;
        int tag;
        double units;
    };
    bool enabled = false;
    std::vector<Rec> recs;
    size_t used = 0;
    void begin(int tag, double units, cudaStream_t st) {
        if (!enabled) return;
        if (used == recs.size()) {
            Rec r;
            cudaEventCreate(&r.a);
            cudaEventCreate(&r.b);
            recs.push_back(r);
        }
        recs[used].tag = tag;
        recs[used].units = units;
        cudaEventRecord(recs[used].a, st);
    }
    void end(cudaStream_t st) {
        if (!enabled) return;
        cudaEventRecord(recs[used].b, st);
        used++;
    }
    // G16_TIMELINE=1: named events on whatever stream, dumped (ms since the first mark) by timeline_dump();
    // unlike `enabled` this leaves the stream overlap of the pipeline untouched
    struct Mark {
        cudaEvent_t ev;
        const char* a;
        const char* b;
    };
    std::vector<Mark> marks;
    bool timeline = getenv("G16_TIMELINE") && atoi(getenv("G16_TIMELINE")) != 0;
    void mark(const char* a, const char* b, cudaStream_t st) {
        if (!timeline) return;
        Mark m;
        cudaEventCreate(&m.ev);
        m.a = a;
        m.b = b;
        cudaEventRecord(m.ev, st);
        marks.push_back(m);
    }
    void timeline_dump() {
        if (marks.empty()) return;
        cudaDeviceSynchronize();
        for (auto& m : marks) {
            float t = 0;
            cudaEventElapsedTime(&t, marks[0].ev, m.ev);
            fprintf(stderr, "[timeline %8.3f ms] %s %s\n", t, m.a, m.b);
        }
        for (auto& m : marks) cudaEventDestroy(m.ev);
        marks.clear();
    }
    // sums per tag (0..7): milliseconds, launches, units
    void read(double ms[8], double launches[8], double units[8]) {
        for (int i = 0; i < 8; i++) ms[i] = launches[i] = units[i] = 0;
        for (size_t i = 0; i < used; i++) {
            cudaEventSynchronize(recs[i].b);
            float t = 0;
            cudaEventElapsedTime(&t, recs[i].a, recs[i].b);
            int g = recs[i].tag & 7;
            ms[g] += t;
            launches[g] += 1;
            units[g] += recs[i].units;
        }
        used = 0;
    }
    ~KernelProfiler() {
        for (auto& r : recs) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
    }
};
enum { PROF_MSM_ACC_G1 = 0, PROF_MSM_ACC_G2 = 1, PROF_NTT = 2, PROF_MSM_OTHER = 3, PROF_SPMV = 4 };

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace g16
