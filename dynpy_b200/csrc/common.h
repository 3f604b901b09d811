// Shared host-side helpers of the C-ABI library (error text, per-device state).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dynpy_b200.h"

namespace dynpy {


void set_error(const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* where);

struct DeviceState {
    bool ready;
    int n_sm;
    double2* tw4096;  // W_4096^j
};
// Per-device state for the current device (allocates + fills the twiddle table on first use).
int get_device_state(DeviceState** out, cudaStream_t st);

// Optional per-kernel-class timing (CUDA events on the launching stream), used by bench.py for
// the live roofline numbers.  Classes: 0 dipolar word generator, 1 column pass, 2 row pass, 3 other.
enum ProfClass { PROF_WORDS = 0, PROF_COLS = 1, PROF_ROWS = 2, PROF_OTHER = 3, PROF_NCLS = 4 };
void prof_begin(int cls, cudaStream_t st);
void prof_end(int cls, cudaStream_t st);
struct ProfScope {
    int cls; cudaStream_t st;
    ProfScope(int c, cudaStream_t s) : cls(c), st(s) { prof_begin(cls, st); }
    ~ProfScope() { prof_end(cls, st); }
};

#define DYNPY_CHECK(call, where)                                   \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return dynpy::fail_cuda(_e, where); \
    } while (0)

}  // namespace dynpy
