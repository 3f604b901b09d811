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

#define DYNPY_CHECK(call, where)                                   \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return dynpy::fail_cuda(_e, where); \
    } while (0)

}  // namespace dynpy
