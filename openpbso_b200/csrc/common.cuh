// Shared host-side plumbing for libpbso_b200.so: status codes, error text, CUDA checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <string>
#include "../../include/pbso_b200.h"

namespace pbso {

std::string& last_error();                       // thread-local storage (common.cu)
int set_error(int code, const char* fmt, ...);   // formats into last_error(), returns code
int check_device();                              // PBSO_OK or PBSO_ERR_NO_DEVICE

#define PBSO_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return pbso::set_error(PBSO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                                   cudaGetErrorString(_e), __FILE__, __LINE__);           \
    } while (0)

#define PBSO_REQUIRE(cond, code, msg)                                                     \
    do {                                                                                  \
        if (!(cond)) return pbso::set_error(code, "%s: %s", __func__, msg);               \
    } while (0)

// Binds the calling thread to a handle's device for the scope of one API call.
struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace pbso
