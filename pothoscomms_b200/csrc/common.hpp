// Shared host-side helpers for libb200comms.so (error text, CUDA status checks).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>

#include "../../include/b200comms.h"

namespace b200c {

void set_error(const char *fmt, ...);

inline bool dtype_valid(int dt) { return dt >= 0 && dt <= B200C_CI64; }
inline bool dtype_is_complex(int dt) { return (dt & 1) != 0; }
inline bool dtype_is_float(int dt) { return (dt >> 1) < 2; }
inline size_t dtype_scalar_bytes(int dt)
{
    switch (dt >> 1) { case 0: return 4; case 1: return 8; case 2: return 1; case 3: return 2; case 4: return 4; default: return 8; }
}
inline size_t dtype_bytes(int dt) { return dtype_scalar_bytes(dt) * (dtype_is_complex(dt) ? 2 : 1); }

// RAII device switch
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define B200C_CUDA_TRY(expr)                                                                         \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            ::b200c::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            (void)cudaGetLastError();                                                                \
            return B200C_ERR_CUDA;                                                                   \
        }                                                                                            \
    } while (0)

} // namespace b200c
