// Error plumbing shared by all translation units of libesr_b200.so.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/esr_b200.h"

namespace esr {

void set_error(const char* fmt, ...);

#define ESR_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ::esr::set_error(__VA_ARGS__);  \
            return ESR_ERR_INVALID;         \
        }                                   \
    } while (0)

#define ESR_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::esr::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return ESR_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return ESR_ERR_CUDA;
    }
    return ESR_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One-time set-up that CUDA keeps PER DEVICE (cudaFuncSetAttribute, attribute queries): run `body` the first time the
// calling thread's current device reaches this point.  The bodies are idempotent, so two host threads racing on the
// same device only repeat the work; a second GPU in the same process gets its own set-up (a process-wide static
// flag left its kernels without the opt-in shared-memory size and every launch there failed).
inline int current_device_ordinal() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & 63;
}
#define ESR_ONCE_PER_DEVICE(...)                                                        \
    do {                                                                                \
        static std::atomic<unsigned long long> once_mask__{0ull};                       \
        const int dev__ = ::esr::current_device_ordinal();                              \
        if (!((once_mask__.load(std::memory_order_acquire) >> dev__) & 1ull)) {         \
            __VA_ARGS__                                                                 \
            once_mask__.fetch_or(1ull << dev__, std::memory_order_release);             \
        }                                                                               \
    } while (0)

}  // namespace esr
