// Error plumbing shared by all translation units of libesr_b200.so.
#pragma once
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

}  // namespace esr
