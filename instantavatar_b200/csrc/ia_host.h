// ia_host.h -- host-side error plumbing shared by the translation units of libia_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/ia_b200.h"

char* ia_err_buf();  // thread-local, defined in ia_kernels.cu

inline int ia_set_err(int code, const char* fmt, const char* detail = "") {
    snprintf(ia_err_buf(), 512, fmt, detail);
    return code;
}
#define set_err ia_set_err
#define IA_CHECK_CUDA(expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return ia_set_err(IA_ECUDA, #expr ": %s", cudaGetErrorString(_e)); \
    } while (0)
#define IA_REQUIRE(cond)                                                        \
    do {                                                                        \
        if (!(cond)) return ia_set_err(IA_EINVAL, "invalid argument: %s", #cond); \
    } while (0)
