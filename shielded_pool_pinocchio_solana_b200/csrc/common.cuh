// common.cuh -- error plumbing and small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/g16b200.h"

namespace g16 {

// Thread-local last error, surfaced through g16_last_error() (include/g16b200.h).
void set_error(const std::string& msg);
const char* get_error();


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

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace g16
