// Shared host-side helpers of the C-ABI layer: status codes, thread-local error text.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/robchar_b200.h"

namespace rc {

char* last_error_buffer();            // thread local, 512 bytes
int set_error(int code, const char* fmt, ...);
int device_sm_count();                // SMs of the current device (cached per device)

#define RC_CUDA_TRY(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return rc::set_error(RC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                          \
    } while (0)

}  // namespace rc
