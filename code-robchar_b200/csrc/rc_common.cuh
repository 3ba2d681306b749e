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
void note_launch();                   // one kernel of this library was launched (process-wide counter, rc_launch_count)
// rc_fidelity.cu: shared implementation of rc_fidelity_mc, also launched per sigma chunk by the host sweep
int fidelity_mc_impl(const char* who, const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                     const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed, int64_t c_offset,
                     int64_t b_offset, const double* replay_dev, double* fids_dev, unsigned long long* nonconv_dev,
                     int s_offset, cudaStream_t st, double* amps_dev = nullptr);
// rc_stats.cu: sort-free statistics of segments [0, nseg_chunk) of fids_dev, written to columns
// stats_dev[row * stat_stride + seg] (rc_stats_unsorted; per sigma chunk in the host sweep)
int stats_unsorted_impl(const double* fids_dev, int64_t nseg_chunk, int64_t B, double dkw_eps, double* stats_dev,
                        int64_t stat_stride, unsigned long long* illegal_dev, cudaStream_t st);

#define RC_CUDA_TRY(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return rc::set_error(RC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                          \
    } while (0)

}  // namespace rc
