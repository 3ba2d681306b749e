// Peer exchange of the per-rank statistics blocks over NVLink, without NCCL on the data path.
//
// The reference has no multi-device path; the sweep shards by controller block, one process per GPU,
// and the only exchange is "every rank gets every rank's [15][S][C_local] statistics" (MCDataSim.get_metrics_dict's
// tensors for the whole controller set, mcsim.py:463-510, assembled from the shards).  Each rank owns one exchange
// buffer (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by its peers); a rank PUSHES its block straight
// into the column range [lo, lo + C_local) of every peer's [15][S][C_total] tensor with the copy engines
// (cudaMemcpy2DAsync on a side stream: no SM is taken from the evolution kernel of the next step, which the
// push overlaps), then raises its sequence number in every peer's flag array; a consumer waits, on its own stream,
// until all flags of its buffer have reached the sequence number it needs.  No pad, no concatenation, no
// collective launch.
#include <stdint.h>
#include <string.h>
#include "rc_common.cuh"

namespace rc {

constexpr int PEER_MAX_WORLD = 64;

struct PeerPtrs { unsigned long long* p[PEER_MAX_WORLD]; };

// one thread per peer: publish `seq` in slot `rank` of that peer's flag array (the copies that precede this
// kernel on the stream have completed, so the data is visible before the flag)
__global__ void peer_signal_kernel(PeerPtrs flags, int world, int rank, unsigned long long seq) {
    const int r = threadIdx.x;
    if (r < world) {
        __threadfence_system();
        *(volatile unsigned long long*)(flags.p[r] + rank) = seq;
    }
}

// one thread per rank slot: spin until the slot has reached `seq`; bounded (timeout_ns) so that a dead peer turns
// into an error code instead of a hung GPU
__global__ void peer_wait_kernel(const unsigned long long* flags, int world, unsigned long long seq,
                                 unsigned long long timeout_ns, unsigned long long* timed_out) {
    const int r = threadIdx.x;
    if (r >= world) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const volatile unsigned long long* f = flags + r;
    while (*f < seq) {
        __nanosleep(200);
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            if (timed_out) atomicAdd(timed_out, 1ull);
            break;
        }
    }
    __threadfence_system();
}

}  // namespace rc

using namespace rc;

extern "C" int rc_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle64_out) {
    if (!dev_ptr_out || !handle64_out) return set_error(RC_ERR_NULL, "rc_peer_alloc: null output");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    RC_CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 1));
    cudaError_t e = cudaMemset(p, 0, bytes ? bytes : 1);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return set_error(RC_ERR_CUDA, "rc_peer_alloc: %s", cudaGetErrorString(e));
    }
    memcpy(handle64_out, &h, 64);
    *dev_ptr_out = p;
    return RC_OK;
}

extern "C" int rc_peer_open(const unsigned char* handle64, void** dev_ptr_out) {
    if (!handle64 || !dev_ptr_out) return set_error(RC_ERR_NULL, "rc_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    RC_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return RC_OK;
}

extern "C" int rc_peer_close(void* dev_ptr) {
    if (dev_ptr) RC_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return RC_OK;
}

extern "C" int rc_peer_free(void* dev_ptr) {
    if (dev_ptr) RC_CUDA_TRY(cudaFree(dev_ptr));
    return RC_OK;
}

// local [rows][c_local] (dense) -> columns [col_offset, col_offset + c_local) of the [rows][c_total] tensor that
// starts at peer_tensors[r] on every rank r (this rank's own buffer included), one 2-D copy per destination.
extern "C" int rc_peer_push_columns(void* const* peer_tensors, int world, const double* local_dev, int64_t rows,
                                    int64_t c_local, int64_t c_total, int64_t col_offset, void* stream) {
    if (world < 1 || world > PEER_MAX_WORLD) return set_error(RC_ERR_BAD_ARG, "rc_peer_push_columns: world=%d", world);
    if (rows < 0 || c_local < 0 || col_offset < 0 || col_offset + c_local > c_total)
        return set_error(RC_ERR_BAD_ARG, "rc_peer_push_columns: bad block [%lld, %lld) of %lld", (long long)col_offset,
                         (long long)(col_offset + c_local), (long long)c_total);
    if (rows == 0 || c_local == 0) return RC_OK;
    if (!peer_tensors || !local_dev) return set_error(RC_ERR_NULL, "rc_peer_push_columns: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    for (int k = 0; k < world; ++k) {
        if (!peer_tensors[k]) return set_error(RC_ERR_NULL, "rc_peer_push_columns: null peer tensor %d", k);
        double* dst = (double*)peer_tensors[k] + col_offset;
        RC_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)c_total * 8, local_dev, (size_t)c_local * 8, (size_t)c_local * 8, (size_t)rows,
                                      cudaMemcpyDeviceToDevice, st));
    }
    return RC_OK;
}

extern "C" int rc_peer_signal(void* const* peer_flags, int world, int rank, uint64_t seq, void* stream) {
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
        return set_error(RC_ERR_BAD_ARG, "rc_peer_signal: world=%d rank=%d", world, rank);
    if (!peer_flags) return set_error(RC_ERR_NULL, "rc_peer_signal: null flags");
    PeerPtrs f = {};
    for (int k = 0; k < world; ++k) {
        if (!peer_flags[k]) return set_error(RC_ERR_NULL, "rc_peer_signal: null flag array %d", k);
        f.p[k] = (unsigned long long*)peer_flags[k];
    }
    peer_signal_kernel<<<1, PEER_MAX_WORLD, 0, (cudaStream_t)stream>>>(f, world, rank, seq); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}

extern "C" int rc_peer_wait(const void* local_flags, int world, uint64_t seq, double timeout_s, void* timed_out_dev,
                            void* stream) {
    if (world < 1 || world > PEER_MAX_WORLD) return set_error(RC_ERR_BAD_ARG, "rc_peer_wait: world=%d", world);
    if (!local_flags) return set_error(RC_ERR_NULL, "rc_peer_wait: null flags");
    if (!(timeout_s > 0.0)) timeout_s = 10.0;
    peer_wait_kernel<<<1, PEER_MAX_WORLD, 0, (cudaStream_t)stream>>>((const unsigned long long*)local_flags, world, seq,
                                                                   (unsigned long long)(timeout_s * 1e9),
                                                                   (unsigned long long*)timed_out_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}
