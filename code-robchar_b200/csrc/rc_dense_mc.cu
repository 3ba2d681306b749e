// Batched Monte-Carlo sweep for Hamiltonians that are NOT tridiagonal: topo="ring" (noise_model.py:83-85,
// qnewton.py:145-147: the chain closed by unit couplings between sites 0 and N-1) under the structured
// perturbation (noise_model.py:122-147 / qnewton.py:366-379, which never touches the corner elements).
// The reference evaluates these one at a time with scipy.linalg.expm (noise_model.py:98-109); here every
// (sigma, controller, draw) of a sweep gets its dense -iTH built on the device (same Philox counters / same
// replay layout as rc_fidelity_mc), the whole tile goes through the Pade-13 expm kernel (rc_expm.cu, one CTA
// per matrix) and |U[out,in]|^2 is extracted — same fids[S][C][B] layout, so rc_stats / rc_stats_unsorted and
// the ranking stage apply unchanged.  Throughput is that of the dense exponential (a generality path, not a
// roofline path): the tridiagonal kernels remain the fast path for open chains.
// rc_directional_fidelity_mc: the same pipeline for directional_perturbation (noise_model.py:150-201).
#include <cuComplex.h>
#include "rc_common.cuh"
#include "rc_fidelity.cuh"

using namespace rc;

namespace rc {

struct DenseArgs {
    FidArgs f;
    long long e0, count;      // evaluations [e0, e0 + count) of the sweep form this tile
    int ring;
    cuDoubleComplex* A;       // [count][N][N]  -i T H
    double* draws_out;        // directional sweep, Philox mode: optional [S][C][B][3] record of (direction, n0, n1)
};

template <int MODEL, bool REPLAY>
__global__ void __launch_bounds__(128) dense_build_kernel(DenseArgs g) {
    const FidArgs& a = g.f;
    const int n = a.N;
    constexpr int P = draws_per_site(MODEL);
    const long long K = (long long)P * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < g.count; t += (long long)gridDim.x * blockDim.x) {
        const long long ev = g.e0 + t;
        const EvalIndex ix = decode_eval(ev, a.C, a.B);
        const double* x = a.ctrl + ix.c * (n + 1);
        const double sigma = __ldg(a.sigma + ix.s);
        const double T = fabs(__ldg(x + n));
        double z[3 * MAX_N];
        if (REPLAY) {
            for (int j = 0; j < K; ++j) z[j] = __ldg(a.replay + ev * K + j);
        } else {
            for (int j = 0; j < K; ++j) z[j] = 0.0;
            normals_fill(noise_key(a, ix.s, ix.c, ix.b), (int)K - (P - 1), a.zig,
                         [&](int jc) -> double& { return z[jc == 0 ? 0 : jc + (P - 1)]; });
        }
        cuDoubleComplex* A = g.A + t * n * n;
        for (int j = 0; j < n * n; ++j) A[j] = make_cuDoubleComplex(0.0, 0.0);
        for (int i = 0; i < n; ++i) {
            const double base = a.zz ? zz_diag(i, n) - ((g.ring && (i == 0 || i == n - 1)) ? 1.0 : 0.0) + (g.ring ? 0.5 : 0.0) : 0.0;
            const double hii = __dadd_rn(__dadd_rn(base, __dmul_rn(sigma, z[P * i])), __ldg(x + i));
            A[i * n + i] = make_cuDoubleComplex(0.0, -T * hii);
            if (i >= 1) {
                const double re = __dadd_rn(1.0, __dmul_rn(sigma, z[P * i + 1]));
                const double im = MODEL == MODEL_COMPLEX3 ? __dmul_rn(sigma, z[P * i + 2]) : 0.0;
                // H[i][i-1] = re + i im, H[i-1][i] = re - i im;  A = -i T H
                A[i * n + (i - 1)] = make_cuDoubleComplex(T * im, -T * re);
                A[(i - 1) * n + i] = make_cuDoubleComplex(-T * im, -T * re);
            }
        }
        if (g.ring && n > 2) {
            A[(n - 1) * n] = make_cuDoubleComplex(A[(n - 1) * n].x, A[(n - 1) * n].y - T);
            A[n - 1] = make_cuDoubleComplex(A[n - 1].x, A[n - 1].y - T);
        }
    }
}


// directional_perturbation (noise_model.py:150-201): ONE Hermitian pair of entries of H is perturbed per evaluation.
// directions (noise_model.py:155-163), index k of 3N: 0 -> (0,0); 1 -> (N-1,N-1); 2 + 3(d-1) + (o+1) -> (d, d+o) for
// d = 1..N-2, o = -1,0,1; 3N-4 -> (0,1); 3N-3 -> (1,0); 3N-2 -> (N-2,N-1); 3N-1 -> (N-1,N-2).
// z[i][j] = v, then z[j][i] = conj(v) with v = sigma (n0 + i n1) (noise_model.py:196-199): on a diagonal direction the
// second assignment wins, H[i][i] gets the complex number conj(v) and H is not Hermitian — hence the dense path.
// Draws per evaluation (replay layout [S][C][B][3]): the direction index as a double, then the two standard normals;
// Philox mode: index = (first word of block sub=127) * 3N >> 32, normals = compact draws 0, 1 of the primary stream.
constexpr uint32_t DIRECTION_SUBSTREAM = 127u;   // between the primary blocks (< 49) and the completion streams (>= 128)

__device__ __forceinline__ void direction_of(int k, int n, int* i, int* j) {
    if (k == 0) { *i = 0; *j = 0; }
    else if (k == 1) { *i = n - 1; *j = n - 1; }
    else if (k < 3 * n - 4) { const int m = k - 2, d = 1 + m / 3; *i = d; *j = d + (m - 3 * (m / 3)) - 1; }
    else if (k == 3 * n - 4) { *i = 0; *j = 1; }
    else if (k == 3 * n - 3) { *i = 1; *j = 0; }
    else if (k == 3 * n - 2) { *i = n - 2; *j = n - 1; }
    else { *i = n - 1; *j = n - 2; }
}

template <bool REPLAY>
__global__ void __launch_bounds__(128) directional_build_kernel(DenseArgs g) {
    const FidArgs& a = g.f;
    const int n = a.N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < g.count; t += (long long)gridDim.x * blockDim.x) {
        const long long ev = g.e0 + t;
        const EvalIndex ix = decode_eval(ev, a.C, a.B);
        const double* x = a.ctrl + ix.c * (n + 1);
        const double sigma = __ldg(a.sigma + ix.s);
        const double T = fabs(__ldg(x + n));
        double kd, nv[2];
        if (REPLAY) {
            kd = __ldg(a.replay + ev * 3);
            nv[0] = __ldg(a.replay + ev * 3 + 1);
            nv[1] = __ldg(a.replay + ev * 3 + 2);
        } else {
            const NoiseKey key = noise_key(a, ix.s, ix.c, ix.b);
            const Philox4 r = philox_block(key, DIRECTION_SUBSTREAM);
            kd = (double)(uint32_t)(((uint64_t)r.x * (uint64_t)(3 * n)) >> 32);
            nv[0] = nv[1] = 0.0;
            normals_fill(key, 2, a.zig, [&](int jc) -> double& { return nv[jc]; });
            if (g.draws_out) { g.draws_out[ev * 3] = kd; g.draws_out[ev * 3 + 1] = nv[0]; g.draws_out[ev * 3 + 2] = nv[1]; }
        }
        cuDoubleComplex* A = g.A + t * n * n;
        const bool ok = kd >= 0.0 && kd < (double)(3 * n) && kd == floor(kd);
        if (!ok) {   // not a direction index: NaN fidelity
            for (int j = 0; j < n * n; ++j) A[j] = make_cuDoubleComplex(NAN, NAN);
            continue;
        }
        int pi, pj;
        direction_of((int)kd, n, &pi, &pj);
        const double vr = __dmul_rn(sigma, nv[0]), vi = __dmul_rn(sigma, nv[1]);
        for (int j = 0; j < n * n; ++j) A[j] = make_cuDoubleComplex(0.0, 0.0);
        for (int i = 0; i < n; ++i) {
            const double base = a.zz ? zz_diag(i, n) - ((g.ring && (i == 0 || i == n - 1)) ? 1.0 : 0.0) + (g.ring ? 0.5 : 0.0) : 0.0;
            // (HH_ii + z_ii) + x_i, z_ii = conj(v) on a diagonal direction (noise_model.py:100-104)
            const bool hit = pi == pj && pi == i;
            const double hr = __dadd_rn(__dadd_rn(base, hit ? vr : 0.0), __ldg(x + i));
            const double hi = hit ? -vi : 0.0;
            A[i * n + i] = make_cuDoubleComplex(T * hi, -T * hr);          // -i T (hr + i hi)
            if (i >= 1) {
                // H[i][i-1] = 1 + z[i][i-1], H[i-1][i] = 1 + z[i-1][i]
                double lr = 1.0, li = 0.0, ur = 1.0, ui = 0.0;
                if (pi == i && pj == i - 1) { lr = __dadd_rn(1.0, vr); li = vi; ur = lr; ui = -vi; }
                if (pi == i - 1 && pj == i) { ur = __dadd_rn(1.0, vr); ui = vi; lr = ur; li = -vi; }
                A[i * n + (i - 1)] = make_cuDoubleComplex(T * li, -T * lr);
                A[(i - 1) * n + i] = make_cuDoubleComplex(T * ui, -T * ur);
            }
        }
        if (g.ring && n > 2) {
            A[(n - 1) * n] = make_cuDoubleComplex(A[(n - 1) * n].x, A[(n - 1) * n].y - T);
            A[n - 1] = make_cuDoubleComplex(A[n - 1].x, A[n - 1].y - T);
        }
    }
}

__global__ void dense_extract_kernel(const cuDoubleComplex* __restrict__ U, long long count, int n, int in, int out,
                                     double* __restrict__ fids) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (long long)gridDim.x * blockDim.x) {
        const cuDoubleComplex u = U[t * n * n + (long long)out * n + in];
        fids[t] = fma(u.x, u.x, u.y * u.y);
    }
}

cudaError_t zig_tables_device(ZigTables* t);

}  // namespace rc

extern "C" size_t rc_dense_fidelity_mc_workspace_bytes(int nspin, int64_t tile) {
    if (nspin < 2 || nspin > RC_MAX_NSPIN || tile < 1) return 256;
    return (size_t)2 * tile * nspin * nspin * sizeof(cuDoubleComplex) + 512;
}

extern "C" int rc_dense_fidelity_mc(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                                    const double* sigma_dev, int S, int64_t B, int model, int zz, int ring, uint64_t seed,
                                    int64_t c_offset, int64_t b_offset, const double* replay_dev, double* fids_dev,
                                    void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (nspin < 2 || nspin > RC_MAX_NSPIN) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, RC_MAX_NSPIN);
    if (inspin < 0 || inspin >= nspin || outspin < 0 || outspin >= nspin)
        return set_error(RC_ERR_BAD_ARG, "inspin=%d / outspin=%d outside [0,%d)", inspin, outspin, nspin);
    if (C < 0 || S < 0 || B < 0) return set_error(RC_ERR_BAD_ARG, "negative size C=%lld S=%d B=%lld", (long long)C, S, (long long)B);
    if (model != RC_MODEL_COMPLEX3 && model != RC_MODEL_REAL2) return set_error(RC_ERR_BAD_ARG, "unknown model %d", model);
    const long long total = (long long)S * C * B;
    if (total == 0) return RC_OK;
    if (!ctrl_dev || !sigma_dev || !fids_dev || !workspace_dev) return set_error(RC_ERR_NULL, "rc_dense_fidelity_mc: null pointer");
    const size_t per = (size_t)2 * nspin * nspin * sizeof(cuDoubleComplex);
    long long tile = workspace_bytes > 512 ? (long long)((workspace_bytes - 512) / per) : 0;
    if (tile < 1) return set_error(RC_ERR_WORKSPACE, "rc_dense_fidelity_mc: workspace holds no matrix (need %zu bytes each)", per);
    if (tile > total) tile = total;
    cudaStream_t st = (cudaStream_t)stream;
    DenseArgs g = {};
    FidArgs& a = g.f;
    a.ctrl = ctrl_dev; a.sigma = sigma_dev; a.replay = replay_dev; a.C = C; a.B = B; a.S = S; a.N = nspin;
    a.in = inspin; a.out = outspin; a.model = model; a.zz = zz;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.c_offset = c_offset; a.b_offset = b_offset;
    RC_CUDA_TRY(zig_tables_device(&a.zig));
    g.ring = (ring && nspin > 2) ? 1 : 0;
    g.A = (cuDoubleComplex*)(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
    cuDoubleComplex* U = g.A + (size_t)tile * nspin * nspin;
    const int sm = device_sm_count();
    for (long long e0 = 0; e0 < total; e0 += tile) {
        g.e0 = e0;
        g.count = total - e0 < tile ? total - e0 : tile;
        long long blocks = (g.count + 127) / 128;
        if (blocks > (long long)sm * 16) blocks = (long long)sm * 16;
        const bool replay = replay_dev != nullptr;
        if (model == RC_MODEL_COMPLEX3) {
            if (replay) dense_build_kernel<MODEL_COMPLEX3, true><<<(unsigned)blocks, 128, 0, st>>>(g);
            else dense_build_kernel<MODEL_COMPLEX3, false><<<(unsigned)blocks, 128, 0, st>>>(g);
        } else {
            if (replay) dense_build_kernel<MODEL_REAL2, true><<<(unsigned)blocks, 128, 0, st>>>(g);
            else dense_build_kernel<MODEL_REAL2, false><<<(unsigned)blocks, 128, 0, st>>>(g);
        }
        rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        int rcode = rc_expm_batch((const double*)g.A, g.count, nspin, (double*)U, stream);
        if (rcode) return rcode;
        dense_extract_kernel<<<(unsigned)blocks, 128, 0, st>>>(U, g.count, nspin, inspin, outspin, fids_dev + e0);
        rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
    }
    return RC_OK;
}

extern "C" int rc_directional_fidelity_mc(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                                          const double* sigma_dev, int S, int64_t B, int zz, int ring, uint64_t seed,
                                          int64_t c_offset, int64_t b_offset, const double* replay_dev, double* fids_dev,
                                          double* draws_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (nspin < 2 || nspin > RC_MAX_NSPIN) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, RC_MAX_NSPIN);
    if (inspin < 0 || inspin >= nspin || outspin < 0 || outspin >= nspin)
        return set_error(RC_ERR_BAD_ARG, "inspin=%d / outspin=%d outside [0,%d)", inspin, outspin, nspin);
    if (C < 0 || S < 0 || B < 0) return set_error(RC_ERR_BAD_ARG, "negative size C=%lld S=%d B=%lld", (long long)C, S, (long long)B);
    const long long total = (long long)S * C * B;
    if (total == 0) return RC_OK;
    if (!ctrl_dev || !sigma_dev || !fids_dev || !workspace_dev) return set_error(RC_ERR_NULL, "rc_directional_fidelity_mc: null pointer");
    const size_t per = (size_t)2 * nspin * nspin * sizeof(cuDoubleComplex);
    long long tile = workspace_bytes > 512 ? (long long)((workspace_bytes - 512) / per) : 0;
    if (tile < 1) return set_error(RC_ERR_WORKSPACE, "rc_directional_fidelity_mc: workspace holds no matrix (need %zu bytes each)", per);
    if (tile > total) tile = total;
    cudaStream_t st = (cudaStream_t)stream;
    DenseArgs g = {};
    FidArgs& a = g.f;
    a.ctrl = ctrl_dev; a.sigma = sigma_dev; a.replay = replay_dev; a.C = C; a.B = B; a.S = S; a.N = nspin;
    a.in = inspin; a.out = outspin; a.model = MODEL_COMPLEX3; a.zz = zz;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.c_offset = c_offset; a.b_offset = b_offset;
    RC_CUDA_TRY(zig_tables_device(&a.zig));
    g.ring = (ring && nspin > 2) ? 1 : 0;
    g.draws_out = replay_dev ? nullptr : draws_out_dev;
    g.A = (cuDoubleComplex*)(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
    cuDoubleComplex* U = g.A + (size_t)tile * nspin * nspin;
    const int sm = device_sm_count();
    for (long long e0 = 0; e0 < total; e0 += tile) {
        g.e0 = e0;
        g.count = total - e0 < tile ? total - e0 : tile;
        long long blocks = (g.count + 127) / 128;
        if (blocks > (long long)sm * 16) blocks = (long long)sm * 16;
        if (replay_dev) directional_build_kernel<true><<<(unsigned)blocks, 128, 0, st>>>(g);
        else directional_build_kernel<false><<<(unsigned)blocks, 128, 0, st>>>(g);
        rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        int rcode = rc_expm_batch((const double*)g.A, g.count, nspin, (double*)U, stream);
        if (rcode) return rcode;
        dense_extract_kernel<<<(unsigned)blocks, 128, 0, st>>>(U, g.count, nspin, inspin, outspin, fids_dev + e0);
        rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
    }
    return RC_OK;
}
