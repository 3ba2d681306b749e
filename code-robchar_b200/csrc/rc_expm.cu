// Dense complex matrix exponential, one CTA per matrix (M <= 32): the generality path.
//
// Stands behind every `scipy.linalg.expm` call of the reference whose argument is NOT a Hermitian
// tridiagonal matrix and therefore cannot take the eigensolver fast path: `topo="ring"`
// (noise_model.py:83-85), directional_perturbation's complex diagonal entries (noise_model.py:196-199),
// arbitrary `perturbation()` overrides, and the 2N x 2N block matrices of the analytic gradient
// (qnewton.py:186-196).  Algorithm: scaling and squaring with the [13/13] Pade approximant
// (Higham 2005, the algorithm behind scipy.linalg.expm; published coefficients b_0..b_13 and
// theta_13 = 5.37), always degree 13: A <- A / 2^s with s = max(0, ceil(log2(|A|_1 / theta_13))),
// R = (V - U)^{-1} (V + U) by Gauss-Jordan elimination with partial pivoting, then s squarings.
// Thread (i, j) owns element (i, j) of every M x M product; matrices live in shared memory.
#include <cuComplex.h>
#include "rc_common.cuh"

namespace rc {

typedef cuDoubleComplex cplx;
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_cuDoubleComplex(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {
    return make_cuDoubleComplex(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cplx cscale(double s, cplx a) { return make_cuDoubleComplex(s * a.x, s * a.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_cuDoubleComplex(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_cuDoubleComplex(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cdivz(cplx a, cplx b) {  // a / b, scaled to avoid overflow
    const double s = fmax(fabs(b.x), fabs(b.y));
    const double br = b.x / s, bi = b.y / s, den = (br * br + bi * bi) * s;
    return make_cuDoubleComplex((a.x * br + a.y * bi) / den, (a.y * br - a.x * bi) / den);
}

// C = X * Y for the element owned by this thread
__device__ __forceinline__ cplx mm_elem(const cplx* X, const cplx* Y, int M, int i, int j) {
    cplx acc = make_cuDoubleComplex(0.0, 0.0);
    for (int k = 0; k < M; ++k) acc = cfma(X[i * M + k], Y[k * M + j], acc);
    return acc;
}

__global__ void __launch_bounds__(1024) expm_pade13_kernel(const cplx* __restrict__ Ain, long long batch, int M,
                                                           cplx* __restrict__ out) {
    extern __shared__ cplx smc[];
    const int MM = M * M;
    cplx* A = smc;            // scaled input
    cplx* A2 = A + MM;
    cplx* A4 = A2 + MM;
    cplx* A6 = A4 + MM;
    cplx* W = A6 + MM;        // work / U
    cplx* P = W + MM;         // V - U, becomes identity
    cplx* Q = P + MM;         // V + U, becomes the result
    __shared__ double colsum[32];
    __shared__ int piv_s;
    __shared__ int squarings;
    const int tid = threadIdx.x;
    const bool act = tid < MM;
    const int i = act ? tid / M : 0, j = act ? tid % M : 0;
    const double b[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
                          10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};
    for (long long mat = blockIdx.x; mat < batch; mat += gridDim.x) {
        __syncthreads();
        if (act) A[tid] = Ain[mat * MM + tid];
        __syncthreads();
        // 1-norm = max column sum of |a_ij|
        if (tid < M) {
            double sres = 0.0;
            for (int r = 0; r < M; ++r) sres += hypot(A[r * M + tid].x, A[r * M + tid].y);
            colsum[tid] = sres;
        }
        __syncthreads();
        if (tid == 0) {
            double nrm = 0.0;
            bool bad = false;
            for (int c = 0; c < M; ++c) { nrm = fmax(nrm, colsum[c]); bad |= !(colsum[c] <= 1.79e308); }
            int s = 0;
            if (!bad && nrm > 5.371920351148152) s = (int)ceil(log2(nrm / 5.371920351148152));
            if (s > 60) s = 60;
            squarings = bad ? -1 : s;
        }
        __syncthreads();
        const int s = squarings;
        if (s < 0) {  // NaN / Inf input -> NaN output (controller NaN sentinel, mcsim.py:443)
            if (act) out[mat * MM + tid] = make_cuDoubleComplex(NAN, NAN);
            continue;
        }
        const double scale = ldexp(1.0, -s);
        if (act) A[tid] = cscale(scale, A[tid]);
        __syncthreads();
        cplx t;
        if (act) t = mm_elem(A, A, M, i, j);
        if (act) A2[tid] = t;
        __syncthreads();
        if (act) A4[tid] = mm_elem(A2, A2, M, i, j);
        __syncthreads();
        if (act) A6[tid] = mm_elem(A4, A2, M, i, j);
        __syncthreads();
        // U = A (A6 (b13 A6 + b11 A4 + b9 A2) + b7 A6 + b5 A4 + b3 A2 + b1 I)
        if (act) W[tid] = cadd(cadd(cscale(b[13], A6[tid]), cscale(b[11], A4[tid])), cscale(b[9], A2[tid]));
        __syncthreads();
        cplx u1;
        if (act) {
            u1 = mm_elem(A6, W, M, i, j);
            u1 = cadd(u1, cadd(cadd(cscale(b[7], A6[tid]), cscale(b[5], A4[tid])), cscale(b[3], A2[tid])));
            if (i == j) u1.x += b[1];
        }
        __syncthreads();
        if (act) W[tid] = u1;
        __syncthreads();
        cplx U;
        if (act) U = mm_elem(A, W, M, i, j);
        __syncthreads();
        // V = A6 (b12 A6 + b10 A4 + b8 A2) + b6 A6 + b4 A4 + b2 A2 + b0 I
        if (act) W[tid] = cadd(cadd(cscale(b[12], A6[tid]), cscale(b[10], A4[tid])), cscale(b[8], A2[tid]));
        __syncthreads();
        if (act) {
            cplx V = mm_elem(A6, W, M, i, j);
            V = cadd(V, cadd(cadd(cscale(b[6], A6[tid]), cscale(b[4], A4[tid])), cscale(b[2], A2[tid])));
            if (i == j) V.x += b[0];
            P[tid] = csub(V, U);
            Q[tid] = cadd(V, U);
        }
        __syncthreads();
        // Gauss-Jordan: P X = Q, partial pivoting; thread (i, j) updates P[i][j] and Q[i][j]
        for (int k = 0; k < M; ++k) {
            if (tid == 0) {
                int pv = k;
                double best = -1.0;
                for (int r = k; r < M; ++r) {
                    double mag = fabs(P[r * M + k].x) + fabs(P[r * M + k].y);
                    if (mag > best) { best = mag; pv = r; }
                }
                piv_s = pv;
            }
            __syncthreads();
            const int pv = piv_s;
            if (pv != k && act && i == k) {  // row k threads swap rows k and pv (both matrices)
                cplx a = P[k * M + j]; P[k * M + j] = P[pv * M + j]; P[pv * M + j] = a;
                cplx q = Q[k * M + j]; Q[k * M + j] = Q[pv * M + j]; Q[pv * M + j] = q;
            }
            __syncthreads();
            const cplx pkk = P[k * M + k];
            cplx pkj, qkj, fac;
            if (act) {
                pkj = cdivz(P[k * M + j], pkk);
                qkj = cdivz(Q[k * M + j], pkk);
                fac = P[i * M + k];
            }
            __syncthreads();
            if (act) {
                if (i == k) { P[tid] = pkj; Q[tid] = qkj; }
                else { P[tid] = csub(P[tid], cmul(fac, pkj)); Q[tid] = csub(Q[tid], cmul(fac, qkj)); }
            }
            __syncthreads();
        }
        // squarings: Q <- Q^2, s times (ping-pong with W)
        cplx* X = Q;
        cplx* Y = W;
        for (int r = 0; r < s; ++r) {
            if (act) Y[tid] = mm_elem(X, X, M, i, j);
            __syncthreads();
            cplx* tswap = X; X = Y; Y = tswap;
        }
        if (act) out[mat * MM + tid] = X[tid];
    }
}

}  // namespace rc

using namespace rc;

extern "C" int rc_expm_batch(const double* A_dev, int64_t batch, int M, double* out_dev, void* stream) {
    if (M < 1 || M > 32) return set_error(RC_ERR_BAD_ARG, "rc_expm_batch: M=%d outside [1,32]", M);
    if (batch < 0) return set_error(RC_ERR_BAD_ARG, "rc_expm_batch: negative batch");
    if (batch == 0) return RC_OK;
    if (!A_dev || !out_dev) return set_error(RC_ERR_NULL, "rc_expm_batch: null pointer");
    int threads = ((M * M + 31) / 32) * 32;
    size_t smem = (size_t)7 * M * M * sizeof(cplx);
    RC_CUDA_TRY(cudaFuncSetAttribute(expm_pade13_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    RC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, expm_pade13_kernel, threads, smem));
    if (occ < 1) occ = 1;
    long long grid = (long long)device_sm_count() * occ;
    if (grid > batch) grid = batch;
    expm_pade13_kernel<<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>((const cplx*)A_dev, batch, M, (cplx*)out_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}
