// Infidelity and its analytic gradient w.r.t. the biases and the evolution time, from the eigendecomposition.
//
// Reference: LBFGS.eval_static_fidelity_gradient (qnewton.py:162-212), the L-BFGS inner call (qnewton.py:497,513).
// Upstream evaluates N + 1 dense matrix exponentials per call — expm(-iTH) and, per bias l, the 2N x 2N block
// matrix expm([[-iTH, 0], [-iT C_l, -iTH]]) whose lower-left block is dU/dx_l.  With H = V diag(lambda) V^T (real
// symmetric tridiagonal in the single-excitation subspace) the same quantities are
//
//   phi            = U[out,in]        = sum_k V[out,k] V[in,k] e_k,              e_k = exp(-i lambda_k T)
//   (H U)[out,in]                     = sum_k lambda_k V[out,k] V[in,k] e_k
//   (dU/dx_l)[out,in]                 = sum_{k,m} V[out,k] V[l,k] Phi_km V[l,m] V[in,m]
//   Phi_km = (e_k - e_m)/(lambda_k - lambda_m) = -i T h_k h_m sinc(delta),   h_k = exp(-i lambda_k T/2),
//            delta = (lambda_k - lambda_m) T / 2,  sin(delta) = -Im(h_k conj(h_m))      (Phi_kk = -i T e_k)
//
//   err = 1 - |phi|^2,  grad[l] = -2 Re((dU/dx_l)[out,in] conj(phi)),  grad[N] = -2 Im((H U)[out,in] conj(phi)).
//
// The sinc form is an identity, stable for coincident eigenvalues, and needs no trigonometric call beyond the N
// half-angle phases.  One WARP per controller, any N <= 32: lane k holds d_k / e_k of the tridiagonal matrix in
// registers (the QL chase reads them by shuffle; every lane runs the scalar recurrence redundantly, so there is no
// cross-lane memory traffic to order), lane j owns row j of V in shared memory and applies every rotation to it.
#include <stdlib.h>
#include <string.h>
#include "rc_common.cuh"
#include "rc_fidelity.cuh"

using namespace rc;

namespace rc {

struct GradArgs {
    const double* x;      // [C][N+1]
    const double* rows;   // [C][2N] explicit perturbations in the real 2-draw replay layout (sigma = 1), or nullptr
    double* err;          // [C]
    double* grad;         // [C][N+1]
    unsigned long long* nonconv;
    long long C;
    int N, in, out, zz;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int GRAD_WARPS = 4;

__host__ __device__ inline size_t grad_warp_doubles(int n) {
    const int P = n | 1;
    return (size_t)3 * n * P + 3 * n;   // V, PhiRe, PhiIm rows (pitch P) + hr, hi, lambda
}

__global__ void __launch_bounds__(32 * GRAD_WARPS) fidelity_grad_kernel(GradArgs a) {
    extern __shared__ double sm[];
    const int n = a.N, P = n | 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* V = sm + (size_t)warp * grad_warp_doubles(n);
    double* PhiRe = V + (size_t)n * P;
    double* PhiIm = PhiRe + (size_t)n * P;
    double* hr = PhiIm + (size_t)n * P;
    double* hi = hr + n;
    double* lam = hi + n;
    for (long long c = (long long)blockIdx.x * GRAD_WARPS + warp; c < a.C; c += (long long)gridDim.x * GRAD_WARPS) {
        const double* x = a.x + c * (n + 1);
        const double* row = a.rows ? a.rows + c * 2 * n : nullptr;
        const double T = fabs(__ldg(x + n));
        // ---- build: lane i holds d_i and e_i (coupling between sites i and i+1) --------------------------------
        double dl = 0.0, el = 0.0;
        if (lane < n) {
            const double base = a.zz ? zz_diag(lane, n) : 0.0;
            dl = __dadd_rn(__dadd_rn(base, row ? __ldg(row + 2 * lane) : 0.0), __ldg(x + lane));
            if (lane < n - 1) el = __dadd_rn(1.0, row ? __ldg(row + 2 * (lane + 1) + 1) : 0.0);
            for (int i = 0; i < n; ++i) V[(size_t)lane * P + i] = (i == lane) ? 1.0 : 0.0;
        }
        double anorm = fabs(dl) + fabs(el), chk = dl + el;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) anorm = fmax(anorm, __shfl_xor_sync(0xffffffffu, anorm, o));
        chk = warp_sum(chk) + T;
        const bool finite = fabs(chk) <= DBL_MAX;
        const double tol = DBL_EPSILON * anorm;
        const int tolhi = threshold_hi(tol);
        const double tiny = fmin(tol, 1e-280);
        // ---- implicit QL, eigenvectors accumulated: warp-uniform control flow ---------------------------------
        int l = 0, it = 0;
        bool bad = false;
        while (finite && l < n - 1) {
            const unsigned negl = __ballot_sync(0xffffffffu, lane < n - 1 && negligible_hi(el, tolhi)) | (1u << (n - 1));
            const int m = __ffs(negl >> l) - 1 + l;       // first negligible coupling at or after l
            if (m == l) { ++l; it = 0; continue; }
            if (++it > QL_MAX_SWEEPS) { bad = true; break; }
            const double d_l = __shfl_sync(0xffffffffu, dl, l), d_l1 = __shfl_sync(0xffffffffu, dl, l + 1);
            const double e_l = __shfl_sync(0xffffffffu, el, l);
            double d_up = __shfl_sync(0xffffffffu, dl, m);
            double g = wilkinson_g(d_l, d_l1, e_l, d_up);
            double s = 1.0, cc = 1.0, p = 0.0;
            double* vrow = V + (size_t)lane * P;
            double z_up = lane < n ? vrow[m] : 0.0;
            for (int i = m - 1; i >= l; --i) {
                const double ei = __shfl_sync(0xffffffffu, el, i), di = __shfl_sync(0xffffffffu, dl, i);
                const double f = s * ei, b = cc * ei;
                const double h = fma(f, f, fma(g, g, tiny));
                const double rinv = rc_rsqrt(h);
                double r = h * rinv;
                if (lane == i + 1) el = r;
                s = f * rinv;
                cc = g * rinv;
                g = d_up - p;
                r = fma(di - g, s, (2.0 * cc) * b);
                p = s * r;
                if (lane == i + 1) dl = g + p;
                g = cc * r - b;
                d_up = di;
                if (lane < n) {
                    const double zi = vrow[i];
                    vrow[i + 1] = fma(s, zi, cc * z_up);
                    z_up = fma(cc, zi, -(s * z_up));
                }
            }
            if (lane < n) vrow[l] = z_up;
            if (lane == l) { dl = d_up - p; el = g; }
            if (lane == m) el = 0.0;
        }
        if (!finite || bad) {
            if (bad && a.nonconv && lane == 0) atomicAdd(a.nonconv, 1ull);
            if (lane == 0) a.err[c] = NAN;
            if (lane <= n) a.grad[c * (n + 1) + lane] = NAN;
            __syncwarp();
            continue;
        }
        // ---- phases: h_k = exp(-i lambda_k T / 2), e_k = h_k^2 ------------------------------------------------
        double hkr = 1.0, hki = 0.0;
        if (lane < n) {
            double sn, cs;
            rc_sincos_tab(0.5 * dl * T, &sn, &cs);
            hkr = cs; hki = -sn;
            hr[lane] = hkr; hi[lane] = hki; lam[lane] = dl;
        }
        __syncwarp();
        const double ekr = fma(hkr, hkr, -(hki * hki)), eki = 2.0 * hkr * hki;
        const double wk = lane < n ? V[(size_t)a.out * P + lane] * V[(size_t)a.in * P + lane] : 0.0;
        const double phr = warp_sum(wk * ekr), phi_i = warp_sum(wk * eki);
        const double hur = warp_sum(dl * wk * ekr), hui = warp_sum(dl * wk * eki);
        // ---- divided differences: row k of Phi in lane k ---------------------------------------------------------
        if (lane < n) {
            for (int m = 0; m < n; ++m) {
                const double lm = lam[m];
                const double hmr = hr[m], hmi = hi[m];
                const double pr = fma(hkr, hmr, -(hki * hmi)), pi = fma(hkr, hmi, hki * hmr);     // h_k h_m
                const double sind = -fma(hki, hmr, -(hkr * hmi));                                  // sin(delta)
                const double delta = 0.5 * (dl - lm) * T;
                const double sinc = fabs(delta) < 1e-4 ? fma(-delta * delta, 1.0 / 6.0, 1.0) : sind / delta;
                // Phi = -i T (pr + i pi) sinc = T sinc (pi - i pr)
                PhiRe[(size_t)lane * P + m] = T * sinc * pi;
                PhiIm[(size_t)lane * P + m] = -(T * sinc * pr);
            }
        }
        __syncwarp();
        // ---- bias derivatives ---------------------------------------------------------------------------------------
        for (int lsite = 0; lsite < n; ++lsite) {
            double gr = 0.0, gi = 0.0;
            if (lane < n) {
                const double ak = V[(size_t)a.out * P + lane] * V[(size_t)lsite * P + lane];
                double ir = 0.0, ii = 0.0;
                for (int m = 0; m < n; ++m) {
                    const double bm = V[(size_t)lsite * P + m] * V[(size_t)a.in * P + m];
                    ir = fma(PhiRe[(size_t)lane * P + m], bm, ir);
                    ii = fma(PhiIm[(size_t)lane * P + m], bm, ii);
                }
                gr = ak * ir; gi = ak * ii;
            }
            gr = warp_sum(gr); gi = warp_sum(gi);
            // z = G conj(phi); grad[l] = -2 Re z
            if (lane == 0) a.grad[c * (n + 1) + lsite] = -2.0 * fma(gr, phr, gi * phi_i);
        }
        if (lane == 0) {
            a.err[c] = 1.0 - fma(phr, phr, phi_i * phi_i);
            // z = (HU)[out,in] conj(phi); grad[N] = -2 Im z
            a.grad[c * (n + 1) + n] = -2.0 * fma(hui, phr, -(hur * phi_i));
        }
        __syncwarp();
    }
}

}  // namespace rc

extern "C" int rc_fidelity_grad(const double* x_dev, int64_t C, int nspin, int inspin, int outspin, const double* rows_dev,
                                int zz, double* err_dev, double* grad_dev, unsigned long long* nonconv_dev, void* stream) {
    if (nspin < 2 || nspin > RC_MAX_NSPIN) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, RC_MAX_NSPIN);
    if (inspin < 0 || inspin >= nspin || outspin < 0 || outspin >= nspin)
        return set_error(RC_ERR_BAD_ARG, "inspin=%d / outspin=%d outside [0,%d)", inspin, outspin, nspin);
    if (C < 0) return set_error(RC_ERR_BAD_ARG, "rc_fidelity_grad: C=%lld", (long long)C);
    if (C == 0) return RC_OK;
    if (!x_dev || !err_dev || !grad_dev) return set_error(RC_ERR_NULL, "rc_fidelity_grad: null x / err / grad pointer");
    GradArgs a;
    a.x = x_dev; a.rows = rows_dev; a.err = err_dev; a.grad = grad_dev; a.nonconv = nonconv_dev;
    a.C = C; a.N = nspin; a.in = inspin; a.out = outspin; a.zz = zz;
    const size_t smem = grad_warp_doubles(nspin) * GRAD_WARPS * sizeof(double);
    if (smem > 40 * 1024)
        RC_CUDA_TRY(cudaFuncSetAttribute(fidelity_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (C + GRAD_WARPS - 1) / GRAD_WARPS;
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    fidelity_grad_kernel<<<(unsigned)blocks, 32 * GRAD_WARPS, smem, (cudaStream_t)stream>>>(a); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}

extern "C" int rc_fidelity_grad_host(const double* x_host, int64_t C, int nspin, int inspin, int outspin,
                                     const double* rows_host, int zz, double* err_host, double* grad_host, void* stream) {
    if (C < 0) return set_error(RC_ERR_BAD_ARG, "rc_fidelity_grad_host: C=%lld", (long long)C);
    if (C == 0) return RC_OK;
    if (nspin < 2 || nspin > RC_MAX_NSPIN) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, RC_MAX_NSPIN);
    if (!x_host || !err_host || !grad_host) return set_error(RC_ERR_NULL, "rc_fidelity_grad_host: null x / err / grad pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nx = (size_t)C * (nspin + 1), nr = rows_host ? (size_t)C * 2 * nspin : 0;
    double* dev = nullptr;
    RC_CUDA_TRY(cudaMallocAsync((void**)&dev, (2 * nx + nr + (size_t)C + 1) * 8, st));
    double* dx = dev;
    double* drows = dx + nx;
    double* dgrad = drows + nr;
    double* derr = dgrad + nx;
    unsigned long long* dcnt = (unsigned long long*)(derr + C);
    int rcode = RC_OK;
    cudaError_t e = cudaMemsetAsync(dcnt, 0, 8, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x_host, nx * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nr) e = cudaMemcpyAsync(drows, rows_host, nr * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rcode = rc_fidelity_grad(dx, C, nspin, inspin, outspin, nr ? drows : nullptr, zz, derr, dgrad, dcnt, st);
    unsigned long long cnt = 0;
    if (e == cudaSuccess && rcode == RC_OK) e = cudaMemcpyAsync(err_host, derr, (size_t)C * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rcode == RC_OK) e = cudaMemcpyAsync(grad_host, dgrad, nx * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rcode == RC_OK) e = cudaMemcpyAsync(&cnt, dcnt, 8, cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(dev, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return set_error(RC_ERR_CUDA, "rc_fidelity_grad_host: %s", cudaGetErrorString(e));
    if (rcode) return rcode;
    if (cnt) return set_error(RC_ERR_NONCONV, "eigensolver did not converge for %llu controllers (NaN written)", cnt);
    return RC_OK;
}
