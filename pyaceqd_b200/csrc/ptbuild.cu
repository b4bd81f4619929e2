// On-device construction of the uniform process tensor of a Gaussian bath with diagonal coupling.
//
// Replaces the PT-generation run of pyaceqd/general_system/general_system.py:159-192 (`ACE <generate.param>` with
// `use_Gaussian_infinite true`, `threshold`, `Boson_J_type QDPhonon`, `write_PT`).  Same algorithm as the host
// reference implementation pyaceqd_b200/pt_builder.py (iTEBD contraction of the time-translation-invariant
// influence-functional network; Link, Tu, Strunz, PRL 132, 200403), with every tensor resident in HBM:
//   k_eta            QUAPI coefficients eta_k by trapezoid quadrature of the tabulated spectral density
//   per level k      ZGEMM (two-site tensor)  ->  k_gate (swap, influence weights, bond weights; writes the matrix the
//                    SVD factorises, transposed when it is wider than tall)  ->  cuSOLVER SVD  ->  k_take_right
//                    (kept right vectors become the new right site)  ->  ZGEMM (new left site)
//   cap              d ZGEMMs f[c] = i0[c] * B_x[:, c, :] B_y[:, c, :]
// Site tensors are stored as T[a + A (i + d b)] (left bond fastest, then the class index, then the right bond): the
// same bytes are the (A d) x B matrix a site is as LEFT factor of a product and the A x (d B) matrix it is as RIGHT
// factor, so no tensor is ever re-laid out between gates.  The SVDs are >95 % of the time (cuSOLVER, library code like
// cuBLAS); the host only reads the singular values of every gate to choose the kept rank.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/aceqd_ptbuild.h"

namespace {

thread_local char g_err[512] = "";
void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define PB_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);       \
            return -3;                                                                                 \
        }                                                                                              \
    } while (0)
#define PB_LIB(call, ok)                                                                               \
    do {                                                                                               \
        int s_ = (int)(call);                                                                          \
        if (s_ != (int)(ok)) {                                                                         \
            set_err("%s failed with status %d (%s:%d)", #call, s_, __FILE__, __LINE__);                \
            return -3;                                                                                 \
        }                                                                                              \
    } while (0)

typedef cuDoubleComplex zc;

struct DevBuf {   // grow-only device buffer
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            set_err("cudaMalloc of %zu bytes failed", want);
            return -4;
        }
        cap = want;
        return 0;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

// ---------------------------------------------------------------------------------------------------- eta_k
// One CTA per k.  Trapezoid rule on the caller's grid, integrand evaluated on the fly.
__global__ void __launch_bounds__(256) k_eta(int n_w, const double* __restrict__ w, const double* __restrict__ J, double dt,
                                             double h2kT, double2* __restrict__ eta) {
    const int k = blockIdx.x;
    auto f = [&](int i) -> double2 {
        const double ww = w[i];
        if (!(ww > 0.0)) return make_double2(0.0, 0.0);
        const double jw2 = J[i] / (ww * ww);
        double coth = 1.0;
        if (h2kT > 0.0) coth = 1.0 / tanh(h2kT * ww);
        double jc = jw2 * coth;
        if (!isfinite(jc)) jc = 0.0;
        const double wd = ww * dt;
        double s, c;
        sincos(wd, &s, &c);
        if (k == 0) return make_double2(jc * (1.0 - c), jw2 * (s - wd));
        double sk, ck;
        sincos((double)k * wd, &sk, &ck);
        const double omc = 2.0 * (1.0 - c);
        return make_double2(jc * omc * ck, -jw2 * omc * sk);
    };
    double ar = 0.0, ai = 0.0;
    for (int i = threadIdx.x; i + 1 < n_w; i += blockDim.x) {
        const double2 a = f(i), b = f(i + 1);
        const double h = 0.5 * (w[i + 1] - w[i]);
        ar += h * (a.x + b.x);
        ai += h * (a.y + b.y);
    }
    __shared__ double sr[256], si[256];
    sr[threadIdx.x] = ar;
    si[threadIdx.x] = ai;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sr[threadIdx.x] += sr[threadIdx.x + o];
            si[threadIdx.x] += si[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) eta[k] = make_double2(sr[0], si[0]);
}

// ---------------------------------------------------------------------------------------------------- gate kernels
// C [(l + chi_l i) + M (j + d r)]  ->  Cp[(l + chi_l a) + M (b + d r)] = C[l, i = b, j = a, r] * W[b][a]   (swap)
//                                                                         C[l, i = a, j = b, r] * W[b][a]   (no swap)
// and theta = lam[l] * Cp, written as theta (M x N) or as its conjugate transpose (N x M) for the SVD.
__global__ void k_gate(int chi_l, int d, int chi_r, const zc* __restrict__ C, const double* __restrict__ lam,
                       const zc* __restrict__ W, int swap, int transposed, zc* __restrict__ Cp, zc* __restrict__ theta) {
    const long long M = (long long)chi_l * d, N = (long long)d * chi_r;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const long long row = idx % M, col = idx / M;
    const int l = (int)(row % chi_l), a = (int)(row / chi_l);
    const int b = (int)(col % d), r = (int)(col / d);
    zc v = swap ? C[(l + (long long)chi_l * b) + M * (a + (long long)d * r)] : C[idx];
    if (W) v = cuCmul(v, W[b * d + a]);
    Cp[idx] = v;
    const double s = lam[l];
    v.x *= s;
    v.y *= s;
    if (transposed) theta[col + N * row] = make_cuDoubleComplex(v.x, -v.y);
    else theta[idx] = v;
}

// kept right singular vectors -> new right site By[k + keep (c)] , c = b + d r
//   from_vt: VT[k + ld c]       else: conj(Vr[c + ld k])
__global__ void k_take_right(int keep, long long N, const zc* __restrict__ V, long long ld, int from_vt, zc* __restrict__ By) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)keep * N) return;
    const long long k = idx % keep, c = idx / keep;
    if (from_vt) By[idx] = V[k + ld * c];
    else {
        const zc v = V[c + ld * k];
        By[idx] = make_cuDoubleComplex(v.x, -v.y);
    }
}

__global__ void k_fill(zc* p, long long n, double re) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) p[idx] = make_cuDoubleComplex(re, 0.0);
}

struct Site {
    DevBuf buf;
    int l = 1, r = 1;
    zc* p() { return (zc*)buf.p; }
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" const char* aceqd_ptbuild_last_error(void) { return g_err; }

extern "C" int aceqd_ptbuild_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int aceqd_ptbuild_eta(int device, int n_w, const double* w, const double* J, double dt, int K,
                                 double hbar_over_2kT, double* eta_out) {
    if (n_w < 2 || !w || !J || K < 0 || !eta_out) {
        set_err("ptbuild_eta: bad arguments");
        return -1;
    }
    PB_CUDA(cudaSetDevice(device));
    DevBuf dw, dj, de;
    if (dw.reserve(sizeof(double) * n_w) || dj.reserve(sizeof(double) * n_w) || de.reserve(16 * (size_t)(K + 1))) return -4;
    PB_CUDA(cudaMemcpy(dw.p, w, sizeof(double) * n_w, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(dj.p, J, sizeof(double) * n_w, cudaMemcpyHostToDevice));
    k_eta<<<K + 1, 256>>>(n_w, (const double*)dw.p, (const double*)dj.p, dt, hbar_over_2kT, (double2*)de.p);
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaMemcpy(eta_out, de.p, 16 * (size_t)(K + 1), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int aceqd_ptbuild_uniform(int device, int d, int K, const double* weights, const double* i0, double threshold,
                                     int chi_max, int svd_method, double* f_out, int* chi_out, double* stats) {
    if (d < 1 || d > 64 || K < 1 || !weights || !i0 || chi_max < 1 || !f_out || !chi_out || svd_method < 0 || svd_method > 1) {
        set_err("ptbuild_uniform: bad arguments");
        return -1;
    }
    PB_CUDA(cudaSetDevice(device));
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t sol = nullptr;
    cusolverDnParams_t par = nullptr;
    PB_LIB(cublasCreate(&blas), CUBLAS_STATUS_SUCCESS);
    PB_LIB(cusolverDnCreate(&sol), CUSOLVER_STATUS_SUCCESS);
    PB_LIB(cusolverDnCreateParams(&par), CUSOLVER_STATUS_SUCCESS);
    struct Guard {
        cublasHandle_t& b;
        cusolverDnHandle_t& s;
        cusolverDnParams_t& p;
        ~Guard() {
            if (p) cusolverDnDestroyParams(p);
            if (s) cusolverDnDestroy(s);
            if (b) cublasDestroy(b);
        }
    } guard{blas, sol, par};

    const double t_begin = now_ms();
    double ms_svd = 0.0, ms_gemm = 0.0, max_dim = 0.0;
    DevBuf dW, dC, dCp, dTheta, dS, dU, dV, dWork, dInfo, dLam[2], dF;
    Site B[2];
    if (dW.reserve(16 * (size_t)(K + 1) * d * d) || dInfo.reserve(sizeof(int))) return -4;
    PB_CUDA(cudaMemcpy(dW.p, weights, 16 * (size_t)(K + 1) * d * d, cudaMemcpyHostToDevice));
    for (int s = 0; s < 2; ++s) {
        if (B[s].buf.reserve(16 * (size_t)d) || dLam[s].reserve(sizeof(double))) return -4;
        k_fill<<<1, 64>>>(B[s].p(), d, 1.0 / sqrt((double)d));
        const double one = 1.0;
        PB_CUDA(cudaMemcpy(dLam[s].p, &one, sizeof(double), cudaMemcpyHostToDevice));
    }
    std::vector<double> S_host, lam_host;
    std::vector<char> work_host;
    const zc z_one = make_cuDoubleComplex(1.0, 0.0), z_zero = make_cuDoubleComplex(0.0, 0.0);

    // one two-site gate on the bond between site x (left) and site 1-x (right)
    auto gate = [&](int x, const zc* W, int swap) -> int {
        const int y = 1 - x;
        const int chi_l = B[x].l, chi_m = B[x].r, chi_r = B[y].r;
        if (chi_m != B[y].l) {
            set_err("ptbuild: bond mismatch %d vs %d", chi_m, B[y].l);
            return -5;
        }
        const long long M = (long long)chi_l * d, N = (long long)d * chi_r, mn = M < N ? M : N;
        const size_t mat = 16 * (size_t)M * N;
        if (dC.reserve(mat) || dCp.reserve(mat) || dTheta.reserve(mat) || dS.reserve(sizeof(double) * mn)) return -4;
        double t0 = now_ms();
        PB_LIB(cublasZgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)M, (int)N, chi_m, &z_one, B[x].p(), (int)M, B[y].p(), chi_m,
                           &z_zero, (zc*)dC.p, (int)M),
               CUBLAS_STATUS_SUCCESS);
        const int transposed = M < N;   // the SVD routines want tall matrices: factorise theta^H instead
        const long long tot = M * N;
        k_gate<<<(unsigned)((tot + 255) / 256), 256>>>(chi_l, d, chi_r, (const zc*)dC.p, (const double*)dLam[x].p, W, swap,
                                                       transposed, (zc*)dCp.p, (zc*)dTheta.p);
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaDeviceSynchronize());
        ms_gemm += now_ms() - t0;
        // ---- SVD of theta (m x n = M x N) or theta^H (N x M): right singular vectors of theta wanted
        const long long m = transposed ? N : M, n = transposed ? M : N;   // m >= n = mn
        max_dim = std::max(max_dim, (double)m);
        t0 = now_ms();
        const zc* right = nullptr;    // N x mn (ld N) unless from_vt
        long long ld_right = N;
        int from_vt = 0;
        size_t wd = 0, wh = 0;
        if (svd_method == 0) {
            // QR-iteration SVD.  theta: jobu = N, jobvt = S -> VT (n x n);  theta^H: jobu = S -> U (N x M) = right vectors
            const signed char jobu = transposed ? 'S' : 'N', jobvt = transposed ? 'N' : 'S';
            if (dU.reserve(transposed ? 16 * (size_t)m * n : 16) || dV.reserve(transposed ? 16 : 16 * (size_t)n * n)) return -4;
            PB_LIB(cusolverDnXgesvd_bufferSize(sol, par, jobu, jobvt, m, n, CUDA_C_64F, dTheta.p, m, CUDA_R_64F, dS.p, CUDA_C_64F,
                                               dU.p, m, CUDA_C_64F, dV.p, n, CUDA_C_64F, &wd, &wh),
                   CUSOLVER_STATUS_SUCCESS);
            if (dWork.reserve(wd + 16)) return -4;
            if (work_host.size() < wh + 16) work_host.resize(wh + 16);
            PB_LIB(cusolverDnXgesvd(sol, par, jobu, jobvt, m, n, CUDA_C_64F, dTheta.p, m, CUDA_R_64F, dS.p, CUDA_C_64F, dU.p, m,
                                    CUDA_C_64F, dV.p, n, CUDA_C_64F, dWork.p, wd, work_host.data(), wh, (int*)dInfo.p),
                   CUSOLVER_STATUS_SUCCESS);
            if (transposed) {
                right = (const zc*)dU.p;
            } else {
                right = (const zc*)dV.p;
                ld_right = n;
                from_vt = 1;
            }
        } else {
            // polar-decomposition SVD: A = U S V^H with U (m x n), V (n x n)
            if (dU.reserve(16 * (size_t)m * n) || dV.reserve(16 * (size_t)n * n)) return -4;
            PB_LIB(cusolverDnXgesvdp_bufferSize(sol, par, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, CUDA_C_64F, dTheta.p, m, CUDA_R_64F,
                                                dS.p, CUDA_C_64F, dU.p, m, CUDA_C_64F, dV.p, n, CUDA_C_64F, &wd, &wh),
                   CUSOLVER_STATUS_SUCCESS);
            if (dWork.reserve(wd + 16)) return -4;
            if (work_host.size() < wh + 16) work_host.resize(wh + 16);
            double err_sigma = 0.0;
            PB_LIB(cusolverDnXgesvdp(sol, par, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, CUDA_C_64F, dTheta.p, m, CUDA_R_64F, dS.p,
                                     CUDA_C_64F, dU.p, m, CUDA_C_64F, dV.p, n, CUDA_C_64F, dWork.p, wd, work_host.data(), wh,
                                     (int*)dInfo.p, &err_sigma),
                   CUSOLVER_STATUS_SUCCESS);
            right = transposed ? (const zc*)dU.p : (const zc*)dV.p;   // theta^H = U' S V'^H  =>  theta = V' S U'^H
            ld_right = transposed ? m : n;                            // both equal N
        }
        int info = 0;
        PB_CUDA(cudaMemcpy(&info, dInfo.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (info != 0) {
            set_err("ptbuild: SVD of a %lld x %lld gate did not converge (info = %d)", m, n, info);
            return -6;
        }
        S_host.resize((size_t)mn);
        PB_CUDA(cudaMemcpy(S_host.data(), dS.p, sizeof(double) * mn, cudaMemcpyDeviceToHost));
        ms_svd += now_ms() - t0;
        int keep = 0;
        for (long long i = 0; i < mn; ++i) keep += S_host[(size_t)i] > threshold * S_host[0];
        keep = std::max(1, std::min(keep, chi_max));
        double nrm = 0.0;
        for (int i = 0; i < keep; ++i) nrm += S_host[i] * S_host[i];
        nrm = sqrt(nrm);
        lam_host.assign(S_host.begin(), S_host.begin() + keep);
        for (double& v : lam_host) v /= nrm;
        if (dLam[y].reserve(sizeof(double) * keep)) return -4;
        PB_CUDA(cudaMemcpy(dLam[y].p, lam_host.data(), sizeof(double) * keep, cudaMemcpyHostToDevice));
        // ---- new right site = kept right vectors, new left site = C' . (right vectors) / nrm
        t0 = now_ms();
        if (B[y].buf.reserve(16 * (size_t)keep * N) || B[x].buf.reserve(16 * (size_t)M * keep)) return -4;
        k_take_right<<<(unsigned)(((long long)keep * N + 255) / 256), 256>>>(keep, N, right, ld_right, from_vt, B[y].p());
        PB_CUDA(cudaGetLastError());
        const zc alpha = make_cuDoubleComplex(1.0 / nrm, 0.0);
        PB_LIB(cublasZgemm(blas, CUBLAS_OP_N, from_vt ? CUBLAS_OP_C : CUBLAS_OP_N, (int)M, keep, (int)N, &alpha, (const zc*)dCp.p,
                           (int)M, right, (int)ld_right, &z_zero, B[x].p(), (int)M),
               CUBLAS_STATUS_SUCCESS);
        PB_CUDA(cudaDeviceSynchronize());
        ms_gemm += now_ms() - t0;
        B[y].l = keep;
        B[y].r = chi_r;
        B[x].l = chi_l;
        B[x].r = keep;
        return 0;
    };

    const zc* Wd = (const zc*)dW.p;
    int bond = 0, rc;
    if ((rc = gate(bond, Wd + (size_t)K * d * d, 0))) return rc;
    for (int k = K - 1; k >= 1; --k) {
        bond = 1 - bond;
        if ((rc = gate(bond, Wd + (size_t)k * d * d, 1))) return rc;
    }
    bond = 1 - bond;
    if ((rc = gate(bond, nullptr, 1))) return rc;
    // cap: f[c, l, r] = i0[c] sum_m B_x[l, c, m] B_y[m, c, r]
    const int x = bond, y = 1 - bond;
    const int chi_l = B[x].l, chi_m = B[x].r, chi_r = B[y].r;
    if (chi_l != chi_r || chi_l > chi_max) {
        set_err("ptbuild: uniform tensor is %d x %d (chi_max %d)", chi_l, chi_r, chi_max);
        return -5;
    }
    if (dF.reserve(16 * (size_t)d * chi_l * chi_r)) return -4;
    for (int c = 0; c < d; ++c) {
        const zc alpha = make_cuDoubleComplex(i0[2 * c], i0[2 * c + 1]);
        PB_LIB(cublasZgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, chi_l, chi_r, chi_m, &alpha, B[x].p() + (size_t)chi_l * c, chi_l * d,
                           B[y].p() + (size_t)chi_m * c, chi_m * d, &z_zero, (zc*)dF.p + (size_t)c * chi_l * chi_r, chi_l),
               CUBLAS_STATUS_SUCCESS);
    }
    std::vector<double> F((size_t)2 * d * chi_l * chi_r);
    PB_CUDA(cudaMemcpy(F.data(), dF.p, 16 * (size_t)d * chi_l * chi_r, cudaMemcpyDeviceToHost));
    for (int c = 0; c < d; ++c)
        for (int l = 0; l < chi_l; ++l)
            for (int r = 0; r < chi_r; ++r) {
                const size_t src = 2 * ((size_t)c * chi_l * chi_r + l + (size_t)chi_l * r);
                const size_t dst = 2 * (((size_t)c * chi_max + l) * chi_max + r);
                f_out[dst] = F[src];
                f_out[dst + 1] = F[src + 1];
            }
    *chi_out = chi_l;
    if (stats) {
        stats[0] = now_ms() - t_begin;
        stats[1] = ms_svd;
        stats[2] = ms_gemm;
        stats[3] = max_dim;
    }
    return 0;
}
