/*
 * CPU ORACLE, C restatement (test infrastructure + reported CPU baseline; NOT a product path).
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the external ACE code
 * (github.com/mcygorek/ACE, unpinned by the reference), absent from /root/reference; the
 * reference holds no numeric golden vector.  This file restates SURVEY App. D.2
 * literally -- per step: M1 = exp(L(t_n + off1 dt) dt/2), PT slice, M2 = exp(L(t_n + off2 dt) dt/2),
 * closure, outputs -- as driven by pyaceqd/general_system/general_system.py:227-290
 * (use_symmetric_Trotter :234, add_PT :236, add_Pulse :255/:279, add_Output :289).
 * It is validated against oracle/oracle.py (scipy Pade expm) in tests/test_oracle_c.py.
 *
 * expm: Pade-13 scaling and squaring (Higham 2005) with LU, deliberately a different
 * algorithm from the GPU's Taylor-Horner kernel.  One trajectory per OpenMP thread, like the
 * reference's one-ACE-process-per-job fan-out (two_time/correlations.py:153-170).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

static void matmul(int n, const cplx* a, const cplx* b, cplx* c) {
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) c[i * n + j] = 0;
        for (int k = 0; k < n; ++k) {
            const cplx aik = a[i * n + k];
            for (int j = 0; j < n; ++j) c[i * n + j] += aik * b[k * n + j];
        }
    }
}

/* solve A X = B in place (B <- X), A destroyed; partial pivoting */
static int lu_solve(int n, cplx* a, cplx* b) {
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = cabs(a[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (cabs(a[i * n + k]) > best) { best = cabs(a[i * n + k]); p = i; }
        if (best == 0.0) return -1;
        if (p != k)
            for (int j = 0; j < n; ++j) {
                cplx t = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = t;
                t = b[k * n + j]; b[k * n + j] = b[p * n + j]; b[p * n + j] = t;
            }
        const cplx inv = 1.0 / a[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            const cplx f = a[i * n + k] * inv;
            if (f == 0) continue;
            for (int j = k + 1; j < n; ++j) a[i * n + j] -= f * a[k * n + j];
            for (int j = 0; j < n; ++j) b[i * n + j] -= f * b[k * n + j];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        const cplx inv = 1.0 / a[k * n + k];
        for (int j = 0; j < n; ++j) {
            cplx s = b[k * n + j];
            for (int i = k + 1; i < n; ++i) s -= a[k * n + i] * b[i * n + j];
            b[k * n + j] = s * inv;
        }
    }
    return 0;
}

/* r = exp(a); work: 6*n*n cplx */
static int expm_pade13(int n, const cplx* a, cplx* r, cplx* work) {
    static const double b[14] = {64764752532480000., 32382376266240000., 7771770303897600.,
                                 1187353796428800., 129060195264000., 10559470521600.,
                                 670442572800., 33522128640., 1323241920., 40840800., 960960.,
                                 16380., 182., 1.};
    const int n2 = n * n;
    cplx *A = work, *A2 = A + n2, *A4 = A2 + n2, *A6 = A4 + n2, *U = A6 + n2, *V = U + n2;
    double nrm = 0;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += cabs(a[i * n + j]);
        if (s > nrm) nrm = s;
    }
    int s = 0;
    const double theta13 = 5.371920351148152;
    if (nrm > theta13) s = (int)ceil(log2(nrm / theta13));
    const double sc = ldexp(1.0, -s);
    for (int i = 0; i < n2; ++i) A[i] = a[i] * sc;
    matmul(n, A, A, A2);
    matmul(n, A2, A2, A4);
    matmul(n, A4, A2, A6);
    /* U = A [A6 (b13 A6 + b11 A4 + b9 A2) + b7 A6 + b5 A4 + b3 A2 + b1 I] */
    cplx* T = r; /* scratch */
    for (int i = 0; i < n2; ++i) T[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    matmul(n, A6, T, V);
    for (int i = 0; i < n2; ++i) V[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i];
    for (int i = 0; i < n; ++i) V[i * n + i] += b[1];
    matmul(n, A, V, U);
    /* V = A6 (b12 A6 + b10 A4 + b8 A2) + b6 A6 + b4 A4 + b2 A2 + b0 I */
    for (int i = 0; i < n2; ++i) T[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    matmul(n, A6, T, V);
    for (int i = 0; i < n2; ++i) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i];
    for (int i = 0; i < n; ++i) V[i * n + i] += b[0];
    /* r = (V - U)^-1 (V + U) */
    for (int i = 0; i < n2; ++i) { A2[i] = V[i] - U[i]; r[i] = V[i] + U[i]; }
    if (lu_solve(n, A2, r)) return -1;
    for (int q = 0; q < s; ++q) {
        matmul(n, r, r, A2);
        memcpy(r, A2, sizeof(cplx) * n2);
    }
    return 0;
}

static cplx sample(const cplx* v, int n, double x) {
    if (n <= 0) return 0;
    if (x <= 0) return v[0];
    if (x >= n - 1) return v[n - 1];
    const int j = (int)floor(x);
    const double w = x - j;
    return (1.0 - w) * v[j] + w * v[j + 1];
}

int oracle_c_expm(int n, int count, const double* a, double* out) {
    cplx* work = (cplx*)malloc(sizeof(cplx) * 6 * n * n);
    if (!work) return -1;
    for (int i = 0; i < count; ++i)
        if (expm_pade13(n, (const cplx*)a + (size_t)i * n * n, (cplx*)out + (size_t)i * n * n, work)) {
            free(work);
            return -2;
        }
    free(work);
    return 0;
}

/*
 * Propagate n_traj MTO-free trajectories that start at the PT origin.
 *   tables[set][table][sample] complex;  set_of_traj / n_steps / out_off per trajectory;
 *   out block of a trajectory: [n_steps+1][n_out] complex at out_off (complex elements).
 * Returns 0, or a negative code.  n_threads <= 0 -> OpenMP default.
 */
int oracle_c_propagate(int NL, int n_fields, int n_out, const double* L0_, const double* LA_,
                       const double* LB_, const int* field_table, const double* out_w_,
                       const int* block_of_alpha, int n_cls, int n_slices, int n_initial,
                       const int* chi_in, const int* chi_out, const double* const* slices,
                       const double* const* closures, int n_traj, const int* set_of_traj,
                       const int* n_steps, double dt, double t0, int n_sets, int n_tables,
                       int n_samples, double tab_t0, double tab_dt, const double* tables_,
                       const double* rho0_, double off1, double off2, const long long* out_off,
                       double* out_, int n_threads) {
    (void)n_sets; (void)n_cls;
    const cplx *L0 = (const cplx*)L0_, *LA = (const cplx*)LA_, *LB = (const cplx*)LB_;
    const cplx *ow = (const cplx*)out_w_, *tables = (const cplx*)tables_, *rho0 = (const cplx*)rho0_;
    cplx* out = (cplx*)out_;
    const int n2 = NL * NL;
    int chi_max = 1;
    for (int s = 0; s < n_slices; ++s) {
        if (chi_in[s] > chi_max) chi_max = chi_in[s];
        if (chi_out[s] > chi_max) chi_max = chi_out[s];
    }
    /* split PT slices into re / im planes once (vectorisable inner loops) */
    double** sre = (double**)calloc(n_slices, sizeof(double*));
    double** sim = (double**)calloc(n_slices, sizeof(double*));
    for (int s = 0; s < n_slices; ++s) {
        const size_t cnt = (size_t)n_cls * chi_in[s] * chi_out[s];
        sre[s] = (double*)malloc(sizeof(double) * cnt);
        sim[s] = (double*)malloc(sizeof(double) * cnt);
        for (size_t i = 0; i < cnt; ++i) { sre[s][i] = slices[s][2 * i]; sim[s][i] = slices[s][2 * i + 1]; }
    }
    const int n_rep = n_slices - n_initial;
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        cplx* Lm = (cplx*)malloc(sizeof(cplx) * n2);
        cplx* M = (cplx*)malloc(sizeof(cplx) * n2);
        cplx* work = (cplx*)malloc(sizeof(cplx) * 6 * n2);
        double* xr = (double*)malloc(sizeof(double) * NL * chi_max);
        double* xi = (double*)malloc(sizeof(double) * NL * chi_max);
        double* yr = (double*)malloc(sizeof(double) * NL * chi_max);
        double* yi = (double*)malloc(sizeof(double) * NL * chi_max);
        cplx* rho = (cplx*)malloc(sizeof(cplx) * NL);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < n_traj; ++b) {
            const int set = set_of_traj[b];
            int chi = 1;
            memset(xr, 0, sizeof(double) * NL * chi_max);
            memset(xi, 0, sizeof(double) * NL * chi_max);
            for (int a = 0; a < NL; ++a) { xr[a * chi_max] = creal(rho0[a]); xi[a * chi_max] = cimag(rho0[a]); }
            cplx* ob = out + out_off[b];
            for (int j = 0; j < n_out; ++j) {
                cplx acc = 0;
                for (int a = 0; a < NL; ++a) acc += ow[j * NL + a] * rho0[a];
                ob[j] = acc;
            }
            for (int n = 0; n < n_steps[b]; ++n) {
                const double t_n = t0 + n * dt;
                const int s = n < n_initial ? n : n_initial + (n - n_initial) % n_rep;
                const int din = chi_in[s], dout = chi_out[s];
                for (int half = 0; half < 2; ++half) {
                    const double te = t_n + (half ? off2 : off1) * dt;
                    const double x = (te - tab_t0) / tab_dt;
                    memcpy(Lm, L0, sizeof(cplx) * n2);
                    for (int k = 0; k < n_fields; ++k) {
                        const int tb = field_table[k];
                        if (tb < 0 || tb >= n_tables) continue;
                        const cplx f = sample(tables + ((size_t)set * n_tables + tb) * n_samples, n_samples, x);
                        const cplx fc = conj(f);
                        for (int i = 0; i < n2; ++i) Lm[i] += f * LA[(size_t)k * n2 + i] + fc * LB[(size_t)k * n2 + i];
                    }
                    for (int i = 0; i < n2; ++i) Lm[i] *= 0.5 * dt;
                    if (expm_pade13(NL, Lm, M, work)) { err = -2; }
                    /* state <- M state   (y = M x over the system index) */
                    const int w = half ? dout : chi;
                    for (int a = 0; a < NL; ++a) {
                        double* pr = yr + a * chi_max; double* pi = yi + a * chi_max;
                        for (int d = 0; d < w; ++d) { pr[d] = 0; pi[d] = 0; }
                        for (int k = 0; k < NL; ++k) {
                            const double mr = creal(M[a * NL + k]), mi = cimag(M[a * NL + k]);
                            if (mr == 0 && mi == 0) continue;
                            const double* qr = xr + k * chi_max; const double* qi = xi + k * chi_max;
                            for (int d = 0; d < w; ++d) {
                                pr[d] += mr * qr[d] - mi * qi[d];
                                pi[d] += mr * qi[d] + mi * qr[d];
                            }
                        }
                    }
                    if (half == 0) {
                        /* PT slice: x[a, d2] = sum_d1 y[a, d1] A[blk(a)][d1, d2]; state narrower than din is zero-extended */
                        for (int a = 0; a < NL; ++a) {
                            const double* Ar = sre[s] + (size_t)block_of_alpha[a] * din * dout;
                            const double* Ai = sim[s] + (size_t)block_of_alpha[a] * din * dout;
                            double* pr = xr + a * chi_max; double* pi = xi + a * chi_max;
                            for (int d = 0; d < dout; ++d) { pr[d] = 0; pi[d] = 0; }
                            const int kin = chi < din ? chi : din;
                            for (int d1 = 0; d1 < kin; ++d1) {
                                const double vr = yr[a * chi_max + d1], vi = yi[a * chi_max + d1];
                                const double* br = Ar + (size_t)d1 * dout; const double* bi = Ai + (size_t)d1 * dout;
                                for (int d = 0; d < dout; ++d) {
                                    pr[d] += vr * br[d] - vi * bi[d];
                                    pi[d] += vr * bi[d] + vi * br[d];
                                }
                            }
                        }
                    } else {
                        double* t1 = xr; xr = yr; yr = t1;
                        double* t2 = xi; xi = yi; yi = t2;
                    }
                }
                chi = dout;
                const cplx* q = (const cplx*)closures[s];
                for (int a = 0; a < NL; ++a) {
                    cplx acc = 0;
                    for (int d = 0; d < dout; ++d) acc += (xr[a * chi_max + d] + I * xi[a * chi_max + d]) * q[d];
                    rho[a] = acc;
                }
                for (int j = 0; j < n_out; ++j) {
                    cplx acc = 0;
                    for (int a = 0; a < NL; ++a) acc += ow[j * NL + a] * rho[a];
                    ob[(size_t)(n + 1) * n_out + j] = acc;
                }
            }
        }
        free(Lm); free(M); free(work); free(xr); free(xi); free(yr); free(yi); free(rho);
    }
    for (int s = 0; s < n_slices; ++s) { free(sre[s]); free(sim[s]); }
    free(sre); free(sim);
    return err;
}

int oracle_c_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
