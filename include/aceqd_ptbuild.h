/* aceqd_ptbuild.h -- C ABI of the on-device process-tensor builder (libaceqd_ptbuild.so).
 *
 * Replaces the PT-generation run pyaceqd delegates to the external ACE binary
 * (pyaceqd/general_system/general_system.py:159-192: `Boson_SysOp`, `Boson_J_type QDPhonon` or `Boson_J_from_file`,
 * `temperature`, `threshold`, `use_Gaussian_infinite`, `write_PT`): the uniform (time-translation-invariant) process
 * tensor of a Gaussian bosonic bath with diagonal coupling, contracted by iTEBD on the GPU.  SURVEY 8(f) rank 2.
 *
 * A library of its own because it links cuBLAS (plain ZGEMMs) and cuSOLVER (the SVD of every iTEBD gate), which the
 * propagation library libaceqd.so does not need.  Conventions as in aceqd.h: 0 on success, negative status
 * otherwise, aceqd_ptbuild_last_error() has the message; the caller owns every buffer (host pointers); complex =
 * interleaved double[2].
 */
#ifndef ACEQD_PTBUILD_H
#define ACEQD_PTBUILD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* aceqd_ptbuild_last_error(void);
/* Number of CUDA devices this process can use (0 on a CPU-only box: the builder then refuses to run). */
int aceqd_ptbuild_device_count(void);

/* QUAPI influence coefficients eta_0 .. eta_K of a tabulated spectral density on the device (one CTA per k,
 * trapezoid rule on the caller's frequency grid):
 *   eta_0 = int dw J/w^2 [coth(hbar w/2kT)(1 - cos w dt) + i (sin w dt - w dt)]
 *   eta_k = int dw J/w^2 2(1 - cos w dt) [coth cos(k w dt) - i sin(k w dt)]
 * w, J: [n_w] host arrays (1/ps); hbar_over_2kT in ps (<= 0: zero temperature); eta_out: [K+1] complex. */
int aceqd_ptbuild_eta(int device, int n_w, const double* w, const double* J, double dt, int K,
                      double hbar_over_2kT, double* eta_out);

/* iTEBD contraction of the uniform influence-functional network (what `use_Gaussian_infinite true` asks ACE for,
 * general_system.py:165-167).
 *   d          number of coupling classes (pairs (l+, l-) of coupling eigenvalues)
 *   K          memory steps; weights: [K+1][d][d] complex, weights[k][later][earlier] = I_k(c_later, c_earlier)
 *              (weights[0] is not used: the k = 0 factor enters through i0)
 *   i0         [d] complex, the on-site factor (I_0 diagonal times the polaron-shift phase)
 *   threshold  singular values below threshold * s_max are dropped at every gate; at most chi_max are kept
 *   svd_method 0: cusolverDnXgesvd (QR iteration)   1: cusolverDnXgesvdp (polar decomposition)
 *   f_out      [d][chi_max][chi_max] complex, row-major per class; on return the leading chi x chi corner of
 *              f_out[c] (row stride chi_max) is the one-site tensor f[c, l, r] of the uniform MPS
 *   chi_out    bond dimension reached
 *   stats      optional [4]: total ms, ms inside the SVDs, ms inside ZGEMMs, largest SVD dimension */
int aceqd_ptbuild_uniform(int device, int d, int K, const double* weights, const double* i0, double threshold,
                          int chi_max, int svd_method, double* f_out, int* chi_out, double* stats);

#ifdef __cplusplus
}
#endif
#endif
