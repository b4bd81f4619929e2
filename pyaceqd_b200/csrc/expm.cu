// Batched Liouvillian assembly + matrix exponential (scaling and squaring, Taylor-Horner) and the
// per-step operator builder (SURVEY 2.4 kernels K2/K3).  Replaces ACE's FreePropagator
// (`fprop.update(t, dt); fprop.M`, pyaceqd/general_system/general_system.py:324-327) and the
// apply_Operator bookkeeping of general_system.py:281-286.
//
// For every output row n of a trajectory the step kernel needs two small matrices:
//     V_n  = S_before(t_n) * M2_{n-1}                    (second half step of the previous step)
//     W_n  = M1_n * S_after(t_n) * V_n                   (first half step of the next step)
//     OV_n = out_w * V_n                                  (output functionals pulled through V_n)
// with M1_n = exp(L(t_n + off1*dt) dt/2), M2_{n-1} = exp(L(t_{n-1} + off2*dt) dt/2).  Closure and
// system operators act on different indices, so rho(t_n) = V_n * (Y . q) and the next PT input is
// X = W_n * Y, where Y is the bond state right after the previous PT slice.
//
// One thread group (32..256 threads) owns one entry; matrices live in shared memory.
#include <algorithm>

#include "common.cuh"

namespace aceqd {

namespace {

// exp(A) by scaling and squaring around a degree-12 Taylor polynomial evaluated in
// Paterson-Stockmeyer form (5 matrix products): with ||A/2^s||_1 <= THETA = 0.25 the truncation
// error 0.25^13/13! = 2.4e-18 is below double rounding.
constexpr double THETA = 0.25;
constexpr int EXPM_BUFS = 6;       // A, A2, A3, A4, P0, P1

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

template <int G>
__device__ __forceinline__ void group_sync(int gid) {
    if (G == 16)
        __syncwarp(0xFFFFu << ((gid & 1) * 16));
    else if (G == 32)
        __syncwarp();
    else
        asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(G) : "memory");
}

// Matrix products of a thread group: every thread owns 2 x 2 blocks of C, so that one k step costs 2 + 2 operand
// loads for 4 complex FMAs -- the products of these small matrices are bound by shared-memory traffic (one
// element per thread costs 2 loads per complex FMA: measured 14.5 ms of operator building next to a 23 ms step
// kernel for a biexciton area sweep, profiles/r03x_shapes.jsonl).
// C = scale * A*B (+ I)
template <int G, bool BLK>
__device__ void gmm(double2* C, const double2* A, const double2* B, int n, int tid, int gid,
                    double scale, bool add_identity) {
    if constexpr (!BLK) {   // small matrices: more threads per entry beat operand reuse (measured, r03z)
        for (int e = tid; e < n * n; e += G) {
            const int i = e / n, j = e - i * n;
            double2 acc = make_double2(0.0, 0.0);
            for (int k = 0; k < n; ++k) cfma(acc, A[i * n + k], B[k * n + j]);
            acc.x *= scale;
            acc.y *= scale;
            if (add_identity && i == j) acc.x += 1.0;
            C[e] = acc;
        }
        group_sync<G>(gid);
        return;
    }
    else {
    const int nb = (n + 1) >> 1;
    for (int blk = tid; blk < nb * nb; blk += G) {
        const int i0 = 2 * (blk / nb), j0 = 2 * (blk - (blk / nb) * nb);
        const bool i1 = i0 + 1 < n, j1 = j0 + 1 < n;
        const int ia = i0 * n, ib = (i1 ? i0 + 1 : i0) * n, ja = j0, jb = j1 ? j0 + 1 : j0;
        double2 c00 = make_double2(0.0, 0.0), c01 = c00, c10 = c00, c11 = c00;
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const double2 a0 = A[ia + k], a1 = A[ib + k], b0 = B[k * n + ja], b1 = B[k * n + jb];
            cfma(c00, a0, b0);
            cfma(c01, a0, b1);
            cfma(c10, a1, b0);
            cfma(c11, a1, b1);
        }
        auto put = [&](int i, int j, double2 v) {
            v.x *= scale;
            v.y *= scale;
            if (add_identity && i == j) v.x += 1.0;
            C[i * n + j] = v;
        };
        put(i0, j0, c00);
        if (j1) put(i0, j0 + 1, c01);
        if (i1) put(i0 + 1, j0, c10);
        if (i1 && j1) put(i0 + 1, j0 + 1, c11);
    }
    group_sync<G>(gid);
    }
}

// C = A*B + (c0 I + c1 X1 + c2 X2 + c3 X3)   (any Xi may be null)
template <int G, bool BLK>
__device__ void gmm_poly(double2* C, const double2* A, const double2* B, int n, int tid, int gid,
                         double c0, double c1, const double2* X1, double c2, const double2* X2,
                         double c3, const double2* X3) {
    if constexpr (!BLK) {
        for (int e = tid; e < n * n; e += G) {
            const int i = e / n, j = e - i * n;
            double2 acc = make_double2(i == j ? c0 : 0.0, 0.0);
            if (X1) { acc.x = fma(c1, X1[e].x, acc.x); acc.y = fma(c1, X1[e].y, acc.y); }
            if (X2) { acc.x = fma(c2, X2[e].x, acc.x); acc.y = fma(c2, X2[e].y, acc.y); }
            if (X3) { acc.x = fma(c3, X3[e].x, acc.x); acc.y = fma(c3, X3[e].y, acc.y); }
            for (int k = 0; k < n; ++k) cfma(acc, A[i * n + k], B[k * n + j]);
            C[e] = acc;
        }
        group_sync<G>(gid);
        return;
    }
    else {
    const int nb = (n + 1) >> 1;
    for (int blk = tid; blk < nb * nb; blk += G) {
        const int i0 = 2 * (blk / nb), j0 = 2 * (blk - (blk / nb) * nb);
        const bool i1 = i0 + 1 < n, j1 = j0 + 1 < n;
        const int ia = i0 * n, ib = (i1 ? i0 + 1 : i0) * n, ja = j0, jb = j1 ? j0 + 1 : j0;
        auto init = [&](int i, int j) {
            const int e = i * n + j;
            double2 acc = make_double2(i == j ? c0 : 0.0, 0.0);
            if (X1) { acc.x = fma(c1, X1[e].x, acc.x); acc.y = fma(c1, X1[e].y, acc.y); }
            if (X2) { acc.x = fma(c2, X2[e].x, acc.x); acc.y = fma(c2, X2[e].y, acc.y); }
            if (X3) { acc.x = fma(c3, X3[e].x, acc.x); acc.y = fma(c3, X3[e].y, acc.y); }
            return acc;
        };
        double2 c00 = init(i0, j0), c01 = init(i0, jb), c10 = init(i1 ? i0 + 1 : i0, j0),
                c11 = init(i1 ? i0 + 1 : i0, jb);
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const double2 a0 = A[ia + k], a1 = A[ib + k], b0 = B[k * n + ja], b1 = B[k * n + jb];
            cfma(c00, a0, b0);
            cfma(c01, a0, b1);
            cfma(c10, a1, b0);
            cfma(c11, a1, b1);
        }
        C[i0 * n + j0] = c00;
        if (j1) C[i0 * n + j0 + 1] = c01;
        if (i1) C[(i0 + 1) * n + j0] = c10;
        if (i1 && j1) C[(i0 + 1) * n + j0 + 1] = c11;
    }
    group_sync<G>(gid);
    }
}

// exp(A) for the n x n matrix in buf[0..n2) (destroyed).  `buf` holds EXPM_BUFS matrices; the
// result pointer is one of them.  `red` is a scratch of >= n doubles.
template <int G, bool BLK>
__device__ double2* expm_group(double2* buf, double* red, int n, int tid, int gid) {
    const int n2 = n * n;
    double2 *A = buf, *A2 = A + n2, *A3 = A2 + n2, *A4 = A3 + n2, *P0 = A4 + n2, *P1 = P0 + n2;
    for (int j = tid; j < n; j += G) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double2 v = A[i * n + j];
            s += sqrt(v.x * v.x + v.y * v.y);
        }
        red[j] = s;
    }
    group_sync<G>(gid);
    double nrm = 0.0;
    for (int j = 0; j < n; ++j) nrm = fmax(nrm, red[j]);
    int s = 0;
    if (nrm > THETA) {
        int ex;
        frexp(nrm / THETA, &ex);  // nrm/THETA = m * 2^ex, m in [0.5, 1)
        s = ex > 60 ? 60 : ex;
    }
    const double sc = ldexp(1.0, -s);
    group_sync<G>(gid);  // everyone has read `red`
    for (int e = tid; e < n2; e += G) {
        double2 v = A[e];
        A[e] = make_double2(v.x * sc, v.y * sc);
    }
    group_sync<G>(gid);
    constexpr double c2 = 1.0 / 2, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720,
                     c7 = 1.0 / 5040, c8 = 1.0 / 40320, c9 = 1.0 / 362880, c10 = 1.0 / 3628800,
                     c11 = 1.0 / 39916800, c12 = 1.0 / 479001600;
    gmm_poly<G, BLK>(A2, A, A, n, tid, gid, 0.0, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr);
    gmm_poly<G, BLK>(A3, A2, A, n, tid, gid, 0.0, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr);
    // P0 = c8 I + c9 A + c10 A2 + c11 A3 + c12 A4 ,  A4 = A2*A2  (one product gives both)
    gmm_poly<G, BLK>(A4, A2, A2, n, tid, gid, 0.0, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr);
    for (int e = tid; e < n2; e += G) {
        const int i = e / n, j = e - i * n;
        double2 v = make_double2(i == j ? c8 : 0.0, 0.0);
        v.x += c9 * A[e].x + c10 * A2[e].x + c11 * A3[e].x + c12 * A4[e].x;
        v.y += c9 * A[e].y + c10 * A2[e].y + c11 * A3[e].y + c12 * A4[e].y;
        P0[e] = v;
    }
    group_sync<G>(gid);
    // P1 = (c4 I + c5 A + c6 A2 + c7 A3) + A4 P0 ;  P0 = (I + A + c2 A2 + c3 A3) + A4 P1
    gmm_poly<G, BLK>(P1, A4, P0, n, tid, gid, c4, c5, A, c6, A2, c7, A3);
    gmm_poly<G, BLK>(P0, A4, P1, n, tid, gid, 1.0, 1.0, A, c2, A2, c3, A3);
    double2* cur = P0;
    double2* nxt = P1;
    for (int q = 0; q < s; ++q) {
        gmm_poly<G, BLK>(nxt, cur, cur, n, tid, gid, 0.0, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr);
        double2* t = cur;
        cur = nxt;
        nxt = t;
    }
    return cur;
}

__device__ __forceinline__ double2 sample_table(const double2* v, int n, double x) {
    if (n <= 0) return make_double2(0.0, 0.0);
    if (x <= 0.0) return v[0];
    if (x >= (double)(n - 1)) return v[n - 1];
    const int j = (int)floor(x);
    const double w = x - j;
    const double2 a = v[j], b = v[j + 1];
    return make_double2((1.0 - w) * a.x + w * b.x, (1.0 - w) * a.y + w * b.y);
}

// A = (L0 + sum_k f_k LA_k + conj(f_k) LB_k) * delta   at time t, drive set `set`
template <int G>
__device__ void assemble(double2* A, const OpBuildParams& p, int set, double t, double delta,
                         int tid, int gid, int ns) {
    const int n = p.prob.NL, n2 = n * n;
    const double2* L0 = reinterpret_cast<const double2*>(p.prob.L0);
    const double2* LA = reinterpret_cast<const double2*>(p.prob.LA);
    const double2* LB = reinterpret_cast<const double2*>(p.prob.LB);
    const double2* tabs = reinterpret_cast<const double2*>(p.tables);
    const double x = (t - p.tab_t0) / p.tab_dt;
    for (int e = tid; e < n2; e += G) {
        double2 acc = L0[e];
        for (int k = 0; k < p.prob.n_fields; ++k) {
            const int tb = p.prob.field_table[k];
            if (tb < 0 || tb >= p.n_tables) continue;
            const double2 f =
                sample_table(tabs + ((size_t)set * p.n_tables + tb) * p.n_samples, ns, x);
            cfma(acc, f, LA[(size_t)k * n2 + e]);
            cfma(acc, make_double2(f.x, -f.y), LB[(size_t)k * n2 + e]);
        }
        A[e] = make_double2(acc.x * delta, acc.y * delta);
    }
    group_sync<G>(gid);
}

template <int G, bool BLK>
__global__ void __launch_bounds__(256) k_opbuild(OpBuildParams p) {
    extern __shared__ double2 sm[];
    const int n = p.prob.NL, n2 = n * n;
    const int groups = blockDim.x / G;
    const int gid = threadIdx.x / G, tid = threadIdx.x - gid * G;
    const size_t per_group = (size_t)(EXPM_BUFS + 2) * n2 + MAX_NL;  // expm buffers, V, X + scratch
    // large Liouville spaces (NL > ~40): the work matrices live in a per-CTA slab of global memory
    double2* base = p.scratch ? reinterpret_cast<double2*>(p.scratch) + (size_t)blockIdx.x * groups * per_group : sm;
    double2* A = base + gid * per_group;
    double2* V = A + (size_t)EXPM_BUFS * n2;
    double2* X = V + n2;
    double* red = reinterpret_cast<double*>(X + n2);

    const long long total = p.e_end > p.e_begin ? p.e_end : p.n_seq_entries + p.n_entries;
    const double half = 0.5 * p.dt;
    for (long long e = p.e_begin + (long long)blockIdx.x * groups + gid; e < total;
         e += (long long)gridDim.x * groups) {
        int set, step, sb = -1, sa = -1, has_prev, ns = p.n_samples;
        if (e < p.n_seq_entries) {
            int lo = 0, hi = p.n_seq;  // seq_base[lo] <= e < seq_base[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (p.seq_base[mid] <= e) lo = mid; else hi = mid;
            }
            const aceqd_seq sq = p.seqs[lo];
            const int i = (int)(e - p.seq_base[lo]);
            set = sq.set;
            step = sq.step0 + i;
            has_prev = (i > 0) || sq.first_has_prev;
        } else {
            const aceqd_entry en = p.entries[e - p.n_seq_entries];
            set = en.set; step = en.step; sb = en.sb; sa = en.sa; has_prev = en.has_prev;
            if (en.clamp > 0) ns = min(en.clamp, p.n_samples);   // this row sees only the first `clamp` samples of its drive
        }
        const double t_n = p.t0 + (double)step * p.dt;
        const double2* mto = reinterpret_cast<const double2*>(p.mto_mats);

        // ---- V = Sb * M2_{n-1}
        if (has_prev) {
            assemble<G>(A, p, set, t_n - p.dt + p.eval_off2 * p.dt, half, tid, gid, ns);
            double2* M2 = expm_group<G, BLK>(A, red, n, tid, gid);
            if (sb >= 0) {
                gmm<G, BLK>(V, mto + (size_t)sb * n2, M2, n, tid, gid, 1.0, false);
            } else {
                for (int q = tid; q < n2; q += G) V[q] = M2[q];
                group_sync<G>(gid);
            }
        } else {
            for (int q = tid; q < n2; q += G) {
                const int i = q / n, j = q - i * n;
                V[q] = (sb >= 0) ? mto[(size_t)sb * n2 + q] : make_double2(i == j ? 1.0 : 0.0, 0.0);
            }
            group_sync<G>(gid);
        }
        // ---- OV = out_w * V
        {
            const double2* ow = reinterpret_cast<const double2*>(p.prob.out_w);
            double2* ov = reinterpret_cast<double2*>(p.OV + (size_t)e * p.prob.ov_doubles);
            const int cnt = p.prob.n_out * n;
            for (int q = tid; q < cnt; q += G) {
                const int j = q / n, a = q - j * n;
                double2 acc = make_double2(0.0, 0.0);
                for (int k = 0; k < n; ++k) cfma(acc, ow[j * n + k], V[k * n + a]);
                ov[q] = acc;
            }
        }
        // ---- X = Sa * V
        const double2* Xp = V;
        if (sa >= 0) {
            gmm<G, BLK>(X, mto + (size_t)sa * n2, V, n, tid, gid, 1.0, false);
            Xp = X;
        }
        // ---- W = M1_n * X   (zero padded to [NLp8][NLp4])
        assemble<G>(A, p, set, t_n + p.eval_off1 * p.dt, half, tid, gid, ns);
        double2* M1 = expm_group<G, BLK>(A, red, n, tid, gid);
        {
            double2* w = reinterpret_cast<double2*>(p.W + (size_t)e * p.prob.w_doubles);
            const int ld = p.prob.NLp4, cnt = p.prob.NLp8 * ld;
            for (int q = tid; q < cnt; q += G) {
                const int i = q / ld, j = q - i * ld;
                double2 acc = make_double2(0.0, 0.0);
                if (i < n && j < n)
                    for (int k = 0; k < n; ++k) cfma(acc, M1[i * n + k], Xp[k * n + j]);
                w[q] = acc;
            }
        }
        group_sync<G>(gid);  // buffers are reused by the next entry
    }
}

// -------------------------------------------------------------------------------------------
// Register-resident operator builder for small Liouville spaces (NL = N <= 4: two-level emitters, i.e. the
// pulse-parameter sweeps of two_level_system/rabi_rotations.py:172-198).  The group kernel above is bound by
// shared-memory traffic there (one LDS.128 per complex FMA); here ONE THREAD owns one entry, the matrices live in
// registers and the exponential is the same degree-12 Taylor polynomial, evaluated by Horner's rule with RIGHT
// multiplications   P <- I + (P A)/k ,  k = 12 .. 1   so that row i of the new P needs only row i of the old one
// (in place, one row of temporaries).
template <int N>
__device__ __forceinline__ void expm_reg(double2 (&A)[N][N], double2 (&P)[N][N]) {
    double nrm = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double cs = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) cs += sqrt(A[i][j].x * A[i][j].x + A[i][j].y * A[i][j].y);
        nrm = fmax(nrm, cs);
    }
    int s = 0;
    if (nrm > THETA) {
        int ex;
        frexp(nrm / THETA, &ex);
        s = ex > 60 ? 60 : ex;
    }
    const double sc = ldexp(1.0, -s);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            A[i][j].x *= sc;
            A[i][j].y *= sc;
            P[i][j] = make_double2(A[i][j].x * (1.0 / 12.0) + (i == j ? 1.0 : 0.0), A[i][j].y * (1.0 / 12.0));
        }
#pragma unroll 1
    for (int k = 11; k >= 1; --k) {
        const double inv = 1.0 / (double)k;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double2 row[N];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int m = 0; m < N; ++m) cfma(acc, P[i][m], A[m][j]);
                row[j] = acc;
            }
#pragma unroll
            for (int j = 0; j < N; ++j)
                P[i][j] = make_double2(row[j].x * inv + (i == j ? 1.0 : 0.0), row[j].y * inv);
        }
    }
#pragma unroll 1
    for (int q = 0; q < s; ++q) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) A[i][j] = P[i][j];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int m = 0; m < N; ++m) cfma(acc, A[i][m], A[m][j]);
                P[i][j] = acc;
            }
    }
}

template <int N>
__device__ __forceinline__ void assemble_reg(double2 (&A)[N][N], const OpBuildParams& p, int set, double t,
                                             double delta, int ns) {
    const double2* L0 = reinterpret_cast<const double2*>(p.prob.L0);
    const double2* LA = reinterpret_cast<const double2*>(p.prob.LA);
    const double2* LB = reinterpret_cast<const double2*>(p.prob.LB);
    const double2* tabs = reinterpret_cast<const double2*>(p.tables);
    const double x = (t - p.tab_t0) / p.tab_dt;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) A[i][j] = __ldg(L0 + i * N + j);
    for (int k = 0; k < p.prob.n_fields; ++k) {
        const int tb = p.prob.field_table[k];
        if (tb < 0 || tb >= p.n_tables) continue;
        const double2 f = sample_table(tabs + ((size_t)set * p.n_tables + tb) * p.n_samples, ns, x);
        const double2 fc = make_double2(f.x, -f.y);
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                cfma(A[i][j], f, __ldg(LA + (size_t)k * N * N + i * N + j));
                cfma(A[i][j], fc, __ldg(LB + (size_t)k * N * N + i * N + j));
            }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            A[i][j].x *= delta;
            A[i][j].y *= delta;
        }
}

constexpr int OPREG_THREADS = 128;

template <int N>
__global__ void __launch_bounds__(OPREG_THREADS) k_opbuild_reg(OpBuildParams p) {
    __shared__ double2 park[N * N * OPREG_THREADS];   // X = Sa V of every thread, element-major (conflict-free)
    const long long total = p.e_end > p.e_begin ? p.e_end : p.n_seq_entries + p.n_entries;
    const double half = 0.5 * p.dt;
    const double2* mto = reinterpret_cast<const double2*>(p.mto_mats);
    double2* mine = park + threadIdx.x;
    const int ld = p.prob.NLp4, n_out = p.prob.n_out;
    for (long long e = p.e_begin + (long long)blockIdx.x * OPREG_THREADS + threadIdx.x; e < total;
         e += (long long)gridDim.x * OPREG_THREADS) {
        int set, step, sb = -1, sa = -1, has_prev, ns = p.n_samples;
        if (e < p.n_seq_entries) {
            int lo = 0, hi = p.n_seq;  // seq_base[lo] <= e < seq_base[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (p.seq_base[mid] <= e) lo = mid; else hi = mid;
            }
            const aceqd_seq sq = p.seqs[lo];
            const int i = (int)(e - p.seq_base[lo]);
            set = sq.set;
            step = sq.step0 + i;
            has_prev = (i > 0) || sq.first_has_prev;
        } else {
            const aceqd_entry en = p.entries[e - p.n_seq_entries];
            set = en.set; step = en.step; sb = en.sb; sa = en.sa; has_prev = en.has_prev;
            if (en.clamp > 0) ns = min(en.clamp, p.n_samples);   // this row sees only the first `clamp` samples of its drive
        }
        const double t_n = p.t0 + (double)step * p.dt;
        double2 A[N][N], P[N][N];
        // ---- V = Sb * M2_{n-1}
        if (has_prev) {
            assemble_reg<N>(A, p, set, t_n - p.dt + p.eval_off2 * p.dt, half, ns);
            expm_reg<N>(A, P);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) P[i][j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
        }
        if (sb >= 0) {   // rare (multi-time operator rows)
            const double2* S = mto + (size_t)sb * N * N;
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                    for (int m = 0; m < N; ++m) cfma(acc, S[i * N + m], P[m][j]);
                    A[i][j] = acc;
                }
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) P[i][j] = A[i][j];
        }
        // ---- OV = out_w * V
        {
            const double2* ow = reinterpret_cast<const double2*>(p.prob.out_w);
            double2* ov = reinterpret_cast<double2*>(p.OV + (size_t)e * p.prob.ov_doubles);
            for (int j = 0; j < n_out; ++j) {
                double2 o4[N];
#pragma unroll
                for (int a = 0; a < N; ++a) o4[a] = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double2 w = __ldg(ow + j * N + k);
#pragma unroll
                    for (int a = 0; a < N; ++a) cfma(o4[a], w, P[k][a]);
                }
#pragma unroll
                for (int a = 0; a < N; ++a) ov[j * N + a] = o4[a];
            }
        }
        // ---- X = Sa * V  -> parked in shared memory while M1 is built
        if (sa >= 0) {
            const double2* S = mto + (size_t)sa * N * N;
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                    for (int m = 0; m < N; ++m) cfma(acc, S[i * N + m], P[m][j]);
                    mine[(i * N + j) * OPREG_THREADS] = acc;
                }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) mine[(i * N + j) * OPREG_THREADS] = P[i][j];
        }
        // ---- W = M1_n * X   (zero padded to [NLp8][NLp4])
        assemble_reg<N>(A, p, set, t_n + p.eval_off1 * p.dt, half, ns);
        expm_reg<N>(A, P);
        double2* w = reinterpret_cast<double2*>(p.W + (size_t)e * p.prob.w_doubles);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double2 xc[N];
#pragma unroll
            for (int k = 0; k < N; ++k) xc[k] = mine[(k * N + j) * OPREG_THREADS];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) cfma(acc, P[i][k], xc[k]);
                A[i][j] = acc;
            }
        }
        // NLp8 x NLp4 = 8 x 4 for N = 4 (checked by the launcher)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) w[i * ld + j] = i < N ? A[i < N ? i : 0][j] : make_double2(0.0, 0.0);
    }
}

// -------------------------------------------------------------------------------------------
// Tensor-core operator builder for mid-size Liouville spaces (4 < NL <= 16: three-level models, the biexciton of
// four_level_system/tpe_rotations.py:182-207).  ONE WARP owns an entry; matrices live in shared memory as split
// re/im planes [NPAD][NPAD + 4] (zero padded to a multiple of 8), and every product of the scaling-and-squaring
// chain is a complex DMMA.8x8x4 GEMM of the warp: C fragments go back to shared memory and return as A / B
// fragments of the next product after a __syncwarp().  Same polynomial and scaling rule as expm_group.
constexpr int WM_BUFS = 7;    // A, A2, A3, A4, P0, P1, V
constexpr int WM_MAX_WARPS = 4;   // warps (entries in flight) per CTA: one per SM sub-partition

__device__ __forceinline__ void dmma8(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

template <int NP>   // NP = NPAD / 8 tiles per dimension
struct WMat {
    static constexpr int NPAD = 8 * NP, LD = NPAD + 4, PLANE = NPAD * LD;
    double* re;
    __device__ __forceinline__ double* im() const { return re + PLANE; }
};

// C = A * B + Init, all NPAD x NPAD complex.  init(i, j, &re, &im) supplies the additive term of element (i, j).
template <int NP, class Init>
__device__ __forceinline__ void wmm(WMat<NP> C, WMat<NP> A, WMat<NP> B, int lane, Init init) {
    constexpr int LD = WMat<NP>::LD, NPAD = WMat<NP>::NPAD;
    const int g = lane >> 2, tq = lane & 3;
    double cr[NP][NP][2], ci[NP][NP][2];
#pragma unroll
    for (int mi = 0; mi < NP; ++mi)
#pragma unroll
        for (int ni = 0; ni < NP; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) init(8 * mi + g, 8 * ni + 2 * tq + h, cr[mi][ni][h], ci[mi][ni][h]);
#pragma unroll
    for (int ks = 0; ks < NPAD / 4; ++ks) {
        double ar[NP], ai[NP], br[NP], bi[NP];
#pragma unroll
        for (int mi = 0; mi < NP; ++mi) {
            ar[mi] = A.re[(8 * mi + g) * LD + 4 * ks + tq];
            ai[mi] = A.im()[(8 * mi + g) * LD + 4 * ks + tq];
        }
#pragma unroll
        for (int ni = 0; ni < NP; ++ni) {
            br[ni] = B.re[(4 * ks + tq) * LD + 8 * ni + g];
            bi[ni] = B.im()[(4 * ks + tq) * LD + 8 * ni + g];
        }
#pragma unroll
        for (int mi = 0; mi < NP; ++mi)
#pragma unroll
            for (int ni = 0; ni < NP; ++ni) {
                dmma8(cr[mi][ni][0], cr[mi][ni][1], ar[mi], br[ni]);
                dmma8(ci[mi][ni][0], ci[mi][ni][1], ar[mi], bi[ni]);
            }
#pragma unroll
        for (int mi = 0; mi < NP; ++mi)
#pragma unroll
            for (int ni = 0; ni < NP; ++ni) {
                dmma8(cr[mi][ni][0], cr[mi][ni][1], -ai[mi], bi[ni]);
                dmma8(ci[mi][ni][0], ci[mi][ni][1], ai[mi], br[ni]);
            }
    }
#pragma unroll
    for (int mi = 0; mi < NP; ++mi)
#pragma unroll
        for (int ni = 0; ni < NP; ++ni) {
            const int o = (8 * mi + g) * LD + 8 * ni + 2 * tq;
            *reinterpret_cast<double2*>(C.re + o) = make_double2(cr[mi][ni][0], cr[mi][ni][1]);
            *reinterpret_cast<double2*>(C.im() + o) = make_double2(ci[mi][ni][0], ci[mi][ni][1]);
        }
    __syncwarp();
}

// plane offsets of the (at most 8) elements e = lane + 32 k < n*n a lane touches in element-wise passes (no divisions there)
struct LaneOffs {
    int o[8];
    bool diag[8];
    int cnt;
};

// exp(A) (A destroyed); buffers M[0..5] = A, A2, A3, A4, P0, P1; returns the index of the result buffer (4 or 5)
template <int NP>
__device__ int expm_warp(WMat<NP>* M, int n, int lane, const LaneOffs& lo) {
    constexpr int LD = WMat<NP>::LD;
    WMat<NP> A = M[0], A2 = M[1], A3 = M[2], A4 = M[3], P0 = M[4], P1 = M[5];
    double cs = 0.0;
    if (lane < n)
        for (int i = 0; i < n; ++i) {
            const double x = A.re[i * LD + lane], y = A.im()[i * LD + lane];
            cs += fabs(x) + fabs(y);     // >= |a|, <= sqrt(2) |a|: a bound is all the scaling rule needs, and a
                                         // double-precision square root is ~40 instructions on the only warp of its sub-partition
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cs = fmax(cs, __shfl_xor_sync(0xffffffffu, cs, o));
    int s = 0;
    if (cs > THETA) {
        int ex;
        frexp(cs / THETA, &ex);
        s = ex > 60 ? 60 : ex;
    }
    const double sc = ldexp(1.0, -s);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < lo.cnt) {
            A.re[lo.o[k]] *= sc;
            A.im()[lo.o[k]] *= sc;
        }
    __syncwarp();
    constexpr double c2 = 1.0 / 2, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720,
                     c7 = 1.0 / 5040, c8 = 1.0 / 40320, c9 = 1.0 / 362880, c10 = 1.0 / 3628800,
                     c11 = 1.0 / 39916800, c12 = 1.0 / 479001600;
    auto zero = [](int, int, double& r, double& i) { r = 0.0; i = 0.0; };
    wmm<NP>(A2, A, A, lane, zero);
    wmm<NP>(A3, A2, A, lane, zero);
    wmm<NP>(A4, A2, A2, lane, zero);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < lo.cnt) {
            const int o = lo.o[k];
            P0.re[o] = (lo.diag[k] ? c8 : 0.0) + c9 * A.re[o] + c10 * A2.re[o] + c11 * A3.re[o] + c12 * A4.re[o];
            P0.im()[o] = c9 * A.im()[o] + c10 * A2.im()[o] + c11 * A3.im()[o] + c12 * A4.im()[o];
        }
    __syncwarp();
    // P1 = (c4 I + c5 A + c6 A2 + c7 A3) + A4 P0 ;  P0 = (I + A + c2 A2 + c3 A3) + A4 P1
    wmm<NP>(P1, A4, P0, lane, [&](int i, int j, double& r, double& im_) {
        const int o = i * LD + j;
        r = ((i == j && i < n) ? c4 : 0.0) + c5 * A.re[o] + c6 * A2.re[o] + c7 * A3.re[o];
        im_ = c5 * A.im()[o] + c6 * A2.im()[o] + c7 * A3.im()[o];
    });
    wmm<NP>(P0, A4, P1, lane, [&](int i, int j, double& r, double& im_) {
        const int o = i * LD + j;
        r = ((i == j && i < n) ? 1.0 : 0.0) + A.re[o] + c2 * A2.re[o] + c3 * A3.re[o];
        im_ = A.im()[o] + c2 * A2.im()[o] + c3 * A3.im()[o];
    });
    int cur = 4, nxt = 5;
    for (int q = 0; q < s; ++q) {
        wmm<NP>(M[nxt], M[cur], M[cur], lane, zero);
        const int t = cur;
        cur = nxt;
        nxt = t;
    }
    return cur;
}

// Lsm: L0 | LA[0..n_fields) | LB[0..n_fields) as row-major complex n x n (shared memory copy, or the global arrays)
template <int NP>
__device__ void assemble_warp(WMat<NP> A, const OpBuildParams& p, const double2* L0, const double2* LA,
                              const double2* LB, int set, double t, double delta, int lane, int ns) {
    constexpr int LD = WMat<NP>::LD;
    const int n = p.prob.NL, n2 = n * n, nf = p.prob.n_fields;
    const double2* tabs = reinterpret_cast<const double2*>(p.tables);
    const double x = (t - p.tab_t0) / p.tab_dt;
    // lane k samples drive field k once for the whole matrix
    double2 fl = make_double2(0.0, 0.0);
    if (lane < nf) {
        const int tb = p.prob.field_table[lane];
        if (tb >= 0 && tb < p.n_tables)
            fl = sample_table(tabs + ((size_t)set * p.n_tables + tb) * p.n_samples, ns, x);
    }
    for (int e0 = 0; e0 < n2; e0 += 32) {          // warp-uniform trip count: the shuffles below need every lane
        const int e = e0 + lane;
        const bool on = e < n2;
        double2 acc = on ? L0[e] : make_double2(0.0, 0.0);
        for (int k = 0; k < nf; ++k) {
            const double2 f = make_double2(__shfl_sync(0xffffffffu, fl.x, k), __shfl_sync(0xffffffffu, fl.y, k));
            if (on) {
                cfma(acc, f, LA[(size_t)k * n2 + e]);
                cfma(acc, make_double2(f.x, -f.y), LB[(size_t)k * n2 + e]);
            }
        }
        if (on) {
            const int i = e / n;
            const int o = i * LD + (e - i * n);
            A.re[o] = acc.x * delta;
            A.im()[o] = acc.y * delta;
        }
    }
    __syncwarp();
}

template <int NP>
__global__ void __launch_bounds__(32 * WM_MAX_WARPS) k_opbuild_dmma(OpBuildParams p, int warps_per_cta, int lsm) {
    extern __shared__ __align__(16) double wm_smem[];
    constexpr int LD = WMat<NP>::LD, PLANE = WMat<NP>::PLANE;
    const int n = p.prob.NL, n2 = n * n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* base = wm_smem + (size_t)warp * WM_BUFS * 2 * PLANE;
    WMat<NP> M[WM_BUFS];
#pragma unroll
    for (int b = 0; b < WM_BUFS; ++b) M[b].re = base + (size_t)b * 2 * PLANE;
    for (int e = lane; e < WM_BUFS * 2 * PLANE; e += 32) base[e] = 0.0;    // the zero padding stays zero throughout
    // Liouvillian pieces: one copy per CTA in shared memory when it fits (they are read twice per entry)
    const double2* L0 = reinterpret_cast<const double2*>(p.prob.L0);
    const double2* LA = reinterpret_cast<const double2*>(p.prob.LA);
    const double2* LB = reinterpret_cast<const double2*>(p.prob.LB);
    if (lsm) {
        double2* ls = reinterpret_cast<double2*>(wm_smem + (size_t)warps_per_cta * WM_BUFS * 2 * PLANE);
        const int nf = p.prob.n_fields;
        for (int e = threadIdx.x; e < n2; e += blockDim.x) ls[e] = L0[e];
        for (int e = threadIdx.x; e < nf * n2; e += blockDim.x) {
            ls[n2 + e] = LA[e];
            ls[(1 + nf) * n2 + e] = LB[e];
        }
        L0 = ls;
        LA = ls + n2;
        LB = ls + (size_t)(1 + nf) * n2;
        __syncthreads();
    }
    __syncwarp();
    WMat<NP> V = M[6];
    LaneOffs lo;
    lo.cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int e = lane + 32 * k;
        const int i = e / n, j = e - i * n;
        lo.o[k] = e < n2 ? i * LD + j : 0;
        lo.diag[k] = e < n2 && i == j;
        if (e < n2) lo.cnt = k + 1;
    }
    auto load_global = [&](WMat<NP> D, const double2* src) {    // row-major complex n x n -> planes
        for (int e = lane; e < n2; e += 32) {
            const int o = (e / n) * LD + e % n;
            D.re[o] = src[e].x;
            D.im()[o] = src[e].y;
        }
        __syncwarp();
    };
    auto copy = [&](WMat<NP> D, WMat<NP> S) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < lo.cnt) {
                D.re[lo.o[k]] = S.re[lo.o[k]];
                D.im()[lo.o[k]] = S.im()[lo.o[k]];
            }
        __syncwarp();
    };
    auto zero = [](int, int, double& r, double& i) { r = 0.0; i = 0.0; };
    const long long total = p.e_end > p.e_begin ? p.e_end : p.n_seq_entries + p.n_entries;
    const double half = 0.5 * p.dt;
    const double2* mto = reinterpret_cast<const double2*>(p.mto_mats);
    for (long long e = p.e_begin + (long long)blockIdx.x * warps_per_cta + warp; e < total;
         e += (long long)gridDim.x * warps_per_cta) {
        int set, step, sb = -1, sa = -1, has_prev, ns = p.n_samples;
        if (e < p.n_seq_entries) {
            int lo = 0, hi = p.n_seq;  // seq_base[lo] <= e < seq_base[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (p.seq_base[mid] <= e) lo = mid; else hi = mid;
            }
            const aceqd_seq sq = p.seqs[lo];
            const int i = (int)(e - p.seq_base[lo]);
            set = sq.set;
            step = sq.step0 + i;
            has_prev = (i > 0) || sq.first_has_prev;
        } else {
            const aceqd_entry en = p.entries[e - p.n_seq_entries];
            set = en.set; step = en.step; sb = en.sb; sa = en.sa; has_prev = en.has_prev;
            if (en.clamp > 0) ns = min(en.clamp, p.n_samples);   // this row sees only the first `clamp` samples of its drive
        }
        const double t_n = p.t0 + (double)step * p.dt;
        // ---- V = Sb * M2_{n-1}
        if (has_prev) {
            assemble_warp<NP>(M[0], p, L0, LA, LB, set, t_n - p.dt + p.eval_off2 * p.dt, half, lane, ns);
            const int r = expm_warp<NP>(M, n, lane, lo);
            if (sb >= 0) {
                load_global(M[0], mto + (size_t)sb * n2);
                wmm<NP>(V, M[0], M[r], lane, zero);
            } else {
                copy(V, M[r]);
            }
        } else {
            for (int q = lane; q < n2; q += 32) {
                const int i = q / n, j = q - i * n, o = i * LD + j;
                const double2 v = (sb >= 0) ? mto[(size_t)sb * n2 + q] : make_double2(i == j ? 1.0 : 0.0, 0.0);
                V.re[o] = v.x;
                V.im()[o] = v.y;
            }
            __syncwarp();
        }
        // ---- OV = out_w * V
        {
            const double2* ow = reinterpret_cast<const double2*>(p.prob.out_w);
            double2* ov = reinterpret_cast<double2*>(p.OV + (size_t)e * p.prob.ov_doubles);
            const int cnt = p.prob.n_out * n;
            for (int q = lane; q < cnt; q += 32) {
                const int j = q / n, a = q - j * n;
                double2 acc = make_double2(0.0, 0.0);
                for (int k = 0; k < n; ++k) cfma(acc, ow[j * n + k], make_double2(V.re[k * LD + a], V.im()[k * LD + a]));
                ov[q] = acc;
            }
        }
        // ---- X = Sa * V (in V's place, through a work buffer: it must outlive the second exponential)
        if (sa >= 0) {
            __syncwarp();     // OV has read V
            load_global(M[0], mto + (size_t)sa * n2);
            wmm<NP>(M[1], M[0], V, lane, zero);
            copy(V, M[1]);
        }
        // ---- W = M1_n * X   (zero padded to [NLp8][NLp4])
        assemble_warp<NP>(M[0], p, L0, LA, LB, set, t_n + p.eval_off1 * p.dt, half, lane, ns);
        const int r1 = expm_warp<NP>(M, n, lane, lo);
        WMat<NP> Wm = M[r1 == 4 ? 5 : 4];
        wmm<NP>(Wm, M[r1], V, lane, zero);
        {
            double2* w = reinterpret_cast<double2*>(p.W + (size_t)e * p.prob.w_doubles);
            const int ld = p.prob.NLp4, cnt = p.prob.NLp8 * ld;
            for (int q = lane; q < cnt; q += 32) {
                const int i = q / ld, j = q - i * ld;
                w[q] = (i < n && j < n) ? make_double2(Wm.re[i * LD + j], Wm.im()[i * LD + j]) : make_double2(0.0, 0.0);
            }
        }
        __syncwarp();  // buffers are reused by the next entry
    }
}

// -------------------------------------------------------------------------------------------
// Tensor-core operator builder for the large Liouville spaces (16 < NL <= 40: the five- and six-level models,
// pyaceqd/four_level_system/dark_model.py:34-55, six_level_system/linear.py:28-72).  ONE CTA of 16 warps owns an entry:
// the matrices of the scaling-and-squaring chain live in shared memory as split re/im planes [NPAD][NPAD + 4] (seven of
// them: 197 KB for NL = 36), every product is a complex DMMA.8x8x4 GEMM whose 8x8 output tiles are dealt round-robin to
// the warps (tile t -> warp t mod 16: the four warps of a sub-partition carry 6-7 of the 25 tiles of NL = 36), element-wise
// passes run on all 512 threads, block barriers in between.  Same polynomial, scaling rule and operator algebra as the
// warp-per-entry builder above, whose matrices would leave room for a single warp per SM at these sizes.
constexpr int BM_BUFS = 7;       // A, A2, A3, A4, P0, P1, V
constexpr int BM_WARPS = 16;
constexpr int BM_THREADS = 32 * BM_WARPS;
constexpr int BM_RED = 64 + BM_THREADS;   // norm reduction workspace (doubles): column sums + partial sums
constexpr int BM_EPT = 4;        // elements per thread in element-wise passes: ceil(40 * 40 / 512)

// `ksn` = ceil(n / 4) DMMA k-steps: the columns of A / rows of B beyond n are zero padding
template <int NP, class Init>
__device__ __forceinline__ void bmm(WMat<NP> C, WMat<NP> A, WMat<NP> B, int tid, int ksn, Init init) {
    constexpr int LD = WMat<NP>::LD, NPAD = WMat<NP>::NPAD, NT = NP * NP, TPW = (NT + BM_WARPS - 1) / BM_WARPS;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    double cr[TPW][2], ci[TPW][2];
    int ao[TPW], bo[TPW];
    bool on[TPW];
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
        const int t = warp + BM_WARPS * q;
        on[q] = t < NT;                                   // warp-uniform
        const int mi = on[q] ? t / NP : 0, ni = on[q] ? t - (t / NP) * NP : 0;
        ao[q] = (8 * mi + g) * LD + tq;                   // A fragment of k-step 0
        bo[q] = tq * LD + 8 * ni + g;                     // B fragment of k-step 0
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            cr[q][h] = ci[q][h] = 0.0;
            if (on[q]) init(8 * mi + g, 8 * ni + 2 * tq + h, cr[q][h], ci[q][h]);
        }
    }
#pragma unroll 2
    for (int ks = 0; ks < ksn; ++ks) {
        double ar[TPW], ai[TPW], br[TPW], bi[TPW];
#pragma unroll
        for (int q = 0; q < TPW; ++q) {
            ar[q] = ai[q] = br[q] = bi[q] = 0.0;
            if (on[q]) {
                ar[q] = A.re[ao[q] + 4 * ks];
                ai[q] = A.im()[ao[q] + 4 * ks];
                br[q] = B.re[bo[q] + 4 * ks * LD];
                bi[q] = B.im()[bo[q] + 4 * ks * LD];
            }
        }
#pragma unroll
        for (int q = 0; q < TPW; ++q)
            if (on[q]) {
                dmma8(cr[q][0], cr[q][1], ar[q], br[q]);
                dmma8(ci[q][0], ci[q][1], ar[q], bi[q]);
            }
#pragma unroll
        for (int q = 0; q < TPW; ++q)
            if (on[q]) {
                dmma8(cr[q][0], cr[q][1], -ai[q], bi[q]);
                dmma8(ci[q][0], ci[q][1], ai[q], br[q]);
            }
    }
#pragma unroll
    for (int q = 0; q < TPW; ++q)
        if (on[q]) {
            const int t = warp + BM_WARPS * q;
            const int o = (8 * (t / NP) + g) * LD + 8 * (t - (t / NP) * NP) + 2 * tq;
            *reinterpret_cast<double2*>(C.re + o) = make_double2(cr[q][0], cr[q][1]);
            *reinterpret_cast<double2*>(C.im() + o) = make_double2(ci[q][0], ci[q][1]);
        }
    __syncthreads();
}

struct BlockOffs {    // plane offsets of the elements e = tid + 512 k < n*n a thread touches in element-wise passes
    int o[BM_EPT];
    bool diag[BM_EPT];
    int cnt;
};

// exp(A) (A destroyed); buffers M[0..5] = A, A2, A3, A4, P0, P1; `red`: BM_RED doubles; returns the result buffer (4 or 5)
template <int NP>
__device__ int expm_block(WMat<NP>* M, double* red, int n, int tid, const BlockOffs& bf) {
    constexpr int LD = WMat<NP>::LD;
    WMat<NP> A = M[0], A2 = M[1], A3 = M[2], A4 = M[3], P0 = M[4], P1 = M[5];
    const int ksn = (n + 3) / 4;
    // 1-norm (largest column sum) on all threads: 8 threads share a column (a single thread per column with its chain of
    // n square roots was 18 % of the kernel: profiles/r07g_*), partial sums in red[64 .. 64 + 512), column sums in red[0 .. n)
    {
        const int j = tid & 63, part = tid >> 6;
        double ps = 0.0;
        if (j < n)
            for (int i = part; i < n; i += BM_THREADS / 64) {
                const double x = A.re[i * LD + j], y = A.im()[i * LD + j];
                ps += fabs(x) + fabs(y);     // a bound on |a| (see expm_warp)
            }
        red[64 + part * 64 + j] = ps;
    }
    __syncthreads();
    if (tid < n) {
        double cs = 0.0;
#pragma unroll
        for (int part = 0; part < BM_THREADS / 64; ++part) cs += red[64 + part * 64 + tid];
        red[tid] = cs;
    }
    __syncthreads();
    double cs = 0.0;
    for (int j = 0; j < n; ++j) cs = fmax(cs, red[j]);
    int s = 0;
    if (cs > THETA) {
        int ex;
        frexp(cs / THETA, &ex);
        s = ex > 60 ? 60 : ex;
    }
    const double sc = ldexp(1.0, -s);
#pragma unroll
    for (int k = 0; k < BM_EPT; ++k)
        if (k < bf.cnt) {
            A.re[bf.o[k]] *= sc;
            A.im()[bf.o[k]] *= sc;
        }
    __syncthreads();
    constexpr double c2 = 1.0 / 2, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720,
                     c7 = 1.0 / 5040, c8 = 1.0 / 40320, c9 = 1.0 / 362880, c10 = 1.0 / 3628800,
                     c11 = 1.0 / 39916800, c12 = 1.0 / 479001600;
    auto zero = [](int, int, double& r, double& i) { r = 0.0; i = 0.0; };
    bmm<NP>(A2, A, A, tid, ksn, zero);
    bmm<NP>(A3, A2, A, tid, ksn, zero);
    bmm<NP>(A4, A2, A2, tid, ksn, zero);
#pragma unroll
    for (int k = 0; k < BM_EPT; ++k)
        if (k < bf.cnt) {
            const int o = bf.o[k];
            P0.re[o] = (bf.diag[k] ? c8 : 0.0) + c9 * A.re[o] + c10 * A2.re[o] + c11 * A3.re[o] + c12 * A4.re[o];
            P0.im()[o] = c9 * A.im()[o] + c10 * A2.im()[o] + c11 * A3.im()[o] + c12 * A4.im()[o];
        }
    __syncthreads();
    // P1 = (c4 I + c5 A + c6 A2 + c7 A3) + A4 P0 ;  P0 = (I + A + c2 A2 + c3 A3) + A4 P1
    bmm<NP>(P1, A4, P0, tid, ksn, [&](int i, int j, double& r, double& im_) {
        const int o = i * LD + j;
        r = ((i == j && i < n) ? c4 : 0.0) + c5 * A.re[o] + c6 * A2.re[o] + c7 * A3.re[o];
        im_ = c5 * A.im()[o] + c6 * A2.im()[o] + c7 * A3.im()[o];
    });
    bmm<NP>(P0, A4, P1, tid, ksn, [&](int i, int j, double& r, double& im_) {
        const int o = i * LD + j;
        r = ((i == j && i < n) ? 1.0 : 0.0) + A.re[o] + c2 * A2.re[o] + c3 * A3.re[o];
        im_ = A.im()[o] + c2 * A2.im()[o] + c3 * A3.im()[o];
    });
    int cur = 4, nxt = 5;
    for (int q = 0; q < s; ++q) {
        bmm<NP>(M[nxt], M[cur], M[cur], tid, ksn, zero);
        const int t = cur;
        cur = nxt;
        nxt = t;
    }
    return cur;
}

// A = delta * (L0 + sum_k f_k LA_k + conj(f_k) LB_k) at time t; `fsm`: 2 * n_fields doubles of shared memory
template <int NP>
__device__ void assemble_block(WMat<NP> A, const OpBuildParams& p, const double2* L0, const double2* LA, const double2* LB,
                               double* fsm, int set, double t, double delta, int tid, int ns) {
    constexpr int LD = WMat<NP>::LD;
    const int n = p.prob.NL, n2 = n * n, nf = p.prob.n_fields;
    if (tid < nf) {       // thread k samples drive field k once for the whole matrix
        const double2* tabs = reinterpret_cast<const double2*>(p.tables);
        const double x = (t - p.tab_t0) / p.tab_dt;
        double2 fl = make_double2(0.0, 0.0);
        const int tb = p.prob.field_table[tid];
        if (tb >= 0 && tb < p.n_tables) fl = sample_table(tabs + ((size_t)set * p.n_tables + tb) * p.n_samples, ns, x);
        fsm[2 * tid] = fl.x;
        fsm[2 * tid + 1] = fl.y;
    }
    __syncthreads();
    for (int e = tid; e < n2; e += BM_THREADS) {
        double2 acc = L0[e];
        for (int k = 0; k < nf; ++k) {
            const double2 f = make_double2(fsm[2 * k], fsm[2 * k + 1]);
            cfma(acc, f, LA[(size_t)k * n2 + e]);
            cfma(acc, make_double2(f.x, -f.y), LB[(size_t)k * n2 + e]);
        }
        const int i = e / n;
        const int o = i * LD + (e - i * n);
        A.re[o] = acc.x * delta;
        A.im()[o] = acc.y * delta;
    }
    __syncthreads();
}

constexpr int BM_MAX_FIELDS = 16;

template <int NP>
__global__ void __launch_bounds__(BM_THREADS) k_opbuild_dmma_cta(OpBuildParams p, int lsm) {
    extern __shared__ __align__(16) double wm_smem[];
    constexpr int LD = WMat<NP>::LD, PLANE = WMat<NP>::PLANE;
    const int n = p.prob.NL, n2 = n * n, ksn = (n + 3) / 4;
    const int tid = threadIdx.x;
    WMat<NP> M[BM_BUFS];
#pragma unroll
    for (int b = 0; b < BM_BUFS; ++b) M[b].re = wm_smem + (size_t)b * 2 * PLANE;
    double* red = wm_smem + (size_t)BM_BUFS * 2 * PLANE;      // [BM_RED] norm reduction workspace
    double* fsm = red + BM_RED;                                // [2 * BM_MAX_FIELDS] drive fields of the current matrix
    for (int e = tid; e < BM_BUFS * 2 * PLANE; e += BM_THREADS) wm_smem[e] = 0.0;    // the zero padding stays zero throughout
    // Liouvillian pieces: a copy in shared memory when it fits (NL <= 32; they are read twice per entry)
    const double2* L0 = reinterpret_cast<const double2*>(p.prob.L0);
    const double2* LA = reinterpret_cast<const double2*>(p.prob.LA);
    const double2* LB = reinterpret_cast<const double2*>(p.prob.LB);
    if (lsm) {
        double2* ls = reinterpret_cast<double2*>(fsm + 2 * BM_MAX_FIELDS);
        const int nf = p.prob.n_fields;
        for (int e = tid; e < n2; e += BM_THREADS) ls[e] = L0[e];
        for (int e = tid; e < nf * n2; e += BM_THREADS) {
            ls[n2 + e] = LA[e];
            ls[(1 + nf) * n2 + e] = LB[e];
        }
        L0 = ls;
        LA = ls + n2;
        LB = ls + (size_t)(1 + nf) * n2;
    }
    WMat<NP> V = M[6];
    BlockOffs bf;
    bf.cnt = 0;
#pragma unroll
    for (int k = 0; k < BM_EPT; ++k) {
        const int e = tid + BM_THREADS * k;
        const int i = e / n, j = e - i * n;
        bf.o[k] = e < n2 ? i * LD + j : 0;
        bf.diag[k] = e < n2 && i == j;
        if (e < n2) bf.cnt = k + 1;
    }
    __syncthreads();
    auto load_global = [&](WMat<NP> D, const double2* src) {    // row-major complex n x n -> planes
        for (int e = tid; e < n2; e += BM_THREADS) {
            const int o = (e / n) * LD + e % n;
            D.re[o] = src[e].x;
            D.im()[o] = src[e].y;
        }
        __syncthreads();
    };
    auto copy = [&](WMat<NP> D, WMat<NP> S) {
#pragma unroll
        for (int k = 0; k < BM_EPT; ++k)
            if (k < bf.cnt) {
                D.re[bf.o[k]] = S.re[bf.o[k]];
                D.im()[bf.o[k]] = S.im()[bf.o[k]];
            }
        __syncthreads();
    };
    auto zero = [](int, int, double& r, double& i) { r = 0.0; i = 0.0; };
    const long long total = p.e_end > p.e_begin ? p.e_end : p.n_seq_entries + p.n_entries;
    const double half = 0.5 * p.dt;
    const double2* mto = reinterpret_cast<const double2*>(p.mto_mats);
    for (long long e = p.e_begin + blockIdx.x; e < total; e += gridDim.x) {
        int set, step, sb = -1, sa = -1, has_prev, ns = p.n_samples;
        if (e < p.n_seq_entries) {
            int lo = 0, hi = p.n_seq;  // seq_base[lo] <= e < seq_base[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (p.seq_base[mid] <= e) lo = mid; else hi = mid;
            }
            const aceqd_seq sq = p.seqs[lo];
            const int i = (int)(e - p.seq_base[lo]);
            set = sq.set;
            step = sq.step0 + i;
            has_prev = (i > 0) || sq.first_has_prev;
        } else {
            const aceqd_entry en = p.entries[e - p.n_seq_entries];
            set = en.set; step = en.step; sb = en.sb; sa = en.sa; has_prev = en.has_prev;
            if (en.clamp > 0) ns = min(en.clamp, p.n_samples);   // this row sees only the first `clamp` samples of its drive
        }
        const double t_n = p.t0 + (double)step * p.dt;
        // ---- V = Sb * M2_{n-1}
        if (has_prev) {
            assemble_block<NP>(M[0], p, L0, LA, LB, fsm, set, t_n - p.dt + p.eval_off2 * p.dt, half, tid, ns);
            const int r = expm_block<NP>(M, red, n, tid, bf);
            if (sb >= 0) {
                load_global(M[0], mto + (size_t)sb * n2);
                bmm<NP>(V, M[0], M[r], tid, ksn, zero);
            } else {
                copy(V, M[r]);
            }
        } else {
            for (int q = tid; q < n2; q += BM_THREADS) {
                const int i = q / n, j = q - i * n, o = i * LD + j;
                const double2 v = (sb >= 0) ? mto[(size_t)sb * n2 + q] : make_double2(i == j ? 1.0 : 0.0, 0.0);
                V.re[o] = v.x;
                V.im()[o] = v.y;
            }
            __syncthreads();
        }
        // ---- OV = out_w * V
        {
            const double2* ow = reinterpret_cast<const double2*>(p.prob.out_w);
            double2* ov = reinterpret_cast<double2*>(p.OV + (size_t)e * p.prob.ov_doubles);
            const int cnt = p.prob.n_out * n;
            for (int q = tid; q < cnt; q += BM_THREADS) {
                const int j = q / n, a = q - j * n;
                double2 acc = make_double2(0.0, 0.0);
                for (int k = 0; k < n; ++k) cfma(acc, ow[j * n + k], make_double2(V.re[k * LD + a], V.im()[k * LD + a]));
                ov[q] = acc;
            }
        }
        // ---- X = Sa * V (in V's place, through a work buffer: V must outlive the second exponential)
        if (sa >= 0) {
            __syncthreads();     // OV has read V
            load_global(M[0], mto + (size_t)sa * n2);
            bmm<NP>(M[1], M[0], V, tid, ksn, zero);
            copy(V, M[1]);
        }
        // ---- W = M1_n * X   (zero padded to [NLp8][NLp4])
        assemble_block<NP>(M[0], p, L0, LA, LB, fsm, set, t_n + p.eval_off1 * p.dt, half, tid, ns);
        const int r1 = expm_block<NP>(M, red, n, tid, bf);
        WMat<NP> Wm = M[r1 == 4 ? 5 : 4];
        bmm<NP>(Wm, M[r1], V, tid, ksn, zero);
        {
            double2* w = reinterpret_cast<double2*>(p.W + (size_t)e * p.prob.w_doubles);
            const int ld = p.prob.NLp4, cnt = p.prob.NLp8 * ld;
            for (int q = tid; q < cnt; q += BM_THREADS) {
                const int i = q / ld, j = q - i * ld;
                w[q] = (i < n && j < n) ? make_double2(Wm.re[i * LD + j], Wm.im()[i * LD + j]) : make_double2(0.0, 0.0);
            }
        }
        __syncthreads();  // buffers are reused by the next entry
    }
}

template <int G, bool BLK>
__global__ void __launch_bounds__(256) k_expm_batch(int n, int count, const double* a,
                                                    double* out, double* scratch) {
    extern __shared__ double2 sm[];
    const int n2 = n * n;
    const int groups = blockDim.x / G;
    const int gid = threadIdx.x / G, tid = threadIdx.x - gid * G;
    const size_t per_group = (size_t)EXPM_BUFS * n2 + MAX_NL;
    double2* base = scratch ? reinterpret_cast<double2*>(scratch) + (size_t)blockIdx.x * groups * per_group : sm;
    double2* A = base + gid * per_group;
    double* red = reinterpret_cast<double*>(A + (size_t)EXPM_BUFS * n2);
    for (int e = blockIdx.x * groups + gid; e < count; e += gridDim.x * groups) {
        const double2* src = reinterpret_cast<const double2*>(a) + (size_t)e * n2;
        for (int q = tid; q < n2; q += G) A[q] = src[q];
        group_sync<G>(gid);
        double2* R = expm_group<G, BLK>(A, red, n, tid, gid);
        double2* dst = reinterpret_cast<double2*>(out) + (size_t)e * n2;
        for (int q = tid; q < n2; q += G) dst[q] = R[q];
        group_sync<G>(gid);
    }
}

int pick_group(int n) {   // threads per entry (measured optima: gpurun_out/r03z_call.log)
    if (const char* env = getenv("ACEQD_EXPM_GROUP")) {
        const int g = atoi(env);
        if (g == 16 || g == 32 || g == 64 || g == 128 || g == 256) return g;
    }
    if (n <= 4) return 16;
    if (n <= 8) return 32;
    if (n <= 12) return 64;
    if (n <= 20) return 128;
    if (n <= 28) return 256;
    return 128;   // NL = 36: two groups per CTA on the global workspace (2 CTAs per SM) beat one group in shared memory
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        ACEQD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)bytes));
    return ACEQD_OK;
}

}  // namespace

constexpr int SCRATCH_CTAS = 296;   // two CTAs per SM when the work matrices live in global memory

size_t opbuild_scratch_bytes(int NL, int* ctas) {
    const int groups = 256 / pick_group(NL);
    const size_t per_cta = groups * ((size_t)(EXPM_BUFS + 2) * NL * NL + MAX_NL) * sizeof(double2);
    if (ctas) *ctas = SCRATCH_CTAS;
    return per_cta > (size_t)SMEM_BUDGET ? per_cta * SCRATCH_CTAS : 0;
}

int launch_opbuild(const OpBuildParams& p, cudaStream_t s, LaunchLog* log) {
    const long long total = p.e_end > p.e_begin ? p.e_end - p.e_begin : p.n_seq_entries + p.n_entries;
    if (total <= 0) return ACEQD_OK;
    const int n = p.prob.NL;
    if (n > MAX_NL) {
        set_error("NL=%d exceeds MAX_NL=%d", n, MAX_NL);
        return ACEQD_ERR_CAPACITY;
    }
    if (n == 4 && p.prob.NLp8 == 8 && p.prob.NLp4 == 4 && !getenv("ACEQD_OPBUILD_GROUP")) {   // register-resident builder, one thread per entry
        long long blocks = (total + OPREG_THREADS - 1) / OPREG_THREADS;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        k_opbuild_reg<4><<<(int)blocks, OPREG_THREADS, 0, s>>>(p);
        ++log->count;
        log_name(log->opbuild, "k_opbuild_reg<4>");
        ACEQD_CUDA(cudaGetLastError());
        return ACEQD_OK;
    }
    if (n > 4 && n <= 16 && !getenv("ACEQD_OPBUILD_GROUP")) {   // tensor-core builder, one warp per entry
        const int np = (n + 7) / 8;
        const size_t per_warp = (size_t)WM_BUFS * 2 * (8 * np) * (8 * np + 4) * sizeof(double);
        // four warps per CTA = one per SM sub-partition.  More entries in flight do NOT help: 5 / 6 warps per SM (seven
        // instead of eight matrices per warp make room) measured 9.65 / 9.80 ms against 9.39 ms for the biexciton sweep
        // (gpurun_out/r7i): the kernel is bound by the shared-memory pipe (fragment loads, 2-way conflicts of the C
        // stores at LD = 20), not by the latency of one warp's chain.
        int wpc = (int)std::max<size_t>(1, std::min<size_t>(WM_MAX_WARPS, (size_t)SMEM_BUDGET / per_warp));
        if (const char* ev = getenv("ACEQD_OPB_WARPS")) wpc = std::max(1, std::min(wpc, atoi(ev)));   // tuning
        const size_t l_bytes = (size_t)(1 + 2 * p.prob.n_fields) * n * n * sizeof(double2);
        const int lsm = (size_t)wpc * per_warp + l_bytes <= (size_t)SMEM_BUDGET ? 1 : 0;
        const size_t smem = (size_t)wpc * per_warp + (lsm ? l_bytes : 0);
        const long long per_sm = std::max<long long>(1, (long long)((size_t)SMEM_BUDGET / smem));
        long long blocks = (total + wpc - 1) / wpc;
        if (blocks > 148LL * per_sm) blocks = 148LL * per_sm;
        int rc;
        if (np == 1) {
            if ((rc = set_smem(k_opbuild_dmma<1>, smem))) return rc;
            k_opbuild_dmma<1><<<(int)blocks, 32 * wpc, smem, s>>>(p, wpc, lsm);
        } else {
            if ((rc = set_smem(k_opbuild_dmma<2>, smem))) return rc;
            k_opbuild_dmma<2><<<(int)blocks, 32 * wpc, smem, s>>>(p, wpc, lsm);
        }
        ++log->count;
        log_name(log->opbuild, "k_opbuild_dmma<%d>", np);
        ACEQD_CUDA(cudaGetLastError());
        return ACEQD_OK;
    }
    if (n > 16 && n <= 40 && p.prob.n_fields <= BM_MAX_FIELDS && !getenv("ACEQD_OPBUILD_GROUP")) {
        // tensor-core builder, one CTA per entry (five- and six-level models)
        const int np = (n + 7) / 8;
        const size_t m_bytes = ((size_t)BM_BUFS * 2 * (8 * np) * (8 * np + 4) + BM_RED + 2 * BM_MAX_FIELDS) * sizeof(double);
        const size_t l_bytes = (size_t)(1 + 2 * p.prob.n_fields) * n * n * sizeof(double2);
        const int lsm = m_bytes + l_bytes <= (size_t)SMEM_BUDGET ? 1 : 0;
        const size_t smem = m_bytes + (lsm ? l_bytes : 0);
        // 512 threads of <= 128 registers: the register file holds one CTA per SM
        const long long blocks = std::min<long long>(total, 148LL);
        int rc;
        if (np == 3) {
            if ((rc = set_smem(k_opbuild_dmma_cta<3>, smem))) return rc;
            k_opbuild_dmma_cta<3><<<(int)blocks, BM_THREADS, smem, s>>>(p, lsm);
        } else if (np == 4) {
            if ((rc = set_smem(k_opbuild_dmma_cta<4>, smem))) return rc;
            k_opbuild_dmma_cta<4><<<(int)blocks, BM_THREADS, smem, s>>>(p, lsm);
        } else {
            if ((rc = set_smem(k_opbuild_dmma_cta<5>, smem))) return rc;
            k_opbuild_dmma_cta<5><<<(int)blocks, BM_THREADS, smem, s>>>(p, lsm);
        }
        ++log->count;
        log_name(log->opbuild, "k_opbuild_dmma_cta<%d>", np);
        ACEQD_CUDA(cudaGetLastError());
        return ACEQD_OK;
    }
    const int G = pick_group(n);
    const int groups = 256 / G;
    size_t smem = groups * ((size_t)(EXPM_BUFS + 2) * n * n + MAX_NL) * sizeof(double2);
    long long blocks = (total + groups - 1) / groups;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    if (smem > (size_t)SMEM_BUDGET) {
        if (!p.scratch || p.scratch_ctas <= 0) {
            set_error("operator builder: NL=%d needs %zu B of work matrices and no global workspace was given", n, smem);
            return ACEQD_ERR_CAPACITY;
        }
        smem = 0;
        blocks = std::min<long long>(blocks, p.scratch_ctas);
    }
    int rc = ACEQD_OK;
    const bool blk = n > 16;   // 2 x 2 register blocks per thread for the larger matrices only (measured)
#define ACEQD_OPB(GV)                                                                  \
    do {                                                                               \
        if (blk) {                                                                     \
            if ((rc = set_smem(k_opbuild<GV, true>, smem))) return rc;                 \
            k_opbuild<GV, true><<<(int)blocks, 256, smem, s>>>(p);                     \
        } else {                                                                       \
            if ((rc = set_smem(k_opbuild<GV, false>, smem))) return rc;                \
            k_opbuild<GV, false><<<(int)blocks, 256, smem, s>>>(p);                    \
        }                                                                              \
    } while (0)
    switch (G) {
        case 16: ACEQD_OPB(16); break;
        case 32: ACEQD_OPB(32); break;
        case 64: ACEQD_OPB(64); break;
        case 128: ACEQD_OPB(128); break;
        default: ACEQD_OPB(256); break;
    }
#undef ACEQD_OPB
    ++log->count;
    log_name(log->opbuild, "k_opbuild<%d,%s>", G, blk ? "true" : "false");
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

int launch_expm_batch(int n, int count, const double* a_dev, double* out_dev, double* scratch,
                      cudaStream_t s, LaunchLog* log) {
    if (count <= 0) return ACEQD_OK;
    if (n > MAX_NL) {
        set_error("n=%d exceeds MAX_NL=%d", n, MAX_NL);
        return ACEQD_ERR_CAPACITY;
    }
    const int G = pick_group(n);
    const int groups = 256 / G;
    size_t smem = groups * ((size_t)EXPM_BUFS * n * n + MAX_NL) * sizeof(double2);
    int blocks = (count + groups - 1) / groups;
    if (blocks > 148 * 64) blocks = 148 * 64;
    if (smem > (size_t)SMEM_BUDGET) {
        if (!scratch) {
            set_error("expm: n=%d needs a global workspace", n);
            return ACEQD_ERR_CAPACITY;
        }
        smem = 0;
        blocks = std::min(blocks, SCRATCH_CTAS);
    } else {
        scratch = nullptr;
    }
    int rc = ACEQD_OK;
    const bool blk = n > 16;
#define ACEQD_EXB(GV)                                                                              \
    do {                                                                                           \
        if (blk) {                                                                                 \
            if ((rc = set_smem(k_expm_batch<GV, true>, smem))) return rc;                          \
            k_expm_batch<GV, true><<<blocks, 256, smem, s>>>(n, count, a_dev, out_dev, scratch);   \
        } else {                                                                                   \
            if ((rc = set_smem(k_expm_batch<GV, false>, smem))) return rc;                         \
            k_expm_batch<GV, false><<<blocks, 256, smem, s>>>(n, count, a_dev, out_dev, scratch);  \
        }                                                                                          \
    } while (0)
    switch (G) {
        case 16: ACEQD_EXB(16); break;
        case 32: ACEQD_EXB(32); break;
        case 64: ACEQD_EXB(64); break;
        case 128: ACEQD_EXB(128); break;
        default: ACEQD_EXB(256); break;
    }
#undef ACEQD_EXB
    ++log->count;
    log_name(log->other, "k_expm_batch<%d,%s>", G, blk ? "true" : "false");
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

}  // namespace aceqd
