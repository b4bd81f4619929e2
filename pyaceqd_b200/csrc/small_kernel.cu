// Step kernel for the SMALL-BOND regime of two-level sweeps (NL = 4, chi_pad <= 32: SURVEY 8d "bandwidth-bound
// small-chi regime").  The persistent tile kernel (step_kernel.cu) spreads the bond columns of a tile over 8 warps
// and pays block barriers per step; with 2..4 n-tiles most of its warps idle and the per-step phases dominate.
// Here ONE WARP owns 8 trajectories (an "octet") for all their steps and nothing is shared between warps but the
// read-only process tensor, which is small enough to live in shared memory as a whole:
//   A  outputs      out[j] = OV_n[j] . r[j]      r = closure of the previous PT slice, kept in registers
//   B  system       X[a',j,c] = sum_a W_n[j][a',a] Y[a,j,c]      FP64 FMA; lane (j, c mod 4); in place
//   C  PT slice     Y[a,j,:] = X[a,j,:] A_n[beta(a)]             DMMA.8x8x4, m-tile = the 8 trajectories of one
//                   Liouville row a, B fragments from the resident PT; closure partials by quad shuffles
// The DMMA A fragment of lane (j, tq) holds columns 4ks+tq, exactly the columns the same lane wrote in phase B,
// so the only synchronisation of a step is one __syncwarp() after the C-fragment stores.  The per-row operators
// (512 B + outputs per trajectory-row: the HBM stream of this regime) are prefetched one row ahead with cp.async.
//
// Same inputs, same outputs and same arithmetic order per element as k_step_dmma's NL <= 4 path up to the
// summation order of the closure; replaces the inner loop of ACE's Simulation.run for
// pyaceqd/two_level_system/rabi_rotations.py:172-198 style sweeps at small bond dimension.
#include "kernel_common.cuh"

namespace aceqd {

namespace {

constexpr int OCT = 8;           // trajectories per warp
constexpr int SMALL_NL = 4;
constexpr int W_USED = 16;       // complex entries of W actually read (rows 0..3 of the [8][4] padded block)

struct SmallLayout {
    size_t pt, clo, meta, warp0, per_warp, state_plane, wov_buf, total;
};

__host__ __device__ inline SmallLayout small_layout(long long pt_doubles, int n_slices, int chi_pad, int n_out,
                                                    int warps) {
    SmallLayout L;
    const size_t strideA = chi_pad + 4;
    size_t o = 0;
    L.pt = o;    o += (size_t)pt_doubles * 8;
    o = (o + 15) / 16 * 16;
    L.clo = o;   o += (size_t)n_slices * chi_pad * 16;
    L.meta = o;  o += ((size_t)n_slices * 16 + 15) / 16 * 16;     // kin, nout (int), off (long long)
    o = (o + 127) / 128 * 128;
    L.warp0 = o;
    L.state_plane = (size_t)SMALL_NL * OCT * strideA;             // doubles per plane
    L.wov_buf = (size_t)OCT * (W_USED + (size_t)n_out * SMALL_NL) * 2;   // doubles per staging buffer
    size_t w = 2 * L.state_plane * 8 + 2 * L.wov_buf * 8 + OCT * sizeof(aceqd_traj);
    L.per_warp = (w + 127) / 128 * 128;
    L.total = o + (size_t)warps * L.per_warp;
    return L;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(256) k_step_small(const __grid_constant__ StepParams p, long long pt_doubles,
                                                   int warps_per_cta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the bond dimension is a template parameter: every state / PT offset below folds into an immediate (a single
    // warp per sub-partition cannot hide address arithmetic behind other warps)
    constexpr int chi_pad = 8 * NT, strideA = chi_pad + 4, strideB = chi_pad + 4, CHUNK = 2 * KC * strideB;
    const int n_out = p.prob.n_out, n_slices = p.pt.n_slices;
    const SmallLayout L = small_layout(pt_doubles, n_slices, chi_pad, n_out, warps_per_cta);
    double* pt_s = reinterpret_cast<double*>(smem_raw + L.pt);
    double2* clo_s = reinterpret_cast<double2*>(smem_raw + L.clo);
    int* kin_s = reinterpret_cast<int*>(smem_raw + L.meta);
    int* nout_s = kin_s + n_slices;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;      // g = trajectory of the octet, tq = column class / DMMA k index

    // ---- the whole process tensor, closures and slice metadata: once per CTA
    {
        const double2* src = reinterpret_cast<const double2*>(p.pt.blob);
        double2* dst = reinterpret_cast<double2*>(pt_s);
        for (long long e = tid; e < pt_doubles / 2; e += blockDim.x) dst[e] = src[e];
        const double2* cs = reinterpret_cast<const double2*>(p.pt.closure);
        for (int e = tid; e < n_slices * chi_pad; e += blockDim.x) clo_s[e] = cs[e];
        for (int e = tid; e < n_slices; e += blockDim.x) {
            kin_s[e] = p.pt.kin_pad[e];
            nout_s[e] = p.pt.nout_pad[e];
        }
    }
    __syncthreads();
    const int oct = blockIdx.x * warps_per_cta + warp;
    if (oct >= p.n_tiles) return;     // no block-level synchronisation below

    unsigned char* wbase = smem_raw + L.warp0 + (size_t)warp * L.per_warp;
    double* Xre = reinterpret_cast<double*>(wbase);
    double* Xim = Xre + L.state_plane;
    double* stg = Xim + L.state_plane;                       // [2][OCT][W_USED + n_out*4] complex
    aceqd_traj* trj = reinterpret_cast<aceqd_traj*>(stg + 2 * L.wov_buf);
    const int per_traj = 2 * (W_USED + n_out * SMALL_NL);    // doubles of one trajectory's staged operators

    for (int j = lane; j < OCT; j += 32) {
        const int idx = p.tile_traj[(size_t)oct * OCT + j];
        if (idx >= 0) {
            trj[j] = p.trajs[idx];
        } else {
            aceqd_traj z;
            memset(&z, 0, sizeof(z));
            z.n_steps = -1;
            trj[j] = z;
        }
    }
    for (size_t e = lane; e < 2 * L.state_plane; e += 32) Xre[e] = 0.0;
    __syncwarp();
    const aceqd_traj t = trj[g];            // this lane's trajectory
    const bool valid = t.n_steps >= 0;
    int n_begin = valid ? t.step0 : 0x7fffffff, n_end = valid ? t.step0 + t.n_steps : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_begin = min(n_begin, __shfl_xor_sync(0xffffffffu, n_begin, o));
        n_end = max(n_end, __shfl_xor_sync(0xffffffffu, n_end, o));
    }
    if (n_end < 0) return;
    // state offset of (row a, trajectory j, column c) inside a plane
    auto soff = [&](int a, int j, int c) -> int { return (a * OCT + j) * strideA + c; };
    // initial states
    if (valid) {
        if (t.init_kind == 0) {
            const double2* r0 = reinterpret_cast<const double2*>(p.rho0s) + (size_t)t.init_index * SMALL_NL;
            if (tq == 0)
                for (int a = 0; a < SMALL_NL; ++a) {
                    Xre[soff(a, g, 0)] = r0[a].x;
                    Xim[soff(a, g, 0)] = r0[a].y;
                }
        } else {
            const double2* sn = reinterpret_cast<const double2*>(p.snaps) + (size_t)t.init_index * SMALL_NL * chi_pad;
            for (int a = 0; a < SMALL_NL; ++a)
                for (int c = tq; c < chi_pad; c += 4) {
                    const double2 v = sn[a * chi_pad + c];
                    Xre[soff(a, g, c)] = v.x;
                    Xim[soff(a, g, c)] = v.y;
                }
        }
    }
    __syncwarp();

    // stage W_n | OV_n of the 8 trajectories for output row n (lane (g, tq) copies row tq of W and its share of OV)
    auto stage_row = [&](int n) {
        if (valid && n >= t.step0 && n <= t.step0 + t.n_steps) {
            const long long e = entry_of(trj[g], n - t.step0, p.ovr_base);
            double* dst = stg + (size_t)(n & 1) * L.wov_buf + (size_t)g * per_traj;
            const double2* w = reinterpret_cast<const double2*>(p.W + (size_t)e * p.prob.w_doubles);
            double2* dw = reinterpret_cast<double2*>(dst);
#pragma unroll
            for (int k = 0; k < 4; ++k) cp_async16(dw + tq * 4 + k, w + tq * p.prob.NLp4 + k);
            const double2* ov = reinterpret_cast<const double2*>(p.OV + (size_t)e * p.prob.ov_doubles);
            for (int q = tq; q < n_out * SMALL_NL; q += 4) cp_async16(dw + W_USED + q, ov + q);
        }
        cp_async_commit();
    };
    stage_row(n_begin);

    int blk_of[SMALL_NL];
#pragma unroll
    for (int a = 0; a < SMALL_NL; ++a) blk_of[a] = p.prob.block_of_alpha[a];
    double2 r[SMALL_NL];                 // closure rho[a] of this lane's trajectory at the current row
#pragma unroll
    for (int a = 0; a < SMALL_NL; ++a) r[a] = make_double2(0.0, 0.0);

    for (int n = n_begin; n <= n_end; ++n) {
        if (n < n_end) {
            stage_row(n + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const int i = n - t.step0;                                   // local row of this lane's trajectory
        const bool row_on = valid && i >= 0 && i <= t.n_steps;       // writes output row n
        const bool step_on = valid && i >= 0 && i < t.n_steps;       // steps n -> n+1
        const double2* sw = reinterpret_cast<const double2*>(stg + (size_t)(n & 1) * L.wov_buf + (size_t)g * per_traj);
        // ---------------- A: closure of a trajectory that starts at this row, then the outputs
        if (valid && i == 0) {
            if (t.init_kind == 0) {
#pragma unroll
                for (int a = 0; a < SMALL_NL; ++a) r[a] = make_double2(Xre[soff(a, g, 0)], Xim[soff(a, g, 0)]);
            } else {
                const double2* q = clo_s + (size_t)slice_of(p.pt, n - 1) * chi_pad;
#pragma unroll
                for (int a = 0; a < SMALL_NL; ++a) {
                    double2 acc = make_double2(0.0, 0.0);
                    for (int c = tq; c < chi_pad; c += 4) {
                        const double x = Xre[soff(a, g, c)], y = Xim[soff(a, g, c)];
                        acc.x += x * q[c].x - y * q[c].y;
                        acc.y += x * q[c].y + y * q[c].x;
                    }
                    r[a] = acc;
                }
            }
        }
        {   // snapshot-started rows hold quad-partial closures: complete them (uniform shuffles, selected per lane)
            const bool part = valid && i == 0 && t.init_kind != 0;
            if (__any_sync(0xffffffffu, part)) {
#pragma unroll
                for (int a = 0; a < SMALL_NL; ++a) {
                    double sx = r[a].x, sy = r[a].y;
                    sx += __shfl_xor_sync(0xffffffffu, sx, 1);
                    sy += __shfl_xor_sync(0xffffffffu, sy, 1);
                    sx += __shfl_xor_sync(0xffffffffu, sx, 2);
                    sy += __shfl_xor_sync(0xffffffffu, sy, 2);
                    if (part) r[a] = make_double2(sx, sy);
                }
            }
        }
        if (row_on && i >= t.out_from) {
            double2* out = reinterpret_cast<double2*>(p.out) + t.out_off + (long long)(i - t.out_from) * n_out;
            for (int o = tq; o < n_out; o += 4) {
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int a = 0; a < SMALL_NL; ++a) {
                    const double2 w = sw[W_USED + o * SMALL_NL + a];
                    acc.x += w.x * r[a].x - w.y * r[a].y;
                    acc.y += w.x * r[a].y + w.y * r[a].x;
                }
                out[o] = acc;
            }
        }
        if (n == n_end) break;
        // ---------------- B: X = W_n Y for the columns c = tq (mod 4) of this lane's trajectory, in place
        if (step_on) {
            double2 w[SMALL_NL][SMALL_NL];
#pragma unroll
            for (int a = 0; a < SMALL_NL; ++a)
#pragma unroll
                for (int k = 0; k < SMALL_NL; ++k) w[a][k] = sw[a * 4 + k];
#pragma unroll
            for (int cc = 0; cc < 2 * NT; ++cc) {          // chi_pad = 8 NT columns, every fourth one is this lane's
                const int c = tq + 4 * cc;
                double2 y[SMALL_NL];
#pragma unroll
                for (int k = 0; k < SMALL_NL; ++k) y[k] = make_double2(Xre[soff(k, g, c)], Xim[soff(k, g, c)]);
#pragma unroll
                for (int a = 0; a < SMALL_NL; ++a) {
                    double xr = 0.0, xi = 0.0;
#pragma unroll
                    for (int k = 0; k < SMALL_NL; ++k) {
                        xr = fma(w[a][k].x, y[k].x, xr);
                        xr = fma(-w[a][k].y, y[k].y, xr);
                        xi = fma(w[a][k].x, y[k].y, xi);
                        xi = fma(w[a][k].y, y[k].x, xi);
                    }
                    Xre[soff(a, g, c)] = xr;
                    Xim[soff(a, g, c)] = xi;
                }
            }
        }
        // ---------------- C: PT slice (DMMA: rows = the 8 trajectories, one Liouville row a at a time)
        const int s = slice_of(p.pt, n);
        const int nks = kin_s[s] / 4, nout = nout_s[s], nch = kin_s[s] / KC;
        const double* sl = pt_s + p.pt.off[s];
        const double2* q = clo_s + (size_t)s * chi_pad;
#pragma unroll
        for (int a0 = 0; a0 < SMALL_NL; a0 += 2) {      // two Liouville rows side by side: independent DMMA chains
            double cre[2][NT][2], cim[2][NT][2];
            const double* xre[2];
            const double* xim[2];
            const double* blk[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) cre[h][nt][0] = cre[h][nt][1] = cim[h][nt][0] = cim[h][nt][1] = 0.0;
                xre[h] = Xre + soff(a0 + h, g, tq);
                xim[h] = Xim + soff(a0 + h, g, tq);
                blk[h] = sl + (size_t)blk_of[a0 + h] * nch * CHUNK;
            }
#pragma unroll
            for (int ks = 0; ks < 2 * NT; ++ks) {       // kin_pad <= chi_pad = 8 NT
                if (ks < nks) {                         // warp-uniform
                    double a_re[2], a_im[2], b_re[2][NT], b_im[2][NT];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        a_re[h] = xre[h][4 * ks];
                        a_im[h] = xim[h][4 * ks];
                        const double* bre = blk[h] + (size_t)(ks >> 1) * CHUNK + ((ks & 1) * 4 + tq) * strideB + g;
                        const double* bim = bre + KC * strideB;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            b_re[h][nt] = 8 * nt < nout ? bre[8 * nt] : 0.0;
                            b_im[h][nt] = 8 * nt < nout ? bim[8 * nt] : 0.0;
                        }
                    }
                    // two sweeps so that consecutive DMMAs never share an accumulator (8 * nt < nout is warp-uniform)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            if (8 * nt < nout) {
                                dmma(cre[h][nt][0], cre[h][nt][1], a_re[h], b_re[h][nt]);
                                dmma(cim[h][nt][0], cim[h][nt][1], a_re[h], b_im[h][nt]);
                            }
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            if (8 * nt < nout) {
                                dmma(cre[h][nt][0], cre[h][nt][1], -a_im[h], b_im[h][nt]);
                                dmma(cim[h][nt][0], cim[h][nt][1], a_im[h], b_re[h][nt]);
                            }
                }
            }
            // new rows a0, a0+1 of trajectory g (columns 8nt + 2tq, +1) and their closures
            double pr[2] = {0.0, 0.0}, pi[2] = {0.0, 0.0};
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                if (8 * nt < nout) {
                    const int c0 = 8 * nt + 2 * tq;
                    const double2 q0 = q[c0], q1 = q[c0 + 1];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        pr[h] += (cre[h][nt][0] * q0.x - cim[h][nt][0] * q0.y) + (cre[h][nt][1] * q1.x - cim[h][nt][1] * q1.y);
                        pi[h] += (cre[h][nt][0] * q0.y + cim[h][nt][0] * q0.x) + (cre[h][nt][1] * q1.y + cim[h][nt][1] * q1.x);
                        if (step_on) {
                            *reinterpret_cast<double2*>(Xre + soff(a0 + h, g, c0)) = make_double2(cre[h][nt][0], cre[h][nt][1]);
                            *reinterpret_cast<double2*>(Xim + soff(a0 + h, g, c0)) = make_double2(cim[h][nt][0], cim[h][nt][1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                pr[h] += __shfl_xor_sync(0xffffffffu, pr[h], 1);
                pi[h] += __shfl_xor_sync(0xffffffffu, pi[h], 1);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                pr[h] += __shfl_xor_sync(0xffffffffu, pr[h], 2);
                pi[h] += __shfl_xor_sync(0xffffffffu, pi[h], 2);
                if (step_on) r[a0 + h] = make_double2(pr[h], pi[h]);
            }
        }
        __syncwarp();   // the C-fragment columns of a lane are read by the other lanes of its quad in the next phase B
    }
}

}  // namespace

size_t small_smem_bytes(long long pt_doubles, int n_slices, int chi_pad, int n_out, int warps) {
    return small_layout(pt_doubles, n_slices, chi_pad, n_out, warps).total;
}

// p.tile_traj must hold OCTETS (8 trajectory indices per entry, -1 = none), p.n_tiles their number.
int launch_step_small(const StepParams& p, long long pt_doubles, int warps_per_cta, size_t smem_bytes, cudaStream_t s,
                      LaunchLog* log) {
    if (p.n_tiles <= 0) return ACEQD_OK;
    const int nt = p.pt.chi_pad / 8;
    const int grid = (p.n_tiles + warps_per_cta - 1) / warps_per_cta;
#define ACEQD_SMALL(NTV)                                                                                          \
    do {                                                                                                          \
        ACEQD_CUDA(cudaFuncSetAttribute(k_step_small<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                        (int)smem_bytes));                                                        \
        k_step_small<NTV><<<grid, 32 * warps_per_cta, smem_bytes, s>>>(p, pt_doubles, warps_per_cta);             \
    } while (0)
    switch (nt) {
        case 1: ACEQD_SMALL(1); break;
        case 2: ACEQD_SMALL(2); break;
        case 3: ACEQD_SMALL(3); break;
        case 4: ACEQD_SMALL(4); break;
        default:
            set_error("small-bond kernel: chi_pad=%d not supported", p.pt.chi_pad);
            return ACEQD_ERR_CAPACITY;
    }
#undef ACEQD_SMALL
    ++log->count;
    log_name(log->step, "k_step_small<%d> warps=%d", nt, warps_per_cta);
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

}  // namespace aceqd
