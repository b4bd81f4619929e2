// Step-synchronous streaming variant of the PT step kernel for LARGE Liouville spaces (NL >= 9).
//
// The persistent kernel (step_kernel.cu) keeps the bond state of T trajectories in one CTA's shared
// memory.  At NL = 16..36 and chi = 128 only T = 1..4 trajectories fit, so a GEMM pass multiplies a
// PT block (chi x chi complex, 270 KB) with 4..16 state rows: 4 flop per streamed PT byte, and the
// pass is bound by the L2 -> shared-memory ingest rate of one SM (~30 GB/s, profiles/r01m_*), not by
// the tensor pipe; m-tiles are mostly padding (44..80 % filled).
//
// Here the state lives in HBM/L2 ([trajectory][alpha position][chi] split re/im planes, 32 KB per
// trajectory at NL = 16) and every absolute time step is two grid-wide phases of one cooperative
// persistent kernel:
//   phase 1 (one CTA task per active trajectory): closure + outputs (+ snapshots) and the small
//           system product X = W_n Y;
//   phase 2 (one CTA task per 16 rows of ONE coupling class, rows gathered ACROSS trajectories):
//           Y = X A_n[class] with the same DMMA pass / cp.async.bulk chunk pipeline as the persistent
//           kernel -- full m-tiles regardless of T, and every PT byte feeds 16 rows.
// Trajectories may start and end at different absolute steps (G2(t,tau) branches): they are sorted by
// start step, the active set is a contiguous window that slides with n.  Same inputs, same outputs,
// same per-row operators (k_opbuild) as the persistent kernel.
#include "kernel_common.cuh"

namespace aceqd {

namespace {

constexpr int SM_ROWS = 16;     // rows of one phase-2 task (two DMMA m-tiles)

struct StreamSmem {
    size_t bar, q, r, wov, state, chunks, total;
};

__host__ __device__ inline size_t s_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline StreamSmem stream_layout(int NL, int chi_pad, int stages, int wov_doubles) {
    StreamSmem L;
    const size_t strideA = chi_pad + 4;
    const size_t rows = NL > SM_ROWS ? NL : SM_ROWS;
    size_t o = 0;
    L.bar = o;    o += 128;
    L.q = o;      o += s_align(16 * (size_t)chi_pad, 16);
    L.r = o;      o += s_align(16 * (size_t)MAX_NL, 16);
    L.wov = o;    o += s_align((size_t)wov_doubles * 8, 16);
    o = s_align(o, 128);
    L.state = o;  o += 2 * rows * strideA * 8;
    o = s_align(o, 128);
    L.chunks = o; o += (size_t)stages * 2 * KC * strideA * 8;
    L.total = o;
    return L;
}

// all CTAs are co-resident (cooperative launch): monotone counter, one arrival per CTA per barrier
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned n_cta, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        const unsigned target = epoch * n_cta;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

template <int NB>
__global__ void __launch_bounds__(STEP_THREADS, 1) k_step_stream(const __grid_constant__ StreamParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NL = p.prob.NL, chi_pad = p.pt.chi_pad, n_out = p.prob.n_out;
    const int strideA = chi_pad + 4, strideB = p.pt.strideB, stages = p.stages;
    const int NLp4 = p.prob.NLp4;
    const StreamSmem L = stream_layout(NL, chi_pad, stages, p.prob.w_doubles + p.prob.ov_doubles);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    double2* qbuf = reinterpret_cast<double2*>(smem_raw + L.q);
    double2* rvec = reinterpret_cast<double2*>(smem_raw + L.r);
    double2* Wsm = reinterpret_cast<double2*>(smem_raw + L.wov);
    double2* OVsm = Wsm + p.prob.w_doubles / 2;
    double* Sre = reinterpret_cast<double*>(smem_raw + L.state);
    const size_t plane = (size_t)(NL > SM_ROWS ? NL : SM_ROWS) * strideA;
    double* Sim = Sre + plane;
    double* chunks = reinterpret_cast<double*>(smem_raw + L.chunks);

    __shared__ int pos_s[MAX_NL], apos_s[MAX_NL];   // alpha -> position, position -> alpha
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned G = gridDim.x;
    for (int a = tid; a < NL; a += blockDim.x) {
        pos_s[a] = p.pos_of_alpha[a];
        apos_s[a] = p.alpha_of_pos[a];
    }
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + MAX_STAGES);
    const bool producer = warp == N_COMPUTE_WARPS;
    const size_t row_doubles = (size_t)chi_pad;            // one state row in global memory

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, N_COMPUTE_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    unsigned epoch = 0;
    int stage = 0;
    uint32_t phase = 0;
    int lo = 0, hi = 0;
    const int g = lane >> 2, tq = lane & 3;

    for (int n = p.n_begin; n <= p.n_end; ++n) {
        // ---- active window over the trajectories sorted by start step
        while (hi < p.n_traj && p.trajs[p.order[hi]].step0 <= n) ++hi;
        while (lo < hi && p.trajs[p.order[lo]].step0 + p.trajs[p.order[lo]].n_steps < n) ++lo;

        // ================================================================= phase 1: per trajectory
        if (!producer) {
            for (int k = lo + (int)blockIdx.x; k < hi; k += (int)G) {
                const aceqd_traj t = p.trajs[p.order[k]];
                const int i = n - t.step0;
                if (i < 0 || i > t.n_steps) continue;      // block-uniform
                const long long e = entry_of(t, i, p.ovr_base);
                // operators of this output row
                {
                    const double2* Wg = reinterpret_cast<const double2*>(p.W + (size_t)e * p.prob.w_doubles);
                    const double2* Og = reinterpret_cast<const double2*>(p.OV + (size_t)e * p.prob.ov_doubles);
                    for (int x = tid; x < p.prob.w_doubles / 2; x += N_COMPUTE_WARPS * 32) Wsm[x] = __ldg(Wg + x);
                    for (int x = tid; x < p.prob.ov_doubles / 2; x += N_COMPUTE_WARPS * 32) OVsm[x] = __ldg(Og + x);
                }
                // bond state Y of this trajectory -> shared memory (rows in alpha-position order)
                if (i == 0) {
                    if (t.init_kind == 0) {
                        const double2* r0 = reinterpret_cast<const double2*>(p.rho0s) + (size_t)t.init_index * NL;
                        for (int x = tid; x < NL * chi_pad; x += N_COMPUTE_WARPS * 32) {
                            const int ps = x / chi_pad, d = x - ps * chi_pad;
                            const double2 v = d == 0 ? r0[apos_s[ps]] : make_double2(0.0, 0.0);
                            Sre[(size_t)ps * strideA + d] = v.x;
                            Sim[(size_t)ps * strideA + d] = v.y;
                        }
                    } else {
                        const double2* sn = reinterpret_cast<const double2*>(p.snaps) + (size_t)t.init_index * NL * chi_pad;
                        for (int x = tid; x < NL * chi_pad; x += N_COMPUTE_WARPS * 32) {
                            const int ps = x / chi_pad, d = x - ps * chi_pad;
                            const double2 v = __ldcg(sn + (size_t)apos_s[ps] * chi_pad + d);
                            Sre[(size_t)ps * strideA + d] = v.x;
                            Sim[(size_t)ps * strideA + d] = v.y;
                        }
                    }
                } else {
                    const double* yr = p.Yre + (size_t)k * NL * row_doubles;
                    const double* yi = p.Yim + (size_t)k * NL * row_doubles;
                    for (int x = tid; x < NL * chi_pad / 2; x += N_COMPUTE_WARPS * 32) {
                        const int ps = (2 * x) / chi_pad, d = 2 * x - ps * chi_pad;
                        const double2 a = __ldcg(reinterpret_cast<const double2*>(yr + (size_t)ps * row_doubles + d));
                        const double2 b = __ldcg(reinterpret_cast<const double2*>(yi + (size_t)ps * row_doubles + d));
                        *reinterpret_cast<double2*>(Sre + (size_t)ps * strideA + d) = a;
                        *reinterpret_cast<double2*>(Sim + (size_t)ps * strideA + d) = b;
                    }
                }
                // closure vector of the slice that produced Y
                const bool fresh = (i == 0 && t.init_kind == 0);
                if (!fresh) {
                    const double2* cl = reinterpret_cast<const double2*>(p.pt.closure) +
                                        (size_t)slice_of(p.pt, n - 1) * chi_pad;
                    for (int d = tid; d < chi_pad; d += N_COMPUTE_WARPS * 32) qbuf[d] = __ldg(cl + d);
                }
                compute_bar();
                // closure rho[ps] = Y[ps, :] . q
                for (int ps = warp; ps < NL; ps += N_COMPUTE_WARPS) {
                    double2 acc = make_double2(0.0, 0.0);
                    const double* xr = Sre + (size_t)ps * strideA;
                    const double* xi = Sim + (size_t)ps * strideA;
                    if (fresh) {
                        if (lane == 0) acc = make_double2(xr[0], xi[0]);
                    } else {
                        for (int d = lane; d < chi_pad; d += 32) {
                            const double2 q = qbuf[d];
                            const double a = xr[d], b = xi[d];
                            acc.x += a * q.x - b * q.y;
                            acc.y += a * q.y + b * q.x;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                    }
                    if (lane == 0) rvec[ps] = acc;
                }
                compute_bar();
                // outputs
                if (i >= t.out_from)
                    for (int o = tid; o < n_out; o += N_COMPUTE_WARPS * 32) {
                        double2 acc = make_double2(0.0, 0.0);
                        for (int a = 0; a < NL; ++a) {
                            const double2 w = OVsm[(size_t)o * NL + a];
                            const double2 r = rvec[pos_s[a]];
                            acc.x += w.x * r.x - w.y * r.y;
                            acc.y += w.x * r.y + w.y * r.x;
                        }
                        reinterpret_cast<double2*>(p.out)[t.out_off + (long long)(i - t.out_from) * n_out + o] = acc;
                    }
                // snapshots of the bond state (natural alpha order, interleaved complex)
                for (int sn = 0; sn < t.snap_cnt; ++sn) {
                    if (p.snap_steps[t.snap_off + sn] != i) continue;
                    double2* dst = reinterpret_cast<double2*>(p.snaps) + (size_t)(t.snap_slot0 + sn) * NL * chi_pad;
                    for (int x = tid; x < NL * chi_pad; x += N_COMPUTE_WARPS * 32) {
                        const int a = x / chi_pad, d = x - a * chi_pad;
                        const size_t o = (size_t)pos_s[a] * strideA + d;
                        dst[x] = make_double2(Sre[o], Sim[o]);
                    }
                }
                // system product X = W Y (FP64 FMA: thread = bond column x half of the output rows)
                if (i < t.n_steps) {
                    double* xr = p.Xre + (size_t)k * NL * row_doubles;
                    double* xi = p.Xim + (size_t)k * NL * row_doubles;
                    const int n_grp = (N_COMPUTE_WARPS * 32) / chi_pad > 0 ? (N_COMPUTE_WARPS * 32) / chi_pad : 1;
                    const int grp = tid / chi_pad;           // which slice of the output rows
                    const int c0 = tid - grp * chi_pad;
                    for (int c = c0; c < chi_pad && grp < n_grp; c += chi_pad) {
                        for (int a0 = 8 * grp; a0 < NL; a0 += 8 * n_grp) {
                            double2 acc[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[j] = make_double2(0.0, 0.0);
                            for (int kk = 0; kk < NL; ++kk) {
                                const size_t so = (size_t)pos_s[kk] * strideA + c;
                                const double yr_ = Sre[so], yi_ = Sim[so];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    if (a0 + j < NL) {
                                        const double2 w = Wsm[(size_t)(a0 + j) * NLp4 + kk];
                                        acc[j].x = fma(w.x, yr_, acc[j].x);
                                        acc[j].x = fma(-w.y, yi_, acc[j].x);
                                        acc[j].y = fma(w.x, yi_, acc[j].y);
                                        acc[j].y = fma(w.y, yr_, acc[j].y);
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (a0 + j < NL) {
                                    const size_t go = (size_t)pos_s[a0 + j] * row_doubles + c;
                                    xr[go] = acc[j].x;
                                    xi[go] = acc[j].y;
                                }
                        }
                    }
                }
                compute_bar();     // shared state / operators are reused by the next task
            }
        }
        grid_barrier(p.barrier, G, epoch);

        // ================================================================= phase 2: class-batched PT GEMM
        if (n < p.n_end) {
            const int s = slice_of(p.pt, n);
            const int nch = p.pt.kin_pad[s] / KC;
            const int nout = p.pt.nout_pad[s];
            const int width = hi - lo;
            // tasks: for every class c, ceil(width * rc / 16) groups of 16 rows
            int u0 = 0;
            for (int c = 0; c < p.n_classes; ++c) {
                const int rc = p.cls_rc[c];
                const int rows_c = width * rc;
                const int nu = (rows_c + SM_ROWS - 1) / SM_ROWS;
                // first task of this class that belongs to this CTA
                int u = (int)blockIdx.x - (u0 % (int)G);
                if (u < 0) u += (int)G;
                for (; u < nu; u += (int)G) {
                    const int q0 = u * SM_ROWS;
                    const int nrows = min(SM_ROWS, rows_c - q0);
                    if (producer) {
                        if (lane == 0) {
                            const double* src = p.pt.blob + p.pt.off[s] + (size_t)p.cls_blk[c] * nch * p.pt.chunk_doubles;
                            const uint32_t bytes = (uint32_t)p.pt.chunk_doubles * 8u;
                            for (int j = 0; j < nch; ++j) {
                                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                                mbar_expect_tx(bar_full + 8 * stage, bytes);
                                bulk_g2s(smem_u32(chunks + (size_t)stage * p.pt.chunk_doubles),
                                         src + (size_t)j * p.pt.chunk_doubles, bytes, bar_full + 8 * stage);
                                if (++stage == stages) { stage = 0; phase ^= 1u; }
                            }
                        }
                        __syncwarp();
                        if (lane != 0) {   // keep the other lanes' pipeline bookkeeping in step
                            for (int j = 0; j < nch; ++j)
                                if (++stage == stages) { stage = 0; phase ^= 1u; }
                        }
                        continue;
                    }
                    // ---- gather the 16 rows (zero rows for trajectories that do not take this step)
                    for (int rr = warp; rr < SM_ROWS; rr += N_COMPUTE_WARPS) {
                        const int q = q0 + rr;
                        bool valid = false;
                        const double *gr = nullptr, *gi = nullptr;
                        if (rr < nrows) {
                            const int k = lo + q / rc, ps = p.cls_p0[c] + q % rc;
                            const aceqd_traj& t = p.trajs[p.order[k]];
                            const int i = n - t.step0;
                            valid = i >= 0 && i < t.n_steps;
                            gr = p.Xre + ((size_t)k * NL + ps) * row_doubles;
                            gi = p.Xim + ((size_t)k * NL + ps) * row_doubles;
                        }
                        double* sr = Sre + (size_t)rr * strideA;
                        double* si = Sim + (size_t)rr * strideA;
                        for (int d = 2 * lane; d < chi_pad; d += 64) {
                            const double2 a = valid ? __ldcg(reinterpret_cast<const double2*>(gr + d)) : make_double2(0.0, 0.0);
                            const double2 b = valid ? __ldcg(reinterpret_cast<const double2*>(gi + d)) : make_double2(0.0, 0.0);
                            *reinterpret_cast<double2*>(sr + d) = a;
                            *reinterpret_cast<double2*>(si + d) = b;
                        }
                    }
                    compute_bar();
                    // ---- GEMM pass over the PT block
                    bool nbv[NB];
                    bool allnb = true, anynb = false;
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        nbv[nb] = 8 * (warp + N_COMPUTE_WARPS * nb) < nout;
                        allnb &= nbv[nb];
                        anynb |= nbv[nb];
                    }
                    double cre[MC][NB][2], cim[MC][NB][2];
#pragma unroll
                    for (int mc = 0; mc < MC; ++mc)
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            cre[mc][nb][0] = cre[mc][nb][1] = 0.0;
                            cim[mc][nb][0] = cim[mc][nb][1] = 0.0;
                        }
                    const double* are[MC];
                    const double* aim[MC];
                    bool aval[MC];
#pragma unroll
                    for (int mc = 0; mc < MC; ++mc) {
                        aval[mc] = true;      // missing rows are zero rows
                        are[mc] = Sre + (size_t)(8 * mc + g) * strideA + tq;
                        aim[mc] = Sim + (size_t)(8 * mc + g) * strideA + tq;
                    }
                    const bool two = nrows > 8;
                    if (!anynb) {
                        for (int jc = 0; jc < nch; ++jc) {
                            mbar_wait(bar_full + 8 * stage, phase);
                            if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
                            if (++stage == stages) { stage = 0; phase ^= 1u; }
                        }
                    } else if (allnb && two)
                        gemm_pass<NB, MC, true>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                                warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
                    else if (allnb)
                        gemm_pass<NB, 1, true>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                               warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
                    else if (two)
                        gemm_pass<NB, MC, false>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                                 warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
                    else
                        gemm_pass<NB, 1, false>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                                warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
                    // ---- epilogue: new Y rows to global memory (columns beyond the slice are zero)
#pragma unroll
                    for (int mc = 0; mc < MC; ++mc) {
                        const int rr = 8 * mc + g;
                        if (rr >= nrows) continue;
                        const int q = q0 + rr;
                        const int k = lo + q / rc, ps = p.cls_p0[c] + q % rc;
                        const aceqd_traj& t = p.trajs[p.order[k]];
                        const int i = n - t.step0;
                        if (i < 0 || i >= t.n_steps) continue;
                        double* yr = p.Yre + ((size_t)k * NL + ps) * row_doubles;
                        double* yi = p.Yim + ((size_t)k * NL + ps) * row_doubles;
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            const int c0 = 8 * (warp + N_COMPUTE_WARPS * nb) + 2 * tq;
                            if (c0 >= chi_pad) continue;
                            const bool v = nbv[nb] && (mc == 0 || two);
                            __stcg(reinterpret_cast<double2*>(yr + c0),
                                   v ? make_double2(cre[mc][nb][0], cre[mc][nb][1]) : make_double2(0.0, 0.0));
                            __stcg(reinterpret_cast<double2*>(yi + c0),
                                   v ? make_double2(cim[mc][nb][0], cim[mc][nb][1]) : make_double2(0.0, 0.0));
                        }
                    }
                    compute_bar();     // the row buffer is refilled by the next task
                }
                u0 += nu;
            }
        }
        grid_barrier(p.barrier, G, epoch);
    }
}

}  // namespace

size_t stream_smem_bytes(int NL, int chi_pad, int stages, int wov_doubles) {
    return stream_layout(NL, chi_pad, stages, wov_doubles).total;
}

template <int NB>
static int launch_stream_nb(const StreamParams& p, size_t smem, cudaStream_t s) {
    ACEQD_CUDA(cudaFuncSetAttribute(k_step_stream<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {const_cast<StreamParams*>(&p)};
    ACEQD_CUDA(cudaLaunchCooperativeKernel((const void*)k_step_stream<NB>, dim3((unsigned)p.grid), dim3(STEP_THREADS),
                                           args, smem, s));
    return ACEQD_OK;
}

int launch_step_stream(const StreamParams& p, size_t smem, cudaStream_t s, long long* launches) {
    const int chi = p.pt.chi_pad;
    if (chi > 256) {
        set_error("chi_pad=%d exceeds the step kernel's 256 limit", chi);
        return ACEQD_ERR_CAPACITY;
    }
    if (p.n_traj <= 0) return ACEQD_OK;
    int rc;
    if (chi <= 64) rc = launch_stream_nb<1>(p, smem, s);
    else if (chi <= 128) rc = launch_stream_nb<2>(p, smem, s);
    else rc = launch_stream_nb<4>(p, smem, s);
    if (rc) return rc;
    ++*launches;
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

}  // namespace aceqd
