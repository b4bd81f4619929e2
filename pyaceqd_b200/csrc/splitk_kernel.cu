// Split-K cluster variant of the fused PT step kernel (SURVEY 2.4 kernel K1) for LARGE Liouville spaces.
//
// The tile kernel (step_kernel.cu) keeps the whole bond state of a tile in ONE CTA's shared memory, which limits a
// tile to T = 4 (NL = 16, chi = 128), 2 (NL = 36) or 1 (NL = 25, chi = 256) trajectories -- and with so few
// trajectories the rows that share a PT block fill the 8-row DMMA m-tiles badly (31 % for the five-level model).
// Here a thread-block cluster of C CTAs shares a tile of G trajectories and splits the BOND index: CTA r holds the
// bond columns K_r = [r NR, (r+1) NR) of every row, so G grows C-fold at the same shared-memory footprint:
//   A  outputs      the owner CTA of a trajectory (j mod C) sums the closure partials of all CTAs: out = OV_n rho
//   B  system       X[:, K_r] = W_n Y[:, K_r]            column-local: nothing is replicated, nothing exchanged
//   C  PT slice     P_r[rows, :] = X[rows, K_r] A_n[beta][K_r, :]   DMMA; CTA r streams only rows K_r of every PT
//                   block (the chunks [r NR/8, (r+1) NR/8) of the existing PT blob); the partial products P_r of all
//                   CTAs are reduce-scattered through distributed shared memory: each accumulator fragment is
//                   stored straight into the receive ring of the CTA that owns its bond columns
//                   (st.shared::cluster), which sums the C partials one pass later, writes the new rows in place
//                   and accumulates the closure partials of its columns.
// Passes of chi_pad > 128 are cut into panels of 128 output columns (a second PT blob ordered by panel).
// Exchange protocol: st.async stores that complete on the owner's mbarrier (tx-count), so no cluster-scope fence
// is needed (a fence.acq_rel.cluster / release-arrive costs ~2k cycles; with plain remote stores + one fence per pass
// the kernel was 5x slower, profiles/r05e..r05j).  Measured and NOT kept: dedicated reducer warps (+ a signal warp doing
// the fence off the GEMM warps' path): three or two reducer warps cannot keep up with eight GEMM warps, the pipeline
// then waits for "slot consumed" (profiles/r05k_*, r05l_*).
// Replaces the inner loop of ACE's Simulation.run (pyaceqd/general_system/general_system.py:331) for the batches
// of two_time/correlations.py:135-184, pol_entanglement/G2.py:439-533 and timebin/twophoton_new.py:515-557.
#include "kernel_common.cuh"

namespace aceqd {

namespace {

constexpr int SK_SKEW = 4;       // doubles of padding after every alpha block of G rows (bank skew for phase B)
constexpr int SK_RS = 2;         // largest receive-ring depth
constexpr int SK_MAXP = 2;       // panels per pass (chi_pad <= 256)
constexpr int SK_META = 256;
constexpr int SK_THREADS = (N_COMPUTE_WARPS + 1) * 32;                    // 8 compute warps + the PT chunk producer

struct SkLayout {
    size_t bar, traj, pass, pos, snapn, rown, rx, rall, q, meta, state, recv, chunks, total;
    size_t plane;   // doubles per state plane
    size_t rslot;   // doubles per receive slot
    int Gown;
};

__host__ __device__ inline size_t sk_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SkLayout sk_layout(int NL, int chi_pad, int G, int C, int NR, int stages, int rslots) {
    SkLayout L;
    const size_t R = (size_t)G * NL;
    const size_t strideA = NR + 4;
    const size_t strideB = (chi_pad > PANEL ? PANEL : chi_pad) + 4;
    L.Gown = (G + C - 1) / C;
    size_t o = 0;
    L.bar = o;    o += 256;
    L.traj = o;   o += sk_align(sizeof(aceqd_traj) * G, 16);
    L.pass = o;   o += sk_align(sizeof(PassDesc) * MAX_PASSES, 16);
    L.pos = o;    o += sk_align(sizeof(int) * MAX_NL, 16);
    L.snapn = o;  o += sk_align(sizeof(int) * MAX_TILE_T, 16);
    L.rown = o;   o += 16 * R;                                   // closure partials of this CTA's columns, every row
    L.rx = o;     o += 16 * (size_t)C * L.Gown * NL;             // partials of all CTAs for the trajectories owned here
    L.rall = o;   o += 16 * (size_t)L.Gown * NL;                 // closure rho of the owned trajectories
    L.q = o;      o += 16 * (size_t)NR;
    L.meta = o;   o += sk_align(sizeof(int) * 2 * SK_META, 16);
    o = sk_align(o, 128);
    L.plane = R * strideA + (size_t)NL * SK_SKEW;
    L.state = o;  o += 2 * L.plane * 8;
    o = sk_align(o, 128);
    L.rslot = (size_t)C * 2 * 16 * NR;
    L.recv = o;   o += (size_t)rslots * L.rslot * 8;
    o = sk_align(o, 128);
    L.chunks = o; o += (size_t)stages * 2 * KC * strideB * 8;
    L.total = o;
    return L;
}

__device__ __forceinline__ void st_cluster_v2(uint32_t cluster_addr, double a, double b) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(cluster_addr), "d"(a), "d"(b) : "memory");
}

template <int NB, int KSU_T>
__global__ void __launch_bounds__(SK_THREADS, 1) k_step_splitk(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int G = p.T, NL = p.prob.NL, R = G * NL, C = p.cluster, NR = p.NR;
    const int chi_pad = p.pt.chi_pad;
    const bool paneled = p.pt.pblob != nullptr;
    const int n_pan = paneled ? p.pt.n_panels : 1;
    const int PW = paneled ? PANEL : 64 * NB;             // output columns of one GEMM pass
    const int strideA = NR + 4;
    const int strideB = (paneled ? PANEL : chi_pad) + 4;
    const int chunk_doubles = 2 * KC * strideB;
    const int stages = p.stages;
    const int RS = p.rslots;                              // receive-ring depth (1 with panels: the panels alternate)
    const SkLayout L = sk_layout(NL, chi_pad, G, C, NR, stages, RS);
    const int Gown = L.Gown;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    aceqd_traj* trj = reinterpret_cast<aceqd_traj*>(smem_raw + L.traj);
    PassDesc* passes = reinterpret_cast<PassDesc*>(smem_raw + L.pass);
    int* pos = reinterpret_cast<int*>(smem_raw + L.pos);
    int* snapn = reinterpret_cast<int*>(smem_raw + L.snapn);
    double2* rown = reinterpret_cast<double2*>(smem_raw + L.rown);
    double2* rx = reinterpret_cast<double2*>(smem_raw + L.rx);
    double2* rall = reinterpret_cast<double2*>(smem_raw + L.rall);
    double2* qbuf = reinterpret_cast<double2*>(smem_raw + L.q);
    int* smeta = reinterpret_cast<int*>(smem_raw + L.meta);
    double* Xre = reinterpret_cast<double*>(smem_raw + L.state);
    double* Xim = Xre + L.plane;
    double* recv = reinterpret_cast<double*>(smem_raw + L.recv);
    double* chunks = reinterpret_cast<double*>(smem_raw + L.chunks);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank();
    const int col0 = (int)crank * NR;                     // first (global) bond column held here
    const int my_q = col0 / PW;                           // the panel whose partial products are reduced here
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 4);
    const uint32_t bar_rfull = smem_u32(bars + 8);        // [SK_RS]   partials of one pass have arrived from every CTA
    const uint32_t bar_rempty = smem_u32(bars + 10);      // [SK_MAXP][SK_RS]  the owners of a panel have consumed a slot
    const uint32_t bar_r = smem_u32(bars + 14);           // closure partials of a step have arrived from every CTA
    auto rowoff = [&](int ps, int j) -> size_t { return (size_t)(ps * G + j) * strideA + (size_t)ps * SK_SKEW; };
    auto rowoff_r = [&](int row) -> size_t { return (size_t)row * strideA + (size_t)(row / G) * SK_SKEW; };
    // owners of a panel: CTAs whose columns lie inside it
    auto owners_of = [&](int q, int& o_lo, int& o_hi) {
        o_lo = q * PW / NR;
        o_hi = min(C, (q + 1) * PW / NR);
    };

    // ------------------------------------------------------------------ setup
    for (int j = tid; j < p.n_pass; j += blockDim.x) passes[j] = p.passes[j];
    for (int j = tid; j < NL; j += blockDim.x) pos[j] = p.prob.pos_of_alpha[j];
    for (int j = tid; j < min(p.pt.n_slices, SK_META); j += blockDim.x) {
        smeta[2 * j] = p.pt.kin_pad[j];
        smeta[2 * j + 1] = p.pt.nout_pad[j];
    }
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, N_COMPUTE_WARPS);
        }
        for (int s = 0; s < SK_RS; ++s) mbar_init(bar_rfull + 8 * s, 1);   // the owner's expect_tx; the partials arrive as st.async
        for (int q = 0; q < SK_MAXP; ++q) {
            int o_lo, o_hi;
            owners_of(q, o_lo, o_hi);
            for (int s = 0; s < SK_RS; ++s) mbar_init(bar_rempty + 8 * (q * SK_RS + s), (uint32_t)max(1, o_hi - o_lo));
        }
        mbar_init(bar_r, 1);      // the owner's expect_tx; closure partials arrive as st.async
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int tile = blockIdx.x / C;
    for (int j = tid; j < G; j += blockDim.x) {
        const int idx = p.tile_traj[(size_t)tile * G + j];
        if (idx >= 0) {
            trj[j] = p.trajs[idx];
        } else {
            aceqd_traj z;
            memset(&z, 0, sizeof(z));
            z.n_steps = -1;
            trj[j] = z;
        }
        snapn[j] = 0;
    }
    for (size_t e = tid; e < 2 * L.plane; e += blockDim.x) Xre[e] = 0.0;
    for (int e = tid; e < R; e += blockDim.x) rown[e] = make_double2(0.0, 0.0);
    for (int e = tid; e < C * Gown * NL; e += blockDim.x) rx[e] = make_double2(0.0, 0.0);
    for (int e = tid; e < Gown * NL; e += blockDim.x) rall[e] = make_double2(0.0, 0.0);
    __syncthreads();
    int n_begin = 0x7fffffff, n_end = -1;
    bool has_snap = false;
    for (int j = 0; j < G; ++j) {
        if (trj[j].n_steps < 0) continue;
        n_begin = min(n_begin, trj[j].step0);
        n_end = max(n_end, trj[j].step0 + trj[j].n_steps);
        has_snap |= trj[j].snap_cnt > 0;
    }
    // initial states: the columns of this CTA
    for (int j = 0; j < G && n_end >= 0; ++j) {
        if (trj[j].n_steps < 0) continue;
        if (trj[j].init_kind == 0) {
            if (crank == 0) {
                const double2* r0 = reinterpret_cast<const double2*>(p.rho0s) + (size_t)trj[j].init_index * NL;
                for (int a = tid; a < NL; a += blockDim.x) {
                    const size_t o = rowoff(pos[a], j);
                    Xre[o] = r0[a].x;
                    Xim[o] = r0[a].y;
                }
            }
        } else {
            const double2* sn = reinterpret_cast<const double2*>(p.snaps) + (size_t)trj[j].init_index * NL * chi_pad;
            for (int e = tid; e < NL * NR; e += blockDim.x) {
                const int a = e / NR, lc = e - a * NR;
                if (col0 + lc < chi_pad) {
                    const double2 v = sn[(size_t)a * chi_pad + col0 + lc];
                    const size_t o = rowoff(pos[a], j) + lc;
                    Xre[o] = v.x;
                    Xim[o] = v.y;
                }
            }
        }
    }
    cluster_sync_all();   // peers' barriers are initialised and their buffers zeroed before anyone writes to them

    if (n_end >= 0) {
    // ------------------------------------------------------------------ producer warp: PT chunks of rows K_r
    if (warp == N_COMPUTE_WARPS) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)chunk_doubles * 8u;
            const int j0 = col0 / KC;
            for (int n = n_begin; n < n_end; ++n) {
                const int s = slice_of(p.pt, n);
                const int nch_s = (s < SK_META ? smeta[2 * s] : p.pt.kin_pad[s]) / KC;
                const int j1 = min(j0 + NR / KC, nch_s);
                for (int m = 0; m < p.n_pass; ++m)
                    for (int q = 0; q < n_pan; ++q) {
                        const double* src = paneled
                            ? p.pt.pblob + p.pt.poff[s] + ((size_t)passes[m].blk * p.pt.n_panels + q) * nch_s * chunk_doubles
                            : p.pt.blob + p.pt.off[s] + (size_t)passes[m].blk * nch_s * chunk_doubles;
                        for (int j = j0; j < j1; ++j) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                            mbar_expect_tx(bar_full + 8 * stage, bytes);
                            bulk_g2s(smem_u32(chunks + (size_t)stage * chunk_doubles), src + (size_t)j * chunk_doubles, bytes,
                                     bar_full + 8 * stage);
                            if (++stage == stages) { stage = 0; phase ^= 1u; }
                        }
                    }
            }
        }
    } else {
    // ------------------------------------------------------------------ compute warps
    const int g = lane >> 2, tq = lane & 3;
    const int n_out = p.prob.n_out;
    const int NLp4 = p.prob.NLp4, MTU = p.prob.NLp8 / 8, KSU = NLp4 / 4;
    const int NTc = NR / 8;                                   // n-tiles of this CTA's columns (system product)
    const int LPR = min(NR, 32), CPL = NR / LPR;              // reduce: lanes per row, columns per lane
    int stage = 0;
    uint32_t phase = 0;
    unsigned mglob = 0u;                                      // m-passes done so far (receive-ring position)
    auto active = [&](const aceqd_traj& t, int n) -> bool { return t.n_steps >= 0 && n >= t.step0 && n < t.step0 + t.n_steps; };

    // optional phase clock (debug): CTA 0 / thread 0 accumulates the cycles between consecutive marks
    long long tick_prev = 0;
    int tick_last = -1;
#define SK_TICK(k)                                                                        \
    do {                                                                                  \
        if (p.ticks && blockIdx.x == 0 && tid == 0) {                                     \
            const long long now_ = clock64();                                             \
            if (tick_last >= 0) p.ticks[tick_last] += now_ - tick_prev;                   \
            tick_prev = now_;                                                             \
            tick_last = (k);                                                              \
        }                                                                                 \
    } while (0)
    for (int n = n_begin; n <= n_end; ++n) {
        SK_TICK(0);
        // ---------------- phase A: closures of the owned trajectories, outputs, snapshots
        if (n > n_begin) {
            mbar_wait(bar_r, (uint32_t)(n - n_begin - 1) & 1u);
            for (int e = tid; e < Gown * NL; e += N_COMPUTE_WARPS * 32) {
                double2 r = rx[e];
                for (int src = 1; src < C; ++src) {
                    const double2 r2 = rx[(size_t)src * Gown * NL + e];
                    r.x += r2.x;
                    r.y += r2.y;
                }
                rall[e] = r;
            }
        }
        compute_bar();
        bool any_snap = false;
        for (int j = 0; j < G; ++j) {
            const aceqd_traj& t = trj[j];
            if (t.n_steps < 0) continue;
            const int i = n - t.step0;
            if (has_snap)
                any_snap |= (snapn[j] < t.snap_cnt && i >= 0 && i <= t.n_steps && p.snap_steps[t.snap_off + snapn[j]] == i);
            if (i == 0 && (uint32_t)(j % C) == crank) {   // a trajectory that starts at this row: its closure is known
                const double2* src = t.init_kind == 0
                    ? reinterpret_cast<const double2*>(p.rho0s) + (size_t)t.init_index * NL
                    : reinterpret_cast<const double2*>(p.snap_r) + (size_t)t.init_index * NL;
                for (int a = tid; a < NL; a += N_COMPUTE_WARPS * 32) rall[(j / C) * NL + pos[a]] = src[a];
            }
        }
        compute_bar();
        for (int it = tid; it < G * n_out; it += N_COMPUTE_WARPS * 32) {
            const int j = it / n_out, o = it - j * n_out;
            const aceqd_traj& t = trj[j];
            if ((uint32_t)(j % C) != crank) continue;
            if (t.n_steps < 0 || n < t.step0 + t.out_from || n > t.step0 + t.n_steps) continue;
            const int i = n - t.step0;
            const long long e = entry_of(t, i, p.ovr_base);
            const double2* ov = reinterpret_cast<const double2*>(p.OV + (size_t)e * p.prob.ov_doubles) + (size_t)o * NL;
            const double2* rj = rall + (j / C) * NL;
            double2 acc = make_double2(0.0, 0.0);
            for (int a = 0; a < NL; ++a) {
                const double2 w = __ldg(ov + a);
                const double2 r = rj[pos[a]];
                acc.x += w.x * r.x - w.y * r.y;
                acc.y += w.x * r.y + w.y * r.x;
            }
            reinterpret_cast<double2*>(p.out)[t.out_off + (long long)(i - t.out_from) * n_out + o] = acc;
        }
        if (any_snap) {
            for (int j = 0; j < G; ++j) {
                const aceqd_traj& t = trj[j];
                if (t.n_steps < 0 || snapn[j] >= t.snap_cnt) continue;
                const int i = n - t.step0;
                if (i < 0 || i > t.n_steps || p.snap_steps[t.snap_off + snapn[j]] != i) continue;
                const size_t slot = (size_t)(t.snap_slot0 + snapn[j]);
                double2* dst = reinterpret_cast<double2*>(p.snaps) + slot * NL * chi_pad;
                for (int e = tid; e < NL * NR; e += N_COMPUTE_WARPS * 32) {
                    const int a = e / NR, lc = e - a * NR;
                    if (col0 + lc < chi_pad) {
                        const size_t o = rowoff(pos[a], j) + lc;
                        dst[(size_t)a * chi_pad + col0 + lc] = make_double2(Xre[o], Xim[o]);
                    }
                }
                if (p.snap_r && (uint32_t)(j % C) == crank)
                    for (int a = tid; a < NL; a += N_COMPUTE_WARPS * 32)
                        reinterpret_cast<double2*>(p.snap_r)[slot * NL + a] = rall[(j / C) * NL + pos[a]];
            }
        }
        if (n == n_end) break;
        if (any_snap) {
            compute_bar();   // snapshot reads of the state precede the in-place system product
            if (tid < G) {
                const aceqd_traj& t = trj[tid];
                const int i = n - t.step0;
                if (t.n_steps >= 0 && snapn[tid] < t.snap_cnt && i >= 0 && i <= t.n_steps &&
                    p.snap_steps[t.snap_off + snapn[tid]] == i)
                    snapn[tid] += 1;
            }
        }

        SK_TICK(1);
        // ---------------- phase B: X[:, K_r] = W_n Y[:, K_r]; unit = (trajectory, n-tile of this CTA's columns)
        for (int u = warp; u < G * NTc; u += N_COMPUTE_WARPS) {
            const int j = u / NTc, nt = u - j * NTc;
            const aceqd_traj& t = trj[j];
            if (!active(t, n)) continue;      // warp-uniform
            const long long e = entry_of(t, n - t.step0, p.ovr_base);
            const double2* Wp = reinterpret_cast<const double2*>(p.W + (size_t)e * p.prob.w_doubles);
            double yre[KSU_T], yim[KSU_T];
#pragma unroll
            for (int ks = 0; ks < KSU_T; ++ks) {
                const int a = 4 * ks + tq;
                const bool ld = ks < KSU && a < NL;
                const size_t o = ld ? rowoff(pos[a], j) + 8 * nt + g : 0;
                yre[ks] = ld ? Xre[o] : 0.0;
                yim[ks] = ld ? Xim[o] : 0.0;
            }
            __syncwarp();       // every lane holds its Y fragments before any row of this unit is overwritten
            for (int mt = 0; mt < MTU; mt += 2) {
                double cr[2][2], ci[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) cr[h][0] = cr[h][1] = ci[h][0] = ci[h][1] = 0.0;
                const bool two = mt + 1 < MTU;
#pragma unroll
                for (int ks = 0; ks < KSU_T; ++ks) {
                    if (ks < KSU) {
                        double2 w[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            w[h] = (h == 0 || two) ? __ldg(Wp + (size_t)(8 * (mt + h) + g) * NLp4 + tq + 4 * ks) : make_double2(0.0, 0.0);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (h == 0 || two) {
                                dmma(cr[h][0], cr[h][1], w[h].x, yre[ks]);
                                dmma(ci[h][0], ci[h][1], w[h].x, yim[ks]);
                            }
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (h == 0 || two) {
                                dmma(cr[h][0], cr[h][1], -w[h].y, yim[ks]);
                                dmma(ci[h][0], ci[h][1], w[h].y, yre[ks]);
                            }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int a = 8 * (mt + h) + g;
                    if ((h == 0 || two) && a < NL) {
                        const size_t o = rowoff(pos[a], j) + 8 * nt + 2 * tq;
                        *reinterpret_cast<double2*>(Xre + o) = make_double2(cr[h][0], cr[h][1]);
                        *reinterpret_cast<double2*>(Xim + o) = make_double2(ci[h][0], ci[h][1]);
                    }
                }
            }
        }
        // ---------------- phase C: PT slice
        const int s = slice_of(p.pt, n);
        const int nch_s = (s < SK_META ? smeta[2 * s] : p.pt.kin_pad[s]) / KC;
        const int nout = s < SK_META ? smeta[2 * s + 1] : p.pt.nout_pad[s];
        const int nch = max(0, min(col0 / KC + NR / KC, nch_s) - col0 / KC);     // chunks of this CTA's rows K_r
        {   // closure of this slice, the columns held here
            const double2* cl = reinterpret_cast<const double2*>(p.pt.closure) + (size_t)s * chi_pad;
            for (int d = tid; d < NR; d += N_COMPUTE_WARPS * 32)
                qbuf[d] = col0 + d < chi_pad ? cl[col0 + d] : make_double2(0.0, 0.0);
        }
        compute_bar();      // system product done everywhere in this CTA

        // reduce one pass: sum the C partial products of the slot, new rows in place, closure partials of these columns
        auto reduce_pass = [&](int m, unsigned mg) {
            const PassDesc& pd = passes[m];
            const int sl = (int)(mg % (unsigned)RS);
            SK_TICK(6);
            mbar_wait(bar_rfull + 8 * sl, (mg / (unsigned)RS) & 1u);
            SK_TICK(7);
            const double* base = recv + (size_t)sl * L.rslot;
            const int lr = tid % LPR;
            for (int mrow = tid / LPR; mrow < 16; mrow += N_COMPUTE_WARPS * 32 / LPR) {
                const int mc = mrow >> 3, gi = mrow & 7;
                const bool valid = gi < pd.nvalid[mc];         // uniform over the LPR lanes of the row
                const int row = pd.row0[mc] + (valid ? gi : 0);
                const aceqd_traj& t = trj[row - (row / G) * G];
                const bool wr = valid && active(t, n);
                double ax = 0.0, ay = 0.0;
                if (wr) {
                    const size_t so = rowoff_r(row) + (size_t)lr * CPL;
                    for (int c = 0; c < CPL; ++c) {
                        double vr = 0.0, vi = 0.0;
                        for (int src = 0; src < C; ++src) {
                            const double* pr = base + ((size_t)(src * 2) * 16 + mrow) * NR + lr * CPL + c;
                            vr += pr[0];
                            vi += pr[(size_t)16 * NR];
                        }
                        Xre[so + c] = vr;
                        Xim[so + c] = vi;
                        const double2 q = qbuf[lr * CPL + c];
                        ax += vr * q.x - vi * q.y;
                        ay += vr * q.y + vi * q.x;
                    }
                }
                for (int o = LPR >> 1; o > 0; o >>= 1) {
                    ax += __shfl_xor_sync(0xffffffffu, ax, o);
                    ay += __shfl_xor_sync(0xffffffffu, ay, o);
                }
                if (wr && lr == 0) rown[row] = make_double2(ax, ay);
            }
            compute_bar();      // every thread has read the slot (the sums are in the state already)
            // ... so every CTA may write into it again: a pure signal, no data travels with it
            if (warp == 0 && lane < C) mbar_arrive_remote_relaxed(mapa(bar_rempty + 8 * (my_q * SK_RS + sl), (uint32_t)lane));
        };

        const int n_pp = p.n_pass * n_pan;
        for (int i = 0; i < n_pp; ++i) {
            const int m = i / n_pan, q = i - m * n_pan;
            const PassDesc pd = passes[m];
            const unsigned mg = mglob + (unsigned)m;
            const int sl = (int)(mg % (unsigned)RS);
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
            bool aval[MC], nbv[NB];
#pragma unroll
            for (int mc = 0; mc < MC; ++mc) {
                aval[mc] = g < pd.nvalid[mc];
                const size_t o = rowoff_r(pd.row0[mc] + (aval[mc] ? g : 0)) + tq;
                are[mc] = Xre + o;
                aim[mc] = Xim + o;
            }
            bool allnb = true, anynb = false;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                nbv[nb] = q * PW + 8 * (warp + N_COMPUTE_WARPS * nb) < nout;
                allnb &= nbv[nb];
                anynb |= nbv[nb];
            }
            const int mcn = pd.nvalid[MC - 1] > 0 ? MC : 1;
            SK_TICK(2);
            if (!anynb) {
                for (int jc = 0; jc < nch; ++jc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            } else if (allnb && mcn == MC)
                gemm_pass<NB, MC, true>(cre, cim, are, aim, aval, nbv, chunks, chunk_doubles, strideB, nch, warp, g, tq,
                                        bar_full, bar_empty, stage, phase, stages, lane);
            else if (allnb)
                gemm_pass<NB, 1, true>(cre, cim, are, aim, aval, nbv, chunks, chunk_doubles, strideB, nch, warp, g, tq,
                                       bar_full, bar_empty, stage, phase, stages, lane);
            else if (mcn == MC)
                gemm_pass<NB, MC, false>(cre, cim, are, aim, aval, nbv, chunks, chunk_doubles, strideB, nch, warp, g, tq,
                                         bar_full, bar_empty, stage, phase, stages, lane);
            else
                gemm_pass<NB, 1, false>(cre, cim, are, aim, aval, nbv, chunks, chunk_doubles, strideB, nch, warp, g, tq,
                                        bar_full, bar_empty, stage, phase, stages, lane);
            // ---- send the partial products to the owners of their columns (the slot's previous contents consumed)
            SK_TICK(3);
            if (mg >= (unsigned)RS) mbar_wait(bar_rempty + 8 * (q * SK_RS + sl), (mg / (unsigned)RS - 1u) & 1u);
            SK_TICK(4);
            // the owners of this panel expect C partial products of the pass's valid rows x their NR columns
            if (q == my_q && tid == 0)
                mbar_expect_tx(bar_rfull + 8 * sl, (uint32_t)C * (uint32_t)(pd.nvalid[0] + pd.nvalid[1]) * (uint32_t)NR * 16u);
            const uint32_t slot_base = smem_u32(recv + (size_t)sl * L.rslot + (size_t)crank * 2 * 16 * NR);
#pragma unroll
            for (int mc = 0; mc < MC; ++mc) {
                if (g >= pd.nvalid[mc]) continue;
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const int gc = q * PW + 8 * (warp + N_COMPUTE_WARPS * nb) + 2 * tq;     // global bond column
                    const int o = gc / NR;
                    if (o >= C) continue;
                    const uint32_t a_re = slot_base + (uint32_t)(((mc * 8 + g) * NR + (gc - o * NR)) * 8);
                    const uint32_t a_im = a_re + (uint32_t)(16 * NR * 8);
                    const uint32_t rb = mapa(bar_rfull + 8 * sl, (uint32_t)o);
                    st_async_v2(mapa(a_re, (uint32_t)o), cre[mc][nb][0], cre[mc][nb][1], rb);
                    st_async_v2(mapa(a_im, (uint32_t)o), cim[mc][nb][0], cim[mc][nb][1], rb);
                }
            }
            compute_bar();      // the X rows of this pass have been read by every warp
            SK_TICK(5);
            // ---- reduce the PREVIOUS pass if its columns live here: its partials have had a whole GEMM pass to land,
            //      and its X rows were last read by the pass just finished (the other panel of the same rows)
            if (i >= 1) {
                const int mp = (i - 1) / n_pan, qp = (i - 1) - mp * n_pan;
                if (qp == my_q) reduce_pass(mp, mglob + (unsigned)mp);
            }
        }
        if (n_pp >= 1 && (n_pp - 1) % n_pan == my_q) reduce_pass((n_pp - 1) / n_pan, mglob + (unsigned)((n_pp - 1) / n_pan));
        mglob += (unsigned)p.n_pass;
        compute_bar();          // closure partials of every row are in rown[]
        // ---------------- closure partials to the owner CTA of each trajectory (counted by its barrier)
        if (tid == 0) {
            const int n_owned = (int)crank < G ? (G - (int)crank + C - 1) / C : 0;
            mbar_expect_tx(bar_r, (uint32_t)C * (uint32_t)n_owned * (uint32_t)NL * 16u);
        }
        for (int row = tid; row < R; row += N_COMPUTE_WARPS * 32) {
            const int ps = row / G, j = row - ps * G;
            const uint32_t o = (uint32_t)(j % C);
            double2* dst = rx + ((size_t)crank * Gown + j / C) * NL + ps;
            const double2 v = rown[row];
            st_async_v2(mapa(smem_u32(dst), o), v.x, v.y, mapa(bar_r, o));
        }
    }
#undef SK_TICK
    }   // compute warps
    }   // non-empty tile
    cluster_sync_all();   // no CTA of a cluster exits while a peer may still address it
}

}  // namespace

int splitk_columns(int chi_pad, int C) {
    int need = (chi_pad + C - 1) / C, nr = 8;
    while (nr < need) nr *= 2;
    return nr;
}

size_t splitk_smem_bytes(int NL, int chi_pad, int G, int C, int stages) {
    if (C != 2 && C != 4 && C != 8) return 0;
    if (G < 1 || G > MAX_TILE_T || chi_pad > 2 * PANEL || chi_pad % 8) return 0;
    const int NR = splitk_columns(chi_pad, C);
    const int PW = chi_pad > PANEL ? PANEL : (chi_pad <= 64 ? 64 : 128);
    if (NR > PW || PW % NR) return 0;
    return sk_layout(NL, chi_pad, G, C, NR, stages, chi_pad > PANEL ? 1 : 2).total;
}

template <int NB>
static int launch_sk(const StepParams& p, size_t smem_bytes, cudaStream_t s) {
    const int ksu = p.prob.NLp4 / 4;
#define ACEQD_LAUNCH_SK(KS)                                                                          \
    do {                                                                                             \
        ACEQD_CUDA(cudaFuncSetAttribute(k_step_splitk<NB, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)smem_bytes));                                           \
        cudaLaunchConfig_t cfg = {};                                                                 \
        cfg.gridDim = dim3((unsigned)(p.n_tiles * p.cluster), 1, 1);                                 \
        cfg.blockDim = dim3(SK_THREADS, 1, 1);                                                       \
        cfg.dynamicSmemBytes = smem_bytes;                                                           \
        cfg.stream = s;                                                                              \
        cudaLaunchAttribute attr[1];                                                                 \
        attr[0].id = cudaLaunchAttributeClusterDimension;                                            \
        attr[0].val.clusterDim.x = (unsigned)p.cluster;                                              \
        attr[0].val.clusterDim.y = 1;                                                                \
        attr[0].val.clusterDim.z = 1;                                                                \
        cfg.attrs = attr;                                                                            \
        cfg.numAttrs = 1;                                                                            \
        ACEQD_CUDA(cudaLaunchKernelEx(&cfg, k_step_splitk<NB, KS>, p));                              \
    } while (0)
    if (ksu <= 4) ACEQD_LAUNCH_SK(4);
    else if (ksu <= 9) ACEQD_LAUNCH_SK(9);
    else ACEQD_LAUNCH_SK(16);
#undef ACEQD_LAUNCH_SK
    return ACEQD_OK;
}

int launch_step_splitk(const StepParams& p, size_t smem_bytes, cudaStream_t s, LaunchLog* log) {
    if (p.n_tiles <= 0) return ACEQD_OK;
    const int chi = p.pt.chi_pad;
    int rc;
    if (chi <= 64) rc = launch_sk<1>(p, smem_bytes, s);
    else rc = launch_sk<2>(p, smem_bytes, s);
    if (rc) return rc;
    ++log->count;
    const int ksu = p.prob.NLp4 / 4;
    log_name(log->step, "k_step_splitk<%d,%d> G=%d cluster=%d NR=%d panels=%d", chi <= 64 ? 1 : 2,
             ksu <= 4 ? 4 : (ksu <= 9 ? 9 : 16), p.T, p.cluster, p.NR, p.pt.pblob ? p.pt.n_panels : 1);
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

}  // namespace aceqd
