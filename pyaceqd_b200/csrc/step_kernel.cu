// The fused PT step kernel (SURVEY 2.4 kernel K1) for sm_100a.
//
// One persistent CTA owns a tile of T trajectories for their whole time evolution.  The bond
// state of the tile ([T*NL rows] x [chi_pad] complex, split re/im planes) never leaves shared
// memory; per absolute time step n the CTA does
//   A  closure   r = Y . q_n ; outputs out = OV_n r ; optional snapshot of Y
//   B  system    X = W_n Y          (DMMA, per trajectory, column-local -> no block barrier)
//   C  PT slice  Y[alpha,:] = X[alpha,:] A_n[beta(alpha)]   (DMMA; PT chunks streamed L2 -> smem
//                by a producer warp with cp.async.bulk + mbarrier full/empty pipeline)
// Rows are stored alpha-major in coupling-class-sorted order (row = pos(alpha)*T + traj) so that
// the 8 rows of a DMMA m-tile share one PT block.  Complex GEMM = 4 real DMMA.8x8x4 per tile.
//
// Replaces the inner loop of ACE's Simulation.run (pyaceqd/general_system/general_system.py:331
// / the `ACE <param>` subprocess of :339-341) for a whole batch of trajectories.
#include "kernel_common.cuh"

namespace aceqd {

namespace {

struct SmemLayout {
    size_t bar, traj, pass, pos, r, rall, own, brow, q, snapn, meta, wov, state, chunks, total;
    size_t plane;  // doubles per state plane
};

constexpr int PB_BUDGET = 16;    // (trajectories x n-tiles x k-steps) of Y fragments a warp keeps in flight in the system product
constexpr int META_SLICES = 256;  // kin/nout of this many slices are cached in shared memory
constexpr int SKEW = 4;  // extra doubles after every alpha block of T rows (bank skew for phase B)

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline SmemLayout make_layout(int NL, int chi_pad, int T, int stages,
                                                  int wov_doubles, int wbufs) {
    SmemLayout L;
    const size_t R = (size_t)T * NL;
    const size_t strideA = chi_pad + 4;
    size_t o = 0;
    L.bar = o;   o += 128;
    L.traj = o;  o += align_up(sizeof(aceqd_traj) * T, 16);
    L.pass = o;  o += align_up(sizeof(PassDesc) * MAX_PASSES, 16);
    L.pos = o;   o += align_up(sizeof(int) * MAX_NL, 16);
    L.snapn = o; o += align_up(sizeof(int) * 2 * MAX_TILE_T, 16);   // snapshot cursors + the step of each next snapshot
    L.r = o;     o += align_up(16 * R * N_COMPUTE_WARPS, 16);   // per-warp partial closures
    L.rall = o;  o += align_up(16 * R, 16);                     // closure rho[row] of the current output row
    L.own = o;   o += align_up(sizeof(int) * MAX_NL, 16);       // alpha position computed by this CTA?
    L.brow = o;  o += align_up(sizeof(int) * (MAX_NL + 8), 16); // the alphas whose rows this CTA computes, in m-tiles of 8
    L.meta = o;  o += align_up(sizeof(int) * 2 * META_SLICES, 16);
    L.q = o;     o += align_up(16 * (size_t)chi_pad, 16);
    o = align_up(o, 128);
    L.wov = o;   o += (size_t)wbufs * T * wov_doubles * 8;   // staged W|OV of the tile (0 = global mode)
    o = align_up(o, 128);
    L.plane = R * strideA + (size_t)NL * SKEW;
    L.state = o; o += 2 * L.plane * 8;
    o = align_up(o, 128);
    L.chunks = o; o += (size_t)stages * 2 * KC * strideA * 8;  // strideB == strideA
    L.total = o;
    return L;
}

// NB   = n-tiles (8 bond columns) per compute warp; KSU_T = compile-time bound on the number of
// DMMA k-steps of the system-operator product (ceil(NL/4) <= KSU_T).
// GPT  = the tile's bond states leave no room for a PT chunk ring: PT fragments come from global memory / L2
//        (gemm_pass_global), no producer warps (256 threads, so that the register-staged fragments fit without spills).
template <int NB, int KSU_T, bool GPT>
__global__ void __launch_bounds__(GPT ? N_COMPUTE_WARPS * 32 : STEP_THREADS, 1) k_step_dmma(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int T = p.T, NL = p.prob.NL, R = T * NL;
    const int chi_pad = p.pt.chi_pad;
    const int strideA = chi_pad + 4;
    const int strideB = p.pt.strideB;
    const int stages = GPT ? 0 : p.stages;
    const int wov = GPT ? 0 : p.wov_doubles;          // 0: operators are read from global memory
    const bool wsm = wov > 0;
    const int wbufs = p.wbufs;              // 2: W(n+1) is prefetched during step n; 1: during phase C of step n
    const SmemLayout L = make_layout(NL, chi_pad, T, stages, wov, wbufs);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    aceqd_traj* trj = reinterpret_cast<aceqd_traj*>(smem_raw + L.traj);
    PassDesc* passes = reinterpret_cast<PassDesc*>(smem_raw + L.pass);
    int* pos = reinterpret_cast<int*>(smem_raw + L.pos);
    int* snapn = reinterpret_cast<int*>(smem_raw + L.snapn);
    int* snapnx = snapn + MAX_TILE_T;   // step (relative to the trajectory's start) of its next snapshot, -1: none left
    double2* rpart = reinterpret_cast<double2*>(smem_raw + L.r);   // [warp][row]
    double2* rall = reinterpret_cast<double2*>(smem_raw + L.rall); // [row]
    int* own_pos = reinterpret_cast<int*>(smem_raw + L.own);       // [alpha position] rows computed here?
    int* brow = reinterpret_cast<int*>(smem_raw + L.brow);         // [m-tile of the system product][8] alpha or -1
    double2* qbuf = reinterpret_cast<double2*>(smem_raw + L.q);
    int* smeta = reinterpret_cast<int*>(smem_raw + L.meta);
    double* Wst = reinterpret_cast<double*>(smem_raw + L.wov);
    double* Xre = reinterpret_cast<double*>(smem_raw + L.state);
    double* Xim = Xre + L.plane;
    double* chunks = reinterpret_cast<double*>(smem_raw + L.chunks);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // a tile may be shared by a cluster of C CTAs: GEMM passes (row blocks of one PT block) are split
    // between the CTAs, every CTA keeps a full copy of the bond state and receives the rows its peers
    // computed through distributed shared memory (cp.async.bulk shared::cta -> shared::cluster)
    const int C = p.cluster;
    const uint32_t crank = C > 1 ? cluster_ctarank() : 0u;
    const uint32_t bar_y = smem_u32(bars + 2 * MAX_STAGES + 4), bar_free = smem_u32(bars + 2 * MAX_STAGES + 5);
    // "the rows I pushed in the previous step have landed in every peer": the bulk copies read their SOURCE rows
    // asynchronously, and the next system product overwrites those rows in place -- it must not start before every
    // peer has seen its exchange barrier complete (found with NL = 36, T = 1 on 8-CTA clusters: 14 copies of 17 KB
    // were still being read when the next step began)
    const uint32_t bar_landed = smem_u32(bars + 2 * MAX_STAGES + 6);
    // hand-over between the compute warps and the pusher warp of a cluster launch (monotonic counters, so that neither side
    // can miss a phase): rows_done = epilogues finished (one count per compute warp and own pass), free_seen = steps
    // for which every peer has read the previous contents of the rows it holds of mine
    uint32_t* ctr = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 7);
    const uint32_t ctr_rows = smem_u32(ctr), ctr_free = smem_u32(ctr + 1);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + MAX_STAGES);
    const uint32_t bar_wfull = smem_u32(bars + 2 * MAX_STAGES), bar_wempty = smem_u32(bars + 2 * MAX_STAGES + 2);
    // offset (doubles) of state row (pos, j) inside a plane; rows are alpha-major with a skew
    auto rowoff = [&](int ps, int j) -> size_t { return (size_t)(ps * T + j) * strideA + (size_t)ps * SKEW; };
    // row / T for row < 1024, T <= 16 by a reciprocal multiplication (exact in that range)
    const unsigned invT = ((1u << 20) + (unsigned)T - 1u) / (unsigned)T;
    auto divT = [&](int row) -> int { return (int)(((unsigned)row * invT) >> 20); };
    auto rowoff_r = [&](int row) -> size_t { return (size_t)row * strideA + (size_t)divT(row) * SKEW; };

    // ------------------------------------------------------------------ setup (once per CTA)
    for (int j = tid; j < p.n_pass; j += blockDim.x) passes[j] = p.passes[j];
    for (int j = tid; j < NL; j += blockDim.x) {
        pos[j] = p.prob.pos_of_alpha[j];
        own_pos[j] = 0;
    }
    for (int j = tid; j < min(p.pt.n_slices, META_SLICES); j += blockDim.x) {
        smeta[2 * j] = p.pt.kin_pad[j];
        smeta[2 * j + 1] = p.pt.nout_pad[j];
    }
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, N_COMPUTE_WARPS);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_wfull + 8 * s, 1);
            mbar_init(bar_wempty + 8 * s, N_COMPUTE_WARPS);
        }
        mbar_init(bar_y, 1);                                 // own expect_tx; rows (bulk copies) and closures (st.async) count bytes
        mbar_init(bar_free, (uint32_t)(C > 1 ? C - 1 : 1));  // "I have read your rows" from every peer
        mbar_init(bar_landed, (uint32_t)(C > 1 ? C - 1 : 1));
        ctr[0] = ctr[1] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    // which alpha positions (blocks of T rows) this CTA computes, and how many bytes its peers push per step
    // (T divides the 8-row m-tile or is a multiple of it, so an alpha block never straddles two passes)
    // the rows of a pass are consecutive in a plane (alpha blocks of T rows separated by SKEW doubles of padding), so a
    // pass travels as ONE bulk copy per plane and peer, padding included
    auto pass_rows = [&](const PassDesc& pd) -> int {
        int n = 0;
#pragma unroll
        for (int mc = 0; mc < MC; ++mc) n += pd.nvalid[mc];
        return n;
    };
    auto pass_plane_bytes = [&](const PassDesc& pd) -> uint32_t {
        const int r0 = pd.row0[0], r1 = r0 + pass_rows(pd) - 1;
        return (uint32_t)(((size_t)r1 * strideA + (size_t)(r1 / T) * SKEW + strideA) -
                          ((size_t)r0 * strideA + (size_t)(r0 / T) * SKEW)) * 8u;
    };
    auto pass_push_bytes = [&](const PassDesc& pd) -> uint32_t { return 2u * pass_plane_bytes(pd); };   // re + im planes
    uint32_t rx_bytes = 0;
    for (int ps = 0; ps < p.n_pass; ++ps) {
        const PassDesc& pd = passes[ps];
        if (pd.owner == (int)crank) {
            if (tid == 0)
                for (int mc = 0; mc < MC; ++mc)
                    for (int r = pd.row0[mc]; r < pd.row0[mc] + pd.nvalid[mc]; r += 1) own_pos[r / T] = 1;
        } else {
            rx_bytes += pass_push_bytes(pd) + 16u * (uint32_t)pass_rows(pd);   // rows + their closures
        }
    }
    if (C > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
    else __syncthreads();
    // system product: a CTA of a cluster needs X only for the rows it multiplies itself.  Its alphas are gathered into
    // consecutive m-tiles (rows of W are read through this table), so that C CTAs share the product instead of each
    // computing nearly all of it (own alphas are scattered over the natural 8-row groups).
    if (tid == 0) {
        int k = 0;
        for (int a = 0; a < NL; ++a)
            if (C == 1 || own_pos[pos[a]]) brow[k++] = a;
        while (k & 7) brow[k++] = -1;
        brow[MAX_NL + 7] = k / 8;     // m-tiles of this CTA's system product
    }
    __syncthreads();
    const int MTB = brow[MAX_NL + 7];
    // Sending the rows of one finished pass to the peers (ONE thread: lane 0 of the pusher warp; thread 0 in the
    // ring-less instantiation, which has no producer warps): wait until every compute warp has written them, before the
    // first copy of a step until every peer has read the previous contents of its copy, then one bulk copy per plane and peer
    uint32_t fph = 0u;         // phase of the "rows read" barrier
    uint32_t push_ctr = 0u;    // epilogue counts of the passes sent so far
    auto push_wait = [&](bool& free_waited) {
        push_ctr += N_COMPUTE_WARPS;
        ctr_wait_ge(ctr_rows, push_ctr);
        if (!free_waited) {
            mbar_wait(bar_free, fph);
            free_waited = true;
            ctr_add_release(ctr_free);
        }
    };
    auto push_copy = [&](const PassDesc& pd, uint32_t peer) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int r0 = pd.row0[0];
        const size_t off = (size_t)r0 * strideA + (size_t)(r0 / T) * SKEW;
        const uint32_t bytes = pass_plane_bytes(pd);
        const uint32_t rb = mapa(bar_y, peer);
        bulk_s2s(mapa(smem_u32(Xre + off), peer), smem_u32(Xre + off), bytes, rb);
        bulk_s2s(mapa(smem_u32(Xim + off), peer), smem_u32(Xim + off), bytes, rb);
    };
    auto push_rows = [&](const PassDesc& pd, bool& free_waited) {      // one thread serves every peer
        push_wait(free_waited);
        for (uint32_t peer = 0; peer < (uint32_t)C; ++peer)
            if (peer != crank) push_copy(pd, peer);
    };
    auto push_step_end = [&](bool free_waited) {
        if (!free_waited) {              // keep the phase in step without own passes
            mbar_wait(bar_free, fph);
            ctr_add_release(ctr_free);
        }
        fph ^= 1u;
    };

    // A CTA works through a list of SEGMENTS (tile, step range): with more tiles than SMs the host lays
    // the tiles end to end and cuts the line into equal pieces, one per CTA (wrap-around rule), so a tile
    // may be started by one CTA (which saves the bond state of the tile to HBM) and finished by the next.
    // Without a segment list a CTA (cluster) runs exactly one whole tile.
    const int seg_first = p.segs ? p.seg_off[blockIdx.x] : 0;
    const int n_seg = p.segs ? p.seg_off[blockIdx.x + 1] - seg_first : 1;
    // pipeline positions persist across segments (producer and consumers advance identically)
    int stage = 0;
    uint32_t phase = 0, wph0 = 0u, wph1 = 0u;
    unsigned chunk_ctr = 0u;   // producers: chunks issued by all producers together so far
    uint32_t yph = 0u, lph = 0u;   // phases of the row-exchange barriers
    uint32_t step_ctr = 0u;    // compute warps: steps with a PT contraction so far
    for (int si = 0; si < n_seg; ++si) {
    SegDesc sg;
    if (p.segs) {
        sg = p.segs[seg_first + si];
    } else {
        sg.tile = blockIdx.x / C;
        sg.n_lo = -0x7fffffff;
        sg.n_hi = 0x7fffffff;
        sg.save_slot = sg.load_slot = -1;
    }
    const int tile = sg.tile;
    if (si > 0) __syncthreads();   // every warp (producer included) has left the previous segment
    for (int j = tid; j < T; j += blockDim.x) {
        const int idx = p.tile_traj[(size_t)tile * T + j];
        if (idx >= 0) {
            trj[j] = p.trajs[idx];
        } else {
            aceqd_traj z;
            memset(&z, 0, sizeof(z));
            z.n_steps = -1;
            trj[j] = z;
        }
        snapn[j] = 0;
        snapnx[j] = (idx >= 0 && p.trajs[idx].snap_cnt > 0) ? p.snap_steps[p.trajs[idx].snap_off] : -1;
    }
    if (sg.load_slot < 0)
        for (size_t e = tid; e < 2 * L.plane; e += blockDim.x) Xre[e] = 0.0;
    __syncthreads();
    int n_begin = 0x7fffffff, n_end = -1;
    for (int j = 0; j < T; ++j) {
        if (trj[j].n_steps < 0) continue;
        n_begin = min(n_begin, trj[j].step0);
        n_end = max(n_end, trj[j].step0 + trj[j].n_steps);
    }
    if (n_end < 0) continue;  // empty tile
    // tile-level facts that let the step loop skip per-trajectory checks
    int s0_max = -1, e_min = 0x7fffffff;
    bool all_valid = true, has_snap = false;
    for (int j = 0; j < T; ++j) {
        if (trj[j].n_steps < 0) {
            all_valid = false;
            continue;
        }
        s0_max = max(s0_max, trj[j].step0);
        e_min = min(e_min, trj[j].step0 + trj[j].n_steps);
        has_snap |= trj[j].snap_cnt > 0;
    }
    const int n_lo = max(n_begin, sg.n_lo), n_hi = min(n_end, sg.n_hi);
    const bool final_seg = n_hi == n_end;   // else: stop BEFORE output row n_hi and save the state
    if (sg.load_slot >= 0) {
        // resume a tile another CTA started: wait for its state, then restore planes, closures, cursors
        if (tid == 0) {
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.seg_flags + sg.load_slot) : "memory");
                if (v != p.seg_epoch) __nanosleep(200);
            } while (v != p.seg_epoch);
        }
        __syncthreads();
        const double2* src = reinterpret_cast<const double2*>(p.seg_state + (size_t)sg.load_slot * p.seg_slot_doubles);
        double2* dst = reinterpret_cast<double2*>(Xre);
        for (size_t e = tid; e < L.plane; e += blockDim.x) dst[e] = src[e];
        for (int e = tid; e < R; e += blockDim.x) rall[e] = src[L.plane + e];
        const int* sn = reinterpret_cast<const int*>(src + L.plane + R);
        for (int e = tid; e < T; e += blockDim.x) {
            snapn[e] = sn[e];
            snapnx[e] = (trj[e].n_steps >= 0 && sn[e] < trj[e].snap_cnt) ? p.snap_steps[trj[e].snap_off + sn[e]] : -1;
        }
    } else {
    // initial states (Y form)
    for (int j = 0; j < T; ++j) {
        if (trj[j].n_steps < 0) continue;
        if (trj[j].init_kind == 0) {
            const double2* r0 = reinterpret_cast<const double2*>(p.rho0s) + (size_t)trj[j].init_index * NL;
            for (int a = tid; a < NL; a += blockDim.x) {
                const size_t o = rowoff(pos[a], j);
                Xre[o] = r0[a].x;
                Xim[o] = r0[a].y;
            }
        } else {
            const double2* sn = reinterpret_cast<const double2*>(p.snaps) +
                                (size_t)trj[j].init_index * NL * chi_pad;
            for (int e = tid; e < NL * chi_pad; e += blockDim.x) {
                const int a = e / chi_pad, d = e - a * chi_pad;
                const size_t o = rowoff(pos[a], j) + d;
                const double2 v = sn[e];
                Xre[o] = v.x;
                Xim[o] = v.y;
            }
        }
    }
    }
    // closure of the slice before the first row (used by snapshot-started trajectories)
    if (n_lo > 0) {
        const double2* cl = reinterpret_cast<const double2*>(p.pt.closure) +
                            (size_t)slice_of(p.pt, n_lo - 1) * chi_pad;
        for (int d = tid; d < chi_pad; d += blockDim.x) qbuf[d] = cl[d];
    }
    __syncthreads();

    // ------------------------------------------------------------------ producer warps
    // warps 8 .. 8+P-1 feed the PT chunk ring: chunk number c (counted over passes, steps and segments) goes to
    // stage c mod S and is issued by producer c mod P, P <= S (with P <= S a producer can never be two ring
    // generations ahead of the consumers, so the parity wait on the stage's empty barrier is unambiguous);
    // the last producer warp stages the per-row operators W_n | OV_n.
    if (warp >= N_COMPUTE_WARPS) {
        const int pw = warp - N_COMPUTE_WARPS;
        const int P = min(N_CHUNK_PRODUCERS, stages);
        if (lane == 0 && pw < P) {
            const uint32_t bytes = (uint32_t)p.pt.chunk_doubles * 8u;
            for (int n = n_lo; n < n_hi; ++n) {
                const int s = slice_of(p.pt, n);
                const int nch = (s < META_SLICES ? smeta[2 * s] : p.pt.kin_pad[s]) / KC;
                const double* sl = p.pt.blob + p.pt.off[s];
                for (int ps = 0; ps < p.n_pass; ++ps) {
                    if (passes[ps].owner != (int)crank) continue;
                    const double* src = sl + (size_t)passes[ps].blk * nch * p.pt.chunk_doubles;
                    for (int j = (int)((pw + P - chunk_ctr % P) % P); j < nch; j += P) {
                        const unsigned c = chunk_ctr + (unsigned)j;
                        const int st = (int)(c % (unsigned)stages);
                        const uint32_t ph = (c / (unsigned)stages) & 1u;
                        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
                        mbar_expect_tx(bar_full + 8 * st, bytes);
                        bulk_g2s(smem_u32(chunks + (size_t)st * p.pt.chunk_doubles),
                                 src + (size_t)j * p.pt.chunk_doubles, bytes, bar_full + 8 * st);
                    }
                    chunk_ctr += (unsigned)nch;
                }
            }
        } else if (lane == 0 && pw == N_CHUNK_PRODUCERS && wsm) {
            const uint32_t w_bytes = (uint32_t)p.prob.w_doubles * 8u, ov_bytes = (uint32_t)p.prob.ov_doubles * 8u;
            // stage the per-row operators W_n | OV_n of every active trajectory into buffer n & 1 (or the single
            // buffer, which the consumers hand back after the system product of row n)
            auto issue_wov = [&](int n) {
                const int buf = wbufs == 2 ? (n & 1) : 0;
                mbar_wait(bar_wempty + 8 * buf, (buf ? wph1 : wph0) ^ 1u);
                if (buf) wph1 ^= 1u; else wph0 ^= 1u;
                uint32_t total = 0;
                for (int j = 0; j < T; ++j) {
                    const aceqd_traj& t = trj[j];
                    if (t.n_steps >= 0 && n >= t.step0 && n <= t.step0 + t.n_steps) total += w_bytes + ov_bytes;
                }
                mbar_expect_tx(bar_wfull + 8 * buf, total);
                for (int j = 0; j < T; ++j) {
                    const aceqd_traj& t = trj[j];
                    if (t.n_steps < 0 || n < t.step0 || n > t.step0 + t.n_steps) continue;
                    const long long e = entry_of(t, n - t.step0, p.ovr_base);
                    double* dst = Wst + (size_t)(buf * T + j) * wov;
                    bulk_g2s(smem_u32(dst), p.W + (size_t)e * p.prob.w_doubles, w_bytes, bar_wfull + 8 * buf);
                    bulk_g2s(smem_u32(dst + p.prob.w_doubles), p.OV + (size_t)e * p.prob.ov_doubles, ov_bytes,
                             bar_wfull + 8 * buf);
                }
            };
            issue_wov(n_lo);
            for (int n = n_lo; n < n_hi; ++n)
                if (n + 1 < n_hi || final_seg) issue_wov(n + 1);   // row n_hi of an unfinished tile belongs to the next segment
        } else if (pw == N_CHUNK_PRODUCERS + 1 && C > 1) {
            // row pusher: as soon as all compute warps have written the new rows of one of this CTA's passes, copy them
            // into every peer's state (cp.async.bulk shared::cta -> shared::cluster, completing on the peer's exchange
            // barrier).  Issuing a bulk copy costs its thread ~500 cycles (profiles/r05b_bulk_latency.txt): on a warp
            // of its own that is off the compute warps' critical path, and the rows leave one pass earlier than when
            // thread 0 sent them at the next block barrier.
            // Lane 0 waits, lane r copies to peer r: bulk copies issued by different lanes overlap, a thread's own
            // consecutive copies do not (profiles/r05b_bulk_latency.txt).
            for (int n = n_lo; n < n_hi; ++n) {
                bool free_waited = false;
                for (int ps = 0; ps < p.n_pass; ++ps) {
                    if (passes[ps].owner != (int)crank) continue;
                    if (lane == 0) push_wait(free_waited);
                    __syncwarp();
                    if (lane < C && (uint32_t)lane != crank) push_copy(passes[ps], (uint32_t)lane);
                }
                if (lane == 0) push_step_end(free_waited);
            }
        }
        continue;   // next segment (all lanes meet the compute warps at its first barrier)
    }

    // ------------------------------------------------------------------ compute warps
    const int g = lane >> 2, tq = lane & 3;  // DMMA fragment coordinates
    const int NT = chi_pad / 8;
    const int n_out = p.prob.n_out;
    const int NLp4 = p.prob.NLp4, MTU = p.prob.NLp8 / 8, KSU = NLp4 / 4;

    // optional phase clock (debug): CTA 0 / thread 0 accumulates the cycles between consecutive marks
    long long tick_prev = 0;
    int tick_last = -1;
#define TICK(k)                                                                           \
    do {                                                                                  \
        if (p.ticks && blockIdx.x == 0 && tid == 0) {                                     \
            const long long now_ = clock64();                                             \
            if (tick_last >= 0) p.ticks[tick_last] += now_ - tick_prev;                   \
            tick_prev = now_;                                                             \
            tick_last = (k);                                                              \
        }                                                                                 \
    } while (0)
    for (int n = n_lo; n <= n_hi; ++n) {
        if (n == n_hi && !final_seg) break;   // the next segment of this tile starts with output row n_hi
        const int buf = wbufs == 2 ? (n & 1) : 0;
        if (C > 1) {
            if (n > n_lo) {           // rows and closures computed by the peers in step n-1 have landed
                mbar_wait(bar_y, yph);
                yph ^= 1u;
                // tell every peer that ITS rows have landed here (a pure signal: no data travels with it)
                if (warp == 0 && lane < C && (uint32_t)lane != crank) mbar_arrive_remote_relaxed(mapa(bar_landed, (uint32_t)lane));
            }
            if (n < n_end && tid == 0) mbar_expect_tx(bar_y, rx_bytes);   // arm this step's exchange
        }
        TICK(0);
        // Operators that cannot be staged in shared memory (large Liouville spaces) are read by every warp straight from
        // global memory in the system product: pull this row's W into L1 now, so that those reads find it there instead
        // of waiting for L2 two k-steps ahead of their DMMAs (all 8 warps read the same 11-23 KB per trajectory).
        if (!wsm && KSU_T > 4 && n < n_end) {
            for (int j = 0; j < T; ++j) {
                const aceqd_traj& t = trj[j];
                if (t.n_steps < 0 || n < t.step0 || n >= t.step0 + t.n_steps) continue;
                const char* wp = reinterpret_cast<const char*>(p.W + (size_t)entry_of(t, n - t.step0, p.ovr_base) * p.prob.w_doubles);
                for (int o = tid * 128; o < p.prob.w_doubles * 8; o += N_COMPUTE_WARPS * 32 * 128)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + o));
            }
        }
        // ---------------- phase A: outputs (closures rall[] come from the previous GEMM epilogue)
        // rows of trajectories that START at this row have no closure yet: generic closure pass
        const bool full_act = all_valid && n >= s0_max && n < e_min;   // every trajectory of the tile steps n -> n+1
        bool any_start = false, any_snap = false;
        if (n <= s0_max || has_snap)
        for (int j = 0; j < T; ++j) {
            const aceqd_traj& t = trj[j];
            if (t.n_steps < 0) continue;
            any_start |= (n == t.step0);
            const int i = n - t.step0;
            any_snap |= (snapn[j] < t.snap_cnt && i >= 0 && i <= t.n_steps && snapnx[j] == i);
        }
        if (any_start) {
            for (int row = warp; row < R; row += N_COMPUTE_WARPS) {
                const int j = row - divT(row) * T;
                const aceqd_traj& t = trj[j];
                if (t.n_steps < 0 || n != t.step0) continue;
                double2 acc = make_double2(0.0, 0.0);
                const double* xr = Xre + rowoff_r(row);
                const double* xi = Xim + rowoff_r(row);
                if (t.init_kind == 0) {
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
                if (lane == 0) rall[row] = acc;
            }
            compute_bar();
        }
        if (wsm) mbar_wait(bar_wfull + 8 * buf, buf ? wph1 : wph0);
        TICK(1);
        // out[j][o] = OV_n[o] . rho_j: for large Liouville spaces LPO lanes share one functional (a serial sum over NL = 16
        // or more on a handful of threads took 2.4k cycles per step of the cfg3 branch launch, profiles/r05o_*)
        {
            const int LPO = NL >= 16 ? 16 : 1;
            const int items = T * n_out * LPO;
            for (int base = 0; base < items; base += N_COMPUTE_WARPS * 32) {     // block-uniform trip count
                const int idx = base + tid;
                const int it = idx / LPO, l = idx - it * LPO;
                const int j = it / n_out, o = it - j * n_out;
                bool on = idx < items;
                if (on && C > 1 && (uint32_t)(j % C) != crank) on = false;   // every CTA holds all closures: split the writes
                double2 acc = make_double2(0.0, 0.0);
                int i = 0;
                if (on) {
                    const aceqd_traj& t = trj[j];
                    on = !(t.n_steps < 0 || n < t.step0 + t.out_from || n > t.step0 + t.n_steps);
                    if (on) {
                        i = n - t.step0;
                        const double2* ov;
                        if (wsm) {
                            ov = reinterpret_cast<const double2*>(Wst + (size_t)(buf * T + j) * wov + p.prob.w_doubles) +
                                 (size_t)o * NL;
                        } else {
                            const long long e = entry_of(t, i, p.ovr_base);
                            ov = reinterpret_cast<const double2*>(p.OV + (size_t)e * p.prob.ov_doubles) + (size_t)o * NL;
                        }
                        for (int a = l; a < NL; a += LPO) {
                            const double2 w = ov[a];
                            const double2 r = rall[pos[a] * T + j];
                            acc.x += w.x * r.x - w.y * r.y;
                            acc.y += w.x * r.y + w.y * r.x;
                        }
                    }
                }
                for (int sh = LPO >> 1; sh > 0; sh >>= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sh);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sh);
                }
                if (on && l == 0) {
                    const aceqd_traj& t = trj[j];
                    reinterpret_cast<double2*>(p.out)[t.out_off + (long long)(i - t.out_from) * n_out + o] = acc;
                }
            }
        }
        if (any_snap) {
            // every CTA of a cluster holds the whole bond state: each writes its share of the snapshot (the trunk of a
            // G2 map snapshots at every step -- 32 KB per step written by rank 0 alone were 11 % of the trunk's step)
            for (int j = 0; j < T; ++j) {
                const aceqd_traj& t = trj[j];
                if (t.n_steps < 0 || snapn[j] >= t.snap_cnt) continue;
                const int i = n - t.step0;
                if (i < 0 || i > t.n_steps || snapnx[j] != i) continue;
                double2* dst = reinterpret_cast<double2*>(p.snaps) +
                               (size_t)(t.snap_slot0 + snapn[j]) * NL * chi_pad;
                for (int e = tid + (int)crank * N_COMPUTE_WARPS * 32; e < NL * chi_pad; e += C * N_COMPUTE_WARPS * 32) {
                    const int a = e / chi_pad, d = e - a * chi_pad;
                    const size_t o = rowoff(pos[a], j) + d;
                    dst[e] = make_double2(Xre[o], Xim[o]);
                }
                // the closure of this row travels with the snapshot (column-distributed kernels start from it)
                if (p.snap_r && crank == 0)
                    for (int a = tid; a < NL; a += N_COMPUTE_WARPS * 32)
                        reinterpret_cast<double2*>(p.snap_r)[(size_t)(t.snap_slot0 + snapn[j]) * NL + a] = rall[pos[a] * T + j];
            }
        }
        if (n == n_end) {
            if (wsm) {  // last row: hand the buffer back (the CTA may go on with another segment)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_wempty + 8 * buf);
                if (buf) wph1 ^= 1u; else wph0 ^= 1u;
            }
            break;
        }
        if (any_snap) {
            compute_bar();   // snapshot reads of the state precede phase B's in-place update
            if (tid < T) {   // advance snapshot cursors (read again only after later barriers)
                const aceqd_traj& t = trj[tid];
                const int i = n - t.step0;
                if (t.n_steps >= 0 && snapn[tid] < t.snap_cnt && i >= 0 && i <= t.n_steps && snapnx[tid] == i) {
                    const int k = snapn[tid] + 1;
                    snapn[tid] = k;
                    snapnx[tid] = k < t.snap_cnt ? p.snap_steps[t.snap_off + k] : -1;   // the only global read, off the step's path
                }
            }
        }

        if (C > 1 && n > n_lo) {      // my pushes of step n-1 have been read out of my rows: they may be overwritten
            mbar_wait(bar_landed, lph);
            lph ^= 1u;
        }
        TICK(2);
        // ---------------- phase B: X = W_n Y.  Warp w owns bond columns of its n-tiles for every
        // trajectory (column-local, in place, no block barrier); JU trajectories x NBB n-tiles are
        // kept in flight for instruction-level parallelism (bounded by the register budget).
        if constexpr (KSU_T == 1) {
            // NL <= 4 (two-level system): a DMMA m-tile would be half empty and its operand traffic heavy.
            // The FP64 FMA pipe has the same peak on this part (profiles/r01_fp64_peaks.log): warp w takes
            // trajectories w, w+8, ..., keeps W_n in registers, lane = bond column, 16 complex FMAs per column.
            // With more than one trajectory per warp the two half-warps take one trajectory each (16 bond columns per
            // sweep) so that the operator loads and the loop overhead of both are paid once, side by side.
            const int LPT = T > N_COMPUTE_WARPS ? 16 : 32;        // lanes per trajectory
            const int per_warp = 32 / LPT;
            const int sub = lane / LPT, cl = lane - sub * LPT;
            for (int jb = warp * per_warp; jb < T; jb += N_COMPUTE_WARPS * per_warp) {
                const int j = jb + sub;
                bool actj = false;
                if (j < T) {
                    const aceqd_traj& t = trj[j];
                    actj = full_act || (t.n_steps >= 0 && n >= t.step0 && n < t.step0 + t.n_steps);
                }
                if (!actj) continue;     // no warp-level synchronisation inside this loop
                const aceqd_traj& t = trj[j];
                const double2* Wp;
                if (wsm) {
                    Wp = reinterpret_cast<const double2*>(Wst + (size_t)(buf * T + j) * wov);
                } else {
                    const long long e = entry_of(t, n - t.step0, p.ovr_base);
                    Wp = reinterpret_cast<const double2*>(p.W + (size_t)e * p.prob.w_doubles);
                }
                double2 w[4][4];
                int ro[4];
                bool own[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    ro[a] = a < NL ? (int)rowoff(pos[a], j) : 0;
                    own[a] = a < NL && (C == 1 || own_pos[pos[a]]);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        w[a][k] = (a < NL && k < NL) ? (wsm ? Wp[a * NLp4 + k] : __ldg(Wp + a * NLp4 + k))
                                                     : make_double2(0.0, 0.0);
                }
                for (int c = cl; c < chi_pad; c += LPT) {
                    double2 y[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        y[k] = k < NL ? make_double2(Xre[ro[k] + c], Xim[ro[k] + c]) : make_double2(0.0, 0.0);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        double xr = 0.0, xi = 0.0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            xr = fma(w[a][k].x, y[k].x, xr);
                            xr = fma(-w[a][k].y, y[k].y, xr);
                            xi = fma(w[a][k].x, y[k].y, xi);
                            xi = fma(w[a][k].y, y[k].x, xi);
                        }
                        if (own[a]) {
                            Xre[ro[a] + c] = xr;
                            Xim[ro[a] + c] = xi;
                        }
                    }
                }
            }
        } else if constexpr (KSU_T > 4) {
            // Large Liouville spaces (NL = 25, 36: KSU_T = 9, 16).  k-steps outermost: the accumulators of ALL m-tiles of
            // this CTA's rows are live at once (2 x MTB independent DMMA chains instead of 2), the Y fragment of a
            // k-step is read from shared memory when it is needed, and the W fragments -- which come from global memory /
            // L2 when the bond states leave no room to stage the operators -- are prefetched WS - 1 k-steps ahead.  (The
            // m-tile-outer loop below re-reads W per n-tile with two dependent chains: 140k cycles per step for NL = 36,
            // T = 2 against a DMMA floor of 23k, profiles/r06l_ticks_shapes.txt; with this loop 52k, sixls shape
            // 33.8 -> 23.3 ms.)
            constexpr int MTM = (KSU_T * 4 + 7) / 8;          // m-tiles of NLp8 rows
            constexpr int WS = KSU_T > 9 ? 2 : 3;             // W fragment sets in flight
            // ring-less chi = 256 tiles: their register-staged GEMM fills the register file, so the W fragments of only two
            // m-tiles are in registers at a time (the Y fragments are read once per group instead of once)
            constexpr int MG = (GPT && NB == 4) ? 2 : MTM;
            int arow[MTM];
#pragma unroll
            for (int mt = 0; mt < MTM; ++mt) arow[mt] = mt < MTB ? brow[8 * mt + g] : -1;
            for (int j = 0; j < T; ++j) {
                const aceqd_traj& t = trj[j];
                if (!(full_act || (t.n_steps >= 0 && n >= t.step0 && n < t.step0 + t.n_steps))) continue;   // warp-uniform
                const double2* Wp;
                if (wsm) {
                    Wp = reinterpret_cast<const double2*>(Wst + (size_t)(buf * T + j) * wov);
                } else {
                    const long long e = entry_of(t, n - t.step0, p.ovr_base);
                    Wp = reinterpret_cast<const double2*>(p.W + (size_t)e * p.prob.w_doubles);
                }
                for (int nb = 0; nb < NB; ++nb) {
                    const int nt = warp + N_COMPUTE_WARPS * nb;
                    if (nt >= NT) break;                      // warp-uniform
                    const int ncol = 8 * nt;
                    double cr[MTM][2], ci[MTM][2];
#pragma unroll
                    for (int mt = 0; mt < MTM; ++mt) cr[mt][0] = cr[mt][1] = ci[mt][0] = ci[mt][1] = 0.0;
#pragma unroll
                    for (int m0 = 0; m0 < MTM; m0 += MG) {       // groups of MG m-tiles share the W fragment registers
                        if (m0 >= MTB) break;                     // warp-uniform
                        double2 w[WS][MG];
                        auto loadW = [&](int slot, int ks) {
#pragma unroll
                            for (int mg = 0; mg < MG; ++mg) {
                                w[slot][mg] = make_double2(0.0, 0.0);
                                if (m0 + mg < MTM && arow[m0 + mg < MTM ? m0 + mg : 0] >= 0) {
                                    const double2* wp = Wp + (size_t)arow[m0 + mg] * NLp4 + tq + 4 * ks;
                                    w[slot][mg] = wsm ? *wp : __ldg(wp);
                                }
                            }
                        };
#pragma unroll
                        for (int q = 0; q < WS - 1; ++q)
                            if (q < KSU) loadW(q, q);
#pragma unroll
                        for (int ks = 0; ks < KSU_T; ++ks) {
                            if (ks < KSU) {
                                if (ks + WS - 1 < KSU) loadW((ks + WS - 1) % WS, ks + WS - 1);
                                const int a = 4 * ks + tq;
                                double yr = 0.0, yi = 0.0;
                                if (a < NL) {
                                    const size_t o = rowoff(pos[a], j) + g + ncol;
                                    yr = Xre[o];
                                    yi = Xim[o];
                                }
#pragma unroll
                                for (int mg = 0; mg < MG; ++mg)
                                    if (m0 + mg < MTM && m0 + mg < MTB) {
                                        dmma(cr[m0 + mg][0], cr[m0 + mg][1], w[ks % WS][mg].x, yr);
                                        dmma(ci[m0 + mg][0], ci[m0 + mg][1], w[ks % WS][mg].x, yi);
                                    }
#pragma unroll
                                for (int mg = 0; mg < MG; ++mg)
                                    if (m0 + mg < MTM && m0 + mg < MTB) {
                                        dmma(cr[m0 + mg][0], cr[m0 + mg][1], -w[ks % WS][mg].y, yi);
                                        dmma(ci[m0 + mg][0], ci[m0 + mg][1], w[ks % WS][mg].y, yr);
                                    }
                            }
                        }
                    }
                    __syncwarp();       // every lane has read the Y rows of these columns: write X over them
#pragma unroll
                    for (int mt = 0; mt < MTM; ++mt)
                        if (arow[mt] >= 0) {
                            const size_t o = rowoff(pos[arow[mt]], j) + ncol + 2 * tq;
                            *reinterpret_cast<double2*>(Xre + o) = make_double2(cr[mt][0], cr[mt][1]);
                            *reinterpret_cast<double2*>(Xim + o) = make_double2(ci[mt][0], ci[mt][1]);
                        }
                    __syncwarp();
                }
            }
        } else {
        constexpr int NBB = (KSU_T * NB <= PB_BUDGET) ? NB : 1;
        constexpr int JU_ = PB_BUDGET / (KSU_T * NBB);
        constexpr int JU = JU_ < 1 ? 1 : (JU_ > 4 ? 4 : JU_);
        for (int j0 = 0; j0 < T; j0 += JU) {
            bool act[JU];
            const double2* Wp[JU];
            bool any = false;
#pragma unroll
            for (int jj = 0; jj < JU; ++jj) {
                const int j = j0 + jj;
                act[jj] = false;
                Wp[jj] = nullptr;
                if (j < T) {
                    const aceqd_traj& t = trj[j];
                    act[jj] = t.n_steps >= 0 && n >= t.step0 && n < t.step0 + t.n_steps;
                    if (act[jj]) {
                        if (wsm) {
                            Wp[jj] = reinterpret_cast<const double2*>(Wst + (size_t)(buf * T + j) * wov);
                        } else {
                            const long long e = entry_of(t, n - t.step0, p.ovr_base);
                            Wp[jj] = reinterpret_cast<const double2*>(p.W + (size_t)e * p.prob.w_doubles);
                        }
                    }
                }
                any |= act[jj];
            }
            if (!any) continue;
            for (int nb0 = 0; nb0 < NB; nb0 += NBB) {
                bool nbB[NBB];
                int ncol[NBB];
#pragma unroll
                for (int nb = 0; nb < NBB; ++nb) {
                    const int nt = warp + N_COMPUTE_WARPS * (nb0 + nb);
                    nbB[nb] = nt < NT;
                    ncol[nb] = 8 * nt;
                }
                double yre[JU][NBB][KSU_T], yim[JU][NBB][KSU_T];
#pragma unroll
                for (int jj = 0; jj < JU; ++jj)
#pragma unroll
                    for (int ks = 0; ks < KSU_T; ++ks) {
                        const int a = 4 * ks + tq;
                        const bool ld = act[jj] && ks < KSU && a < NL;
                        const size_t o = ld ? rowoff(pos[a], j0 + jj) + g : 0;
#pragma unroll
                        for (int nb = 0; nb < NBB; ++nb) {
                            const bool l2 = ld && nbB[nb];
                            yre[jj][nb][ks] = l2 ? Xre[o + ncol[nb]] : 0.0;
                            yim[jj][nb][ks] = l2 ? Xim[o + ncol[nb]] : 0.0;
                        }
                    }
                __syncwarp();
                for (int mt = 0; mt < MTB; ++mt) {     // m-tiles of the rows this CTA feeds into its own GEMM passes
                    const int arow = brow[8 * mt + g];  // alpha of this lane's W row (and output row), -1 = padding
                    double cr[JU][NBB][2], ci[JU][NBB][2];
#pragma unroll
                    for (int jj = 0; jj < JU; ++jj)
#pragma unroll
                        for (int nb = 0; nb < NBB; ++nb)
                            cr[jj][nb][0] = cr[jj][nb][1] = ci[jj][nb][0] = ci[jj][nb][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KSU_T; ++ks) {
                        if (ks < KSU) {
                            double2 w[JU];
#pragma unroll
                            for (int jj = 0; jj < JU; ++jj) {
                                w[jj] = make_double2(0.0, 0.0);
                                if (act[jj] && arow >= 0) {
                                    const double2* wp = Wp[jj] + (size_t)arow * NLp4 + tq + 4 * ks;
                                    w[jj] = wsm ? *wp : __ldg(wp);
                                }
                            }
#pragma unroll
                            for (int jj = 0; jj < JU; ++jj)
#pragma unroll
                                for (int nb = 0; nb < NBB; ++nb)
                                    if (act[jj] && nbB[nb]) {
                                        dmma(cr[jj][nb][0], cr[jj][nb][1], w[jj].x, yre[jj][nb][ks]);
                                        dmma(ci[jj][nb][0], ci[jj][nb][1], w[jj].x, yim[jj][nb][ks]);
                                    }
#pragma unroll
                            for (int jj = 0; jj < JU; ++jj)
#pragma unroll
                                for (int nb = 0; nb < NBB; ++nb)
                                    if (act[jj] && nbB[nb]) {
                                        dmma(cr[jj][nb][0], cr[jj][nb][1], -w[jj].y, yim[jj][nb][ks]);
                                        dmma(ci[jj][nb][0], ci[jj][nb][1], w[jj].y, yre[jj][nb][ks]);
                                    }
                        }
                    }
                    const int a = arow;
                    if (a >= 0) {
#pragma unroll
                        for (int jj = 0; jj < JU; ++jj)
#pragma unroll
                            for (int nb = 0; nb < NBB; ++nb)
                                if (act[jj] && nbB[nb]) {
                                    const size_t o = rowoff(pos[a], j0 + jj) + ncol[nb] + 2 * tq;
                                    *reinterpret_cast<double2*>(Xre + o) = make_double2(cr[jj][nb][0], cr[jj][nb][1]);
                                    *reinterpret_cast<double2*>(Xim + o) = make_double2(ci[jj][nb][0], ci[jj][nb][1]);
                                }
                    }
                }
                __syncwarp();
            }
        }
        }
        if (wsm) {  // this row's operators are consumed: hand the buffer back to the producer
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_wempty + 8 * buf);
            if (buf) wph1 ^= 1u; else wph0 ^= 1u;
        }
        compute_bar();
        TICK(3);
        // every warp of this CTA has read the peers' rows (their values are in registers): they may overwrite them
        if (C > 1 && warp == 0 && lane < C && (uint32_t)lane != crank) mbar_arrive_remote_relaxed(mapa(bar_free, (uint32_t)lane));

        // ---------------- phase C: PT slice, Y = X A_n[beta]
        const int s = slice_of(p.pt, n);
        const int nch = (s < META_SLICES ? smeta[2 * s] : p.pt.kin_pad[s]) / KC;
        const int nout = s < META_SLICES ? smeta[2 * s + 1] : p.pt.nout_pad[s];
        {   // stage the closure of this slice: consumed by the GEMM epilogue below
            const double2* cl = reinterpret_cast<const double2*>(p.pt.closure) + (size_t)s * chi_pad;
            for (int d = tid; d < chi_pad; d += N_COMPUTE_WARPS * 32) qbuf[d] = cl[d];
        }
        bool nbv[NB];
#pragma unroll
        // ring-less tiles and two n-tiles per warp: a warp owns 8 NB CONSECUTIVE columns (vector loads of the PT fragments
        // in gemm_pass_global / gemm_pass), all or none of them inside the slice
        // (not for the two-level instantiation KSU_T = 1: its passes always carry two full m-tiles and it measured 0.3 %
        // slower with the vector loads, 31.30 against 31.19 ms on cfg2, where the thin passes of larger systems gain 1-2 %)
        constexpr bool CC = GPT || (NB == 2 && KSU_T > 1);
        for (int nb = 0; nb < NB; ++nb)
            nbv[nb] = (CC ? 8 * NB * warp : 8 * (warp + N_COMPUTE_WARPS * nb)) < nout;
        bool gpt_free_waited = false;    // ring-less instantiation: thread 0 is the pusher
        // epilogue of one pass: new rows into the state (in place) + this warp's closure partials
        auto epilogue = [&](const PassDesc& pd, const double (&cre)[MC][NB][2], const double (&cim)[MC][NB][2]) {
            // both m-tiles side by side (independent dependency chains: the FP64 latency is exposed here)
            double pr[MC], pi[MC];
            int row[MC];
            bool wr[MC];
#pragma unroll
            for (int mc = 0; mc < MC; ++mc) {
                const bool av = g < pd.nvalid[mc];
                row[mc] = pd.row0[mc] + (av ? g : 0);
                const aceqd_traj& t = trj[row[mc] - divT(row[mc]) * T];
                wr[mc] = av && (full_act || (t.n_steps >= 0 && n >= t.step0 && n < t.step0 + t.n_steps));
                pr[mc] = pi[mc] = 0.0;
            }
            if constexpr (GPT || (NB == 2 && KSU_T > 1)) {
                // accumulator (nb, e) of this lane is column cb + NB e + nb: 2 NB consecutive columns per lane
                const int cb = 8 * NB * warp + 2 * NB * tq;
                if (nbv[0]) {
#pragma unroll
                    for (int mc = 0; mc < MC; ++mc) {
                        if (pd.nvalid[mc] <= 0) continue;   // warp-uniform
                        double vr[2 * NB], vi[2 * NB];
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) {
                                vr[e * NB + nb] = cre[mc][nb][e];
                                vi[e * NB + nb] = cim[mc][nb][e];
                            }
#pragma unroll
                        for (int c = 0; c < 2 * NB; ++c) {
                            const double2 q = qbuf[cb + c];
                            pr[mc] += vr[c] * q.x - vi[c] * q.y;
                            pi[mc] += vr[c] * q.y + vi[c] * q.x;
                        }
                        if (wr[mc]) {
                            const size_t o = rowoff_r(row[mc]) + cb;
#pragma unroll
                            for (int c = 0; c < 2 * NB; c += 2) {
                                *reinterpret_cast<double2*>(Xre + o + c) = make_double2(vr[c], vr[c + 1]);
                                *reinterpret_cast<double2*>(Xim + o + c) = make_double2(vi[c], vi[c + 1]);
                            }
                        }
                    }
                }
            } else {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                if (!nbv[nb]) continue;
                const int c0 = 8 * (warp + N_COMPUTE_WARPS * nb) + 2 * tq;
                const double2 q0 = qbuf[c0], q1 = qbuf[c0 + 1];
#pragma unroll
                for (int mc = 0; mc < MC; ++mc) {
                    if (pd.nvalid[mc] <= 0) continue;   // warp-uniform
                    // closure partial of this warp's columns: r[row] += sum_col Y[row, col] q[col]
                    const double r0 = cre[mc][nb][0], r1 = cre[mc][nb][1], i0 = cim[mc][nb][0], i1 = cim[mc][nb][1];
                    pr[mc] += (r0 * q0.x - i0 * q0.y) + (r1 * q1.x - i1 * q1.y);
                    pi[mc] += (r0 * q0.y + i0 * q0.x) + (r1 * q1.y + i1 * q1.x);
                    if (wr[mc]) {
                        const size_t o = rowoff_r(row[mc]) + c0;
                        *reinterpret_cast<double2*>(Xre + o) = make_double2(r0, r1);
                        *reinterpret_cast<double2*>(Xim + o) = make_double2(i0, i1);
                    }
                }
            }
            }
#pragma unroll
            for (int mc = 0; mc < MC; ++mc) {
                pr[mc] += __shfl_xor_sync(0xffffffffu, pr[mc], 1);
                pi[mc] += __shfl_xor_sync(0xffffffffu, pi[mc], 1);
            }
#pragma unroll
            for (int mc = 0; mc < MC; ++mc) {
                if (pd.nvalid[mc] <= 0) continue;
                pr[mc] += __shfl_xor_sync(0xffffffffu, pr[mc], 2);
                pi[mc] += __shfl_xor_sync(0xffffffffu, pi[mc], 2);
                if (wr[mc] && tq == 0) rpart[warp * R + row[mc]] = make_double2(pr[mc], pi[mc]);
            }
            // the rows written above are read by the bulk-copy engine (async proxy) when the pusher warp sends them to
            // the peers: proxy fence by every writer, then one count per warp
            if (C > 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) ctr_add_release(ctr_rows);
                if (GPT && tid == 0) push_rows(pd, gpt_free_waited);
            }
        };
        for (int ps = 0; ps < p.n_pass; ++ps) {
            const PassDesc pd = passes[ps];
            if (pd.owner != (int)crank) continue;
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
            bool aval[MC], mcv[MC];
#pragma unroll
            for (int mc = 0; mc < MC; ++mc) {
                aval[mc] = g < pd.nvalid[mc];
                mcv[mc] = pd.nvalid[mc] > 0;
                const size_t o = rowoff_r(pd.row0[mc] + (aval[mc] ? g : 0)) + tq;
                are[mc] = Xre + o;
                aim[mc] = Xim + o;
            }
            // warp-uniform dispatch to a main loop without predicated-off DMMAs (a nullified
            // DMMA.8x8x4 still occupies the tensor pipe: profiles/r01f_cfg3_step_kernel_ncu.txt)
            const int mcn = mcv[MC - 1] ? MC : 1;
            bool allnb = true, anynb = false;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                allnb &= nbv[nb];
                anynb |= nbv[nb];
            }
            TICK(3);
            if constexpr (GPT) {
                // no chunk ring (the bond states fill shared memory): B fragments come from global memory / L2
                const double* blk = p.pt.blob + p.pt.off[s] + (size_t)pd.blk * nch * p.pt.chunk_doubles;
                if (!anynb) {
                } else if (allnb && mcn == MC)
                    gemm_pass_global<NB, MC, true>(cre, cim, are, aim, aval, nbv, blk, p.pt.chunk_doubles, strideB, nch, warp, g, tq);
                else if (allnb)
                    gemm_pass_global<NB, 1, true>(cre, cim, are, aim, aval, nbv, blk, p.pt.chunk_doubles, strideB, nch, warp, g, tq);
                else if (mcn == MC)
                    gemm_pass_global<NB, MC, false>(cre, cim, are, aim, aval, nbv, blk, p.pt.chunk_doubles, strideB, nch, warp, g, tq);
                else
                    gemm_pass_global<NB, 1, false>(cre, cim, are, aim, aval, nbv, blk, p.pt.chunk_doubles, strideB, nch, warp, g, tq);
            } else {
            if (!anynb) {  // this warp owns no bond column of the slice: keep the pipeline moving only
                for (int jc = 0; jc < nch; ++jc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            } else if (allnb && mcn == MC)
                gemm_pass<NB, MC, true, (NB == 2 && KSU_T > 1)>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                        warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
            else if (allnb)
                gemm_pass<NB, 1, true, (NB == 2 && KSU_T > 1)>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                       warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
            else if (mcn == MC)
                gemm_pass<NB, MC, false, (NB == 2 && KSU_T > 1)>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                         warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
            else
                gemm_pass<NB, 1, false, (NB == 2 && KSU_T > 1)>(cre, cim, are, aim, aval, nbv, chunks, p.pt.chunk_doubles, strideB, nch,
                                        warp, g, tq, bar_full, bar_empty, stage, phase, stages, lane);
            }
            TICK(7);
            compute_bar();  // every warp has finished reading this pass's X rows
            epilogue(pd, cre, cim);
        }
        TICK(4);
        compute_bar();
        TICK(5);
        // closure of the rows computed here: sum the per-warp partials; peers get a copy -- once they have left the
        // outputs of this row behind (they read the closures there): the pusher has seen their "rows read" signals
        ++step_ctr;
        if (GPT && C > 1 && tid == 0) push_step_end(gpt_free_waited);
        if (C > 1) ctr_wait_ge(ctr_free, step_ctr);
        for (int row = tid; row < R; row += N_COMPUTE_WARPS * 32) {
            if (C > 1 && !own_pos[divT(row)]) continue;
            double2 r = rpart[row];
#pragma unroll
            for (int w8 = 1; w8 < N_COMPUTE_WARPS; ++w8) {
                const double2 r2 = rpart[w8 * R + row];
                r.x += r2.x;
                r.y += r2.y;
            }
            rall[row] = r;
            for (uint32_t peer = 0; C > 1 && peer < (uint32_t)C; ++peer)   // counted by the peer's exchange barrier
                if (peer != crank) st_async_v2(mapa(smem_u32(rall + row), peer), r.x, r.y, mapa(bar_y, peer));
        }
        compute_bar();
        TICK(6);
    }
    if (!final_seg && sg.save_slot >= 0) {
        // unfinished tile: bond state, closures of row n_hi and snapshot cursors go to HBM for the CTA that resumes it
        double2* dst = reinterpret_cast<double2*>(p.seg_state + (size_t)sg.save_slot * p.seg_slot_doubles);
        const double2* src = reinterpret_cast<const double2*>(Xre);
        for (size_t e = tid; e < L.plane; e += N_COMPUTE_WARPS * 32) dst[e] = src[e];
        for (int e = tid; e < R; e += N_COMPUTE_WARPS * 32) dst[L.plane + e] = rall[e];
        int* sn = reinterpret_cast<int*>(dst + L.plane + R);
        for (int e = tid; e < T; e += N_COMPUTE_WARPS * 32) sn[e] = snapn[e];
        __threadfence();
        compute_bar();
        if (tid == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.seg_flags + sg.save_slot), "r"(p.seg_epoch) : "memory");
    }
#undef TICK
    }   // segments
    if (C > 1) cluster_sync_all();   // no CTA of a cluster exits while a peer may still address it
}

// -------------------------------------------------------------------------------------------
// Plain-FMA check kernel (SURVEY 7.1 step 5 "v0"): one CTA per trajectory, state in global
// memory, natural alpha order, no tensor cores, no pipeline.  Independent of the DMMA kernel's
// tiling; used by the GPU parity tests to localise faults.  Same inputs, same outputs.
__global__ void __launch_bounds__(256) k_step_check(const StepParams p, double* scratch) {
    const int b = blockIdx.x;
    const aceqd_traj t = p.trajs[b];
    const int NL = p.prob.NL, chi_pad = p.pt.chi_pad, n_out = p.prob.n_out;
    const int strideB = p.pt.strideB;
    double2* Y = reinterpret_cast<double2*>(scratch) + (size_t)b * 2 * NL * chi_pad;
    double2* X = Y + (size_t)NL * chi_pad;
    __shared__ double2 r[MAX_NL];
    const int tid = threadIdx.x;
    const int tot = NL * chi_pad;
    if (t.init_kind == 0) {
        const double2* r0 = reinterpret_cast<const double2*>(p.rho0s) + (size_t)t.init_index * NL;
        for (int e = tid; e < tot; e += blockDim.x) {
            const int a = e / chi_pad, d = e - a * chi_pad;
            Y[e] = d == 0 ? r0[a] : make_double2(0.0, 0.0);
        }
    } else {
        const double2* sn = reinterpret_cast<const double2*>(p.snaps) + (size_t)t.init_index * tot;
        for (int e = tid; e < tot; e += blockDim.x) Y[e] = sn[e];
    }
    __syncthreads();
    int snapn = 0;
    for (int i = 0; i <= t.n_steps; ++i) {
        const int n = t.step0 + i;
        const long long e_ = entry_of(t, i, p.ovr_base);
        for (int a = tid; a < NL; a += blockDim.x) {
            double2 acc = make_double2(0.0, 0.0);
            if (i == 0 && t.init_kind == 0) {
                acc = Y[(size_t)a * chi_pad];
            } else {
                const double2* q = reinterpret_cast<const double2*>(p.pt.closure) +
                                   (size_t)slice_of(p.pt, n - 1) * chi_pad;
                for (int d = 0; d < chi_pad; ++d) {
                    const double2 y = Y[(size_t)a * chi_pad + d];
                    acc.x += y.x * q[d].x - y.y * q[d].y;
                    acc.y += y.x * q[d].y + y.y * q[d].x;
                }
            }
            r[a] = acc;
        }
        __syncthreads();
        for (int o = tid; o < n_out && i >= t.out_from; o += blockDim.x) {
            const double2* ov = reinterpret_cast<const double2*>(p.OV + (size_t)e_ * p.prob.ov_doubles) +
                                (size_t)o * NL;
            double2 acc = make_double2(0.0, 0.0);
            for (int a = 0; a < NL; ++a) {
                acc.x += ov[a].x * r[a].x - ov[a].y * r[a].y;
                acc.y += ov[a].x * r[a].y + ov[a].y * r[a].x;
            }
            reinterpret_cast<double2*>(p.out)[t.out_off + (long long)(i - t.out_from) * n_out + o] = acc;
        }
        if (snapn < t.snap_cnt && p.snap_steps[t.snap_off + snapn] == i) {
            double2* dst = reinterpret_cast<double2*>(p.snaps) + (size_t)(t.snap_slot0 + snapn) * tot;
            for (int e = tid; e < tot; e += blockDim.x) dst[e] = Y[e];
            if (p.snap_r)
                for (int a = tid; a < NL; a += blockDim.x)
                    reinterpret_cast<double2*>(p.snap_r)[(size_t)(t.snap_slot0 + snapn) * NL + a] = r[a];
            ++snapn;
        }
        if (i == t.n_steps) break;
        const double2* W = reinterpret_cast<const double2*>(p.W + (size_t)e_ * p.prob.w_doubles);
        for (int e = tid; e < tot; e += blockDim.x) {
            const int a = e / chi_pad, d = e - a * chi_pad;
            double2 acc = make_double2(0.0, 0.0);
            for (int k = 0; k < NL; ++k) {
                const double2 w = W[(size_t)a * p.prob.NLp4 + k];
                const double2 y = Y[(size_t)k * chi_pad + d];
                acc.x += w.x * y.x - w.y * y.y;
                acc.y += w.x * y.y + w.y * y.x;
            }
            X[e] = acc;
        }
        __syncthreads();
        const int s = slice_of(p.pt, n);
        const int nch = p.pt.kin_pad[s] / KC;
        for (int e = tid; e < tot; e += blockDim.x) {
            const int a = e / chi_pad, d2 = e - a * chi_pad;
            const double* blk = p.pt.blob + p.pt.off[s] +
                                (size_t)p.prob.block_of_alpha[a] * nch * p.pt.chunk_doubles;
            double2 acc = make_double2(0.0, 0.0);
            if (d2 < p.pt.nout_pad[s]) {
                for (int d1 = 0; d1 < p.pt.kin_pad[s]; ++d1) {
                    const double* ch = blk + (size_t)(d1 / KC) * p.pt.chunk_doubles;
                    const double br = ch[(d1 % KC) * strideB + d2];
                    const double bi = ch[KC * strideB + (d1 % KC) * strideB + d2];
                    const double2 x = X[(size_t)a * chi_pad + d1];
                    acc.x += x.x * br - x.y * bi;
                    acc.y += x.x * bi + x.y * br;
                }
                Y[e] = acc;
            }
        }
        __syncthreads();
    }
}

}  // namespace

size_t step_smem_bytes(int NL, int chi_pad, int T, int stages, int wov_doubles, int wbufs) {
    return make_layout(NL, chi_pad, T, stages, wov_doubles, wbufs).total;
}

size_t step_seg_slot_doubles(int NL, int chi_pad, int T) {
    const SmemLayout L = make_layout(NL, chi_pad, T, 2, 0, 1);
    return 2 * L.plane + 2 * (size_t)T * NL + 8;   // planes, closures, <= 16 snapshot cursors
}

// query != nullptr: do not launch, report how many clusters of this configuration can be resident at once
template <int NB, bool GPT>
static int launch_nb(const StepParams& p, size_t smem_bytes, cudaStream_t s, int* query = nullptr) {
    const int ksu = p.prob.NLp4 / 4;
#define ACEQD_LAUNCH(KS)                                                                        \
    do {                                                                                        \
        ACEQD_CUDA(cudaFuncSetAttribute(k_step_dmma<NB, KS, GPT>,                               \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                        (int)smem_bytes));                                      \
        if (p.cluster > 8) /* 16 CTAs: one whole GPC's worth, beyond the portable limit */      \
            ACEQD_CUDA(cudaFuncSetAttribute(k_step_dmma<NB, KS, GPT>,                           \
                                            cudaFuncAttributeNonPortableClusterSizeAllowed, 1));\
        cudaLaunchConfig_t cfg = {};                                                            \
        cfg.gridDim = dim3((unsigned)(p.segs ? p.n_ctas : p.n_tiles * p.cluster), 1, 1);       \
        cfg.blockDim = dim3(GPT ? N_COMPUTE_WARPS * 32 : STEP_THREADS, 1, 1);                   \
        cfg.dynamicSmemBytes = smem_bytes;                                                      \
        cfg.stream = s;                                                                         \
        cudaLaunchAttribute attr[1];                                                            \
        cfg.attrs = attr;                                                                       \
        cfg.numAttrs = 0;                                                                       \
        if (p.cluster > 1) {                                                                    \
            attr[0].id = cudaLaunchAttributeClusterDimension;                                   \
            attr[0].val.clusterDim.x = (unsigned)p.cluster;                                     \
            attr[0].val.clusterDim.y = 1;                                                       \
            attr[0].val.clusterDim.z = 1;                                                       \
            cfg.numAttrs = 1;                                                                   \
        } else if (p.segs) {                                                                    \
            /* CTAs of a segment schedule wait for each other (bond-state hand-over): a         \
             * cooperative launch guarantees that all of them are resident at once */           \
            attr[0].id = cudaLaunchAttributeCooperative;                                        \
            attr[0].val.cooperative = 1;                                                        \
            cfg.numAttrs = 1;                                                                   \
        }                                                                                       \
        if (query) {                                                                            \
            *query = 0;                                                                         \
            if (cudaOccupancyMaxActiveClusters(query, k_step_dmma<NB, KS, GPT>, &cfg) !=        \
                cudaSuccess) {                                                                  \
                cudaGetLastError();                                                             \
                *query = 0;                                                                     \
            }                                                                                   \
            return ACEQD_OK;                                                                    \
        }                                                                                       \
        ACEQD_CUDA(cudaLaunchKernelEx(&cfg, k_step_dmma<NB, KS, GPT>, p));                      \
    } while (0)
    if (ksu <= 1) ACEQD_LAUNCH(1);
    else if (ksu <= 4) ACEQD_LAUNCH(4);
    else if (ksu <= 9) ACEQD_LAUNCH(9);
    else ACEQD_LAUNCH(16);
#undef ACEQD_LAUNCH
    return ACEQD_OK;
}

// clusters of p.cluster CTAs of this configuration that fit the device at once (0: such a cluster cannot be scheduled)
int step_max_active_clusters(const StepParams& p, size_t smem_bytes) {
    const int chi = p.pt.chi_pad;
    int ncl = 0, rc;
    if (p.stages == 0) {
        if (chi <= 64) rc = launch_nb<1, true>(p, smem_bytes, nullptr, &ncl);
        else if (chi <= 128) rc = launch_nb<2, true>(p, smem_bytes, nullptr, &ncl);
        else rc = launch_nb<4, true>(p, smem_bytes, nullptr, &ncl);
    } else if (chi <= 64) rc = launch_nb<1, false>(p, smem_bytes, nullptr, &ncl);
    else if (chi <= 128) rc = launch_nb<2, false>(p, smem_bytes, nullptr, &ncl);
    else rc = launch_nb<4, false>(p, smem_bytes, nullptr, &ncl);
    return rc ? 0 : ncl;
}

int launch_step_dmma(const StepParams& p, size_t smem_bytes, cudaStream_t s, LaunchLog* log) {
    const int chi = p.pt.chi_pad;
    if (chi > 256) {
        set_error("chi_pad=%d exceeds the step kernel's 256 limit", chi);
        return ACEQD_ERR_CAPACITY;
    }
    if (p.n_tiles <= 0) return ACEQD_OK;
    int rc;
    if (p.stages == 0) {
        if (chi <= 64) rc = launch_nb<1, true>(p, smem_bytes, s);
        else if (chi <= 128) rc = launch_nb<2, true>(p, smem_bytes, s);
        else rc = launch_nb<4, true>(p, smem_bytes, s);
    } else if (chi <= 64) rc = launch_nb<1, false>(p, smem_bytes, s);
    else if (chi <= 128) rc = launch_nb<2, false>(p, smem_bytes, s);
    else rc = launch_nb<4, false>(p, smem_bytes, s);
    if (rc) return rc;
    ++log->count;
    {
        const int ksu = p.prob.NLp4 / 4;
        log_name(log->step, "k_step_dmma<%d,%d> T=%d cluster=%d segments=%d%s", chi <= 64 ? 1 : (chi <= 128 ? 2 : 4),
                 ksu <= 1 ? 1 : (ksu <= 4 ? 4 : (ksu <= 9 ? 9 : 16)), p.T, p.cluster, p.segs ? 1 : 0,
                 p.stages == 0 ? " pt=global" : "");
    }
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

int launch_step_check(const StepParams& p, double* scratch, cudaStream_t s, LaunchLog* log) {
    // n_tiles carries the trajectory count for this kernel
    if (p.n_tiles <= 0) return ACEQD_OK;
    k_step_check<<<p.n_tiles, 256, 0, s>>>(p, scratch);
    ++log->count;
    log_name(log->step, "k_step_check");
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

}  // namespace aceqd
