// Batched time-local dynamical-map chains.
//
// Replaces the reference's two Fortran/OpenMP/BLAS helper modules,
//   pyaceqd/two_time/propagate_tau.f90   (propagate_tau :3-19, calc_onetime* :43-295,
//                                         calc_twotime_phonon_block :374-536)
//   pyaceqd/timebin/timebin_tl.f90       (fast_propagate :23-47, propagate_tb :50-77,
//                                         four_time :145-214, four_time_8op :216-303, dynamics_* :305-397)
// which all do the same thing: push Liouville vectors through long chains of NL x NL matrices
// (zgemv), with operator insertions (also NL x NL matrices) and traces (dot products) in between.
// Here a chain is a PROGRAM of segments over one matrix pool: segment (start, count, stride) applies
// pool[start], pool[start+stride], ... (count matrices; stride 0 repeats one matrix); segments flagged `emit` write the output
// functionals after each of their steps.  One warp owns one chain: the vector lives in shared
// memory, lane a owns rows a (and a+32), matrices stream from L2 (neighbouring chains of a (t,tau)
// grid walk the same matrices a few steps apart, so they hit L1/L2).  Bound: L2 bandwidth /
// load latency (16*NL^2 bytes and 8*NL^2 flops per step: 0.5 flop/B).
#include "common.cuh"

namespace aceqd {

namespace {

constexpr int TL_MAX_WARPS = 4;   // chains per CTA
constexpr int TL_MAX_NL = 64;
constexpr int TL_MAX_W = 64;

struct TlParams {
    int NL, n_chains, n_w, n_emit_max;
    int warps, nbuf;           // chains per CTA; matrix buffers per chain in shared memory (2: next matrix prefetched)
    const double2* pool;       // [n_mats][NL][NL]
    const double2* v0;         // [n_chains][NL]
    const long long* seg_off;  // [n_chains+1]
    const aceqd_tlseg* segs;
    const double2* w;          // [n_w][NL]
    double2* out;              // [n_chains][n_emit_max][n_w]
    double2* final_v;          // [n_chains][NL] or null
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Position in a chain program: segment s, step k of it.
struct TlPos {
    long long s;
    int k;
    aceqd_tlseg sg;
};

// The matrix of a step is copied into the warp's shared memory with coalesced 16-byte cp.async (lane e mod 32 takes element
// e) while the previous step computes; the product then reads row `lane` from shared memory (stride NL * 16 bytes: no bank
// conflicts for odd NL, 2-way at most otherwise).  The first version read every row straight from global memory inside
// the FMA loop: 25 strided loads per step in groups of four, each group a round trip to L2 -- 8.6k cycles per step of a
// 25 x 25 product (profiles/r07j_tlmap_chains_before_ncu.txt: 3.6 % of the issue slots used).
__global__ void __launch_bounds__(TL_MAX_WARPS * 32) k_tlmap_chains(const TlParams p) {
    extern __shared__ __align__(16) unsigned char tl_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NL = p.NL, n_w = p.n_w, N2 = NL * NL;
    double2* ws = reinterpret_cast<double2*>(tl_smem);                           // [TL_MAX_W * TL_MAX_NL / 4] functionals
    double2* vs = ws + TL_MAX_W * TL_MAX_NL / 4;                                  // [warps][TL_MAX_NL] vectors
    double2* mb = vs + (size_t)p.warps * TL_MAX_NL + (size_t)warp * p.nbuf * N2;  // [nbuf][NL][NL] matrices of this warp
    const bool w_sm = n_w * NL <= TL_MAX_W * TL_MAX_NL / 4;
    if (w_sm)
        for (int e = threadIdx.x; e < n_w * NL; e += blockDim.x) ws[e] = p.w[e];
    __syncthreads();
    const int chain = blockIdx.x * p.warps + warp;
    if (chain >= p.n_chains) return;
    double2* v = vs + (size_t)warp * TL_MAX_NL;
    for (int a = lane; a < NL; a += 32) v[a] = p.v0[(size_t)chain * NL + a];
    __syncwarp();
    const int r0 = lane, r1 = lane + 32;
    const bool has0 = r0 < NL, has1 = r1 < NL;
    double2* out = p.out ? p.out + (size_t)chain * p.n_emit_max * n_w : nullptr;
    int emitted = 0;
    const long long s_end = p.seg_off[chain + 1];
    auto settle = [&](TlPos& q) {         // skip empty segments; false at the end of the program
        while (q.s < s_end) {
            q.sg = p.segs[q.s];
            if (q.k < q.sg.count) return true;
            ++q.s;
            q.k = 0;
        }
        return false;
    };
    auto mat_of = [&](const TlPos& q) { return p.pool + ((size_t)q.sg.start + (size_t)q.k * q.sg.stride) * N2; };
    auto fetch = [&](const double2* m, double2* dst) {
        for (int e = lane; e < N2; e += 32) cp_async16(dst + e, m + e);
        cp_async_commit();
    };
    TlPos cur;
    cur.s = p.seg_off[chain];
    cur.k = 0;
    bool live = settle(cur);
    int b = 0;
    const double2* m_cur = nullptr;
    if (live) {
        m_cur = mat_of(cur);
        fetch(m_cur, mb);
    }
    while (live) {
        TlPos nxt = cur;
        ++nxt.k;
        const bool more = settle(nxt);
        const double2* m_nxt = more ? mat_of(nxt) : nullptr;
        const bool same = more && m_nxt == m_cur;           // a repeated matrix (stride 0) stays where it is
        bool prefetched = false;
        if (more && !same && p.nbuf == 2) {
            fetch(m_nxt, mb + (size_t)(b ^ 1) * N2);
            prefetched = true;
        }
        if (prefetched) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const double2* mm = mb + (size_t)b * N2;
        double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
        if (has0) {
            const double2* row = mm + (size_t)r0 * NL;
            double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);     // two partial sums: shorter FMA chains
            int q = 0;
            for (; q + 1 < NL; q += 2) {
                const double2 e = row[q], x = v[q], f = row[q + 1], y = v[q + 1];
                c0.x = fma(e.x, x.x, c0.x); c0.x = fma(-e.y, x.y, c0.x);
                c0.y = fma(e.x, x.y, c0.y); c0.y = fma(e.y, x.x, c0.y);
                c1.x = fma(f.x, y.x, c1.x); c1.x = fma(-f.y, y.y, c1.x);
                c1.y = fma(f.x, y.y, c1.y); c1.y = fma(f.y, y.x, c1.y);
            }
            if (q < NL) {
                const double2 e = row[q], x = v[q];
                c0.x = fma(e.x, x.x, c0.x); c0.x = fma(-e.y, x.y, c0.x);
                c0.y = fma(e.x, x.y, c0.y); c0.y = fma(e.y, x.x, c0.y);
            }
            a0 = make_double2(c0.x + c1.x, c0.y + c1.y);
        }
        if (has1) {
            const double2* row = mm + (size_t)r1 * NL;
            double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
            int q = 0;
            for (; q + 1 < NL; q += 2) {
                const double2 e = row[q], x = v[q], f = row[q + 1], y = v[q + 1];
                c0.x = fma(e.x, x.x, c0.x); c0.x = fma(-e.y, x.y, c0.x);
                c0.y = fma(e.x, x.y, c0.y); c0.y = fma(e.y, x.x, c0.y);
                c1.x = fma(f.x, y.x, c1.x); c1.x = fma(-f.y, y.y, c1.x);
                c1.y = fma(f.x, y.y, c1.y); c1.y = fma(f.y, y.x, c1.y);
            }
            if (q < NL) {
                const double2 e = row[q], x = v[q];
                c0.x = fma(e.x, x.x, c0.x); c0.x = fma(-e.y, x.y, c0.x);
                c0.y = fma(e.x, x.y, c0.y); c0.y = fma(e.y, x.x, c0.y);
            }
            a1 = make_double2(c0.x + c1.x, c0.y + c1.y);
        }
        __syncwarp();
        if (has0) v[r0] = a0;
        if (has1) v[r1] = a1;
        __syncwarp();
        if (cur.sg.emit && out && emitted < p.n_emit_max) {
            for (int j = 0; j < n_w; ++j) {
                const double2* wj = (w_sm ? ws : p.w) + (size_t)j * NL;
                double2 acc = make_double2(0.0, 0.0);
                if (has0) {
                    const double2 c = wj[r0];
                    acc.x = c.x * a0.x - c.y * a0.y;
                    acc.y = c.x * a0.y + c.y * a0.x;
                }
                if (has1) {
                    const double2 c = wj[r1];
                    acc.x += c.x * a1.x - c.y * a1.y;
                    acc.y += c.x * a1.y + c.y * a1.x;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                }
                if (lane == 0) out[(size_t)emitted * n_w + j] = acc;
            }
            ++emitted;
        }
        // next step
        if (more && !same) {
            if (prefetched) {
                b ^= 1;
            } else {                       // single buffer: every lane has read the matrix (barriers above), refill it
                fetch(m_nxt, mb + (size_t)b * N2);
            }
            m_cur = m_nxt;
        }
        cur = nxt;
        live = more;
    }
    if (p.final_v)
        for (int a = lane; a < NL; a += 32) p.final_v[(size_t)chain * NL + a] = v[a];
}

}  // namespace

int launch_tlmap(int NL, int n_chains, int n_w, int n_emit_max, const double* pool, const double* v0,
                 const long long* seg_off, const aceqd_tlseg* segs, const double* w, double* out,
                 double* final_v, cudaStream_t s, LaunchLog* log) {
    if (NL > TL_MAX_NL) {
        set_error("tlmap: NL=%d exceeds %d", NL, TL_MAX_NL);
        return ACEQD_ERR_CAPACITY;
    }
    if (n_chains <= 0) return ACEQD_OK;
    TlParams p{};
    p.NL = NL;
    p.n_chains = n_chains;
    p.n_w = n_w;
    p.n_emit_max = n_emit_max;
    p.pool = (const double2*)pool;
    p.v0 = (const double2*)v0;
    p.seg_off = seg_off;
    p.segs = segs;
    p.w = (const double2*)w;
    p.out = (double2*)out;
    p.final_v = (double2*)final_v;
    // chains per CTA and matrix buffers per chain: the most warps, then double buffering, that leave two CTAs per SM
    const size_t fixed = (size_t)(TL_MAX_W * TL_MAX_NL / 4) * sizeof(double2);
    size_t smem = 0;
    p.warps = 0;
    for (const int budget : {SMEM_BUDGET / 2 - 1024, SMEM_BUDGET}) {
        for (const int wn : {4, 2, 1}) {
            for (const int nb : {2, 1}) {
                const size_t need = fixed + (size_t)wn * (TL_MAX_NL + (size_t)nb * NL * NL) * sizeof(double2);
                if (need <= (size_t)budget) {
                    p.warps = wn;
                    p.nbuf = nb;
                    smem = need;
                    break;
                }
            }
            if (p.warps) break;
        }
        if (p.warps) break;
    }
    if (!p.warps) {
        set_error("tlmap: NL=%d does not fit shared memory", NL);
        return ACEQD_ERR_CAPACITY;
    }
    if (smem > 48 * 1024)
        ACEQD_CUDA(cudaFuncSetAttribute(k_tlmap_chains, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tlmap_chains<<<(n_chains + p.warps - 1) / p.warps, p.warps * 32, smem, s>>>(p);
    ++log->count;
    log_name(log->other, "k_tlmap_chains");
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

// -------------------------------------------------------------------------------------------
// Fused tail reduction: the tau integral of G2(t, tau) per t, computed where the step kernel left
// its outputs.  Replaces the host loop of pol_entanglement/G2.py:507-533 (np.trapz over the last n_t2 + 1 rows of every
// run) -- only n_traj x n_pairs numbers cross PCIe instead of the whole (t, tau) map.
namespace {
__global__ void __launch_bounds__(128) k_tail_reduce(const aceqd_traj* trajs, int n_traj, int n_out, const double2* out,
                                                     int n_reduce, const int* reduce_ch, double spacing, double2* result) {
    const int b = blockIdx.x;
    if (b >= n_traj) return;
    const aceqd_traj t = trajs[b];
    const int rows = t.n_steps + 1 - t.out_from;            // kept rows: tau = 0 .. m
    const double2* o = out + t.out_off;
    __shared__ double2 part[4];
    for (int p = 0; p < n_reduce; ++p) {
        const int ch_tau = reduce_ch[2 * p], ch_zero = reduce_ch[2 * p + 1];
        double2 acc = make_double2(0.0, 0.0);
        for (int k = threadIdx.x; k < rows; k += blockDim.x) {
            const double2 v = o[(size_t)k * n_out + (k == 0 ? ch_zero : ch_tau)];
            const double w = (k == 0 || k == rows - 1) ? 0.5 : 1.0;
            acc.x += w * v.x;
            acc.y += w * v.y;
        }
        for (int sh = 16; sh > 0; sh >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sh);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sh);
        }
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double2 r = part[0];
            for (int w = 1; w < 4; ++w) {
                r.x += part[w].x;
                r.y += part[w].y;
            }
            if (rows < 2) r = make_double2(0.0, 0.0);      // a single sample spans no interval
            result[(size_t)b * n_reduce + p] = make_double2(spacing * r.x, spacing * r.y);
        }
        __syncthreads();
    }
}
}  // namespace

int launch_tail_reduce(const aceqd_traj* trajs, int n_traj, int n_out, const double* out, int n_reduce,
                       const int* reduce_ch, double spacing, double* result, cudaStream_t s, LaunchLog* log) {
    if (n_traj <= 0 || n_reduce <= 0) return ACEQD_OK;
    k_tail_reduce<<<n_traj, 128, 0, s>>>(trajs, n_traj, n_out, reinterpret_cast<const double2*>(out), n_reduce, reduce_ch,
                                         spacing, reinterpret_cast<double2*>(result));
    ++log->count;
    log_name(log->other, "k_tail_reduce");
    ACEQD_CUDA(cudaGetLastError());
    return ACEQD_OK;
}

}  // namespace aceqd
