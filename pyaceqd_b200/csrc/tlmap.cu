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

constexpr int TL_WARPS = 4;       // chains per CTA
constexpr int TL_MAX_NL = 64;
constexpr int TL_MAX_W = 64;

struct TlParams {
    int NL, n_chains, n_w, n_emit_max;
    const double2* pool;       // [n_mats][NL][NL]
    const double2* v0;         // [n_chains][NL]
    const long long* seg_off;  // [n_chains+1]
    const aceqd_tlseg* segs;
    const double2* w;          // [n_w][NL]
    double2* out;              // [n_chains][n_emit_max][n_w]
    double2* final_v;          // [n_chains][NL] or null
};

__global__ void __launch_bounds__(TL_WARPS * 32) k_tlmap_chains(const TlParams p) {
    __shared__ double2 vs[TL_WARPS][TL_MAX_NL];
    __shared__ double2 ws[TL_MAX_W * TL_MAX_NL / 4];   // up to n_w*NL <= 1024 complex
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NL = p.NL, n_w = p.n_w;
    const bool w_sm = n_w * NL <= TL_MAX_W * TL_MAX_NL / 4;
    if (w_sm)
        for (int e = threadIdx.x; e < n_w * NL; e += blockDim.x) ws[e] = p.w[e];
    __syncthreads();
    const int chain = blockIdx.x * TL_WARPS + warp;
    if (chain >= p.n_chains) return;
    double2* v = vs[warp];
    for (int a = lane; a < NL; a += 32) v[a] = p.v0[(size_t)chain * NL + a];
    __syncwarp();
    const int r0 = lane, r1 = lane + 32;
    const bool has0 = r0 < NL, has1 = r1 < NL;
    double2* out = p.out ? p.out + (size_t)chain * p.n_emit_max * n_w : nullptr;
    int emitted = 0;
    for (long long s = p.seg_off[chain]; s < p.seg_off[chain + 1]; ++s) {
        const aceqd_tlseg sg = p.segs[s];
        const double2* m = p.pool + (size_t)sg.start * NL * NL;
        for (int k = 0; k < sg.count; ++k, m += (size_t)sg.stride * NL * NL) {
            double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
            if (has0) {
                const double2* row = m + (size_t)r0 * NL;
#pragma unroll 4
                for (int b = 0; b < NL; ++b) {
                    const double2 e = __ldg(row + b), x = v[b];
                    a0.x = fma(e.x, x.x, a0.x); a0.x = fma(-e.y, x.y, a0.x);
                    a0.y = fma(e.x, x.y, a0.y); a0.y = fma(e.y, x.x, a0.y);
                }
            }
            if (has1) {
                const double2* row = m + (size_t)r1 * NL;
#pragma unroll 4
                for (int b = 0; b < NL; ++b) {
                    const double2 e = __ldg(row + b), x = v[b];
                    a1.x = fma(e.x, x.x, a1.x); a1.x = fma(-e.y, x.y, a1.x);
                    a1.y = fma(e.x, x.y, a1.y); a1.y = fma(e.y, x.x, a1.y);
                }
            }
            __syncwarp();
            if (has0) v[r0] = a0;
            if (has1) v[r1] = a1;
            __syncwarp();
            if (sg.emit && out && emitted < p.n_emit_max) {
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
        }
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
    k_tlmap_chains<<<(n_chains + TL_WARPS - 1) / TL_WARPS, TL_WARPS * 32, 0, s>>>(p);
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
