// Device-side building blocks shared by the persistent step kernel (step_kernel.cu) and the
// step-synchronous streaming kernel (stream_kernel.cu): mbarrier / bulk-copy / cluster primitives,
// the FP64 DMMA wrapper and the GEMM pass over the k-chunks of one PT block.
#pragma once
#include "common.cuh"

#ifndef ACEQD_GPT_BUFS
#define ACEQD_GPT_BUFS 2
#endif

namespace aceqd {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// ---- thread-block cluster / distributed shared memory primitives
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {  // same offset in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// One cluster-scope release fence followed by RELAXED remote arrivals: a `mbarrier.arrive.release.cluster` per peer
// pays a cluster-scope fence each time (measured: ~1.5k cycles per arrive, 16 serialised arrives per GEMM pass made
// the split-K kernel 5x slower than the tile kernel; profiles/r05e_*).  Call from lanes 0..C-1 of ONE warp after a CTA
// barrier: the fence executes once for the warp, every lane signals one peer.
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Remote store that completes on the DESTINATION CTA's mbarrier (tx-count): data and signal travel together through
// the async proxy, so neither side needs a cluster-scope fence (which costs ~2k cycles: profiles/r05i_splitk_ticks.txt).
__device__ __forceinline__ void st_async_v2(uint32_t dst_cluster, double a, double b, uint32_t mbar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(dst_cluster),
                 "d"(a), "d"(b), "r"(mbar_cluster)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_CWAIT:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_CDONE;\n"
        "bra LAB_CWAIT;\n"
        "LAB_CDONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_cluster_c128(uint32_t cluster_addr, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(cluster_addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// monotonic shared-memory counters (CTA scope): producer side adds with release, consumer side polls with acquire
__device__ __forceinline__ void ctr_add_release(uint32_t addr) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void ctr_wait_ge(uint32_t addr, uint32_t want) {
    uint32_t v;
    do {
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    } while ((int32_t)(v - want) < 0);
}
__device__ __forceinline__ void compute_bar() {  // the 8 compute warps only
    asm volatile("bar.sync 1, %0;" ::"n"(N_COMPUTE_WARPS * 32) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
// N consecutive doubles through the read-only path in one instruction (N = 4: a 256-bit load, new with sm_100)
template <int N>
__device__ __forceinline__ void ldg_vec(const double* p, double (&v)[N]) {
    static_assert(N == 1 || N == 2 || N == 4, "1, 2 or 4 doubles");
    if constexpr (N == 1) {
        v[0] = __ldg(p);
    } else if constexpr (N == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        v[0] = t.x;
        v[1] = t.y;
    } else {
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    }
}
__device__ __forceinline__ int slice_of(const PtDev& pt, int n) {
    return n < pt.n_initial ? n : pt.n_initial + (n - pt.n_initial) % pt.n_repeat;
}
__device__ __forceinline__ long long entry_of(const aceqd_traj& t, int i, long long ovr_base) {
    long long e = t.ent0 + i;
    for (int q = 0; q < t.n_ovr; ++q)
        if (t.ovr_step[q] == i) e = ovr_base + t.ovr_ent[q];
    return e;
}

// Main loop of one GEMM pass: MCV (<= MC) m-tiles x NB n-tiles of this warp over all k-chunks of
// one PT block.  ALLNB: every n-tile of the warp is inside the slice (no predicates at all).
// CC: the warp owns 8 NB consecutive bond columns (see the fragment loads) instead of the n-tiles w, w + 8, ...
template <int NB, int MCV, bool ALLNB, bool CC = false>
__device__ __forceinline__ void gemm_pass(double (&cre)[MC][NB][2], double (&cim)[MC][NB][2],
                                          const double* const (&are)[MC], const double* const (&aim)[MC],
                                          const bool (&aval)[MC], const bool (&nbv)[NB],
                                          const double* chunks, int chunk_doubles, int strideB, int nch,
                                          int warp, int g, int tq, uint32_t bar_full, uint32_t bar_empty,
                                          int& stage, uint32_t& phase, int stages, int lane) {
    static_assert(KC == 8, "the main loop is written for two DMMA k-steps per chunk");
    // Software-pipelined over k-steps: the fragments of the NEXT k-step are loaded before the DMMAs of the current
    // one are issued -- across the chunk boundary too (wait for the next stage, load its first fragments, release
    // the current stage, THEN issue the last DMMA batch of the current chunk), so a warp's DMMA stream has no gap
    // at a chunk boundary (the two warps of a sub-partition alternate on the tensor pipe in lockstep and would
    // otherwise idle it together there).
    double a_re[2][MCV], a_im[2][MCV], b_re[2][NB], b_im[2][NB];
    auto load = [&](int buf, const double* bre, int ks, int k) {
        const double* bim = bre + KC * strideB;
#pragma unroll
        for (int mc = 0; mc < MCV; ++mc) {
            a_re[buf][mc] = aval[mc] ? are[mc][k] : 0.0;
            a_im[buf][mc] = aval[mc] ? aim[mc][k] : 0.0;
        }
        if constexpr (CC && NB == 2) {
            // two n-tiles per warp: the warp owns 16 CONSECUTIVE bond columns, column 16 w + 2 g + nb is n-index g of its
            // n-tile nb, and the two fragments of a lane are one 16-byte load per plane (conflict-free: the four k-rows of
            // a quarter-warp sit 32 bytes apart modulo 128)
            const int bo = (4 * ks + tq) * strideB + 2 * (8 * warp + g);
            if (ALLNB || nbv[0]) {
                const double2 vr = *reinterpret_cast<const double2*>(bre + bo);
                const double2 vi = *reinterpret_cast<const double2*>(bim + bo);
                b_re[buf][0] = vr.x; b_re[buf][1] = vr.y;
                b_im[buf][0] = vi.x; b_im[buf][1] = vi.y;
            } else {
                b_re[buf][0] = b_re[buf][1] = b_im[buf][0] = b_im[buf][1] = 0.0;
            }
        } else {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const int bo = (4 * ks + tq) * strideB + 8 * (warp + N_COMPUTE_WARPS * nb) + g;
            b_re[buf][nb] = (ALLNB || nbv[nb]) ? bre[bo] : 0.0;
            b_im[buf][nb] = (ALLNB || nbv[nb]) ? bim[bo] : 0.0;
        }
        }
    };
    auto batch = [&](int buf) {
        // two sweeps so that consecutive DMMAs never share an accumulator
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int mc = 0; mc < MCV; ++mc)
                if (ALLNB || nbv[nb]) {
                    dmma(cre[mc][nb][0], cre[mc][nb][1], a_re[buf][mc], b_re[buf][nb]);
                    dmma(cim[mc][nb][0], cim[mc][nb][1], a_re[buf][mc], b_im[buf][nb]);
                }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int mc = 0; mc < MCV; ++mc)
                if (ALLNB || nbv[nb]) {
                    dmma(cre[mc][nb][0], cre[mc][nb][1], -a_im[buf][mc], b_im[buf][nb]);
                    dmma(cim[mc][nb][0], cim[mc][nb][1], a_im[buf][mc], b_re[buf][nb]);
                }
    };
    if (nch <= 0) return;
    mbar_wait(bar_full + 8 * stage, phase);
    load(0, chunks + (size_t)stage * chunk_doubles, 0, 0);
    for (int jc = 0; jc < nch; ++jc) {
        const double* bre = chunks + (size_t)stage * chunk_doubles;
        load(1, bre, 1, jc * KC + 4);
        batch(0);
        const int cur = stage;
        if (++stage == stages) { stage = 0; phase ^= 1u; }
        if (jc + 1 < nch) {
            mbar_wait(bar_full + 8 * stage, phase);
            load(0, chunks + (size_t)stage * chunk_doubles, 0, (jc + 1) * KC);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * cur);     // every fragment of this chunk has been read
        batch(1);
    }
}

// The same pass with the PT block read straight from global memory / L2 (no shared-memory ring, no producer): used
// when the bond states of a tile leave no room for chunk stages (NL = 25, chi = 256: one trajectory plus a ring, or
// TWO trajectories without one -- and two trajectories fill the 8-row DMMA m-tiles 1.8x better).  The B fragments of
// a whole chunk (two k-steps) travel through registers one chunk ahead of the DMMAs that use them; the chunk layout
// in HBM is the shared-memory stage layout, so the fragment addressing is unchanged.  Every B element is read by
// exactly one warp (warps own disjoint bond columns): the L2 -> SM traffic equals the ring's.
template <int NB, int MCV, bool ALLNB>
__device__ __forceinline__ void gemm_pass_global(double (&cre)[MC][NB][2], double (&cim)[MC][NB][2],
                                                 const double* const (&are)[MC], const double* const (&aim)[MC],
                                                 const bool (&aval)[MC], const bool (&nbv)[NB], const double* blk,
                                                 int chunk_doubles, int strideB, int nch, int warp, int g, int tq) {
    static_assert(KC == 8, "two DMMA k-steps per chunk");
    // chunk buffers in registers: loads run NBUF - 1 chunks ahead of their DMMAs.  Measured at NL = 25, chi = 256, T = 2
    // (gpurun_out/r7b_*): three buffers in every pass 36.3 -> 51.2 ms, three buffers in the thin (one m-tile) passes only
    // 49.4 ms -- more loads in flight do not help (and the register file is full at two), so the fragments' way from
    // L2 into the SM is bound by its throughput, not by its latency.
    constexpr int NBUF = ACEQD_GPT_BUFS;
    double b_re[NBUF][2][NB], b_im[NBUF][2][NB];
    // Warp w owns the bond columns [8 NB w, 8 NB (w + 1)); column 8 NB w + NB g + nb is n-index g of its n-tile nb, so the
    // NB fragments of a lane are NB consecutive doubles of one PT row: ONE load of 8 NB bytes (LDG.256 for chi = 256)
    // where the strided assignment of the ring kernel needs NB loads of 8 bytes that touch four cache lines each.
    auto loadB = [&](int buf, const double* bre) {
        const double* bim = bre + KC * strideB;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int bo = (4 * ks + tq) * strideB + NB * (8 * warp + g);
            if (ALLNB || nbv[0]) {
                ldg_vec<NB>(bre + bo, b_re[buf][ks]);
                ldg_vec<NB>(bim + bo, b_im[buf][ks]);
            } else {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) b_re[buf][ks][nb] = b_im[buf][ks][nb] = 0.0;
            }
        }
    };
    auto compute = [&](int buf, int jc) {
        double a_re[2][MCV], a_im[2][MCV];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int mc = 0; mc < MCV; ++mc) {
                a_re[ks][mc] = aval[mc] ? are[mc][jc * KC + 4 * ks] : 0.0;
                a_im[ks][mc] = aval[mc] ? aim[mc][jc * KC + 4 * ks] : 0.0;
            }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int mc = 0; mc < MCV; ++mc)
                    if (ALLNB || nbv[nb]) {
                        dmma(cre[mc][nb][0], cre[mc][nb][1], a_re[ks][mc], b_re[buf][ks][nb]);
                        dmma(cim[mc][nb][0], cim[mc][nb][1], a_re[ks][mc], b_im[buf][ks][nb]);
                    }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int mc = 0; mc < MCV; ++mc)
                    if (ALLNB || nbv[nb]) {
                        dmma(cre[mc][nb][0], cre[mc][nb][1], -a_im[ks][mc], b_im[buf][ks][nb]);
                        dmma(cim[mc][nb][0], cim[mc][nb][1], a_im[ks][mc], b_re[buf][ks][nb]);
                    }
        }
    };
    if (nch <= 0) return;
#pragma unroll
    for (int q = 0; q < NBUF - 1; ++q)
        if (q < nch) loadB(q, blk + (size_t)q * chunk_doubles);
    for (int jc = 0; jc < nch; jc += NBUF) {
#pragma unroll
        for (int q = 0; q < NBUF; ++q) {
            const int pre = jc + q + NBUF - 1;      // chunk fetched now, into the buffer compute(q - 1) released
            if (pre < nch) loadB((q + NBUF - 1) % NBUF, blk + (size_t)pre * chunk_doubles);
            if (jc + q < nch) compute(q, jc + q);
        }
    }
}

}  // namespace
}  // namespace aceqd
