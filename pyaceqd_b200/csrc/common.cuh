// Shared declarations of libaceqd (sm_100a only).  See include/aceqd.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aceqd.h"

namespace aceqd {

constexpr int KC = 8;             // k rows of one PT chunk (two DMMA k-steps)
constexpr int MC = 2;             // m-tiles (8 rows each) accumulated per pass
constexpr int N_COMPUTE_WARPS = 8;
// A thread that issues a cp.async.bulk right after mbarrier traffic pays ~450 cycles before its next copy can go
// out (scripts/micro/l2_ingest.cu: one chunk per 257 ns whatever its size or the ring depth), while copies of
// DIFFERENT warps overlap (scripts/micro/bulk_latency.cu: 67 B/clk per SM with 64+ KB in flight).  So the PT chunk
// ring is fed by several producer warps (chunk c belongs to producer c mod P) and the per-row operators by their own.
constexpr int N_CHUNK_PRODUCERS = 1;   // measured: more producers do not speed up the tile kernel (profiles/r05c_*)
constexpr int N_PRODUCER_WARPS = N_CHUNK_PRODUCERS + 2;   // + the W|OV stager + the row pusher of cluster launches
constexpr int STEP_THREADS = (N_COMPUTE_WARPS + N_PRODUCER_WARPS) * 32;
constexpr int MAX_NL = 64;
constexpr int MAX_PASSES = 96;
constexpr int MAX_TILE_T = 16;
constexpr int MAX_STAGES = 4;
constexpr int SMEM_BUDGET = 227 * 1024;

void set_error(const char* fmt, ...);

// What the context has launched: a counter (bench: gpu_launches) and the names of the last step kernel /
// operator builder instantiation (tests and smoke assert that the kernel they name is the one that ran).
struct LaunchLog {
    long long count = 0;
    char step[128] = "";
    char opbuild[128] = "";
    char other[128] = "";
};
void log_name(char (&dst)[128], const char* fmt, ...);

#define ACEQD_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            aceqd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                             __FILE__, __LINE__);                                     \
            return ACEQD_ERR_CUDA;                                                    \
        }                                                                             \
    } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------- device-side PT description
struct PtDev {
    int n_cls, n_slices, n_initial, n_repeat;
    int chi_pad;        // max padded bond dimension (multiple of 8)
    int strideB;        // doubles per k-row of a chunk plane: chi_pad + 4  (== 4 mod 8)
    int chunk_doubles;  // 2 * KC * strideB  (re plane then im plane)
    int pad_;
    const int* kin_pad;      // [n_slices] padded input bond (multiple of KC)
    const int* nout_pad;     // [n_slices] padded output bond (multiple of 8)
    const long long* off;    // [n_slices] offset (doubles) of chunk (cls 0, j 0)
    const double* blob;      // chunks ordered [slice][cls][chunk]
    const double* closure;   // [n_slices][2*chi_pad] interleaved complex, zero padded
    // the same slices cut into panels of PANEL output columns (split-K cluster kernel, chi_pad > PANEL only):
    // chunks ordered [slice][cls][panel][chunk], row stride PANEL + 4
    const double* pblob;
    const long long* poff;   // [n_slices]
    int n_panels;            // panels of the widest slice (1 when pblob is null)
    int pad2_;
};
constexpr int PANEL = 128;        // output columns of one split-K GEMM pass (8 warps x 2 n-tiles)

// ---------------------------------------------------------------- device-side problem
struct ProbDev {
    int NL, NLp8, NLp4, n_out, n_fields;
    int w_doubles;   // doubles per W entry: 2*NLp8*NLp4
    int ov_doubles;  // doubles per OV entry: 2*n_out*NL
    int pad_;
    const double* L0;        // [NL][NL] complex
    const double* LA;        // [n_fields][NL][NL]
    const double* LB;
    const int* field_table;  // [n_fields]
    const double* out_w;     // [n_out][NL]
    const int* pos_of_alpha; // [NL]
    const int* block_of_alpha;
};

// One GEMM pass of the step kernel: up to MC m-tiles of rows that share one PT block.
struct PassDesc {
    int blk;
    int row0[MC];
    int nvalid[MC];  // 0 = absent
    int owner;       // cluster rank that computes this pass (0 without clusters)
};

// One piece of work of a persistent CTA: steps [n_lo, n_hi) of a tile (clipped to the tile's own range).
// A tile cut in two is started by one CTA, which saves the bond state to slot `save_slot` of the segment
// workspace and raises the slot's flag, and finished by another one, which waits on `load_slot`.
struct SegDesc {
    int tile, n_lo, n_hi;
    int save_slot, load_slot;   // -1: none
};

struct StepParams {
    PtDev pt;
    ProbDev prob;
    int T;           // trajectories per tile
    int n_pass;
    int stages;      // chunk pipeline depth
    int wov_doubles; // per-trajectory smem staging of W|OV (0: read operators from global)
    int wbufs;       // staging buffers: 2 = double buffered, 1 = refilled during phase C
    int n_tiles;
    int cluster;     // CTAs per tile (1, 2 or 4): the tile's GEMM passes are split over a thread-block cluster
    const PassDesc* passes;   // [n_pass]
    const aceqd_traj* trajs;
    const int* tile_traj;     // [n_tiles][T]
    const double* W;          // entry pool
    const double* OV;
    long long ovr_base;       // entry index of explicit entry 0 in the pools
    const double* rho0s;      // [n_rho0][NL] complex
    const int* snap_steps;
    double* snaps;            // [slots][NL][chi_pad] complex
    double* out;
    // segment schedule (null: CTA i runs tile i / cluster as a whole)
    const SegDesc* segs;      // segments of CTA i: segs[seg_off[i] .. seg_off[i+1])
    const int* seg_off;
    double* seg_state;        // [slot][seg_slot_doubles]: state planes, closures, snapshot cursors
    size_t seg_slot_doubles;
    unsigned* seg_flags;      // [slot] == seg_epoch once the slot has been written in this launch
    unsigned seg_epoch;
    int n_ctas;               // grid size with a segment schedule
    long long* ticks;         // debug: [8] phase cycle counters of CTA 0 (null = off)
    double* snap_r;           // [slots][NL] complex: closure rho of the snapshot row (written with every snapshot)
    // split-K cluster kernel (splitk_kernel.cu): `cluster` CTAs hold NR bond columns each of T trajectories
    int NR;                   // bond columns per CTA (multiple of 8)
    int rslots;               // receive-ring depth for the partial products (2)
};

// ---------------------------------------------------------------- operator builder
struct OpBuildParams {
    ProbDev prob;
    double dt, t0, eval_off1, eval_off2;
    int n_sets, n_tables, n_samples, n_seq;
    double tab_t0, tab_dt;
    const double* tables;
    const aceqd_seq* seqs;
    const long long* seq_base;   // [n_seq+1] prefix sums of len
    long long n_seq_entries;
    int n_entries;               // explicit entries
    int n_mto_mats;
    const aceqd_entry* entries;
    const double* mto_mats;
    double* W;
    double* OV;
    double* scratch;             // global workspace for NL too large for shared memory (else null)
    int scratch_ctas;            // CTAs the workspace was sized for
    long long e_begin, e_end;    // entries built by this launch (0, 0 = all)
};

// bytes of global workspace the operator builder / expm kernel need for this NL (0: shared memory suffices)
size_t opbuild_scratch_bytes(int NL, int* ctas);
int launch_opbuild(const OpBuildParams& p, cudaStream_t s, LaunchLog* log);
int launch_expm_batch(int n, int count, const double* a_dev, double* out_dev, double* scratch,
                      cudaStream_t s, LaunchLog* log);
int launch_step_dmma(const StepParams& p, size_t smem_bytes, cudaStream_t s, LaunchLog* log);
int step_max_active_clusters(const StepParams& p, size_t smem_bytes);
int launch_step_check(const StepParams& p, double* scratch, cudaStream_t s, LaunchLog* log);
size_t step_smem_bytes(int NL, int chi_pad, int T, int stages, int wov_doubles, int wbufs);
size_t step_seg_slot_doubles(int NL, int chi_pad, int T);
// split-K cluster kernel: shared-memory bytes for (NL, chi_pad, n_out) with G trajectories on a cluster of C CTAs and
// `stages` PT chunk stages (0 if the combination is not supported)
size_t splitk_smem_bytes(int NL, int chi_pad, int G, int C, int stages);
int splitk_columns(int chi_pad, int C);   // NR
int launch_step_splitk(const StepParams& p, size_t smem_bytes, cudaStream_t s, LaunchLog* log);
// small-bond kernel (small_kernel.cu): one warp per 8 trajectories, process tensor resident in shared memory
size_t small_smem_bytes(long long pt_doubles, int n_slices, int chi_pad, int n_out, int warps);
int launch_step_small(const StepParams& p, long long pt_doubles, int warps_per_cta, size_t smem_bytes, cudaStream_t s,
                      LaunchLog* log);
int launch_tlmap(int NL, int n_chains, int n_w, int n_emit_max, const double* pool, const double* v0,
                 const long long* seg_off, const aceqd_tlseg* segs, const double* w, double* out,
                 double* final_v, cudaStream_t s, LaunchLog* log);
// fused tail reduction of a batch's outputs (workflow-level fusion): result[traj][pair] = spacing * trapezoid
int launch_tail_reduce(const aceqd_traj* trajs, int n_traj, int n_out, const double* out, int n_reduce,
                       const int* reduce_ch, double spacing, double* result, cudaStream_t s, LaunchLog* log);
int launch_fp64_peak(int kind, int iters, double* sink_dev, int* blocks, int* threads,
                     cudaStream_t s, LaunchLog* log);

}  // namespace aceqd
