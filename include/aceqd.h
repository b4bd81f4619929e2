/*
 * aceqd.h -- C ABI of the B200-native process-tensor propagation engine (libaceqd.so).
 *
 * The reference (tbracht/pyaceqd) has no FFI for this path: it crosses a PROCESS boundary,
 * writing an ACE parameter file + pulse files and parsing ACE's text output
 * (pyaceqd/general_system/general_system.py:227-343), or -- in calc_dynmap mode -- calls the
 * ACEutils pybind objects (general_system.py:313-336).  Each entry point below names the
 * piece of that interface it replaces.  Conventions:
 *   - return 0 on success, a negative aceqd_status otherwise; never throws;
 *     aceqd_last_error() gives the message of the calling thread's last failure;
 *   - the caller owns every buffer; handles are opaque and created/destroyed by the library;
 *   - complex numbers are interleaved double[2] (re, im); matrices are row-major;
 *   - all kernels are launched on the context's stream; nothing is global;
 *   - there is NO CPU fallback: every compute entry fails with ACEQD_ERR_CUDA without a GPU.
 */
#ifndef ACEQD_H
#define ACEQD_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ACEQD_OK = 0,
    ACEQD_ERR_ARG = -1,      /* invalid argument                                   */
    ACEQD_ERR_CUDA = -2,     /* CUDA runtime error / no device                      */
    ACEQD_ERR_NOMEM = -3,    /* device or host allocation failed                    */
    ACEQD_ERR_CAPACITY = -4  /* problem does not fit the kernel's on-chip budget    */
} aceqd_status;

typedef struct aceqd_ctx aceqd_ctx;         /* device + stream + workspace + counters           */
typedef struct aceqd_pt aceqd_pt;           /* process tensor resident in HBM (kernel layout)   */
typedef struct aceqd_problem aceqd_problem; /* Liouvillian pieces + output functionals in HBM   */

const char* aceqd_last_error(void);
const char* aceqd_version(void);

/* Context.  `stream` is a cudaStream_t (or NULL for a library-owned non-blocking stream). */
int aceqd_ctx_create(int device, void* stream, aceqd_ctx** out);
void aceqd_ctx_destroy(aceqd_ctx* ctx);
int aceqd_ctx_sync(aceqd_ctx* ctx);
/* Number of kernels of THIS library launched on the context so far (bench: gpu_launches). */
long long aceqd_launch_count(const aceqd_ctx* ctx);
/* Name of the kernel instantiation the last aceqd_run_steps / aceqd_build_operators / service call (expm, tl-map
 * chains) launched on this context, e.g. "k_step_small<2> warps=4", "k_step_dmma<2,4> T=4 cluster=2 segments=0",
 * "k_opbuild_dmma<2>".  Tests and the smoke run assert that the kernel they name is the one that ran. */
const char* aceqd_last_step_kernel(const aceqd_ctx* ctx);
const char* aceqd_last_opbuild_kernel(const aceqd_ctx* ctx);
const char* aceqd_last_other_kernel(const aceqd_ctx* ctx);
/* Device time (ms, CUDA events on the context's stream) of the last step-kernel launch and of
 * the last operator-builder launch; valid after aceqd_ctx_sync(). */
int aceqd_last_timings(aceqd_ctx* ctx, float* step_kernel_ms, float* opbuild_kernel_ms);

/*
 * Process tensor.  Replaces ACE's `add_PT <file>` (general_system.py:236) / ProcessTensors(param)
 * (general_system.py:328).  Slices are given as host arrays:
 *   slices[s]   : [n_cls][chi_in[s]][chi_out[s]] complex   (A_n[beta, d1, d2])
 *   closures[s] : [chi_out[s]] complex                      (environment closure after slice s)
 * Slices 0..n_initial-1 are used once, the remaining ones are cycled ("repeat" PTs,
 * general_system.py:174,194).  The library re-tiles them into its HBM layout.
 */
int aceqd_pt_create(aceqd_ctx* ctx, int n_cls, int n_slices, int n_initial,
                    const int* chi_in, const int* chi_out,
                    const double* const* slices, const double* const* closures,
                    aceqd_pt** out);
void aceqd_pt_destroy(aceqd_pt* pt);
int aceqd_pt_chi_pad(const aceqd_pt* pt);

/*
 * Problem.  Replaces the param-file keys initial/add_Hamiltonian/add_Pulse/add_Lindblad/
 * add_Output (general_system.py:239-289) after operator parsing:
 *   L(t) = L0 + sum_k f_k(t) LA[k] + conj(f_k(t)) LB[k]          (NL x NL each)
 *   field_table[k] : which drive table (column of `tables`, see aceqd_batch) feeds field k, or -1
 *   out_w          : [n_out][NL] output functionals,  <O_j> = out_w[j] . vec(rho)
 *   pos_of_alpha   : [NL] position of Liouville index alpha in coupling-class-sorted order
 *   block_of_alpha : [NL] PT block (beta) that multiplies row alpha            
 */
int aceqd_problem_create(aceqd_ctx* ctx, int NL, int n_fields, int n_out,
                         const double* L0, const double* LA, const double* LB,
                         const int* field_table, const double* out_w,
                         const int* pos_of_alpha, const int* block_of_alpha,
                         aceqd_problem** out);
void aceqd_problem_destroy(aceqd_problem* p);

/* One sequence of per-step operators: `len` consecutive absolute steps of one drive set. */
typedef struct {
    int32_t set;    /* drive set (row block of aceqd_batch.tables)                            */
    int32_t step0;  /* absolute step index of entry 0 (time = t0 + step0*dt)                  */
    int32_t len;    /* number of entries (steps + 1 output rows)                              */
    int32_t first_has_prev; /* 0: entry 0 starts a fresh trajectory (no preceding half step)  */
} aceqd_seq;

/* An explicit per-step operator entry (steps that carry multi-time operators). */
typedef struct {
    int32_t set;
    int32_t step;     /* absolute step index                                                  */
    int32_t sb;       /* index into mto_mats applied BEFORE the output row (applyBefore), or -1 */
    int32_t sa;       /* index into mto_mats applied AFTER the output row, or -1               */
    int32_t has_prev; /* a half step precedes this row                                         */
    int32_t clamp;    /* > 0: this row's drive ends after `clamp` samples (end value held beyond):     */
                      /* the reference writes one pulse file PER run on np.arange(t_start, t_end, dt)  */
                      /* (general_system.py:55-71,213), so the last step of a run that shares a longer */
                      /* table with other runs must not see samples past its own t_end - dt            */
} aceqd_entry;

#define ACEQD_MAX_OVR 6
#define ACEQD_SMALL_MIN_TRAJ 2048   /* kernel 0: smallest NL = 4 / chi_pad <= 32 batch that takes k_step_small */

/* One trajectory = one `system(t_start, t_end, ...)` call of the reference. */
typedef struct {
    int64_t ent0;     /* operator entry of its local step 0: seq_base[seq] + offset            */
    int64_t out_off;  /* offset (complex elements) of its output block                         */
                      /*   [n_steps+1-out_from][n_out]: rows out_from..n_steps                 */
    int32_t step0;    /* absolute step index of local step 0 (selects PT slices)               */
    int32_t n_steps;
    int32_t init_kind;  /* 0: rho0s[init_index] in bond column 0; 1: snapshot slot init_index  */
    int32_t init_index;
    int32_t n_ovr;      /* local steps whose operator entry is replaced (MTO rows)             */
    int32_t ovr_step[ACEQD_MAX_OVR];
    int32_t ovr_ent[ACEQD_MAX_OVR];  /* index into the explicit entry list                     */
    int32_t snap_off;   /* offset into snap_steps of this trajectory's snapshot requests       */
    int32_t snap_cnt;
    int32_t snap_slot0; /* snapshot j goes to slot snap_slot0 + j                              */
    int32_t out_from;   /* first local row that is written out (consumers index results from   */
                        /* the END: two_time/correlations.py:182-183), 0 = every row            */
} aceqd_traj;

/*
 * A batch of trajectories (the unit the reference fans out over a ThreadPoolExecutor, e.g.
 * two_time/correlations.py:153-170, two_level_system/rabi_rotations.py:172-198).
 * All pointers are HOST pointers unless `device_resident` is set, in which case `tables` and
 * `out` are device pointers (HBM-resident timing leg of bench.py); descriptors are always host.
 */
typedef struct {
    double dt;          /* time step                                                           */
    double t0;          /* time of absolute step 0 (the PT's origin, ACE's `ta`)               */
    double eval_off1;   /* first half step evaluates L at t_n + eval_off1*dt   (0.25)          */
    double eval_off2;   /* second half step evaluates L at t_n + eval_off2*dt  (0.75)          */
    /* drive tables: values[set][table][sample] complex, sample j at tab_t0 + j*tab_dt
     * (the content of the pulse / rf files of general_system.py:55-102)                       */
    int32_t n_sets, n_tables, n_samples;
    double tab_t0, tab_dt;
    const double* tables;
    /* operator sequences + explicit entries */
    int32_t n_seq;
    const aceqd_seq* seqs;
    int32_t n_entries;
    const aceqd_entry* entries;
    int32_t n_mto_mats;
    const double* mto_mats;   /* [n_mto_mats][NL][NL] complex superoperators                   */
    /* initial states */
    int32_t n_rho0;
    const double* rho0s;      /* [n_rho0][NL] complex                                          */
    /* trajectories, tiled: tile i owns trajectories tile_traj[i*tile_T .. +tile_T) (-1 = none) */
    int32_t n_traj;
    const aceqd_traj* trajs;
    int32_t tile_T, n_tiles;
    const int32_t* tile_traj;
    int32_t n_snap_steps;
    const int32_t* snap_steps; /* local step indices at which the bond state is snapshotted    */
    int32_t n_snap_slots;      /* snapshot pool size needed (slots of NL x chi_pad complex)    */
    /* output */
    int64_t out_elems;         /* total complex elements of `out`                              */
    double* out;
    int32_t device_resident;
    int32_t kernel;            /* 0: persistent DMMA kernels, library's choice between the tile kernel    */
                               /*    (k_step_dmma) and, for NL = 4 / chi_pad <= 32 batches of at least    */
                               /*    ACEQD_SMALL_MIN_TRAJ trajectories, the small-bond kernel k_step_small */
                               /* 1: plain-FMA check kernel (tests)                                        */
                               /* 3: split-K cluster kernel k_step_splitk: `cluster` CTAs hold             */
                               /*    chi_pad/cluster bond columns each of tile_T trajectories (large NL)   */
                               /* 4: k_step_small or ACEQD_ERR_CAPACITY;  5: k_step_dmma always            */
    int32_t cluster;           /* CTAs per tile (0/1, 2, 4, 8 or 16): a thread-block cluster shares one tile */
                               /* kernel 0/5: its GEMM passes are split, rows exchanged through            */
                               /* distributed shared memory, every CTA keeps the full bond state;          */
                               /* kernel 3: the bond columns are split (see above)                         */
    int32_t n_reduce;          /* > 0: fused tail reduction (workflow-level fusion, pol_entanglement/G2.py:507-533):   */
                               /* the outputs stay in HBM and only, per trajectory and pair p, the trapezoid over its   */
                               /* kept rows   s * (f_0/2 + f_1 + ... + f_{m-1} + f_m/2),   f_0 = out[row 0][zero_p],     */
                               /* f_k = out[row k][tau_p]   comes back (the tau integral of G2(t, tau) per t)            */
    const int32_t* reduce_ch;  /* [n_reduce][2] = (tau_p, zero_p) output channels                                       */
    double reduce_spacing;     /* s: spacing of the tau axis                                                            */
    double* reduce_out;        /* host [n_traj][n_reduce] complex                                                       */
} aceqd_batch;

/*
 * Propagate a batch: builds the per-step operators exp(L dt/2) (batched scaling-and-squaring
 * kernel; replaces ACE's FreePropagator, general_system.py:324-327), then runs the fused
 * PT-step kernel (half step -> PT slice -> half step -> MTO -> closure/outputs; replaces
 * Simulation.run, general_system.py:331 / the `ACE <param>` subprocess of :339-341).
 * With host buffers the H2D/D2H copies happen inside this call (the end-to-end path).
 */
int aceqd_propagate_batch(aceqd_ctx* ctx, const aceqd_problem* prob, const aceqd_pt* pt,
                          const aceqd_batch* batch);

/* Stage 1 only: build (or rebuild) the per-step operators of `batch` into the context's
 * workspace.  Stage 2 only: run the step kernel on operators built by the previous stage-1
 * call with the same batch.  aceqd_propagate_batch = stage 1 + stage 2 (+ copies). */
int aceqd_build_operators(aceqd_ctx* ctx, const aceqd_problem* prob, const aceqd_batch* batch);
int aceqd_run_steps(aceqd_ctx* ctx, const aceqd_problem* prob, const aceqd_pt* pt,
                    const aceqd_batch* batch);

/* Copy one snapshot slot (NL x chi_pad complex, natural alpha order) to the host (tests). */
int aceqd_snapshot_read(aceqd_ctx* ctx, int slot, int NL, int chi_pad, double* host_out);

/* Batched matrix exponential exp(A_i) of n x n complex matrices (host in/out); the kernel
 * behind the operator builder.  Replaces `fprop.update(t, dt); fprop.M` (general_system.py:324-327). */
int aceqd_expm_batch(aceqd_ctx* ctx, int n, int count, const double* a_host, double* out_host);

/*
 * Time-local dynamical-map chains.  Replaces the reference's f2py/OpenMP/BLAS modules
 * `propagate_tau_module` (pyaceqd/two_time/propagate_tau.f90:3-536) and `timebin_tl`
 * (pyaceqd/timebin/timebin_tl.f90:23-397), which push Liouville vectors through chains of
 * NL x NL matrices (zgemv) with operator insertions and traces.  A chain is a program of
 * segments over one matrix pool `mats[n_mats][NL][NL]` (time-local maps, binary powers of the
 * stationary map, operator superoperators): segment (start, count, stride) applies
 * mats[start], mats[start+stride], ... (count matrices, stride 0 or 1) in order; after every step of a segment with emit != 0 the functionals
 * w[n_w][NL] of the vector are appended to the chain's output row.
 *   v0      [n_chains][NL]   start vectors        seg_off [n_chains+1] first segment of each chain
 *   out     [n_chains][n_emit_max][n_w] (may be NULL)   final_v [n_chains][NL] (may be NULL)
 * All pointers are host pointers; copies happen inside the call.
 */
typedef struct {
    int32_t start;  /* first matrix of the segment                                             */
    int32_t count;  /* consecutive matrices applied                                            */
    int32_t emit;   /* != 0: emit the output functionals after each step                       */
    int32_t stride; /* 1: consecutive matrices; 0: mats[start] applied `count` times            */
} aceqd_tlseg;

int aceqd_tlmap_run(aceqd_ctx* ctx, int NL, int n_mats, const double* mats, int n_chains,
                    const double* v0, const int64_t* seg_off, int64_t n_segs, const aceqd_tlseg* segs,
                    int n_w, const double* w, int n_emit_max, double* out, double* final_v);
/* Device time (ms) of the last aceqd_tlmap_run kernel. */
int aceqd_tlmap_last_ms(aceqd_ctx* ctx, float* ms);

/* Page-locked host memory for the end-to-end path (H2D of drive tables, D2H of outputs). */
int aceqd_host_alloc(size_t bytes, void** out);
void aceqd_host_free(void* p);

/* sizeof() of aceqd_seq, aceqd_entry, aceqd_traj, aceqd_batch as compiled (binding self-check). */
void aceqd_struct_sizes(int32_t out[4]);

/* Largest trajectories-per-tile T for which (NL, chi_pad) fits the step kernel's shared
 * memory budget (0 if even T=1 does not fit). */
int aceqd_max_tile(int NL, int chi_pad);
/* The same without a shared-memory ring for the PT chunks: tiles this large read the PT fragments from global
 * memory / L2 (launch name "... pt=global").  NL = 25, chi = 256: 2 instead of 1. */
int aceqd_max_tile_global_pt(int NL, int chi_pad);

/* Planner helper for aceqd_batch.kernel = 3: shared-memory bytes the split-K cluster kernel needs for tile_T = G
 * trajectories on a cluster of C CTAs (0: unsupported combination or does not fit). */
long long aceqd_splitk_fit(int NL, int chi_pad, int G, int C);

/* With more tiles than SMs the step kernel does not run in waves: the tiles are laid end to end and cut into
 * one equal piece of steps per SM; a tile that straddles a cut is started by one CTA and finished by the next
 * (bond state handed over through HBM).  This returns that schedule for `n_sm` CTAs (host only, no GPU):
 * segs_out = 5 ints per segment (tile, n_lo, n_hi, save_slot, load_slot; +-0x7fffffff = unbounded, -1 = none)
 * in execution order, seg_off_out = n_ctas+1 offsets into it. */
int aceqd_segment_plan(const aceqd_batch* batch, int n_sm, int max_segs, int32_t* segs_out,
                       int32_t* seg_off_out, int32_t* n_ctas, int32_t* n_slots);

/* Debug aid: per-phase cycle counters of the step kernel's CTA 0 (enable != 0 switches the clock on and zeroes it;
 * out8, if not NULL, receives the counters accumulated so far). */
int aceqd_debug_phase_ticks(aceqd_ctx* ctx, int enable, long long* out8);

/* DMMA m-tiles (8 rows) the most loaded CTA computes per step when a tile of T trajectories is shared
 * by a cluster of `cluster` CTAs (planner cost model; -1 on error). */
int aceqd_pass_load(const aceqd_problem* prob, int T, int cluster);

/* Register-resident FP64 micro-benchmarks (roofline denominators):
 * kind 0 = DMMA.8x8x4 tensor pipe, kind 1 = DFMA.  Returns TFLOP/s in *tflops. */
int aceqd_fp64_peak(aceqd_ctx* ctx, int kind, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* ACEQD_H */
