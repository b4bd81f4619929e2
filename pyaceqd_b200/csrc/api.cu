// C ABI of libaceqd.so (include/aceqd.h): context, PT re-tiling into the HBM layout the step
// kernel streams, problem upload, batch staging (H2D / D2H for the end-to-end path) and the
// two-stage launch  [operator builder] -> [persistent DMMA step kernel].
#include <cstdarg>
#include <algorithm>
#include <new>

#include "common.cuh"

namespace aceqd {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

void log_name(char (&dst)[128], const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, sizeof(dst), fmt, ap);
    va_end(ap);
}

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return ACEQD_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            p = nullptr;
            return ACEQD_ERR_NOMEM;
        }
        cap = want;
        return ACEQD_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace aceqd

using namespace aceqd;

struct aceqd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;   // device->host copy of finished waves while later waves run
    cudaEvent_t ev_wave = nullptr;
    cudaStream_t build_stream = nullptr;  // low priority: operators of the later waves, built on the SMs the
    cudaEvent_t ev_built = nullptr;       // first (partial) wave leaves idle
    // wave split decided by aceqd_propagate_batch for the current batch (0 = none)
    int split_tiles = 0;
    long long split_entries = 0, split_out = 0;
    bool split_ops_pending = false;
    LaunchLog log;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // step, opbuild, tlmap start/stop
    bool have_step = false, have_op = false, have_tl = false;
    DevBuf W, OV, tables, seqs, seq_base, entries, mto, rho0s, trajs, tiles, snap_steps, snaps, snap_r,
        out, passes, scratch, misc, octets, segs, seg_off, seg_state, seg_flags, opscratch, tl_pool, tl_v0, tl_segoff, tl_segs, tl_w, tl_out, tl_final;
    long long* ticks = nullptr;           // debug phase clock of the step kernel (aceqd_debug_phase_ticks)
    // layout of the operators currently in the workspace
    long long n_seq_entries = 0;
};

struct aceqd_pt {
    aceqd_ctx* ctx = nullptr;
    PtDev d{};
    void *blob = nullptr, *closure = nullptr, *kin = nullptr, *nout = nullptr, *off = nullptr;
    void *pblob = nullptr, *poff = nullptr;   // panel-ordered copy for the split-K kernel (chi_pad > PANEL only)
    long long blob_doubles = 0;
};

struct aceqd_problem {
    aceqd_ctx* ctx = nullptr;
    ProbDev d{};
    void* mem = nullptr;  // one allocation holding every array
    std::vector<int> pos_of_alpha, block_of_alpha;
};

extern "C" {

const char* aceqd_last_error(void) { return g_err.c_str(); }
const char* aceqd_version(void) { return "aceqd-b200 0.1 (sm_100a)"; }

int aceqd_ctx_create(int device, void* stream, aceqd_ctx** out) {
    if (!out) {
        set_error("aceqd_ctx_create: out is NULL");
        return ACEQD_ERR_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("no CUDA device available (%s); libaceqd has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return ACEQD_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (count %d)", device, count);
        return ACEQD_ERR_ARG;
    }
    ACEQD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ACEQD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device,
                  prop.major, prop.minor);
        return ACEQD_ERR_CUDA;
    }
    aceqd_ctx* c = new (std::nothrow) aceqd_ctx();
    if (!c) return ACEQD_ERR_NOMEM;
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        // high priority: the side stream that builds later waves' operators must only fill idle SMs
        int least = 0, greatest = 0;
        ACEQD_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        ACEQD_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatest));
        c->own_stream = true;
    }
    for (auto& ev : c->ev) ACEQD_CUDA(cudaEventCreate(&ev));
    *out = c;
    return ACEQD_OK;
}

void aceqd_ctx_destroy(aceqd_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->W, &c->OV, &c->tables, &c->seqs, &c->seq_base, &c->entries, &c->mto,
                      &c->rho0s, &c->trajs, &c->tiles, &c->snap_steps, &c->snaps, &c->snap_r, &c->out,
                      &c->passes, &c->scratch, &c->misc, &c->octets, &c->segs, &c->seg_off, &c->seg_state, &c->seg_flags, &c->opscratch, &c->tl_pool, &c->tl_v0, &c->tl_segoff,
                      &c->tl_segs, &c->tl_w, &c->tl_out, &c->tl_final})
        b->release();
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->ev_wave) cudaEventDestroy(c->ev_wave);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_built) cudaEventDestroy(c->ev_built);
    if (c->build_stream) cudaStreamDestroy(c->build_stream);
    if (c->ticks) cudaFree(c->ticks);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int aceqd_ctx_sync(aceqd_ctx* c) {
    if (!c) return ACEQD_ERR_ARG;
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    return ACEQD_OK;
}

long long aceqd_launch_count(const aceqd_ctx* c) { return c ? c->log.count : 0; }
const char* aceqd_last_step_kernel(const aceqd_ctx* c) { return c ? c->log.step : ""; }
const char* aceqd_last_opbuild_kernel(const aceqd_ctx* c) { return c ? c->log.opbuild : ""; }
const char* aceqd_last_other_kernel(const aceqd_ctx* c) { return c ? c->log.other : ""; }

int aceqd_last_timings(aceqd_ctx* c, float* step_ms, float* op_ms) {
    if (!c) return ACEQD_ERR_ARG;
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    if (step_ms) {
        *step_ms = 0.f;
        if (c->have_step) ACEQD_CUDA(cudaEventElapsedTime(step_ms, c->ev[0], c->ev[1]));
    }
    if (op_ms) {
        *op_ms = 0.f;
        if (c->have_op) ACEQD_CUDA(cudaEventElapsedTime(op_ms, c->ev[2], c->ev[3]));
    }
    return ACEQD_OK;
}

// ------------------------------------------------------------------------------- PT
int aceqd_pt_create(aceqd_ctx* c, int n_cls, int n_slices, int n_initial, const int* chi_in,
                    const int* chi_out, const double* const* slices,
                    const double* const* closures, aceqd_pt** out) {
    if (!c || !out || n_cls <= 0 || n_slices <= 0 || n_initial < 0 || n_initial >= n_slices ||
        !chi_in || !chi_out || !slices || !closures) {
        set_error("aceqd_pt_create: invalid argument");
        return ACEQD_ERR_ARG;
    }
    *out = nullptr;
    ACEQD_CUDA(cudaSetDevice(c->device));
    int chi_max = 1;
    for (int s = 0; s < n_slices; ++s) {
        if (chi_in[s] <= 0 || chi_out[s] <= 0) {
            set_error("aceqd_pt_create: slice %d has non-positive bond dimension", s);
            return ACEQD_ERR_ARG;
        }
        chi_max = std::max(chi_max, std::max(chi_in[s], chi_out[s]));
    }
    const int chi_pad = round_up(chi_max, 8);
    const int strideB = chi_pad + 4;
    const int chunk_doubles = 2 * KC * strideB;
    std::vector<int> kin(n_slices), nout(n_slices);
    std::vector<long long> off(n_slices);
    long long total = 0;
    for (int s = 0; s < n_slices; ++s) {
        kin[s] = round_up(chi_in[s], KC);
        nout[s] = round_up(chi_out[s], 8);
        off[s] = total;
        total += (long long)n_cls * (kin[s] / KC) * chunk_doubles;
    }
    std::vector<double> blob((size_t)total, 0.0);
    std::vector<double> clo((size_t)n_slices * 2 * chi_pad, 0.0);
    for (int s = 0; s < n_slices; ++s) {
        const int din = chi_in[s], dout = chi_out[s];
        const int nch = kin[s] / KC;
        for (int b = 0; b < n_cls; ++b) {
            const double* src = slices[s] + (size_t)b * din * dout * 2;
            for (int d1 = 0; d1 < din; ++d1) {
                double* ch = blob.data() + off[s] + ((size_t)b * nch + d1 / KC) * chunk_doubles;
                double* re = ch + (size_t)(d1 % KC) * strideB;
                double* im = re + (size_t)KC * strideB;
                const double* row = src + (size_t)d1 * dout * 2;
                for (int d2 = 0; d2 < dout; ++d2) {
                    re[d2] = row[2 * d2];
                    im[d2] = row[2 * d2 + 1];
                }
            }
        }
        for (int d = 0; d < dout; ++d) {
            clo[((size_t)s * chi_pad + d) * 2] = closures[s][2 * d];
            clo[((size_t)s * chi_pad + d) * 2 + 1] = closures[s][2 * d + 1];
        }
    }
    aceqd_pt* pt = new (std::nothrow) aceqd_pt();
    if (!pt) return ACEQD_ERR_NOMEM;
    pt->ctx = c;
    auto fail = [&](int rc) {
        aceqd_pt_destroy(pt);
        return rc;
    };
#define PT_UP(dst, vec)                                                                   \
    do {                                                                                  \
        size_t bytes_ = (vec).size() * sizeof((vec)[0]);                                  \
        if (cudaMalloc(&(dst), bytes_ ? bytes_ : 16) != cudaSuccess) {                    \
            set_error("aceqd_pt_create: cudaMalloc(%zu) failed", bytes_);                 \
            return fail(ACEQD_ERR_NOMEM);                                                 \
        }                                                                                 \
        if (cudaMemcpyAsync((dst), (vec).data(), bytes_, cudaMemcpyHostToDevice,          \
                            c->stream) != cudaSuccess) {                                  \
            set_error("aceqd_pt_create: upload failed");                                  \
            return fail(ACEQD_ERR_CUDA);                                                  \
        }                                                                                 \
    } while (0)
    PT_UP(pt->blob, blob);
    PT_UP(pt->closure, clo);
    PT_UP(pt->kin, kin);
    PT_UP(pt->nout, nout);
    PT_UP(pt->off, off);
    // chi_pad > PANEL: a second copy cut into panels of PANEL output columns, [slice][cls][panel][chunk], row stride
    // PANEL + 4 (what one split-K GEMM pass streams)
    const int n_panels = (chi_pad + PANEL - 1) / PANEL;
    std::vector<double> pblob;
    std::vector<long long> poff(n_slices, 0);
    if (chi_pad > PANEL) {
        const int pstride = PANEL + 4, pchunk = 2 * KC * pstride;
        long long ptotal = 0;
        for (int s = 0; s < n_slices; ++s) {
            poff[s] = ptotal;
            ptotal += (long long)n_cls * n_panels * (kin[s] / KC) * pchunk;
        }
        pblob.assign((size_t)ptotal, 0.0);
        for (int s = 0; s < n_slices; ++s) {
            const int din = chi_in[s], dout = chi_out[s], nch = kin[s] / KC;
            for (int b = 0; b < n_cls; ++b) {
                const double* src = slices[s] + (size_t)b * din * dout * 2;
                for (int d1 = 0; d1 < din; ++d1)
                    for (int d2 = 0; d2 < dout; ++d2) {
                        const int q = d2 / PANEL, c = d2 - q * PANEL;
                        double* ch = pblob.data() + poff[s] + (((size_t)b * n_panels + q) * nch + d1 / KC) * pchunk;
                        ch[(size_t)(d1 % KC) * pstride + c] = src[((size_t)d1 * dout + d2) * 2];
                        ch[(size_t)KC * pstride + (size_t)(d1 % KC) * pstride + c] = src[((size_t)d1 * dout + d2) * 2 + 1];
                    }
            }
        }
        PT_UP(pt->pblob, pblob);
        PT_UP(pt->poff, poff);
    }
#undef PT_UP
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
        set_error("aceqd_pt_create: sync failed");
        return fail(ACEQD_ERR_CUDA);
    }
    pt->blob_doubles = total;
    PtDev& d = pt->d;
    d.n_cls = n_cls;
    d.n_slices = n_slices;
    d.n_initial = n_initial;
    d.n_repeat = n_slices - n_initial;
    d.chi_pad = chi_pad;
    d.strideB = strideB;
    d.chunk_doubles = chunk_doubles;
    d.kin_pad = (const int*)pt->kin;
    d.nout_pad = (const int*)pt->nout;
    d.off = (const long long*)pt->off;
    d.blob = (const double*)pt->blob;
    d.closure = (const double*)pt->closure;
    d.pblob = (const double*)pt->pblob;
    d.poff = (const long long*)pt->poff;
    d.n_panels = pt->pblob ? n_panels : 1;
    *out = pt;
    return ACEQD_OK;
}

void aceqd_pt_destroy(aceqd_pt* pt) {
    if (!pt) return;
    if (pt->ctx) cudaSetDevice(pt->ctx->device);
    for (void* p : {pt->blob, pt->closure, pt->kin, pt->nout, pt->off, pt->pblob, pt->poff})
        if (p) cudaFree(p);
    delete pt;
}

int aceqd_pt_chi_pad(const aceqd_pt* pt) { return pt ? pt->d.chi_pad : 0; }

// ------------------------------------------------------------------------------- problem
int aceqd_problem_create(aceqd_ctx* c, int NL, int n_fields, int n_out, const double* L0,
                         const double* LA, const double* LB, const int* field_table,
                         const double* out_w, const int* pos_of_alpha, const int* block_of_alpha,
                         aceqd_problem** out) {
    if (!c || !out || NL <= 0 || NL > MAX_NL || n_fields < 0 || n_out < 0 || !L0 ||
        !pos_of_alpha || !block_of_alpha || (n_fields && (!LA || !LB || !field_table)) ||
        (n_out && !out_w)) {
        set_error("aceqd_problem_create: invalid argument (NL must be 1..%d)", MAX_NL);
        return ACEQD_ERR_ARG;
    }
    *out = nullptr;
    // pos_of_alpha must be a permutation whose block sequence is grouped
    std::vector<int> blk_of_pos(NL, -1);
    for (int a = 0; a < NL; ++a) {
        const int ps = pos_of_alpha[a];
        if (ps < 0 || ps >= NL || blk_of_pos[ps] != -1 || block_of_alpha[a] < 0) {
            set_error("aceqd_problem_create: pos_of_alpha is not a permutation / negative block");
            return ACEQD_ERR_ARG;
        }
        blk_of_pos[ps] = block_of_alpha[a];
    }
    ACEQD_CUDA(cudaSetDevice(c->device));
    const size_t n2 = (size_t)NL * NL;
    const size_t b_L0 = n2 * 16, b_LA = (size_t)n_fields * n2 * 16, b_ow = (size_t)n_out * NL * 16;
    const size_t b_int = (size_t)(n_fields + 2 * NL) * sizeof(int);
    const size_t total = b_L0 + 2 * b_LA + b_ow + b_int + 64;
    aceqd_problem* p = new (std::nothrow) aceqd_problem();
    if (!p) return ACEQD_ERR_NOMEM;
    p->ctx = c;
    if (cudaMalloc(&p->mem, total) != cudaSuccess) {
        delete p;
        set_error("aceqd_problem_create: cudaMalloc failed");
        return ACEQD_ERR_NOMEM;
    }
    std::vector<unsigned char> host(total, 0);
    size_t o = 0;
    auto put = [&](const void* src, size_t bytes) {
        size_t at = o;
        if (bytes) memcpy(host.data() + o, src, bytes);
        o += (bytes + 15) / 16 * 16;
        return at;
    };
    const size_t o_L0 = put(L0, b_L0), o_LA = put(LA, b_LA), o_LB = put(LB, b_LA),
                 o_ow = put(out_w, b_ow);
    const size_t o_ft = put(field_table, n_fields * sizeof(int));
    const size_t o_pos = put(pos_of_alpha, NL * sizeof(int));
    const size_t o_blk = put(block_of_alpha, NL * sizeof(int));
    if (cudaMemcpy(p->mem, host.data(), total, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(p->mem);
        delete p;
        set_error("aceqd_problem_create: upload failed");
        return ACEQD_ERR_CUDA;
    }
    unsigned char* base = (unsigned char*)p->mem;
    ProbDev& d = p->d;
    d.NL = NL;
    d.NLp8 = round_up(NL, 8);
    d.NLp4 = round_up(NL, 4);
    d.n_out = n_out;
    d.n_fields = n_fields;
    d.w_doubles = 2 * d.NLp8 * d.NLp4;
    d.ov_doubles = 2 * n_out * NL;
    d.L0 = (const double*)(base + o_L0);
    d.LA = (const double*)(base + o_LA);
    d.LB = (const double*)(base + o_LB);
    d.out_w = (const double*)(base + o_ow);
    d.field_table = (const int*)(base + o_ft);
    d.pos_of_alpha = (const int*)(base + o_pos);
    d.block_of_alpha = (const int*)(base + o_blk);
    p->pos_of_alpha.assign(pos_of_alpha, pos_of_alpha + NL);
    p->block_of_alpha.assign(block_of_alpha, block_of_alpha + NL);
    *out = p;
    return ACEQD_OK;
}

void aceqd_problem_destroy(aceqd_problem* p) {
    if (!p) return;
    if (p->ctx) cudaSetDevice(p->ctx->device);
    if (p->mem) cudaFree(p->mem);
    delete p;
}

// ------------------------------------------------------------------------------- batch
static int check_batch(const aceqd_problem* prob, const aceqd_batch* b) {
    if (!prob || !b) {
        set_error("batch: NULL argument");
        return ACEQD_ERR_ARG;
    }
    if (b->n_traj <= 0 || !b->trajs || b->n_seq < 0 || (b->n_seq && !b->seqs) ||
        b->n_entries < 0 || (b->n_entries && !b->entries) || b->n_rho0 < 0 ||
        (b->n_rho0 && !b->rho0s) || !b->out || b->out_elems <= 0 || b->dt <= 0.0) {
        set_error("batch: invalid sizes or missing pointers");
        return ACEQD_ERR_ARG;
    }
    if (b->n_tables > 0 && (b->n_sets <= 0 || b->n_samples <= 0 || !b->tables || b->tab_dt <= 0)) {
        set_error("batch: inconsistent drive tables");
        return ACEQD_ERR_ARG;
    }
    return ACEQD_OK;
}

#define UP(buf, src, bytes)                                                                   \
    do {                                                                                      \
        int rc_ = (buf).reserve(bytes);                                                       \
        if (rc_) return rc_;                                                                  \
        if ((bytes) > 0)                                                                      \
            ACEQD_CUDA(cudaMemcpyAsync((buf).p, (src), (bytes), cudaMemcpyHostToDevice,       \
                                       c->stream));                                           \
    } while (0)

int aceqd_build_operators(aceqd_ctx* c, const aceqd_problem* prob, const aceqd_batch* b) {
    int rc = check_batch(prob, b);
    if (rc) return rc;
    if (!c) return ACEQD_ERR_ARG;
    ACEQD_CUDA(cudaSetDevice(c->device));
    const ProbDev& pd = prob->d;
    std::vector<long long> base(b->n_seq + 1, 0);
    for (int q = 0; q < b->n_seq; ++q) {
        if (b->seqs[q].len <= 0 || b->seqs[q].set < 0 ||
            (b->n_tables > 0 && b->seqs[q].set >= b->n_sets)) {
            set_error("batch: sequence %d invalid", q);
            return ACEQD_ERR_ARG;
        }
        base[q + 1] = base[q] + b->seqs[q].len;
    }
    for (int e = 0; e < b->n_entries; ++e) {
        const aceqd_entry& en = b->entries[e];
        if (en.sb >= b->n_mto_mats || en.sa >= b->n_mto_mats || en.set < 0 ||
            (b->n_tables > 0 && en.set >= b->n_sets)) {
            set_error("batch: explicit entry %d invalid", e);
            return ACEQD_ERR_ARG;
        }
    }
    const long long n_ent = base[b->n_seq] + b->n_entries;
    c->n_seq_entries = base[b->n_seq];
    const size_t n2 = (size_t)pd.NL * pd.NL;
    if (b->device_resident) {
        // tables already in HBM: use the caller's pointer directly
    } else {
        UP(c->tables, b->tables, (size_t)b->n_sets * b->n_tables * b->n_samples * 16);
    }
    UP(c->seqs, b->seqs, (size_t)b->n_seq * sizeof(aceqd_seq));
    UP(c->seq_base, base.data(), base.size() * sizeof(long long));
    UP(c->entries, b->entries, (size_t)b->n_entries * sizeof(aceqd_entry));
    UP(c->mto, b->mto_mats, (size_t)b->n_mto_mats * n2 * 16);
    if ((rc = c->W.reserve((size_t)n_ent * pd.w_doubles * 8))) return rc;
    if ((rc = c->OV.reserve((size_t)n_ent * pd.ov_doubles * 8 + 16))) return rc;

    OpBuildParams op{};
    op.prob = pd;
    op.dt = b->dt;
    op.t0 = b->t0;
    op.eval_off1 = b->eval_off1;
    op.eval_off2 = b->eval_off2;
    op.n_sets = b->n_sets;
    op.n_tables = b->n_tables;
    op.n_samples = b->n_samples;
    op.n_seq = b->n_seq;
    op.tab_t0 = b->tab_t0;
    op.tab_dt = b->tab_dt > 0 ? b->tab_dt : 1.0;
    op.tables = b->device_resident ? b->tables : (const double*)c->tables.p;
    op.seqs = (const aceqd_seq*)c->seqs.p;
    op.seq_base = (const long long*)c->seq_base.p;
    op.n_seq_entries = base[b->n_seq];
    op.n_entries = b->n_entries;
    op.n_mto_mats = b->n_mto_mats;
    op.entries = (const aceqd_entry*)c->entries.p;
    op.mto_mats = (const double*)c->mto.p;
    op.W = (double*)c->W.p;
    op.OV = (double*)c->OV.p;
    {
        int ctas = 0;
        const size_t need = opbuild_scratch_bytes(pd.NL, &ctas);
        if (need) {
            if ((rc = c->opscratch.reserve(need))) return rc;
            op.scratch = (double*)c->opscratch.p;
            op.scratch_ctas = ctas;
        }
    }
    ACEQD_CUDA(cudaEventRecord(c->ev[2], c->stream));
    c->split_ops_pending = false;
    if (c->split_entries > 0 && c->split_entries < n_ent && b->n_entries == 0) {
        // operators of the first (partial) wave now; those of the later waves on a low-priority stream
        // that fills the SMs the first wave leaves idle (aceqd_run_steps waits for them before wave 2)
        if (!c->build_stream) {
            int lo_prio = 0, hi_prio = 0;
            ACEQD_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            ACEQD_CUDA(cudaStreamCreateWithPriority(&c->build_stream, cudaStreamNonBlocking, lo_prio));
            ACEQD_CUDA(cudaEventCreateWithFlags(&c->ev_built, cudaEventDisableTiming));
        }
        OpBuildParams first = op, rest = op;
        first.e_begin = 0;
        first.e_end = c->split_entries;
        rest.e_begin = c->split_entries;
        rest.e_end = n_ent;
        if ((rc = launch_opbuild(first, c->stream, &c->log))) return rc;
        ACEQD_CUDA(cudaEventRecord(c->ev[3], c->stream));
        // the uploads above were enqueued on the main stream: order the side stream after them
        ACEQD_CUDA(cudaEventRecord(c->ev_built, c->stream));
        ACEQD_CUDA(cudaStreamWaitEvent(c->build_stream, c->ev_built, 0));
        if ((rc = launch_opbuild(rest, c->build_stream, &c->log))) return rc;
        ACEQD_CUDA(cudaEventRecord(c->ev_built, c->build_stream));
        c->split_ops_pending = true;
        c->have_op = true;
        return ACEQD_OK;
    }
    if ((rc = launch_opbuild(op, c->stream, &c->log))) return rc;
    ACEQD_CUDA(cudaEventRecord(c->ev[3], c->stream));
    c->have_op = true;
    return ACEQD_OK;
}

static int build_passes(const aceqd_problem* prob, int T, int cluster, std::vector<PassDesc>& passes) {
    const int NL = prob->d.NL;
    std::vector<int> blk_of_pos(NL);
    for (int a = 0; a < NL; ++a) blk_of_pos[prob->pos_of_alpha[a]] = prob->block_of_alpha[a];
    passes.clear();
    int p0 = 0;
    while (p0 < NL) {
        int p1 = p0;
        while (p1 < NL && blk_of_pos[p1] == blk_of_pos[p0]) ++p1;
        // rows [p0*T, p1*T) share PT block blk_of_pos[p0]
        int row = p0 * T;
        const int row_end = p1 * T;
        while (row < row_end) {
            PassDesc pd{};
            pd.blk = blk_of_pos[p0];
            for (int mc = 0; mc < MC; ++mc) {
                pd.row0[mc] = row < row_end ? row : 0;
                pd.nvalid[mc] = std::max(0, std::min(8, row_end - row));
                row += 8;
            }
            passes.push_back(pd);
        }
        p0 = p1;
    }
    if ((int)passes.size() > MAX_PASSES) {
        set_error("tile needs %zu GEMM passes (max %d)", passes.size(), MAX_PASSES);
        return ACEQD_ERR_CAPACITY;
    }
    // owners: heaviest pass first onto the least loaded CTA of the cluster (load = m-tiles)
    std::vector<int> order(passes.size()), load(std::max(1, cluster), 0);
    auto weight = [&](int i) {
        int w = 0;
        for (int mc = 0; mc < MC; ++mc) w += passes[i].nvalid[mc] > 0;
        return w;
    };
    // ties in m-tiles go to the CTA with fewer ROWS so far: the system product of a CTA costs ceil(own alphas / 8)
    // m-tiles per trajectory, and an uneven split (10 + 6 alphas for the biexciton at T = 4) doubles it on one CTA
    auto rows_of = [&](int i) {
        int r = 0;
        for (int mc = 0; mc < MC; ++mc) r += passes[i].nvalid[mc];
        return r;
    };
    std::vector<int> rows(load.size(), 0);
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return weight(a) != weight(b) ? weight(a) > weight(b) : rows_of(a) > rows_of(b);
    });
    for (int i : order) {
        int best = 0;
        for (int r = 1; r < (int)load.size(); ++r)
            if (load[r] < load[best] || (load[r] == load[best] && rows[r] < rows[best])) best = r;
        passes[i].owner = best;
        load[best] += weight(i);
        rows[best] += rows_of(i);
    }
    return ACEQD_OK;
}

// More tiles than SMs: instead of running the persistent CTAs in waves (the last one partial), lay the
// tiles end to end on a line of step counts and cut it into n_sm equal pieces (McNaughton's wrap-around
// rule for preemptive scheduling: makespan = max(longest tile, total / n_sm)).  A tile that straddles a
// cut is STARTED by the left CTA and FINISHED by the right one; every CTA runs its piece right to left,
// so the head of a cut tile is the first thing its CTA does and the tail the last thing the next CTA
// does (never a circular wait; the producer of a slot has the lower block index).
static int segment_ctas(const aceqd_ctx* c) {   // CTAs of a segmented launch: one per SM (ACEQD_SEG_SMS overrides: tests)
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
    const char* env = getenv("ACEQD_SEG_SMS");
    if (env && atoi(env) > 0) n_sm = std::min(n_sm, atoi(env));
    return n_sm;
}

static bool use_segments(const aceqd_ctx* c, const aceqd_batch* b) {
    if (!c || !b || (b->kernel != 0 && b->kernel != 5) || b->cluster > 1 || !b->tile_traj || b->tile_T < 1) return false;
    const char* env = getenv("ACEQD_SEGMENTS");
    if (env && env[0] == '0') return false;
    return b->n_tiles > segment_ctas(c);
}

static bool plan_segments(const aceqd_batch* b, int n_sm, std::vector<SegDesc>& segs, std::vector<int>& seg_off,
                          int& n_slots) {
    const int T = b->tile_T, MINSEG = 8;
    const int INF = 0x7fffffff;
    std::vector<int> t_begin(b->n_tiles, INF), t_len(b->n_tiles, 0);
    long long total = 0;
    int longest = 1;
    for (int i = 0; i < b->n_tiles; ++i) {
        int e_max = -1;
        for (int j = 0; j < T; ++j) {
            const int idx = b->tile_traj[(size_t)i * T + j];
            if (idx < 0) continue;
            const aceqd_traj& t = b->trajs[idx];
            t_begin[i] = std::min(t_begin[i], t.step0);
            e_max = std::max(e_max, t.step0 + t.n_steps);
        }
        if (e_max < 0) continue;                       // empty tile: no work
        t_len[i] = std::max(1, e_max - t_begin[i]);    // a zero-step tile still writes its row 0
        total += t_len[i];
        longest = std::max(longest, t_len[i]);
    }
    if (total == 0) return false;
    struct Piece { int cta, tile, lo, hi, save, load; };
    std::vector<Piece> pieces;
    long long piece = std::max<long long>(longest, (total + n_sm - 1) / n_sm);
    for (int attempt = 0; attempt < 64; ++attempt, piece += MINSEG) {
        pieces.clear();
        n_slots = 0;
        int cta = 0;
        long long room = piece;
        for (int i = 0; i < b->n_tiles; ++i) {
            if (t_len[i] == 0) continue;
            int done = 0, pending_slot = -1;
            while (done < t_len[i]) {
                if (room <= 0) {
                    ++cta;
                    room = piece;
                }
                const int r = t_len[i] - done;
                long long take = std::min<long long>(r, room);
                if (take < r) {                        // a cut inside the tile
                    if (take < MINSEG || done > 0) {   // too short a head (or a third piece): start in the next CTA
                        room = 0;
                        if (done > 0) take = r;        // (cannot happen: r <= piece) keep the tail whole
                        else continue;
                    } else if (r - take < MINSEG) {
                        take = r;                      // too short a tail: overfill this CTA slightly
                    }
                }
                Piece pc{cta, i, -INF, INF, -1, -1};
                if (done > 0) {
                    pc.lo = t_begin[i] + done;
                    pc.load = pending_slot;
                }
                if (take < r) {
                    pc.hi = t_begin[i] + done + (int)take;
                    pc.save = pending_slot = n_slots++;
                }
                pieces.push_back(pc);
                done += (int)take;
                room -= take;
            }
        }
        if (cta < n_sm) break;
        if (attempt == 63) return false;
    }
    const int n_ctas = pieces.back().cta + 1;
    seg_off.assign(n_ctas + 1, 0);
    for (const Piece& pc : pieces) ++seg_off[pc.cta + 1];
    for (int k = 0; k < n_ctas; ++k) seg_off[k + 1] += seg_off[k];
    segs.assign(pieces.size(), SegDesc{});
    std::vector<int> fill(n_ctas, 0);
    for (const Piece& pc : pieces) {                   // right to left within a CTA
        const int cnt = seg_off[pc.cta + 1] - seg_off[pc.cta];
        const int at = seg_off[pc.cta] + (cnt - 1 - fill[pc.cta]++);
        segs[at] = SegDesc{pc.tile, pc.lo, pc.hi, pc.save, pc.load};
    }
    return true;
}

/* m-tiles the most loaded CTA of a `cluster` computes per step for tile size T (planner cost model) */
extern "C" int aceqd_pass_load(const aceqd_problem* prob, int T, int cluster) {
    if (!prob || T < 1 || cluster < 1) return -1;
    std::vector<PassDesc> passes;
    if (build_passes(prob, T, cluster, passes)) return -1;
    std::vector<int> load(cluster, 0);
    for (auto& pd : passes)
        for (int mc = 0; mc < MC; ++mc) load[pd.owner] += pd.nvalid[mc] > 0;
    return *std::max_element(load.begin(), load.end());
}

/* The segment schedule of a batch on n_sm CTAs (host only; tests).  segs_out: 5 ints per segment
 * (tile, n_lo, n_hi, save_slot, load_slot) in execution order per CTA; seg_off_out: n_ctas+1 offsets. */
int aceqd_segment_plan(const aceqd_batch* b, int n_sm, int max_segs, int32_t* segs_out, int32_t* seg_off_out,
                       int32_t* n_ctas, int32_t* n_slots) {
    if (!b || !b->tile_traj || !b->trajs || b->tile_T < 1 || n_sm < 1 || !segs_out || !seg_off_out || !n_ctas ||
        !n_slots) {
        set_error("aceqd_segment_plan: invalid argument");
        return ACEQD_ERR_ARG;
    }
    std::vector<SegDesc> segs;
    std::vector<int> off;
    int slots = 0;
    if (!plan_segments(b, n_sm, segs, off, slots)) {
        set_error("aceqd_segment_plan: no schedule");
        return ACEQD_ERR_CAPACITY;
    }
    if ((int)segs.size() > max_segs) {
        set_error("aceqd_segment_plan: %zu segments exceed the caller's buffer", segs.size());
        return ACEQD_ERR_CAPACITY;
    }
    for (size_t i = 0; i < segs.size(); ++i) {
        const int32_t v[5] = {segs[i].tile, segs[i].n_lo, segs[i].n_hi, segs[i].save_slot, segs[i].load_slot};
        memcpy(segs_out + 5 * i, v, sizeof(v));
    }
    for (size_t i = 0; i < off.size(); ++i) seg_off_out[i] = off[i];
    *n_ctas = (int)off.size() - 1;
    *n_slots = slots;
    return ACEQD_OK;
}

/* Debug: phase clock of the step kernel.  enable != 0 switches it on (and zeroes it); out8 (may be NULL) gets the
 * cycles CTA 0 spent after each of its 7 marks: [0] outputs wait, [1] outputs, [2] system product + barrier,
 * [3] PT GEMM passes, [4] barrier after the passes, [5] closure sums + barrier, [6] step tail. */
int aceqd_debug_phase_ticks(aceqd_ctx* c, int enable, long long* out8) {
    if (!c) return ACEQD_ERR_ARG;
    ACEQD_CUDA(cudaSetDevice(c->device));
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    if (out8 && c->ticks) ACEQD_CUDA(cudaMemcpy(out8, c->ticks, 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    else if (out8) memset(out8, 0, 8 * sizeof(long long));
    if (enable) {
        if (!c->ticks) ACEQD_CUDA(cudaMalloc(&c->ticks, 8 * sizeof(long long)));
        ACEQD_CUDA(cudaMemset(c->ticks, 0, 8 * sizeof(long long)));
    } else if (c->ticks) {
        cudaFree(c->ticks);
        c->ticks = nullptr;
    }
    return ACEQD_OK;
}

/* Shared-memory bytes of the split-K cluster kernel for G trajectories on a cluster of C CTAs (deepest chunk ring
 * that fits; 0 if the combination is unsupported or does not fit). */
long long aceqd_splitk_fit(int NL, int chi_pad, int G, int C) {
    for (int st = MAX_STAGES; st >= 2; --st) {
        const size_t smem = splitk_smem_bytes(NL, chi_pad, G, C, st);
        if (smem && smem <= (size_t)SMEM_BUDGET) return (long long)smem;
    }
    return 0;
}

int aceqd_max_tile(int NL, int chi_pad) {
    for (int T = MAX_TILE_T; T >= 1; T >>= 1)
        if (step_smem_bytes(NL, chi_pad, T, 2, 0, 1) <= (size_t)SMEM_BUDGET) return T;
    return 0;
}

int aceqd_max_tile_global_pt(int NL, int chi_pad) {
    for (int T = MAX_TILE_T; T >= 1; T >>= 1)
        if (step_smem_bytes(NL, chi_pad, T, 0, 0, 1) <= (size_t)SMEM_BUDGET) return T;
    return 0;
}

// End of a launch with host buffers: either the outputs go back (staged copy unless the kernel wrote them straight
// into page-locked memory), or -- fused tail reduction -- they stay in HBM and only the per-trajectory trapezoids do.
static int finish_outputs(aceqd_ctx* c, const aceqd_batch* b, int n_out, double* out_dev, bool zero_copy, bool reduce) {
    if (b->device_resident) return ACEQD_OK;
    if (reduce) {
        int rc;
        const size_t res_bytes = (size_t)b->n_traj * b->n_reduce * 16;
        if ((rc = c->misc.reserve(res_bytes + (size_t)b->n_reduce * 2 * sizeof(int) + 64))) return rc;
        int* ch_dev = (int*)((char*)c->misc.p + (res_bytes + 15) / 16 * 16);
        ACEQD_CUDA(cudaMemcpyAsync(ch_dev, b->reduce_ch, (size_t)b->n_reduce * 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        if ((rc = launch_tail_reduce((const aceqd_traj*)c->trajs.p, b->n_traj, n_out, out_dev, b->n_reduce, ch_dev,
                                     b->reduce_spacing, (double*)c->misc.p, c->stream, &c->log)))
            return rc;
        ACEQD_CUDA(cudaMemcpyAsync(b->reduce_out, c->misc.p, res_bytes, cudaMemcpyDeviceToHost, c->stream));
    } else if (!zero_copy) {
        ACEQD_CUDA(cudaMemcpyAsync(b->out, out_dev, (size_t)b->out_elems * 16, cudaMemcpyDeviceToHost, c->stream));
    }
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    return ACEQD_OK;
}

int aceqd_run_steps(aceqd_ctx* c, const aceqd_problem* prob, const aceqd_pt* pt,
                    const aceqd_batch* b) {
    int rc = check_batch(prob, b);
    if (rc) return rc;
    if (!c || !pt) return ACEQD_ERR_ARG;
    ACEQD_CUDA(cudaSetDevice(c->device));
    const ProbDev& pd = prob->d;
    const int chi_pad = pt->d.chi_pad;
    for (int a = 0; a < pd.NL; ++a)
        if (prob->block_of_alpha[a] >= pt->d.n_cls) {
            set_error("problem references PT block %d, PT has %d", prob->block_of_alpha[a],
                      pt->d.n_cls);
            return ACEQD_ERR_ARG;
        }
    const int T = b->tile_T;
    if (b->kernel != 1) {
        if (T < 1 || T > MAX_TILE_T || b->n_tiles <= 0 || !b->tile_traj) {
            set_error("batch: tile_T must be 1..%d with a tile list", MAX_TILE_T);
            return ACEQD_ERR_ARG;
        }
    }
    // validate trajectories
    const size_t snap_bytes = (size_t)pd.NL * chi_pad * 16;
    for (int i = 0; i < b->n_traj; ++i) {
        const aceqd_traj& t = b->trajs[i];
        const char* why = nullptr;
        if (t.n_steps < 0 || t.step0 < 0) why = "negative step count";
        else if (t.ent0 < 0 || t.ent0 + t.n_steps >= c->n_seq_entries)
            why = "operator entries beyond the sequence pool";
        else if (t.n_ovr < 0 || t.n_ovr > ACEQD_MAX_OVR) why = "too many override entries";
        else if (t.out_from < 0 || t.out_from > t.n_steps + 1) why = "out_from outside the trajectory";
        else if (t.out_off < 0 ||
                 t.out_off + (long long)(t.n_steps + 1 - t.out_from) * pd.n_out > b->out_elems)
            why = "output block outside the output buffer";
        else if (t.init_kind == 0 && (t.init_index < 0 || t.init_index >= b->n_rho0))
            why = "initial state index out of range";
        else if (t.init_kind == 1 && (t.init_index < 0 ||
                 ((size_t)(t.init_index + 1) * snap_bytes > c->snaps.cap &&
                  t.init_index >= b->n_snap_slots)))
            why = "snapshot slot not available";
        else if (t.init_kind != 0 && t.init_kind != 1) why = "unknown init_kind";
        else if (t.snap_cnt < 0 || (t.snap_cnt > 0 && (t.snap_off < 0 ||
                 t.snap_off + t.snap_cnt > b->n_snap_steps || t.snap_slot0 < 0 ||
                 t.snap_slot0 + t.snap_cnt > b->n_snap_slots)))
            why = "snapshot request out of range";
        for (int q = 0; !why && q < t.n_ovr; ++q)
            if (t.ovr_ent[q] < 0 || t.ovr_ent[q] >= b->n_entries) why = "override entry out of range";
        if (why) {
            set_error("batch: trajectory %d: %s", i, why);
            return ACEQD_ERR_ARG;
        }
    }
    UP(c->rho0s, b->rho0s, (size_t)b->n_rho0 * pd.NL * 16);
    UP(c->trajs, b->trajs, (size_t)b->n_traj * sizeof(aceqd_traj));
    UP(c->snap_steps, b->snap_steps, (size_t)b->n_snap_steps * sizeof(int32_t));
    if (b->n_snap_slots > 0) {
        // growing discards old snapshots; writers re-create them in this call
        if ((rc = c->snaps.reserve((size_t)b->n_snap_slots * pd.NL * chi_pad * 16))) return rc;
    }
    // closures of the snapshot rows (same slots); kept as large as the snapshot pool
    if (c->snaps.cap > 0 && (rc = c->snap_r.reserve(c->snaps.cap / (size_t)chi_pad + 64))) return rc;
    double* out_dev = nullptr;
    bool zero_copy = false;   // page-locked host output: the kernel writes its rows straight through PCIe
    const bool reduce = b->n_reduce > 0 && !b->device_resident && b->reduce_ch && b->reduce_out;
    if (reduce) {
        for (int p = 0; p < 2 * b->n_reduce; ++p)
            if (b->reduce_ch[p] < 0 || b->reduce_ch[p] >= pd.n_out) {
                set_error("batch: reduce_ch[%d] = %d is not an output channel", p, b->reduce_ch[p]);
                return ACEQD_ERR_ARG;
            }
        if ((rc = c->out.reserve((size_t)b->out_elems * 16))) return rc;
        out_dev = (double*)c->out.p;      // the outputs stay in HBM
        c->split_tiles = 0;               // (no early copy of finished waves either)
    } else if (b->device_resident) {
        out_dev = b->out;
    } else {
        cudaPointerAttributes at{};
        const char* env = getenv("ACEQD_NO_ZEROCOPY");
        if (!(env && env[0] == '1') && cudaPointerGetAttributes(&at, b->out) == cudaSuccess &&
            at.type == cudaMemoryTypeHost && at.devicePointer) {
            out_dev = (double*)at.devicePointer;
            zero_copy = true;
        } else {
            cudaGetLastError();
            if ((rc = c->out.reserve((size_t)b->out_elems * 16))) return rc;
            out_dev = (double*)c->out.p;
        }
    }

    StepParams sp{};
    sp.pt = pt->d;
    sp.prob = pd;
    sp.trajs = (const aceqd_traj*)c->trajs.p;
    sp.W = (const double*)c->W.p;
    sp.OV = (const double*)c->OV.p;
    sp.ovr_base = c->n_seq_entries;
    sp.rho0s = (const double*)c->rho0s.p;
    sp.snap_steps = (const int*)c->snap_steps.p;
    sp.snaps = (double*)c->snaps.p;
    sp.snap_r = (double*)c->snap_r.p;
    sp.out = out_dev;
    sp.ticks = c->ticks;
    if (const char* tc = getenv("ACEQD_TICK_CLUSTER"))   // debug clock only for launches of this cluster size
        if (atoi(tc) != (b->cluster <= 1 ? 1 : b->cluster)) sp.ticks = nullptr;

    if (b->kernel == 1) {
        sp.T = 1;
        sp.n_tiles = b->n_traj;
        if ((rc = c->scratch.reserve((size_t)b->n_traj * 2 * pd.NL * chi_pad * 16))) return rc;
        ACEQD_CUDA(cudaEventRecord(c->ev[0], c->stream));
        if ((rc = launch_step_check(sp, (double*)c->scratch.p, c->stream, &c->log))) return rc;
        ACEQD_CUDA(cudaEventRecord(c->ev[1], c->stream));
    } else {
        for (long long i = 0; i < (long long)b->n_tiles * T; ++i)
            if (b->tile_traj[i] >= b->n_traj) {
                set_error("batch: tile list references trajectory %d", b->tile_traj[i]);
                return ACEQD_ERR_ARG;
            }
        int cluster = b->cluster <= 1 ? 1 : b->cluster;
        if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8 && cluster != 16) {
            set_error("batch: cluster must be 0/1, 2, 4, 8 or 16");
            return ACEQD_ERR_ARG;
        }
        // ---- small-bond regime of two-level sweeps: one warp per 8 trajectories, PT resident in shared memory.
        // kernel 4 forces it (error if the batch is not eligible), kernel 5 forces the tile kernel; kernel 0 takes it
        // for batches large enough to fill the sub-partitions (one warp per octet: a small batch is latency-bound
        // there and runs faster on tiles shared by clusters).
        {
            const char* env = getenv("ACEQD_SMALL");
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
            const bool eligible = pd.NL == 4 && chi_pad <= 32 && b->n_snap_steps == 0 && pd.n_out <= 16 && pd.NLp4 == 4 &&
                small_smem_bytes(pt->blob_doubles, pt->d.n_slices, chi_pad, pd.n_out, 1) <= (size_t)SMEM_BUDGET;
            bool take = eligible && cluster == 1 && b->n_traj >= ACEQD_SMALL_MIN_TRAJ;
            if (env && env[0] == '0') take = false;
            if (env && env[0] == '1') take = eligible && cluster == 1;
            if (b->kernel == 5) take = false;
            if (b->kernel == 4) {
                if (!eligible) {
                    set_error("batch: the small-bond kernel needs NL == 4, chi_pad <= 32, n_out <= 16 and no snapshot "
                              "requests (NL=%d chi_pad=%d n_out=%d snapshots=%d)", pd.NL, chi_pad, pd.n_out, b->n_snap_steps);
                    return ACEQD_ERR_CAPACITY;
                }
                take = true;
            }
            if (take) {
                std::vector<int32_t> oct;
                oct.reserve((size_t)b->n_tiles * T + 8);
                for (long long i = 0; i < (long long)b->n_tiles * T; ++i)
                    if (b->tile_traj[i] >= 0) oct.push_back(b->tile_traj[i]);
                const int n_oct = (int)((oct.size() + 7) / 8);
                oct.resize((size_t)n_oct * 8, -1);
                // fewest warps per CTA for which every CTA is resident at once (the PT copy is per CTA); else 4
                int wpc = 4;
                for (int w : {1, 2, 4, 8}) {
                    const size_t sm = small_smem_bytes(pt->blob_doubles, pt->d.n_slices, chi_pad, pd.n_out, w);
                    if (sm > (size_t)SMEM_BUDGET) break;
                    const long long per_sm = std::min<long long>((long long)((size_t)SMEM_BUDGET / sm), 2048 / (32 * w));
                    if ((long long)(n_oct + w - 1) / w <= per_sm * n_sm) {
                        wpc = w;
                        break;
                    }
                    wpc = w;
                }
                UP(c->octets, oct.data(), oct.size() * sizeof(int32_t));
                sp.T = 8;
                sp.n_tiles = n_oct;
                sp.tile_traj = (const int*)c->octets.p;
                sp.cluster = 1;
                const size_t smem = small_smem_bytes(pt->blob_doubles, pt->d.n_slices, chi_pad, pd.n_out, wpc);
                if (c->split_ops_pending) {      // operators of a planned wave split: all of them are needed here
                    ACEQD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_built, 0));
                    c->split_ops_pending = false;
                }
                c->split_tiles = 0;
                ACEQD_CUDA(cudaEventRecord(c->ev[0], c->stream));
                if ((rc = launch_step_small(sp, pt->blob_doubles, wpc, smem, c->stream, &c->log))) return rc;
                ACEQD_CUDA(cudaEventRecord(c->ev[1], c->stream));
                c->have_step = true;
                if ((rc = finish_outputs(c, b, pd.n_out, out_dev, zero_copy, reduce))) return rc;
                return ACEQD_OK;
            }
        }
        if (b->kernel == 3) {
            // ---- split-K cluster kernel: `cluster` CTAs hold chi_pad/cluster bond columns each of T trajectories
            int stages = 0;
            size_t smem = 0;
            for (int st = MAX_STAGES; st >= 2; --st) {
                smem = splitk_smem_bytes(pd.NL, chi_pad, T, cluster, st);
                if (smem && smem <= (size_t)SMEM_BUDGET) {
                    stages = st;
                    break;
                }
            }
            if (!stages) {
                set_error("split-K kernel: NL=%d chi_pad=%d T=%d cluster=%d is not supported or does not fit %d B of "
                          "shared memory", pd.NL, chi_pad, T, cluster, SMEM_BUDGET);
                return ACEQD_ERR_CAPACITY;
            }
            std::vector<PassDesc> passes;
            if ((rc = build_passes(prob, T, 1, passes))) return rc;
            UP(c->passes, passes.data(), passes.size() * sizeof(PassDesc));
            UP(c->tiles, b->tile_traj, (size_t)b->n_tiles * T * sizeof(int32_t));
            sp.T = T;
            sp.n_pass = (int)passes.size();
            sp.stages = stages;
            sp.n_tiles = b->n_tiles;
            sp.cluster = cluster;
            sp.NR = splitk_columns(chi_pad, cluster);
            sp.rslots = chi_pad > PANEL ? 1 : 2;   // with panels the two panels of a row block alternate on one slot
            sp.passes = (const PassDesc*)c->passes.p;
            sp.tile_traj = (const int*)c->tiles.p;
            if (c->split_ops_pending) {
                ACEQD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_built, 0));
                c->split_ops_pending = false;
            }
            c->split_tiles = 0;
            ACEQD_CUDA(cudaEventRecord(c->ev[0], c->stream));
            if ((rc = launch_step_splitk(sp, smem, c->stream, &c->log))) return rc;
            ACEQD_CUDA(cudaEventRecord(c->ev[1], c->stream));
            c->have_step = true;
            if ((rc = finish_outputs(c, b, pd.n_out, out_dev, zero_copy, reduce))) return rc;
            return ACEQD_OK;
        }
        std::vector<PassDesc> passes;
        if ((rc = build_passes(prob, T, cluster, passes))) return rc;
        UP(c->passes, passes.data(), passes.size() * sizeof(PassDesc));
        UP(c->tiles, b->tile_traj, (size_t)b->n_tiles * T * sizeof(int32_t));
        // prefer staging the per-row operators in shared memory (hides their DRAM latency) if at
        // least 3 chunk stages still fit; otherwise read them from global memory
        // prefer staging the per-row operators in shared memory (hides their DRAM latency): double
        // buffered if >= 3 chunk stages still fit, else single buffered (refilled during phase C),
        // else read them from global memory
        const int wov_full = pd.w_doubles + pd.ov_doubles;
        int stages = 0, wov = 0, wbufs = 1;
        struct Cand { int wov, bufs, min_stages; };
        int wstage_max = 2;      // debug/tuning: ACEQD_WSTAGE=0 reads the operators from global memory, 1 forbids double buffering
        if (const char* ws = getenv("ACEQD_WSTAGE")) wstage_max = atoi(ws);
        for (const Cand cd : {Cand{wov_full, 2, 3}, Cand{wov_full, 1, 3}, Cand{wov_full, 1, 2}, Cand{0, 1, 2}}) {
            if ((cd.wov ? cd.bufs : 0) > wstage_max) continue;
            for (int st = MAX_STAGES; st >= cd.min_stages; --st)
                if (step_smem_bytes(pd.NL, chi_pad, T, st, cd.wov, cd.bufs) <= (size_t)SMEM_BUDGET) {
                    stages = st;
                    wov = cd.wov;
                    wbufs = cd.bufs;
                    break;
                }
            if (stages) break;
        }
        if (stages < 2) {
            // last resort: no chunk ring at all, PT fragments read from global memory / L2 (gemm_pass_global)
            stages = wov = 0;
            wbufs = 1;
            if (step_smem_bytes(pd.NL, chi_pad, T, 0, 0, 1) > (size_t)SMEM_BUDGET) {
                set_error("NL=%d chi_pad=%d T=%d does not fit %d B of shared memory", pd.NL, chi_pad,
                          T, SMEM_BUDGET);
                return ACEQD_ERR_CAPACITY;
            }
        }
        sp.T = T;
        sp.n_pass = (int)passes.size();
        sp.stages = stages;
        sp.wov_doubles = wov;
        sp.wbufs = wbufs;
        sp.n_tiles = b->n_tiles;
        sp.cluster = cluster;
        sp.passes = (const PassDesc*)c->passes.p;
        sp.tile_traj = (const int*)c->tiles.p;
        const size_t smem = step_smem_bytes(pd.NL, chi_pad, T, stages, wov, wbufs);
        if (cluster > 8 && (step_max_active_clusters(sp, smem) < 1 || getenv("ACEQD_CLUSTER16_UNSCHEDULABLE"))) {   // (env: tests)
            // 16 CTAs are beyond the portable cluster size: where the device (a partition of it, a smaller part) cannot
            // place such a cluster the tile runs on 8 CTAs instead; the launch name says which
            cluster = 8;
            if ((rc = build_passes(prob, T, cluster, passes))) return rc;
            UP(c->passes, passes.data(), passes.size() * sizeof(PassDesc));
            sp.n_pass = (int)passes.size();
            sp.cluster = cluster;
            sp.passes = (const PassDesc*)c->passes.p;
        }
        if (use_segments(c, b)) {
            int n_slots = 0;
            const int n_sm = segment_ctas(c);
            std::vector<SegDesc> segs;
            std::vector<int> seg_off;
            if (plan_segments(b, n_sm, segs, seg_off, n_slots)) {
                UP(c->segs, segs.data(), segs.size() * sizeof(SegDesc));
                UP(c->seg_off, seg_off.data(), seg_off.size() * sizeof(int));
                const size_t slot = step_seg_slot_doubles(pd.NL, chi_pad, T);
                if ((rc = c->seg_state.reserve(std::max<size_t>(1, n_slots) * slot * 8))) return rc;
                if ((rc = c->seg_flags.reserve(std::max<size_t>(1, n_slots) * sizeof(unsigned)))) return rc;
                ACEQD_CUDA(cudaMemsetAsync(c->seg_flags.p, 0, std::max<size_t>(1, n_slots) * sizeof(unsigned), c->stream));
                sp.segs = (const SegDesc*)c->segs.p;
                sp.seg_off = (const int*)c->seg_off.p;
                sp.seg_state = (double*)c->seg_state.p;
                sp.seg_slot_doubles = slot;
                sp.seg_flags = (unsigned*)c->seg_flags.p;
                sp.seg_epoch = 1u;
                sp.n_ctas = (int)seg_off.size() - 1;
            }
        }
        // More tiles than SMs: the persistent CTAs run in waves.  aceqd_propagate_batch may have planned a
        // split (first the partial wave, then the full waves): the later waves' operators are built on a
        // side stream meanwhile, and on the end-to-end path the first wave's outputs are copied to the
        // host while the later waves compute.
        ACEQD_CUDA(cudaEventRecord(c->ev[0], c->stream));
        if (c->split_tiles > 0 && c->split_tiles < b->n_tiles && cluster == 1) {
            const int n_a = c->split_tiles;
            StepParams sa = sp, sb2 = sp;
            sa.n_tiles = n_a;
            sb2.n_tiles = b->n_tiles - n_a;
            sb2.tile_traj = sp.tile_traj + (size_t)n_a * T;
            if ((rc = launch_step_dmma(sa, smem, c->stream, &c->log))) return rc;
            if (c->split_ops_pending) {
                ACEQD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_built, 0));
                c->split_ops_pending = false;
            }
            const bool copy_early = !b->device_resident && !zero_copy && c->split_out > 0;
            if (copy_early) {
                if (!c->copy_stream) {
                    ACEQD_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
                    ACEQD_CUDA(cudaEventCreateWithFlags(&c->ev_wave, cudaEventDisableTiming));
                }
                ACEQD_CUDA(cudaEventRecord(c->ev_wave, c->stream));
            }
            if ((rc = launch_step_dmma(sb2, smem, c->stream, &c->log))) return rc;
            ACEQD_CUDA(cudaEventRecord(c->ev[1], c->stream));
            c->have_step = true;
            c->split_tiles = 0;
            if (copy_early) {
                ACEQD_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_wave, 0));
                ACEQD_CUDA(cudaMemcpyAsync(b->out, out_dev, (size_t)c->split_out * 16, cudaMemcpyDeviceToHost,
                                           c->copy_stream));
                ACEQD_CUDA(cudaMemcpyAsync(b->out + 2 * c->split_out, out_dev + 2 * c->split_out,
                                           (size_t)(b->out_elems - c->split_out) * 16, cudaMemcpyDeviceToHost,
                                           c->stream));
                ACEQD_CUDA(cudaStreamSynchronize(c->stream));
                ACEQD_CUDA(cudaStreamSynchronize(c->copy_stream));
                return ACEQD_OK;
            }
            if ((rc = finish_outputs(c, b, pd.n_out, out_dev, zero_copy, reduce))) return rc;
            return ACEQD_OK;
        }
        if (c->split_ops_pending) {      // a planned split that this launch does not use: just order the streams
            ACEQD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_built, 0));
            c->split_ops_pending = false;
        }
        if ((rc = launch_step_dmma(sp, smem, c->stream, &c->log))) return rc;
        ACEQD_CUDA(cudaEventRecord(c->ev[1], c->stream));
    }
    c->have_step = true;
    if ((rc = finish_outputs(c, b, pd.n_out, out_dev, zero_copy, reduce))) return rc;
    return ACEQD_OK;
}

// Decide whether the batch runs as "partial wave first, full waves after": needs more tiles than SMs, no
// explicit (MTO) entries, and the first group's operator entries and output rows to lie entirely before the
// second group's (true for sweeps, whose trajectories are tiled in order).
static void plan_wave_split(aceqd_ctx* c, const aceqd_problem* prob, const aceqd_batch* b) {
    c->split_tiles = 0;
    c->split_entries = c->split_out = 0;
    if (!c || !prob || !b || (b->kernel != 0 && b->kernel != 5) || b->cluster > 1 || b->n_entries != 0 || !b->tile_traj ||
        b->tile_T < 1 || !b->trajs)
        return;
    if (use_segments(c, b)) return;   // one balanced launch instead of waves
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
    if (b->n_tiles <= n_sm) return;
    const int T = b->tile_T;
    const int n_a = b->n_tiles - (b->n_tiles - 1) / n_sm * n_sm;      // the partial wave goes first
    if (n_a <= 0 || n_a >= b->n_tiles) return;
    long long ent_a = 0, ent_b = -1, out_a = 0, out_b = -1;
    for (long long i = 0; i < (long long)b->n_tiles * T; ++i) {
        const int idx = b->tile_traj[i];
        if (idx < 0 || idx >= b->n_traj) continue;
        const aceqd_traj& t = b->trajs[idx];
        if (t.n_ovr != 0 || t.n_steps < 0) return;
        const long long e0 = t.ent0, e1 = t.ent0 + t.n_steps + 1;
        const long long o0 = t.out_off, o1 = o0 + (long long)(t.n_steps + 1 - t.out_from) * prob->d.n_out;
        if (i < (long long)n_a * T) {
            ent_a = std::max(ent_a, e1);
            out_a = std::max(out_a, o1);
        } else {
            ent_b = ent_b < 0 ? e0 : std::min(ent_b, e0);
            out_b = out_b < 0 ? o0 : std::min(out_b, o0);
        }
    }
    if (ent_b < 0 || ent_a > ent_b) return;
    c->split_tiles = n_a;
    c->split_entries = ent_a;
    c->split_out = (out_b >= 0 && out_a <= out_b) ? out_a : 0;
}

int aceqd_propagate_batch(aceqd_ctx* c, const aceqd_problem* prob, const aceqd_pt* pt,
                          const aceqd_batch* b) {
    if (c && check_batch(prob, b) == ACEQD_OK) plan_wave_split(c, prob, b);
    int rc = aceqd_build_operators(c, prob, b);
    if (rc) {
        if (c) c->split_tiles = 0;
        return rc;
    }
    rc = aceqd_run_steps(c, prob, pt, b);
    if (c) {
        c->split_tiles = 0;
        c->split_entries = 0;
    }
    return rc;
}

int aceqd_snapshot_read(aceqd_ctx* c, int slot, int NL, int chi_pad, double* host_out) {
    if (!c || !host_out || slot < 0 || NL <= 0 || chi_pad <= 0) return ACEQD_ERR_ARG;
    const size_t bytes = (size_t)NL * chi_pad * 16;
    if ((size_t)(slot + 1) * bytes > c->snaps.cap) {
        set_error("snapshot slot %d not allocated", slot);
        return ACEQD_ERR_ARG;
    }
    ACEQD_CUDA(cudaSetDevice(c->device));
    ACEQD_CUDA(cudaMemcpyAsync(host_out, (char*)c->snaps.p + (size_t)slot * bytes, bytes,
                               cudaMemcpyDeviceToHost, c->stream));
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    return ACEQD_OK;
}

int aceqd_expm_batch(aceqd_ctx* c, int n, int count, const double* a_host, double* out_host) {
    if (!c || n <= 0 || count < 0 || !a_host || !out_host) return ACEQD_ERR_ARG;
    ACEQD_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)count * n * n * 16;
    int rc;
    if ((rc = c->misc.reserve(2 * bytes + 32))) return rc;
    double* a_dev = (double*)c->misc.p;
    double* o_dev = (double*)((char*)c->misc.p + (bytes + 15) / 16 * 16);
    ACEQD_CUDA(cudaMemcpyAsync(a_dev, a_host, bytes, cudaMemcpyHostToDevice, c->stream));
    double* scratch = nullptr;
    {
        int ctas = 0;
        const size_t need = opbuild_scratch_bytes(n, &ctas);   // sized for the larger operator-builder footprint
        if (need) {
            if ((rc = c->opscratch.reserve(need))) return rc;
            scratch = (double*)c->opscratch.p;
        }
    }
    if ((rc = launch_expm_batch(n, count, a_dev, o_dev, scratch, c->stream, &c->log))) return rc;
    ACEQD_CUDA(cudaMemcpyAsync(out_host, o_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    return ACEQD_OK;
}

int aceqd_tlmap_run(aceqd_ctx* c, int NL, int n_mats, const double* mats, int n_chains,
                    const double* v0, const int64_t* seg_off, int64_t n_segs, const aceqd_tlseg* segs,
                    int n_w, const double* w, int n_emit_max, double* out, double* final_v) {
    if (!c || NL <= 0 || n_mats <= 0 || !mats || n_chains < 0 || !v0 || !seg_off || n_segs < 0 ||
        (n_segs && !segs) || n_w < 0 || (n_w && !w) || n_emit_max < 0 || (!out && !final_v)) {
        set_error("aceqd_tlmap_run: invalid argument");
        return ACEQD_ERR_ARG;
    }
    if (n_chains == 0) return ACEQD_OK;
    if (seg_off[0] != 0 || seg_off[n_chains] != n_segs) {
        set_error("aceqd_tlmap_run: seg_off must run from 0 to n_segs");
        return ACEQD_ERR_ARG;
    }
    for (int i = 0; i < n_chains; ++i)
        if (seg_off[i + 1] < seg_off[i]) {
            set_error("aceqd_tlmap_run: seg_off not monotone at chain %d", i);
            return ACEQD_ERR_ARG;
        }
    for (int64_t s = 0; s < n_segs; ++s)
        if (segs[s].start < 0 || segs[s].count < 0 || segs[s].stride < 0 || segs[s].stride > 1 ||
            (long long)segs[s].start + (long long)(segs[s].count > 0 ? segs[s].count - 1 : 0) * segs[s].stride >= n_mats) {
            set_error("aceqd_tlmap_run: segment %lld outside the matrix pool", (long long)s);
            return ACEQD_ERR_ARG;
        }
    ACEQD_CUDA(cudaSetDevice(c->device));
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    const size_t vec = (size_t)NL * 16;
    UP(c->tl_pool, mats, (size_t)n_mats * NL * vec);
    UP(c->tl_v0, v0, (size_t)n_chains * vec);
    UP(c->tl_segoff, seg_off, (size_t)(n_chains + 1) * sizeof(int64_t));
    UP(c->tl_segs, segs, (size_t)n_segs * sizeof(aceqd_tlseg));
    UP(c->tl_w, w, (size_t)n_w * vec);
    int rc;
    const size_t out_bytes = (size_t)n_chains * n_emit_max * n_w * 16;
    if (out && (rc = c->tl_out.reserve(out_bytes ? out_bytes : 16))) return rc;
    if (final_v && (rc = c->tl_final.reserve((size_t)n_chains * vec))) return rc;
    if (out && out_bytes) ACEQD_CUDA(cudaMemsetAsync(c->tl_out.p, 0, out_bytes, c->stream));
    ACEQD_CUDA(cudaEventRecord(c->ev[4], c->stream));
    if ((rc = launch_tlmap(NL, n_chains, n_w, n_emit_max, (const double*)c->tl_pool.p,
                           (const double*)c->tl_v0.p, (const long long*)c->tl_segoff.p,
                           (const aceqd_tlseg*)c->tl_segs.p, (const double*)c->tl_w.p,
                           out ? (double*)c->tl_out.p : nullptr,
                           final_v ? (double*)c->tl_final.p : nullptr, c->stream, &c->log)))
        return rc;
    ACEQD_CUDA(cudaEventRecord(c->ev[5], c->stream));
    c->have_tl = true;
    if (out && out_bytes)
        ACEQD_CUDA(cudaMemcpyAsync(out, c->tl_out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (final_v)
        ACEQD_CUDA(cudaMemcpyAsync(final_v, c->tl_final.p, (size_t)n_chains * vec,
                                   cudaMemcpyDeviceToHost, c->stream));
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    return ACEQD_OK;
}

int aceqd_tlmap_last_ms(aceqd_ctx* c, float* ms) {
    if (!c || !ms) return ACEQD_ERR_ARG;
    *ms = 0.f;
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    if (c->have_tl) ACEQD_CUDA(cudaEventElapsedTime(ms, c->ev[4], c->ev[5]));
    return ACEQD_OK;
}

int aceqd_host_alloc(size_t bytes, void** out) {
    if (!out) return ACEQD_ERR_ARG;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return ACEQD_ERR_NOMEM;
    }
    return ACEQD_OK;
}

void aceqd_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

void aceqd_struct_sizes(int32_t out[4]) {
    out[0] = (int32_t)sizeof(aceqd_seq);
    out[1] = (int32_t)sizeof(aceqd_entry);
    out[2] = (int32_t)sizeof(aceqd_traj);
    out[3] = (int32_t)sizeof(aceqd_batch);
}

int aceqd_fp64_peak(aceqd_ctx* c, int kind, int iters, double* tflops) {
    if (!c || !tflops || iters <= 0) return ACEQD_ERR_ARG;
    ACEQD_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = c->misc.reserve(64))) return rc;
    int blocks = 0, threads = 0;
    cudaEvent_t e0, e1;
    ACEQD_CUDA(cudaEventCreate(&e0));
    ACEQD_CUDA(cudaEventCreate(&e1));
    // warm-up
    if ((rc = launch_fp64_peak(kind, iters / 8 + 1, (double*)c->misc.p, &blocks, &threads,
                               c->stream, &c->log)))
        return rc;
    ACEQD_CUDA(cudaEventRecord(e0, c->stream));
    if ((rc = launch_fp64_peak(kind, iters, (double*)c->misc.p, &blocks, &threads, c->stream,
                               &c->log)))
        return rc;
    ACEQD_CUDA(cudaEventRecord(e1, c->stream));
    ACEQD_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    ACEQD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double warps = (double)blocks * threads / 32.0;
    double flops;
    if (kind == 0 || kind >= 10)
        flops = warps * (double)iters * 8.0 * 512.0;  // 8 DMMA.8x8x4 per iteration per warp
    else
        flops = (double)blocks * threads * (double)iters * 16.0 * 2.0;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return ACEQD_OK;
}

}  // extern "C"
