// llama2_b200.cu — C ABI (include/llama2_b200.h) over the sm_100a kernels in l2b_device.cuh.
//
// Replaces transformer() of the reference (src/main.zig:285-430) and the device-side
// equivalents of Weights.init (:73-115) / RunState.init (:137-154).  Pure CUDA runtime:
// no torch, no CPU fallback.  One context = one GPU = one stream; a decode step is a CUDA
// graph of 5 kernels per layer + classifier, replayed per token.
#include "../../include/llama2_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>   // types only; the library is dlopen()ed on first multi-GPU use
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "l2b_device.cuh"
#include "l2b_mega.cuh"

using namespace l2b;

// ---------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU use never needs libnccl to be present.
// ---------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool nccl_load(std::string *err) {
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        if (err) *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return false;
    }
#define L2B_SYM(field, name)                                                        \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);                         \
    if (!g_nccl.field) {                                                            \
        if (err) *err = std::string("dlsym failed: ") + name;                       \
        return false;                                                               \
    }
    L2B_SYM(GetUniqueId, "ncclGetUniqueId");
    L2B_SYM(CommInitRank, "ncclCommInitRank");
    L2B_SYM(CommDestroy, "ncclCommDestroy");
    L2B_SYM(AllReduce, "ncclAllReduce");
    L2B_SYM(AllGather, "ncclAllGather");
    L2B_SYM(GetErrorString, "ncclGetErrorString");
#undef L2B_SYM
    g_nccl.ok = true;
    return true;
}
}  // namespace

// ---------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------
struct l2b_ctx {
    l2b_config cfg{};
    int rank = 0, world = 1, device = 0, num_sms = 0;
    // derived sizes (local = this rank's shard)
    int dim = 0, hidden = 0, head_size = 0, kv_mul = 0;
    int q_dim = 0, kv_dim = 0;              // global
    int q_loc = 0, kv_loc = 0, hid_loc = 0, heads_loc = 0, vocab_loc = 0;
    // device weights (Weights, src/main.zig:53-71), this rank's slices
    float *emb = nullptr, *rms_att = nullptr, *rms_ffn = nullptr, *rms_final = nullptr;
    float *wq = nullptr, *wk = nullptr, *wv = nullptr, *wo = nullptr;
    float *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *wcls = nullptr;
    bool wcls_owned = false;
    // device run state (RunState, :119-135)
    float *X[2] = {nullptr, nullptr};       // residual stream, ping-pong
    float *delta_a = nullptr, *delta_f = nullptr;  // pending residual from wo / w2 (all-reduced in TP)
    float *q = nullptr, *xb = nullptr, *hb = nullptr, *logits = nullptr, *logits_loc = nullptr;
    float *kcache = nullptr, *vcache = nullptr;     // (L, seq_len, kv_loc)
    float *rope_cos = nullptr, *rope_sin = nullptr; // (seq_len, head_size/2)
    float *part_o = nullptr, *part_ml = nullptr;
    unsigned int *counters = nullptr;
    int *ctl = nullptr;
    unsigned long long *amax = nullptr;
    int *gen_forced = nullptr, *gen_out = nullptr, *gen_ndone = nullptr;
    int final_x = 0;                        // which X[] holds x after the step
    // host pinned
    float *h_logits = nullptr;
    int *h_ints = nullptr;                  // [0]=next, [1]=n_done
    int *h_gen = nullptr;                   // seq_len ints
    // attention launch shape
    int nsplit = 1, min_chunk = 256, attn_smem = 0;   // timeline splits of >= 256 positions (A/B: 64 -> 256 = +14% on stories15M)
    // streams / graphs
    cudaStream_t stream = nullptr;
    cudaGraphExec_t graph_logits = nullptr, graph_argmax = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool use_graphs = true;
    bool use_pdl = true;
    bool use_mega = false;                   // one persistent cooperative kernel per step
    unsigned long long *gbar = nullptr;      // its grid-barrier counter
    int mega_nstage = 0, mega_xs_floats = 0, mega_smem = 0;
    int last_grid = 0;                       // grid of the most recent GEMV launch
    // fused all-reduce over peer memory (world > 1): exchange buffer [slots][world][dim] + counters
    bool use_p2p = false;
    float *xchg = nullptr;
    unsigned int *xflags = nullptr;
    float *peer_xchg[MAX_TP] = {};
    unsigned int *peer_flags[MAX_TP] = {};
    std::vector<void *> ipc_opened;
    std::vector<int> xgrid;                  // producer CTAs per reduce point (same on every rank)
    bool attn_flash = true;                  // flash-decoding attention (false: 3-pass kernel)
    unsigned long long *trace = nullptr;     // L2B_TRACE=1: [launch][TRACE_MAX_CTAS][TRACE_SLOTS] timeline
    int trace_launches = 0;
    int pf_bytes = 0;                        // per-CTA L2 prefetch of the next GEMV's first rows (L2B_PF_KB; measured: no gain)
    int tma_ctas_per_sm = 1;                 // CTAs of ONE TMA kernel per SM (the other half-SM is for its successor)
    int tma_stages = 0;                      // 0 = auto (two CTAs per SM); else forced ring depth
    bool big_kernel_tma = true;              // bandwidth-bound GEMVs: TMA-ring kernel (false: register-fed 8-row kernel)
    long long gemv8_min_bytes = 8ll << 20;   // >= this many weight bytes (and n >= 1024): 8-row kernel; -1 = never
    int launches_per_step = 0;
    // comm
    ncclComm_t comm = nullptr;
    // bookkeeping
    int n_appended = 0;
    float last_ms = 0.0f;
    int last_launches = 0;
    std::string err;
    std::vector<void *> owned;              // device allocations to free
    // per-kernel profiling (l2b_profile_step)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<l2b_kernel_time> prof_rec;
};

namespace {

const char *kNoError = "";
thread_local std::string g_create_error;

#define L2B_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            char buf__[512];                                                                  \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call,                     \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                            \
            (ctx)->err = buf__;                                                               \
            return (e__ == cudaErrorMemoryAllocation) ? L2B_ERR_OOM : L2B_ERR_CUDA;           \
        }                                                                                     \
    } while (0)

#define L2B_NCCL(ctx, call)                                                                   \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            char buf__[512];                                                                  \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call,                     \
                     g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?", __FILE__,      \
                     __LINE__);                                                               \
            (ctx)->err = buf__;                                                               \
            return L2B_ERR_COMM;                                                              \
        }                                                                                     \
    } while (0)

int fail(l2b_ctx *ctx, int code, const char *msg) {
    if (ctx) ctx->err = msg;
    return code;
}

template <typename T>
int dev_alloc(l2b_ctx *ctx, T **p, size_t count) {
    void *q = nullptr;
    L2B_CUDA(ctx, cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
    ctx->owned.push_back(q);
    *p = static_cast<T *>(q);
    return L2B_OK;
}

// ---- shape validation (the reference's asserts, src/main.zig:433-434,:534-540,:658-661) ----
int validate_config(const l2b_config *c, int world, std::string *why) {
    auto bad = [&](const char *m) { *why = m; return L2B_ERR_UNSUPPORTED; };
    if (c->dim <= 0 || c->hidden_dim <= 0 || c->n_layers <= 0 || c->n_heads <= 0 ||
        c->n_kv_heads <= 0 || c->vocab_size <= 0 || c->seq_len <= 0) {
        *why = "config fields must be positive";
        return L2B_ERR_INVALID_ARG;
    }
    if (c->dim % c->n_heads) return bad("dim % n_heads != 0");
    if (c->n_heads % c->n_kv_heads) return bad("n_heads % n_kv_heads != 0");
    const int hs = c->dim / c->n_heads;
    if (c->dim % 4 || c->hidden_dim % 4) return bad("dim and hidden_dim must be multiples of 4");
    if (hs % 4) return bad("head_size must be a multiple of 4");
    if (hs / 4 > NT) return bad("head_size too large");
    if (world != 1 && world != 2 && world != 4 && world != 8) return bad("world_size must be 1, 2, 4 or 8");
    if (c->n_kv_heads % world) return bad("n_kv_heads % world_size != 0");
    if ((c->hidden_dim / world) % 4 || c->hidden_dim % world) return bad("hidden_dim/world_size must be a multiple of 4");
    if (c->vocab_size % world) return bad("vocab_size % world_size != 0");
    return L2B_OK;
}

uint64_t checkpoint_floats(const l2b_config *c) {
    const uint64_t dim = c->dim, hid = c->hidden_dim, L = c->n_layers, V = c->vocab_size, S = c->seq_len;
    const uint64_t hs = dim / c->n_heads, kvd = hs * c->n_kv_heads;
    uint64_t n = V * dim + L * dim + L * dim * dim + 2 * L * dim * kvd + L * dim * dim + L * dim +
                 3 * L * dim * hid + dim + 2 * (S * hs / 2);
    if (!c->shared_weights) n += V * dim;
    return n;
}

// ---- tensor materialisation: upload a (possibly sharded) window, or synthesise it in place ----
enum ShardMode { SH_NONE, SH_ROWS, SH_COLS };
struct Dist { double mean, sigma; float lo, hi; };
constexpr double kSynthInvStd = 2.6428997921303014e-05;  // 1 / sqrt(4*(65536^2-1)/12)

struct Source {
    const float *host = nullptr;   // checkpoint payload or nullptr => synthetic
    uint64_t seed = 0;
};

int materialize(l2b_ctx *ctx, float **dptr, const Source &src, uint64_t payload_off, int tensor_id,
                uint64_t L, uint64_t rows, uint64_t cols, ShardMode mode, Dist dist) {
    const uint64_t g = ctx->world, r = ctx->rank;
    const uint64_t rows_loc = (mode == SH_ROWS) ? rows / g : rows;
    const uint64_t cols_loc = (mode == SH_COLS) ? cols / g : cols;
    const uint64_t row0 = (mode == SH_ROWS) ? r * rows_loc : 0;
    const uint64_t col0 = (mode == SH_COLS) ? r * cols_loc : 0;
    int rc = dev_alloc(ctx, dptr, L * rows_loc * cols_loc);
    if (rc) return rc;
    const uint64_t tensor_seed = mix64(src.seed * 1000003ull + (uint64_t)tensor_id);
    const bool whole = (rows_loc == rows && cols_loc == cols);
    const uint64_t nl = whole ? 1 : L;            // one shot when nothing is cut
    const uint64_t rr = whole ? L * rows : rows_loc;
    for (uint64_t l = 0; l < nl; ++l) {
        float *dst = *dptr + l * rows_loc * cols_loc;
        const uint64_t first = l * rows * cols + row0 * cols + col0;
        if (src.host && whole) {
            L2B_CUDA(ctx, cudaMemcpy(dst, src.host + payload_off, L * rows * cols * sizeof(float),
                                     cudaMemcpyHostToDevice));
        } else if (src.host) {
            L2B_CUDA(ctx, cudaMemcpy2D(dst, cols_loc * sizeof(float), src.host + payload_off + first,
                                       cols * sizeof(float), cols_loc * sizeof(float), rr,
                                       cudaMemcpyHostToDevice));
        } else {
            const uint64_t total = rr * cols_loc;
            int blocks = (int)((total + 255) / 256 < (uint64_t)(ctx->num_sms * 16)
                                   ? (total + 255) / 256
                                   : (uint64_t)(ctx->num_sms * 16));
            if (blocks < 1) blocks = 1;
            synth_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(dst, rr, cols_loc, first, cols, tensor_seed,
                                                               dist.mean, dist.sigma * kSynthInvStd,
                                                               dist.lo, dist.hi);
            L2B_CUDA(ctx, cudaGetLastError());
        }
    }
    return L2B_OK;
}

// ---- GEMV launch ---------------------------------------------------------------------------
typedef void (*gemv_fn)(const GemvParams);

template <int EPI>
gemv_fn gemv_pick(int tpr) {
    switch (tpr) {
    case 8: return gemv_kernel<8, EPI>;
    case 16: return gemv_kernel<16, EPI>;
    case 32: return gemv_kernel<32, EPI>;
    case 64: return gemv_kernel<64, EPI>;
    case 128: return gemv_kernel<128, EPI>;
    default: return gemv_kernel<256, EPI>;
    }
}
gemv_fn gemv_pick(int epi, int tpr) {
    switch (epi) {
    case EPI_XCHG: return gemv_pick<EPI_XCHG>(tpr);
    case EPI_STORE: return gemv_pick<EPI_STORE>(tpr);
    case EPI_ARGMAX: return gemv_pick<EPI_ARGMAX>(tpr);
    case EPI_QKV: return gemv_pick<EPI_QKV>(tpr);
    default: return gemv_pick<EPI_SILU>(tpr);
    }
}

constexpr int kMaxDynSmem = 200 * 1024;
constexpr int kMaxSmemOptin = 227 * 1024 - 2048;   // static __shared__ of the kernels stays under 2 KB

int gemv_tpr(int n) {
    const int n4 = n / 4;
    int tpr = 8;
    while (tpr < 256 && tpr * 2 <= n4 / 2) tpr *= 2;
    return tpr;
}

int prof_mark(l2b_ctx *ctx, const char *name, int layer, uint64_t bytes, cudaStream_t st) {
    if (!ctx->profiling) return L2B_OK;
    cudaEvent_t ev;
    L2B_CUDA(ctx, cudaEventCreate(&ev));
    L2B_CUDA(ctx, cudaEventRecord(ev, st));
    ctx->prof_ev.push_back(ev);
    l2b_kernel_time r{};
    snprintf(r.name, sizeof r.name, "%s", name);
    r.layer = layer;
    r.bytes = bytes;
    ctx->prof_rec.push_back(r);
    return L2B_OK;
}

gemv_fn gemv_tma_pick(int epi) {
    switch (epi) {
    case EPI_XCHG: return gemv_tma_kernel<EPI_XCHG>;
    case EPI_STORE: return gemv_tma_kernel<EPI_STORE>;
    case EPI_ARGMAX: return gemv_tma_kernel<EPI_ARGMAX>;
    case EPI_QKV: return gemv_tma_kernel<EPI_QKV>;
    default: return gemv_tma_kernel<EPI_SILU>;
    }
}

gemv_fn gemv8_pick(int epi) {
    switch (epi) {
    case EPI_XCHG: return gemv8_kernel<EPI_XCHG>;
    case EPI_STORE: return gemv8_kernel<EPI_STORE>;
    case EPI_ARGMAX: return gemv8_kernel<EPI_ARGMAX>;
    case EPI_QKV: return gemv8_kernel<EPI_QKV>;
    default: return gemv8_kernel<EPI_SILU>;
    }
}

int launch_gemv(l2b_ctx *ctx, int epi, const GemvParams &p, cudaStream_t st, const char *name = "gemv",
                int layer = -1) {
    {
        int prc = prof_mark(ctx, name, layer, (uint64_t)p.total_rows * p.n * 4ull, st);
        if (prc) return prc;
    }
    const size_t xbytes = (size_t)p.n * 4 * (1 + (p.delta ? 1 : 0) + (p.gamma ? 1 : 0));
    // bandwidth-bound shapes take the TMA-ring kernel (or the register-fed 8-row kernel when
    // L2B_GEMV_BIG=ldg), latency-bound ones the fine-grained kernel
    const bool big = ctx->gemv8_min_bytes >= 0 && p.n >= 1024 &&
                     (uint64_t)p.total_rows * p.n * 4ull >= (uint64_t)ctx->gemv8_min_bytes;
    // TMA-ring kernel: only x is staged in shared memory, the rest of the SM's 227 KB is the ring
    // (measured: 5-6 stages of 32 KB on ONE CTA per SM beat 2 x 3 stages and beat leaving half
    // the SM to the successor kernel's pre-fill; profiles/r01_tma_ring_variants.md)
    const size_t stage_bytes = (size_t)TMA_STAGE_FLOATS * 4;
    const size_t xonly = (size_t)p.n * 4;
    int nstage = ctx->tma_stages > 0 ? ctx->tma_stages : (int)(((size_t)kMaxSmemOptin - xonly) / stage_bytes);
    if (nstage > TMA_MAX_STAGES) nstage = TMA_MAX_STAGES;
    const bool fused_ok = !(p.delta || p.gamma || p.xparts) || p.n <= 4 * TMA_THREADS * 4;   // register slices of delta/gamma
    const bool tma = big && ctx->big_kernel_tma && nstage >= 2 && fused_ok && p.head_size <= 256 &&
                     (size_t)nstage * stage_bytes + xonly <= (size_t)kMaxSmemOptin;
    const size_t smem = tma ? (size_t)nstage * stage_bytes + xonly : xbytes;
    if (smem > (size_t)kMaxSmemOptin) return fail(ctx, L2B_ERR_UNSUPPORTED, "activation vector too large for shared memory");
    const int tpr = gemv_tpr(p.n);
    gemv_fn fn = tma ? gemv_tma_pick(epi) : big ? gemv8_pick(epi) : gemv_pick(epi, tpr);
    static bool attr_done[5][8][16] = {};
    const int ti = tma ? 7 : big ? 6 : tpr == 8 ? 0 : tpr == 16 ? 1 : tpr == 32 ? 2 : tpr == 64 ? 3 : tpr == 128 ? 4 : 5;
    if (!attr_done[epi][ti][ctx->device & 15]) {
        L2B_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin));
        attr_done[epi][ti][ctx->device & 15] = true;
    }
    const int threads = tma ? TMA_THREADS : NT;
    int occ = 0;
    L2B_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
    if (occ < 1) occ = 1;
    if (tma && occ > ctx->tma_ctas_per_sm) occ = ctx->tma_ctas_per_sm;   // leave room for the next kernel's CTA
    int grid = ctx->num_sms * occ;
    if (big) {
        const int npairs = (p.total_rows + 1) / 2;
        if (grid > npairs) grid = npairs;
    } else {
        const int tile_rows = (NT / tpr) * GEMV_R;
        const int ntiles = (p.total_rows + tile_rows - 1) / tile_rows;
        if (grid > ntiles) grid = ntiles;
    }
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(threads);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = ctx->use_pdl ? 1 : 0;
    GemvParams pp = p;
    pp.nstage = nstage;
    pp.trace = (ctx->trace && !ctx->profiling) ? ctx->trace + (size_t)(ctx->last_launches % ctx->trace_launches) * TRACE_MAX_CTAS * TRACE_SLOTS : nullptr;
    ctx->last_grid = grid;
    L2B_CUDA(ctx, cudaLaunchKernelEx(&lc, fn, pp));
    ++ctx->last_launches;
    return L2B_OK;
}

// flash-decoding form when head_size/4 splits into 1..8 float4 per lane, else the 3-pass kernel
typedef void (*attn_fn)(const AttnParams);
attn_fn pick_attention(int head_size, bool flash, size_t *smem) {
    const int hs4 = head_size / 4;
    const int lpr = (hs4 % 8 == 0) ? 8 : (hs4 % 4 == 0) ? 4 : (hs4 % 2 == 0) ? 2 : 1;
    const int nf = hs4 / lpr;
    attn_fn f2 = nullptr;
    if (flash) {
        switch (nf) {
        case 1: f2 = attention_flash_kernel<1>; break;
        case 2: f2 = attention_flash_kernel<2>; break;
        case 3: f2 = attention_flash_kernel<3>; break;
        case 4: f2 = attention_flash_kernel<4>; break;
        case 5: f2 = attention_flash_kernel<5>; break;
        case 6: f2 = attention_flash_kernel<6>; break;
        case 8: f2 = attention_flash_kernel<8>; break;
        default: break;
        }
    }
    if (!f2) return attention_kernel;
    const int ng = NWARP * (32 / lpr);
    *smem = ((size_t)ng * head_size + 3 * (size_t)ng) * sizeof(float);
    return f2;
}

int launch_attention(l2b_ctx *ctx, int layer, cudaStream_t st) {
    if (ctx->profiling) {
        int hpos = 0;   // the host knows pos only through the last set_ctl; stored in n_appended-1
        hpos = ctx->n_appended > 0 ? ctx->n_appended - 1 : 0;
        int prc = prof_mark(ctx, "attention", layer, 2ull * (uint64_t)(hpos + 1) * ctx->kv_loc * 4ull, st);
        if (prc) return prc;
    }
    AttnParams a{};
    a.ctl = ctx->ctl;
    a.q = ctx->q;
    const size_t loff = (size_t)layer * ctx->cfg.seq_len * ctx->kv_loc;   // :354
    a.kcache = ctx->kcache + loff;
    a.vcache = ctx->vcache + loff;
    a.xb = ctx->xb;
    a.part_o = ctx->part_o;
    a.part_ml = ctx->part_ml;
    a.counters = ctx->counters;
    a.head_size = ctx->head_size;
    a.kv_dim = ctx->kv_loc;
    a.kv_mul = ctx->kv_mul;
    a.nsplit = ctx->nsplit;
    a.min_chunk = ctx->min_chunk;
    a.trace = (ctx->trace && !ctx->profiling) ? ctx->trace + (size_t)(ctx->last_launches % ctx->trace_launches) * TRACE_MAX_CTAS * TRACE_SLOTS : nullptr;
    size_t smem = ctx->attn_smem;
    attn_fn fn = pick_attention(ctx->head_size, ctx->attn_flash, &smem);
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(ctx->heads_loc, ctx->nsplit);
    lc.blockDim = dim3(NT);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = ctx->use_pdl ? 1 : 0;
    L2B_CUDA(ctx, cudaLaunchKernelEx(&lc, fn, a));
    ++ctx->last_launches;
    return L2B_OK;
}

// One decode step on `st` (transformer(), src/main.zig:285-430).  want_argmax selects the
// classifier epilogue.  All (token, pos) dependence is through ctx->ctl, so the sequence is
// capturable once and replayed.
int enqueue_step(l2b_ctx *ctx, cudaStream_t st, bool want_argmax) {
    const l2b_config &c = ctx->cfg;
    const int dim = ctx->dim;
    int cur = 0;
    // weight map of the GEMV that follows a kernel, for its L2 prefetch (see GemvParams::pf_*)
    auto set_pf = [&](GemvParams &g, int epi, const float *w0, const float *w1, const float *w2, int rows0,
                      int rows1, int total, int n) {
        g.pf_epi = epi; g.pf_w0 = w0; g.pf_w1 = w1; g.pf_w2 = w2;
        g.pf_rows0 = rows0; g.pf_rows1 = rows1; g.pf_total_rows = total; g.pf_n = n;
        g.pf_bytes = ctx->pf_bytes;
    };
    auto pf_qkv = [&](GemvParams &g, int l) {
        set_pf(g, EPI_QKV, ctx->wq + (size_t)l * ctx->q_loc * dim, ctx->wk + (size_t)l * ctx->kv_loc * dim,
               ctx->wv + (size_t)l * ctx->kv_loc * dim, ctx->q_loc, ctx->kv_loc, ctx->q_loc + 2 * ctx->kv_loc, dim);
    };
    for (int l = 0; l < c.n_layers; ++l) {
        // ---- rmsnorm + q,k,v + RoPE + KV append (:305-358)
        GemvParams p{};
        p.ctl = ctx->ctl;
        p.n = dim;
        const bool p2p = ctx->world > 1 && ctx->use_p2p;
        auto consume_slot = [&](GemvParams &g, int slot) {      // pending residual = sum of all ranks' partials
            g.xparts = ctx->xchg + (size_t)slot * ctx->world * dim;
            g.xflags = ctx->xflags + (size_t)slot * ctx->world;
            g.xworld = ctx->world;
            g.xcount_per_step = (dim + 1) / 2;     // counters count row pairs landed
        };
        auto produce_slot = [&](GemvParams &g, int slot) {      // my partial rows go to every rank
            g.xworld = ctx->world;
            for (int r = 0; r < ctx->world; ++r) {
                g.xout_peer[r] = ctx->peer_xchg[r] + ((size_t)slot * ctx->world + ctx->rank) * dim;
                g.xflag_peer[r] = ctx->peer_flags[r] + (size_t)slot * ctx->world + ctx->rank;
            }
        };
        if (l == 0) {
            p.emb = ctx->emb;                 // :295-296
            p.bump_epoch = 1;
        } else {
            p.x_in = ctx->X[cur];
            if (p2p) consume_slot(p, 2 * (l - 1) + 1);
            else p.delta = ctx->delta_f;      // pending :422 of the previous layer
        }
        p.gamma = ctx->rms_att + (size_t)l * dim;
        p.x_out = ctx->X[cur ^ 1];
        cur ^= 1;
        p.w0 = ctx->wq + (size_t)l * ctx->q_loc * dim;
        p.w1 = ctx->wk + (size_t)l * ctx->kv_loc * dim;
        p.w2 = ctx->wv + (size_t)l * ctx->kv_loc * dim;
        p.rows0 = ctx->q_loc; p.rows1 = ctx->kv_loc; p.rows2 = ctx->kv_loc;
        p.total_rows = ctx->q_loc + 2 * ctx->kv_loc;
        p.out0 = ctx->q;
        const size_t loff = (size_t)l * c.seq_len * ctx->kv_loc;
        p.kcache = ctx->kcache + loff;
        p.vcache = ctx->vcache + loff;
        p.rope_cos = ctx->rope_cos; p.rope_sin = ctx->rope_sin;
        p.head_size = ctx->head_size; p.kv_dim = ctx->kv_loc;
        set_pf(p, EPI_STORE, ctx->wo + (size_t)l * dim * ctx->q_loc, nullptr, nullptr, dim, 0, dim, ctx->q_loc);
        int rc = launch_gemv(ctx, EPI_QKV, p, st, "qkv_rope", l);
        if (rc) return rc;

        // ---- attention (:361-389)
        rc = launch_attention(ctx, l, st);
        if (rc) return rc;

        // ---- wo (:392); the residual add (:395) is applied by the next kernel's prologue
        GemvParams o{};
        o.ctl = ctx->ctl;
        o.n = ctx->q_loc;
        o.x_in = ctx->xb;
        o.w0 = ctx->wo + (size_t)l * dim * ctx->q_loc;
        o.total_rows = dim; o.rows0 = dim;
        o.out0 = ctx->delta_a;
        if (p2p) produce_slot(o, 2 * l);
        set_pf(o, EPI_SILU, ctx->w1 + (size_t)l * ctx->hid_loc * dim, ctx->w3 + (size_t)l * ctx->hid_loc * dim, nullptr,
               ctx->hid_loc, 0, 2 * ctx->hid_loc, dim);
        rc = launch_gemv(ctx, p2p ? EPI_XCHG : EPI_STORE, o, st, "wo", l);
        if (rc) return rc;
        if (p2p) ctx->xgrid[2 * l] = ctx->last_grid;
        else if (ctx->world > 1)
            L2B_NCCL(ctx, g_nccl.AllReduce(ctx->delta_a, ctx->delta_a, dim, ncclFloat, ncclSum, ctx->comm, st));

        // ---- residual + rmsnorm + w1,w3 + SiLU*mul (:395-416)
        GemvParams f{};
        f.ctl = ctx->ctl;
        f.n = dim;
        f.x_in = ctx->X[cur];
        if (p2p) consume_slot(f, 2 * l);
        else f.delta = ctx->delta_a;
        f.gamma = ctx->rms_ffn + (size_t)l * dim;
        f.x_out = ctx->X[cur ^ 1];
        cur ^= 1;
        f.w0 = ctx->w1 + (size_t)l * ctx->hid_loc * dim;
        f.w1 = ctx->w3 + (size_t)l * ctx->hid_loc * dim;
        f.rows0 = ctx->hid_loc;
        f.total_rows = 2 * ctx->hid_loc;
        f.out0 = ctx->hb;
        set_pf(f, EPI_STORE, ctx->w2 + (size_t)l * dim * ctx->hid_loc, nullptr, nullptr, dim, 0, dim, ctx->hid_loc);
        rc = launch_gemv(ctx, EPI_SILU, f, st, "w13_silu", l);
        if (rc) return rc;

        // ---- w2 (:419); residual (:422) deferred likewise
        GemvParams d{};
        d.ctl = ctx->ctl;
        d.n = ctx->hid_loc;
        d.x_in = ctx->hb;
        d.w0 = ctx->w2 + (size_t)l * dim * ctx->hid_loc;
        d.total_rows = dim; d.rows0 = dim;
        d.out0 = ctx->delta_f;
        if (p2p) produce_slot(d, 2 * l + 1);
        if (l + 1 < c.n_layers) pf_qkv(d, l + 1);
        else set_pf(d, EPI_STORE, ctx->wcls, nullptr, nullptr, ctx->vocab_loc, 0, ctx->vocab_loc, dim);
        rc = launch_gemv(ctx, p2p ? EPI_XCHG : EPI_STORE, d, st, "w2", l);
        if (rc) return rc;
        if (p2p) ctx->xgrid[2 * l + 1] = ctx->last_grid;
        else if (ctx->world > 1)
            L2B_NCCL(ctx, g_nccl.AllReduce(ctx->delta_f, ctx->delta_f, dim, ncclFloat, ncclSum, ctx->comm, st));
    }
    // ---- final residual + rmsnorm + classifier (:422-429)
    GemvParams k{};
    k.ctl = ctx->ctl;
    k.n = dim;
    k.x_in = ctx->X[cur];
    if (ctx->world > 1 && ctx->use_p2p) {
        const int slot = 2 * (c.n_layers - 1) + 1;
        k.xparts = ctx->xchg + (size_t)slot * ctx->world * dim;
        k.xflags = ctx->xflags + (size_t)slot * ctx->world;
        k.xworld = ctx->world;
        k.xcount_per_step = (dim + 1) / 2;
    } else {
        k.delta = ctx->delta_f;
    }
    k.gamma = ctx->rms_final;
    k.x_out = ctx->X[cur ^ 1];
    cur ^= 1;
    ctx->final_x = cur;
    k.w0 = ctx->wcls;
    k.total_rows = ctx->vocab_loc; k.rows0 = ctx->vocab_loc;
    k.out0 = (ctx->world > 1) ? ctx->logits_loc : ctx->logits;
    k.amax = ctx->amax;
    k.row_base = ctx->rank * ctx->vocab_loc;
    pf_qkv(k, 0);   // the next token starts with layer 0's q/k/v rows
    int rc = launch_gemv(ctx, want_argmax ? EPI_ARGMAX : EPI_STORE, k, st, "classifier", -1);
    if (rc) return rc;
    if (ctx->world > 1) {
        if (want_argmax)
            L2B_NCCL(ctx, g_nccl.AllReduce(ctx->amax, ctx->amax, 1, ncclUint64, ncclMax, ctx->comm, st));
        else
            L2B_NCCL(ctx, g_nccl.AllGather(ctx->logits_loc, ctx->logits, ctx->vocab_loc, ncclFloat, ctx->comm, st));
    }
    return L2B_OK;
}

int launch_mega(l2b_ctx *ctx, cudaStream_t st, bool want_argmax, bool do_advance) {
    MegaParams mp{};
    mp.emb = ctx->emb; mp.rms_att = ctx->rms_att; mp.rms_ffn = ctx->rms_ffn; mp.rms_final = ctx->rms_final;
    mp.wq = ctx->wq; mp.wk = ctx->wk; mp.wv = ctx->wv; mp.wo = ctx->wo;
    mp.w1 = ctx->w1; mp.w2 = ctx->w2; mp.w3 = ctx->w3; mp.wcls = ctx->wcls;
    mp.X0 = ctx->X[0]; mp.X1 = ctx->X[1]; mp.delta_a = ctx->delta_a; mp.delta_f = ctx->delta_f;
    mp.q = ctx->q; mp.xb = ctx->xb; mp.hb = ctx->hb;
    mp.logits = (ctx->world > 1) ? ctx->logits_loc : ctx->logits;
    mp.kcache = ctx->kcache; mp.vcache = ctx->vcache;
    mp.rope_cos = ctx->rope_cos; mp.rope_sin = ctx->rope_sin;
    mp.part_o = ctx->part_o; mp.part_ml = ctx->part_ml; mp.counters = ctx->counters;
    mp.ctl = ctx->ctl; mp.amax = ctx->amax; mp.gbar = ctx->gbar;
    mp.dim = ctx->dim; mp.hid_loc = ctx->hid_loc; mp.q_loc = ctx->q_loc; mp.kv_loc = ctx->kv_loc;
    mp.heads_loc = ctx->heads_loc; mp.vocab_loc = ctx->vocab_loc; mp.n_layers = ctx->cfg.n_layers;
    mp.seq_len = ctx->cfg.seq_len; mp.head_size = ctx->head_size; mp.kv_mul = ctx->kv_mul;
    mp.nsplit = ctx->nsplit; mp.min_chunk = ctx->min_chunk;
    mp.nstage = ctx->mega_nstage; mp.xs_floats = ctx->mega_xs_floats;
    mp.want_argmax = want_argmax ? 1 : 0;
    mp.row_base = ctx->rank * ctx->vocab_loc;
    mp.do_advance = do_advance ? 1 : 0;
    mp.world = ctx->world; mp.rank = ctx->rank;
    for (int r = 0; r < ctx->world && r < MAX_TP; ++r) { mp.peer_xchg[r] = ctx->peer_xchg[r]; mp.peer_flags[r] = ctx->peer_flags[r]; }
    mp.xchg = ctx->xchg; mp.xflags = ctx->xflags;
    mp.forced = ctx->gen_forced; mp.out_next = ctx->gen_out; mp.n_done = ctx->gen_ndone;
    void *args[] = {&mp};
    L2B_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)mega_step_kernel, dim3(ctx->num_sms), dim3(MEGA_THREADS),
                                              args, (size_t)ctx->mega_smem, st));
    ++ctx->last_launches;
    return L2B_OK;
}

// one decode step through the megakernel (+ the collectives a sharded context still needs)
int enqueue_mega_step(l2b_ctx *ctx, cudaStream_t st, bool want_argmax, bool advance_after) {
    const bool tp = ctx->world > 1;
    int rc = launch_mega(ctx, st, want_argmax, advance_after && !tp);
    if (rc) return rc;
    if (tp) {
        if (want_argmax) {
            L2B_NCCL(ctx, g_nccl.AllReduce(ctx->amax, ctx->amax, 1, ncclUint64, ncclMax, ctx->comm, st));
            if (advance_after) {
                advance_kernel<<<1, 1, 0, st>>>(ctx->ctl, ctx->amax, ctx->gen_forced, ctx->gen_out, ctx->gen_ndone);
                L2B_CUDA(ctx, cudaGetLastError());
                ++ctx->last_launches;
            }
        } else {
            L2B_NCCL(ctx, g_nccl.AllGather(ctx->logits_loc, ctx->logits, ctx->vocab_loc, ncclFloat, ctx->comm, st));
        }
    }
    return L2B_OK;
}

int build_graphs(l2b_ctx *ctx) {
    for (int which = 0; which < 2; ++which) {
        cudaGraph_t g = nullptr;
        L2B_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        ctx->last_launches = 0;
        int rc = enqueue_step(ctx, ctx->stream, which == 1);
        if (rc == L2B_OK) {
            if (which == 0) {
                cudaError_t e = cudaMemcpyAsync(ctx->h_logits, ctx->logits, (size_t)ctx->cfg.vocab_size * 4,
                                                cudaMemcpyDeviceToHost, ctx->stream);
                if (e != cudaSuccess) rc = fail(ctx, L2B_ERR_CUDA, cudaGetErrorString(e));
            } else {
                advance_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, ctx->amax, ctx->gen_forced, ctx->gen_out,
                                                         ctx->gen_ndone);
                ++ctx->last_launches;
            }
        }
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        L2B_CUDA(ctx, e);
        cudaGraphExec_t *dst = which == 0 ? &ctx->graph_logits : &ctx->graph_argmax;
        L2B_CUDA(ctx, cudaGraphInstantiate(dst, g, 0));
        L2B_CUDA(ctx, cudaGraphDestroy(g));
        if (which == 0) ctx->launches_per_step = ctx->last_launches;
    }
    return L2B_OK;
}

int check_step_args(l2b_ctx *ctx, int token, int pos) {
    if (!ctx) return L2B_ERR_INVALID_ARG;
    if (token < 0 || token >= ctx->cfg.vocab_size) return fail(ctx, L2B_ERR_INVALID_ARG, "token out of range");
    if (pos < 0 || pos >= ctx->cfg.seq_len) return fail(ctx, L2B_ERR_INVALID_ARG, "pos out of range");
    if (pos > ctx->n_appended) return fail(ctx, L2B_ERR_STATE, "pos skips ahead of the KV cache");
    return L2B_OK;
}

int common_create(l2b_ctx **out, const l2b_config *cfg, const Source &src, uint64_t n_floats,
                  const float *rope_cos, const float *rope_sin, const l2b_shard *shard) {
    if (!out || !cfg) { g_create_error = "NULL argument"; return L2B_ERR_INVALID_ARG; }
    *out = nullptr;
    const int world = shard ? shard->world_size : 1;
    const int rank = shard ? shard->rank : 0;
    std::string why;
    int rc = validate_config(cfg, world, &why);
    if (rc) { g_create_error = why; return rc; }
    if (rank < 0 || rank >= world) { g_create_error = "rank out of range"; return L2B_ERR_INVALID_ARG; }
    if (src.host && n_floats < checkpoint_floats(cfg)) {
        g_create_error = "host_weights shorter than the checkpoint layout requires";
        return L2B_ERR_INVALID_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device";
        return L2B_ERR_NO_DEVICE;
    }
    const int device = shard ? shard->device : 0;
    if (device < 0 || device >= ndev) { g_create_error = "device ordinal out of range"; return L2B_ERR_INVALID_ARG; }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_create_error = "cudaGetDeviceProperties failed"; return L2B_ERR_CUDA; }
    if (prop.major != 10) { g_create_error = "device is not compute capability 10.x (built for sm_100a only)"; return L2B_ERR_NO_DEVICE; }

    l2b_ctx *ctx = new l2b_ctx();
    ctx->cfg = *cfg;
    ctx->rank = rank; ctx->world = world; ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->dim = cfg->dim; ctx->hidden = cfg->hidden_dim;
    ctx->head_size = cfg->dim / cfg->n_heads;
    ctx->kv_mul = cfg->n_heads / cfg->n_kv_heads;
    ctx->q_dim = cfg->dim;
    ctx->kv_dim = ctx->head_size * cfg->n_kv_heads;
    ctx->q_loc = ctx->q_dim / world; ctx->kv_loc = ctx->kv_dim / world;
    ctx->hid_loc = cfg->hidden_dim / world;
    ctx->heads_loc = cfg->n_heads / world;
    ctx->vocab_loc = cfg->vocab_size / world;

#define L2B_TRY(expr)                                  \
    do {                                               \
        int rc__ = (expr);                             \
        if (rc__) {                                    \
            g_create_error = ctx->err;                 \
            l2b_destroy(ctx);                          \
            return rc__;                               \
        }                                              \
    } while (0)
    auto cuda_try = [&](cudaError_t e, const char *what) -> int {
        if (e == cudaSuccess) return L2B_OK;
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? L2B_ERR_OOM : L2B_ERR_CUDA;
    };

    L2B_TRY(cuda_try(cudaSetDevice(device), "cudaSetDevice"));
    L2B_TRY(cuda_try(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate"));
    L2B_TRY(cuda_try(cudaEventCreate(&ctx->ev0), "cudaEventCreate"));
    L2B_TRY(cuda_try(cudaEventCreate(&ctx->ev1), "cudaEventCreate"));

    const uint64_t dim = cfg->dim, hid = cfg->hidden_dim, L = cfg->n_layers, V = cfg->vocab_size, S = cfg->seq_len;
    const uint64_t hs = ctx->head_size, kvd = ctx->kv_dim;
    const double sd = sqrt(288.0 / (double)dim), sh = sqrt(768.0 / (double)hid);
    const float BIG = 3.0e38f;
    // payload offsets in checkpoint order (src/main.zig:85-112)
    uint64_t off = 0;
    const uint64_t o_emb = off;  off += V * dim;
    const uint64_t o_ratt = off; off += L * dim;
    const uint64_t o_wq = off;   off += L * dim * dim;
    const uint64_t o_wk = off;   off += L * kvd * dim;
    const uint64_t o_wv = off;   off += L * kvd * dim;
    const uint64_t o_wo = off;   off += L * dim * dim;
    const uint64_t o_rffn = off; off += L * dim;
    const uint64_t o_w1 = off;   off += L * hid * dim;
    const uint64_t o_w2 = off;   off += L * dim * hid;
    const uint64_t o_w3 = off;   off += L * hid * dim;
    const uint64_t o_rfin = off; off += dim;
    off += 2 * (S * hs / 2);     // freq_cis_real/imag: not used by transformer() (:67-69, :298-300)
    const uint64_t o_wcls = off;

    L2B_TRY(materialize(ctx, &ctx->emb, src, o_emb, 1, 1, V, dim, SH_NONE, {0.0, 0.04, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->rms_att, src, o_ratt, 2, 1, L, dim, SH_NONE, {1.35, 0.35, 0.25f, 2.4f}));
    L2B_TRY(materialize(ctx, &ctx->wq, src, o_wq, 3, L, dim, dim, SH_ROWS, {0.0, 0.04 * sd, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->wk, src, o_wk, 4, L, kvd, dim, SH_ROWS, {0.0, 0.04 * sd, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->wv, src, o_wv, 5, L, kvd, dim, SH_ROWS, {0.0, 0.02 * sd, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->wo, src, o_wo, 6, L, dim, dim, SH_COLS, {0.0, 0.02 * sd, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->rms_ffn, src, o_rffn, 7, 1, L, dim, SH_NONE, {1.35, 0.35, 0.25f, 2.4f}));
    L2B_TRY(materialize(ctx, &ctx->w1, src, o_w1, 8, L, hid, dim, SH_ROWS, {0.0, 0.026 * sd, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->w2, src, o_w2, 9, L, dim, hid, SH_COLS, {0.0, 0.026 * sh, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->w3, src, o_w3, 10, L, hid, dim, SH_ROWS, {0.0, 0.026 * sd, -BIG, BIG}));
    L2B_TRY(materialize(ctx, &ctx->rms_final, src, o_rfin, 11, 1, 1, dim, SH_NONE, {7.1, 0.6, 3.0f, 10.0f}));
    if (cfg->shared_weights) {
        ctx->wcls = ctx->emb + (size_t)rank * ctx->vocab_loc * dim;   // :112
    } else {
        L2B_TRY(materialize(ctx, &ctx->wcls, src, o_wcls, 14, 1, V, dim, SH_ROWS, {0.0, 0.04, -BIG, BIG}));
        ctx->wcls_owned = true;
    }

    // ---- run state
    L2B_TRY(dev_alloc(ctx, &ctx->X[0], dim));
    L2B_TRY(dev_alloc(ctx, &ctx->X[1], dim));
    L2B_TRY(dev_alloc(ctx, &ctx->delta_a, dim));
    L2B_TRY(dev_alloc(ctx, &ctx->delta_f, dim));
    L2B_TRY(dev_alloc(ctx, &ctx->q, (size_t)ctx->q_loc));
    L2B_TRY(dev_alloc(ctx, &ctx->xb, (size_t)ctx->q_loc));
    L2B_TRY(dev_alloc(ctx, &ctx->hb, (size_t)ctx->hid_loc));
    L2B_TRY(dev_alloc(ctx, &ctx->logits, V));
    if (world > 1) L2B_TRY(dev_alloc(ctx, &ctx->logits_loc, (size_t)ctx->vocab_loc));
    L2B_TRY(dev_alloc(ctx, &ctx->kcache, L * S * ctx->kv_loc));
    L2B_TRY(dev_alloc(ctx, &ctx->vcache, L * S * ctx->kv_loc));
    L2B_TRY(dev_alloc(ctx, &ctx->rope_cos, S * hs / 2));
    L2B_TRY(dev_alloc(ctx, &ctx->rope_sin, S * hs / 2));
    L2B_TRY(dev_alloc(ctx, &ctx->ctl, (size_t)CTL_WORDS));
    L2B_TRY(dev_alloc(ctx, &ctx->amax, (size_t)1));
    L2B_TRY(dev_alloc(ctx, &ctx->gen_forced, S));
    L2B_TRY(dev_alloc(ctx, &ctx->gen_out, S));
    L2B_TRY(dev_alloc(ctx, &ctx->gen_ndone, (size_t)1));
    L2B_TRY(cuda_try(cudaMemsetAsync(ctx->ctl, 0, CTL_WORDS * sizeof(int), ctx->stream), "memset"));
    L2B_TRY(cuda_try(cudaMemsetAsync(ctx->amax, 0, sizeof(unsigned long long), ctx->stream), "memset"));
    L2B_TRY(cuda_try(cudaMemsetAsync(ctx->gen_forced, 0xff, S * sizeof(int), ctx->stream), "memset"));
    L2B_TRY(cuda_try(cudaMemsetAsync(ctx->kcache, 0, L * S * ctx->kv_loc * sizeof(float), ctx->stream), "memset"));
    L2B_TRY(cuda_try(cudaMemsetAsync(ctx->vcache, 0, L * S * ctx->kv_loc * sizeof(float), ctx->stream), "memset"));

    // ---- attention launch shape: enough (head, split) CTAs to cover the SMs, chunks >= min_chunk positions
    {
        const char *envc = getenv("L2B_ATTN_MIN_CHUNK");
        if (envc && atoi(envc) >= 8) ctx->min_chunk = atoi(envc);
        int ns = (2 * ctx->num_sms + ctx->heads_loc - 1) / ctx->heads_loc;
        const int max_by_len = (int)((S + ctx->min_chunk - 1) / ctx->min_chunk);
        if (ns > max_by_len) ns = max_by_len;
        if (ns > 32) ns = 32;
        if (ns < 1) ns = 1;
        ctx->nsplit = ns;
        int cap = (int)((S + ns - 1) / ns);
        if (cap < ctx->min_chunk) cap = ctx->min_chunk;
        const int G = NT / (int)(hs / 4);
        ctx->attn_smem = (int)(((uint64_t)G * hs + cap) * sizeof(float));
        if (ctx->attn_smem > kMaxDynSmem) {
            ctx->err = "seq_len too large for the attention kernel's shared memory";
            g_create_error = ctx->err;
            l2b_destroy(ctx);
            return L2B_ERR_UNSUPPORTED;
        }
        L2B_TRY(cuda_try(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem), "cudaFuncSetAttribute"));
        L2B_TRY(dev_alloc(ctx, &ctx->part_o, (size_t)ctx->heads_loc * ns * hs));
        L2B_TRY(dev_alloc(ctx, &ctx->part_ml, (size_t)ctx->heads_loc * ns * 2));
        L2B_TRY(dev_alloc(ctx, &ctx->counters, (size_t)ctx->heads_loc));
        L2B_TRY(cuda_try(cudaMemsetAsync(ctx->counters, 0, ctx->heads_loc * sizeof(unsigned int), ctx->stream), "memset"));
    }

    // ---- persistent megakernel: shared-memory budget = ring + activation/attention scratch
    {
        L2B_TRY(dev_alloc(ctx, &ctx->gbar, (size_t)64));
        L2B_TRY(cuda_try(cudaMemsetAsync(ctx->gbar, 0, 64 * sizeof(unsigned long long), ctx->stream), "memset"));
        size_t xs = (size_t)ctx->dim;
        if ((size_t)ctx->hid_loc > xs) xs = ctx->hid_loc;
        if ((size_t)ctx->q_loc > xs) xs = ctx->q_loc;
        const size_t attn_need = (size_t)ctx->attn_smem / sizeof(float);
        if (attn_need > xs) xs = attn_need;
        xs = (xs + 3) & ~(size_t)3;
        const size_t stage_bytes = (size_t)TMA_STAGE_FLOATS * 4;
        int ns = (int)(((size_t)kMaxSmemOptin - xs * 4) / stage_bytes);
        if (ns > TMA_MAX_STAGES) ns = TMA_MAX_STAGES;
        ctx->mega_nstage = ns;
        ctx->mega_xs_floats = (int)xs;
        ctx->mega_smem = (int)(ns * stage_bytes + xs * 4);
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        // opt-in: measured slower than the CUDA-graph + PDL chain on every workload so far
        // (profiles/r01_megakernel.md) — the per-phase grid barrier + activation staging chain
        // (~6 us) costs more than a PDL kernel boundary (~3.4 us)
        const char *envm = getenv("L2B_MEGA");
        const bool want = envm && envm[0] == '1';
        ctx->use_mega = want && coop && ns >= 2 && hs <= 256;
        if (ctx->use_mega) {
            L2B_TRY(cuda_try(cudaFuncSetAttribute(mega_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin), "cudaFuncSetAttribute(mega)"));
            int occ = 0;
            L2B_TRY(cuda_try(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mega_step_kernel, MEGA_THREADS, ctx->mega_smem), "occupancy(mega)"));
            if (occ < 1) ctx->use_mega = false;
        }
    }

    // ---- RoPE table (:338-342); the host's own libm values when the caller passes them
    {
        std::vector<float> hc(S * hs / 2), hsn(S * hs / 2);
        if (rope_cos && rope_sin) {
            memcpy(hc.data(), rope_cos, hc.size() * sizeof(float));
            memcpy(hsn.data(), rope_sin, hsn.size() * sizeof(float));
        } else {
            for (uint64_t p = 0; p < S; ++p)
                for (uint64_t j = 0; j < hs / 2; ++j) {
                    const float head_dim = (float)(2 * j);                                   // i % head_size
                    const float freq = 1.0f / powf(10000.0f, head_dim / (float)hs);          // :339
                    const float val = (float)p * freq;                                       // :340
                    hc[p * (hs / 2) + j] = cosf(val);                                        // :341
                    hsn[p * (hs / 2) + j] = sinf(val);                                       // :342
                }
        }
        L2B_TRY(cuda_try(cudaMemcpy(ctx->rope_cos, hc.data(), hc.size() * sizeof(float), cudaMemcpyHostToDevice), "rope upload"));
        L2B_TRY(cuda_try(cudaMemcpy(ctx->rope_sin, hsn.data(), hsn.size() * sizeof(float), cudaMemcpyHostToDevice), "rope upload"));
    }

    // ---- pinned host buffers
    L2B_TRY(cuda_try(cudaHostAlloc((void **)&ctx->h_logits, V * sizeof(float), cudaHostAllocDefault), "cudaHostAlloc"));
    L2B_TRY(cuda_try(cudaHostAlloc((void **)&ctx->h_ints, 16 * sizeof(int), cudaHostAllocDefault), "cudaHostAlloc"));
    L2B_TRY(cuda_try(cudaHostAlloc((void **)&ctx->h_gen, S * sizeof(int), cudaHostAllocDefault), "cudaHostAlloc"));

    // ---- communicator
    if (world > 1) {
        std::string e;
        if (!nccl_load(&e)) { ctx->err = e; g_create_error = e; l2b_destroy(ctx); return L2B_ERR_COMM; }
        ncclUniqueId id;
        memcpy(&id, shard->comm_id, sizeof id);
        ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
        if (r != ncclSuccess) {
            ctx->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
            g_create_error = ctx->err;
            l2b_destroy(ctx);
            return L2B_ERR_COMM;
        }
    }

    // ---- fused all-reduce over peer memory: exchange buffers shared through CUDA IPC
    if (world > 1) {
        const char *mode = getenv("L2B_TP");
        const bool want = !(mode && strcmp(mode, "nccl") == 0);
        const size_t slots = 2 * (size_t)L;
        ctx->xgrid.assign(slots, 0);
        int ok = want ? 1 : 0;
        cudaIpcMemHandle_t hx{}, hf{};
        if (ok) {
            L2B_TRY(dev_alloc(ctx, &ctx->xchg, slots * world * dim));
            L2B_TRY(dev_alloc(ctx, &ctx->xflags, slots * world));
            L2B_TRY(cuda_try(cudaMemsetAsync(ctx->xflags, 0, slots * world * sizeof(unsigned int), ctx->stream), "memset"));
            L2B_TRY(cuda_try(cudaMemsetAsync(ctx->xchg, 0, slots * world * dim * sizeof(float), ctx->stream), "memset"));
            if (cudaIpcGetMemHandle(&hx, ctx->xchg) != cudaSuccess || cudaIpcGetMemHandle(&hf, ctx->xflags) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            }
        }
        // every rank learns every rank's handles (and whether it could make them) through NCCL
        struct Msg { cudaIpcMemHandle_t hx, hf; int ok; int pad[3]; };
        static_assert(sizeof(Msg) % 16 == 0, "Msg must be 16-byte sized");
        Msg mine{hx, hf, ok, {0, 0, 0}};
        Msg *d_all = nullptr;
        L2B_TRY(dev_alloc(ctx, &d_all, (size_t)world));
        L2B_TRY(cuda_try(cudaMemcpyAsync(d_all + rank, &mine, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream), "msg upload"));
        {
            ncclResult_t r = g_nccl.AllGather(d_all + rank, d_all, sizeof(Msg), ncclChar, ctx->comm, ctx->stream);
            if (r != ncclSuccess) { ctx->err = "ncclAllGather(ipc handles) failed"; g_create_error = ctx->err; l2b_destroy(ctx); return L2B_ERR_COMM; }
        }
        std::vector<Msg> all(world);
        L2B_TRY(cuda_try(cudaMemcpyAsync(all.data(), d_all, sizeof(Msg) * world, cudaMemcpyDeviceToHost, ctx->stream), "msg download"));
        L2B_TRY(cuda_try(cudaStreamSynchronize(ctx->stream), "ipc exchange"));
        for (int r = 0; r < world; ++r) ok = ok && all[r].ok;
        if (ok) {
            for (int r = 0; r < world && ok; ++r) {
                if (r == rank) { ctx->peer_xchg[r] = ctx->xchg; ctx->peer_flags[r] = ctx->xflags; continue; }
                void *px = nullptr, *pf = nullptr;
                if (cudaIpcOpenMemHandle(&px, all[r].hx, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                    cudaIpcOpenMemHandle(&pf, all[r].hf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = 0;
                    break;
                }
                ctx->ipc_opened.push_back(px);
                ctx->ipc_opened.push_back(pf);
                ctx->peer_xchg[r] = (float *)px;
                ctx->peer_flags[r] = (unsigned int *)pf;
            }
        }
        // all ranks must take the same path: agree through a max-reduce of the failure bit
        int *d_bad = nullptr;
        L2B_TRY(dev_alloc(ctx, &d_bad, (size_t)1));
        int bad = ok ? 0 : 1;
        L2B_TRY(cuda_try(cudaMemcpyAsync(d_bad, &bad, sizeof(int), cudaMemcpyHostToDevice, ctx->stream), "flag upload"));
        {
            ncclResult_t r = g_nccl.AllReduce(d_bad, d_bad, 1, ncclInt, ncclMax, ctx->comm, ctx->stream);
            if (r != ncclSuccess) { ctx->err = "ncclAllReduce(ipc agreement) failed"; g_create_error = ctx->err; l2b_destroy(ctx); return L2B_ERR_COMM; }
        }
        L2B_TRY(cuda_try(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "flag download"));
        L2B_TRY(cuda_try(cudaStreamSynchronize(ctx->stream), "ipc agreement"));
        ctx->use_p2p = (bad == 0);
    }

    L2B_TRY(cuda_try(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize"));
    {
        const char *env = getenv("L2B_NO_GRAPH");
        ctx->use_graphs = !(env && env[0] == '1');
        const char *env2 = getenv("L2B_NO_PDL");
        ctx->use_pdl = !(env2 && env2[0] == '1');
        const char *env3 = getenv("L2B_GEMV8_MIN_BYTES");
        if (env3) ctx->gemv8_min_bytes = atoll(env3);
        const char *env4 = getenv("L2B_GEMV_BIG");
        if (env4 && strcmp(env4, "ldg") == 0) ctx->big_kernel_tma = false;
        const char *env5 = getenv("L2B_TMA_STAGES");
        if (env5) ctx->tma_stages = atoi(env5);
        const char *enva = getenv("L2B_ATTN");
        if (enva && strcmp(enva, "3pass") == 0) ctx->attn_flash = false;
        const char *envt = getenv("L2B_TRACE");
        if (envt && envt[0] == '1') {
            ctx->trace_launches = 5 * cfg->n_layers + 2;
            L2B_TRY(dev_alloc(ctx, &ctx->trace, (size_t)ctx->trace_launches * TRACE_MAX_CTAS * TRACE_SLOTS));
            L2B_TRY(cuda_try(cudaMemset(ctx->trace, 0, (size_t)ctx->trace_launches * TRACE_MAX_CTAS * TRACE_SLOTS * 8), "memset"));
        }
        const char *env7 = getenv("L2B_PF_KB");
        if (env7) ctx->pf_bytes = atoi(env7) * 1024;
        const char *env6 = getenv("L2B_TMA_CTAS");
        if (env6) ctx->tma_ctas_per_sm = atoi(env6) > 0 ? atoi(env6) : 1;
    }
    // one eager step of each flavour: sets function attributes outside capture and surfaces
    // launch errors before a graph hides them (it scribbles on KV row 0, rewritten by step 0)
    {
        set_ctl_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, 0, 0, 0, ctx->amax);
        L2B_TRY(enqueue_step(ctx, ctx->stream, false));
        L2B_TRY(enqueue_step(ctx, ctx->stream, true));
        advance_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, ctx->amax, ctx->gen_forced, ctx->gen_out, ctx->gen_ndone);
        L2B_TRY(cuda_try(cudaStreamSynchronize(ctx->stream), "warm-up step"));
        L2B_TRY(cuda_try(cudaGetLastError(), "warm-up step"));
        if (ctx->use_mega) {
            set_ctl_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, 0, 0, 0, ctx->amax);
            L2B_TRY(enqueue_mega_step(ctx, ctx->stream, false, false));
            L2B_TRY(enqueue_mega_step(ctx, ctx->stream, true, true));
            L2B_TRY(cuda_try(cudaStreamSynchronize(ctx->stream), "warm-up megakernel step"));
            L2B_TRY(cuda_try(cudaGetLastError(), "warm-up megakernel step"));
        }
    }
    if (ctx->use_graphs) L2B_TRY(build_graphs(ctx));
#undef L2B_TRY
    *out = ctx;
    return L2B_OK;
}

// run one step; which: 0 = logits to pinned host, 1 = argmax
int run_step(l2b_ctx *ctx, int token, int pos, int which) {
    L2B_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->last_launches = 0;
    L2B_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (which == 1)   // a previous generate() may have left a forced token in slot 0
        L2B_CUDA(ctx, cudaMemsetAsync(ctx->gen_forced, 0xff, sizeof(int), ctx->stream));
    set_ctl_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, token, pos, 0, ctx->amax);
    L2B_CUDA(ctx, cudaGetLastError());
    if (ctx->use_mega) {
        int rc = enqueue_mega_step(ctx, ctx->stream, which == 1, which == 1);
        if (rc) return rc;
        if (which == 0)
            L2B_CUDA(ctx, cudaMemcpyAsync(ctx->h_logits, ctx->logits, (size_t)ctx->cfg.vocab_size * 4,
                                          cudaMemcpyDeviceToHost, ctx->stream));
    } else if (ctx->use_graphs) {
        L2B_CUDA(ctx, cudaGraphLaunch(which == 0 ? ctx->graph_logits : ctx->graph_argmax, ctx->stream));
        ctx->last_launches = ctx->launches_per_step + (which ? 1 : 0);
    } else {
        int rc = enqueue_step(ctx, ctx->stream, which == 1);
        if (rc) return rc;
        if (which == 0) {
            L2B_CUDA(ctx, cudaMemcpyAsync(ctx->h_logits, ctx->logits, (size_t)ctx->cfg.vocab_size * 4,
                                          cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            advance_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, ctx->amax, ctx->gen_forced, ctx->gen_out,
                                                     ctx->gen_ndone);
            L2B_CUDA(ctx, cudaGetLastError());
            ++ctx->last_launches;
        }
    }
    ++ctx->last_launches;  // set_ctl
    if (which == 1)
        L2B_CUDA(ctx, cudaMemcpyAsync(ctx->h_ints, ctx->gen_out, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    L2B_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    L2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    L2B_CUDA(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    if (pos + 1 > ctx->n_appended) ctx->n_appended = pos + 1;
    return L2B_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int32_t l2b_abi_version(void) { return L2B_ABI_VERSION; }

const char *l2b_status_string(int32_t s) {
    switch (s) {
    case L2B_OK: return "ok";
    case L2B_ERR_INVALID_ARG: return "invalid argument";
    case L2B_ERR_UNSUPPORTED: return "unsupported shape";
    case L2B_ERR_CUDA: return "CUDA error";
    case L2B_ERR_NO_DEVICE: return "no usable sm_100 device";
    case L2B_ERR_OOM: return "out of memory";
    case L2B_ERR_COMM: return "communication error";
    case L2B_ERR_STATE: return "call order violated";
    default: return "unknown status";
    }
}

const char *l2b_last_error(const l2b_ctx *ctx) {
    if (!ctx) return g_create_error.c_str();
    return ctx->err.empty() ? kNoError : ctx->err.c_str();
}

uint64_t l2b_checkpoint_floats(const l2b_config *cfg) { return cfg ? checkpoint_floats(cfg) : 0; }

int32_t l2b_create(l2b_ctx **out, const l2b_config *cfg, const float *host_weights, uint64_t n_floats,
                   const float *rope_cos, const float *rope_sin, int32_t n_gpus) {
    if (!host_weights) { g_create_error = "host_weights is NULL"; return L2B_ERR_INVALID_ARG; }
    if (n_gpus != 1) {
        g_create_error = "l2b_create drives one GPU; use l2b_create_sharded (one process per GPU) for 2/4/8";
        return L2B_ERR_UNSUPPORTED;
    }
    Source s; s.host = host_weights;
    return common_create(out, cfg, s, n_floats, rope_cos, rope_sin, nullptr);
}

int32_t l2b_create_sharded(l2b_ctx **out, const l2b_config *cfg, const float *host_weights,
                           uint64_t n_floats, const float *rope_cos, const float *rope_sin,
                           const l2b_shard *shard) {
    if (!host_weights || !shard) { g_create_error = "NULL argument"; return L2B_ERR_INVALID_ARG; }
    Source s; s.host = host_weights;
    return common_create(out, cfg, s, n_floats, rope_cos, rope_sin, shard);
}

int32_t l2b_create_synthetic(l2b_ctx **out, const l2b_config *cfg, uint64_t seed, const l2b_shard *shard) {
    Source s; s.host = nullptr; s.seed = seed;
    return common_create(out, cfg, s, 0, nullptr, nullptr, shard);
}

void l2b_destroy(l2b_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    // graphs first: NCCL keeps a communicator alive (ncclCommDestroy spins) while captured
    // graphs still reference it
    if (ctx->graph_logits) cudaGraphExecDestroy(ctx->graph_logits);
    if (ctx->graph_argmax) cudaGraphExecDestroy(ctx->graph_argmax);
    cudaDeviceSynchronize();
    if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
    for (void *p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->h_logits) cudaFreeHost(ctx->h_logits);
    if (ctx->h_ints) cudaFreeHost(ctx->h_ints);
    if (ctx->h_gen) cudaFreeHost(ctx->h_gen);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int32_t l2b_reset(l2b_ctx *ctx) {
    if (!ctx) return L2B_ERR_INVALID_ARG;
    ctx->n_appended = 0;
    return L2B_OK;
}

int32_t l2b_forward_pinned(l2b_ctx *ctx, int32_t token, int32_t pos, const float **logits) {
    int rc = check_step_args(ctx, token, pos);
    if (rc) return rc;
    if (!logits) return fail(ctx, L2B_ERR_INVALID_ARG, "logits is NULL");
    rc = run_step(ctx, token, pos, 0);
    if (rc) return rc;
    *logits = ctx->h_logits;
    return L2B_OK;
}

int32_t l2b_forward(l2b_ctx *ctx, int32_t token, int32_t pos, float *host_logits) {
    int rc = check_step_args(ctx, token, pos);
    if (rc) return rc;
    if (!host_logits) return fail(ctx, L2B_ERR_INVALID_ARG, "host_logits is NULL");
    rc = run_step(ctx, token, pos, 0);
    if (rc) return rc;
    // state.logits of the reference is one long-lived buffer (src/main.zig:149): copy with a
    // streaming memcpy from the pinned landing buffer (the D2H DMA itself is part of the graph)
    memcpy(host_logits, ctx->h_logits, (size_t)ctx->cfg.vocab_size * sizeof(float));
    return L2B_OK;
}

int32_t l2b_forward_argmax(l2b_ctx *ctx, int32_t token, int32_t pos, int32_t *next) {
    int rc = check_step_args(ctx, token, pos);
    if (rc) return rc;
    if (!next) return fail(ctx, L2B_ERR_INVALID_ARG, "next is NULL");
    rc = run_step(ctx, token, pos, 1);
    if (rc) return rc;
    *next = ctx->h_ints[0];
    return L2B_OK;
}

int32_t l2b_generate_argmax(l2b_ctx *ctx, int32_t token, int32_t pos, int32_t n_steps,
                            const int32_t *forced, int32_t stop_on_bos, int32_t *out_next,
                            int32_t *n_done) {
    int rc = check_step_args(ctx, token, pos);
    if (rc) return rc;
    if (!out_next || !n_done || n_steps < 0) return fail(ctx, L2B_ERR_INVALID_ARG, "bad generate arguments");
    if (n_steps > ctx->cfg.seq_len - pos) n_steps = ctx->cfg.seq_len - pos;   // :992-993
    *n_done = 0;
    if (n_steps == 0) return L2B_OK;
    L2B_CUDA(ctx, cudaSetDevice(ctx->device));
    // forced tokens (prompt forcing, :999-1000); -1 = free-running
    for (int i = 0; i < n_steps; ++i) ctx->h_gen[i] = forced ? forced[i] : -1;
    L2B_CUDA(ctx, cudaMemcpyAsync(ctx->gen_forced, ctx->h_gen, (size_t)n_steps * sizeof(int),
                                  cudaMemcpyHostToDevice, ctx->stream));
    L2B_CUDA(ctx, cudaMemsetAsync(ctx->gen_ndone, 0, sizeof(int), ctx->stream));
    ctx->last_launches = 0;
    L2B_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    set_ctl_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, token, pos, stop_on_bos ? 1 : 0, ctx->amax);
    L2B_CUDA(ctx, cudaGetLastError());
    int launches = 1;
    for (int i = 0; i < n_steps; ++i) {
        // the whole step depends on (token,pos) only through ctl, which advance_kernel updates,
        // so the same graph is replayed back to back with no host round trip
        if (ctx->use_mega) {
            ctx->last_launches = 0;
            rc = enqueue_mega_step(ctx, ctx->stream, true, true);
            if (rc) return rc;
            launches += ctx->last_launches;
        } else if (ctx->use_graphs) {
            L2B_CUDA(ctx, cudaGraphLaunch(ctx->graph_argmax, ctx->stream));
            launches += ctx->launches_per_step + 1;
        } else {
            ctx->last_launches = 0;
            rc = enqueue_step(ctx, ctx->stream, true);
            if (rc) return rc;
            advance_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, ctx->amax, ctx->gen_forced, ctx->gen_out,
                                                     ctx->gen_ndone);
            L2B_CUDA(ctx, cudaGetLastError());
            launches += ctx->last_launches + 1;
        }
    }
    L2B_CUDA(ctx, cudaMemcpyAsync(ctx->h_gen, ctx->gen_out, (size_t)n_steps * sizeof(int),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    L2B_CUDA(ctx, cudaMemcpyAsync(ctx->h_ints + 1, ctx->gen_ndone, sizeof(int), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    L2B_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    L2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    L2B_CUDA(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    const int done = ctx->h_ints[1];
    for (int i = 0; i < done; ++i) out_next[i] = ctx->h_gen[i];
    *n_done = done;
    ctx->last_launches = launches;
    if (pos + done > ctx->n_appended) ctx->n_appended = pos + done;
    return L2B_OK;
}

int32_t l2b_read_state(l2b_ctx *ctx, int32_t which, float *dst, uint64_t n, uint64_t *n_out) {
    if (!ctx || !dst) return L2B_ERR_INVALID_ARG;
    const float *src = nullptr;
    uint64_t cnt = 0;
    const uint64_t L = ctx->cfg.n_layers, S = ctx->cfg.seq_len;
    switch (which) {
    case 0: src = ctx->X[ctx->final_x]; cnt = ctx->dim; break;
    case 1: src = ctx->xb; cnt = ctx->q_loc; break;
    case 2: src = ctx->hb; cnt = ctx->hid_loc; break;
    case 3: src = ctx->q; cnt = ctx->q_loc; break;
    case 4: src = ctx->kcache; cnt = L * S * ctx->kv_loc; break;
    case 5: src = ctx->vcache; cnt = L * S * ctx->kv_loc; break;
    case 6: src = ctx->logits; cnt = ctx->cfg.vocab_size; break;
    default: return fail(ctx, L2B_ERR_INVALID_ARG, "unknown state id");
    }
    if (n < cnt) return fail(ctx, L2B_ERR_INVALID_ARG, "dst too small");
    L2B_CUDA(ctx, cudaSetDevice(ctx->device));
    L2B_CUDA(ctx, cudaMemcpy(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    if (n_out) *n_out = cnt;
    return L2B_OK;
}

int32_t l2b_last_timing(const l2b_ctx *ctx, float *device_ms, int32_t *kernel_launches) {
    if (!ctx) return L2B_ERR_INVALID_ARG;
    if (device_ms) *device_ms = ctx->last_ms;
    if (kernel_launches) *kernel_launches = ctx->last_launches;
    return L2B_OK;
}

int32_t l2b_profile_step(l2b_ctx *ctx, int32_t token, int32_t pos, l2b_kernel_time *out, int32_t cap,
                         int32_t *n_out) {
    int rc = check_step_args(ctx, token, pos);
    if (rc) return rc;
    if (!out || !n_out || cap <= 0) return fail(ctx, L2B_ERR_INVALID_ARG, "bad profile arguments");
    L2B_CUDA(ctx, cudaSetDevice(ctx->device));
    if (pos + 1 > ctx->n_appended) ctx->n_appended = pos + 1;   // attention bytes use this
    const int saved_appended = ctx->n_appended;
    ctx->n_appended = pos + 1;
    set_ctl_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, token, pos, 0, ctx->amax);
    L2B_CUDA(ctx, cudaGetLastError());
    ctx->profiling = true;
    ctx->prof_ev.clear();
    ctx->prof_rec.clear();
    rc = enqueue_step(ctx, ctx->stream, false);
    ctx->profiling = false;
    ctx->n_appended = saved_appended;
    cudaEvent_t last = nullptr;
    if (rc == L2B_OK) {
        if (cudaEventCreate(&last) != cudaSuccess || cudaEventRecord(last, ctx->stream) != cudaSuccess)
            rc = fail(ctx, L2B_ERR_CUDA, "event record failed");
    }
    if (rc == L2B_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        rc = fail(ctx, L2B_ERR_CUDA, "profile step failed");
    const int n = (int)ctx->prof_rec.size();
    if (rc == L2B_OK) {
        for (int i = 0; i < n; ++i) {
            cudaEvent_t b = (i + 1 < n) ? ctx->prof_ev[i + 1] : last;
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, ctx->prof_ev[i], b);
            ctx->prof_rec[i].ms = ms;
            if (i < cap) out[i] = ctx->prof_rec[i];
        }
        *n_out = n;
    }
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    if (last) cudaEventDestroy(last);
    ctx->prof_ev.clear();
    return rc;
}

// debug: copy the in-kernel timeline of the last step (L2B_TRACE=1) to the host
int32_t l2b_debug_trace(l2b_ctx *ctx, unsigned long long *dst, uint64_t cap_words, uint64_t *n_words) {
    if (!ctx || !dst) return L2B_ERR_INVALID_ARG;
    if (!ctx->trace) return fail(ctx, L2B_ERR_STATE, "tracing not enabled (L2B_TRACE=1)");
    const uint64_t n = (uint64_t)ctx->trace_launches * TRACE_MAX_CTAS * TRACE_SLOTS;
    if (cap_words < n) return fail(ctx, L2B_ERR_INVALID_ARG, "dst too small");
    L2B_CUDA(ctx, cudaMemcpy(dst, ctx->trace, n * 8, cudaMemcpyDeviceToHost));
    if (n_words) *n_words = n;
    return L2B_OK;
}

int32_t l2b_step_bytes(const l2b_ctx *ctx, int32_t pos, uint64_t *weight_bytes, uint64_t *kv_bytes) {
    if (!ctx) return L2B_ERR_INVALID_ARG;
    const uint64_t dim = ctx->dim, L = ctx->cfg.n_layers;
    // every weight touched once per token (+ the embedding row), this rank's shard
    uint64_t w = L * ((uint64_t)ctx->q_loc * dim + 2ull * ctx->kv_loc * dim + dim * ctx->q_loc +
                      3ull * ctx->hid_loc * dim + 2 * dim) +
                 dim + (uint64_t)ctx->vocab_loc * dim;
    if (weight_bytes) *weight_bytes = 4 * w;
    if (kv_bytes) *kv_bytes = 4ull * L * 2 * ctx->kv_loc * (uint64_t)(pos + 1);
    return L2B_OK;
}

int32_t l2b_comm_unique_id(uint8_t id[128]) {
    if (!id) return L2B_ERR_INVALID_ARG;
    std::string e;
    if (!nccl_load(&e)) { g_create_error = e; return L2B_ERR_COMM; }
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return L2B_ERR_COMM; }
    memcpy(id, &u, 128);
    return L2B_OK;
}

// ---- synthetic generator, host mirror ------------------------------------------------------
void l2b_synth_fill_host(float *dst, uint64_t first, uint64_t count, uint64_t tensor_seed, double mean,
                         double sigma, float lo, float hi) {
    const double scale = sigma * kSynthInvStd;
    for (uint64_t j = 0; j < count; ++j) {
        const int32_t s = synth_irwin_hall(tensor_seed, first + j);
        volatile double prod = (double)s * scale;   // no a*b+c contraction
        float v = (float)(prod + mean);
        v = fminf(fmaxf(v, lo), hi);
        dst[j] = v;
    }
}

int32_t l2b_synth_checkpoint_host(const l2b_config *cfg, uint64_t seed, float *data, uint64_t n_floats) {
    if (!cfg || !data) return L2B_ERR_INVALID_ARG;
    if (n_floats < checkpoint_floats(cfg)) return L2B_ERR_INVALID_ARG;
    const uint64_t dim = cfg->dim, hid = cfg->hidden_dim, L = cfg->n_layers, V = cfg->vocab_size, S = cfg->seq_len;
    const uint64_t hs = dim / cfg->n_heads, kvd = hs * cfg->n_kv_heads;
    const double sd = sqrt(288.0 / (double)dim), sh = sqrt(768.0 / (double)hid);
    const float BIG = 3.0e38f;
    float *p = data;
    auto T = [&](int id, uint64_t count, double mean, double sigma, float lo, float hi) {
        l2b_synth_fill_host(p, 0, count, mix64(seed * 1000003ull + (uint64_t)id), mean, sigma, lo, hi);
        p += count;
    };
    T(1, V * dim, 0.0, 0.04, -BIG, BIG);
    T(2, L * dim, 1.35, 0.35, 0.25f, 2.4f);
    T(3, L * dim * dim, 0.0, 0.04 * sd, -BIG, BIG);
    T(4, L * kvd * dim, 0.0, 0.04 * sd, -BIG, BIG);
    T(5, L * kvd * dim, 0.0, 0.02 * sd, -BIG, BIG);
    T(6, L * dim * dim, 0.0, 0.02 * sd, -BIG, BIG);
    T(7, L * dim, 1.35, 0.35, 0.25f, 2.4f);
    T(8, L * hid * dim, 0.0, 0.026 * sd, -BIG, BIG);
    T(9, L * dim * hid, 0.0, 0.026 * sh, -BIG, BIG);
    T(10, L * hid * dim, 0.0, 0.026 * sd, -BIG, BIG);
    T(11, dim, 7.1, 0.6, 3.0f, 10.0f);
    memset(p, 0, sizeof(float) * 2 * (S * hs / 2));
    p += 2 * (S * hs / 2);
    if (!cfg->shared_weights) T(14, V * dim, 0.0, 0.04, -BIG, BIG);
    return L2B_OK;
}

// ---- single ops with host buffers (unit-test surface) ----------------------------------------
namespace {
struct OpScope {
    int rc = L2B_OK;
    std::vector<void *> bufs;
    explicit OpScope(int device) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); rc = L2B_ERR_NO_DEVICE; g_create_error = "no CUDA device"; return; }
        if (device < 0 || device >= ndev) { rc = L2B_ERR_INVALID_ARG; g_create_error = "device out of range"; return; }
        if (cudaSetDevice(device) != cudaSuccess) { rc = L2B_ERR_CUDA; g_create_error = "cudaSetDevice failed"; }
    }
    float *up(const float *h, size_t n) {
        if (rc) return nullptr;
        void *d = nullptr;
        if (cudaMalloc(&d, (n ? n : 1) * sizeof(float)) != cudaSuccess) { rc = L2B_ERR_OOM; g_create_error = "cudaMalloc failed"; return nullptr; }
        bufs.push_back(d);
        if (h && cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { rc = L2B_ERR_CUDA; g_create_error = "H2D failed"; }
        return (float *)d;
    }
    void down(float *h, const float *d, size_t n) {
        if (rc) return;
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { rc = L2B_ERR_CUDA; g_create_error = std::string("op failed: ") + cudaGetErrorString(e); }
    }
    ~OpScope() { for (void *p : bufs) cudaFree(p); }
};
}  // namespace

int32_t l2b_op_matmul(int32_t device, float *xout, const float *x, const float *w, int32_t d, int32_t n) {
    if (!xout || !x || !w || d <= 0 || n <= 0) return L2B_ERR_INVALID_ARG;   // asserts :534-536
    OpScope s(device);
    float *dx = s.up(x, n), *dw = s.up(w, (size_t)d * n), *dout = s.up(nullptr, d);
    int *ctl = (int *)s.up(nullptr, CTL_WORDS);
    if (s.rc) return s.rc;
    cudaMemset(ctl, 0, CTL_WORDS * sizeof(int));
    if (n % 4 == 0) {
        l2b_ctx tmp;   // only for launch_gemv's bookkeeping
        tmp.device = device;
        cudaDeviceProp prop{};
        cudaGetDeviceProperties(&prop, device);
        tmp.num_sms = prop.multiProcessorCount;
        GemvParams p{};
        p.ctl = ctl; p.n = n; p.x_in = dx; p.w0 = dw; p.total_rows = d; p.rows0 = d; p.out0 = dout;
        int rc = launch_gemv(&tmp, EPI_STORE, p, 0);
        if (rc) { g_create_error = tmp.err; return rc; }
    } else {
        int blocks = (d + NWARP - 1) / NWARP;
        gemv_scalar_kernel<<<blocks, NT>>>(dout, dx, dw, d, n);
    }
    s.down(xout, dout, d);
    return s.rc;
}

int32_t l2b_op_rmsnorm(int32_t device, float *o, const float *x, const float *w, int32_t n) {
    if (!o || !x || !w || n <= 0) return L2B_ERR_INVALID_ARG;
    OpScope s(device);
    float *dx = s.up(x, n), *dw = s.up(w, n), *dout = s.up(nullptr, n);
    if (s.rc) return s.rc;
    rmsnorm_kernel<<<1, NT>>>(dout, dx, dw, n);
    s.down(o, dout, n);
    return s.rc;
}

int32_t l2b_op_softmax(int32_t device, float *x, int32_t n) {
    if (!x || n <= 0) return L2B_ERR_INVALID_ARG;   // assert :688
    OpScope s(device);
    float *dx = s.up(x, n);
    if (s.rc) return s.rc;
    softmax_kernel<<<1, NT>>>(dx, n);
    s.down(x, dx, n);
    return s.rc;
}

int32_t l2b_op_weighted_sum_rows(int32_t device, float *xout, int32_t out_len, const float *rows,
                                 int32_t row_stride, const float *weights, int32_t n_weights) {
    if (!xout || !rows || !weights || out_len <= 0 || n_weights <= 0 || row_stride < out_len)
        return L2B_ERR_INVALID_ARG;   // asserts :658-661
    OpScope s(device);
    const size_t nrows = (size_t)(n_weights - 1) * row_stride + out_len;
    float *dr = s.up(rows, nrows), *dw = s.up(weights, n_weights), *dout = s.up(nullptr, out_len);
    if (s.rc) return s.rc;
    weighted_rows_kernel<<<(out_len + NT - 1) / NT, NT>>>(dout, out_len, dr, row_stride, dw, n_weights);
    s.down(xout, dout, out_len);
    return s.rc;
}

int32_t l2b_op_attention_head(int32_t device, float *out, const float *q, const float *keys,
                              const float *values, int32_t head_size, int32_t kv_stride, int32_t n_pos) {
    if (!out || !q || !keys || !values || head_size <= 0 || n_pos <= 0 || kv_stride < head_size)
        return L2B_ERR_INVALID_ARG;
    if (head_size % 4 || kv_stride % 4 || head_size / 4 > NT) { g_create_error = "head_size/kv_stride must be multiples of 4"; return L2B_ERR_UNSUPPORTED; }
    OpScope s(device);
    const size_t nkv = (size_t)n_pos * kv_stride;
    float *dq = s.up(q, head_size), *dk = s.up(keys, nkv), *dv = s.up(values, nkv), *dout = s.up(nullptr, head_size);
    const int nsplit = n_pos > 256 ? 4 : 1;
    float *po = s.up(nullptr, (size_t)nsplit * head_size), *pml = s.up(nullptr, (size_t)nsplit * 2);
    unsigned int *cnt = (unsigned int *)s.up(nullptr, 1);
    int *ctl = (int *)s.up(nullptr, CTL_WORDS);
    if (s.rc) return s.rc;
    int hctl[CTL_WORDS] = {0, n_pos - 1, 0, 0, 0, 0, 0, 0};
    cudaMemcpy(ctl, hctl, sizeof hctl, cudaMemcpyHostToDevice);
    cudaMemset(cnt, 0, sizeof(unsigned int));
    AttnParams a{};
    a.ctl = ctl; a.q = dq; a.kcache = dk; a.vcache = dv; a.xb = dout; a.part_o = po; a.part_ml = pml;
    a.counters = cnt; a.head_size = head_size; a.kv_dim = kv_stride; a.kv_mul = 1; a.nsplit = nsplit;
    a.min_chunk = 64;
    int cap = (n_pos + nsplit - 1) / nsplit;
    if (cap < 64) cap = 64;
    const int G = NT / (head_size / 4);
    const size_t smem = ((size_t)G * head_size + cap) * sizeof(float);
    if (smem > (size_t)kMaxDynSmem) { g_create_error = "n_pos too large"; return L2B_ERR_UNSUPPORTED; }
    cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    {
        const char *enva = getenv("L2B_ATTN");
        size_t smem2 = smem;
        attn_fn fn = pick_attention(head_size, !(enva && strcmp(enva, "3pass") == 0), &smem2);
        fn<<<dim3(1, nsplit), NT, smem2>>>(a);
    }
    s.down(out, dout, head_size);
    return s.rc;
}

}  // extern "C"
