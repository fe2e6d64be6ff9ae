// l2b_mega.cuh — one persistent cooperative kernel per decode step (transformer(),
// src/main.zig:285-430), instead of 5*L+1 dependent kernel launches.
//
// Why: per kernel boundary the launch chain costs ~3.4 us on the small models (31 launches per
// stories15M token) and ~4-5 us of ring drain/refill on llama2-7B (161 boundaries, 13% of the
// step; worse under tensor parallelism where every kernel is 1/g the size).  Here one CTA per SM
// lives for the whole step:
//   * warp 8 (producer) streams the weight rows of EVERY phase, in order, through one shared
//     memory ring with cp.async.bulk; weights are immutable, so it never waits for a phase
//     boundary — only for free ring slots — and HBM stays busy across boundaries;
//   * warps 0-7 (consumers) and warp 9 (epilogue) walk the phases; between phases a grid-wide
//     barrier (one release-add per CTA, acquire-polled) gates only the ACTIVATIONS;
//   * activations are read with ld.global.cg (L2) because they are rewritten by other SMs inside
//     the same kernel; weights and RoPE tables stay on the read-only path.
// Phase list: for each layer qkv_rope, attention, wo, w13_silu, w2; then the classifier.
#pragma once

#include "l2b_device.cuh"

namespace l2b {

constexpr int MEGA_THREADS = NT + 64;          // 8 consumer warps + producer warp + epilogue warp
constexpr int MEGA_WORKERS = NT + 32;          // consumers + epilogue warp (named barrier 2)

enum { CTL_MEGA_EPOCH = 6 };                   // completed megakernel steps (grid-barrier base)

struct MegaParams {
    // weights (this rank's shards), all layers
    const float *emb, *rms_att, *rms_ffn, *rms_final, *wq, *wk, *wv, *wo, *w1, *w2, *w3, *wcls;
    // run state
    float *X0, *X1, *delta_a, *delta_f, *q, *xb, *hb, *logits;
    float *kcache, *vcache;
    const float *rope_cos, *rope_sin;
    float *part_o, *part_ml;
    unsigned int *counters;
    int *ctl;
    unsigned long long *amax;
    unsigned long long *gbar;                  // grid barrier: [0] arrival count, [32] released generation
                                               // (separate 128-byte lines: pollers never touch the counter)
    // dims
    int dim, hid_loc, q_loc, kv_loc, heads_loc, vocab_loc, n_layers, seq_len, head_size, kv_mul;
    int nsplit, min_chunk;
    int nstage;                                // ring depth
    int xs_floats;                             // size of the activation / attention scratch area
    int want_argmax, row_base, do_advance;
    // tensor parallel exchange (fused all-reduce), world == 1 => unused
    int world, rank;
    float *peer_xchg[MAX_TP];
    unsigned int *peer_flags[MAX_TP];
    float *xchg;
    unsigned int *xflags;
    // on-device generation loop bookkeeping (advance step)
    const int *forced;
    int *out_next;
    int *n_done;
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Grid barrier, arrival side (ONE thread per CTA, after the CTA's writes are fenced): the last CTA
// to arrive publishes the new generation on a separate line that everyone else polls.
__device__ __forceinline__ void grid_arrive(unsigned long long *gbar, unsigned long long target_count,
                                            unsigned long long generation) {
    __threadfence();
    const unsigned long long old = atomicAdd(gbar, 1ull);
    if (old + 1ull == target_count) {
        __threadfence();
        st_release_gpu_u64(gbar + 32, generation);
    }
}
__device__ __forceinline__ void grid_wait(const unsigned long long *gbar, unsigned long long generation) {
    while (ld_acquire_gpu_u64(gbar + 32) < generation) { }
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

enum MegaKind { MK_QKV = 0, MK_ATTN = 1, MK_WO = 2, MK_W13 = 3, MK_W2 = 4, MK_CLS = 5 };

// Phase ph of the step -> kind, layer, and (for GEMV phases) the GemvParams the standalone kernels
// would have been launched with.
__device__ __forceinline__ int mega_kind(const MegaParams &mp, int ph, int &layer) {
    if (ph == 5 * mp.n_layers) { layer = -1; return MK_CLS; }
    layer = ph / 5;
    return ph - layer * 5;
}

__device__ __forceinline__ void mega_make_gemv(const MegaParams &mp, int kind, int l, GemvParams &g) {
    const int dim = mp.dim;
    g.ctl = mp.ctl;
    g.emb = nullptr; g.delta = nullptr; g.gamma = nullptr; g.x_out = nullptr; g.x_in = nullptr;
    g.xparts = nullptr; g.xflags = nullptr; g.xworld = mp.world; g.xcount_per_step = 0; g.bump_epoch = 0;
    g.amax = mp.amax; g.row_base = mp.row_base;
    g.head_size = mp.head_size; g.kv_dim = mp.kv_loc;
    g.rope_cos = mp.rope_cos; g.rope_sin = mp.rope_sin;
    const bool tp = mp.world > 1;
    if (kind == MK_QKV) {
        g.n = dim;
        if (l == 0) g.emb = mp.emb;
        else {
            g.x_in = mp.X0;
            if (tp) { g.xparts = mp.xchg + (size_t)(2 * (l - 1) + 1) * mp.world * dim; g.xflags = mp.xflags + (size_t)(2 * (l - 1) + 1) * mp.world; }
            else g.delta = mp.delta_f;
        }
        g.gamma = mp.rms_att + (size_t)l * dim;
        g.x_out = mp.X1;
        g.w0 = mp.wq + (size_t)l * mp.q_loc * dim;
        g.w1 = mp.wk + (size_t)l * mp.kv_loc * dim;
        g.w2 = mp.wv + (size_t)l * mp.kv_loc * dim;
        g.rows0 = mp.q_loc; g.rows1 = mp.kv_loc; g.rows2 = mp.kv_loc;
        g.total_rows = mp.q_loc + 2 * mp.kv_loc;
        g.out0 = mp.q;
        const size_t loff = (size_t)l * mp.seq_len * mp.kv_loc;
        g.kcache = mp.kcache + loff;
        g.vcache = mp.vcache + loff;
    } else if (kind == MK_WO) {
        g.n = mp.q_loc;
        g.x_in = mp.xb;
        g.w0 = mp.wo + (size_t)l * dim * mp.q_loc;
        g.rows0 = dim; g.total_rows = dim;
        g.out0 = mp.delta_a;
        if (tp) for (int r = 0; r < mp.world; ++r) {
            g.xout_peer[r] = mp.peer_xchg[r] + ((size_t)(2 * l) * mp.world + mp.rank) * dim;
            g.xflag_peer[r] = mp.peer_flags[r] + (size_t)(2 * l) * mp.world + mp.rank;
        }
    } else if (kind == MK_W13) {
        g.n = dim;
        g.x_in = mp.X1;
        if (tp) { g.xparts = mp.xchg + (size_t)(2 * l) * mp.world * dim; g.xflags = mp.xflags + (size_t)(2 * l) * mp.world; }
        else g.delta = mp.delta_a;
        g.gamma = mp.rms_ffn + (size_t)l * dim;
        g.x_out = mp.X0;
        g.w0 = mp.w1 + (size_t)l * mp.hid_loc * dim;
        g.w1 = mp.w3 + (size_t)l * mp.hid_loc * dim;
        g.rows0 = mp.hid_loc;
        g.total_rows = 2 * mp.hid_loc;
        g.out0 = mp.hb;
    } else if (kind == MK_W2) {
        g.n = mp.hid_loc;
        g.x_in = mp.hb;
        g.w0 = mp.w2 + (size_t)l * dim * mp.hid_loc;
        g.rows0 = dim; g.total_rows = dim;
        g.out0 = mp.delta_f;
        if (tp) for (int r = 0; r < mp.world; ++r) {
            g.xout_peer[r] = mp.peer_xchg[r] + ((size_t)(2 * l + 1) * mp.world + mp.rank) * dim;
            g.xflag_peer[r] = mp.peer_flags[r] + (size_t)(2 * l + 1) * mp.world + mp.rank;
        }
    } else {  // MK_CLS
        const int last = 2 * (mp.n_layers - 1) + 1;
        g.n = dim;
        g.x_in = mp.X0;
        if (tp) { g.xparts = mp.xchg + (size_t)last * mp.world * dim; g.xflags = mp.xflags + (size_t)last * mp.world; }
        else g.delta = mp.delta_f;
        g.gamma = mp.rms_final;
        g.x_out = mp.X1;
        g.w0 = mp.wcls;
        g.rows0 = mp.vocab_loc; g.total_rows = mp.vocab_loc;
        g.out0 = mp.logits;
    }
}

// row pointer of virtual row v for a phase kind (same maps as gemv_row_ptr<EPI>)
__device__ __forceinline__ const float *mega_row_ptr(const GemvParams &g, int kind, int v) {
    if (kind == MK_QKV) return gemv_row_ptr<EPI_QKV>(g, v);
    if (kind == MK_W13) return gemv_row_ptr<EPI_SILU>(g, v);
    return gemv_row_ptr<EPI_STORE>(g, v);
}

// this CTA's contiguous, balanced range of rows for a GEMV phase
__device__ __forceinline__ void mega_row_range(int total_rows, int &r0, int &r1) {
    const int npairs = (total_rows + 1) >> 1;
    const int G = (int)gridDim.x, b = (int)blockIdx.x;
    const int base = npairs / G, rem = npairs % G;
    const int pair0 = b * base + min(b, rem);
    const int pair1 = pair0 + base + (b < rem ? 1 : 0);
    r0 = pair0 * 2;
    r1 = min(pair1 * 2, total_rows);
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) mega_step_kernel(const MegaParams mp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NSTAGE = mp.nstage;
    float *ring = reinterpret_cast<float *>(smem_raw);                 // NSTAGE x 32 KB
    float *xs = ring + (size_t)NSTAGE * TMA_STAGE_FLOATS;              // activation / attention scratch
    __shared__ uint64_t full[TMA_MAX_STAGES], empty[TMA_MAX_STAGES], tile_full[2], tile_free[2];
    __shared__ float scratch[NWARP + 2];
    __shared__ float red[2][NWARP][GEMV8_R];
    __shared__ float rope_s[2][128];
    __shared__ int is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = (int)gridDim.x;
    const int nph = 5 * mp.n_layers + 1;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWARP); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tile_full[s], NWARP); mbar_init(&tile_free[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    if (mp.ctl[CTL_DONE]) return;   // uniform: nothing issued yet
    const unsigned long long gen_base = (unsigned long long)mp.ctl[CTL_MEGA_EPOCH] * (unsigned long long)nph;
    const unsigned long long bar_base = gen_base * (unsigned long long)G;   // arrivals before this step
    const unsigned int tp_epoch = (unsigned int)mp.ctl[CTL_EPOCH] + 1u;   // this step's index for the TP counters
    const int pos = mp.ctl[CTL_POS];

    // =====================================================================================
    // producer warp: all GEMV phases, in order, never blocked by phase boundaries
    // =====================================================================================
    if (warp == NWARP) {
        int p_stage = 0;
        uint32_t p_phase = 0;
        for (int ph = 0; ph < nph; ++ph) {
            int l;
            const int kind = mega_kind(mp, ph, l);
            if (kind == MK_ATTN) continue;
            GemvParams g;
            mega_make_gemv(mp, kind, l, g);
            const int n4 = g.n >> 2;
            const int nsteps = (n4 + NT - 1) / NT;
            int r0, r1;
            mega_row_range(g.total_rows, r0, r1);
            const float *p_row = nullptr;
            for (int v0 = r0; v0 < r1; v0 += GEMV8_R) {
                const int rows = min(GEMV8_R, r1 - v0);
                if (lane < rows) p_row = mega_row_ptr(g, kind, v0 + lane);
                for (int st = 0; st < nsteps; ++st) {
                    const int c0 = st * NT;
                    const int cols4 = min(NT, n4 - c0);
                    if (lane == 0) {
                        mbar_wait(&empty[p_stage], p_phase ^ 1);
                        mbar_expect_tx(&full[p_stage], (uint32_t)(rows * cols4 * 16));
                    }
                    __syncwarp();
                    if (lane < rows)
                        tma_load_1d(ring + (size_t)p_stage * TMA_STAGE_FLOATS + lane * NT * 4, p_row + (size_t)c0 * 4,
                                    (uint32_t)(cols4 * 16), &full[p_stage]);
                    if (++p_stage == NSTAGE) { p_stage = 0; p_phase ^= 1; }
                }
            }
        }
        return;
    }

    // =====================================================================================
    // consumers (warps 0-7) and epilogue warp (warp 9)
    // =====================================================================================
    const bool is_epi = (warp == NWARP + 1);
    const int wtid = is_epi ? NT + lane : tid;             // 0..287 among the workers
    int c_stage = 0;
    uint32_t c_phase = 0;
    unsigned int tile_seq = 0;                             // running count of 8-row tiles (red[] hand-off)
    unsigned long long best = 0ull;

    for (int ph = 0; ph < nph; ++ph) {
        int l;
        const int kind = mega_kind(mp, ph, l);
        // ---- wait until every CTA has finished the previous phase
        if (ph > 0) {
            if (wtid == 0) grid_wait(mp.gbar, gen_base + (unsigned long long)ph);
            bar_sync_named(2, MEGA_WORKERS);
        }

        if (kind == MK_ATTN) {
            // ------------------------------------------------------------------ attention (:361-389)
            if (!is_epi) {
                const int hs = mp.head_size, hs4 = hs >> 2;
                const int AG = NT / hs4;
                float *ared = xs;
                float *sc = xs + AG * hs;
                const size_t loff = (size_t)l * mp.seq_len * mp.kv_loc;
                const float *kc = mp.kcache + loff, *vc = mp.vcache + loff;
                const int T = pos + 1;
                int chunk = (T + mp.nsplit - 1) / mp.nsplit;
                if (chunk < mp.min_chunk) chunk = mp.min_chunk;
                const int active = (T + chunk - 1) / chunk;
                const int nitems = mp.heads_loc * active;
                for (int item = blockIdx.x; item < nitems; item += G) {
                    const int h = item / active, s = item - h * active;
                    const int t0 = s * chunk, t1 = min(T, t0 + chunk), len = t1 - t0;
                    const size_t hoff = (size_t)(h / mp.kv_mul) * hs;
                    // scores
                    {
                        const int lpr = attn_lanes_per_row(hs4);
                        const int nf = hs4 / lpr, rows_per_warp = 32 / lpr;
                        const int lr = lane % lpr, rw = lane / lpr;
                        const float4 *q4 = reinterpret_cast<const float4 *>(mp.q + (size_t)h * hs);
                        const float root_hs = sqrtf((float)hs);
                        for (int tb = t0 + warp * rows_per_warp; tb < t1; tb += NWARP * rows_per_warp) {
                            const int t = tb + rw;
                            float acc = 0.0f;
                            if (t < t1) {
                                const float4 *k4 = reinterpret_cast<const float4 *>(kc + hoff + (size_t)t * mp.kv_loc);
                                for (int f = 0; f < nf; ++f) {
                                    const int j = lr + f * lpr;
                                    acc = dot4(__ldcg(k4 + j), __ldcg(q4 + j), acc);
                                }
                            }
                            for (int o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                            if (t < t1 && lr == 0) sc[t - t0] = acc / root_hs;      // :372
                        }
                    }
                    bar_sync_named(1, NT);
                    float m = -INFINITY;
                    for (int i = lane; i < len; i += 32) m = fmaxf(m, sc[i]);
                    m = warp_max(m);
                    bar_sync_named(1, NT);                                            // all warps have read the raw scores
                    for (int i = tid; i < len; i += NT) sc[i] = expf(sc[i] - m);      // :699
                    bar_sync_named(1, NT);
                    float lsum = 0.0f;
                    for (int i = lane; i < len; i += 32) lsum += sc[i];
                    lsum = warp_sum(lsum);
                    {
                        const int g = tid / hs4, c = tid % hs4;
                        if (g < AG) {
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int t = t0 + g; t < t1; t += AG) {
                                const float wt = sc[t - t0];
                                const float4 v = __ldcg(reinterpret_cast<const float4 *>(vc + hoff + (size_t)t * mp.kv_loc) + c);
                                a.x = fmaf(v.x, wt, a.x); a.y = fmaf(v.y, wt, a.y);
                                a.z = fmaf(v.z, wt, a.z); a.w = fmaf(v.w, wt, a.w);
                            }
                            reinterpret_cast<float4 *>(ared)[g * hs4 + c] = a;
                        }
                    }
                    bar_sync_named(1, NT);
                    float o = 0.0f;
                    if (tid < hs)
                        for (int gg = 0; gg < AG; ++gg) o += ared[gg * hs + tid];
                    if (active == 1) {
                        if (tid < hs) mp.xb[(size_t)h * hs + tid] = o / lsum;
                    } else {
                        if (tid < hs) mp.part_o[((size_t)h * mp.nsplit + s) * hs + tid] = o;
                        if (tid == 0) {
                            mp.part_ml[((size_t)h * mp.nsplit + s) * 2 + 0] = m;
                            mp.part_ml[((size_t)h * mp.nsplit + s) * 2 + 1] = lsum;
                        }
                        __threadfence();
                        bar_sync_named(1, NT);
                        if (tid == 0) {
                            const unsigned int prev = atomicAdd(&mp.counters[h], 1u);
                            is_last = (prev == (unsigned int)(active - 1));
                        }
                        bar_sync_named(1, NT);
                        if (is_last) {
                            __threadfence();
                            if (tid == 0) mp.counters[h] = 0u;
                            const volatile float *ml = mp.part_ml + (size_t)h * mp.nsplit * 2;
                            float M = -INFINITY;
                            for (int j = 0; j < active; ++j) M = fmaxf(M, ml[j * 2]);
                            float Lsum = 0.0f;
                            for (int j = 0; j < active; ++j) Lsum += expf(ml[j * 2] - M) * ml[j * 2 + 1];
                            if (tid < hs) {
                                const volatile float *pb = mp.part_o + (size_t)h * mp.nsplit * hs;
                                float acc = 0.0f;
                                for (int j = 0; j < active; ++j) acc = fmaf(expf(ml[j * 2] - M), pb[(size_t)j * hs + tid], acc);
                                mp.xb[(size_t)h * hs + tid] = acc / Lsum;
                            }
                        }
                    }
                    bar_sync_named(1, NT);   // scratch reuse by the next item
                }
                __threadfence();
            }
            bar_sync_named(2, MEGA_WORKERS);
            if (is_epi && lane == 0)
                grid_arrive(mp.gbar, bar_base + (unsigned long long)(ph + 1) * G, gen_base + (unsigned long long)(ph + 1));
            continue;
        }

        // ---------------------------------------------------------------------- GEMV phase
        GemvParams g;
        mega_make_gemv(mp, kind, l, g);
        const int n4 = g.n >> 2;
        const int nsteps = (n4 + NT - 1) / NT;
        int r0, r1;
        mega_row_range(g.total_rows, r0, r1);

        // ---- prologue: activation vector (+ pending residual [+ peers' partials], + rmsnorm)
        {
            float4 *xs4w = reinterpret_cast<float4 *>(xs);
            const float *xsrc = g.emb ? g.emb + (size_t)mp.ctl[CTL_TOKEN] * g.n : g.x_in;
            const float4 *x4 = reinterpret_cast<const float4 *>(xsrc);
            if (g.xparts) {
                // fused all-reduce, consumer half: every rank's partial rows must have landed
                const unsigned int per_step = (unsigned int)((mp.dim + 1) >> 1);   // row pairs per reduce point
                if (wtid < mp.world) {
                    const unsigned int want = tp_epoch * per_step;
                    while ((int)(ld_acquire_sys(g.xflags + wtid) - want) < 0) { }
                }
                bar_sync_named(2, MEGA_WORKERS);
            }
            if (kind == MK_QKV && is_epi) {
                const int half = mp.head_size >> 1;
                for (int i = lane; i < half; i += 32) {
                    rope_s[0][i] = mp.rope_cos[(size_t)pos * half + i];
                    rope_s[1][i] = mp.rope_sin[(size_t)pos * half + i];
                }
            }
            float ssq = 0.0f;
            const float4 *d4 = reinterpret_cast<const float4 *>(g.delta);
            const float4 *pp = reinterpret_cast<const float4 *>(g.xparts);
            const float4 *g4 = reinterpret_cast<const float4 *>(g.gamma);
            constexpr int GK = 4;                               // gain slices kept in registers (n <= 4608)
            float4 gv[GK];
            const bool gain_in_regs = g.gamma && n4 <= GK * MEGA_WORKERS;
            int k = 0;
            for (int i = wtid; i < n4; i += MEGA_WORKERS, ++k) {
                float4 v = g.emb ? __ldg(x4 + i) : __ldcg(x4 + i);
                if (gain_in_regs && k < GK) gv[k] = __ldg(g4 + i);             // issued with x: one round trip
                if (g.delta) {
                    const float4 d = __ldcg(d4 + i);
                    v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;               // accum(), :708-713
                } else if (g.xparts) {
                    for (int r = 0; r < mp.world; ++r) {                           // fixed rank order
                        const float4 d = __ldcg(pp + (size_t)r * n4 + i);
                        v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
                    }
                }
                if (g.x_out && blockIdx.x == 0) reinterpret_cast<float4 *>(g.x_out)[i] = v;
                xs4w[i] = v;
                ssq = fmaf(v.x, v.x, ssq); ssq = fmaf(v.y, v.y, ssq);
                ssq = fmaf(v.z, v.z, ssq); ssq = fmaf(v.w, v.w, ssq);
            }
            if (g.gamma) {
                ssq = warp_sum(ssq);
                const int wslot = is_epi ? NWARP : warp;
                if (lane == 0) scratch[wslot] = ssq;
                bar_sync_named(2, MEGA_WORKERS);
                float ss = (lane < NWARP + 1) ? scratch[lane] : 0.0f;
                ss = warp_sum(ss);
                ss /= (float)g.n;            // :452
                ss += 1e-5f;                 // :453
                const float sc = 1.0f / sqrtf(ss);  // :454
                k = 0;
                for (int i = wtid; i < n4; i += MEGA_WORKERS, ++k) {               // same thread wrote xs4w[i]
                    float4 v = xs4w[i];
                    float4 gg;
                    if (gain_in_regs) {
                        gg = gv[0];
#pragma unroll
                        for (int j = 1; j < GK; ++j) if (k == j) gg = gv[j];
                    } else {
                        gg = __ldg(g4 + i);
                    }
                    v.x = __fmul_rn(__fmul_rn(v.x, sc), gg.x);   // (x*s)*w, :462
                    v.y = __fmul_rn(__fmul_rn(v.y, sc), gg.y);
                    v.z = __fmul_rn(__fmul_rn(v.z, sc), gg.z);
                    v.w = __fmul_rn(__fmul_rn(v.w, sc), gg.w);
                    xs4w[i] = v;
                }
            }
            bar_sync_named(2, MEGA_WORKERS);
        }

        if (is_epi) {
            // ---- epilogue warp: lanes 0..3 own the four row pairs of each tile
            for (int v0 = r0; v0 < r1; v0 += GEMV8_R, ++tile_seq) {
                const int par = tile_seq & 1, use = tile_seq >> 1;
                mbar_wait(&tile_full[par], use & 1);
                float s0 = 0.0f, s1 = 0.0f;
                if (lane < GEMV8_R / 2) {
#pragma unroll
                    for (int w8 = 0; w8 < NWARP; ++w8) {       // fixed order
                        s0 += red[par][w8][2 * lane];
                        s1 += red[par][w8][2 * lane + 1];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&tile_free[par]);
                const int vp = v0 + 2 * lane;
                if (lane < GEMV8_R / 2 && vp < r1) {
                    if (kind == MK_QKV) {
                        if (vp < g.rows0 + g.rows1) {
                            const bool is_q = vp < g.rows0;
                            const int i = is_q ? vp : vp - g.rows0;
                            const int pr = (i % mp.head_size) >> 1;                                 // :338
                            const float fcr = rope_s[0][pr], fci = rope_s[1][pr];
                            const float q0 = __fsub_rn(__fmul_rn(s0, fcr), __fmul_rn(s1, fci));      // :348
                            const float q1 = __fadd_rn(__fmul_rn(s0, fci), __fmul_rn(s1, fcr));      // :349
                            float *dst = is_q ? g.out0 + i : g.kcache + (size_t)pos * g.kv_dim + i;  // :355,:357
                            *reinterpret_cast<float2 *>(dst) = make_float2(q0, q1);
                        } else {
                            const int i = vp - g.rows0 - g.rows1;
                            *reinterpret_cast<float2 *>(g.vcache + (size_t)pos * g.kv_dim + i) = make_float2(s0, s1);
                        }
                    } else if (kind == MK_W13) {
                        gemv_epilogue_pair<EPI_SILU>(g, vp, s0, s1, pos, best);
                    } else if (kind == MK_CLS) {
                        if (mp.want_argmax) gemv_epilogue_pair<EPI_ARGMAX>(g, vp, s0, s1, pos, best);
                        else gemv_epilogue_pair<EPI_STORE>(g, vp, s0, s1, pos, best);
                    } else {   // wo / w2
                        if (mp.world > 1) gemv_epilogue_pair<EPI_XCHG>(g, vp, s0, s1, pos, best);
                        else gemv_epilogue_pair<EPI_STORE>(g, vp, s0, s1, pos, best);
                    }
                }
            }
            if (kind == MK_CLS && mp.want_argmax) {
#pragma unroll
                for (int o = 2; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                    best = other > best ? other : best;
                }
                if (lane == 0 && best) atomicMax(mp.amax, best);
            }
            // ---- this CTA's part of the phase is written: publish
            if (mp.world > 1 && (kind == MK_WO || kind == MK_W2) && r1 > r0) {
                __threadfence_system();
                __syncwarp();
                if (lane < mp.world) red_release_sys_add(g.xflag_peer[lane], (unsigned int)((r1 - r0 + 1) >> 1));
            }
            __threadfence();
            __syncwarp();
            if (lane == 0)
                grid_arrive(mp.gbar, bar_base + (unsigned long long)(ph + 1) * G, gen_base + (unsigned long long)(ph + 1));
        } else {
            // ---- consumers
            const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
            float acc[GEMV8_R];
#pragma unroll
            for (int r = 0; r < GEMV8_R; ++r) acc[r] = 0.0f;
            for (int v0 = r0; v0 < r1; v0 += GEMV8_R, ++tile_seq) {
                const int rows = min(GEMV8_R, r1 - v0);
                for (int st = 0; st < nsteps; ++st) {
                    const int c = st * NT + tid;
                    mbar_wait(&full[c_stage], c_phase);
                    if (c < n4) {
                        const float4 *w4 = reinterpret_cast<const float4 *>(ring + (size_t)c_stage * TMA_STAGE_FLOATS) + tid;
                        const float4 xv = xs4[c];
#pragma unroll
                        for (int r = 0; r < GEMV8_R; ++r)
                            if (r < rows) acc[r] = dot4(w4[r * NT], xv, acc[r]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[c_stage]);
                    if (++c_stage == NSTAGE) { c_stage = 0; c_phase ^= 1; }
                }
                // transposing butterfly: lane L (L % 4 == 0) ends with row L / 4
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float send = (lane & 16) ? acc[i] : acc[i + 4];
                    const float keep = (lane & 16) ? acc[i + 4] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float send = (lane & 8) ? acc[i] : acc[i + 2];
                    const float keep = (lane & 8) ? acc[i + 2] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                {
                    const float send = (lane & 4) ? acc[0] : acc[1];
                    const float keep = (lane & 4) ? acc[1] : acc[0];
                    acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
                const int par = tile_seq & 1, use = tile_seq >> 1;
                mbar_wait(&tile_free[par], (use & 1) ^ 1);
                if ((lane & 3) == 0) red[par][warp][lane >> 2] = acc[0];
                __syncwarp();
                if (lane == 0) mbar_arrive(&tile_full[par]);
#pragma unroll
                for (int r = 0; r < GEMV8_R; ++r) acc[r] = 0.0f;
            }
        }
    }

    // ---- end of the step: CTA 0's epilogue warp closes the books once everyone is done
    if (is_epi && blockIdx.x == 0) {
        if (lane == 0) {
            grid_wait(mp.gbar, gen_base + (unsigned long long)nph);
            int *ctl = mp.ctl;
            ctl[CTL_EPOCH] = ctl[CTL_EPOCH] + 1;
            ctl[CTL_MEGA_EPOCH] = ctl[CTL_MEGA_EPOCH] + 1;
            if (mp.do_advance) {
                // src/main.zig:999-1041 at temperature 0 (same as advance_kernel)
                const int step = ctl[CTL_STEP];
                int next = (int)(0xFFFFFFFFu - (unsigned int)(*mp.amax & 0xFFFFFFFFull));
                if (mp.forced && mp.forced[step] >= 0) next = mp.forced[step];
                mp.out_next[step] = next;
                *mp.n_done = step + 1;
                *mp.amax = 0ull;
                if (ctl[CTL_STOP_ON_BOS] && next == 1) {
                    ctl[CTL_DONE] = 1;
                } else {
                    ctl[CTL_TOKEN] = next;
                    ctl[CTL_POS] = ctl[CTL_POS] + 1;
                    ctl[CTL_STEP] = step + 1;
                }
            }
        }
    }
}

}  // namespace l2b
