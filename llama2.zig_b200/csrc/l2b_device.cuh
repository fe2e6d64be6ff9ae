// l2b_device.cuh — sm_100a device building blocks for the llama2.zig decode step.
//
// Everything here is fp32 and HBM/L2-bandwidth bound (batch-1 decode is a vector
// contraction: 0.5 flop per weight byte), so there is deliberately no tensor-core code.
// What matters on B200 for this path: 128-bit coalesced streaming loads with many of them
// in flight per SM, the activation vector staged once per CTA into shared memory by a TMA
// bulk copy (cp.async.bulk + mbarrier), persistent CTAs sized from the SM count, and
// warp-shuffle reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace l2b {

constexpr int NT = 256;          // threads per CTA for every kernel in this file
constexpr int NWARP = NT / 32;
constexpr int GEMV_R = 2;        // weight rows per thread-group (pairs: RoPE (i,i+1), SiLU (w1,w3))
constexpr int GEMV_U = 4;        // 128-bit loads per row per chunk => R*U = 8 loads in flight/thread

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Streaming 128-bit load of immutable weights: read-only path, do not allocate in L1.
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p));
    return r;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the init visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier.
// dst, src 16-byte aligned; bytes a multiple of 16.  SASS: UBLKCP.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------
// Block-level reductions (warp shuffles + one smem hop).  Result returned to all threads.
// scratch: >= NWARP + 1 floats of shared memory.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float block_sum(float v, float *scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();  // scratch may still be read from a previous reduction
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float t = (lane < NWARP) ? scratch[lane] : 0.0f;
    t = warp_sum(t);
    return t;
}
__device__ __forceinline__ float block_max(float v, float *scratch) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float t = (lane < NWARP) ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    return t;
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}

// Orderable key for a device-side argmax with the reference's tie rule (src/main.zig:715-726:
// strict '>' so the FIRST maximum wins): larger value wins, then smaller index.
__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(idx));
}

// ---------------------------------------------------------------------------------------
// Step control block (device memory, 8 ints):
//   [0] token  [1] pos  [2] done flag (generation loop saw BOS)  [3] step index in generate
//   [4] stop-on-BOS enabled
// ---------------------------------------------------------------------------------------
enum { CTL_TOKEN = 0, CTL_POS = 1, CTL_DONE = 2, CTL_STEP = 3, CTL_STOP_ON_BOS = 4, CTL_WORDS = 8 };

// ---------------------------------------------------------------------------------------
// GEMV: out = W(rows, n) . xs(n), W row-major fp32 (src/main.zig:485-498, :530-605).
//
// Work decomposition: a group of TPR threads owns GEMV_R consecutive (virtual) rows and walks
// their columns in 128-bit steps of stride TPR, so a group reads TPR*16 contiguous bytes per
// row per step (fully coalesced, every 128-byte line used whole).  A tile is (NT/TPR)*GEMV_R
// rows; CTAs are persistent and stride over tiles.  Loads of the next chunk are issued before
// the reduction/epilogue of the current one so HBM requests stay in flight across tiles.
//
// Prologue (once per CTA): the input vector (and, when fused, the pending residual delta and
// the RMSNorm gain) are staged into shared memory by TMA bulk copies; optional fused
// rmsnorm (src/main.zig:432-468) with the reference's rounding points (x*s)*w.
// ---------------------------------------------------------------------------------------
enum GemvEpi {
    EPI_STORE = 0,   // out[v] = acc                                   (wo, w2, wcls; :392,:419,:429)
    EPI_ARGMAX = 1,  // EPI_STORE + device argmax                       (:715-726 fused, 8f.1)
    EPI_QKV = 2,     // RoPE on (even,odd) pairs + KV-cache append     (:308-358)
    EPI_SILU = 3     // hb[i] = silu(w1.x) * (w3.x)                     (:405-416)
};

struct GemvParams {
    // ---- input vector / prologue
    const float *x_in;      // n floats; when emb != nullptr: row `token` of emb is used instead
    const float *emb;       // token embedding table (layer 0: x = emb[token], :295-296) or nullptr
    const float *delta;     // pending residual (n floats) added to x_in before use, or nullptr
    const float *gamma;     // rmsnorm gain (n floats) => fused rmsnorm, or nullptr => plain staging
    float *x_out;           // if non-null, CTA 0 writes the (residual-updated, un-normalised) x here
    const int *ctl;         // control block (token, pos, done)
    int n;                  // columns (multiple of 4)
    // ---- matrices (virtual row space depends on the epilogue)
    const float *w0, *w1, *w2;
    int rows0, rows1, rows2;  // EPI_QKV: q/k/v rows.  EPI_SILU: rows0 = hidden (virtual rows = 2*hidden)
    int total_rows;           // virtual rows
    // ---- outputs
    float *out0;            // STORE: out; QKV: q; SILU: hb
    float *kcache, *vcache; // QKV: this layer's (seq_len, kv_dim) caches
    const float *rope_cos, *rope_sin;  // (seq_len, head_size/2)
    int head_size, kv_dim;
    unsigned long long *amax;  // ARGMAX: packed running maximum (must be 0 before the launch)
    int row_base;              // ARGMAX: global index of out row 0 (vocab shard offset)
};

template <int EPI>
__device__ __forceinline__ const float *gemv_row_ptr(const GemvParams &p, int v) {
    if (EPI == EPI_QKV) {
        if (v < p.rows0) return p.w0 + (size_t)v * p.n;
        v -= p.rows0;
        if (v < p.rows1) return p.w1 + (size_t)v * p.n;
        v -= p.rows1;
        return p.w2 + (size_t)v * p.n;
    } else if (EPI == EPI_SILU) {
        return ((v & 1) ? p.w1 : p.w0) + (size_t)(v >> 1) * p.n;
    } else {
        return p.w0 + (size_t)v * p.n;
    }
}

template <int TPR, int EPI>
__global__ void __launch_bounds__(NT) gemv_kernel(const GemvParams p) {
    constexpr int GROUPS = NT / TPR;           // row groups per CTA
    constexpr int TILE_ROWS = GROUPS * GEMV_R;
    constexpr int WPG = (TPR + 31) / 32;       // warps per group (TPR > 32)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);   // n floats
    float *aux = xs + p.n;                              // delta (n) then gamma (n) when fused
    __shared__ uint64_t bar;
    __shared__ float scratch[NWARP + 1];
    __shared__ float red[NWARP][GEMV_R];
    __shared__ unsigned long long blk_key;

    const int tid = threadIdx.x;
    const int grp = tid / TPR, sub = tid % TPR;
    const int n4 = p.n >> 2;
    const int ntiles = (p.total_rows + TILE_ROWS - 1) / TILE_ROWS;
    const int nchunks = (n4 + TPR * GEMV_U - 1) / (TPR * GEMV_U);

    if (p.ctl[CTL_DONE]) return;  // generation loop already ended (BOS)

    // ---- issue the TMA staging of the activation vector
    const float *xsrc = p.emb ? p.emb + (size_t)p.ctl[CTL_TOKEN] * p.n : p.x_in;
    float *ds = aux;
    float *gs = p.delta ? aux + p.n : aux;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        const uint32_t bytes = (uint32_t)p.n * 4u;
        mbar_expect_tx(&bar, bytes * (1u + (p.delta ? 1u : 0u) + (p.gamma ? 1u : 0u)));
        tma_load_1d(xs, xsrc, bytes, &bar);
        if (p.delta) tma_load_1d(ds, p.delta, bytes, &bar);
        if (p.gamma) tma_load_1d(gs, p.gamma, bytes, &bar);
    }

    // ---- first chunk of weights goes in flight before we wait for the activations
    float4 wv[GEMV_R][GEMV_U];
    int tile = blockIdx.x, chunk = 0;
    auto issue = [&](int t, int ch) {
        const int v0 = t * TILE_ROWS + grp * GEMV_R;
#pragma unroll
        for (int r = 0; r < GEMV_R; ++r) {
            const int v = v0 + r;
            const bool rok = v < p.total_rows;
            const float4 *wr =
                reinterpret_cast<const float4 *>(gemv_row_ptr<EPI>(p, rok ? v : 0));
#pragma unroll
            for (int u = 0; u < GEMV_U; ++u) {
                const int c = (ch * GEMV_U + u) * TPR + sub;
                wv[r][u] = (rok && c < n4) ? ldg_stream(wr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if (tile < ntiles) issue(tile, 0);

    // ---- wait for the staged vector, optional residual add + rmsnorm in shared memory
    __syncthreads();  // barrier init visible to all waiters
    mbar_wait(&bar, 0);
    if (p.delta || p.gamma || p.x_out) {
        float4 *xs4 = reinterpret_cast<float4 *>(xs);
        const float4 *ds4 = reinterpret_cast<const float4 *>(ds);
        float ssq = 0.0f;
        for (int i = tid; i < n4; i += NT) {
            float4 v = xs4[i];
            if (p.delta) {
                const float4 d = ds4[i];
                v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;   // accum(), :708-713
                xs4[i] = v;
            }
            if (p.x_out && blockIdx.x == 0) reinterpret_cast<float4 *>(p.x_out)[i] = v;
            ssq = fmaf(v.x, v.x, ssq); ssq = fmaf(v.y, v.y, ssq);
            ssq = fmaf(v.z, v.z, ssq); ssq = fmaf(v.w, v.w, ssq);
        }
        if (p.gamma) {
            float ss = block_sum(ssq, scratch);
            ss /= (float)p.n;            // :452
            ss += 1e-5f;                 // :453
            const float s = 1.0f / sqrtf(ss);  // :454
            const float4 *gs4 = reinterpret_cast<const float4 *>(gs);
            for (int i = tid; i < n4; i += NT) {
                float4 v = xs4[i];
                const float4 g = gs4[i];
                v.x = __fmul_rn(__fmul_rn(v.x, s), g.x);   // (x*s)*w, :462
                v.y = __fmul_rn(__fmul_rn(v.y, s), g.y);
                v.z = __fmul_rn(__fmul_rn(v.z, s), g.z);
                v.w = __fmul_rn(__fmul_rn(v.w, s), g.w);
                xs4[i] = v;
            }
        }
        __syncthreads();
    }

    const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
    const int pos = p.ctl[CTL_POS];
    unsigned long long best = 0ull;
    float acc[GEMV_R];
#pragma unroll
    for (int r = 0; r < GEMV_R; ++r) acc[r] = 0.0f;

    while (tile < ntiles) {
        // ---- consume the chunk in registers
#pragma unroll
        for (int u = 0; u < GEMV_U; ++u) {
            const int c = (chunk * GEMV_U + u) * TPR + sub;
            if (c < n4) {
                const float4 xv = xs4[c];
#pragma unroll
                for (int r = 0; r < GEMV_R; ++r) acc[r] = dot4(wv[r][u], xv, acc[r]);
            }
        }
        // ---- put the next chunk in flight
        int ntile = tile, nchunk = chunk + 1;
        if (nchunk == nchunks) { ntile = tile + gridDim.x; nchunk = 0; }
        if (ntile < ntiles) issue(ntile, nchunk);

        if (chunk == nchunks - 1) {
            // ---- reduce the GEMV_R partial dot products across the TPR threads of the group
            if (TPR <= 32) {
#pragma unroll
                for (int r = 0; r < GEMV_R; ++r) {
#pragma unroll
                    for (int o = TPR / 2; o > 0; o >>= 1)
                        acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < GEMV_R; ++r) acc[r] = warp_sum(acc[r]);
                __syncthreads();  // red[] free (previous tile's reads done)
                if ((tid & 31) == 0) {
#pragma unroll
                    for (int r = 0; r < GEMV_R; ++r) red[tid >> 5][r] = acc[r];
                }
                __syncthreads();
                if (sub == 0) {
#pragma unroll
                    for (int r = 0; r < GEMV_R; ++r) {
                        float s = 0.0f;
#pragma unroll
                        for (int w = 0; w < WPG; ++w) s += red[grp * WPG + w][r];
                        acc[r] = s;
                    }
                }
            }
            // ---- epilogue, one thread per group
            const int v0 = tile * TILE_ROWS + grp * GEMV_R;
            if (sub == 0 && v0 < p.total_rows) {
                if (EPI == EPI_STORE || EPI == EPI_ARGMAX) {
                    p.out0[v0] = acc[0];
                    if (v0 + 1 < p.total_rows) p.out0[v0 + 1] = acc[1];
                    if (EPI == EPI_ARGMAX) {
                        unsigned long long k0 = argmax_key(acc[0], p.row_base + v0);
                        best = k0 > best ? k0 : best;
                        if (v0 + 1 < p.total_rows) {
                            unsigned long long k1 = argmax_key(acc[1], p.row_base + v0 + 1);
                            best = k1 > best ? k1 : best;
                        }
                    }
                } else if (EPI == EPI_QKV) {
                    // rows (v0, v0+1) are an adjacent pair of q, k or v (segment sizes are even)
                    if (v0 < p.rows0 + p.rows1) {
                        const bool is_q = v0 < p.rows0;
                        const int i = is_q ? v0 : v0 - p.rows0;       // index within q / k
                        const int pr = (i % p.head_size) >> 1;          // :338 (i % head_size)
                        const float fcr = p.rope_cos[(size_t)pos * (p.head_size >> 1) + pr];
                        const float fci = p.rope_sin[(size_t)pos * (p.head_size >> 1) + pr];
                        const float a = acc[0], b = acc[1];
                        // :348-349, evaluated without FMA contraction like the reference
                        const float r0 = __fsub_rn(__fmul_rn(a, fcr), __fmul_rn(b, fci));
                        const float r1 = __fadd_rn(__fmul_rn(a, fci), __fmul_rn(b, fcr));
                        float *dst = is_q ? p.out0 + i : p.kcache + (size_t)pos * p.kv_dim + i; // :355,:357
                        *reinterpret_cast<float2 *>(dst) = make_float2(r0, r1);
                    } else {
                        const int i = v0 - p.rows0 - p.rows1;
                        *reinterpret_cast<float2 *>(p.vcache + (size_t)pos * p.kv_dim + i) =
                            make_float2(acc[0], acc[1]);                                     // :356,:358
                    }
                } else {  // EPI_SILU
                    const float h = acc[0];
                    const float sg = __fmul_rn(h, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-h))));  // :412
                    p.out0[v0 >> 1] = __fmul_rn(sg, acc[1]);                                    // :416
                }
            }
#pragma unroll
            for (int r = 0; r < GEMV_R; ++r) acc[r] = 0.0f;
        }
        tile = ntile;
        chunk = nchunk;
    }

    if (EPI == EPI_ARGMAX) {
        // CTA-level max, then one 64-bit atomicMax per CTA
        if (tid == 0) blk_key = 0ull;
        __syncthreads();
        if (best) atomicMax(&blk_key, best);
        __syncthreads();
        if (tid == 0 && blk_key) atomicMax(p.amax, blk_key);
    }
}

// Scalar fallback for shapes the vector kernel cannot take (n % 4 != 0 or unaligned rows):
// one warp per row.  Only reachable from l2b_op_matmul (the reference KATs use n = 3, 12).
__global__ void __launch_bounds__(NT) gemv_scalar_kernel(float *out, const float *x, const float *w,
                                                         int d, int n) {
    const int warp = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * NT) >> 5;
    for (int row = warp; row < d; row += nwarps) {
        float acc = 0.0f;
        for (int c = lane; c < n; c += 32) acc = fmaf(w[(size_t)row * n + c], x[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[row] = acc;
    }
}

// ---------------------------------------------------------------------------------------
// Per-head attention over the KV cache (src/main.zig:361-389), split along the timeline.
//
// grid = (n_heads, nsplit).  CTA (h, s) owns positions [s*chunk, min((s+1)*chunk, pos+1)):
//   scores = q_h . K[t] / sqrt(head_size)   (:367-375)  -> shared memory
//   softmax pieces: max, exp, sum           (:687-706)
//   out = sum_t p[t] * V[t]                 (:657-685)
// With one active split the normalised weights are formed first (x/sum, :703-705) exactly as
// the reference does; with several, partial (max, sum, unnormalised out) triples are merged
// by the last CTA of the head to arrive (threadfence + counter), in fixed split order, so the
// result is deterministic.
// ---------------------------------------------------------------------------------------
struct AttnParams {
    const int *ctl;
    const float *q;        // (n_heads * head_size)
    const float *kcache;   // this layer: (seq_len, kv_dim)
    const float *vcache;
    float *xb;             // (n_heads * head_size)
    float *part_o;         // (n_heads, nsplit, head_size)
    float *part_ml;        // (n_heads, nsplit, 2)
    unsigned int *counters;  // (n_heads), zero between launches
    int head_size, kv_dim, kv_mul, nsplit, min_chunk;
};

__device__ __forceinline__ int attn_lanes_per_row(int hs4) {
    return (hs4 % 8 == 0) ? 8 : (hs4 % 4 == 0) ? 4 : (hs4 % 2 == 0) ? 2 : 1;
}

// scores for positions [t0,t1) of one head into sc[0..t1-t0)
__device__ __forceinline__ void attn_scores(float *sc, const float *qs, const float *kbase,
                                            int kv_dim, int head_size, int t0, int t1) {
    const int hs4 = head_size >> 2;
    const int lpr = attn_lanes_per_row(hs4);
    const int nf = hs4 / lpr;
    const int rows_per_warp = 32 / lpr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = lane % lpr, rw = lane / lpr;
    const float4 *q4 = reinterpret_cast<const float4 *>(qs);
    const float root_hs = sqrtf((float)head_size);
    for (int tb = t0 + warp * rows_per_warp; tb < t1; tb += NWARP * rows_per_warp) {
        const int t = tb + rw;
        float acc = 0.0f;
        if (t < t1) {
            const float4 *k4 = reinterpret_cast<const float4 *>(kbase + (size_t)t * kv_dim);
            for (int f = 0; f < nf; ++f) {
                const int j = lr + f * lpr;
                acc = dot4(__ldg(k4 + j), q4[j], acc);
            }
        }
        for (int o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (t < t1 && lr == 0) sc[t - t0] = acc / root_hs;   // score /= sqrt(head_size), :372
    }
}

// out[0..head_size) (+)= sum_t w[t-t0] * V[t]; result left in red[0..head_size) (shared).
// red must hold (NT / (head_size/4)) * head_size floats.
__device__ __forceinline__ void attn_weighted_rows(float *red, const float *w, const float *vbase,
                                                   int kv_dim, int head_size, int t0, int t1) {
    const int hs4 = head_size >> 2;
    const int G = NT / hs4;
    const int tid = threadIdx.x;
    const int g = tid / hs4, c = tid % hs4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < G) {
        for (int t = t0 + g; t < t1; t += G) {
            const float wt = w[t - t0];
            const float4 v = __ldg(reinterpret_cast<const float4 *>(vbase + (size_t)t * kv_dim) + c);
            a.x = fmaf(v.x, wt, a.x); a.y = fmaf(v.y, wt, a.y);
            a.z = fmaf(v.z, wt, a.z); a.w = fmaf(v.w, wt, a.w);
        }
        reinterpret_cast<float4 *>(red)[g * hs4 + c] = a;
    }
    __syncthreads();
    // fold the G partial vectors in fixed order (thread i only touches column i: no hazard)
    if (tid < head_size) {
        float s = 0.0f;
        for (int gg = 0; gg < G; ++gg) s += red[gg * head_size + tid];
        red[tid] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT) attention_kernel(const AttnParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float scratch[NWARP + 1];
    __shared__ int is_last;
    if (p.ctl[CTL_DONE]) return;

    const int h = blockIdx.x, s = blockIdx.y;
    const int hs = p.head_size;
    const int T = p.ctl[CTL_POS] + 1;
    int chunk = (T + p.nsplit - 1) / p.nsplit;
    if (chunk < p.min_chunk) chunk = p.min_chunk;
    const int active = (T + chunk - 1) / chunk;
    if (s >= active) return;
    const int t0 = s * chunk;
    const int t1 = min(T, t0 + chunk);
    const int len = t1 - t0;

    // shared layout: q[hs] | red[G*hs] | sc[chunk_cap]
    const int hs4 = hs >> 2;
    const int G = NT / hs4;
    float *qs = reinterpret_cast<float *>(smem_raw);
    float *red = qs + hs;
    float *sc = red + G * hs;

    const int tid = threadIdx.x;
    const size_t hoff = (size_t)(h / p.kv_mul) * hs;   // :369, :382
    for (int i = tid; i < hs; i += NT) qs[i] = p.q[(size_t)h * hs + i];
    __syncthreads();

    attn_scores(sc, qs, p.kcache + hoff, p.kv_dim, hs, t0, t1);
    __syncthreads();

    // softmax pieces over this split (:690-705)
    float m = -INFINITY;
    for (int i = tid; i < len; i += NT) m = fmaxf(m, sc[i]);
    m = block_max(m, scratch);
    float l = 0.0f;
    for (int i = tid; i < len; i += NT) {
        const float e = expf(sc[i] - m);
        sc[i] = e;
        l += e;
    }
    l = block_sum(l, scratch);
    if (active == 1) {
        for (int i = tid; i < len; i += NT) sc[i] = sc[i] / l;   // :703-705
    }
    __syncthreads();

    attn_weighted_rows(red, sc, p.vcache + hoff, p.kv_dim, hs, t0, t1);

    if (active == 1) {
        if (tid < hs) p.xb[(size_t)h * hs + tid] = red[tid];
        return;
    }

    // ---- publish the partial, last arriver merges
    float *po = p.part_o + ((size_t)h * p.nsplit + s) * hs;
    if (tid < hs) po[tid] = red[tid];
    if (tid == 0) {
        p.part_ml[((size_t)h * p.nsplit + s) * 2 + 0] = m;
        p.part_ml[((size_t)h * p.nsplit + s) * 2 + 1] = l;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(&p.counters[h], 1u);
        is_last = (prev == (unsigned int)(active - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) p.counters[h] = 0u;  // ready for the next launch
    const volatile float *ml = p.part_ml + (size_t)h * p.nsplit * 2;
    float M = -INFINITY;
    for (int j = 0; j < active; ++j) M = fmaxf(M, ml[j * 2]);
    float Lsum = 0.0f;
    for (int j = 0; j < active; ++j) Lsum += expf(ml[j * 2] - M) * ml[j * 2 + 1];
    if (tid < hs) {
        const volatile float *pb = p.part_o + (size_t)h * p.nsplit * hs;
        float o = 0.0f;
        for (int j = 0; j < active; ++j) o = fmaf(expf(ml[j * 2] - M), pb[(size_t)j * hs + tid], o);
        p.xb[(size_t)h * hs + tid] = o / Lsum;
    }
}

// ---------------------------------------------------------------------------------------
// Small kernels
// ---------------------------------------------------------------------------------------
__global__ void set_ctl_kernel(int *ctl, int token, int pos, int stop_on_bos,
                               unsigned long long *amax) {
    ctl[CTL_TOKEN] = token;
    ctl[CTL_POS] = pos;
    ctl[CTL_DONE] = 0;
    ctl[CTL_STEP] = 0;
    ctl[CTL_STOP_ON_BOS] = stop_on_bos;
    *amax = 0ull;
}

// End of one step of the on-device temperature-0 loop (src/main.zig:999-1041):
// choose next = forced[step] or argmax, record it, stop on BOS, advance (token, pos).
__global__ void advance_kernel(int *ctl, unsigned long long *amax, const int *forced, int *out_next,
                               int *n_done) {
    if (ctl[CTL_DONE]) return;
    const int step = ctl[CTL_STEP];
    int next = (int)(0xFFFFFFFFu - (unsigned int)(*amax & 0xFFFFFFFFull));
    if (forced && forced[step] >= 0) next = forced[step];
    out_next[step] = next;
    *n_done = step + 1;
    *amax = 0ull;
    if (ctl[CTL_STOP_ON_BOS] && next == 1) {   // :1017-1019
        ctl[CTL_DONE] = 1;
        return;
    }
    ctl[CTL_TOKEN] = next;
    ctl[CTL_POS] = ctl[CTL_POS] + 1;
    ctl[CTL_STEP] = step + 1;
}

// standalone rmsnorm (unit-test surface; the hot path fuses it into the GEMV prologue)
__global__ void __launch_bounds__(NT) rmsnorm_kernel(float *o, const float *x, const float *w, int n) {
    __shared__ float scratch[NWARP + 1];
    float ssq = 0.0f;
    for (int i = threadIdx.x; i < n; i += NT) ssq = fmaf(x[i], x[i], ssq);
    float ss = block_sum(ssq, scratch);
    ss /= (float)n;
    ss += 1e-5f;
    const float s = 1.0f / sqrtf(ss);
    for (int i = threadIdx.x; i < n; i += NT) o[i] = __fmul_rn(__fmul_rn(x[i], s), w[i]);
}

// standalone softmax (:687-706) over n values in global memory, one CTA (unit-test surface)
__global__ void __launch_bounds__(NT) softmax_kernel(float *x, int n) {
    __shared__ float scratch[NWARP + 1];
    float m = -INFINITY;
    for (int i = threadIdx.x; i < n; i += NT) m = fmaxf(m, x[i]);
    m = block_max(m, scratch);
    float l = 0.0f;
    for (int i = threadIdx.x; i < n; i += NT) {
        const float e = expf(x[i] - m);
        x[i] = e;
        l += e;
    }
    l = block_sum(l, scratch);
    for (int i = threadIdx.x; i < n; i += NT) x[i] = x[i] / l;
}

// standalone vector_weighted_sum_rows (:657-685), arbitrary out_len / stride (unit-test surface)
__global__ void __launch_bounds__(NT) weighted_rows_kernel(float *xout, int out_len, const float *rows,
                                                           int row_stride, const float *weights,
                                                           int n_weights) {
    for (int i = blockIdx.x * NT + threadIdx.x; i < out_len; i += gridDim.x * NT) {
        float s = 0.0f;
        for (int r = 0; r < n_weights; ++r) s = fmaf(rows[(size_t)r * row_stride + i], weights[r], s);
        xout[i] = s;
    }
}

// ---------------------------------------------------------------------------------------
// Synthetic weights: counter-based, integer-only hash so CPU and GPU agree bit for bit.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ int32_t synth_irwin_hall(uint64_t tensor_seed, uint64_t i) {
    const uint64_t h = mix64(tensor_seed + (i + 1) * 0x9E3779B97F4A7C15ull);
    return (int32_t)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + ((h >> 48) & 0xFFFF)) -
           131070;
}
// rows x cols window of a (.., src_cols) tensor starting at element `first`, written densely
__global__ void synth_fill_kernel(float *dst, uint64_t rows, uint64_t cols, uint64_t first,
                                  uint64_t src_cols, uint64_t tensor_seed, double mean, double scale,
                                  float lo, float hi) {
    const uint64_t total = rows * cols;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total;
         j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = j / cols, c = j - r * cols;
        const uint64_t i = first + r * src_cols + c;
        const int32_t s = synth_irwin_hall(tensor_seed, i);
        float v = (float)__dadd_rn(__dmul_rn((double)s, scale), mean);
        v = fminf(fmaxf(v, lo), hi);
        dst[j] = v;
    }
}

}  // namespace l2b
