// read_bw.cu — ground truth for the GEMV design: what pure-READ HBM bandwidth can one B200 reach,
// and with which access machinery?  (MEASURED_PEAKS.json's 6.58 TB/s is a copy: read + write.)
//   A. LDG.128 streaming (ld.global.nc.L1::no_allocate), U loads in flight per thread
//   B. same with default caching
//   C. TMA 1-D bulk copies (cp.async.bulk) into a shared-memory ring, consumers sum from smem
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a read_bw.cu -o read_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ld_nc(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_def(const float4 *p) { return *p; }

// window mode: the grid sweeps the buffer as one moving contiguous window (memcpy-like)
template <int U, bool NC>
__global__ void k_ldg_window(const float4 *__restrict__ src, size_t n4, float *out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = NC ? ld_nc(src + i + u * stride) : ld_def(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { float4 v = ld_nc(src + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

// range mode: every CTA owns one contiguous slice (like the contiguous-row-range GEMV)
template <int U>
__global__ void k_ldg_range(const float4 *__restrict__ src, size_t n4, float *out) {
    float acc = 0.f;
    const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const size_t b = per * blockIdx.x, e = (b + per < n4) ? b + per : n4;
    size_t i = b + threadIdx.x;
    const size_t stride = blockDim.x;
    for (; i + (U - 1) * stride < e; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_nc(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < e; i += stride) { float4 v = ld_nc(src + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

// ---- TMA ring
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma1d(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}

// one producer thread per CTA streams STAGE_BYTES chunks (window order) into an NSTAGE ring
template <int NSTAGE, int STAGE_BYTES>
__global__ void k_tma_ring(const char *__restrict__ src, size_t bytes, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[NSTAGE], empty[NSTAGE];
    const int tid = threadIdx.x;
    const int nconsumers = blockDim.x - 32;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], nconsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = bytes / STAGE_BYTES;
    float acc = 0.f;
    if (tid < 32) {
        if (tid == 0) {
            int s = 0; uint32_t ph = 0;
            for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
                mb_wait(&empty[s], ph ^ 1);
                mb_expect(&full[s], STAGE_BYTES);
                tma1d(smem + (size_t)s * STAGE_BYTES, src + c * STAGE_BYTES, STAGE_BYTES, &full[s]);
                if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
        }
    } else {
        const int ct = tid - 32;
        int s = 0; uint32_t ph = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            mb_wait(&full[s], ph);
            const float4 *p = reinterpret_cast<const float4 *>(smem + (size_t)s * STAGE_BYTES);
            for (int i = ct; i < STAGE_BYTES / 16; i += nconsumers) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
            mb_arrive(&empty[s]);
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float time_it(F f, int iters = 5) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t bytes = (size_t)2048 << 20;   // 2 GiB >> L2
    char *buf; float *out;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(buf, 1, bytes));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    const size_t n4 = bytes / 16;
    printf("device %s, %d SMs, buffer %zu MiB\n", pr.name, sms, bytes >> 20);
    // reference: cudaMemcpy D2D of half the buffer (read+write)
    {
        float ms = time_it([&] { CK(cudaMemcpyAsync(buf, buf + bytes / 2, bytes / 2, cudaMemcpyDeviceToDevice)); });
        printf("memcpy d2d            : %7.1f GB/s (read+write)\n", (double)bytes / ms / 1e6);
    }
#define RUN_W(U, NC, tpb, bps) { float ms = time_it([&] { k_ldg_window<U, NC><<<sms * bps, tpb>>>((const float4 *)buf, n4, out); }); \
    printf("ldg window U=%2d nc=%d tpb=%4d cta/sm=%d : %7.1f GB/s   (%.0f KB in flight/SM)\n", U, NC, tpb, bps, (double)bytes / ms / 1e6, U * 16.0 * tpb * bps / 1024); }
    RUN_W(4, true, 256, 4) RUN_W(8, true, 256, 4) RUN_W(8, true, 256, 2) RUN_W(16, true, 256, 2) RUN_W(16, true, 256, 4)
    RUN_W(8, true, 512, 2) RUN_W(8, true, 1024, 1) RUN_W(8, true, 256, 8) RUN_W(4, true, 256, 8) RUN_W(2, true, 256, 8)
    RUN_W(8, false, 256, 4) RUN_W(16, false, 256, 2)
#define RUN_R(U, tpb, bps) { float ms = time_it([&] { k_ldg_range<U><<<sms * bps, tpb>>>((const float4 *)buf, n4, out); }); \
    printf("ldg range  U=%2d      tpb=%4d cta/sm=%d : %7.1f GB/s\n", U, tpb, bps, (double)bytes / ms / 1e6); }
    RUN_R(8, 256, 4) RUN_R(16, 256, 2) RUN_R(8, 256, 2)
#define RUN_T(NS, SB, tpb, bps) { CK(cudaFuncSetAttribute(k_tma_ring<NS, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * SB)); \
    float ms = time_it([&] { k_tma_ring<NS, SB><<<sms * bps, tpb, NS * SB>>>(buf, bytes, out); }); \
    printf("tma ring stages=%d stage=%3dKB tpb=%4d cta/sm=%d : %7.1f GB/s   (%d KB in flight/SM)\n", NS, SB / 1024, tpb, bps, (double)bytes / ms / 1e6, NS * SB * bps / 1024); }
    RUN_T(4, 16384, 288, 1) RUN_T(4, 32768, 288, 1) RUN_T(6, 32768, 288, 1) RUN_T(3, 65536, 288, 1) RUN_T(4, 16384, 288, 2)
    RUN_T(4, 32768, 160, 1) RUN_T(8, 8192, 288, 2) RUN_T(4, 8192, 288, 4) RUN_T(6, 32768, 544, 1)
    return 0;
}
