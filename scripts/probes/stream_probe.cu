// Micro-benchmark: how fast can 148 CTAs stream a buffer out of HBM, by access mechanism and pattern?
//   mode 0: cp.async.bulk 1-D copies into a shared-memory ring, every CTA walks its OWN contiguous region
//   mode 1: same, but chunks are dealt round-robin (CTA i takes chunks i, i+G, ...)
//   mode 2: plain 16-byte loads (no shared memory), own contiguous region, unrolled x8 per thread
//   mode 3: plain 16-byte loads, grid-stride (fully interleaved)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu ; run: ./stream_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(128) k_bulk(const char* src, size_t total, int chunk, int slots, int interleave, float* sink) {
    extern __shared__ __align__(128) char ring[];
    __shared__ __align__(8) uint64_t full[16];
    const size_t n_chunks = total / chunk;
    const size_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const size_t first = interleave ? blockIdx.x : blockIdx.x * per;
    const size_t step = interleave ? gridDim.x : 1;
    size_t mine = interleave ? (n_chunks > blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0)
                             : (first < n_chunks ? (n_chunks - first < per ? n_chunks - first : per) : 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < slots; ++s) mbar_init(smem_u32(&full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    if (threadIdx.x == 0)
        for (size_t c = 0; c < (size_t)slots && c < mine; ++c) {
            mbar_expect(smem_u32(&full[c]), chunk);
            bulk(smem_u32(ring + c * chunk), src + (first + c * step) * chunk, chunk, smem_u32(&full[c]));
        }
    for (size_t c = 0; c < mine; ++c) {
        const int s = (int)(c % slots);
        mbar_wait(smem_u32(&full[s]), (uint32_t)(c / slots) & 1u);
        acc += reinterpret_cast<const float*>(ring + (size_t)s * chunk)[threadIdx.x];     // touch the data
        __syncthreads();
        if (threadIdx.x == 0 && c + slots < mine) {
            mbar_expect(smem_u32(&full[s]), chunk);
            bulk(smem_u32(ring + (size_t)s * chunk), src + (first + (c + slots) * step) * chunk, chunk, smem_u32(&full[s]));
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

__global__ void __launch_bounds__(512) k_ldg(const float4* src, size_t n4, int interleave, float* sink) {
    float acc = 0.f;
    if (interleave) {
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7 * stride < n4; i += 8 * stride) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
        }
    } else {
        const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
        const size_t b = blockIdx.x * per, e = b + per < n4 ? b + per : n4;
        size_t i = b + threadIdx.x;
        for (; i + 7 * blockDim.x < e; i += 8 * blockDim.x) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + i + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    const size_t total = (size_t)324 << 20;
    char *buf, *flush; float* sink;
    cudaMalloc(&buf, total); cudaMalloc(&flush, (size_t)256 << 20); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, total);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto timeit = [&](const char* name, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaMemset(flush, r, (size_t)256 << 20);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 1 && ms < best) best = ms;
        }
        printf("%-56s %7.1f us  %5.2f TB/s  %s\n", name, best * 1e3, total / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    char name[128];
    for (int interleave = 0; interleave < 2; ++interleave)
        for (int chunk : {8192, 28160, 65536})
            for (int slots : {2, 4, 7})
                for (int ctas_per_sm : {1, 2}) {
                    if ((size_t)chunk * slots * ctas_per_sm > 200 * 1024) continue;
                    snprintf(name, sizeof(name), "bulk chunk %6d slots %d ctas/sm %d %s", chunk, slots, ctas_per_sm, interleave ? "interleaved" : "own region");
                    timeit(name, [&] { k_bulk<<<sms * ctas_per_sm, 128, (size_t)chunk * slots>>>(buf, total, chunk, slots, interleave, sink); });
                }
    for (int interleave = 0; interleave < 2; ++interleave)
        for (int ctas_per_sm : {1, 2, 4}) {
            snprintf(name, sizeof(name), "ldg.128 x8, 512 thr, ctas/sm %d %s", ctas_per_sm, interleave ? "grid-stride" : "own region");
            timeit(name, [&] { k_ldg<<<sms * ctas_per_sm, 512>>>(reinterpret_cast<const float4*>(buf), total / 16, interleave, sink); });
        }
    return 0;
}
