// microbenchmark: HBM read bandwidth of per-WARP 1-D bulk copies (cp.async.bulk) of CHUNK bytes, STAGES deep,
// against per-CTA copies of 8*CHUNK.  Each lane xors its share of every chunk (6 x LDS.128 per 3 KB) so the data is used.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(256) k_warp_stream(const uint8_t* text, size_t per_warp, uint32_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* buf = smem + (size_t)warp * STAGES * CHUNK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)8 * STAGES * CHUNK) + warp * STAGES;
    const uint8_t* src = text + ((size_t)blockIdx.x * 8 + warp) * per_warp;
    const uint32_t n = (uint32_t)(per_warp / CHUNK);
    if (lane == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (lane == 0) for (int s = 0; s < STAGES && s < (int)n; ++s) { mbar_expect(&bars[s], CHUNK); bulk(buf + s * CHUNK, src + (size_t)s * CHUNK, CHUNK, &bars[s]); }
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const int s = i % STAGES;
        mbar_wait(&bars[s], (i / STAGES) & 1);
        const uint4* p = reinterpret_cast<const uint4*>(buf + s * CHUNK);
#pragma unroll
        for (int j = 0; j < CHUNK / 512; ++j) { uint4 v = p[j * 32 + lane]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        __syncwarp();
        if (lane == 0 && i + STAGES < n) { mbar_expect(&bars[s], CHUNK); bulk(buf + s * CHUNK, src + (size_t)(i + STAGES) * CHUNK, CHUNK, &bars[s]); }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(256) k_cta_stream(const uint8_t* text, size_t per_cta, uint32_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK);
    const uint8_t* src = text + (size_t)blockIdx.x * per_cta;
    const uint32_t n = (uint32_t)(per_cta / CHUNK);
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid == 0) for (int s = 0; s < STAGES && s < (int)n; ++s) { mbar_expect(&bars[s], CHUNK); bulk(smem + s * CHUNK, src + (size_t)s * CHUNK, CHUNK, &bars[s]); }
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const int s = i % STAGES;
        mbar_wait(&bars[s], (i / STAGES) & 1);
        const uint4* p = reinterpret_cast<const uint4*>(smem + s * CHUNK);
#pragma unroll
        for (int j = 0; j < CHUNK / 4096; ++j) { uint4 v = p[j * 256 + tid]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        __syncthreads();
        if (tid == 0 && i + STAGES < n) { mbar_expect(&bars[s], CHUNK); bulk(smem + s * CHUNK, src + (size_t)(i + STAGES) * CHUNK, CHUNK, &bars[s]); }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <class K>
void run(const char* name, K kern, int grid, size_t smem, const uint8_t* d, size_t per, uint32_t* out, size_t total) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; ++w) kern<<<grid, 256, smem>>>(d, per, out);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) kern<<<grid, 256, smem>>>(d, per, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s occ %d CTAs/SM  %8.1f GB/s  (%s)\n", name, occ, total * 5 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const size_t total = (size_t)6 << 30;
    uint8_t* d; uint32_t* out; cudaMalloc(&d, total + 4096); cudaMalloc(&out, 4); cudaMemset(d, 1, total);
    const int sms = 148;
    { const int grid = sms * 4; size_t per = total / ((size_t)grid * 8) / 3072 * 3072; run("warp 3072 x2, 4 CTA/SM", k_warp_stream<3072, 2>, grid, 8 * 2 * 3072 + 8 * 2 * 8, d, per, out, per * grid * 8); }
    { const int grid = sms * 4; size_t per = total / ((size_t)grid * 8) / 1536 * 1536; run("warp 1536 x4, 4 CTA/SM", k_warp_stream<1536, 4>, grid, 8 * 4 * 1536 + 8 * 4 * 8, d, per, out, per * grid * 8); }
    { const int grid = sms * 4; size_t per = total / ((size_t)grid * 8) / 3072 * 3072; run("warp 3072 x3 (3 CTA/SM)", k_warp_stream<3072, 3>, sms * 3, 8 * 3 * 3072 + 8 * 3 * 8, d, per, out, per * sms * 3 * 8); }
    { const int grid = sms * 4; size_t per = total / grid / 24576 * 24576; run("cta 24576 x2, 4 CTA/SM", k_cta_stream<24576, 2>, grid, 2 * 24576 + 64, d, per, out, per * grid); }
    { const int grid = sms * 2; size_t per = total / grid / 49152 * 49152; run("cta 49152 x2, 2 CTA/SM", k_cta_stream<49152, 2>, grid, 2 * 49152 + 64, d, per, out, per * grid); }
    return 0;
}
