// Micro-benchmark (round 2): how fast can every CTA of a grid stream the SAME L2-resident buffer (conv weights) into its
// shared memory with cp.async.bulk, and what does multicast over a thread-block cluster change?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/stream_bench tools/micro/stream_bench.cu && gpurun_out/stream_bench
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_mc(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// remote arrive on the same barrier offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *b, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

constexpr int NS = 4;
// SPLIT: bulk copies per chunk; CS: cluster size (1 = no multicast)
template <int CS>
__global__ void stream_kernel(const uint8_t *w, int n_chunks, int chunk_bytes, int split, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);       // [NS]
    uint64_t *empty = full + NS;                                 // [NS]: all CTAs of the cluster have consumed the stage
    uint8_t *buf = smem + 128;
    const int tid = threadIdx.x;
    uint32_t rank = 0;
    if (CS > 1) rank = cg::this_cluster().block_rank();
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, CS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
    if (tid == 0) {  // producer
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NS;
            if (c >= NS) mbar_wait(empty + s, ((c / NS) - 1) & 1);
            mbar_expect(full + s, chunk_bytes);
            const uint8_t *src = w + size_t(c) * chunk_bytes;
            if (CS == 1) {
                const int part = chunk_bytes / split;
                for (int k = 0; k < split; ++k) bulk(buf + size_t(s) * chunk_bytes + k * part, src + k * part, part, full + s);
            } else {  // every rank fetches 1/CS of the chunk and multicasts it to all
                const int part = chunk_bytes / CS;
                bulk_mc(buf + size_t(s) * chunk_bytes + rank * part, src + rank * part, part, full + s, uint16_t((1u << CS) - 1));
            }
        }
    } else if (tid == 32) {  // consumer: touch the data, then release the stage in every CTA of the cluster
        unsigned long long acc = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NS;
            mbar_wait(full + s, (c / NS) & 1);
            acc += *reinterpret_cast<unsigned long long *>(buf + size_t(s) * chunk_bytes + 64);
            if (CS == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
            else for (uint32_t r = 0; r < CS; ++r) mbar_arrive_remote(empty + s, r);
        }
        if (acc == 0x1234567) *sink = acc;
    }
    if (CS > 1) cg::this_cluster().sync();
}

template <int CS>
float run(const uint8_t *w, int grid, int n_chunks, int chunk_bytes, int split, unsigned long long *sink) {
    size_t smem = 128 + size_t(NS) * chunk_bytes;
    cudaFuncSetAttribute(stream_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaLaunchKernelEx(&cfg, stream_kernel<CS>, w, n_chunks, chunk_bytes, split, sink);
    cudaEventRecord(e0);
    const int it = 20;
    for (int i = 0; i < it; ++i) cudaLaunchKernelEx(&cfg, stream_kernel<CS>, w, n_chunks, chunk_bytes, split, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
    return ms / it * 1e3f;
}

int main() {
    const int chunk = 36864, n_chunks = 16;  // the fp16x2 128->128 3x3 conv at 8x8: 16 chunks of 36 KB per CTA
    uint8_t *w; unsigned long long *sink;
    cudaMalloc(&w, size_t(chunk) * n_chunks); cudaMemset(w, 1, size_t(chunk) * n_chunks); cudaMalloc(&sink, 8);
    for (int grid : {8, 32, 64, 128}) {
        float a = run<1>(w, grid, n_chunks, chunk, 1, sink);
        float b = run<1>(w, grid, n_chunks, chunk, 4, sink);
        float c2 = run<2>(w, grid, n_chunks, chunk, 1, sink);
        float c4 = run<4>(w, grid, n_chunks, chunk, 1, sink);
        float c8 = run<8>(w, grid, n_chunks, chunk, 1, sink);
        const double kb = chunk * n_chunks / 1e3;
        printf("grid %3d: %5.0f KB per CTA | plain %6.1f us (%5.1f GB/s per CTA) | 4 copies/chunk %6.1f us | multicast x2 %6.1f us | x4 %6.1f us | x8 %6.1f us\n",
               grid, kb, a, kb / a / 1e3 * 1e3, b, c2, c4, c8);
    }
    // half-size chunks (bf16 mode)
    for (int grid : {128}) {
        float a = run<1>(w, grid, n_chunks, chunk / 2, 1, sink);
        float c4 = run<4>(w, grid, n_chunks, chunk / 2, 1, sink);
        printf("grid %3d, 18 KB chunks: plain %6.1f us | multicast x4 %6.1f us\n", grid, a, c4);
    }
    return 0;
}
