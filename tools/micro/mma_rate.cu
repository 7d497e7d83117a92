// Micro-benchmark (round 2, late): issue rate of tcgen05.mma.kind::f16 (M = 128, K = 16, cta_group::1) on ONE SM as a function
// of N, of the shared-memory operand layout (K-major no-swizzle core matrices -- what conv_tma / attention_tc / vit_linear
// use -- against K-major SWIZZLE_128B) and of where A comes from (shared memory against tensor memory).
// It answers what bounds the fp16x2 convs: "an M = 128 MMA with both operands in shared memory costs ~64
// cycles whatever N is".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/micro/mma_rate.cu && gpurun_out/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// mode 0: A, B in shared memory, K-major no swizzle (LBO = stride between the two 8-channel groups of a K = 16 step, SBO = 128 B)
// mode 1: A, B in shared memory, K-major SWIZZLE_128B (rows of 128 B, SBO = 1024 B; the K = 16 step is 32 B inside the row)
// mode 2: A in tensor memory (columns 256..263), B in shared memory no swizzle
// shift: extra 16-byte rows added to the A start address per MMA (0, or 1 = the "shifted window" of a 3x3 tap)
// n_acc: accumulators the MMAs rotate over (1 = every MMA accumulates onto the previous one's result: a dependent chain)
// n_issue: issuing threads (lane 0 of the first n_issue warps), each with its own accumulator(s) and its own commit barrier
__global__ void __launch_bounds__(128) rate_kernel(int mode, int N, int n_mma, int shift, int n_acc, int n_issue, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);  // [4]
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + 64);
    uint8_t *sA = smem + 1024;             // 64 KB region
    uint8_t *sB = smem + 1024 + 65536;     // 64 KB region
    for (int i = threadIdx.x * 16; i < 2 * 65536; i += 128 * 16) *reinterpret_cast<uint4 *>(sA + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int w = 0; w < 4; ++w) mbar_init(bar + w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    const int w = threadIdx.x >> 5;
    __shared__ long long t_begin[4], t_end[4];
    if ((threadIdx.x & 31) == 0 && w < n_issue) {
        const uint32_t idesc = (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);  // f16 x f16 -> f32, K-major both
        uint64_t a_desc, b_desc;
        if (mode == 1) {
            // SWIZZLE_128B: layout type 2 at bits 61..63, LBO unused (1), SBO = 1024 B
            const uint64_t hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61) | (uint64_t(1) << 16);
            a_desc = hi | uint64_t((smem_u32(sA) & 0x3FFFF) >> 4);
            b_desc = hi | uint64_t((smem_u32(sB) & 0x3FFFF) >> 4);
        } else {
            // no swizzle: plane (8-channel group) stride = 512 rows of 16 B = 8 KB (a conv window of ~500 positions)
            const uint64_t hi = (uint64_t(8) << 32) | (uint64_t(1) << 46) | (uint64_t(512) << 16);
            a_desc = hi | uint64_t((smem_u32(sA) & 0x3FFFF) >> 4);
            b_desc = hi | uint64_t((smem_u32(sB) & 0x3FFFF) >> 4);
        }
        const uint32_t a_tmem = tmem + 496;  // A operand in tensor memory: 8 columns behind every accumulator
        const uint32_t d0 = tmem + uint32_t(w * n_acc * N);
        for (int i = 0; i < 48; ++i) {  // warm-up
            if (mode == 2) mma_ts(d0, a_tmem, b_desc, idesc, 1u);
            else mma_ss(d0, a_desc, b_desc, idesc, 1u);
        }
        commit(bar + w);
        mbar_wait(bar + w, 0);
        // 12 MMAs per loop iteration, operands precomputed: the issuing thread's own instructions must not be what is measured
        const uint64_t a3[3] = {a_desc, a_desc + uint64_t(shift), a_desc + uint64_t(2 * shift)};
        uint32_t dj[12];
        for (int u = 0; u < 12; ++u) dj[u] = d0 + uint32_t((u % n_acc) * N);
        t_begin[w] = clock64();
        for (int i = 0; i < n_mma; i += 12) {
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                if (mode == 2) mma_ts(dj[u], a_tmem, b_desc, idesc, 1u);
                else mma_ss(dj[u], a3[u % 3], b_desc, idesc, 1u);
            }
        }
        commit(bar + w);
        mbar_wait(bar + w, 1);
        t_end[w] = clock64();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b0 = t_begin[0], e0 = t_end[0];
        for (int k = 1; k < n_issue; ++k) {
            b0 = t_begin[k] < b0 ? t_begin[k] : b0;
            e0 = t_end[k] > e0 ? t_end[k] : e0;
        }
        out[0] = (unsigned long long)(e0 - b0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    unsigned long long *d, h;
    cudaMalloc(&d, 8);
    const size_t smem = 1024 + 2 * 65536;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char *names[3] = {"A,B smem  K-major no-swizzle", "A,B smem  K-major SWIZZLE_128B", "A tmem, B smem no-swizzle"};
    const int n_mma = 4092;
    printf("tcgen05.mma.cta_group::1.kind::f16, M = 128, K = 16: cycles per MMA on one SM\n");
    printf("(math floor at 8192 dense fp16 FLOP/clk/SM: N / 2 cycles)\n");
    for (int n_issue = 1; n_issue <= 4; n_issue *= 2) {
        printf("-- %d issuing thread(s), %d MMAs each, one accumulator per thread (N <= 120 so that four fit)\n", n_issue, n_mma);
        for (int mode = 0; mode < 3; ++mode)
            for (int shift = 0; shift <= (mode == 0 ? 1 : 0); ++shift) {
                printf("%-32s%s:", names[mode], shift ? " (A start +0/+1/+2 rows per MMA)" : "");
                for (int N : {16, 32, 64, 96, 120}) {
                    rate_kernel<<<1, 128, smem>>>(mode, N, n_mma, shift, 1, n_issue, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf(" N=%d: %s\n", N, cudaGetErrorString(e));
                        return 1;
                    }
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    printf("  N=%d: %.1f", N, double(h) / (double(n_mma) * n_issue));
                }
                printf("  cycles per MMA (all threads together)\n");
            }
    }
    printf("-- 1 issuing thread, one accumulator, large N\n");
    for (int mode = 0; mode < 3; mode += 2) {
        printf("%-32s:", names[mode]);
        for (int N : {128, 192, 256}) {
            rate_kernel<<<1, 128, smem>>>(mode, N, n_mma, 0, 1, 1, d);
            if (cudaDeviceSynchronize() != cudaSuccess) return 1;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("  N=%d: %.1f", N, double(h) / n_mma);
        }
        printf("  cycles per MMA\n");
    }
    return 0;
}
