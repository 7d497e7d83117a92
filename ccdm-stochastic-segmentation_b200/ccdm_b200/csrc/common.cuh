// Shared device/host helpers of libccdm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/ccdm_b200.h"

namespace ccdm {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
#define CCDM_FAIL(code, ...)            \
    do {                                \
        ::ccdm::set_error(__VA_ARGS__); \
        return (code);                  \
    } while (0)
#define CCDM_CUDA(expr)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) CCDM_FAIL(-100, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)
#define CCDM_LAUNCH_CHECK(name)                                                              \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) CCDM_FAIL(-101, "launch %s: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// ---- internal launchers (one per op kind) -------------------------------------
int launch_conv(const ccdm_op &op, cudaStream_t s);
int launch_attention(const ccdm_op &op, cudaStream_t s);
int launch_head(const ccdm_op &op, cudaStream_t s);
int launch_encode_input(const ccdm_op &op, cudaStream_t s);

// ---- programmatic dependent launch (PDL) -----------------------------------------
// Every kernel of the reverse step is launched with programmatic stream serialization: the next kernel's
// CTAs may be scheduled (and run their prologue: barrier init, TMEM allocation, weight loads) while the
// previous kernel drains, and block in pdl_wait() until it has COMPLETED and its writes are visible.  A
// kernel must call pdl_wait() before its first access to anything another kernel of the step writes or
// reads.  Captured into the step's CUDA graph as programmatic edges.  Opt-in with CCDM_PDL=1.
bool pdl_enabled();
template <typename Kern, typename... Args>
cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr float kGnEps = 1e-5f;  // nn.GroupNorm default (nn.py:93-100 -> GroupNorm32(32, C))
constexpr int kGnGroups = 32;

// ---- storage-type helpers -----------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 load4(const T *p);
template <>
__device__ __forceinline__ float4 load4<float>(const float *p) {
    return *reinterpret_cast<const float4 *>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    uint2 raw = *reinterpret_cast<const uint2 *>(p);
    __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162 *>(&raw.x);
    __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162 *>(&raw.y);
    float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}
// Stores 4 values and returns them as stored (bf16 rounding applied), so that
// GroupNorm statistics describe exactly what the consumer will read back.
template <typename T>
__device__ __forceinline__ float4 store4(T *p, float4 v);
template <>
__device__ __forceinline__ float4 store4<float>(float *p, float4 v) {
    *reinterpret_cast<float4 *>(p) = v;
    return v;
}
template <>
__device__ __forceinline__ float4 store4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t *>(&lo);
    raw.y = *reinterpret_cast<uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(p) = raw;
    float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
// exact-mode SiLU: x * sigmoid(x) with the accurate expf (unet.py:190 nn.SiLU)
__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// ---- Philox4x32-10 (Salmon et al., SC'11; Random123 constants) -----------------
// counter = (pixel in sample, global sample, draw index, class block), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// bits -> Exp(1): u = ((bits >> 9) + 0.5) * 2^-23 is exact in fp32 and in (0,1).
// approximate base-2 exponential / logarithm (MUFU), for the non-exact engine mode only
__device__ __forceinline__ float exp2f_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float log2f_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float bits_to_exponential(uint32_t bits) {
    float u = (static_cast<float>(bits >> 9) + 0.5f) * 1.1920928955078125e-07f;
    return -logf(u);
}

}  // namespace ccdm
