// QKVAttentionLegacy (reference unet.py:343-360) as a flash-style fp32 kernel:
// softmax((q*s)(k*s)^T) v per (sample, head), s = head_dim^-1/4, never
// materialising the [T,T] score matrix.  Exact ("parity") implementation; one
// thread owns one query row, keys/values stream through shared memory in tiles
// of 32 and the softmax is the usual running-max / running-sum recurrence.
//
// qkv layout (NHWC, tokens = pixels): [B, T, 3C] with the reference's legacy
// channel order -- heads first, then q|k|v inside a head (unet.py:353):
//   channel(h, which, d) = h*3*D + which*D + d.
// Output: [B, T, C], channel(h, d) = h*D + d  (unet.py:360).
#include "common.cuh"

namespace ccdm {
namespace {

constexpr int QT = 64;   // queries per CTA (= threads)
constexpr int KT = 32;   // keys per smem tile

template <typename T, int D>
__global__ void __launch_bounds__(QT) attention_ffma_kernel(const T *__restrict__ qkv, T *__restrict__ out, int Ttok,
                                                            int heads, float scale) {
    __shared__ float sK[KT][D];
    __shared__ float sV[KT][D];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * QT + threadIdx.x;
    const int C3 = heads * 3 * D;
    // fp32: token-major [B, T, 3C]; bf16: plane-major [B][3C/8][T][8] (see conv_ffma.cu act_off)
    constexpr bool PM = sizeof(T) == 2;
    auto qoff = [&](int tok, int ch) -> size_t {  // ch = channel inside this head's q|k|v block of 3*D
        const int c = h * 3 * D + ch;
        return PM ? ((size_t(b) * (C3 >> 3) + (c >> 3)) * Ttok + tok) * 8 + (c & 7) : (size_t(b) * Ttok + tok) * C3 + c;
    };

    float q[D], acc[D];
    const bool active = qi < Ttok;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
        float4 v = active ? load4<T>(qkv + qoff(qi, d)) : make_float4(0.f, 0.f, 0.f, 0.f);
        q[d] = v.x * scale; q[d + 1] = v.y * scale; q[d + 2] = v.z * scale; q[d + 3] = v.w * scale;
        acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
    }
    float m = -INFINITY, l = 0.f;

    for (int k0 = 0; k0 < Ttok; k0 += KT) {
        // stage k (scaled, as the reference scales k before the product) and v
        for (int e = threadIdx.x; e < KT * (2 * D / 4); e += QT) {
            int j = e / (2 * D / 4), f = (e % (2 * D / 4)) * 4;  // f in [0, 2D): k then v
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + j < Ttok) v = load4<T>(qkv + qoff(k0 + j, D + f));
            if (f < D) {
                sK[j][f] = v.x * scale; sK[j][f + 1] = v.y * scale; sK[j][f + 2] = v.z * scale; sK[j][f + 3] = v.w * scale;
            } else {
                sV[j][f - D] = v.x; sV[j][f - D + 1] = v.y; sV[j][f - D + 2] = v.z; sV[j][f - D + 3] = v.w;
            }
        }
        __syncthreads();
        float s[KT];
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            float d0 = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) d0 += q[d] * sK[j][d];
            s[j] = (k0 + j < Ttok) ? d0 : -INFINITY;
            tmax = fmaxf(tmax, s[j]);
        }
        const float mnew = fmaxf(m, tmax);
        const float corr = expf(m - mnew);  // m = -inf on the first tile -> 0
        l *= corr;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] *= corr;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const float pj = expf(s[j] - mnew);  // -inf -> 0 for padded keys
            l += pj;
#pragma unroll
            for (int d = 0; d < D; ++d) acc[d] += pj * sV[j][d];
        }
        m = mnew;
        __syncthreads();
    }
    if (active) {
        const float inv = 1.0f / l;
        const int C = heads * D;
#pragma unroll
        for (int d = 0; d < D; d += 4) {
            const int c = h * D + d;
            T *o = out + (PM ? ((size_t(b) * (C >> 3) + (c >> 3)) * Ttok + qi) * 8 + (c & 7) : (size_t(b) * Ttok + qi) * C + c);
            store4<T>(o, make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv));
        }
    }
}

template <typename T, int D>
int launch_t(const ccdm_op &op, cudaStream_t s) {
    const int Ttok = op.Hin * op.Win;
    dim3 grid((Ttok + QT - 1) / QT, op.heads, op.B);
    // unet.py:354  scale = 1 / sqrt(sqrt(ch)), applied to q and to k
    const float scale = float(1.0 / sqrt(sqrt(double(D))));
    CCDM_CUDA(launch_pdl(attention_ffma_kernel<T, D>, grid, dim3(QT), 0, s, (const T *)op.src0, (T *)op.out, Ttok, op.heads, scale));
    CCDM_LAUNCH_CHECK("attention_ffma_kernel");
    return 0;
}

}  // namespace

int launch_attention_ffma(const ccdm_op &op, cudaStream_t s) {
    if (op.heads <= 0 || op.C0 != op.heads * op.head_dim * 3) CCDM_FAIL(-2, "attention: qkv channels %d != 3*heads*head_dim", op.C0);
    const bool bf = op.dtype == CCDM_DT_BF16;
    switch (op.head_dim) {
        case 32: return bf ? launch_t<__nv_bfloat16, 32>(op, s) : launch_t<float, 32>(op, s);
        case 64: return bf ? launch_t<__nv_bfloat16, 64>(op, s) : launch_t<float, 64>(op, s);
        default: CCDM_FAIL(-3, "attention: head_dim %d not supported (32 or 64)", op.head_dim);
    }
}

bool attention_tc_supported(const ccdm_op &op);
int launch_attention_tc(const ccdm_op &op, cudaStream_t s);

int launch_attention(const ccdm_op &op, cudaStream_t s) {
    if (attention_tc_supported(op)) return launch_attention_tc(op, s);
    return launch_attention_ffma(op, s);
}

}  // namespace ccdm
