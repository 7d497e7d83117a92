// DINO ViT condition encoder: the kernels under ccdm_b200.models.condition_encoder.DinoViT.
//
// Reference: ddpm/models/condition_encoder.py:26-46 (DinoViT.forward = ViTExtractor.extract_descriptors),
// ddpm/models/dino.py:211-229 (_extract_features: one forward of the ViT with a hook on blocks[layer].attn),
// :172-176 (the 'key' facet: qkv(norm1(x)) reshaped [B, N, 3, heads, d], component 1), :279-309 (drop the cls token,
// channel = d * heads + h, [B, C, h_p, w_p], bilinear resize), :86-117 (position-embedding interpolation).  The ViT itself
// (facebookresearch/dino main, vision_transformer.py: PatchEmbed -> cls token + position embedding -> pre-LN blocks
// x + proj(MHA(LN(x))), x + fc2(GELU(fc1(LN(x)))) ) is a torch.hub dependency of the reference, restated in oracle/dino_ref.py.
//
// Everything runs in the fp16x2 ("exact") storage format of the sampler: tokens are a plane-major tensor
// [B][C/8][2][T][8] fp16 (hi and lo plane of an 8-channel group adjacent, 16 x value), token 0 = cls -- the layout
// attention_tc_kernel<64, true, 64> reads and writes, so the attention of the encoder IS the sampler's attention kernel
// (CCDM_OP_ATTENTION with Hin = 1, Win = T, head_dim 64; the qkv rows are permuted to its per-head q|k|v order when
// the weights are packed).  What this file adds:
//   vit_linear_kernel     y = x W^T + b (+GELU) (+residual): tcgen05 GEMM on split fp16 operands -- per K step
//                         A_hi x [W_hi; W_lo] (one MMA, N = 256) and A_lo x W_hi into the upper 128 columns, fp32 in
//                         TMEM, the two halves added by the epilogue: fp32-grade products at tensor-core rate.
//   vit_layernorm_kernel  LayerNorm over the channels of a token (two-pass, values held in registers)
//   vit_patch_embed_kernel  PatchEmbed conv (kernel p, stride s) + bias + position embedding, cls token row
//   vit_pos_embed_kernel  bicubic resize of the position-embedding grid (F.interpolate, A = -0.75, given scale factors)
//   vit_descriptor_kernel key facet -> [B, C, Ho, Wo] fp32 NCHW, channel d * heads + h, bilinear (align_corners = False)
#include <math.h>

#include "tc_common.cuh"

namespace ccdm {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// linear layer
// ------------------------------------------------------------------------------------------------------------------
constexpr int VL_M = 128;    // tokens per CTA = threads = MMA M
constexpr int VL_NT = 128;   // output channels per CTA
constexpr int VL_KG = 4;     // 8-channel groups per K chunk (32 channels: two K = 16 steps)
constexpr int VL_NS = 3;     // pipeline stages
constexpr uint32_t VL_A_STAGE = VL_KG * 2 * VL_M * 16;   // [group][hi|lo][128 tokens][16 B] = 16 KB
constexpr uint32_t VL_W_STAGE = VL_KG * 2 * VL_NT * 16;  // [group][hi|lo][128 channels][16 B] = 16 KB
constexpr size_t VL_SMEM = VL_NS * (VL_A_STAGE + VL_W_STAGE) + (2 * VL_NS + 1) * 8 + 16;
constexpr uint32_t VL_TMEM_COLS = 2 * VL_NT;

struct VlP {
    const __half *x;    // [B][Cin/8][2][T][8]
    const __half *w;    // [Cout/NT][Cin/8][2][NT][8]  (2^shift * W, hi rows then lo rows per group)
    const float *bias;  // [Cout]
    const __half *res;  // [B][Cout/8][2][T][8] or null
    __half *out;        // [B][Cout/8][2][T][8]
    int T, Cin, Cout, tiles, gelu;
    float descale;      // 2^-(shift + 4)
    uint32_t idesc_n256, idesc_n128;
};

__device__ __forceinline__ void vl_tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 stored channels (hi + lo rows of one group) -> values
__device__ __forceinline__ void load8_x3(const __half *hi_row, size_t lo_off, float *v) {
    const uint4 h = *reinterpret_cast<const uint4 *>(hi_row);
    const uint4 l = *reinterpret_cast<const uint4 *>(hi_row + lo_off);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
    constexpr float inv = 1.0f / float(1 << CCDM_F16X2_SCALE_LOG2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = unpack_f16x2(hw[i]), b = unpack_f16x2(lw[i]);
        v[2 * i] = (a.x + b.x) * inv;
        v[2 * i + 1] = (a.y + b.y) * inv;
    }
}
__device__ __forceinline__ void store8_x3(__half *hi_row, size_t lo_off, const float *v) {
    uint32_t pk[4], pl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_f16x2(v[2 * i], v[2 * i + 1], pk[i], pl[i]);
    *reinterpret_cast<uint4 *>(hi_row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4 *>(hi_row + lo_off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// One CTA = 128 tokens of one image x 128 output channels.  Thread 0 streams K chunks of 32 channels (activation rows:
// one bulk copy per (group, hi|lo) plane -- a tile of a plane is contiguous; weights: one bulk copy per chunk) through a
// three-stage mbarrier ring and issues the MMAs; afterwards every thread owns one token row of the accumulator.
__global__ void __launch_bounds__(VL_M) vit_linear_kernel(const VlP p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;
    uint8_t *sW = smem + VL_NS * VL_A_STAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + VL_NS * VL_W_STAGE);
    uint64_t *full = bars, *empty = bars + VL_NS, *acc_done = bars + 2 * VL_NS;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * VL_NS + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int tile = blockIdx.x % p.tiles, b = blockIdx.x / p.tiles, cc = blockIdx.y;
    const int T = p.T, tok0 = tile * VL_M;
    const int nq = min(VL_M, T - tok0);
    const int G = p.Cin / 8, n_chunks = G / VL_KG;

    if (warp == 0) tmem_alloc(s_tmem, VL_TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < VL_NS; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, 1);
        }
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    pdl_launch_dependents();
    pdl_wait();

    if (tid == 0) {
        auto load = [&](int kc, int s) {
            mbar_expect_tx(full + s, uint32_t(VL_KG * 2 * nq * 16) + VL_W_STAGE);
#pragma unroll
            for (int g = 0; g < VL_KG * 2; ++g)
                bulk_g2s(sA + s * VL_A_STAGE + g * VL_M * 16, p.x + ((size_t(b) * G * 2 + size_t(kc) * VL_KG * 2 + g) * T + tok0) * 8,
                         uint32_t(nq * 16), full + s);
            bulk_g2s(sW + s * VL_W_STAGE, p.w + (size_t(cc) * G + size_t(kc) * VL_KG) * 2 * VL_NT * 8, VL_W_STAGE, full + s);
        };
        const uint32_t desc_hi = 8u | (1u << 14);  // SBO = 128 B between 8-row groups, descriptor version 1
        for (int s = 0; s < VL_NS && s < n_chunks; ++s) load(s, s);
        for (int kc = 0; kc < n_chunks; ++kc) {
            const int s = kc % VL_NS;
            mbar_wait(full + s, uint32_t(kc / VL_NS) & 1u);
            tc_fence_after();
            // LBO = stride between 8-channel groups = 2 planes of 128 rows (in 16-byte units)
            const uint32_t a_lo = (smem_u32(sA + s * VL_A_STAGE) >> 4) | (uint32_t(2 * VL_M) << 16);
            const uint32_t w_lo = (smem_u32(sW + s * VL_W_STAGE) >> 4) | (uint32_t(2 * VL_NT) << 16);
#pragma unroll
            for (int j = 0; j < VL_KG / 2; ++j) {
                const uint32_t a = a_lo + uint32_t(j * 2 * 2 * VL_M), w = w_lo + uint32_t(j * 2 * 2 * VL_NT);
                // columns [0, 128): hi*hi; [128, 256): hi*lo (this MMA) + lo*hi (the next one)
                umma_bf16(tmem, (uint64_t(desc_hi) << 32) | a, (uint64_t(desc_hi) << 32) | w, p.idesc_n256, (kc > 0 || j > 0) ? 1u : 0u);
                umma_bf16(tmem + VL_NT, (uint64_t(desc_hi) << 32) | (a + VL_M), (uint64_t(desc_hi) << 32) | w, p.idesc_n128, 1u);
            }
            umma_commit(empty + s);
            // refill the stage of the PREVIOUS chunk: its MMAs complete while this chunk's are queued behind them
            if (kc >= 1 && kc - 1 + VL_NS < n_chunks) {
                const int sp = (kc - 1) % VL_NS;
                mbar_wait(empty + sp, uint32_t((kc - 1) / VL_NS) & 1u);
                load(kc - 1 + VL_NS, sp);
            }
        }
        umma_commit(acc_done);
    }
    __syncwarp();
    mbar_wait<256>(acc_done, 0u);
    tc_fence_after();

    const uint32_t trow = uint32_t(warp * 32) << 16;  // this warp's TMEM lane quarter
    const bool live = tid < nq;
    const int tok = tok0 + tid;
    const size_t lo_off = size_t(T) * 8;
    const int Go = p.Cout / 8;
#pragma unroll 1
    for (int c0 = 0; c0 < VL_NT; c0 += 32) {
        float hh[32], cr[32];
        vl_tmem_ld32(tmem + trow + uint32_t(c0), hh);
        vl_tmem_ld32(tmem + trow + uint32_t(VL_NT + c0), cr);
        if (live) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int ch = cc * VL_NT + c0 + 8 * g;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float y = fmaf(hh[8 * g + i] + cr[8 * g + i], p.descale, __ldg(p.bias + ch + i));
                    if (p.gelu) y = 0.5f * y * (1.0f + erff(y * 0.70710678118654752440f));  // nn.GELU (exact erf form)
                    v[i] = y;
                }
                const size_t row = ((size_t(b) * Go + ch / 8) * 2 * T + tok) * 8;
                if (p.res) {
                    float r[8];
                    load8_x3(p.res + row, lo_off, r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += r[i];
                }
                store8_x3(p.out + row, lo_off, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, VL_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm over the channels of every token
// ------------------------------------------------------------------------------------------------------------------
// block (32 tokens, 8 channel slices): a warp = 32 consecutive tokens of one (group, hi|lo) plane -> 512-byte runs; thread
// (tx, ty) keeps groups ty, ty + 8, ... of its token in registers, so the tensor is read once.
template <int GPT>
__global__ void __launch_bounds__(256) vit_layernorm_kernel(const __half *x, const float *gamma, const float *beta, __half *out, int T, int G,
                                                            float eps) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tok = blockIdx.x * 32 + tx, b = blockIdx.y;
    const bool live = tok < T;
    const size_t lo_off = size_t(T) * 8;
    float v[GPT][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < GPT; ++k) {
        const int g = ty + 8 * k;
        if (live && g < G) {
            load8_x3(x + ((size_t(b) * G + g) * 2 * T + tok) * 8, lo_off, v[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[k][i];
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[k][i] = 0.f;
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += red[j][tx];
    const float mean = tot / float(G * 8);
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < GPT; ++k)
        if (ty + 8 * k < G) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = v[k][i] - mean;
                q = fmaf(d, d, q);
            }
        }
    red[ty][tx] = q;
    __syncthreads();
    float qt = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) qt += red[j][tx];
    const float rstd = 1.0f / sqrtf(qt / float(G * 8) + eps);
    if (!live) return;
#pragma unroll
    for (int k = 0; k < GPT; ++k) {
        const int g = ty + 8 * k;
        if (g < G) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf((v[k][i] - mean) * rstd, __ldg(gamma + 8 * g + i), __ldg(beta + 8 * g + i));
            store8_x3(out + ((size_t(b) * G + g) * 2 * T + tok) * 8, lo_off, y);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// patch embedding
// ------------------------------------------------------------------------------------------------------------------
// One block = 16 horizontally adjacent patches of one image row of patches; thread d (and d + 128, ...) owns output channels.
// wt: the conv weight transposed to [K = 3 p p][D]; pos: [1 + hp wp][D] (row 0 = cls), already resized for this image size.
template <int DM>
__global__ void __launch_bounds__(128) vit_patch_embed_kernel(const float *img, const float *wt, const float *bias, const float *cls,
                                                              const float *pos, __half *out, int H, int W, int p, int s, int hp, int wp) {
    extern __shared__ __align__(16) float sm[];
    constexpr int D = DM * 128;
    const int K = 3 * p * p;
    float *patch = sm;            // [K][16]
    float *stage = sm + K * 16;   // [16][D + 4]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * 16, y = blockIdx.y, b = blockIdx.z;
    const int T = 1 + hp * wp, G = D / 8;
    for (int idx = tid; idx < 16 * K; idx += 128) {
        const int k = idx / 16, t = idx % 16;
        const int c = k / (p * p), i = (k / p) % p, j = k % p;
        float v = 0.f;
        if (x0 + t < wp) v = __ldg(img + ((size_t(b) * 3 + c) * H + (y * s + i)) * W + (x0 + t) * s + j);
        patch[k * 16 + t] = v;
    }
    __syncthreads();
    float acc[DM][16];
#pragma unroll
    for (int m = 0; m < DM; ++m)
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[m][t] = 0.f;
    for (int k = 0; k < K; ++k) {
        float pv[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const float4 f = *reinterpret_cast<const float4 *>(patch + k * 16 + 4 * q4);
            pv[4 * q4] = f.x; pv[4 * q4 + 1] = f.y; pv[4 * q4 + 2] = f.z; pv[4 * q4 + 3] = f.w;
        }
#pragma unroll
        for (int m = 0; m < DM; ++m) {
            const float wv = __ldg(wt + size_t(k) * D + m * 128 + tid);
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[m][t] = fmaf(wv, pv[t], acc[m][t]);
        }
    }
#pragma unroll
    for (int m = 0; m < DM; ++m) {
        const int d = m * 128 + tid;
        const float bv = __ldg(bias + d);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            float v = acc[m][t] + bv;
            if (x0 + t < wp) v += __ldg(pos + size_t(1 + y * wp + x0 + t) * D + d);
            stage[t * (D + 4) + d] = v;
        }
    }
    __syncthreads();
    const size_t lo_off = size_t(T) * 8;
    for (int it = tid; it < 16 * G; it += 128) {
        const int t = it % 16, g = it / 16;
        if (x0 + t >= wp) continue;
        const int tok = 1 + y * wp + x0 + t;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = stage[t * (D + 4) + 8 * g + i];
        store8_x3(out + ((size_t(b) * G + g) * 2 * T + tok) * 8, lo_off, v);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid < G) {  // token 0: cls_token + pos_embed[0]
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(cls + 8 * tid + i) + __ldg(pos + 8 * tid + i);
        store8_x3(out + ((size_t(b) * G + tid) * 2 * T) * 8, lo_off, v);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// position-embedding resize (bicubic, A = -0.75, align_corners = False, source index from the GIVEN scale factor)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float *c) {
    const float A = -0.75f;
    const float x1 = t, x2 = 1.0f - t;
    c[0] = ((A * (x1 + 1.0f) - 5.0f * A) * (x1 + 1.0f) + 8.0f * A) * (x1 + 1.0f) - 4.0f * A;
    c[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
    c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
    c[3] = ((A * (x2 + 1.0f) - 5.0f * A) * (x2 + 1.0f) + 8.0f * A) * (x2 + 1.0f) - 4.0f * A;
}
// pos: [1 + n*n][D] (row 0 = cls) -> out: [1 + hp*wp][D]; ratio_* = float(1 / scale_factor) as ATen computes it
__global__ void vit_pos_embed_kernel(const float *pos, int n, int D, int hp, int wp, float ratio_h, float ratio_w, float *out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int tok = blockIdx.y;  // 0 .. hp*wp
    if (d >= D) return;
    if (tok == 0) {
        out[d] = pos[d];
        return;
    }
    const int oy = (tok - 1) / wp, ox = (tok - 1) % wp;
    const float sy = ratio_h * (float(oy) + 0.5f) - 0.5f, sx = ratio_w * (float(ox) + 0.5f) - 0.5f;
    const float fy = floorf(sy), fx = floorf(sx);
    const int iy = int(fy), ix = int(fx);
    float cy[4], cx[4];
    cubic_coeffs(sy - fy, cy);
    cubic_coeffs(sx - fx, cx);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int yy = min(max(iy - 1 + i, 0), n - 1);
        float row = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int xx = min(max(ix - 1 + j, 0), n - 1);
            row = fmaf(cx[j], pos[size_t(1 + yy * n + xx) * D + d], row);
        }
        acc = fmaf(cy[i], row, acc);
    }
    out[size_t(tok) * D + d] = acc;
}

// ------------------------------------------------------------------------------------------------------------------
// key facet -> descriptor map
// ------------------------------------------------------------------------------------------------------------------
// k: [B][C/8][2][T][8] with channel = h * hd + d (token 0 = cls, dropped); out: [B][C][Ho][Wo] fp32 with channel d * heads + h
// (dino.py:295 x.permute(0, 2, 3, 1).flatten(-2, -1)), bilinear from the (hp, wp) patch grid (align_corners = False; the
// identity when the sizes agree: lambda = 0 selects one source token exactly).
__global__ void vit_descriptor_kernel(const __half *k, float *out, int T, int C, int heads, int hd, int hp, int wp, int Ho, int Wo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (idx >= Ho * Wo) return;
    const int oy = idx / Wo, ox = idx % Wo;
    const float rh = float(hp) / float(Ho), rw = float(wp) / float(Wo);
    const float sy = fmaxf(rh * (float(oy) + 0.5f) - 0.5f, 0.f), sx = fmaxf(rw * (float(ox) + 0.5f) - 0.5f, 0.f);
    const int y0 = min(int(sy), hp - 1), x0 = min(int(sx), wp - 1);
    const int y1 = y0 + (y0 < hp - 1 ? 1 : 0), x1 = x0 + (x0 < wp - 1 ? 1 : 0);
    const float ly = sy - float(y0), lx = sx - float(x0), hy = 1.0f - ly, hx = 1.0f - lx;
    const int t00 = 1 + y0 * wp + x0, t01 = 1 + y0 * wp + x1, t10 = 1 + y1 * wp + x0, t11 = 1 + y1 * wp + x1;
    const int G = C / 8;
    const size_t lo_off = size_t(T) * 8;
    for (int g = 0; g < G; ++g) {
        const __half *base = k + (size_t(b) * G + g) * 2 * T * 8;
        float a[8], bq[8], c[8], d[8];
        load8_x3(base + size_t(t00) * 8, lo_off, a);
        load8_x3(base + size_t(t01) * 8, lo_off, bq);
        load8_x3(base + size_t(t10) * 8, lo_off, c);
        load8_x3(base + size_t(t11) * 8, lo_off, d);
        const int h = (8 * g) / hd, d0 = (8 * g) % hd;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = hy * (hx * a[i] + lx * bq[i]) + ly * (hx * c[i] + lx * d[i]);
            out[((size_t(b) * C + size_t(d0 + i) * heads + h) * Ho + oy) * Wo + ox] = v;
        }
    }
}

}  // namespace
}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_vit_linear_nt(void) { return VL_NT; }

extern "C" int ccdm_vit_linear(const void *x, const void *w_packed, const float *bias, const void *residual, int B, int T, int Cin, int Cout,
                               int gelu, int acc_shift, void *out, void *stream) {
    if (B <= 0 || T <= 0) return 0;
    if (!x || !w_packed || !bias || !out) CCDM_FAIL(-2, "vit_linear: missing tensors");
    if (Cin <= 0 || (Cin % (8 * VL_KG)) || Cout <= 0 || (Cout % VL_NT)) CCDM_FAIL(-2, "vit_linear: Cin %d must be a multiple of 32, Cout %d of %d", Cin, Cout, VL_NT);
    if (acc_shift < CCDM_F16X2_SCALE_LOG2 || acc_shift > 40) CCDM_FAIL(-2, "vit_linear: acc_shift %d", acc_shift);
    if (out == x || out == residual) CCDM_FAIL(-2, "vit_linear: the output must not alias an input (other CTAs still read it)");
    VlP p{};
    p.x = (const __half *)x; p.w = (const __half *)w_packed; p.bias = bias; p.res = (const __half *)residual; p.out = (__half *)out;
    p.T = T; p.Cin = Cin; p.Cout = Cout; p.tiles = (T + VL_M - 1) / VL_M; p.gelu = gelu ? 1 : 0;
    p.descale = float(ldexp(1.0, -acc_shift));
    // cute::UMMA::InstrDescriptor: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t base = (1u << 4) | (uint32_t(VL_M >> 4) << 24);
    p.idesc_n256 = base | (uint32_t((2 * VL_NT) >> 3) << 17);
    p.idesc_n128 = base | (uint32_t(VL_NT >> 3) << 17);
    static bool attr_done = false;
    if (!attr_done) {
        CCDM_CUDA(cudaFuncSetAttribute(vit_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(VL_SMEM)));
        attr_done = true;
    }
    const long long gx = (long long)p.tiles * B;
    if (gx > 0x7fffffffll || Cout / VL_NT > 65535) CCDM_FAIL(-2, "vit_linear: grid too large");
    CCDM_CUDA(launch_pdl(vit_linear_kernel, dim3(unsigned(gx), unsigned(Cout / VL_NT)), dim3(VL_M), VL_SMEM, (cudaStream_t)stream, p));
    CCDM_LAUNCH_CHECK("vit_linear_kernel");
    return 0;
}

extern "C" int ccdm_vit_layernorm(const void *x, const float *gamma, const float *beta, int B, int T, int C, float eps, void *out,
                                  void *stream) {
    if (B <= 0 || T <= 0) return 0;
    if (!x || !gamma || !beta || !out) CCDM_FAIL(-2, "vit_layernorm: missing tensors");
    if (C <= 0 || (C % 8) || C > 8 * 8 * 16) CCDM_FAIL(-2, "vit_layernorm: C = %d (multiple of 8, at most 1024)", C);
    if (B > 65535) CCDM_FAIL(-2, "vit_layernorm: batch too large for one launch");
    const int G = C / 8;
    dim3 grid(unsigned((T + 31) / 32), unsigned(B)), block(32, 8);
    cudaStream_t s = (cudaStream_t)stream;
    if (G <= 48) vit_layernorm_kernel<6><<<grid, block, 0, s>>>((const __half *)x, gamma, beta, (__half *)out, T, G, eps);
    else if (G <= 96) vit_layernorm_kernel<12><<<grid, block, 0, s>>>((const __half *)x, gamma, beta, (__half *)out, T, G, eps);
    else vit_layernorm_kernel<16><<<grid, block, 0, s>>>((const __half *)x, gamma, beta, (__half *)out, T, G, eps);
    CCDM_LAUNCH_CHECK("vit_layernorm_kernel");
    return 0;
}

extern "C" int ccdm_vit_patch_embed(const float *image, const float *w_t, const float *bias, const float *cls, const float *pos, int B,
                                    int H, int W, int patch, int stride, int D, void *tokens, void *stream) {
    if (B <= 0) return 0;
    if (!image || !w_t || !bias || !cls || !pos || !tokens) CCDM_FAIL(-2, "vit_patch_embed: missing tensors");
    if (patch < 1 || patch > 16 || stride < 1 || H < patch || W < patch) CCDM_FAIL(-2, "vit_patch_embed: patch %d stride %d image %dx%d", patch, stride, H, W);
    if (D != 384 && D != 768) CCDM_FAIL(-2, "vit_patch_embed: embed_dim %d (384 = ViT-S, 768 = ViT-B)", D);
    const int hp = 1 + (H - patch) / stride, wp = 1 + (W - patch) / stride;
    if (hp > 65535 || B > 65535) CCDM_FAIL(-2, "vit_patch_embed: grid too large");
    const int K = 3 * patch * patch;
    const size_t smem = sizeof(float) * (size_t(K) * 16 + 16 * size_t(D + 4));
    dim3 grid(unsigned((wp + 15) / 16), unsigned(hp), unsigned(B));
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 384) {
        CCDM_CUDA(cudaFuncSetAttribute(vit_patch_embed_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        vit_patch_embed_kernel<3><<<grid, 128, smem, s>>>(image, w_t, bias, cls, pos, (__half *)tokens, H, W, patch, stride, hp, wp);
    } else {
        CCDM_CUDA(cudaFuncSetAttribute(vit_patch_embed_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        vit_patch_embed_kernel<6><<<grid, 128, smem, s>>>(image, w_t, bias, cls, pos, (__half *)tokens, H, W, patch, stride, hp, wp);
    }
    CCDM_LAUNCH_CHECK("vit_patch_embed_kernel");
    return 0;
}

extern "C" int ccdm_vit_pos_embed(const float *pos, int n_side, int D, int hp, int wp, double scale_h, double scale_w, float *out,
                                  void *stream) {
    if (!pos || !out || n_side <= 0 || D <= 0 || hp <= 0 || wp <= 0 || !(scale_h > 0) || !(scale_w > 0))
        CCDM_FAIL(-2, "vit_pos_embed: bad arguments");
    if (1 + (long long)hp * wp > 65535) CCDM_FAIL(-2, "vit_pos_embed: more than 65534 patch tokens");
    // ATen area_pixel_compute_scale with an explicit scale factor: static_cast<float>(1.0 / scale)
    const float rh = float(1.0 / scale_h), rw = float(1.0 / scale_w);
    dim3 grid(unsigned((D + 127) / 128), unsigned(1 + hp * wp));
    vit_pos_embed_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(pos, n_side, D, hp, wp, rh, rw, out);
    CCDM_LAUNCH_CHECK("vit_pos_embed_kernel");
    return 0;
}

extern "C" int ccdm_vit_descriptor(const void *key, int B, int T, int heads, int head_dim, int hp, int wp, int Ho, int Wo, float *out,
                                   void *stream) {
    if (B <= 0) return 0;
    if (!key || !out) CCDM_FAIL(-2, "vit_descriptor: missing tensors");
    if (heads <= 0 || head_dim <= 0 || (head_dim % 8) || hp <= 0 || wp <= 0 || Ho <= 0 || Wo <= 0 || T != 1 + hp * wp || B > 65535)
        CCDM_FAIL(-2, "vit_descriptor: heads %d head_dim %d grid %dx%d T %d", heads, head_dim, hp, wp, T);
    dim3 grid(unsigned((Ho * Wo + 127) / 128), unsigned(B));
    vit_descriptor_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __half *)key, out, T, heads * head_dim, heads, head_dim, hp, wp, Ho, Wo);
    CCDM_LAUNCH_CHECK("vit_descriptor_kernel");
    return 0;
}
