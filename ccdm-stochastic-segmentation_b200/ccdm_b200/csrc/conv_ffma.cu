// Fused GroupNorm + SiLU + conv (3x3 / 1x1, stride 1|2, nearest-x2 in front) with
// bias + timestep-embedding + residual / fused 1x1 skip in the epilogue and the
// GroupNorm statistics of the OUTPUT emitted for the consumer.  fp32 FFMA maths:
// this is the exact ("parity") implementation of every conv-shaped layer on the
// hot path; the tensor-core kernels are checked against it on the GPU.
//
// Replaces (reference, /root/reference/ddpm/models/unet_openai/unet.py):
//   ResBlock._forward :242-262 (two launches: in_layers+emb, out_layers+skip),
//   Downsample.forward :144-146, Upsample.forward :106-116,
//   AttentionBlock qkv / proj_out 1x1 convs :308-311 (with :291 norm, no SiLU),
//   input concat + conv :760,516-518 (src_kind 1: one-hot(labels) ++ image),
//   the output conv :701-705 (Cout = K, fp32 logits).
//
// Layout: activations NHWC; weights [tap][CinP][CoutP]; one CTA = 8x16 output
// pixels x 32 output channels, 128 threads, thread tile 8 pixels (a row segment)
// x 4 channels.  Input channels are streamed in chunks of 8 through shared memory
// in channel-major planes (so a thread's 8 neighbouring pixels are contiguous);
// GroupNorm scale/shift and SiLU are applied while staging, and out-of-image halo
// elements are written as exact zeros AFTER the activation (zero padding happens
// after GN+SiLU in the reference: conv2d(padding=1) sees SiLU(GN(x)) padded with 0).
#include <stdlib.h>

#include "common.cuh"

namespace ccdm {

namespace {

constexpr int TH = 8, TW = 16;      // output tile
constexpr int COT = 32;             // output channels per CTA
constexpr int CK = 8;               // input channels per smem chunk
constexpr int NTHREADS = 128;

struct ConvP {
    const void *src0, *src1;
    const double *stat0, *stat1;
    const float *gamma, *beta;
    const float *weight, *bias, *emb;
    const void *skip0, *skip1;
    const float *skip_w;
    const void *res;
    void *out;
    double *ostat;
    float *part;
    unsigned int *ticket;
    const uint8_t *labels;
    const float *image;
    const ccdm_step_entry *steps;
    const int *step_ptr;
    int B, Hin, Win, Hout, Wout, C0, C1, Cin, CinP, Cout, CoutP;
    int upsample, gn, silu, S0, S1, K, C_img, emb_off, emb_cols, emb_bstride, src_kind, out_f32;
    int tiles_x, tiles_y;
    int img_rep;  // samples per conditioning image (>= 1)
    int gn_cpg, gn_off;  // GroupNorm group size / channel offset of channel 0 inside the normalised concatenation (ccdm_op::gn_cpg, gn_off)
};

// Element offset of channels c..c+3 of pixel `pix` (index inside the sample) of sample b in an activation
// tensor with C channels and HW pixels per sample: fp32 tensors are NHWC, bf16 tensors are plane-major
// [B][C/8][H][W][8] (the tensor-core kernels' layout: one 16-byte row of a UMMA core matrix per pixel).
template <typename T>
__device__ __forceinline__ size_t act_off(int b, size_t HW, int C, size_t pix, int c);
template <>
__device__ __forceinline__ size_t act_off<float>(int b, size_t HW, int C, size_t pix, int c) {
    return (size_t(b) * HW + pix) * C + c;
}
template <>
__device__ __forceinline__ size_t act_off<__nv_bfloat16>(int b, size_t HW, int C, size_t pix, int c) {
    return ((size_t(b) * (C >> 3) + (c >> 3)) * HW + pix) * 8 + (c & 7);
}

template <int KS, int STRIDE>
struct Geo {
    static constexpr int IH = (TH - 1) * STRIDE + KS;
    static constexpr int IW = (TW - 1) * STRIDE + KS;
    static constexpr int PLANE = (IH * IW) | 1;  // odd plane stride: conflict-free transposed stores
    static constexpr int SEG = 7 * STRIDE + KS;  // inputs one thread needs per (channel, tap row)
};

// GroupNorm scale/shift of the concatenated input, folded from per-channel sums.
__device__ void build_gn_affine(const ConvP &p, int b, float *sA, float *sB) {
    const int Cin = p.Cin;
    const int cpg = p.gn_cpg > 0 ? p.gn_cpg : Cin / kGnGroups;
    const int hw_in = p.Hin * p.Win;
    const double n = double(cpg) * double(hw_in);
    for (int c = threadIdx.x; c < Cin; c += NTHREADS) {
        // (a group cut by this op's channel range is incomplete: its channels carry zero weights, ccdm_op::gn_off)
        const int g0 = ((c + p.gn_off) / cpg) * cpg - p.gn_off;
        const int j0 = g0 < 0 ? 0 : g0, j1 = g0 + cpg < Cin ? g0 + cpg : Cin;
        double s = 0.0, q = 0.0;
        for (int cc = j0; cc < j1; ++cc) {
            const double *st = cc < p.C0 ? p.stat0 + (size_t(b) * p.C0 + cc) * 2
                                         : p.stat1 + (size_t(b) * p.C1 + (cc - p.C0)) * 2;
            s += st[0];
            q += st[1];
        }
        double mean = s / n;
        double var = q / n - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        float rstd = float(1.0 / sqrt(var + double(kGnEps)));
        float a = p.gamma[c] * rstd;
        sA[c] = a;
        sB[c] = p.beta[c] - float(mean) * a;
    }
}

template <typename T, int KS, int STRIDE>
__global__ void __launch_bounds__(NTHREADS) conv_ffma_kernel(const ConvP p) {
    using G = Geo<KS, STRIDE>;
    extern __shared__ float smem[];
    float *sIn = smem;                             // [CK][PLANE]
    float *sW = sIn + CK * G::PLANE;               // [KS*KS][CK][COT]
    float *sRed = sW + KS * KS * CK * COT;         // [16][COT][2]
    float *sA = sRed + 16 * COT * 2;               // [Cin]
    float *sB = sA + p.Cin;                        // [Cin]
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const int tx = tid & 7;    // channel quad
    const int ty = tid >> 3;   // pixel segment: row ty>>1, columns (ty&1)*8 .. +8
    const int prow = ty >> 1, pcol0 = (ty & 1) * 8;
    const int tile = blockIdx.x;
    const int oy0 = (tile / p.tiles_x) * TH, ox0 = (tile % p.tiles_x) * TW;
    const int co0 = blockIdx.y * COT;
    const int b = blockIdx.z;
    constexpr int PAD = KS / 2;

    pdl_launch_dependents();
    pdl_wait();
    if (p.gn) build_gn_affine(p, b, sA, sB);
    __syncthreads();

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;

    // conv-input space (after the optional nearest x2): Hc x Wc
    const int Hc = p.upsample ? p.Hin * 2 : p.Hin;
    const int Wc = p.upsample ? p.Win * 2 : p.Win;
    const int iy0 = oy0 * STRIDE - PAD, ix0 = ox0 * STRIDE - PAD;

    for (int c0 = 0; c0 < p.CinP; c0 += CK) {
        // ---- stage CK input channels of the halo tile, GN+SiLU applied ----------
        if (p.src_kind == 0) {
            const bool first = c0 < p.C0;
            const T *src = reinterpret_cast<const T *>(first ? p.src0 : p.src1);
            const int Cs = first ? p.C0 : p.C1;
            const int cs0 = first ? c0 : c0 - p.C0;
            for (int e = tid; e < G::IH * G::IW * 2; e += NTHREADS) {
                int pix = e >> 1, q = e & 1;
                int py = pix / G::IW, px = pix - py * G::IW;
                int iy = iy0 + py, ix = ix0 + px;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (iy >= 0 && iy < Hc && ix >= 0 && ix < Wc) {
                    int sy = p.upsample ? (iy >> 1) : iy, sx = p.upsample ? (ix >> 1) : ix;
                    v = load4<T>(src + act_off<T>(b, size_t(p.Hin) * p.Win, Cs, size_t(sy) * p.Win + sx, cs0 + q * 4));
                    if (p.gn) {
                        int c = c0 + q * 4;
                        v.x = v.x * sA[c] + sB[c];
                        v.y = v.y * sA[c + 1] + sB[c + 1];
                        v.z = v.z * sA[c + 2] + sB[c + 2];
                        v.w = v.w * sA[c + 3] + sB[c + 3];
                    }
                    if (p.silu) {
                        v.x = silu_exact(v.x);
                        v.y = silu_exact(v.y);
                        v.z = silu_exact(v.z);
                        v.w = silu_exact(v.w);
                    }
                }
                float *d = sIn + (q * 4) * G::PLANE + pix;
                d[0] = v.x;
                d[G::PLANE] = v.y;
                d[2 * G::PLANE] = v.z;
                d[3 * G::PLANE] = v.w;
            }
        } else {
            // one-hot(labels) ++ image, never materialised (unet.py:760)
            for (int e = tid; e < G::IH * G::IW * CK; e += NTHREADS) {
                int cc = e / (G::IH * G::IW), pix = e - cc * (G::IH * G::IW);
                int py = pix / G::IW, px = pix - py * G::IW;
                int iy = iy0 + py, ix = ix0 + px;
                int c = c0 + cc;
                float v = 0.f;
                if (iy >= 0 && iy < Hc && ix >= 0 && ix < Wc && c < p.K + p.C_img) {
                    if (c < p.K)
                        v = (p.labels[(size_t(b) * p.Hin + iy) * p.Win + ix] == c) ? 1.f : 0.f;
                    else
                        v = p.image[((size_t(b / p.img_rep) * p.C_img + (c - p.K)) * p.Hin + iy) * p.Win + ix];
                }
                sIn[cc * G::PLANE + pix] = v;
            }
        }
        // ---- stage the weight chunk [tap][CK][COT] --------------------------------
        for (int e = tid; e < KS * KS * CK * COT; e += NTHREADS) {
            int co = e % COT, r = e / COT;
            int cc = r % CK, tap = r / CK;
            sW[e] = p.weight[(size_t(tap) * p.CinP + c0 + cc) * p.CoutP + co0 + co];
        }
        __syncthreads();
        // ---- FFMA -------------------------------------------------------------------
#pragma unroll 2
        for (int cc = 0; cc < CK; ++cc) {
#pragma unroll
            for (int dy = 0; dy < KS; ++dy) {
                const float *row = sIn + cc * G::PLANE + (prow * STRIDE + dy) * G::IW + pcol0 * STRIDE;
                float v[G::SEG];
#pragma unroll
                for (int i = 0; i < G::SEG; ++i) v[i] = row[i];
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                    const float4 w = *reinterpret_cast<const float4 *>(sW + ((dy * KS + dx) * CK + cc) * COT + tx * 4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float x = v[j * STRIDE + dx];
                        acc[j][0] += x * w.x;
                        acc[j][1] += x * w.y;
                        acc[j][2] += x * w.z;
                        acc[j][3] += x * w.w;
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- fused 1x1 skip conv on the raw (un-normalised) block input (unet.py:262) ----
    if (p.S0 > 0) {
        const int Stot = p.S0 + p.S1;
        for (int c0 = 0; c0 < Stot; c0 += CK) {
            const bool first = c0 < p.S0;
            const T *src = reinterpret_cast<const T *>(first ? p.skip0 : p.skip1);
            const int Cs = first ? p.S0 : p.S1;
            const int cs0 = first ? c0 : c0 - p.S0;
            for (int e = tid; e < TH * TW * 2; e += NTHREADS) {
                int pix = e >> 1, q = e & 1;
                int py = pix / TW, px = pix - py * TW;
                int oy = oy0 + py, ox = ox0 + px;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (oy < p.Hout && ox < p.Wout)
                    v = load4<T>(src + act_off<T>(b, size_t(p.Hout) * p.Wout, Cs, size_t(oy) * p.Wout + ox, cs0 + q * 4));
                float *d = sIn + (q * 4) * G::PLANE + pix;
                d[0] = v.x;
                d[G::PLANE] = v.y;
                d[2 * G::PLANE] = v.z;
                d[3 * G::PLANE] = v.w;
            }
            for (int e = tid; e < CK * COT; e += NTHREADS) {
                int co = e % COT, cc = e / COT;
                sW[e] = p.skip_w[(size_t(c0 + cc)) * p.CoutP + co0 + co];
            }
            __syncthreads();
#pragma unroll 2
            for (int cc = 0; cc < CK; ++cc) {
                const float *row = sIn + cc * G::PLANE + prow * TW + pcol0;
                const float4 w = *reinterpret_cast<const float4 *>(sW + cc * COT + tx * 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float x = row[j];
                    acc[j][0] += x * w.x;
                    acc[j][1] += x * w.y;
                    acc[j][2] += x * w.z;
                    acc[j][3] += x * w.w;
                }
            }
            __syncthreads();
        }
    }

    // ---- epilogue: bias + embedding + residual, store, statistics ------------------------
    const int co = co0 + tx * 4;
    float4 add = *reinterpret_cast<const float4 *>(p.bias + co);
    if (p.emb != nullptr) {
        const ccdm_step_entry &se = p.steps[*p.step_ptr];
        const float *er = p.emb + (size_t(se.emb_row) + size_t(b) * p.emb_bstride) * p.emb_cols + p.emb_off + co;
        if (co < p.Cout) {  // emb blocks always have Cout % 32 == 0; guard anyway
            add.x += er[0];
            add.y += er[1];
            add.z += er[2];
            add.w += er[3];
        }
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const int oy = oy0 + prow;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ox = ox0 + pcol0 + j;
        if (oy < p.Hout && ox < p.Wout) {
            float4 v = make_float4(acc[j][0] + add.x, acc[j][1] + add.y, acc[j][2] + add.z, acc[j][3] + add.w);
            const size_t hw = size_t(p.Hout) * p.Wout, pin = size_t(oy) * p.Wout + ox;
            const size_t pix = size_t(b) * hw + pin;
            if (p.res != nullptr) {
                float4 r = load4<T>(reinterpret_cast<const T *>(p.res) + act_off<T>(b, hw, p.Cout, pin, co));
                v.x += r.x;
                v.y += r.y;
                v.z += r.z;
                v.w += r.w;
            }
            // (fp32 NHWC logits: a pixel's row of Cout floats is 16-byte aligned only when Cout % 4 == 0 -- K = 5, 19 take the tail path)
            if (co + 3 < p.Cout && !(p.out_f32 && (p.Cout & 3))) {
                v = p.out_f32 ? store4<float>(reinterpret_cast<float *>(p.out) + pix * p.Cout + co, v)
                              : store4<T>(reinterpret_cast<T *>(p.out) + act_off<T>(b, hw, p.Cout, pin, co), v);
            } else {  // ragged channel tail (Cout = K classes): fp32 logits only
                float vv[4] = {v.x, v.y, v.z, v.w};
                for (int i = 0; i < 4; ++i)
                    if (co + i < p.Cout) reinterpret_cast<float *>(p.out)[pix * p.Cout + co + i] = vv[i];
            }
            s1[0] += v.x; s2[0] += v.x * v.x;
            s1[1] += v.y; s2[1] += v.y * v.y;
            s1[2] += v.z; s2[2] += v.z * v.z;
            s1[3] += v.w; s2[3] += v.w * v.w;
        }
    }
    if (p.ostat == nullptr) return;

    // per-tile partials, summed in a fixed order (deterministic), then the last
    // CTA of the sample folds all tiles in double precision.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sRed[(ty * COT + tx * 4 + i) * 2 + 0] = s1[i];
        sRed[(ty * COT + tx * 4 + i) * 2 + 1] = s2[i];
    }
    __syncthreads();
    const int n_tiles = gridDim.x;
    if (tid < COT * 2) {
        int c = tid >> 1, w = tid & 1;
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) s += sRed[(r * COT + c) * 2 + w];
        p.part[((size_t(b) * n_tiles + tile) * p.CoutP + co0 + c) * 2 + w] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int total = gridDim.x * gridDim.y;
        unsigned int prev = atomicAdd(p.ticket + b, 1u);
        s_last = (prev == total - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = tid; e < p.Cout * 2; e += NTHREADS) {
        int c = e >> 1, w = e & 1;
        double s = 0.0;
        for (int t = 0; t < n_tiles; ++t)
            s += double(__ldcg(p.part + ((size_t(b) * n_tiles + t) * p.CoutP + c) * 2 + w));
        p.ostat[(size_t(b) * p.Cout + c) * 2 + w] = s;
    }
    if (tid == 0) p.ticket[b] = 0u;  // self-reset for the next launch
}

template <typename T, int KS, int STRIDE>
int launch_t(const ConvP &p, cudaStream_t s) {
    using G = Geo<KS, STRIDE>;
    size_t smem = sizeof(float) * (size_t(CK) * G::PLANE + size_t(KS) * KS * CK * COT + 16 * COT * 2 + 2 * size_t(p.Cin));
    auto kern = conv_ffma_kernel<T, KS, STRIDE>;
    // function-level attribute: set once per instantiation to a fixed ceiling (captured graphs replay nodes
    // after later launches; a per-launch value would leave the function with whatever came last)
    constexpr size_t kMaxSmem = 96 * 1024;
    if (smem > kMaxSmem) CCDM_FAIL(-3, "conv_ffma: %zu bytes of shared memory needed (Cin too large)", smem);
    static bool attr_done = false;
    if (!attr_done) {
        CCDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxSmem)));
        attr_done = true;
    }
    dim3 grid(p.tiles_x * p.tiles_y, p.CoutP / COT, p.B);
    CCDM_CUDA(launch_pdl(kern, grid, dim3(NTHREADS), smem, s, p));
    CCDM_LAUNCH_CHECK("conv_ffma_kernel");
    return 0;
}

}  // namespace

size_t conv_part_floats(int B, int Hout, int Wout, int Cout) {
    int tiles = ((Hout + TH - 1) / TH) * ((Wout + TW - 1) / TW);
    int CoutP = (Cout + COT - 1) / COT * COT;
    return size_t(B) * tiles * CoutP * 2;
}

bool conv_tma_supported(const ccdm_op &op);
size_t conv_tma_part_floats(const ccdm_op &op);
int launch_conv_tma(const ccdm_op &op, cudaStream_t s);

// The tensor-core modes (bf16 / fp16x2 storage, op.exact == 0) have exactly one conv kernel, conv_tma.cu.  An op it
// cannot take falls to the FFMA kernel below, which reads the same plane-major bf16 tensors with fp32 [tap][Cin][Cout]
// weights and FOLDED double2 statistics; the engine binds it accordingly and says so (engine.Program.bind).
bool conv_uses_tma(const ccdm_op &op) { return !op.exact && op.kind == CCDM_OP_CONV && conv_tma_supported(op); }
bool conv_uses_tc(const ccdm_op &op) { return conv_uses_tma(op); }

size_t op_part_floats(const ccdm_op &op) {
    if (op.kind != CCDM_OP_CONV && op.kind != CCDM_OP_INPUT_CONV) return 0;
    if (conv_uses_tma(op)) return conv_tma_part_floats(op);
    return conv_part_floats(op.B, op.Hout, op.Wout, op.Cout);
}

int launch_conv(const ccdm_op &op, cudaStream_t s) {
    if (conv_uses_tma(op)) return launch_conv_tma(op, s);
    if (op.dtype == CCDM_DT_F16X2) CCDM_FAIL(-3, "conv: this shape does not fit the tensor-core kernel and the fp16x2 (exact tensor-core) mode has no other");
    if (op.st_slots[0] > 0 || op.st_slots[1] > 0)
        CCDM_FAIL(-2, "conv_ffma: deferred-fold statistics rows (st_slots > 0) are a tensor-core consumer's layout; this kernel reads folded double2 [B,C]");
    ConvP p{};
    p.src0 = (const void *)op.src0; p.src1 = (const void *)op.src1;
    p.stat0 = (const double *)op.stat0; p.stat1 = (const double *)op.stat1;
    p.gamma = (const float *)op.gamma; p.beta = (const float *)op.beta;
    p.weight = (const float *)op.weight; p.bias = (const float *)op.bias; p.emb = (const float *)op.emb;
    p.skip0 = (const void *)op.skip0; p.skip1 = (const void *)op.skip1; p.skip_w = (const float *)op.skip_w;
    p.res = (const void *)op.res; p.out = (void *)op.out; p.ostat = (double *)op.ostat;
    p.part = (float *)op.part; p.ticket = (unsigned int *)op.ticket;
    p.labels = (const uint8_t *)op.labels_in; p.image = (const float *)op.image;
    p.steps = (const ccdm_step_entry *)op.steps; p.step_ptr = (const int *)op.step_ptr;
    p.B = op.B; p.Hin = op.Hin; p.Win = op.Win; p.Hout = op.Hout; p.Wout = op.Wout;
    p.C0 = op.C0; p.C1 = op.C1; p.Cout = op.Cout;
    p.src_kind = op.src_kind;
    p.Cin = op.src_kind == 1 ? op.K + op.C_img : op.C0 + op.C1;
    p.CinP = (p.Cin + CK - 1) / CK * CK;
    p.CoutP = (op.Cout + COT - 1) / COT * COT;
    p.upsample = op.upsample; p.gn = op.gn; p.silu = op.silu; p.S0 = op.S0; p.S1 = op.S1;
    p.K = op.K; p.C_img = op.C_img; p.emb_off = op.emb_off; p.emb_cols = op.emb_cols; p.emb_bstride = op.emb_bstride;
    p.out_f32 = (op.out_dtype == CCDM_DT_F32);
    p.img_rep = op.img_rep > 1 ? op.img_rep : 1;
    p.tiles_x = (op.Wout + TW - 1) / TW; p.tiles_y = (op.Hout + TH - 1) / TH;

    if (op.B <= 0 || op.Cout <= 0 || p.Cin <= 0) CCDM_FAIL(-2, "conv: empty shape");
    if (op.ksize != 1 && op.ksize != 3) CCDM_FAIL(-2, "conv: ksize %d unsupported", op.ksize);
    if (op.stride != 1 && op.stride != 2) CCDM_FAIL(-2, "conv: stride %d unsupported", op.stride);
    if (op.stride == 2 && (op.ksize != 3 || op.upsample)) CCDM_FAIL(-2, "conv: stride 2 needs ksize 3, no upsample");
    if (op.src_kind == 0 && ((op.C0 % 8) || (op.C1 % 8))) CCDM_FAIL(-2, "conv: source channels must be multiples of 8");
    if (op.src_kind == 1 && (op.gn || op.upsample || !op.labels_in || !op.image)) CCDM_FAIL(-2, "conv: bad one-hot input op");
    if (op.gn && op.gn_cpg <= 0 && (p.Cin % kGnGroups)) CCDM_FAIL(-2, "conv: GroupNorm needs Cin %% 32 == 0 (got %d)", p.Cin);
    if (op.gn_cpg < 0 || op.gn_off < 0) CCDM_FAIL(-2, "conv: gn_cpg / gn_off must not be negative");
    p.gn_cpg = op.gn_cpg; p.gn_off = op.gn_off;
    if (op.gn && (!op.stat0 || (op.C1 && !op.stat1) || !op.gamma || !op.beta)) CCDM_FAIL(-2, "conv: gn without stats/affine");
    if ((op.S0 % 8) || (op.S1 % 8)) CCDM_FAIL(-2, "conv: skip channels must be multiples of 8");
    if (op.S0 > 0 && (op.stride != 1 || !op.skip0 || !op.skip_w)) CCDM_FAIL(-2, "conv: bad skip configuration");
    if (op.Cout % 4 && op.out_dtype != CCDM_DT_F32) CCDM_FAIL(-2, "conv: ragged Cout needs fp32 output");
    if (op.ostat && (!op.part || !op.ticket)) CCDM_FAIL(-2, "conv: ostat without scratch");
    if (op.emb && (!op.steps || !op.step_ptr || op.emb_off < 0)) CCDM_FAIL(-2, "conv: emb without step table");
    {
        int expH = op.upsample ? op.Hin * 2 : (op.stride == 2 ? (op.Hin + 1) / 2 : op.Hin);
        int expW = op.upsample ? op.Win * 2 : (op.stride == 2 ? (op.Win + 1) / 2 : op.Win);
        if (expH != op.Hout || expW != op.Wout) CCDM_FAIL(-2, "conv: output %dx%d inconsistent with input %dx%d", op.Hout, op.Wout, op.Hin, op.Win);
    }
    const bool bf = op.dtype == CCDM_DT_BF16;
    if (op.ksize == 1) return bf ? launch_t<__nv_bfloat16, 1, 1>(p, s) : launch_t<float, 1, 1>(p, s);
    if (op.stride == 1) return bf ? launch_t<__nv_bfloat16, 3, 1>(p, s) : launch_t<float, 3, 1>(p, s);
    return bf ? launch_t<__nv_bfloat16, 3, 2>(p, s) : launch_t<float, 3, 2>(p, s);
}

}  // namespace ccdm
