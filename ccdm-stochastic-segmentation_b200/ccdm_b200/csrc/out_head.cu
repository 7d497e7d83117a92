// CCDM_OP_OUT_HEAD: the UNet's output head and the categorical head of the reverse step in ONE launch, for small class
// counts (K <= 4: LIDC) in the fp16x2 ("exact") mode.
//
// Replaces (reference): unet.py:701-707 `out` = GroupNorm32 + SiLU + conv3x3(C -> K) + Softmax, then
// diffusion_denoising.py:197-212 (theta_post_prob, clamp, sample / argmax) -- i.e. the two launches
// CCDM_OP_CONV (C -> K, fp32 logits) + CCDM_OP_HEAD.
//
// Why: on the tensor-core kernel a 32 -> 2 conv costs what a 32 -> 32 conv costs (its time is the A-operand fetch of nine
// shifted windows per 16 input channels, whatever N is): 91 us at LIDC's 128x128, B = 64, plus 21 us for the head reading the
// logits back.  With two output channels the conv is 576 FMAs per pixel: CUDA cores do that at the speed the input can be
// read, and the logits never leave the registers (they are still written, 8 bytes per pixel, for the tests' record mode).
//
// One CTA = a 32 x 16 pixel tile of one sample, 128 threads.  Per half of the input channels (16): the tile + halo is loaded
// (hi + lo rows, 16-byte loads along a tile row), GroupNorm-affine + SiLU applied in fp32 exactly as the tensor-core conv's
// transform does (x / (1 + exp(-x))), positions outside the image set to zero (the reference pads after the activation), and
// staged in shared memory as fp32 [18][37][16 + 4]; then every thread accumulates a horizontal strip of 4 pixels: per input
// row and 4-channel group 6 LDS.128 of activations + 12 broadcast loads of weights feed 48 K FMAs.  Thread t owns strip
// t / 16 of row t % 16, and a tile row is 37 positions of 20 floats: both the staging stores (consecutive positions) and the
// strip loads (8 consecutive rows per quarter warp) are bank-conflict free.  Then the four pixels go through head_pixel --
// the same bit-exact posterior / draw code as head_kernel.
#include "conv_tc_common.cuh"
#include "head_common.cuh"

namespace ccdm {
namespace {

constexpr int OH_TW = 32, OH_TH = 16;            // output tile
constexpr int OH_WP = 37, OH_HP = OH_TH + 2;     // staged positions per tile row (34 used; 37: see above), staged rows
constexpr int OH_CH = 16, OH_PITCH = OH_CH + 4;  // channels per pass, floats per staged position
constexpr int OH_THREADS = 128;
constexpr int OH_MAXC = 64;

struct OhP {
    WsP w;              // the GroupNorm fold's view of the op (gn_build_affine)
    HeadP h;
    const float *weight;  // fp32 [9][ceil8(C)][32] (the FFMA layout of the output conv)
    const float *bias;    // fp32 [32]
    float *logits;        // fp32 [B, H, W, K] or null
    int C, CP, H, W, tiles_x;
};

template <int KMAX>
__global__ void __launch_bounds__(OH_THREADS) out_head_kernel(const OhP p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double2 *sSum = reinterpret_cast<double2 *>(smem_raw);                       // [C]
    float *sAff = reinterpret_cast<float *>(sSum + p.C);                          // [2 C]
    float *sW = sAff + 2 * p.C;                                                   // [9][C][KMAX]
    float *sAct = sW + 9 * p.C * KMAX;                                            // [OH_HP][OH_WP][OH_PITCH]
    const int t = threadIdx.x;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int y0 = (tile / p.tiles_x) * OH_TH, x0 = (tile % p.tiles_x) * OH_TW;
    const int H = p.H, W = p.W, C = p.C, K = p.h.K;

    for (int i = t; i < 9 * C * KMAX; i += OH_THREADS) {
        const int k = i % KMAX, c = (i / KMAX) % C, tap = i / (KMAX * C);
        sW[i] = k < K ? __ldg(p.weight + (size_t(tap) * p.CP + c) * 32 + k) : 0.f;
    }
    pdl_launch_dependents();
    pdl_wait();
    float alpha = 0.f, cum = 1.f;
    int mode = CCDM_DRAW_X0;
    uint32_t draw = 0;
    {
        const ccdm_step_entry se = p.h.steps[*p.h.step_ptr];
        alpha = se.alpha_t; cum = se.cumalpha_tm1; mode = se.mode; draw = se.draw;
    }
    gn_build_affine(p.w, b, sAff, sSum, t, OH_THREADS, 1);  // sAff[c] = gamma * rstd / 16, sAff[C + c] = beta - mean * gamma * rstd

    const int row = t % OH_TH, strip = t / OH_TH;  // this thread's output row and 4-pixel strip of the tile
    float acc[4][KMAX];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int k = 0; k < KMAX; ++k) acc[px][k] = 0.f;

    const __half *src = reinterpret_cast<const __half *>(p.w.src0);
    const size_t plane = size_t(H) * W * 8;  // halves per (group, hi|lo) plane
    const int G = C / 8;
    for (int c0 = 0; c0 < C; c0 += OH_CH) {
        __syncthreads();  // the previous pass's readers are done (first pass: sAff / sW are complete)
        // ---- stage 34 x 18 positions x 16 channels, transformed ------------------------------------------------------
        for (int it = t; it < 2 * OH_HP * (OH_TW + 2); it += OH_THREADS) {
            const int g = it / (OH_HP * (OH_TW + 2)), pos = it - g * (OH_HP * (OH_TW + 2));
            const int r = pos / (OH_TW + 2), c = pos - r * (OH_TW + 2);
            const int gy = y0 - 1 + r, gx = x0 - 1 + c;
            float v[8];
            if (unsigned(gy) < unsigned(H) && unsigned(gx) < unsigned(W)) {
                const int gg = c0 / 8 + g;
                const __half *ph = src + (size_t(b) * G + gg) * 2 * plane + (size_t(gy) * W + gx) * 8;
                const uint4 hi = *reinterpret_cast<const uint4 *>(ph);
                const uint4 lo = *reinterpret_cast<const uint4 *>(ph + plane);
                const uint32_t h4[4] = {hi.x, hi.y, hi.z, hi.w}, l4[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 xh = unpack_f16x2(h4[i]), xl = unpack_f16x2(l4[i]);
                    const int ch = gg * 8 + 2 * i;
                    float h0 = fmaf(xh.x, sAff[ch], fmaf(xl.x, sAff[ch], sAff[C + ch]));
                    float h1 = fmaf(xh.y, sAff[ch + 1], fmaf(xl.y, sAff[ch + 1], sAff[C + ch + 1]));
                    v[2 * i] = __fdividef(h0, 1.0f + __expf(-h0));       // the transform of conv_tma's fp16x2 path (xf_row_x3)
                    v[2 * i + 1] = __fdividef(h1, 1.0f + __expf(-h1));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            float *dst = sAct + (size_t(r) * OH_WP + c) * OH_PITCH + g * 8;
            *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
        // ---- accumulate --------------------------------------------------------------------------------------------------
#pragma unroll 1
        for (int dy = 0; dy < 3; ++dy) {
            const float *arow = sAct + (size_t(row + dy) * OH_WP + strip * 4) * OH_PITCH;
#pragma unroll
            for (int c4 = 0; c4 < OH_CH / 4; ++c4) {
                float4 a[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) a[j] = *reinterpret_cast<const float4 *>(arow + j * OH_PITCH + c4 * 4);
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float *wp = sW + (size_t(dy * 3 + dx) * C + c0 + c4 * 4) * KMAX;
                    float wv[4][KMAX];
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) wv[ch][k] = wp[ch * KMAX + k];
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float4 av = a[px + dx];
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) {
                            acc[px][k] = fmaf(av.x, wv[0][k], acc[px][k]);
                            acc[px][k] = fmaf(av.y, wv[1][k], acc[px][k]);
                            acc[px][k] = fmaf(av.z, wv[2][k], acc[px][k]);
                            acc[px][k] = fmaf(av.w, wv[3][k], acc[px][k]);
                        }
                    }
                }
            }
        }
    }

    // ---- logits -> posterior -> draw -------------------------------------------------------------------------------------
    const int gy = y0 + row;
    if (gy < H) {
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            const int gx = x0 + strip * 4 + px;
            if (gx < W) {
                float v[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) v[k] = k < K ? acc[px][k] + __ldg(p.bias + k) : 0.f;
                const uint32_t i = (uint32_t(b) * uint32_t(H) + uint32_t(gy)) * uint32_t(W) + uint32_t(gx);
                if (p.logits != nullptr) {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k)
                        if (k < K) p.logits[size_t(i) * K + k] = v[k];
                }
                head_pixel<KMAX>(p.h, v, i, alpha, cum, mode, draw);
            }
        }
    }
    head_step_advance(p.h, gridDim.x * gridDim.y);
}

size_t oh_smem(int C, int KMAX) {
    return sizeof(double2) * C + sizeof(float) * (2 * C + 9 * C * KMAX + OH_HP * OH_WP * OH_PITCH);
}

}  // namespace

bool out_head_supported(const ccdm_op &op) {
    return op.kind == CCDM_OP_OUT_HEAD && op.dtype == CCDM_DT_F16X2 && op.K >= 2 && op.K <= 4 && op.C0 > 0 && op.C0 <= OH_MAXC &&
           (op.C0 % OH_CH) == 0 && op.C1 == 0 && op.ksize == 3 && op.stride == 1 && !op.upsample && op.gn && op.silu &&
           op.Hin == op.Hout && op.Win == op.Wout && op.B <= 65535;
}

int launch_out_head(const ccdm_op &op, cudaStream_t s) {
    if (!out_head_supported(op)) CCDM_FAIL(-3, "out_head: unsupported configuration (fp16x2, K <= 4, C <= %d, 3x3, GroupNorm + SiLU)", OH_MAXC);
    OhP p{};
    p.w.src0 = (const __nv_bfloat16 *)op.src0;
    p.w.stat0 = (const double *)op.stat0;
    p.w.gamma = (const float *)op.gamma; p.w.beta = (const float *)op.beta;
    p.w.B = op.B; p.w.Hin = op.Hin; p.w.Win = op.Win; p.w.H = op.Hin; p.w.W = op.Win;
    p.w.C0 = op.C0; p.w.C1 = 0; p.w.Cin = op.C0;
    p.w.gn = 1; p.w.silu = 3;  // (not a tanh form: the fold must not fold a 1/2 into the affine)
    p.w.x3 = 1;
    p.w.gn_cpg = op.gn_cpg; p.w.gn_off = op.gn_off;
    for (int i = 0; i < 2; ++i) {
        p.w.st_slots[i] = op.st_slots[i]; p.w.st_ips[i] = op.st_ips[i]; p.w.st_items[i] = op.st_items[i];
        p.w.st_grid[i] = op.st_grid[i]; p.w.st_rows[i] = op.st_rows[i];
    }
    if (!op.src0 || !op.stat0 || !op.gamma || !op.beta || !op.weight || !op.bias) CCDM_FAIL(-2, "out_head: missing tensors");
    if (op.gn_cpg <= 0 && (op.C0 % kGnGroups)) CCDM_FAIL(-2, "out_head: GroupNorm needs C %% 32 == 0");
    if (op.st_slots[0] > 0 && (op.st_ips[0] <= 0 || op.st_items[0] <= 0 || op.st_grid[0] <= 0 || op.st_rows[0] <= 0))
        CCDM_FAIL(-2, "out_head: bad deferred-statistics layout");
    HeadP &h = p.h;
    h.labels_in = (const uint8_t *)op.labels_in;
    h.labels_out = (uint8_t *)op.labels_out;
    h.noise = (const float *)op.noise;
    h.probs_out = (float *)op.probs_out;
    h.noise_out = (float *)op.noise_out;
    h.steps = (const ccdm_step_entry *)op.steps;
    h.step_ptr = (const int *)op.step_ptr;
    h.seed = op.seed;
    h.sample0 = uint32_t(op.sample0);
    h.n_pix = uint32_t(op.Hin) * uint32_t(op.Win);
    h.n_total = h.n_pix * uint32_t(op.B);
    h.K = op.K;
    h.from_logits = 1;
    h.noise_mode = op.noise_mode;
    h.fast = op.exact ? 0 : 1;
    h.step_advance = (int *)op.step_ptr;
    h.step_ticket = (unsigned int *)op.ticket;
    if (!h.steps || !h.step_ptr || !h.step_ticket) CCDM_FAIL(-2, "out_head: needs the step table and a ticket");
    if (!h.labels_in || !h.labels_out) CCDM_FAIL(-2, "out_head: missing label maps");
    if (op.noise_mode == CCDM_NOISE_TENSOR && !h.noise) CCDM_FAIL(-2, "out_head: tensor noise mode without a noise tensor");
    if (double(op.B) * op.Hin * op.Win >= 4294967296.0) CCDM_FAIL(-2, "out_head: too many pixels");
    p.weight = (const float *)op.weight;
    p.bias = (const float *)op.bias;
    p.logits = (float *)op.out;
    p.C = op.C0; p.CP = (op.C0 + 7) / 8 * 8; p.H = op.Hin; p.W = op.Win;
    p.tiles_x = (op.Win + OH_TW - 1) / OH_TW;
    const int tiles = p.tiles_x * ((op.Hin + OH_TH - 1) / OH_TH);
    dim3 grid(unsigned(tiles), unsigned(op.B));
    if (op.K <= 2) {
        const size_t smem = oh_smem(p.C, 2);
        static bool done2 = false;
        if (!done2) {
            CCDM_CUDA(cudaFuncSetAttribute(out_head_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(oh_smem(OH_MAXC, 2))));
            done2 = true;
        }
        CCDM_CUDA(launch_pdl(out_head_kernel<2>, grid, dim3(OH_THREADS), smem, s, p));
    } else {
        const size_t smem = oh_smem(p.C, 4);
        static bool done4 = false;
        if (!done4) {
            CCDM_CUDA(cudaFuncSetAttribute(out_head_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(oh_smem(OH_MAXC, 4))));
            done4 = true;
        }
        CCDM_CUDA(launch_pdl(out_head_kernel<4>, grid, dim3(OH_THREADS), smem, s, p));
    }
    CCDM_LAUNCH_CHECK("out_head_kernel");
    return 0;
}

}  // namespace ccdm
