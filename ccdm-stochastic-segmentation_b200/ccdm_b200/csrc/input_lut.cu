// input_blocks[0] on the label map itself (tensor-core modes): conv3x3(cat[one_hot(x_t), image]) without ever building the
// concatenated tensor.
//
// Replaces (reference unet.py:760 `th.cat([x, input_condition])` + :516-518 the first conv).  x_t is one-hot, so its
// branch of the conv is a table lookup: per tap, row `label(pixel + tap)` of that tap's [K, Cout] weight slice is added
// (out-of-image taps add nothing: the reference zero-pads the one-hot planes too).  The image branch is C_img FMAs per tap
// and output channel on a tile kept in shared memory.  One thread = one output pixel x 32 output channels in registers;
// weights ([tap][K + C_img][32] fp32, 3-27 KB) and the haloed label / image tile live in shared memory.  K <= 4: both
// class rows are read as broadcasts and selected (no bank conflicts); larger K gathers the row of the pixel's label.
// Output: the mode's activation storage (bf16 or fp16x2 planes) + the GroupNorm statistics of the output as per-CTA
// partial rows in the deferred-fold layout of the tensor-core convs (conv_tc_common.cuh), so the first ResBlock consumes
// it like any other producer's.  SURVEY.md 8f-1 (one-hot branch as a 9-tap lookup on uint8 labels; the image is indexed by
// sample / img_rep, 8f-2).
//
// This op replaces TWO launches of round 1 (encode_input + a 16/32-channel tensor-core conv) and the round trip of the
// materialised [B, 16|32, H, W] tensor.
#include "tc_common.cuh"

namespace ccdm {
namespace {

constexpr int LT_TH = 8, LT_TW = 32;           // output tile: 8 rows x 32 columns = 256 pixels = 256 threads
constexpr int LT_THREADS = LT_TH * LT_TW;
constexpr int LT_HH = LT_TH + 2, LT_HW = LT_TW + 2;  // haloed tile
constexpr int LT_CO = 32;                      // output channels per pass (registers)

struct LutP {
    const uint8_t *labels;
    const float *image;
    const float *weight;  // [9][CinP][CoutP] fp32 (the FFMA kernels' packing), CinP = ceil8(K + C_img)
    const float *bias;    // [CoutP]
    void *out;
    float *part;          // per-CTA partial statistics rows [B][slots][CoutP][2], or nullptr
    int B, H, W, K, C_img, CinP, Cout, CoutP, img_rep, x3;
    int tiles_x, tiles, n_items, ips, slots;  // items = (sample, tile); one CTA walks a contiguous range
};

template <bool X3, bool SMALLK>
__global__ void __launch_bounds__(LT_THREADS) input_lut_kernel(const LutP p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int n_in = p.K + p.C_img;
    float *sW = reinterpret_cast<float *>(smem_raw);                 // [9][n_in][CoutP]
    float *sImg = sW + 9 * n_in * p.CoutP;                           // [C_img][LT_HH][LT_HW]
    float *sAcc = sImg + p.C_img * LT_HH * LT_HW;                    // [8 warps][CoutP][2]
    uint8_t *sLab = reinterpret_cast<uint8_t *>(sAcc + 8 * p.CoutP * 2);  // [LT_HH][LT_HW], 255 = outside the image
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ty = tid / LT_TW, tx = tid - ty * LT_TW;
    pdl_launch_dependents();
    for (int e = tid; e < 9 * n_in * p.CoutP; e += LT_THREADS) {
        const int co = e % p.CoutP, r = e / p.CoutP, ci = r % n_in, tap = r / n_in;
        sW[e] = p.weight[(size_t(tap) * p.CinP + ci) * p.CoutP + co];
    }
    for (int e = tid; e < 8 * p.CoutP * 2; e += LT_THREADS) sAcc[e] = 0.f;
    pdl_wait();  // labels are written by the previous step's head kernel
    const int it_begin = int((long long)blockIdx.x * p.n_items / gridDim.x);
    const int it_end = int((long long)(blockIdx.x + 1) * p.n_items / gridDim.x);
    const size_t hw = size_t(p.H) * p.W;
    int cur_b = -1, n_pending = 0;

    auto flush = [&](int b) {  // one partial row per (CTA, sample): the consumer's GroupNorm prologue folds them
        __syncthreads();
        const int c_first = int((((long long)b * p.ips + 1) * gridDim.x - 1) / p.n_items);
        const int slot = int(blockIdx.x) - c_first;
        for (int e = tid; e < p.CoutP * 2; e += LT_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                s += sAcc[w * p.CoutP * 2 + e];
                sAcc[w * p.CoutP * 2 + e] = 0.f;
            }
            p.part[(size_t(b) * p.slots + slot) * p.CoutP * 2 + e] = s;
        }
        __syncthreads();
    };

    for (int it = it_begin; it < it_end; ++it) {
        const int b = it / p.tiles, tile = it - b * p.tiles;
        if (b != cur_b) {
            if (n_pending > 0 && p.part != nullptr) flush(cur_b);
            cur_b = b;
            n_pending = 0;
        }
        const int y0 = (tile / p.tiles_x) * LT_TH, x0 = (tile % p.tiles_x) * LT_TW;
        const size_t bi = size_t(b / p.img_rep);
        __syncthreads();  // the previous item's readers are done with the tile
        for (int e = tid; e < LT_HH * LT_HW; e += LT_THREADS) {
            const int r = e / LT_HW, c = e - r * LT_HW;
            const int y = y0 + r - 1, x = x0 + c - 1;
            const bool in = unsigned(y) < unsigned(p.H) && unsigned(x) < unsigned(p.W);
            sLab[e] = in ? p.labels[size_t(b) * hw + size_t(y) * p.W + x] : uint8_t(255);
            for (int ci = 0; ci < p.C_img; ++ci) sImg[ci * LT_HH * LT_HW + e] = in ? p.image[(bi * p.C_img + ci) * hw + size_t(y) * p.W + x] : 0.f;
        }
        __syncthreads();
        const int oy = y0 + ty, ox = x0 + tx;
        const bool valid = oy < p.H && ox < p.W;
        for (int cb = 0; cb < p.CoutP; cb += LT_CO) {
            float acc[LT_CO];
#pragma unroll
            for (int i = 0; i < LT_CO; i += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(p.bias + cb + i);
                acc[i] = t.x; acc[i + 1] = t.y; acc[i + 2] = t.z; acc[i + 3] = t.w;
            }
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const int e = (ty + tap / 3) * LT_HW + tx + tap % 3;
                const int lab = sLab[e];
                const float *wt = sW + size_t(tap) * n_in * p.CoutP + cb;
                if (SMALLK) {
                    // every class row is read as a warp-wide broadcast and selected: no bank conflicts
                    for (int k = 0; k < p.K; ++k) {
                        const float sel = lab == k ? 1.f : 0.f;
#pragma unroll
                        for (int i = 0; i < LT_CO; i += 4) {
                            const float4 t = *reinterpret_cast<const float4 *>(wt + k * p.CoutP + i);
                            acc[i] = fmaf(sel, t.x, acc[i]); acc[i + 1] = fmaf(sel, t.y, acc[i + 1]);
                            acc[i + 2] = fmaf(sel, t.z, acc[i + 2]); acc[i + 3] = fmaf(sel, t.w, acc[i + 3]);
                        }
                    }
                } else if (lab != 255) {
                    const float *row = wt + lab * p.CoutP;
#pragma unroll
                    for (int i = 0; i < LT_CO; i += 4) {
                        const float4 t = *reinterpret_cast<const float4 *>(row + i);
                        acc[i] += t.x; acc[i + 1] += t.y; acc[i + 2] += t.z; acc[i + 3] += t.w;
                    }
                }
                for (int ci = 0; ci < p.C_img; ++ci) {
                    const float v = sImg[ci * LT_HH * LT_HW + e];  // 0 outside the image
                    const float *row = wt + (p.K + ci) * p.CoutP;
#pragma unroll
                    for (int i = 0; i < LT_CO; i += 4) {
                        const float4 t = *reinterpret_cast<const float4 *>(row + i);
                        acc[i] = fmaf(v, t.x, acc[i]); acc[i + 1] = fmaf(v, t.y, acc[i + 1]);
                        acc[i + 2] = fmaf(v, t.z, acc[i + 2]); acc[i + 3] = fmaf(v, t.w, acc[i + 3]);
                    }
                }
            }
            // store: 8-channel planes, a warp writes 512 contiguous bytes per plane (fp16x2: hi plane, then lo plane)
            float s1[16], s2[16];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float v = valid ? acc[half * 16 + i] : 0.f;
                    s1[i] = v;
                    s2[i] = v * v;
                }
                if (p.part != nullptr) {
                    const float r1 = warp_transpose_reduce16(s1, lane);
                    const float r2 = warp_transpose_reduce16(s2, lane);
                    if ((lane & 1) == 0) {
                        const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                        float *a = sAcc + (warp * p.CoutP + cb + half * 16 + ch) * 2;
                        a[0] += r1;
                        a[1] += r2;
                    }
                }
            }
            if (valid) {
                const size_t pix = size_t(oy) * p.W + ox;
#pragma unroll
                for (int g = 0; g < LT_CO / 8; ++g) {
                    const int plane = (cb >> 3) + g;
                    if (plane * 8 >= p.Cout) break;
                    if (X3) {
                        uint32_t ph[4], pl[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) split_f16x2(acc[g * 8 + 2 * i], acc[g * 8 + 2 * i + 1], ph[i], pl[i]);
                        __half *o = reinterpret_cast<__half *>(p.out) + ((size_t(b) * (p.Cout >> 3) + plane) * 2 * hw + pix) * 8;
                        *reinterpret_cast<uint4 *>(o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                        *reinterpret_cast<uint4 *>(o + hw * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    } else {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) pk[i] = pack_bf16(acc[g * 8 + 2 * i], acc[g * 8 + 2 * i + 1]);
                        __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(p.out) + ((size_t(b) * (p.Cout >> 3) + plane) * hw + pix) * 8;
                        *reinterpret_cast<uint4 *>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
        }
        ++n_pending;
    }
    if (n_pending > 0 && p.part != nullptr) flush(cur_b);
}

struct LutCfg {
    int tiles_x, tiles, n_items, grid, ips, slots;
    size_t smem;
};

int lut_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else n = 148;
        (void)cudaGetLastError();
    }
    return n;
}

LutCfg lut_configure(const ccdm_op &op) {
    LutCfg c{};
    const int CoutP = (op.Cout + 31) / 32 * 32;
    c.tiles_x = (op.Wout + LT_TW - 1) / LT_TW;
    c.tiles = c.tiles_x * ((op.Hout + LT_TH - 1) / LT_TH);
    c.n_items = op.B * c.tiles;
    const int cap = 2 * lut_num_sms();  // two 256-thread CTAs per SM (<= 45 KB of shared memory each)
    c.grid = c.n_items < cap ? c.n_items : cap;
    c.ips = c.tiles;
    c.slots = 1;
    for (int b = 0; b < op.B; ++b) {
        const int c_first = int((((long long)b * c.ips + 1) * c.grid - 1) / c.n_items);
        const int c_last = int(((long long)(b + 1) * c.ips * c.grid - 1) / c.n_items);
        if (c_last - c_first + 1 > c.slots) c.slots = c_last - c_first + 1;
    }
    c.smem = sizeof(float) * (size_t(9) * (op.K + op.C_img) * CoutP + size_t(op.C_img) * LT_HH * LT_HW + 8 * size_t(CoutP) * 2) + LT_HH * LT_HW + 16;
    return c;
}

}  // namespace

bool input_lut_supported(const ccdm_op &op) {
    if (op.kind != CCDM_OP_INPUT_LUT) return false;
    if (op.dtype != CCDM_DT_BF16 && op.dtype != CCDM_DT_F16X2) return false;
    if (op.out_dtype != op.dtype || (op.Cout % 16) || op.K < 1 || op.K > 254 || op.C_img < 0) return false;
    if (op.Hin != op.Hout || op.Win != op.Wout) return false;
    return lut_configure(op).smem <= 96 * 1024;
}

int input_lut_stat_layout(const ccdm_op &op, int32_t *out5) {
    if (!input_lut_supported(op)) return -1;
    const LutCfg c = lut_configure(op);
    out5[0] = c.slots; out5[1] = c.ips; out5[2] = c.n_items; out5[3] = c.grid; out5[4] = (op.Cout + 31) / 32 * 32;
    return 0;
}

size_t input_lut_part_floats(const ccdm_op &op) {
    const LutCfg c = lut_configure(op);
    return size_t(op.B) * c.slots * ((op.Cout + 31) / 32 * 32) * 2;
}

int launch_input_lut(const ccdm_op &op, cudaStream_t s) {
    if (!input_lut_supported(op)) CCDM_FAIL(-3, "input_lut: unsupported configuration");
    if (!op.labels_in || !op.weight || !op.bias || !op.out || (op.C_img > 0 && !op.image)) CCDM_FAIL(-2, "input_lut: missing tensors");
    const LutCfg c = lut_configure(op);
    LutP p{};
    p.labels = (const uint8_t *)op.labels_in; p.image = (const float *)op.image; p.weight = (const float *)op.weight;
    p.bias = (const float *)op.bias; p.out = (void *)op.out; p.part = (float *)op.part;
    p.B = op.B; p.H = op.Hout; p.W = op.Wout; p.K = op.K; p.C_img = op.C_img; p.CinP = (op.K + op.C_img + 7) / 8 * 8;
    p.Cout = op.Cout; p.CoutP = (op.Cout + 31) / 32 * 32; p.img_rep = op.img_rep > 1 ? op.img_rep : 1; p.x3 = op.dtype == CCDM_DT_F16X2;
    p.tiles_x = c.tiles_x; p.tiles = c.tiles; p.n_items = c.n_items; p.ips = c.ips; p.slots = c.slots;
    if (c.n_items == 0) return 0;
    const bool x3 = op.dtype == CCDM_DT_F16X2, smallk = op.K <= 4;
    void (*kern)(LutP) = x3 ? (smallk ? input_lut_kernel<true, true> : input_lut_kernel<true, false>)
                            : (smallk ? input_lut_kernel<false, true> : input_lut_kernel<false, false>);
    static bool attr_done = false;
    if (!attr_done) {
        CCDM_CUDA(cudaFuncSetAttribute(input_lut_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CCDM_CUDA(cudaFuncSetAttribute(input_lut_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CCDM_CUDA(cudaFuncSetAttribute(input_lut_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CCDM_CUDA(cudaFuncSetAttribute(input_lut_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_done = true;
    }
    CCDM_CUDA(launch_pdl(kern, dim3(c.grid), dim3(LT_THREADS), c.smem, s, p));
    CCDM_LAUNCH_CHECK("input_lut_kernel");
    return 0;
}

}  // namespace ccdm
