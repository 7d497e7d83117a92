// Small kernels around the reverse step: the timestep-embedding table, layout
// converters at the chain boundary, label <-> one-hot conversion.
#include "tc_common.cuh"

namespace ccdm {
namespace {

// One CTA per table row (= one timestep).  Replaces nn.py:103-121
// (timestep_embedding), unet.py:506-510,758 (time_embed MLP) and, for every
// ResBlock, emb_layers = SiLU -> Linear (unet.py:205-211,251).  The embedding
// depends only on (t, weights), so the whole [rows, sum(Cout)] table is built
// once per chain instead of 3 + n_resblocks tiny GEMMs per step.
__global__ void __launch_bounds__(256) time_table_kernel(const float *__restrict__ t, int mc, const float *__restrict__ w0,
                                                         const float *__restrict__ b0, const float *__restrict__ w2,
                                                         const float *__restrict__ b2, const float *__restrict__ w_all,
                                                         const float *__restrict__ b_all, int cols, float *__restrict__ out) {
    extern __shared__ float sm[];
    const int ed = 4 * mc;
    float *e0 = sm;            // [mc]   sinusoid
    float *h1 = e0 + mc;       // [ed]   silu(time_embed.0)
    float *e2 = h1 + ed;       // [ed]   silu(time_embed.2) == silu(emb)
    const int r = blockIdx.x;
    const float tv = t[r];
    const int half = mc / 2;
    for (int i = threadIdx.x; i < mc; i += blockDim.x) {
        float v = 0.f;
        if (i < 2 * half) {
            int k = i < half ? i : i - half;
            // freqs = exp(-ln(10000) * k / half) in fp32 (nn.py:113-116)
            float f = expf(-9.210340371976184f * float(k) / float(half));
            float a = tv * f;
            v = i < half ? cosf(a) : sinf(a);
        }
        e0[i] = v;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < ed; o += blockDim.x) {
        float s = b0[o];
        for (int i = 0; i < mc; ++i) s += w0[o * mc + i] * e0[i];
        h1[o] = silu_exact(s);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < ed; o += blockDim.x) {
        float s = b2[o];
        for (int i = 0; i < ed; ++i) s += w2[o * ed + i] * h1[i];
        e2[o] = silu_exact(s);
    }
    __syncthreads();
    // one warp per output column: coalesced reads of the weight row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int o = warp; o < cols; o += nwarp) {
        float s = 0.f;
        for (int i = lane; i < ed; i += 32) s += w_all[size_t(o) * ed + i] * e2[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) out[size_t(r) * cols + o] = s + b_all[o];
    }
}

__global__ void onehot_to_labels_kernel(const float *__restrict__ x, int64_t sb, int64_t sk, int64_t sh, int64_t sw, int B, int K,
                                        int H, int W, uint8_t *__restrict__ labels) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t n = size_t(B) * H * W;
    if (i >= n) return;
    const int w = int(i % W), h = int((i / W) % H), b = int(i / (size_t(W) * H));
    const float *p = x + b * sb + h * sh + w * sw;
    int best = 0;
    float bv = p[0];
    for (int k = 1; k < K; ++k) {
        float v = p[k * sk];
        if (v > bv) {
            bv = v;
            best = k;
        }
    }
    labels[i] = uint8_t(best);
}

__global__ void labels_to_onehot_i64_kernel(const uint8_t *__restrict__ labels, size_t n_pix, int K, long long *__restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_pix * K) return;
    const size_t p = i / K;
    const int k = int(i - p * K);
    out[i] = labels[p] == k ? 1ll : 0ll;
}

// NCHW fp32 -> NHWC T, one CTA per (32-pixel strip, sample); per-channel sums are
// accumulated in double with one atomicAdd per (CTA, channel): this runs once per
// chain on the feature condition, so its cost and (tiny) order non-determinism in
// the 17th digit of a double do not matter; the sums are rounded through the same
// double -> float path every step.
template <typename T, bool X3 = false>
__global__ void __launch_bounds__(256) nchw_to_nhwc_stats_kernel(const float *__restrict__ src, int C, int HW, T *__restrict__ dst,
                                                                 double *__restrict__ stat, int rep) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, bs = b / rep;  // bs: entry of `src` this sample reads (rep samples per image)
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        int c = c0 + j, p = p0 + tx;
        tile[j][tx] = (c < C && p < HW) ? src[(size_t(bs) * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int p = p0 + j, c = c0 + tx;
        if (p < HW && c < C) {
            float v = tile[tx][j];
            if (X3) {
                // fp16x2 [B][C/8][2][HW][8]: hi plane then lo plane of every 8-channel group; the statistics describe
                // the fp32 value (hi + lo reproduces it to 2^-23)
                uint32_t hi, lo;
                split_f16x2(v, 0.f, hi, lo);
                __half *d = reinterpret_cast<__half *>(dst) + ((size_t(b) * (C >> 3) + (c >> 3)) * 2 * HW + p) * 8 + (c & 7);
                d[0] = __ushort_as_half(static_cast<unsigned short>(hi & 0xFFFFu));
                d[size_t(HW) * 8] = __ushort_as_half(static_cast<unsigned short>(lo & 0xFFFFu));
            } else if (sizeof(T) == 2) {
                __nv_bfloat16 h = __float2bfloat16_rn(v);
                // bf16 activations are plane-major [B][C/8][HW][8]
                reinterpret_cast<__nv_bfloat16 *>(dst)[((size_t(b) * (C >> 3) + (c >> 3)) * HW + p) * 8 + (c & 7)] = h;
                tile[tx][j] = __bfloat162float(h);
            } else {
                reinterpret_cast<float *>(dst)[(size_t(b) * HW + p) * C + c] = v;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int c = c0 + threadIdx.x;
        if (c < C) {
            double s = 0.0, q = 0.0;
            for (int j = 0; j < 32; ++j)
                if (p0 + j < HW) {
                    double v = tile[threadIdx.x][j];
                    s += v;
                    q += v * v;
                }
            atomicAdd(stat + (size_t(b) * C + c) * 2, s);
            atomicAdd(stat + (size_t(b) * C + c) * 2 + 1, q);
        }
    }
}

// one-hot(labels) ++ image -> plane-major bf16 [B][CP/8][HW][8], CP = ceil16(K + C_img), zero padded.
// One thread per (pixel, plane): consecutive threads write consecutive 16-byte rows.
// X3: fp16x2 [B][CP/8][2][HW][8] -- one thread writes the hi row and the lo row of its (pixel, group).
template <bool X3>
__global__ void __launch_bounds__(256) encode_input_kernel(const uint8_t *__restrict__ labels, const float *__restrict__ image, int K,
                                                           int C_img, int planes, size_t HW, size_t total, int img_rep,
                                                           __nv_bfloat16 *__restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= total) return;
    const size_t pix = i % HW;
    const size_t bg = i / HW;
    const int g = int(bg % planes);
    const size_t b = bg / planes;
    const int lab = labels[b * HW + pix];
    const size_t bi = b / size_t(img_rep);  // conditioning image of this sample
    uint32_t pk[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = g * 8 + 2 * j + e;
            v[e] = c < K ? (c == lab ? 1.f : 0.f) : (c < K + C_img ? image[(bi * C_img + (c - K)) * HW + pix] : 0.f);
        }
        if (X3) {
            split_f16x2(v[0], v[1], pk[j], pl[j]);
        } else {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
            pk[j] = *reinterpret_cast<uint32_t *>(&h);
        }
    }
    if (X3) {
        __nv_bfloat16 *o = out + ((b * planes + g) * 2 * HW + pix) * 8;
        *reinterpret_cast<uint4 *>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4 *>(o + HW * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    } else {
        *reinterpret_cast<uint4 *>(out + i * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// Vote over the N samples of an image: one thread per (image, pixel) counts the labels of its N samples (uint8 reads,
// coalesced across the warp) and writes the K class frequencies (NCHW, coalesced per class) and the majority label.
__global__ void __launch_bounds__(256) vote_kernel(const uint8_t *__restrict__ labels, int N, size_t n_pix, int K, size_t total,
                                                   float *__restrict__ freq, uint8_t *__restrict__ majority) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t img = i / n_pix, pix = i - img * n_pix;
    const uint8_t *src = labels + img * size_t(N) * n_pix + pix;
    const float inv = 1.0f / float(N);
    int best = 0, best_n = -1;
    for (int k = 0; k < K; ++k) {  // K passes over N bytes that stay in L1: no per-thread count array, any K <= 255
        int n = 0;
        for (int s = 0; s < N; ++s) n += (src[size_t(s) * n_pix] == k) ? 1 : 0;
        freq[(img * K + k) * n_pix + pix] = float(n) * inv;
        if (n > best_n) {
            best_n = n;
            best = k;
        }
    }
    if (majority != nullptr) majority[i] = uint8_t(best);
}

// One CTA per (image, sample n, sample m): integer counts of |x==c & y==c|, |x==c|, |y==c| for c = 1..K-1 over the pixels.
// K <= 8: warp ballots, counters in registers (uniform across the warp); larger K: shared-memory histograms.
__global__ void __launch_bounds__(256) pairwise_distance_kernel(const uint8_t *__restrict__ x, const uint8_t *__restrict__ y, int N, int M,
                                                                size_t n_pix, int K, double *__restrict__ dist) {
    __shared__ unsigned int s_cnt[3][256];  // [inter | x | y][class]
    const int m = blockIdx.x % M, n = blockIdx.x / M, b = blockIdx.y;
    const uint8_t *px = x + (size_t(b) * N + n) * n_pix, *py = y + (size_t(b) * M + m) * n_pix;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < 3 * 256; e += 256) (&s_cnt[0][0])[e] = 0u;
    __syncthreads();
    if (K <= 8) {
        unsigned int ci[8] = {0}, cx[8] = {0}, cy[8] = {0};
        const size_t n_round = (n_pix + 255) / 256 * 256;  // every lane takes part in every ballot
        for (size_t i = tid; i < n_round; i += 256) {
            const int vx = i < n_pix ? px[i] : 255, vy = i < n_pix ? py[i] : 255;
#pragma unroll
            for (int c = 1; c < 8; ++c)
                if (c < K) {
                    const unsigned bx = __ballot_sync(0xffffffffu, vx == c), by = __ballot_sync(0xffffffffu, vy == c);
                    ci[c] += __popc(bx & by);
                    cx[c] += __popc(bx);
                    cy[c] += __popc(by);
                }
        }
        if (lane == 0)
            for (int c = 1; c < K; ++c) {
                atomicAdd(&s_cnt[0][c], ci[c]);
                atomicAdd(&s_cnt[1][c], cx[c]);
                atomicAdd(&s_cnt[2][c], cy[c]);
            }
    } else {
        for (size_t i = tid; i < n_pix; i += 256) {
            const int vx = px[i], vy = py[i];
            atomicAdd(&s_cnt[1][vx], 1u);
            atomicAdd(&s_cnt[2][vy], 1u);
            if (vx == vy) atomicAdd(&s_cnt[0][vx], 1u);
        }
    }
    __syncthreads();
    if (tid == 0) {
        double acc = 0.0;
        for (int c = 1; c < K; ++c) {
            const unsigned int inter = s_cnt[0][c], uni = s_cnt[1][c] + s_cnt[2][c] - inter;
            acc += uni == 0u ? 1.0 : double(inter) / double(uni);  // utils.py:131: NaN (0/0) -> 1
        }
        dist[(size_t(b) * N + n) * M + m] = 1.0 - acc / double(K - 1);
    }
}

}  // namespace

int launch_encode_input(const ccdm_op &op, cudaStream_t s) {
    const int CP = op.Cout;
    const bool x3 = op.dtype == CCDM_DT_F16X2;
    if ((op.dtype != CCDM_DT_BF16 && !x3) || (CP % 16) || CP < op.K + op.C_img || op.K < 1) CCDM_FAIL(-2, "encode_input: bad channel counts");
    if (!op.labels_in || !op.image || !op.out) CCDM_FAIL(-2, "encode_input: missing tensors");
    const size_t HW = size_t(op.Hin) * op.Win, total = size_t(op.B) * (CP / 8) * HW;
    if (total == 0) return 0;
    CCDM_CUDA(launch_pdl(x3 ? encode_input_kernel<true> : encode_input_kernel<false>, dim3(unsigned((total + 255) / 256)), dim3(256), 0, s,
                         (const uint8_t *)op.labels_in, (const float *)op.image, op.K, op.C_img, CP / 8, HW, total, op.img_rep > 1 ? op.img_rep : 1,
                         (__nv_bfloat16 *)op.out));
    CCDM_LAUNCH_CHECK("encode_input_kernel");
    return 0;
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_time_table(const float *t, int rows, int model_channels, const float *te0_w, const float *te0_b,
                               const float *te2_w, const float *te2_b, const float *w_all, const float *b_all, int cols,
                               float *out, void *stream) {
    if (rows <= 0 || cols <= 0) return 0;
    if (model_channels <= 0 || model_channels > 1024) CCDM_FAIL(-2, "time_table: model_channels=%d", model_channels);
    size_t smem = sizeof(float) * size_t(model_channels) * 9;
    time_table_kernel<<<rows, 256, smem, (cudaStream_t)stream>>>(t, model_channels, te0_w, te0_b, te2_w, te2_b, w_all, b_all, cols, out);
    CCDM_LAUNCH_CHECK("time_table_kernel");
    return 0;
}

extern "C" int ccdm_onehot_to_labels(const float *x, int64_t sb, int64_t sk, int64_t sh, int64_t sw, int B, int K, int H, int W,
                                     uint8_t *labels, void *stream) {
    const size_t n = size_t(B) * H * W;
    if (n == 0) return 0;
    if (K < 1 || K > 255) CCDM_FAIL(-2, "onehot_to_labels: K=%d", K);
    onehot_to_labels_kernel<<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, sb, sk, sh, sw, B, K, H, W, labels);
    CCDM_LAUNCH_CHECK("onehot_to_labels_kernel");
    return 0;
}

extern "C" int ccdm_labels_to_onehot_i64(const uint8_t *labels, size_t n_pix, int K, int64_t *out, void *stream) {
    const size_t n = n_pix * size_t(K);
    if (n == 0) return 0;
    labels_to_onehot_i64_kernel<<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, n_pix, K, (long long *)out);
    CCDM_LAUNCH_CHECK("labels_to_onehot_i64_kernel");
    return 0;
}

extern "C" int ccdm_nchw_to_nhwc_stats(const float *src, int B, int C, int H, int W, int dtype, void *dst, double *stat, int rep,
                                       void *stream) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    if (rep < 1) rep = 1;
    if (B % rep) CCDM_FAIL(-2, "nchw_to_nhwc_stats: batch %d is not a multiple of rep %d", B, rep);
    cudaStream_t s = (cudaStream_t)stream;
    CCDM_CUDA(cudaMemsetAsync(stat, 0, sizeof(double) * 2 * size_t(B) * C, s));
    const int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
    if (dtype == CCDM_DT_F16X2)
        nchw_to_nhwc_stats_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>(src, C, HW, (__nv_bfloat16 *)dst, stat, rep);
    else if (dtype == CCDM_DT_BF16)
        nchw_to_nhwc_stats_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, C, HW, (__nv_bfloat16 *)dst, stat, rep);
    else
        nchw_to_nhwc_stats_kernel<float><<<grid, 256, 0, s>>>(src, C, HW, (float *)dst, stat, rep);
    CCDM_LAUNCH_CHECK("nchw_to_nhwc_stats_kernel");
    return 0;
}

extern "C" int ccdm_vote(const uint8_t *labels, int B_img, int N, size_t n_pix, int K, float *freq, uint8_t *majority, void *stream) {
    if (B_img <= 0 || n_pix == 0) return 0;
    if (N < 1 || K < 1 || K > 255 || !labels || !freq) CCDM_FAIL(-2, "vote: N=%d K=%d", N, K);
    const size_t total = size_t(B_img) * n_pix;
    vote_kernel<<<unsigned((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, N, n_pix, K, total, freq, majority);
    CCDM_LAUNCH_CHECK("vote_kernel");
    return 0;
}

extern "C" int ccdm_pairwise_distance(const uint8_t *x, const uint8_t *y, int B, int N, int M, size_t n_pix, int K, double *dist,
                                      void *stream) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    if (K < 2 || K > 255 || !x || !y || !dist || n_pix == 0) CCDM_FAIL(-2, "pairwise_distance: K=%d n_pix=%zu", K, n_pix);
    if (size_t(N) * M > 0x7fffffffull || B > 65535) CCDM_FAIL(-2, "pairwise_distance: too many pairs for one launch");
    pairwise_distance_kernel<<<dim3(unsigned(N * M), unsigned(B)), 256, 0, (cudaStream_t)stream>>>(x, y, N, M, n_pix, K, dist);
    CCDM_LAUNCH_CHECK("pairwise_distance_kernel");
    return 0;
}
