// The categorical head of a reverse step, one thread per pixel, everything in
// registers: softmax over K logits -> closed-form posterior theta_post_prob ->
// clamp(1e-12) -> normalise -> exponential-race draw (or argmax / probabilities
// on the last step) -> uint8 label.  One read of the logits, one byte written.
//
// Replaces (reference):
//   unet.py:706                       nn.Softmax(dim=1)
//   diffusion_denoising.py:99-128     DiffusionModel.theta_post_prob  ([B,K,K,H,W] einsum,
//                                     here the O(K) closed form of SURVEY.md 8a-10)
//   diffusion_denoising.py:204        torch.clamp(probs, min=1e-12)
//   one_hot_categorical.py:30-54      sample / max_prob_sample / prob_sample
//   torch.distributions.Categorical   probs / probs.sum(-1); torch.multinomial(n=1)
//                                     == argmax(p / E), E ~ Exp(1)
//
// The posterior and draw use explicit round-to-nearest fp32 intrinsics (no FMA
// contraction) in exactly the operation order of oracle/ccdm_oracle.c
// (ccdm_oracle_posterior_closed, ccdm_oracle_draw), so labels and probabilities
// can be compared bit for bit when both sides are fed the same theta and noise.
#include "common.cuh"

namespace ccdm {
namespace {

struct HeadP {
    const float *in;        // logits or theta, [n, K]
    const uint8_t *labels_in;
    uint8_t *labels_out;
    const float *noise;     // [n, K] or null
    float *probs_out;       // [n, K] or null
    float *noise_out;       // [n, K] or null
    const ccdm_step_entry *steps;  // null: use the immediates below
    const int *step_ptr;
    int *step_advance;      // non-null: the last CTA bumps this counter (end of a reverse step)
    unsigned int *step_ticket;
    float alpha_t, cumalpha_tm1;
    int mode;
    uint32_t draw;
    uint64_t seed;
    uint32_t sample0;
    uint32_t n_pix;         // pixels per sample
    uint32_t n_total;       // B * n_pix
    int K;
    int from_logits;
    int noise_mode;
    int fast;  // ccdm_op::exact == 0: sampling steps may use the fast-math path below
};

// Sampling step in fast maths (bf16 engine mode, ccdm_op::exact == 0): the same quantities as the exact path --
// softmax, closed-form posterior, clamp(1e-12), exponential race on the SAME Philox bits -- with approximate
// exp2 / log2 / reciprocal instead of IEEE divisions (the exact path spends ~4 divisions per class), and without
// the final normalisation, a positive factor common to all classes that cannot change the argmax.  Labels differ
// from the exact path only where two race scores agree to ~1e-6 relative, far below what bf16 logits resolve.
template <int KMAX>
__device__ __forceinline__ int head_sample_fast(const float (&x)[KMAX], int K, int lab, float alpha, float cum, uint32_t pix, uint32_t smp,
                                                uint32_t draw, uint64_t seed) {
    constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
    float m = x[0];
#pragma unroll
    for (int c = 1; c < KMAX; ++c)
        if (c < K) m = fmaxf(m, x[c]);
    float v[KMAX];
    float s = 0.f;
    const float mo = m * kLog2e;
#pragma unroll
    for (int c = 0; c < KMAX; ++c)
        if (c < K) {
            v[c] = exp2f_approx(fmaf(x[c], kLog2e, -mo));
            s += v[c];
        }
    const float Kf = float(K);
    const float ua = (1.0f - alpha) / Kf, u = (1.0f - cum) / Kf;
    const float a_hit = alpha + ua, a_miss = ua;
    const float inv_s = __fdividef(1.0f, s);
    const float w_hit = __fdividef(inv_s, fmaf(cum, a_hit, u)), w_miss = __fdividef(inv_s, fmaf(cum, a_miss, u));
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < KMAX; ++c)
        if (c < K) {
            v[c] *= (c == lab) ? w_hit : w_miss;  // softmax / z
            S += v[c];
        }
    const float uS = u * S;
    const uint2 key = make_uint2(uint32_t(seed), uint32_t(seed >> 32));
    int best = 0;
    float bestv = -1.0f;
#pragma unroll
    for (int cb = 0; cb < (KMAX + 3) / 4; ++cb)
        if (cb * 4 < K) {
            const uint4 r = philox4x32_10(make_uint4(pix, smp, draw, uint32_t(cb)), key);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = cb * 4 + j;
                if (c < KMAX && c < K) {
                    const float post = fmaxf(((c == lab) ? a_hit : a_miss) * fmaf(cum, v[c], uS), 1e-12f);
                    const float uu = (static_cast<float>(w[j] >> 9) + 0.5f) * 1.1920928955078125e-07f;  // as bits_to_exponential
                    const float e = -kLn2 * log2f_approx(uu);
                    const float score = __fdividef(post, e);
                    if (score > bestv) {
                        bestv = score;
                        best = c;
                    }
                }
            }
        }
    return best;
}

template <int KMAX>
__global__ void __launch_bounds__(128) head_kernel(const HeadP p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    float alpha = p.alpha_t, cum = p.cumalpha_tm1;
    int mode = p.mode;
    uint32_t draw = p.draw;
    if (p.steps != nullptr) {
        const ccdm_step_entry se = p.steps[*p.step_ptr];
        alpha = se.alpha_t;
        cum = se.cumalpha_tm1;
        mode = se.mode;
        draw = se.draw;
    }
    if (i < p.n_total) {
        const int K = p.K;
        float v[KMAX];
        const float *src = p.in + size_t(i) * K;
        if (KMAX % 4 == 0 && K % 4 == 0) {
#pragma unroll
            for (int c = 0; c < KMAX; c += 4)
                if (c < K) {
                    float4 t = *reinterpret_cast<const float4 *>(src + c);
                    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
                }
        } else if (KMAX == 2 && K == 2) {
            float2 t = *reinterpret_cast<const float2 *>(src);
            v[0] = t.x; v[1] = t.y;
        } else {
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) v[c] = src[c];
        }

        if (p.fast && p.from_logits && mode == CCDM_DRAW_SAMPLE && p.noise_mode != CCDM_NOISE_TENSOR && p.noise_out == nullptr &&
            p.labels_out != nullptr && p.steps != nullptr) {
            const uint32_t smp = i / p.n_pix, pix = i - smp * p.n_pix;
            p.labels_out[i] = uint8_t(head_sample_fast<KMAX>(v, K, int(p.labels_in[i]), alpha, cum, pix, p.sample0 + smp, draw, p.seed));
        } else {
        if (p.from_logits) {  // unet.py:706
            float m = v[0];
#pragma unroll
            for (int c = 1; c < KMAX; ++c)
                if (c < K) m = fmaxf(m, v[c]);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) {
                    v[c] = expf(__fsub_rn(v[c], m));
                    s = (c == 0) ? v[c] : __fadd_rn(s, v[c]);
                }
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) v[c] = __fdiv_rn(v[c], s);
        }

        if (mode != CCDM_DRAW_X0) {
            // closed-form posterior, op order == ccdm_oracle_posterior_closed
            const int lab = p.labels_in[i];
            const float Kf = float(K);
            const float ua = __fdiv_rn(__fsub_rn(1.0f, alpha), Kf);
            const float u = __fdiv_rn(__fsub_rn(1.0f, cum), Kf);
            const float a_hit = __fadd_rn(__fmul_rn(alpha, 1.0f), ua);
            const float a_miss = __fadd_rn(__fmul_rn(alpha, 0.0f), ua);
            const float z_hit = __fadd_rn(__fmul_rn(cum, a_hit), u);
            const float z_miss = __fadd_rn(__fmul_rn(cum, a_miss), u);
            float S = 0.f;
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) {
                    float r = __fdiv_rn(v[c], c == lab ? z_hit : z_miss);
                    v[c] = r;
                    S = (c == 0) ? r : __fadd_rn(S, r);
                }
            const float uS = __fmul_rn(u, S);
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) {
                    float ac = (c == lab) ? a_hit : a_miss;
                    v[c] = __fmul_rn(ac, __fadd_rn(__fmul_rn(cum, v[c]), uS));
                }
            if (mode == CCDM_DRAW_POSTERIOR) {  // raw theta_post_prob output (diffusion_denoising.py:128)
                float *dst = p.probs_out + size_t(i) * K;
#pragma unroll
                for (int c = 0; c < KMAX; ++c)
                    if (c < K) dst[c] = v[c];
                return;
            }
            // clamp + normalise (diffusion_denoising.py:204, Categorical.__init__)
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) {
                    v[c] = v[c] < 1e-12f ? 1e-12f : v[c];
                    s = (c == 0) ? v[c] : __fadd_rn(s, v[c]);
                }
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) v[c] = __fdiv_rn(v[c], s);
        }

        if (p.probs_out != nullptr && (mode == CCDM_DRAW_CONFIDENCE || mode == CCDM_DRAW_X0 || p.steps == nullptr)) {
            float *dst = p.probs_out + size_t(i) * K;
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) dst[c] = v[c];
        }

        if (mode == CCDM_DRAW_SAMPLE) {
            float e[KMAX];
            if (p.noise_mode == CCDM_NOISE_TENSOR) {
                const float *nz = p.noise + size_t(i) * K;
#pragma unroll
                for (int c = 0; c < KMAX; ++c)
                    if (c < K) e[c] = nz[c];
            } else {
                const uint32_t smp = i / p.n_pix, pix = i - smp * p.n_pix;
                const uint2 key = make_uint2(uint32_t(p.seed), uint32_t(p.seed >> 32));
#pragma unroll
                for (int cb = 0; cb < (KMAX + 3) / 4; ++cb)
                    if (cb * 4 < K) {
                        uint4 r = philox4x32_10(make_uint4(pix, p.sample0 + smp, draw, uint32_t(cb)), key);
                        uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (cb * 4 + j < KMAX) e[cb * 4 + j] = bits_to_exponential(w[j]);
                    }
            }
            if (p.noise_out != nullptr) {
                float *dst = p.noise_out + size_t(i) * K;
#pragma unroll
                for (int c = 0; c < KMAX; ++c)
                    if (c < K) dst[c] = e[c];
            }
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K) v[c] = __fdiv_rn(v[c], e[c]);
        }
        if (p.labels_out != nullptr) {
            int best = 0;
            float bestv = -1.0f;
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < K && v[c] > bestv) {
                    bestv = v[c];
                    best = c;
                }
            p.labels_out[i] = uint8_t(best);
        }
        }  // exact path
    }
    if (p.step_advance != nullptr) {
        // End of a reverse step: the last CTA to finish bumps the device step
        // counter, so the next replay of the captured graph reads the next row of
        // the step table.  Every thread of this grid loaded its row above, before
        // its CTA's arrival, so no reader can observe the new value.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned int prev = atomicAdd(p.step_ticket, 1u);
            if (prev == gridDim.x - 1) {
                *p.step_ticket = 0u;
                *p.step_advance = *p.step_advance + 1;
            }
        }
    }
}

template <int KMAX>
int launch_k(const HeadP &p, cudaStream_t s) {
    const int threads = 128;
    const unsigned blocks = (p.n_total + threads - 1) / threads;
    CCDM_CUDA(launch_pdl(head_kernel<KMAX>, dim3(blocks), dim3(threads), 0, s, p));
    CCDM_LAUNCH_CHECK("head_kernel");
    return 0;
}

int launch_headp(const HeadP &p, cudaStream_t s) {
    if (p.K < 2 || p.K > 32) CCDM_FAIL(-2, "head: K=%d unsupported (2..32)", p.K);
    if (p.n_total == 0) return 0;
    if (p.K <= 2) return launch_k<2>(p, s);
    if (p.K <= 4) return launch_k<4>(p, s);
    if (p.K <= 8) return launch_k<8>(p, s);
    if (p.K <= 20) return launch_k<20>(p, s);
    return launch_k<32>(p, s);
}

__global__ void philox_bits_kernel(uint64_t seed, uint32_t draw, uint32_t sample0, uint32_t n_samples, uint32_t n_pix, int K,
                                   uint32_t *bits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples * n_pix) return;
    const uint32_t smp = i / n_pix, pix = i - smp * n_pix;
    const uint2 key = make_uint2(uint32_t(seed), uint32_t(seed >> 32));
    for (int cb = 0; cb * 4 < K; ++cb) {
        uint4 r = philox4x32_10(make_uint4(pix, sample0 + smp, draw, uint32_t(cb)), key);
        uint32_t w[4] = {r.x, r.y, r.z, r.w};
        for (int j = 0; j < 4 && cb * 4 + j < K; ++j) bits[size_t(i) * K + cb * 4 + j] = w[j];
    }
}

// x_T: argmax_c (1/K)/E_c, first maximum wins (== ccdm_oracle_uniform_labels on the same E)
__global__ void uniform_labels_kernel(uint64_t seed, uint32_t draw, uint32_t sample0, uint32_t n_samples, uint32_t n_pix, int K,
                                      uint8_t *labels) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples * n_pix) return;
    const uint32_t smp = i / n_pix, pix = i - smp * n_pix;
    const uint2 key = make_uint2(uint32_t(seed), uint32_t(seed >> 32));
    const float pk = __fdiv_rn(1.0f, float(K));
    int best = 0;
    float bestv = -1.0f;
    for (int cb = 0; cb * 4 < K; ++cb) {
        uint4 r = philox4x32_10(make_uint4(pix, sample0 + smp, draw, uint32_t(cb)), key);
        uint32_t w[4] = {r.x, r.y, r.z, r.w};
        for (int j = 0; j < 4 && cb * 4 + j < K; ++j) {
            float v = __fdiv_rn(pk, bits_to_exponential(w[j]));
            if (v > bestv) {
                bestv = v;
                best = cb * 4 + j;
            }
        }
    }
    labels[i] = uint8_t(best);
}

}  // namespace

int launch_head(const ccdm_op &op, cudaStream_t s) {
    HeadP p{};
    p.in = (const float *)op.src0;
    p.labels_in = (const uint8_t *)op.labels_in;
    p.labels_out = (uint8_t *)op.labels_out;
    p.noise = (const float *)op.noise;
    p.probs_out = (float *)op.probs_out;
    p.noise_out = (float *)op.noise_out;
    p.steps = (const ccdm_step_entry *)op.steps;
    p.step_ptr = (const int *)op.step_ptr;
    p.seed = op.seed;
    p.sample0 = uint32_t(op.sample0);
    p.n_pix = uint32_t(op.Hin) * uint32_t(op.Win);
    p.n_total = p.n_pix * uint32_t(op.B);
    p.K = op.K;
    p.from_logits = 1;
    p.noise_mode = op.noise_mode;
    p.fast = op.exact ? 0 : 1;
    p.step_advance = (int *)op.step_ptr;
    p.step_ticket = (unsigned int *)op.ticket;
    if (!p.steps || !p.step_ptr || !p.step_ticket) CCDM_FAIL(-2, "head: needs the step table and a ticket");
    if (!p.in || !p.labels_in || !p.labels_out) CCDM_FAIL(-2, "head: missing tensors");
    if (op.noise_mode == CCDM_NOISE_TENSOR && !p.noise) CCDM_FAIL(-2, "head: tensor noise mode without a noise tensor");
    return launch_headp(p, s);
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_posterior_draw(const float *theta, const uint8_t *labels_in, size_t n_pix_per_sample, int B, int K,
                                   float alpha_t, float cumalpha_tm1, int mode, int noise_mode, const float *noise,
                                   uint64_t seed, uint32_t draw, uint32_t sample0, uint8_t *labels_out, float *probs_out,
                                   float *noise_out, void *stream) {
    HeadP p{};
    p.in = theta;
    p.labels_in = labels_in;
    p.labels_out = labels_out;
    p.noise = noise;
    p.probs_out = probs_out;
    p.noise_out = noise_out;
    p.alpha_t = alpha_t;
    p.cumalpha_tm1 = cumalpha_tm1;
    p.mode = mode;
    p.draw = draw;
    p.seed = seed;
    p.sample0 = sample0;
    p.n_pix = uint32_t(n_pix_per_sample);
    p.n_total = uint32_t(n_pix_per_sample * size_t(B));
    p.K = K;
    p.from_logits = 0;
    p.noise_mode = noise_mode;
    if (!theta || (mode != CCDM_DRAW_X0 && !labels_in)) CCDM_FAIL(-2, "posterior_draw: missing tensors");
    if (mode == CCDM_DRAW_POSTERIOR && !probs_out) CCDM_FAIL(-2, "posterior_draw: posterior mode needs probs_out");
    if (mode < 0 || mode > CCDM_DRAW_POSTERIOR) CCDM_FAIL(-2, "posterior_draw: bad mode %d", mode);
    if (mode == CCDM_DRAW_SAMPLE && noise_mode == CCDM_NOISE_TENSOR && !noise) CCDM_FAIL(-2, "posterior_draw: no noise tensor");
    if (n_pix_per_sample * size_t(B) >= (size_t(1) << 32)) CCDM_FAIL(-2, "posterior_draw: too many pixels");
    return launch_headp(p, (cudaStream_t)stream);
}

extern "C" int ccdm_philox_bits(uint64_t seed, uint32_t draw, uint32_t sample0, uint32_t n_samples, uint32_t n_pix, int K,
                                uint32_t *bits, void *stream) {
    const uint32_t n = n_samples * n_pix;
    if (n == 0) return 0;
    philox_bits_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, draw, sample0, n_samples, n_pix, K, bits);
    CCDM_LAUNCH_CHECK("philox_bits_kernel");
    return 0;
}

extern "C" int ccdm_uniform_labels(uint64_t seed, uint32_t draw, uint32_t sample0, uint32_t n_samples, uint32_t n_pix, int K,
                                   uint8_t *labels, void *stream) {
    const uint32_t n = n_samples * n_pix;
    if (n == 0) return 0;
    if (K < 2 || K > 255) CCDM_FAIL(-2, "uniform_labels: K=%d", K);
    uniform_labels_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, draw, sample0, n_samples, n_pix, K, labels);
    CCDM_LAUNCH_CHECK("uniform_labels_kernel");
    return 0;
}
