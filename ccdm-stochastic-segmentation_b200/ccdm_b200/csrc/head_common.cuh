// Per-pixel arithmetic of the categorical head, shared by head_kernel (head.cu: logits read from memory) and
// out_head_kernel (out_head.cu: logits computed in registers by the fused output conv).
#pragma once

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

// One pixel: v = the K logits (from_logits) or probabilities of pixel i; everything downstream of the logits of
// head_kernel's original body -- softmax, closed-form posterior, clamp, normalise, draw / argmax, the optional exports.
template <int KMAX>
__device__ __forceinline__ void head_pixel(const HeadP &p, float (&v)[KMAX], uint32_t i, float alpha, float cum, int mode, uint32_t draw) {
        const int K = p.K;
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

// End of a reverse step: the last CTA to finish bumps the device step counter, so the next replay of the captured graph reads
// the next row of the step table.  Every thread of the grid loaded its row before its CTA's arrival, so no reader can observe
// the new value.  Call from every thread of every CTA, after the CTA's last use of the step table.
__device__ __forceinline__ void head_step_advance(const HeadP &p, unsigned int n_ctas) {
    if (p.step_advance == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        unsigned int prev = atomicAdd(p.step_ticket, 1u);
        if (prev == n_ctas - 1) {
            *p.step_ticket = 0u;
            *p.step_advance = *p.step_advance + 1;
        }
    }
}

}  // namespace
}  // namespace ccdm
