// Fused GroupNorm + SiLU + conv (3x3 / 1x1, stride 1, optional nearest-x2 in front) on the
// 5th-generation tensor cores: bf16 activations and weights, fp32 accumulation in TMEM
// (tcgen05.mma, cta_group::1, M = 128), same fusion contract as conv_ffma.cu (bias + timestep
// embedding + identity residual or fused 1x1 skip, GroupNorm statistics of the output).
//
// Replaces the same reference lines as conv_ffma.cu (unet.py:242-262, :106-116, :305-311,
// :701-705) for the bf16 ("fast") precision mode.
//
// Implicit GEMM without im2col -- the "flattened padded tile":
//   A CTA owns a tile of R x Wt output pixels of one sample.  Its input window, (R+2) x (Wt+2)
//   pixels for a 3x3, is staged in shared memory as channel planes of 8 channels:
//       plane[g][q] = 16 bytes = channels 8g..8g+7 of window position q = r*P + c,  P = Wt + 2.
//   That is exactly the UMMA "K-major, no swizzle" canonical layout with SBO = 128 B (8 rows of
//   16 B), LBO = plane stride: MMA row i reads position (start + i), so output position
//   j = o*P + c under tap (dy,dx) reads window position j + dy*P + dx -- a pure shift of the
//   descriptor start address.  One M=128 MMA therefore covers 128 consecutive flattened
//   output positions; the two pad columns per row produce don't-care rows that the epilogue
//   skips.  No im2col, no data duplication, 9 taps x (Cin/16) MMAs per 128 positions.
//   Weights are pre-packed [tap][Cin/8][Cout][8] so a (tap, 16-channel) slice is the same
//   canonical layout with N = Cout rows.
//
// Pipeline (per CTA, 256 threads): input channels stream through two smem buffers in chunks of
// KC channels.  All warps load+normalise+SiLU chunk k+1 (LDG.128 -> fp32 affine + tanh-form SiLU
// -> bf16 -> STS.128) while the tensor core consumes chunk k (one thread issues the MMAs and
// commits to an mbarrier that releases the buffer).  The epilogue reads the accumulators with
// tcgen05.ld (thread = one output position x 16 channels), adds bias/embedding/residual, stores
// bf16 NHWC and reduces the per-channel statistics deterministically.
#include "common.cuh"

namespace ccdm {
namespace {

constexpr int TC_THREADS = 256;
constexpr int CGW = 16;  // accumulator columns per tcgen05.ld in the epilogue

struct TcP {
    const __nv_bfloat16 *src0, *src1;
    const double *stat0, *stat1;
    const float *gamma, *beta;
    const __nv_bfloat16 *weight;  // [tap][Cin/8][CoutP][8]
    const float *bias, *emb;
    const __nv_bfloat16 *skip0, *skip1;
    const __nv_bfloat16 *skip_w;  // [S/8][CoutP][8]
    const __nv_bfloat16 *res;
    void *out;
    double *ostat;
    float *part;
    unsigned int *ticket;
    const ccdm_step_entry *steps;
    const int *step_ptr;
    int B, Hin, Win, H, W;  // H, W: conv-input == output space (after the optional x2)
    int C0, C1, Cin, Cout, CoutP, NT;
    int upsample, gn, silu, S0, S1, emb_off, emb_cols, emb_bstride, out_f32;
    int R, Wt, P, MB, WN, tiles_x, tiles_y, taps, pad;
    int tmem_cols;
    uint32_t a_region;  // bytes reserved for the two A buffers (>= the epilogue's transpose scratch)
    uint32_t idesc;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// K-major, SWIZZLE_NONE smem descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) | (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (uint64_t(1) << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}

// ---- the kernel -------------------------------------------------------------------------------
// PL = planes (8-channel groups) per K chunk: KC = 8*PL channels.
template <int PL>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(const TcP p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int KC = 8 * PL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int y0 = ty * p.R, x0 = tx * p.Wt;
    const int co0 = chunk * p.NT;
    const int NT = p.NT, WN = p.WN, P = p.P;

    // smem carve-up
    const uint32_t a_bytes = uint32_t(PL) * WN * 16;            // one A buffer
    const uint32_t w_bytes = uint32_t(p.taps) * PL * NT * 16;   // one W buffer
    uint8_t *sAbuf = smem_raw;                                   // [2][PL][WN][16B]
    uint8_t *sWbuf = sAbuf + p.a_region;                         // [2][taps][PL][NT][16B]
    float *sAff = reinterpret_cast<float *>(sWbuf + 2 * w_bytes);  // [2][Cin]  GN scale / shift
    float *sAdd = sAff + 2 * p.Cin;                              // [NT] bias (+ embedding)
    float *sRed = sAdd + NT;                                     // [8 warps][NT][2]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRed + 8 * NT * 2);  // free[0], free[1], done
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 3);
    int *s_last = reinterpret_cast<int *>(s_tmem + 1);

    if (warp == 0) tmem_alloc(s_tmem, uint32_t(p.tmem_cols));
    if (tid == 32) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        fence_barrier_init();
    }
    // GroupNorm scale/shift of the (concatenated) input; the SiLU's 0.5 is folded in:
    // silu(x) = h + h*tanh(h), h = x/2.
    if (p.gn) {
        const int cpg = p.Cin / kGnGroups;
        const double n = double(cpg) * double(p.Hin) * double(p.Win);
        const float half = p.silu ? 0.5f : 1.0f;
        for (int c = tid; c < p.Cin; c += TC_THREADS) {
            const int g0 = (c / cpg) * cpg;
            double s = 0.0, q = 0.0;
            for (int j = 0; j < cpg; ++j) {
                const int cc = g0 + j;
                const double *st = cc < p.C0 ? p.stat0 + (size_t(b) * p.C0 + cc) * 2 : p.stat1 + (size_t(b) * p.C1 + (cc - p.C0)) * 2;
                s += st[0];
                q += st[1];
            }
            const double mean = s / n;
            double var = q / n - mean * mean;
            var = var < 0.0 ? 0.0 : var;
            const float rstd = float(1.0 / sqrt(var + double(kGnEps)));
            const float a = p.gamma[c] * rstd;
            sAff[c] = half * a;
            sAff[p.Cin + c] = half * (p.beta[c] - float(mean) * a);
        }
    }
    for (int c = tid; c < NT; c += TC_THREADS) {
        float v = p.bias[co0 + c];
        if (p.emb != nullptr && co0 + c < p.Cout) {
            const ccdm_step_entry &se = p.steps[*p.step_ptr];
            v += p.emb[(size_t(se.emb_row) + size_t(b) * p.emb_bstride) * p.emb_cols + p.emb_off + co0 + c];
        }
        sAdd[c] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int n_main = p.Cin / KC;
    const int n_skip = (p.S0 + p.S1) / KC;
    const int n_chunks = n_main + n_skip;
    const int g = tid & (PL - 1);                 // this thread's channel plane inside a chunk
    const int q_first = tid / PL, q_step = TC_THREADS / PL;
    const int win_positions = (p.R + 2 * p.pad) * P;

    for (int kc = 0; kc < n_chunks; ++kc) {
        const int buf = kc & 1;
        if (kc >= 2) {
            mbar_wait(bars + buf, uint32_t(((kc >> 1) - 1) & 1));
            tc_fence_after();
        }
        const bool is_skip = kc >= n_main;
        const int cbase = is_skip ? (kc - n_main) * KC : kc * KC;   // channel offset inside its concat space
        const int CA = is_skip ? p.S0 : p.C0;
        const bool first = cbase < CA;
        const __nv_bfloat16 *src = is_skip ? (first ? p.skip0 : p.skip1) : (first ? p.src0 : p.src1);
        const int Cs = is_skip ? (first ? p.S0 : p.S1) : (first ? p.C0 : p.C1);
        const int cs = (first ? cbase : cbase - CA) + 8 * g;
        const bool do_gn = p.gn && !is_skip, do_silu = p.silu && !is_skip;
        float fa[8], fb[8];
        if (do_gn) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                fa[i] = sAff[cbase + 8 * g + i];
                fb[i] = sAff[p.Cin + cbase + 8 * g + i];
            }
        } else {
            const float h = do_silu ? 0.5f : 1.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) fa[i] = h, fb[i] = 0.f;
        }
        // ---- A: window of this chunk, normalised + activated, into plane g -------------
        uint8_t *dstA = sAbuf + buf * a_bytes + uint32_t(g) * WN * 16;
        // the skip conv is a 1x1 on the block input at the output resolution (no x2)
        const int srcH = is_skip ? p.H : p.Hin, srcW = is_skip ? p.W : p.Win;
        const bool ups = p.upsample && !is_skip;
        for (int q = q_first; q < win_positions; q += q_step) {
            const int r = q / P, c = q - r * P;
            const int y = y0 - p.pad + r, x = x0 - p.pad + c;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
                const int sy = ups ? (y >> 1) : y, sx = ups ? (x >> 1) : x;
                const uint4 raw = *reinterpret_cast<const uint4 *>(src + ((size_t(b) * srcH + sy) * srcW + sx) * Cs + cs);
                if (do_gn || do_silu) {
                    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
                    uint32_t r4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float2 v = unpack_bf16(w4[i]);
                        float h0 = fmaf(v.x, fa[2 * i], fb[2 * i]);
                        float h1 = fmaf(v.y, fa[2 * i + 1], fb[2 * i + 1]);
                        if (do_silu) {
                            h0 = fmaf(h0, tanh_approx(h0), h0);
                            h1 = fmaf(h1, tanh_approx(h1), h1);
                        }
                        r4[i] = pack_bf16(h0, h1);
                    }
                    o = make_uint4(r4[0], r4[1], r4[2], r4[3]);
                } else {
                    o = raw;
                }
            }
            *reinterpret_cast<uint4 *>(dstA + size_t(q) * 16) = o;
        }
        // ---- W: [tap][PL][NT][8] slice of the packed weights -----------------------------
        {
            uint8_t *dstW = sWbuf + buf * w_bytes;
            const int ntap = is_skip ? 1 : p.taps;
            const __nv_bfloat16 *wsrc = is_skip ? p.skip_w : p.weight;
            const int planes_total = (is_skip ? (p.S0 + p.S1) : p.Cin) / 8;
            const int plane0 = cbase / 8;
            const int items = ntap * PL * NT;
            for (int e = tid; e < items; e += TC_THREADS) {
                const int n = e % NT, rest = e / NT;
                const int pl = rest % PL, tap = rest / PL;
                const uint4 v = *reinterpret_cast<const uint4 *>(wsrc + ((size_t(tap) * planes_total + plane0 + pl) * p.CoutP + co0 + n) * 8);
                *reinterpret_cast<uint4 *>(dstW + (size_t(tap * PL + pl) * NT + n) * 16) = v;
            }
        }
        fence_proxy_async();
        __syncthreads();
        // ---- MMA issue -----------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            const uint32_t aaddr = smem_u32(sAbuf + buf * a_bytes);
            const uint32_t waddr = smem_u32(sWbuf + buf * w_bytes);
            const int ntap = is_skip ? 1 : p.taps;
            for (int mb = 0; mb < p.MB; ++mb) {
                const uint32_t d = tmem_base + uint32_t(mb * NT);
                for (int tap = 0; tap < ntap; ++tap) {
                    int shift;
                    if (is_skip) {
                        shift = p.pad * P + p.pad;  // centre tap
                    } else {
                        const int dy = p.taps == 9 ? tap / 3 : 0, dx = p.taps == 9 ? tap - dy * 3 : 0;
                        shift = dy * P + dx;
                    }
#pragma unroll
                    for (int k16 = 0; k16 < PL / 2; ++k16) {
                        const uint64_t ad = make_desc(aaddr + uint32_t((2 * k16) * WN + mb * 128 + shift) * 16, uint32_t(WN) * 16, 128);
                        const uint64_t bd = make_desc(waddr + uint32_t((tap * PL + 2 * k16) * NT) * 16, uint32_t(NT) * 16, 128);
                        umma_bf16(d, ad, bd, p.idesc, (kc > 0 || tap > 0 || k16 > 0) ? 1u : 0u);
                    }
                }
            }
            umma_commit(bars + buf);
            if (kc == n_chunks - 1) umma_commit(bars + 2);
        }
    }

    // ---- epilogue -------------------------------------------------------------------------------
    mbar_wait(bars + 2, 0);
    tc_fence_after();
    const int wq = warp & 3, half = warp >> 2;
    float *sT = reinterpret_cast<float *>(sAbuf) + warp * (32 * (CGW + 1));  // per-warp transpose scratch (A buffers are idle now)
    for (int cg = 0; cg < NT / CGW; ++cg) {
        float s1[CGW], s2[CGW];
#pragma unroll
        for (int i = 0; i < CGW; ++i) s1[i] = 0.f, s2[i] = 0.f;
        const int cobase = co0 + cg * CGW;
        for (int mb = half; mb < p.MB; mb += 2) {
            float v[CGW];
            tmem_ld16(tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(mb * NT + cg * CGW), v);
            const int j = mb * 128 + wq * 32 + lane;
            const int o = j / P, c = j - o * P;
            const int y = y0 + o, x = x0 + c;
            if (o < p.R && c < p.Wt && y < p.H && x < p.W) {
                const size_t pix = (size_t(b) * p.H + y) * p.W + x;
#pragma unroll
                for (int i = 0; i < CGW; ++i) v[i] += sAdd[cg * CGW + i];
                if (p.res != nullptr) {
                    const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + pix * p.Cout + cobase);
#pragma unroll
                    for (int h2 = 0; h2 < CGW / 8; ++h2) {
                        const uint4 rr = rp[h2];
                        const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float2 f = unpack_bf16(w4[i]);
                            v[h2 * 8 + 2 * i] += f.x;
                            v[h2 * 8 + 2 * i + 1] += f.y;
                        }
                    }
                }
                if (p.out_f32) {
                    float *op = reinterpret_cast<float *>(p.out) + pix * p.Cout + cobase;
#pragma unroll
                    for (int i = 0; i < CGW; ++i)
                        if (cobase + i < p.Cout) op[i] = v[i];
                } else {
                    uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + pix * p.Cout + cobase);
#pragma unroll
                    for (int h2 = 0; h2 < CGW / 8; ++h2) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            pk[i] = pack_bf16(v[h2 * 8 + 2 * i], v[h2 * 8 + 2 * i + 1]);
                            float2 f = unpack_bf16(pk[i]);  // statistics of the values as stored
                            v[h2 * 8 + 2 * i] = f.x;
                            v[h2 * 8 + 2 * i + 1] = f.y;
                        }
                        op[h2] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
#pragma unroll
                for (int i = 0; i < CGW; ++i) {
                    s1[i] += v[i];
                    s2[i] = fmaf(v[i], v[i], s2[i]);
                }
            }
        }
        if (p.ostat != nullptr) {
            // deterministic cross-lane reduction through a padded smem transpose
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < CGW; ++i) sT[lane * (CGW + 1) + i] = w ? s2[i] : s1[i];
                __syncwarp();
                if (lane < CGW) {
                    float s = 0.f;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) s += sT[r * (CGW + 1) + lane];
                    sRed[(warp * NT + cg * CGW + lane) * 2 + w] = s;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, uint32_t(p.tmem_cols));
    if (p.ostat == nullptr) return;

    const int n_tiles = gridDim.x;
    if (tid < NT * 2) {
        const int c = tid >> 1, w = tid & 1;
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += sRed[(r * NT + c) * 2 + w];
        p.part[((size_t(b) * n_tiles + tile) * p.CoutP + co0 + c) * 2 + w] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        const unsigned int prev = atomicAdd(p.ticket + b, 1u);
        *s_last = (prev == total - 1);
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    for (int e = tid; e < p.Cout * 2; e += TC_THREADS) {
        const int c = e >> 1, w = e & 1;
        double s = 0.0;
        for (int t = 0; t < n_tiles; ++t) s += double(__ldcg(p.part + ((size_t(b) * n_tiles + t) * p.CoutP + c) * 2 + w));
        p.ostat[(size_t(b) * p.Cout + c) * 2 + w] = s;
    }
    if (tid == 0) p.ticket[b] = 0u;
}

struct TcCfg {
    int PL, R, Wt, P, MB, WN, NT, nchunks, tmem_cols;
    size_t smem;
};

constexpr size_t kScratchBytes = size_t(8) * 32 * (CGW + 1) * 4;  // per-warp transpose tiles of the epilogue

size_t tc_a_region(int PL, int WN) {
    size_t a = size_t(2) * PL * WN * 16;
    return a < kScratchBytes ? kScratchBytes : a;
}

size_t tc_smem_bytes(int PL, int WN, int taps, int NT, int Cin) {
    return tc_a_region(PL, WN) + size_t(2) * taps * PL * NT * 16 + sizeof(float) * (2 * size_t(Cin) + NT + 16 * size_t(NT)) + 64;
}

// Tile selection: the largest R x Wt tile whose accumulators fit 256 TMEM columns and whose
// buffers fit ~100 KB, so two CTAs share an SM and one's epilogue overlaps the other's loads.
bool tc_configure(int H, int W, int C0, int C1, int S0, int S1, int Cout, int ksize, TcCfg &c) {
    const int Cin = C0 + C1, Sk = S0 + S1;
    const int pad = ksize / 2, taps = ksize * ksize;
    const int CoutP = (Cout + 15) / 16 * 16;
    int n = 1;
    while (CoutP / n > 128 || CoutP % (16 * n)) {
        if (++n > 8) return false;
    }
    c.nchunks = n;
    c.NT = CoutP / n;
    const bool all32 = !(C0 % 32) && !(C1 % 32) && !(S0 % 32) && !(S1 % 32);
    c.PL = (all32 && c.NT <= 32) ? 4 : 2;
    if ((C0 % 16) || (C1 % 16) || (S0 % 16) || (S1 % 16) || Cin == 0) return false;
    (void)Sk;
    c.Wt = W > 64 ? 64 : W;
    c.P = c.Wt + 2 * pad;
    const size_t budget = 112 * 1024;  // two CTAs per SM (228 KB - 1 KB reserved each)
    for (int pl = c.PL; pl >= 2; pl -= 2) {
        for (int R = H; R >= 1; --R) {
            const int MB = (R * c.P + 127) / 128;
            if (MB * c.NT > 256) continue;
            int WN = MB * 128 + 2 * pad * c.P + 2 * pad;
            WN += (10 - (WN & 7)) & 7;  // WN == 2 (mod 8): conflict-free 16-byte stores across the planes
            const size_t smem = tc_smem_bytes(pl, WN, taps, c.NT, Cin);
            if (smem > budget) continue;
            c.PL = pl; c.R = R; c.MB = MB; c.WN = WN; c.smem = smem;
            int cols = 32;
            while (cols < MB * c.NT) cols *= 2;
            c.tmem_cols = cols;
            return true;
        }
    }
    return false;
}

}  // namespace

bool conv_tc_supported(const ccdm_op &op) {
    if (op.dtype != CCDM_DT_BF16 || op.src_kind != 0 || op.stride != 1) return false;
    if (op.ksize != 1 && op.ksize != 3) return false;
    if ((op.C0 % 16) || (op.C1 % 16) || (op.S0 % 16) || (op.S1 % 16)) return false;
    if (op.out_dtype == CCDM_DT_BF16 && (op.Cout % 16)) return false;
    TcCfg c;
    return tc_configure(op.Hout, op.Wout, op.C0, op.C1, op.S0, op.S1, op.Cout, op.ksize, c);
}

size_t conv_tc_part_floats(const ccdm_op &op) {
    TcCfg c;
    if (!tc_configure(op.Hout, op.Wout, op.C0, op.C1, op.S0, op.S1, op.Cout, op.ksize, c)) return 0;
    const int tiles = ((op.Hout + c.R - 1) / c.R) * ((op.Wout + c.Wt - 1) / c.Wt);
    return size_t(op.B) * tiles * ((op.Cout + 15) / 16 * 16) * 2;
}

int launch_conv_tc(const ccdm_op &op, cudaStream_t s) {
    TcCfg c;
    if (!conv_tc_supported(op) || !tc_configure(op.Hout, op.Wout, op.C0, op.C1, op.S0, op.S1, op.Cout, op.ksize, c))
        CCDM_FAIL(-3, "conv_tc: unsupported configuration");
    TcP p{};
    p.src0 = (const __nv_bfloat16 *)op.src0; p.src1 = (const __nv_bfloat16 *)op.src1;
    p.stat0 = (const double *)op.stat0; p.stat1 = (const double *)op.stat1;
    p.gamma = (const float *)op.gamma; p.beta = (const float *)op.beta;
    p.weight = (const __nv_bfloat16 *)op.weight; p.bias = (const float *)op.bias; p.emb = (const float *)op.emb;
    p.skip0 = (const __nv_bfloat16 *)op.skip0; p.skip1 = (const __nv_bfloat16 *)op.skip1;
    p.skip_w = (const __nv_bfloat16 *)op.skip_w; p.res = (const __nv_bfloat16 *)op.res;
    p.out = (void *)op.out; p.ostat = (double *)op.ostat; p.part = (float *)op.part; p.ticket = (unsigned int *)op.ticket;
    p.steps = (const ccdm_step_entry *)op.steps; p.step_ptr = (const int *)op.step_ptr;
    p.B = op.B; p.Hin = op.Hin; p.Win = op.Win; p.H = op.Hout; p.W = op.Wout;
    p.C0 = op.C0; p.C1 = op.C1; p.Cin = op.C0 + op.C1; p.Cout = op.Cout; p.CoutP = (op.Cout + 15) / 16 * 16; p.NT = c.NT;
    p.upsample = op.upsample; p.gn = op.gn; p.silu = op.silu; p.S0 = op.S0; p.S1 = op.S1;
    p.emb_off = op.emb_off; p.emb_cols = op.emb_cols; p.emb_bstride = op.emb_bstride;
    p.out_f32 = op.out_dtype == CCDM_DT_F32;
    p.R = c.R; p.Wt = c.Wt; p.P = c.P; p.MB = c.MB; p.WN = c.WN;
    p.tiles_x = (op.Wout + c.Wt - 1) / c.Wt; p.tiles_y = (op.Hout + c.R - 1) / c.R;
    p.taps = op.ksize * op.ksize; p.pad = op.ksize / 2;
    p.tmem_cols = c.tmem_cols;
    p.a_region = uint32_t(tc_a_region(c.PL, c.WN));
    // cute::UMMA::InstrDescriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 at 17, M>>4 at 24
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(c.NT >> 3) << 17) | (uint32_t(128 >> 4) << 24);

    if (op.gn && (!op.stat0 || (op.C1 && !op.stat1) || !op.gamma || !op.beta)) CCDM_FAIL(-2, "conv_tc: gn without stats/affine");
    if (op.gn && (p.Cin % kGnGroups)) CCDM_FAIL(-2, "conv_tc: GroupNorm needs Cin %% 32 == 0");
    if (op.ostat && (!op.part || !op.ticket)) CCDM_FAIL(-2, "conv_tc: ostat without scratch");
    if (op.emb && (!op.steps || !op.step_ptr || op.emb_off < 0)) CCDM_FAIL(-2, "conv_tc: emb without step table");
    if (op.S0 > 0 && (!op.skip0 || !op.skip_w)) CCDM_FAIL(-2, "conv_tc: bad skip configuration");
    {
        const int expH = op.upsample ? op.Hin * 2 : op.Hin, expW = op.upsample ? op.Win * 2 : op.Win;
        if (expH != op.Hout || expW != op.Wout) CCDM_FAIL(-2, "conv_tc: inconsistent shapes");
    }
    dim3 grid(p.tiles_x * p.tiles_y, c.nchunks, op.B);
    auto kern = c.PL == 4 ? conv_tc_kernel<4> : conv_tc_kernel<2>;
    CCDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(c.smem)));
    kern<<<grid, TC_THREADS, c.smem, s>>>(p);
    CCDM_LAUNCH_CHECK("conv_tc_kernel");
    return 0;
}

}  // namespace ccdm
