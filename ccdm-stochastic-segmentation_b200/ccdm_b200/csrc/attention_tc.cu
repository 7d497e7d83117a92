// QKVAttentionLegacy (reference unet.py:343-360) on the 5th-generation tensor cores, for bf16 activations
// in the plane-major layout and head_dim = 32 (every shipped configuration: num_head_channels = 32) or 64.
//
//   softmax((q*s)(k*s)^T) v,  s = 32^-1/4, per (sample, head), never materialising the [T,T] scores.
//
// One CTA = one (sample, head, tile of 128 queries), 128 threads; key/value tiles of NK keys stream through
// shared memory (double buffered, cp.async.bulk + mbarrier).  Because qkv is stored plane-major,
// [B][3C/8][T][8], a tile of q, k or v is four contiguous runs (one per 8-channel plane) and lands in
// shared memory ALREADY in a tcgen05 operand layout -- no transpose, no staging through registers:
//     q, k : "K-major, no swizzle"  (rows = tokens, 16-byte rows of 8 channels; LBO = plane stride)
//     v    : "MN-major, no swizzle" (N = channels contiguous inside a 16-byte row, K = keys 16 B apart,
//                                    LBO = 128 B between groups of 8 keys, SBO = plane stride)
// Per key tile:  S = Q K^T (tcgen05.mma, fp32 in TMEM)  ->  every thread owns one query row: tcgen05.ld,
// tile max / running sum, P = exp2((S - m) * log2(e)/sqrt(32)) as bf16 into shared memory (K-major A operand)
// ->  O += P V (tcgen05.mma accumulating in TMEM over ALL tiles; m is a reference maximum that may lag the row's
// true maximum by 2^8, a row rescales l and its TMEM lane only when a tile exceeds it by more -- see the main loop).
// Several CTAs are resident per SM (TMEM: 256 of 512 columns each at NK = 128), so one CTA's softmax overlaps
// another's MMAs and loads.
//
// Output: [B][C/8][T][8] bf16, channel(h, d) = h*32 + d (unet.py:360).
//
// (Measured and rejected: software-pipelining S(t+1) under softmax(t) with a double-buffered S accumulator --
// T=256 22.9 vs 15.0 us, T=2048 75.9 vs 66 us.  The kernel is not waiting for the tensor core; with 128 threads per
// CTA and two CTAs per SM it is short of warps for the softmax arithmetic.
// Also measured and rejected: two threads per query row (256 threads; the warps w and w+4 of a TMEM lane quarter split the
// key columns and the output channels, one extra CTA barrier per key tile to agree on the running maximum) -- correct,
// but T=2048 55.7 vs 54.4 us, T=512 14.6 vs 15.7, T=256 19.7 vs 14.7: the barrier costs what the halved per-thread work
// saves.  What bounds a CTA is the chain of round trips per key tile (MMA commit -> mbarrier -> tcgen05.ld -> barrier ->
// MMA commit -> ...), so the next step is more independent work per SM: split the key range over two CTAs when
// B x heads x query tiles < 2 x 148 (Cityscapes B=8), or a second query tile per CTA.)
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

namespace ccdm {
namespace {

// head_dim D: 32 (every shipped UNet configuration) or 64 (ViT-S heads, the DINO condition encoder)
constexpr int AT_QT = 128;     // queries per CTA = threads = MMA M

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
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

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct AtP {
    const __nv_bfloat16 *qkv;  // [B][3C/8][T][8]
    __nv_bfloat16 *out;        // [B][C/8][T][8]
    int T, heads, q_tiles;
    float scale_log2;          // log2(e) / sqrt(head_dim)
    uint32_t idesc_s, idesc_pv;
};

// X3: fp16x2 operands (CCDM_DT_F16X2; the "exact" tensor-core mode).  Every plane of q, k, v and P exists twice (hi, lo,
// adjacent), both products are three fp16 MMAs (hi*lo + lo*hi + hi*hi), the softmax is the same fp32 arithmetic, and P
// is split into hi + lo before the P V product -- fp32-grade attention on the tensor cores.
template <int NK, bool X3, int D = 32>
__global__ void __launch_bounds__(AT_QT, D > 32 ? (X3 ? 1 : 2) : 3) attention_tc_kernel(const AtP p) {  // (head_dim 32: 72 KB of shared memory per CTA in fp16x2, three per SM; 168 registers, 16 bytes spilled)
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int X = X3 ? 2 : 1;
    constexpr int AT_D = D, AT_PLANES = D / 8;
    constexpr uint32_t Q_BYTES = X * AT_PLANES * AT_QT * 16;  // 8 KB
    constexpr uint32_t KV_BYTES = X * AT_PLANES * NK * 16;    // per tensor per buffer
    constexpr uint32_t P_BYTES = X * (NK / 8) * AT_QT * 16;
    uint8_t *sQ = smem;
    uint8_t *sK = sQ + Q_BYTES;        // [2][KV_BYTES]
    uint8_t *sV = sK + 2 * KV_BYTES;   // [KV_BYTES]: ONE buffer -- V(t) is needed only after the softmax of tile t, so it is fetched
                                       // at the top of iteration t (when P V(t-1) has completed) and its latency hides behind the
                                       // softmax; the 8 KB saved make a third CTA fit an SM at head_dim 32 in fp16x2
    uint8_t *sP = sV + KV_BYTES;       // [NK/8 planes][128 rows][16 B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + P_BYTES);
    uint64_t *kv_full = bars, *s_done = bars + 2, *o_done = bars + 3, *v_full = bars + 4;  // kv_full: K tiles (and Q with the first)
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 5);
    constexpr uint32_t TMEM_COLS = NK + D <= 64 ? 64 : (NK + D <= 128 ? 128 : 256);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int qt = blockIdx.x % p.q_tiles, bh = blockIdx.x / p.q_tiles;
    const int h = bh % p.heads, b = bh / p.heads;
    const int T = p.T, q0 = qt * AT_QT;
    const int nq = min(AT_QT, T - q0);
    const int n_tiles = (T + NK - 1) / NK;
    const int planes3 = p.heads * 3 * AT_PLANES;  // planes of the qkv tensor
    // plane g of (which = 0 q | 1 k | 2 v) of this head: channel h*3D + which*D + 8g
    // (fp16x2: tensor plane 2*that + part, part = 0 hi | 1 lo; g then counts (group, part) pairs)
    auto plane_ptr = [&](int which, int g) { return p.qkv + ((size_t(b) * planes3 * X + (h * 3 * AT_PLANES + which * AT_PLANES) * X + g) * T) * 8; };

    if (warp == 0) tmem_alloc(s_tmem, TMEM_COLS);
    if (tid == 0) {
        mbar_init(kv_full + 0, 1);
        mbar_init(kv_full + 1, 1);
        mbar_init(s_done, 1);
        mbar_init(o_done, 1);
        mbar_init(v_full, 1);
        fence_barrier_init();
    }
    // rows of a partial tile that no copy overwrites must hold finite values (0 * NaN would poison P V)
    for (uint32_t i = tid * 16; i < Q_BYTES + 3 * KV_BYTES; i += AT_QT * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_s = *s_tmem, tmem_o = tmem_s + NK;
    pdl_launch_dependents();
    pdl_wait();  // qkv is written by the previous kernel of the step
    const uint32_t trow = uint32_t(warp * 32) << 16;  // this warp's TMEM lane quarter
    auto ld_o = [&](float *dst) {
#pragma unroll
        for (int d0 = 0; d0 < AT_D; d0 += 32) tmem_ld32(tmem_o + trow + uint32_t(d0), dst + d0);
    };
    auto st_o = [&](const float *src) {
#pragma unroll
        for (int d0 = 0; d0 < AT_D; d0 += 32) tmem_st32(tmem_o + trow + uint32_t(d0), src + d0);
    };

    auto load_k = [&](int t, int buf, uint32_t extra_bytes) {  // thread 0
        const int k0 = t * NK, nk = min(NK, T - k0);
        mbar_expect_tx(kv_full + buf, uint32_t(X * AT_PLANES * nk * 16) + extra_bytes);
#pragma unroll
        for (int g = 0; g < X * AT_PLANES; ++g) bulk_g2s(sK + buf * KV_BYTES + g * NK * 16, plane_ptr(1, g) + size_t(k0) * 8, uint32_t(nk * 16), kv_full + buf);
    };
    auto load_v = [&](int t) {  // thread 0; the single V buffer must be free: P V(t-1) has completed
        const int k0 = t * NK, nk = min(NK, T - k0);
        mbar_expect_tx(v_full, uint32_t(X * AT_PLANES * nk * 16));
#pragma unroll
        for (int g = 0; g < X * AT_PLANES; ++g) bulk_g2s(sV + g * NK * 16, plane_ptr(2, g) + size_t(k0) * 8, uint32_t(nk * 16), v_full);
    };
    if (tid == 0) {
        load_k(0, 0, uint32_t(X * AT_PLANES * nq * 16));
        load_v(0);
#pragma unroll
        for (int g = 0; g < X * AT_PLANES; ++g) bulk_g2s(sQ + g * AT_QT * 16, plane_ptr(0, g) + size_t(q0) * 8, uint32_t(nq * 16), kv_full + 0);
    }

    const uint32_t desc_hi = 8u | (1u << 14);                                   // SBO 128 B, version 1
    const uint32_t q_lo = (smem_u32(sQ) >> 4) | (uint32_t(X * AT_QT) << 16);    // LBO = stride between 8-channel groups (X planes of 128 rows)
    const uint32_t p_lo = (smem_u32(sP) >> 4) | (uint32_t(X * AT_QT) << 16);
    const uint32_t v_hi = uint32_t(X * NK) | (1u << 14);                        // MN-major: SBO = stride between 8-channel groups (X planes of NK rows)
    // one product = one bf16 MMA, or hi*lo + lo*hi + hi*hi on the split operands; (a_lo, b_lo) = row offsets of the lo planes
    auto mma = [&](uint32_t d, uint32_t a, uint32_t a_hi, uint32_t a_lo, uint32_t b, uint32_t b_hi, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
        if constexpr (X3) {
            umma_bf16(d, (uint64_t(a_hi) << 32) | a, (uint64_t(b_hi) << 32) | (b + b_lo), idesc, acc);
            umma_bf16(d, (uint64_t(a_hi) << 32) | (a + a_lo), (uint64_t(b_hi) << 32) | b, idesc, 1u);
            umma_bf16(d, (uint64_t(a_hi) << 32) | a, (uint64_t(b_hi) << 32) | b, idesc, 1u);
        } else {
            umma_bf16(d, (uint64_t(a_hi) << 32) | a, (uint64_t(b_hi) << 32) | b, idesc, acc);
        }
    };

    // O accumulates in TMEM over all key tiles (the P V MMAs of tile t add to it) -- it is read back once, at the end.  The
    // exponentials are taken against a REFERENCE maximum m that may lag the true running maximum of the row by up to LAG
    // (log2 units: p <= 2^LAG, far inside bf16 / fp16 range; softmax is shift-invariant, so nothing else changes); only
    // when a tile exceeds it by more does the row rescale what it has accumulated (l, and its O lane through
    // tcgen05.ld / st) -- after the first tile that is rare.  So per tile a thread waits ONCE (for S), and the issuing
    // thread queues P V(t) and S(t+1) back to back.  (Before: o = o * corr + O_tile in registers every tile -- a second
    // commit -> mbarrier -> tcgen05.ld round trip and 32 FMAs per tile.)
    // fp16x2: the tensor core truncates its accumulator on every accumulating MMA (tests/test_gpu_exact.py), so an O that
    // collects all 12 x n_tiles MMAs in TMEM drifts (8e-6 of the output's scale at T = 2048); every FLUSH tiles the row moves
    // what TMEM holds into fp32 registers (round to nearest) and the next P V starts the accumulator afresh.
    constexpr float LAG = 8.0f;
    constexpr int FLUSH = 8;
    float m = -INFINITY, l = 0.f;
    const float c = p.scale_log2;
    float o[AT_D];
    if constexpr (X3) {
#pragma unroll
        for (int d = 0; d < AT_D; ++d) o[d] = 0.f;
    }

    auto issue_s = [&](int t) {  // thread 0: S(t) = Q K(t)^T into the S columns
        const int buf = t & 1;
        mbar_wait(kv_full + buf, uint32_t(t >> 1) & 1u);
        tc_fence_after();
        const uint32_t k_lo = (smem_u32(sK + buf * KV_BYTES) >> 4) | (uint32_t(X * NK) << 16);
#pragma unroll
        for (int j = 0; j < AT_D / 16; ++j)
            mma(tmem_s, q_lo + uint32_t(j * 2 * X * AT_QT), desc_hi, uint32_t(AT_QT), k_lo + uint32_t(j * 2 * X * NK), desc_hi, uint32_t(NK), p.idesc_s,
                j > 0 ? 1u : 0u);
        umma_commit(s_done);
    };
    if (tid == 0) {
        if (n_tiles > 1) load_k(1, 1, 0u);
        issue_s(0);
    }

    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        const int nk = min(NK, T - t * NK);
        mbar_wait(s_done, uint32_t(t) & 1u);  // S(t) is complete -- and with it every MMA issued before it, P V(t-1) included
        tc_fence_after();
        if (tid == 0 && t >= 1) {
            load_v(t);                                      // P V(t-1), the buffer's last reader, has completed
            if (t + 1 < n_tiles) load_k(t + 1, buf ^ 1, 0u);  // likewise S(t-1), the last reader of that K buffer
        }
        const bool flushed = X3 && t > 0 && (t % FLUSH) == 0;  // (uniform) O is quiescent here: P V(t-1) has completed, P V(t) is not issued yet
        if (flushed) {
            float ot[AT_D];
            ld_o(ot);
#pragma unroll
            for (int d = 0; d < AT_D; ++d) o[d] += ot[d];
        }
        // one tile of the online softmax; FULL: all NK keys are real (every tile but possibly the last) -- no per-element predicates
        auto softmax_tile = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            // pass 1: row maximum of the tile
            // (four independent partial maxima / sums: a single running value would be a 128-long dependent chain per
            // tile, and with two CTAs of four warps per SM there is nothing to hide its latency behind)
            float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c0 = 0; c0 < NK; c0 += 32) {
                if (FULL || c0 < nk) {
                    float s[32];
                    tmem_ld32(tmem_s + trow + uint32_t(c0), s);
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (FULL || c0 + i < nk) tm[i & 3] = fmaxf(tm[i & 3], s[i]);
                }
            }
            const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
            const bool grow = (tmax - m) * c > LAG;  // always on the first tile (m = -inf)
            if (__any_sync(0xffffffffu, grow)) {     // warp-uniform: tcgen05.ld / st are warp-wide instructions
                const float m_new = grow ? tmax : m;
                if (t > 0) {
                    const float corr = ex2_approx((m - m_new) * c);  // 1 on the lanes that keep their maximum
                    if (!flushed) {  // (after a flush TMEM holds nothing that counts: the next P V overwrites it)
                        float ot[AT_D];
                        ld_o(ot);
#pragma unroll
                        for (int d = 0; d < AT_D; ++d) ot[d] *= corr;
                        st_o(ot);
                    }
                    if constexpr (X3) {
#pragma unroll
                        for (int d = 0; d < AT_D; ++d) o[d] *= corr;
                    }
                    l *= corr;
                }
                m = m_new;
            }
            const float mc = m * c;
            float rs[4] = {0.f, 0.f, 0.f, 0.f};
            // pass 2: P = exp2(S*c - m*c) as bf16 (fp16x2: hi + lo), K-major rows for the P V product
#pragma unroll
            for (int c0 = 0; c0 < NK; c0 += 32) {
                uint32_t pk[16], pl[16];
                if (FULL || c0 < nk) {
                    float s[32];
                    tmem_ld32(tmem_s + trow + uint32_t(c0), s);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = (FULL || c0 + i < nk) ? ex2_approx(fmaf(s[i], c, -mc)) : 0.f;
                        float p1 = (FULL || c0 + i + 1 < nk) ? ex2_approx(fmaf(s[i + 1], c, -mc)) : 0.f;
                        if (X3) {
                            split_f16x2(p0, p1, pk[i / 2], pl[i / 2]);  // hi + lo reproduces p to 2^-23: the sum the MMA sees is the fp32 sum
                            rs[(i >> 1) & 3] += p0 + p1;
                        } else {
                            pk[i / 2] = pack_bf16(p0, p1);
                            const float2 f = unpack_bf16(pk[i / 2]);  // the sum of what the MMA will actually see
                            rs[(i >> 1) & 3] += f.x + f.y;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u, pl[i] = 0u;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {  // key planes c0/8 .. c0/8+3, this thread's row
                    *reinterpret_cast<uint4 *>(sP + (size_t((c0 / 8 + g) * X) * AT_QT + tid) * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                    if (X3)
                        *reinterpret_cast<uint4 *>(sP + (size_t((c0 / 8 + g) * X + 1) * AT_QT + tid) * 16) = make_uint4(pl[4 * g], pl[4 * g + 1], pl[4 * g + 2], pl[4 * g + 3]);
                }
            }
            l += (rs[0] + rs[1]) + (rs[2] + rs[3]);
        };
        if (nk == NK) softmax_tile(std::true_type{});
        else softmax_tile(std::false_type{});
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();  // P (and a rescaled O) complete and visible to the tensor core; every thread is done reading S
        if (tid == 0) {
            mbar_wait(v_full, uint32_t(t) & 1u);
            tc_fence_after();
            const uint32_t v_lo = (smem_u32(sV) >> 4) | (8u << 16);  // LBO = 128 B between groups of 8 keys
#pragma unroll
            for (int j = 0; j < NK / 16; ++j)
                mma(tmem_o, p_lo + uint32_t(j * 2 * X * AT_QT), desc_hi, uint32_t(AT_QT), v_lo + uint32_t(j * 16), v_hi, uint32_t(NK), p.idesc_pv,
                    ((t > 0 && !flushed) || j > 0) ? 1u : 0u);
            // S(t+1) right behind it: it overwrites S, which every thread finished reading before the barrier above; P is
            // rewritten only after S(t+1) -- and so P V(t) -- has completed
            if (t + 1 < n_tiles) issue_s(t + 1);
            else umma_commit(o_done);
        }
    }
    mbar_wait(o_done, 0u);
    tc_fence_after();
    {
        float ot[AT_D];
        ld_o(ot);
#pragma unroll
        for (int d = 0; d < AT_D; ++d) o[d] = X3 ? o[d] + ot[d] : ot[d];
    }

    if (tid < nq) {
        // fp16x2: o = sum (16 p)(16 v): the stored 16 x value is o / (16 l)
        const float inv = X3 ? 1.0f / (16.0f * l) : 1.0f / l;
        const int planes = p.heads * AT_PLANES;
#pragma unroll
        for (int g = 0; g < AT_PLANES; ++g) {
            uint32_t pk[4], pl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (X3) split_f16x2_raw(o[8 * g + 2 * i] * inv, o[8 * g + 2 * i + 1] * inv, pk[i], pl[i]);
                else pk[i] = pack_bf16(o[8 * g + 2 * i] * inv, o[8 * g + 2 * i + 1] * inv);
            }
            __nv_bfloat16 *dst = p.out + ((size_t(b) * planes + h * AT_PLANES + g) * X * T + q0 + tid) * 8;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (X3) *reinterpret_cast<uint4 *>(dst + size_t(T) * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_s, TMEM_COLS);
    }
}

template <int NK, bool X3, int D = 32>
int launch_nk(const AtP &p, int grid, cudaStream_t s) {
    constexpr int AT_PLANES = D / 8;
    constexpr size_t smem = (X3 ? 2 : 1) * (AT_PLANES * AT_QT * 16 + 3 * AT_PLANES * NK * 16 + (NK / 8) * AT_QT * 16) + 64;
    static bool attr_done = false;
    if (!attr_done) {
        CCDM_CUDA(cudaFuncSetAttribute(attention_tc_kernel<NK, X3, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        attr_done = true;
    }
    CCDM_CUDA(launch_pdl(attention_tc_kernel<NK, X3, D>, dim3(grid), dim3(AT_QT), smem, s, p));
    CCDM_LAUNCH_CHECK("attention_tc_kernel");
    return 0;
}

}  // namespace

bool attention_tc_supported(const ccdm_op &op) {
    return (op.dtype == CCDM_DT_BF16 || op.dtype == CCDM_DT_F16X2) && (op.head_dim == 32 || op.head_dim == 64) && !op.exact && op.heads > 0;
}

int launch_attention_tc(const ccdm_op &op, cudaStream_t s) {
    const int T = op.Hin * op.Win;
    const int AT_D = op.head_dim;
    if (op.C0 != op.heads * AT_D * 3) CCDM_FAIL(-2, "attention_tc: qkv channels %d != 3*heads*%d", op.C0, AT_D);
    AtP p{};
    p.qkv = (const __nv_bfloat16 *)op.src0;
    p.out = (__nv_bfloat16 *)op.out;
    p.T = T;
    p.heads = op.heads;
    p.q_tiles = (T + AT_QT - 1) / AT_QT;
    const bool x3 = op.dtype == CCDM_DT_F16X2;
    // unet.py:354: q and k are each scaled by 32^-1/4; fp16x2 operands are stored as 16 x value: S comes out 256 x too large
    p.scale_log2 = float(1.4426950408889634 / sqrt(double(AT_D))) * (x3 ? 1.0f / 256.0f : 1.0f);
    static const int env_nk = getenv("CCDM_ATT_NK") ? atoi(getenv("CCDM_ATT_NK")) : 0;  // tuning override
    // (head_dim 64 and fp16x2 are instantiated with 64-key tiles only: the S descriptor must describe the tile the kernel streams)
    const int NK = (x3 || AT_D == 64) ? 64 : env_nk == 64 || env_nk == 128 ? env_nk : (T <= 256 ? 64 : 128);  // measured: T=256 16.8 vs 20.9 us, T=2048 76 vs 66 us
    // cute::UMMA::InstrDescriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), b_major = MN (bit 16), N>>3 at 17, M>>4 at 24
    const uint32_t base = (1u << 4) | (x3 ? 0u : ((1u << 7) | (1u << 10))) | (uint32_t(128 >> 4) << 24);  // formats: 0 = f16, 1 = bf16
    p.idesc_s = base | (uint32_t(NK >> 3) << 17);
    p.idesc_pv = base | (uint32_t(AT_D >> 3) << 17) | (1u << 16);
    const int grid = op.B * op.heads * p.q_tiles;
    if (!p.qkv || !p.out || T <= 0 || grid <= 0) CCDM_FAIL(-2, "attention_tc: missing tensors");
    if (AT_D == 64) return x3 ? launch_nk<64, true, 64>(p, grid, s) : launch_nk<64, false, 64>(p, grid, s);  // one CTA per SM (128 / 72 KB)
    if (x3) return launch_nk<64, true>(p, grid, s);  // 72 KB of shared memory per CTA at NK = 64: three CTAs per SM
    return NK == 64 ? launch_nk<64, false>(p, grid, s) : launch_nk<128, false>(p, grid, s);
}

}  // namespace ccdm
