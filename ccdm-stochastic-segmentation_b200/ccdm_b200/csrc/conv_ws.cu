// Fused GroupNorm + SiLU + conv (3x3 / 1x1, stride 1, optional nearest-x2 in front) on the
// 5th-generation tensor cores, as a PERSISTENT, WARP-SPECIALISED kernel: bf16 activations and
// weights, fp32 accumulation in TMEM (tcgen05.mma, cta_group::1, M = 128), fused bias + timestep
// embedding + identity residual or 1x1 skip conv, GroupNorm statistics of the output.
//
// Replaces the same reference lines as conv_ffma.cu (unet.py:242-262 ResBlock, :106-116 Upsample,
// :305-311 attention qkv / proj_out, :701-705 output conv) for the bf16 ("fast") precision mode.
//
// Implicit GEMM without im2col -- the "flattened padded tile":
//   A work item is a tile of R x Wt output pixels of one sample (x one chunk of NT output
//   channels).  Its input window, (R+2) x (Wt+2) pixels for a 3x3, is staged in shared memory as
//   channel planes of 8 channels:
//       plane[g][q] = 16 bytes = channels 8g..8g+7 of window position q = r*P + c,  P = Wt + 2.
//   This is the UMMA "K-major, no swizzle" canonical layout with SBO = 128 B (8 rows of 16 B) and
//   LBO = plane stride: MMA row i reads position (start + i), so output position j = o*P + c
//   under tap (dy,dx) reads window position j + dy*P + dx -- a pure shift of the descriptor start
//   address.  One M=128 MMA covers 128 consecutive flattened output positions; the two pad
//   columns per row produce don't-care rows that the epilogue skips.  No im2col, no duplication.
//
// Roles (448 threads, one CTA per SM, each CTA walks a contiguous range of work items):
//   warps 0-3   epilogue : tcgen05.ld accumulators -> +bias +embedding +residual -> bf16 NHWC
//                          store + per-channel statistics (warp-shuffle transpose-reduce);
//                          double-buffered accumulators let it overlap the next item's MMAs.
//   warps 4-11  producers: LDG.128 (two batches in flight) -> GroupNorm affine + SiLU in fp32
//                          -> bf16 -> STS.128 into the stage ring, K chunks of 16/32 channels.
//   warp 12     MMA      : one thread issues tcgen05.mma for every (M block, tap, K16) of a stage
//                          and commits to the stage's "empty" mbarrier.
//   warp 13     weights  : cp.async.bulk (TMA, 1-D) of the packed weights -- once per CTA when
//                          they fit (resident), else one chunk per stage.
// All hand-offs are mbarriers; nothing in the main loop is a CTA-wide barrier.
#include "common.cuh"
#include "conv_tc_common.cuh"

namespace ccdm {
namespace {

constexpr int PROD_WARPS = 8;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int WARP_MMA = EPI_WARPS + PROD_WARPS, WARP_LOAD = WARP_MMA + 1;
constexpr int WS_THREADS = (WARP_LOAD + 1) * 32;  // 448
constexpr int UB = 4;    // loads per producer batch (two batches in flight)

// ---- the kernel -------------------------------------------------------------------------------
// PL = planes (8-channel groups) per K chunk: KC = 8*PL channels per pipeline stage.
template <int PL>
__global__ void __launch_bounds__(WS_THREADS, 1) conv_ws_kernel(const __grid_constant__ WsP p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int KC = 8 * PL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = p.NT, WN = p.WN, P = p.P, NS = p.NS;

    // smem carve-up
    uint8_t *sA = smem_raw;                                             // [NS][PL][WN][16B]
    uint8_t *sW = sA + size_t(NS) * p.a_stage;                          // resident: main ++ skip; else [NS][w_stage]
    const size_t w_region = p.resident ? size_t(p.w_main_bytes) + p.w_skip_bytes : size_t(NS) * p.w_stage;
    float *sAff = reinterpret_cast<float *>(sW + w_region);             // [2][Cin] GN scale / shift (producers)
    float *sAdd = sAff + 2 * p.Cin;                                     // [NT] bias (+ embedding) (epilogue)
    float *sRed = sAdd + NT;                                            // [4 warps][CoutP][2] running statistics
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRed + EPI_WARPS * p.CoutP * 2);
    uint64_t *full_a = bars, *full_w = bars + MAX_STAGES, *empty = bars + 2 * MAX_STAGES;
    uint64_t *acc_full = bars + 3 * MAX_STAGES, *acc_empty = acc_full + 2, *w_res = acc_empty + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(w_res + 1);
    int *s_last = reinterpret_cast<int *>(s_tmem + 1);
    double2 *sSum = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(s_last + 1) + 15) & ~uintptr_t(15));  // [Cin] GroupNorm fold scratch

    if (warp == WARP_MMA) tmem_alloc(s_tmem, uint32_t(p.tmem_cols));
    if (tid == WARP_LOAD * 32) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full_a + s, PROD_WARPS);
            mbar_init(full_w + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_full + 0, 1);
        mbar_init(acc_full + 1, 1);
        mbar_init(acc_empty + 0, EPI_WARPS);
        mbar_init(acc_empty + 1, EPI_WARPS);
        mbar_init(w_res, 1);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_launch_dependents();
    pdl_wait();  // everything below reads or writes tensors of the step

    // contiguous item range of this CTA (vertically adjacent tiles stay on one SM: halo rows hit L2)
    const int it_begin = int((long long)blockIdx.x * p.n_items / gridDim.x);
    const int it_end = int((long long)(blockIdx.x + 1) * p.n_items / gridDim.x);
    const int n_chunks = p.n_main + p.n_skip;

    if (warp < EPI_WARPS) {
        conv_epilogue_role<EPI_WARPS>(p, sAdd, sRed, s_last, acc_full, acc_empty, tmem_base, it_begin, it_end);
    } else if (warp < WARP_MMA) {
        // =========================== producers ==================================================
        const int pt = tid - EPI_THREADS;
        const int g = pt & (PL - 1);  // this thread's channel plane inside a chunk
        const int qf = pt / PL;
        constexpr int QS = PROD_THREADS / PL;
        const int nq = (p.R + 2 * p.pad) * P;
        int stage = 0, cur_b = -1;
        uint32_t phase = 0;
        for (int it = it_begin; it < it_end; ++it) {
            const Item I = decode_item(p, it);
            if (p.gn && I.b != cur_b) {
                // GroupNorm scale/shift of the (concatenated) input of sample b; the SiLU's 0.5 is
                // folded in: silu(x) = h + h*tanh(h), h = x/2.
                named_bar_sync(1, PROD_THREADS);
                gn_build_affine(p, I.b, sAff, sSum, pt, PROD_THREADS, 1);
                named_bar_sync(1, PROD_THREADS);
                cur_b = I.b;
            }
            for (int kc = 0; kc < n_chunks; ++kc) {
                const bool is_skip = kc >= p.n_main;
                const int cbase = is_skip ? (kc - p.n_main) * KC : kc * KC;  // channel offset inside its concat space
                const int CA = is_skip ? p.S0 : p.C0;
                const bool first = cbase < CA;
                const __nv_bfloat16 *src = is_skip ? (first ? p.skip0 : p.skip1) : (first ? p.src0 : p.src1);
                const int Cs = is_skip ? (first ? p.S0 : p.S1) : (first ? p.C0 : p.C1);
                const int cs = (first ? cbase : cbase - CA) + 8 * g;
                const bool do_gn = p.gn && !is_skip, do_silu = p.silu && !is_skip;
                const bool xform = do_gn || do_silu;
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
                // the skip conv is a 1x1 on the block input at the output resolution (no x2)
                const int srcH = is_skip ? p.H : p.Hin, srcW = is_skip ? p.W : p.Win;
                const bool ups = p.upsample && !is_skip;
                // plane-major activations [B][C/8][H][W][8]: this thread's 8-channel plane of the source
                const __nv_bfloat16 *sbase = src + (size_t(I.b) * (Cs >> 3) + (cs >> 3)) * srcH * srcW * 8;
                const int yb = I.y0 - p.pad, xb = I.x0 - p.pad;

                mbar_wait(empty + stage, phase ^ 1u);
                uint8_t *dstA = sA + size_t(stage) * p.a_stage + size_t(g) * WN * 16;

                auto load_batch = [&](int qb, uint4(&raw)[UB], uint32_t &ok) {
                    ok = 0;
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int q = qb + u * QS;
                        raw[u] = make_uint4(0u, 0u, 0u, 0u);
                        if (q < nq) {
                            const int r = int((uint32_t(q) * p.magicP) >> 20), c = q - r * P;
                            const int y = yb + r, x = xb + c;
                            if (unsigned(y) < unsigned(p.H) && unsigned(x) < unsigned(p.W)) {
                                const int sy = ups ? (y >> 1) : y, sx = ups ? (x >> 1) : x;
                                raw[u] = ldg_nc16(sbase + (size_t(sy) * srcW + sx) * 8);
                                ok |= 1u << u;
                            }
                        }
                    }
                };
                auto xform_store = [&](int qb, const uint4(&raw)[UB], uint32_t ok) {
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int q = qb + u * QS;
                        if (q < nq) {
                            uint4 o = raw[u];  // zeros outside the image: padding happens AFTER GN+SiLU
                            if (xform && ((ok >> u) & 1u)) {
                                const uint32_t w4[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
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
                            }
                            *reinterpret_cast<uint4 *>(dstA + size_t(q) * 16) = o;
                        }
                    }
                };
                // two batches of UB loads in flight per thread
                uint4 rawA[UB], rawB[UB];
                uint32_t okA, okB;
                int qb = qf;
                load_batch(qb, rawA, okA);
                for (;;) {
                    const int q1 = qb + UB * QS;
                    if (q1 < nq) load_batch(q1, rawB, okB);
                    xform_store(qb, rawA, okA);
                    if (q1 >= nq) break;
                    const int q2 = q1 + UB * QS;
                    if (q2 < nq) load_batch(q2, rawA, okA);
                    xform_store(q1, rawB, okB);
                    if (q2 >= nq) break;
                    qb = q2;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_a + stage);
                if (++stage == NS) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // =========================== MMA issue ==================================================
        // One thread issues every tcgen05.mma.  Descriptors are kept as 32-bit halves: hi = SBO (128 B) |
        // version, lo = address/16 | LBO/16 << 16, and since every operand offset is a whole number of
        // 16-byte rows the lo word advances by plain integer adds (no carry into the LBO field: shared
        // memory addresses stay below 2^18).
        {
            int stage = 0, acc_it = 0;
            uint32_t phase = 0;
            if (p.resident) mbar_wait(w_res, 0);
            const uint32_t desc_hi = 8u | (1u << 14);
            const uint32_t a0 = (smem_u32(sA) >> 4) | (uint32_t(WN) << 16);
            const uint32_t w0 = smem_u32(sW) >> 4;
            const uint32_t a_stage16 = p.a_stage >> 4, w_stage16 = p.w_stage >> 4;
            const uint32_t kA = 2u * uint32_t(WN);  // two 8-channel planes per K = 16 step
            for (int it = it_begin; it < it_end; ++it, ++acc_it) {
                const int buf = p.acc2 ? (acc_it & 1) : 0;
                const uint32_t aph = p.acc2 ? uint32_t((acc_it >> 1) & 1) : uint32_t(acc_it & 1);
                mbar_wait(acc_empty + buf, aph ^ 1u);
                tc_fence_after();
                const uint32_t d0 = tmem_base + uint32_t(buf * p.MB * NT);
                for (int kc = 0; kc < n_chunks; ++kc) {
                    const bool is_skip = kc >= p.n_main;
                    const int ntap = is_skip ? 1 : p.taps;
                    mbar_wait(full_a + stage, phase);
                    if (!p.resident) mbar_wait(full_w + stage, phase);
                    tc_fence_after();
                    const uint32_t aaddr = a0 + uint32_t(stage) * a_stage16;
                    uint32_t waddr;
                    if (p.resident)
                        waddr = is_skip ? w0 + (p.w_main_bytes >> 4) + uint32_t((kc - p.n_main) * PL * NT) : w0 + uint32_t(kc * PL * ntap * NT);
                    else
                        waddr = w0 + uint32_t(stage) * w_stage16;
                    waddr |= uint32_t(ntap * NT) << 16;  // LBO of B: one 8-channel plane = ntap * NT rows of 16 bytes
                    const uint32_t kB = 2u * uint32_t(ntap * NT);
                    if (ntap == 9) {
                        for (int mb = 0; mb < p.MB; ++mb) {
                            const uint32_t d = d0 + uint32_t(mb * NT);
                            const uint32_t arow = aaddr + uint32_t(mb * 128);
                            uint32_t acc = kc > 0 ? 1u : 0u;
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint32_t at = arow + uint32_t((tap / 3) * P + (tap % 3));
                                const uint32_t bt = waddr + uint32_t(tap * NT);
#pragma unroll
                                for (int k16 = 0; k16 < PL / 2; ++k16) {
                                    umma_bf16_split(d, at + k16 * kA, desc_hi, bt + k16 * kB, desc_hi, p.idesc, acc);
                                    acc = 1u;
                                }
                            }
                        }
                    } else {
                        // 1x1 conv, or the fused 1x1 skip conv of a 3x3 block (centre tap of the window)
                        const uint32_t shift = is_skip ? uint32_t(p.pad * P + p.pad) : 0u;
                        for (int mb = 0; mb < p.MB; ++mb) {
                            const uint32_t d = d0 + uint32_t(mb * NT);
                            const uint32_t at = aaddr + uint32_t(mb * 128) + shift;
#pragma unroll
                            for (int k16 = 0; k16 < PL / 2; ++k16)
                                umma_bf16_split(d, at + k16 * kA, desc_hi, waddr + k16 * kB, desc_hi, p.idesc, (kc > 0 || k16 > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit_elect(empty + stage);
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_elect(acc_full + buf);
            }
        }
    } else {
        // =========================== weight loader (TMA bulk copies) ============================
        if (lane == 0 && it_begin < it_end) {
            if (p.resident) {
                mbar_expect_tx(w_res, p.w_main_bytes + p.w_skip_bytes);
                bulk_g2s(sW, p.weight, p.w_main_bytes, w_res);
                if (p.w_skip_bytes) bulk_g2s(sW + p.w_main_bytes, p.skip_w, p.w_skip_bytes, w_res);
            } else {
                int stage = 0;
                uint32_t phase = 0;
                const int planes_main = p.Cin / 8, planes_skip = (p.S0 + p.S1) / 8;
                for (int it = it_begin; it < it_end; ++it) {
                    const int cc = it % p.n_cc;
                    for (int kc = 0; kc < n_chunks; ++kc) {
                        const bool is_skip = kc >= p.n_main;
                        const uint32_t bytes = uint32_t(PL * (is_skip ? 1 : p.taps) * NT) * 16;
                        const __nv_bfloat16 *src =
                            is_skip ? p.skip_w + (size_t(cc) * planes_skip + size_t(kc - p.n_main) * PL) * NT * 8
                                    : p.weight + (size_t(cc) * planes_main + size_t(kc) * PL) * p.taps * NT * 8;
                        mbar_wait(empty + stage, phase ^ 1u);
                        mbar_expect_tx(full_w + stage, bytes);
                        bulk_g2s(sW + size_t(stage) * p.w_stage, src, bytes, full_w + stage);
                        if (++stage == NS) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, uint32_t(p.tmem_cols));
    }
}

// ---- host-side configuration --------------------------------------------------------------------
struct WsCfg {
    int PL, R, Wt, P, MB, WN, NT, n_cc, NS, resident, acc2, tmem_cols, tiles_x, tiles, n_main, n_skip, n_items, grid, ips, slots;
    uint32_t a_stage, w_stage, w_main_bytes, w_skip_bytes, magicP;
    size_t smem;
};

constexpr size_t kSmemBudget = 222 * 1024;   // of the 227 KB a CTA may opt in to
constexpr size_t kResidentMax = 80 * 1024;   // weights kept in smem for the whole launch when they fit
constexpr int kNumSMs = 148;

int ws_nt(int Cout) { return tc_nt(Cout); }

size_t ws_fixed_smem(int Cin, int NT, int CoutP) {
    return sizeof(float) * (2 * size_t(Cin) + NT + size_t(EPI_WARPS) * CoutP * 2) + (3 * MAX_STAGES + 5) * 8 + 64 + 16 * size_t(Cin) + 16;
}

bool ws_configure(int B, int H, int W, int C0, int C1, int S0, int S1, int Cout, int ksize, WsCfg &best) {
    const int Cin = C0 + C1, Sk = S0 + S1;
    if (Cin <= 0 || (C0 % 16) || (C1 % 16) || (S0 % 16) || (S1 % 16)) return false;
    const int pad = ksize / 2, taps = ksize * ksize;
    const int CoutP = (Cout + 15) / 16 * 16;
    WsCfg c{};
    c.NT = ws_nt(Cout);
    c.n_cc = CoutP / c.NT;
    const bool all32 = !(C0 % 32) && !(C1 % 32) && !(S0 % 32) && !(S1 % 32);
    c.PL = all32 ? 4 : 2;
    const int KC = 8 * c.PL;
    c.n_main = Cin / KC;
    c.n_skip = Sk / KC;
    const int n_chunks = c.n_main + c.n_skip;
    c.Wt = W > 64 ? 64 : W;
    c.P = c.Wt + 2 * pad;
    c.magicP = uint32_t(((1u << 20) + c.P - 1) / c.P);
    c.tiles_x = (W + c.Wt - 1) / c.Wt;
    c.w_main_bytes = uint32_t(size_t(Cin) * taps * c.NT * 2);
    c.w_skip_bytes = uint32_t(size_t(Sk) * c.NT * 2);
    const size_t w_total = size_t(c.w_main_bytes) + c.w_skip_bytes;
    c.resident = (c.n_cc == 1 && w_total <= kResidentMax) ? 1 : 0;
    c.w_stage = c.resident ? 0u : uint32_t(c.PL * taps * c.NT * 16);
    const size_t fixed = ws_fixed_smem(Cin, c.NT, CoutP) + (c.resident ? w_total : 0);
    double best_cost = 1e300;
    bool found = false;
    for (int R = 1; R <= H; ++R) {
        const int MB = (R * c.P + 127) / 128;
        if (MB * c.NT > 512) break;
        int WN = MB * 128 + 2 * pad * c.P + 2 * pad;
        const int want = 8 / c.PL;  // plane stride == 8/PL (mod 8) 16-byte units: conflict-free STS.128 across the planes
        WN += ((want - (WN & 7)) + 8) & 7;
        bool magic_ok = true;
        for (int q = 0; q < WN + 128; ++q)
            if (int((uint32_t(q) * c.magicP) >> 20) != q / c.P) magic_ok = false;
        if (!magic_ok) continue;
        const size_t a_stage = size_t(c.PL) * WN * 16;
        if (fixed + 2 * (a_stage + c.w_stage) > kSmemBudget) break;
        int NS = int((kSmemBudget - fixed) / (a_stage + c.w_stage));
        NS = NS > MAX_STAGES ? MAX_STAGES : NS;
        const int want_ns = 2 * n_chunks > 3 ? 2 * n_chunks : 3;  // no point in more than two items' worth of stages
        NS = NS > want_ns ? want_ns : NS;
        const int tiles = ((H + R - 1) / R) * c.tiles_x;
        const long long items = (long long)B * tiles * c.n_cc;
        const int grid = int(items < kNumSMs ? items : kNumSMs);
        const long long per_cta = (items + grid - 1) / grid;
        // producer work per item (window positions x channels) + a fixed per-item hand-off cost
        const double item_cost = double((R + 2 * pad) * c.P) * (Cin + 0.5 * Sk) + 160.0 * (Cin + Sk) + 4000.0;
        double cost = double(per_cta) * item_cost;
        if (NS < 3) cost *= 1.3;
        if (2 * MB * c.NT > 512) cost *= 1.15;  // single accumulator buffer: epilogue not overlapped
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
            best.R = R; best.MB = MB; best.WN = WN; best.NS = NS; best.a_stage = uint32_t(a_stage);
            best.acc2 = 2 * MB * c.NT <= 512;
            best.tiles = tiles; best.n_items = int(items); best.grid = grid;
            int cols = 32;
            while (cols < (best.acc2 ? 2 : 1) * MB * c.NT) cols *= 2;
            best.tmem_cols = cols;
            best.smem = fixed + size_t(NS) * (a_stage + c.w_stage);
            found = true;
        }
    }
    if (found) {
        // statistics slots: the most CTAs whose contiguous item range touches one sample
        best.ips = best.tiles * best.n_cc;
        best.slots = 1;
        for (int b = 0; b < B; ++b) {
            const int c_first = int((((long long)b * best.ips + 1) * best.grid - 1) / best.n_items);
            const int c_last = int(((long long)(b + 1) * best.ips * best.grid - 1) / best.n_items);
            if (c_last - c_first + 1 > best.slots) best.slots = c_last - c_first + 1;
        }
    }
    return found;
}

bool ws_configure_op(const ccdm_op &op, WsCfg &c) {
    return ws_configure(op.B, op.Hout, op.Wout, op.C0, op.C1, op.S0, op.S1, op.Cout, op.ksize, c);
}

}  // namespace

bool conv_tc_supported(const ccdm_op &op) {
    if (op.dtype != CCDM_DT_BF16 || op.src_kind != 0 || op.stride != 1) return false;
    if (op.ksize != 1 && op.ksize != 3) return false;
    if ((op.C0 % 16) || (op.C1 % 16) || (op.S0 % 16) || (op.S1 % 16)) return false;
    if (op.out_dtype == CCDM_DT_BF16 && (op.Cout % 16)) return false;
    WsCfg c;
    return ws_configure_op(op, c);
}

int conv_tc_nt(int Cout) { return ws_nt(Cout); }

// {PL, R, Wt, MB, WN, NT, n_cc, NS, resident, acc2, tmem_cols, tiles, n_items, grid, smem bytes, n_chunks}
int conv_tc_config(const ccdm_op &op, int32_t *out) {
    WsCfg c;
    if (!conv_tc_supported(op) || !ws_configure_op(op, c)) return -1;
    const int32_t v[16] = {c.PL, c.R, c.Wt, c.MB, c.WN, c.NT, c.n_cc, c.NS, c.resident, c.acc2, c.tmem_cols, c.tiles, c.n_items, c.grid,
                           int32_t(c.smem), c.n_main + c.n_skip};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    return 0;
}

int conv_tc_stat_layout(const ccdm_op &op, int32_t *out5) {
    WsCfg c;
    if (!conv_tc_supported(op) || !ws_configure_op(op, c)) return -1;
    out5[0] = c.slots; out5[1] = c.ips; out5[2] = c.n_items; out5[3] = c.grid; out5[4] = (op.Cout + 15) / 16 * 16;
    return 0;
}

size_t conv_tc_part_floats(const ccdm_op &op) {
    WsCfg c;
    if (!ws_configure_op(op, c)) return 0;
    return size_t(op.B) * c.slots * ((op.Cout + 15) / 16 * 16) * 2;
}

int launch_conv_tc(const ccdm_op &op, cudaStream_t s) {
    WsCfg c;
    if (!conv_tc_supported(op) || !ws_configure_op(op, c)) CCDM_FAIL(-3, "conv_ws: unsupported configuration");
    WsP p{};
    p.src0 = (const __nv_bfloat16 *)op.src0; p.src1 = (const __nv_bfloat16 *)op.src1;
    p.stat0 = (const double *)op.stat0; p.stat1 = (const double *)op.stat1;
    p.gamma = (const float *)op.gamma; p.beta = (const float *)op.beta;
    p.weight = (const __nv_bfloat16 *)op.weight; p.bias = (const float *)op.bias; p.emb = (const float *)op.emb;
    p.skip0 = (const __nv_bfloat16 *)op.skip0; p.skip1 = (const __nv_bfloat16 *)op.skip1;
    p.skip_w = (const __nv_bfloat16 *)op.skip_w; p.res = (const __nv_bfloat16 *)op.res;
    p.out = (void *)op.out; p.ostat = (double *)op.ostat; p.part = (float *)op.part; p.ticket = (unsigned int *)op.ticket;
    p.steps = (const ccdm_step_entry *)op.steps; p.step_ptr = (const int *)op.step_ptr;
    p.B = op.B; p.Hin = op.Hin; p.Win = op.Win; p.H = op.Hout; p.W = op.Wout;
    p.C0 = op.C0; p.C1 = op.C1; p.Cin = op.C0 + op.C1; p.Cout = op.Cout; p.CoutP = (op.Cout + 15) / 16 * 16;
    p.NT = c.NT; p.n_cc = c.n_cc;
    p.upsample = op.upsample; p.gn = op.gn; p.silu = op.silu; p.S0 = op.S0; p.S1 = op.S1;
    p.emb_off = op.emb_off; p.emb_cols = op.emb_cols; p.emb_bstride = op.emb_bstride;
    p.out_f32 = op.out_dtype == CCDM_DT_F32;
    if (p.out_f32 && (reinterpret_cast<uintptr_t>(op.out) & 15)) CCDM_FAIL(-2, "conv: fp32 output must be 16-byte aligned");
    p.R = c.R; p.Wt = c.Wt; p.P = c.P; p.MB = c.MB; p.WN = c.WN; p.tiles_x = c.tiles_x; p.tiles = c.tiles;
    p.taps = op.ksize * op.ksize; p.pad = op.ksize / 2;
    p.n_main = c.n_main; p.n_skip = c.n_skip; p.NS = c.NS; p.resident = c.resident; p.acc2 = c.acc2;
    p.tmem_cols = c.tmem_cols; p.n_items = c.n_items; p.ips = c.ips; p.slots = c.slots; p.nsub = 1;
    for (int i = 0; i < 2; ++i) {
        p.st_slots[i] = op.st_slots[i]; p.st_ips[i] = op.st_ips[i]; p.st_items[i] = op.st_items[i];
        p.st_grid[i] = op.st_grid[i]; p.st_rows[i] = op.st_rows[i];
    }
    p.a_stage = c.a_stage; p.w_stage = c.w_stage; p.w_main_bytes = c.w_main_bytes; p.w_skip_bytes = c.w_skip_bytes;
    p.magicP = c.magicP;
    // cute::UMMA::InstrDescriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 at 17, M>>4 at 24
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(c.NT >> 3) << 17) | (uint32_t(128 >> 4) << 24);

    if (op.gn && (!op.stat0 || (op.C1 && !op.stat1) || !op.gamma || !op.beta)) CCDM_FAIL(-2, "conv_ws: gn without stats/affine");
    if (op.gn && (p.Cin % kGnGroups)) CCDM_FAIL(-2, "conv_ws: GroupNorm needs Cin %% 32 == 0");
    if (op.ostat && (!op.part || !op.ticket)) CCDM_FAIL(-2, "conv_ws: ostat without scratch");
    if (op.emb && (!op.steps || !op.step_ptr || op.emb_off < 0)) CCDM_FAIL(-2, "conv_ws: emb without step table");
    if (op.S0 > 0 && (!op.skip0 || !op.skip_w)) CCDM_FAIL(-2, "conv_ws: bad skip configuration");
    {
        const int expH = op.upsample ? op.Hin * 2 : op.Hin, expW = op.upsample ? op.Win * 2 : op.Win;
        if (expH != op.Hout || expW != op.Wout) CCDM_FAIL(-2, "conv_ws: inconsistent shapes");
    }
    auto kern = c.PL == 4 ? conv_ws_kernel<4> : conv_ws_kernel<2>;
    // The opt-in shared memory limit is a property of the FUNCTION, not of a launch: set it once to the
    // budget (never per launch -- a captured graph replays nodes long after a later launch lowered it).
    static bool attr_done = false;
    if (!attr_done) {
        CCDM_CUDA(cudaFuncSetAttribute(conv_ws_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBudget)));
        CCDM_CUDA(cudaFuncSetAttribute(conv_ws_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBudget)));
        attr_done = true;
    }
    CCDM_CUDA(launch_pdl(kern, dim3(c.grid), dim3(WS_THREADS), c.smem, s, p));
    CCDM_LAUNCH_CHECK("conv_ws_kernel");
    return 0;
}

}  // namespace ccdm
