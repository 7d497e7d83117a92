// Fused GroupNorm + SiLU + conv (3x3 / 1x1, stride 1) on the 5th-generation tensor cores, fed by TMA.
//
// Implicit GEMM over a "flattened padded tile": the M rows of an MMA are consecutive positions of the (haloed) window
// of a tile, so the 3x3 taps are pure shifts of the A descriptor's start address.  The data path is the one the
// hardware is built for:
//
//   * bf16 activations live in HBM PLANE-MAJOR, [B][C/8][H][W][8]: 16 bytes per (8-channel plane, pixel),
//     i.e. exactly one row of a tcgen05 "K-major, no swizzle" core matrix.  A 4-D TMA box
//     {window width, window rows, planes of the K chunk, sample} therefore lands in shared memory ALREADY
//     in the MMA operand layout (plane stride = LBO), zero-filled outside the image, with one instruction
//     issued by one thread -- no register staging, no address arithmetic, any number of stages in flight.
//   * GroupNorm-apply + SiLU is an IN-PLACE pass over the landed stage (LDS.128 -> fp32 affine, tanh.approx
//     -> bf16 -> STS.128, conflict-free), done by 8 warps that never see global-memory latency.  Raw chunks
//     (1x1 skip conv, convs without a norm) skip the pass: TMA -> MMA directly.
//   * tcgen05.mma is issued by THREE warps (M blocks mw, mw+3, ...): at N = 32 a single issuing thread
//     cannot keep the tensor pipe fed (measured: ~90 cycles of issue per 16-cycle MMA).
//   * epilogue (conv_tc_common.cuh): TMEM -> registers -> +bias +embedding +residual -> plane-major store
//     (a warp writes 512 contiguous bytes per plane) + GroupNorm statistics of the output.
//
// Replaces (reference unet.py): ResBlock convs :242-262 (with GroupNorm32 nn.py:93-100 + SiLU), the 1x1
// skip_connection :221-228,262, attention qkv / proj_out :305-311, the output conv :701-705.
//
// Roles (608 threads, one CTA per SM, each CTA walks a contiguous range of work items):
//   warps 0-7   epilogue | warps 8-15 in-place transform | warps 16-18 MMA issue | warp 19 TMA.
// All hand-offs are mbarriers; nothing in the main loop is a CTA-wide barrier.
#include <cuda.h>
#include <stdlib.h>

namespace ccdm {
namespace {
__device__ void conv_tma_trace_hook(int slot);
__device__ void conv_tma_tl_hook(int item, int edge);
}
}  // namespace ccdm
#define CCDM_EPI_TRACE(slot) conv_tma_trace_hook(slot)
#define CCDM_EPI_TL(item, edge) conv_tma_tl_hook(item, edge)
#include "conv_tc_common.cuh"

namespace ccdm {
namespace {

constexpr int TM_EPI_WARPS = 8, XF_WARPS = 8, MMA_WARPS = 3;
constexpr int XF_THREADS = XF_WARPS * 32;
constexpr int WARP_XF0 = TM_EPI_WARPS, WARP_MMA0 = WARP_XF0 + XF_WARPS, WARP_TMA = WARP_MMA0 + MMA_WARPS;
constexpr int TM_THREADS = (WARP_TMA + 1) * 32;  // 640: five warpgroups of four warps

struct alignas(64) TmP {
    CUtensorMap map[4];  // src0, src1, skip0, skip1 (plane-major tensors, see make_map)
    WsP w;
};

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// GroupNorm affine (+ SiLU as h + h*tanh(h), h = x/2 with the 1/2 folded into fa/fb) of one 16-byte row
// (8 channels of one pixel), fp32 maths, bf16 in and out.
// SILU: 0 = none, 1 = fp32 tanh.approx per element (default), 2 = EXPERIMENT (CCDM_SILU_MODE=2): the affine stays fp32,
// h is rounded to bf16x2 and h + h*tanh(h) is evaluated with tanh.approx.bf16x2 + one packed fma -- 28 instead of 44
// maths instructions per row and half the MUFU work, at the price of two extra bf16 roundings of the activation.
template <int SILU>
__device__ __forceinline__ uint4 xf_row(uint4 raw, const float (&fa)[8], const float (&fb)[8]) {
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t r4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = unpack_bf16(w4[i]);
        float h0 = fmaf(v.x, fa[2 * i], fb[2 * i]);
        float h1 = fmaf(v.y, fa[2 * i + 1], fb[2 * i + 1]);
        if (SILU == 1) {
            h0 = fmaf(h0, tanh_approx(h0), h0);
            h1 = fmaf(h1, tanh_approx(h1), h1);
        }
        if (SILU == 3) {  // x * sigmoid(x) = x / (1 + 2^(-x log2 e)): two full-rate MUFU ops instead of one tanh
            float e0, e1, r0, r1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(h0 * -1.4426950408889634f));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(h1 * -1.4426950408889634f));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.0f + e0));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.0f + e1));
            h0 *= r0;
            h1 *= r1;
        }
        r4[i] = pack_bf16(h0, h1);
        if (SILU == 2) {
            uint32_t t;
            asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(r4[i]));
            asm("fma.rn.bf16x2 %0, %1, %2, %1;" : "=r"(r4[i]) : "r"(r4[i]), "r"(t));
        }
    }
    return make_uint4(r4[0], r4[1], r4[2], r4[3]);
}

// In-place pass of one thread over its rows q0, q0 + STEP, ... < NQ of one plane of a landed stage.  MASK:
// the tile touches the image border; halo rows outside the image were zero-filled by TMA and must STAY
// zero (the reference pads after GroupNorm + SiLU), so they are skipped; (r, c) tracks the window
// coordinates of the current row incrementally.
template <int STEP, bool MASK, int SILU>
__device__ __forceinline__ void xf_pass(uint32_t addr, int q, int NQ, int r, int c, int dr, int dc, int P, int ymin, int xmin, int H,
                                        int W, const float (&fa)[8], const float (&fb)[8]) {
    auto advance = [&]() {
        if (MASK) {
            c += dc;
            r += dr;
            if (c >= P) {
                c -= P;
                ++r;
            }
        }
    };
    auto ok = [&]() -> bool { return !MASK || (unsigned(ymin + r) < unsigned(H) && unsigned(xmin + c) < unsigned(W)); };
    // Branch-free on purpose: the four rows of an iteration are independent, and only without a per-row branch can
    // the compiler interleave their maths (a row alone is a ~100-cycle dependent chain LDS -> FMA -> MUFU -> FMA ->
    // pack -> STS).  Rows outside the image hold zeros from the TMA fill; they are transformed like the others and
    // the result is replaced by zero with four selects.
    for (; q + 3 * STEP < NQ; q += 4 * STEP, addr += 4u * STEP * 16u) {
        uint4 raw[4], res[4];
        bool keep[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) raw[u] = lds128(addr + uint32_t(u) * STEP * 16u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            keep[u] = ok();
            advance();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) res[u] = xf_row<SILU>(raw[u], fa, fb);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MASK) {
                res[u].x = keep[u] ? res[u].x : 0u;
                res[u].y = keep[u] ? res[u].y : 0u;
                res[u].z = keep[u] ? res[u].z : 0u;
                res[u].w = keep[u] ? res[u].w : 0u;
            }
            sts128(addr + uint32_t(u) * STEP * 16u, res[u]);
        }
    }
    for (; q < NQ; q += STEP, addr += uint32_t(STEP) * 16u) {
        uint4 res = xf_row<SILU>(lds128(addr), fa, fb);
        if (MASK && !ok()) res = make_uint4(0u, 0u, 0u, 0u);
        sts128(addr, res);
        advance();
    }
}

// ---- fp16x2 ("exact" tensor-core mode) transform --------------------------------------------------------------
// One position of one 8-channel group = a hi row and a lo row (NQ rows apart): x = (hi + lo) * 2^-4 exactly in fp32,
// GroupNorm affine (the 2^-4 is folded into fa) and SiLU as h / (1 + exp(-h)) with full-rate MUFU ex2 / rcp (a few ulp;
// tanh.approx of the bf16 path is good to 2^-11 only), then split again.
template <bool SILU>
__device__ __forceinline__ void xf_row_x3(uint4 &hi, uint4 &lo, const float (&fa)[8], const float (&fb)[8]) {
    uint32_t h4[4] = {hi.x, hi.y, hi.z, hi.w}, l4[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 xh = unpack_f16x2(h4[i]), xl = unpack_f16x2(l4[i]);
        float h0 = fmaf(xh.x, fa[2 * i], fmaf(xl.x, fa[2 * i], fb[2 * i]));
        float h1 = fmaf(xh.y, fa[2 * i + 1], fmaf(xl.y, fa[2 * i + 1], fb[2 * i + 1]));
        if (SILU) {
            h0 = __fdividef(h0, 1.0f + __expf(-h0));
            h1 = __fdividef(h1, 1.0f + __expf(-h1));
        }
        split_f16x2(h0, h1, h4[i], l4[i]);
    }
    hi = make_uint4(h4[0], h4[1], h4[2], h4[3]);
    lo = make_uint4(l4[0], l4[1], l4[2], l4[3]);
}

// In-place pass over positions q0, q0 + STEP, ... < NQ of one (hi, lo) plane pair; `addr` = the hi row of q0, the lo row
// is lo_off bytes further.  Two positions per iteration (32 live registers of data); halo positions stay zero (see xf_pass).
template <int STEP, bool MASK, bool SILU>
__device__ __forceinline__ void xf_pass_x3(uint32_t addr, uint32_t lo_off, int q, int NQ, int r, int c, int dr, int dc, int P, int ymin,
                                           int xmin, int H, int W, const float (&fa)[8], const float (&fb)[8]) {
    auto advance = [&]() {
        if (MASK) {
            c += dc;
            r += dr;
            if (c >= P) {
                c -= P;
                ++r;
            }
        }
    };
    auto ok = [&]() -> bool { return !MASK || (unsigned(ymin + r) < unsigned(H) && unsigned(xmin + c) < unsigned(W)); };
    for (; q + STEP < NQ; q += 2 * STEP, addr += 2u * STEP * 16u) {
        uint4 hi[2], lo[2];
        bool keep[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            hi[u] = lds128(addr + uint32_t(u) * STEP * 16u);
            lo[u] = lds128(addr + uint32_t(u) * STEP * 16u + lo_off);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            keep[u] = ok();
            advance();
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) xf_row_x3<SILU>(hi[u], lo[u], fa, fb);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (MASK && !keep[u]) hi[u] = lo[u] = make_uint4(0u, 0u, 0u, 0u);
            sts128(addr + uint32_t(u) * STEP * 16u, hi[u]);
            sts128(addr + uint32_t(u) * STEP * 16u + lo_off, lo[u]);
        }
    }
    for (; q < NQ; q += STEP, addr += uint32_t(STEP) * 16u) {
        uint4 hi = lds128(addr), lo = lds128(addr + lo_off);
        xf_row_x3<SILU>(hi, lo, fa, fb);
        if (MASK && !ok()) hi = lo = make_uint4(0u, 0u, 0u, 0u);
        sts128(addr, hi);
        sts128(addr + lo_off, lo);
        advance();
    }
}

// Milestone time stamps of CTA 0 (debug aid, read back with ccdm_debug_conv_trace): one store per milestone.
__device__ unsigned long long g_trace[16];
enum { kTraceStart = 0, kTraceSetup, kTraceAffine, kTraceRaw0, kTraceXf0, kTraceMma0, kTraceEpi0, kTraceFlush, kTraceEnd };
// Steady-state timeline of CTA 0 (CCDM_TRACE builds): begin / end stamps of the first 8 items per role
// {0 TMA issue, 1 transform, 2 MMA warp 0, 3 epilogue warp 0}.
__device__ unsigned long long g_tl[4][8][2];
// start / end stamp of every CTA of the last launch (CCDM_TRACE builds): the spread shows load imbalance
__device__ unsigned long long g_cta[2][160];
__device__ __forceinline__ void cta_stamp(int edge) {
#ifdef CCDM_TRACE
    if (blockIdx.x < 160) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_cta[edge][blockIdx.x] = t;
    }
#else
    (void)edge;
#endif
}
__device__ __forceinline__ void tl(int role, int item, int edge) {
#ifdef CCDM_TRACE
    if (blockIdx.x == 0 && item < 8) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_tl[role][item][edge] = t;
    }
#else
    (void)role; (void)item; (void)edge;
#endif
}
__device__ __forceinline__ void trace(int slot) {
#ifdef CCDM_TRACE
    if (blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[slot] = t;
    }
#else
    (void)slot;
#endif
}

__device__ void conv_tma_trace_hook(int slot) { trace(slot); }
__device__ void conv_tma_tl_hook(int item, int edge) { tl(3, item, edge); }

// PL = planes (8-channel groups) per K chunk: KC = 8*PL channels per pipeline stage.
// UP: the fused nearest-x2 + conv3x3 variant (16 pre-summed sub-pixel taps, 4 accumulator sets per M block); a separate
// instantiation so that the common kernel's code and register allocation are untouched by it.
// XF: which in-place transform the kernel carries -- 0 none (convs without GroupNorm/SiLU: the role idles), 1 GroupNorm +
// SiLU, 2 GroupNorm only (q/k/v convs), 3 both (selected at run time).  Separate instantiations keep each kernel small.
// X3: fp16x2 operands (CCDM_DT_F16X2).  A stage holds 2*PL planes -- (hi, lo) of PL 8-channel groups -- the weights carry
// NT hi rows + NT lo rows per (plane, tap), and every product is three MMAs into the same accumulator:
// A_hi B_hi + A_hi B_lo + A_lo B_hi (the dropped lo*lo term is 2^-22 of the product).
template <int PL, bool UP, bool LEAN, int XF, bool X3 = false>
__global__ void __launch_bounds__(TM_THREADS, 1) conv_tma_kernel(const __grid_constant__ TmP P_) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const WsP &p = P_.w;
    constexpr int KC = 8 * PL;
    constexpr int X = X3 ? 2 : 1;  // smem planes per 8-channel group; B rows per (plane, tap) in units of NT
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = p.NT, P = p.P, NS = p.NS, NQ = p.NQ;
    if (tid == 0) trace(kTraceStart), cta_stamp(0);

    // smem carve-up: [NS] activation stages (+ over-read slack) | weights | GN affine | bias | stats | barriers
    uint8_t *sA = smem_raw;
    const size_t slack = size_t(128 + 2 * p.pad * P + 2 * p.pad) * 16;
    uint8_t *sW = sA + size_t(NS) * p.a_stage + ((slack + 127) & ~size_t(127));
    const size_t w_region = p.resident ? size_t(p.w_main_bytes) + p.w_skip_bytes : size_t(NS) * p.w_stage;
    float *sAff = reinterpret_cast<float *>(sW + w_region);  // [2][Cin]
    float *sAdd = sAff + 2 * p.Cin;                          // [NT]
    float *sRed = sAdd + NT;                                 // [8][CoutP][2]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sRed + TM_EPI_WARPS * p.CoutP * 2 * ((X3 && UP) ? 2 : 1));  // (fp16x2: four rows of doubles = the same bytes; eight when upsampling)
    uint64_t *raw_full = bars, *xf_full = bars + MAX_STAGES, *empty = bars + 2 * MAX_STAGES;
    uint64_t *acc_full = bars + 3 * MAX_STAGES, *acc_empty = acc_full + 2, *w_res = acc_empty + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(w_res + 1);
    int *s_last = reinterpret_cast<int *>(s_tmem + 1);
    double2 *sSum = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(s_last + 1) + 15) & ~uintptr_t(15));  // [Cin] GroupNorm fold scratch

    if (warp == WARP_MMA0) tmem_alloc(s_tmem, uint32_t(p.tmem_cols));
    if (tid == WARP_TMA * 32) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(raw_full + s, 1);
            mbar_init(xf_full + s, XF_WARPS);
            mbar_init(empty + s, MMA_WARPS);
        }
        mbar_init(acc_full + 0, MMA_WARPS);
        mbar_init(acc_full + 1, MMA_WARPS);
        mbar_init(acc_empty + 0, TM_EPI_WARPS);
        mbar_init(acc_empty + 1, TM_EPI_WARPS);
        mbar_init(w_res, 1);
        fence_barrier_init();
        tma_prefetch_desc(&P_.map[0]);
        if (p.C1) tma_prefetch_desc(&P_.map[1]);
        if (p.S0) tma_prefetch_desc(&P_.map[2]);
        if (p.S1) tma_prefetch_desc(&P_.map[3]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    if (tid == 0 && (smem_u32(sA) & 127u)) __trap();  // dynamic shared memory must be 128-byte aligned for the TMA boxes

    const int it_begin = int((long long)blockIdx.x * p.n_items / gridDim.x);
    const int it_end = int((long long)(blockIdx.x + 1) * p.n_items / gridDim.x);
    const int n_chunks = p.n_main + p.n_skip;

    // Programmatic dependent launch: the prologue above overlapped the previous kernel's tail; the weights are
    // constants of the chain and may be fetched before it completes, everything else waits for it.
    pdl_launch_dependents();
    if (warp == WARP_TMA && lane == 0 && it_begin < it_end && p.resident) {
        mbar_expect_tx(w_res, p.w_main_bytes + p.w_skip_bytes);
        bulk_g2s(sW, p.weight, p.w_main_bytes, w_res);
        if (p.w_skip_bytes) bulk_g2s(sW + p.w_main_bytes, p.skip_w, p.w_skip_bytes, w_res);
    }
    pdl_wait();
    if (tid == 0) trace(kTraceSetup);

    // Register budget per role (setmaxnreg works on whole warpgroups of 4 warps): the kernel launches with
    // 96 registers per thread; the epilogue warpgroups grow to 128, the others shrink and donate theirs.
    // The pool is the CTA's own launch allocation (640 x 96): increases must be covered by the decreases of
    // this CTA -- (8 x 16 + 4 x 40) x 32 freed >= 8 x 32 x 32 needed -- or setmaxnreg.inc waits forever.
    // (Measured: 12 transform warps (768 threads, 80/120/64/40 registers) are not faster than 8: 2.65 vs 2.62 ms
    // per LIDC step -- the roles are limited by shared-memory bandwidth and issue slots, not by warp count.)
    if (warp < TM_EPI_WARPS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        conv_epilogue_role<TM_EPI_WARPS, UP ? 4 : 1, LEAN, X3>(p, sAdd, sRed, s_last, acc_full, acc_empty, tmem_base, it_begin, it_end);
    } else if (warp < WARP_MMA0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        // =========================== in-place GroupNorm + SiLU ======================================
        // Warp xw owns plane (xw % PL) of every transformed stage, so its 8 scale/shift pairs are warp
        // uniform and live in registers; the XF_WARPS/PL warps of a plane interleave blocks of 32 positions.
        if (XF != 0 && p.xf) {
            constexpr bool kSilu = XF == 1 || XF == 3, kPlain = XF == 2 || XF == 3;
            const int xw = warp - WARP_XF0, pt = tid - WARP_XF0 * 32;
            const int plane = xw & (PL - 1), sub = xw / PL;
            constexpr int NSUB = XF_WARPS / PL;
            constexpr int STEP = NSUB * 32;
            const int dr = int((uint32_t(STEP) * p.magicP) >> 20), dc = STEP - dr * P;
            int stage = 0, cur_b = -1;
            uint32_t phase = 0;
            const uint32_t sA32 = smem_u32(sA);
            for (int it = it_begin; it < it_end; ++it) {
                const Item I = decode_item(p, it);
                if (p.gn && I.b != cur_b) {
                    // GroupNorm scale/shift of the (concatenated) input of sample b; the SiLU's 0.5 is
                    // folded in: silu(x) = h + h*tanh(h), h = x/2.
                    named_bar_sync(1, XF_THREADS);
                    gn_build_affine(p, I.b, sAff, sSum, pt, XF_THREADS, 1);
                    named_bar_sync(1, XF_THREADS);
                    if (pt == 0 && it == it_begin) trace(kTraceAffine);
                    cur_b = I.b;
                }
                // halo positions outside the image were zero-filled by TMA and must stay zero (the
                // reference pads AFTER GroupNorm + SiLU): interior tiles skip the test altogether
                const int ymin = I.y0 - p.pad, xmin = I.x0 - p.pad;
                const bool need_mask = ymin < 0 || ymin + p.RW > p.H || xmin < 0 || xmin + P > p.W;
                const int q0 = sub * 32 + lane;
                const int r0 = int((uint32_t(q0) * p.magicP) >> 20), c0 = q0 - r0 * P;
                if (pt == 0) tl(1, it - it_begin, 0);
                for (int kc = 0; kc < n_chunks; ++kc) {
                    // every chunk is acknowledged on xf_full (raw skip-conv chunks without touching them), so
                    // that barrier completes exactly one phase per use of the stage, like the others; waiting
                    // for the TMA first also keeps these warps from running ahead of the ring
                    if (kc >= p.n_main) {
                        mbar_wait<64>(raw_full + stage, phase);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(xf_full + stage);
                    } else {
                        float fa[8], fb[8];
                        if (p.gn) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                fa[i] = sAff[kc * KC + 8 * plane + i];
                                fb[i] = sAff[p.Cin + kc * KC + 8 * plane + i];
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) fa[i] = X3 ? 0.0625f : ((p.silu == 1 || p.silu == 2) ? 0.5f : 1.0f), fb[i] = 0.f;
                        }
                        mbar_wait<64>(raw_full + stage, phase);
                        if (pt == 0 && it == it_begin && kc == 0) trace(kTraceRaw0);
                        const uint32_t addr = sA32 + uint32_t(stage) * p.a_stage + (uint32_t(X * plane) * uint32_t(NQ) + uint32_t(q0)) * 16u;
#ifdef CCDM_ABLATE
                        if (p.dbg & 2) {
                        } else
#endif
                        if constexpr (X3) {
                            const uint32_t lo_off = uint32_t(NQ) * 16u;
                            if (need_mask) {
                                if (kSilu && (!kPlain || p.silu)) xf_pass_x3<STEP, true, true>(addr, lo_off, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                                else if (kPlain) xf_pass_x3<STEP, true, false>(addr, lo_off, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            } else {
                                if (kSilu && (!kPlain || p.silu)) xf_pass_x3<STEP, false, true>(addr, lo_off, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                                else if (kPlain) xf_pass_x3<STEP, false, false>(addr, lo_off, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            }
                        } else
                        // (SiLU modes 2 and 3 are measured-and-rejected experiments: compiled only with -DCCDM_SILU_EXPERIMENTS,
                        // they double the code of this role)
                        if (need_mask) {
#ifdef CCDM_SILU_EXPERIMENTS
                            if (p.silu == 3) xf_pass<STEP, true, 3>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            else if (p.silu == 2) xf_pass<STEP, true, 2>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            else
#endif
                            if (kSilu && (!kPlain || p.silu)) xf_pass<STEP, true, 1>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            else if (kPlain) xf_pass<STEP, true, 0>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                        } else {
#ifdef CCDM_SILU_EXPERIMENTS
                            if (p.silu == 3) xf_pass<STEP, false, 3>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            else if (p.silu == 2) xf_pass<STEP, false, 2>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            else
#endif
                            if (kSilu && (!kPlain || p.silu)) xf_pass<STEP, false, 1>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                            else if (kPlain) xf_pass<STEP, false, 0>(addr, q0, NQ, r0, c0, dr, dc, P, ymin, xmin, p.H, p.W, fa, fb);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(xf_full + stage);
                        if (pt == 0 && it == it_begin && kc == 0) trace(kTraceXf0);
                    }
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (pt == 0) tl(1, it - it_begin, 1);
            }
        }
    } else if (warp < WARP_TMA) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // =========================== MMA issue (three warps: M blocks mw, mw+3, ...) ================
        const int mw = warp - WARP_MMA0;
        int stage = 0, acc_it = 0;
        uint32_t phase = 0;
        if (p.resident) mbar_wait(w_res, 0);
        const uint32_t desc_hi = 8u | (1u << 14);  // SBO = 128 B, descriptor version 1
        const uint32_t a0 = (smem_u32(sA) >> 4) | (uint32_t(X * NQ) << 16);  // LBO of A: stride between 8-channel groups = X planes of NQ rows
        const uint32_t w0 = smem_u32(sW) >> 4;
        const uint32_t a_stage16 = p.a_stage >> 4, w_stage16 = p.w_stage >> 4;
        const uint32_t kA = 2u * uint32_t(X * NQ);
        // one product: a single bf16 MMA, or the three fp16 MMAs of the split operands (lo plane of A: NQ rows after the
        // hi plane; lo rows of B: NT rows after the hi rows)
        const uint32_t a_lo = uint32_t(NQ), b_lo = uint32_t(NT);
        // fp16x2: an accumulator is 2 NT columns.  A_hi x [B_hi; B_lo] (the NT hi rows and the NT lo rows of a (plane, tap)
        // are adjacent: one MMA with N = 2 NT) puts hi*hi into columns [0, NT) and hi*lo into [NT, 2 NT); A_lo x B_hi adds
        // lo*hi to [NT, 2 NT).  Two instructions instead of three, and the large products accumulate on their own: the
        // tensor core truncates on every accumulation (measured: error ~ number of accumulating MMAs x 2^-24 |acc|), so the
        // small terms must not triple the count on the main accumulator.  The epilogue adds the two halves in fp32.
        auto mma = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
#ifdef CCDM_ABLATE
            if (p.dbg & 1) return;
#endif
            if constexpr (X3 && !UP) {
                umma_bf16_split(d, a, desc_hi, b, desc_hi, p.idesc2, acc);
                umma_bf16_split(d + b_lo, a + a_lo, desc_hi, b, desc_hi, p.idesc, 1u);
            } else if constexpr (X3) {
                // upsampling conv: four parity accumulators per M block already; the three split products share one accumulator
                // (N = NT) so that two sets of them still fit the 512 TMEM columns and the epilogue overlaps the next item's MMAs
                // (measured with single-buffered N-concatenated accumulators: 5.7 us per item of pure hand-off latency)
                umma_bf16_split(d, a, desc_hi, b + b_lo, desc_hi, p.idesc, acc);
                umma_bf16_split(d, a + a_lo, desc_hi, b, desc_hi, p.idesc, 1u);
                umma_bf16_split(d, a, desc_hi, b, desc_hi, p.idesc, 1u);
            } else {
                umma_bf16_split(d, a, desc_hi, b, desc_hi, p.idesc, acc);
            }
        };
        for (int it = it_begin; it < it_end; ++it, ++acc_it) {
            const int buf = p.acc2 ? (acc_it & 1) : 0;
            const uint32_t aph = p.acc2 ? uint32_t((acc_it >> 1) & 1) : uint32_t(acc_it & 1);
            mbar_wait<128>(acc_empty + buf, aph ^ 1u);
            tc_fence_after();
            if (mw == 0 && lane == 0) tl(2, it - it_begin, 0);
            constexpr int XA = (X3 && !UP) ? 2 : 1;  // accumulator columns per (M block, parity) in units of NT
            const uint32_t d0 = tmem_base + uint32_t(buf * p.MB * NT * XA * (UP ? 4 : 1));
            for (int kc = 0; kc < n_chunks; ++kc) {
                const bool is_skip = kc >= p.n_main;
                const int ntap = is_skip ? 1 : p.taps;
                mbar_wait<64>((p.xf ? xf_full : raw_full) + stage, phase);
                tc_fence_after();
                const uint32_t aaddr = a0 + uint32_t(stage) * a_stage16;
                uint32_t waddr;
                if (p.resident)
                    waddr = is_skip ? w0 + (p.w_main_bytes >> 4) + uint32_t((kc - p.n_main) * PL * NT * X) : w0 + uint32_t(kc * PL * ntap * NT * X);
                else
                    waddr = w0 + uint32_t(stage) * w_stage16;
                waddr |= uint32_t(ntap * NT * X) << 16;  // LBO of B: one 8-channel plane = ntap * (X * NT) rows of 16 bytes
                const uint32_t kB = 2u * uint32_t(ntap * NT * X);
                if (UP && ntap == 16) {
                    // nearest x2 + conv3x3 as four 2x2 convs on the LOW-resolution window, one per output parity
                    // (py, px): output (2y+py, 2x+px) reads low-res rows y-1+py+ry, columns x-1+px+rx, (ry, rx) in
                    // {0,1}^2, with the 3x3 weights that fall on the same low-res pixel pre-summed on the host.
                    // 16 tap matrices instead of 9, but a quarter of the positions: 2.25x fewer MMAs, 4x less staging.
                    for (int mb = mw; mb < p.MB; mb += MMA_WARPS) {
                        const uint32_t arow = aaddr + uint32_t(mb * 128);
#pragma unroll
                        for (int par = 0; par < 4; ++par) {
                            const uint32_t d = d0 + uint32_t((mb * 4 + par) * NT * XA);
                            uint32_t acc = kc > 0 ? 1u : 0u;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const uint32_t at = arow + uint32_t(((par >> 1) + (t >> 1)) * P + (par & 1) + (t & 1));
                                const uint32_t bt = waddr + uint32_t((par * 4 + t) * NT * X);
#pragma unroll
                                for (int k16 = 0; k16 < PL / 2; ++k16) {
                                    mma(d, at + k16 * kA, bt + k16 * kB, acc);
                                    acc = 1u;
                                }
                            }
                        }
                    }
                } else if (ntap == 9) {
                    for (int mb = mw; mb < p.MB; mb += MMA_WARPS) {
                        const uint32_t d = d0 + uint32_t(mb * NT * XA);
                        const uint32_t arow = aaddr + uint32_t(mb * 128);
                        uint32_t acc = kc > 0 ? 1u : 0u;
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            // stride 1: window position + (dy, dx).  stride 2: input row 2y+dy-1 lives in the
                            // parity sub-image ((dy+1)&1, (dx+1)&1) at sub-row y-1+(dy>0), sub-column x-1+(dx>0),
                            // and every sub-image starts one row / column before the tile
                            const uint32_t at = arow + (p.stride2 ? uint32_t(((((tap / 3) + 1) & 1) * 2 + (((tap % 3) + 1) & 1)) * p.blk16 +
                                                                             ((tap / 3) > 0 ? P : 0) + ((tap % 3) > 0 ? 1 : 0))
                                                                  : uint32_t((tap / 3) * P + (tap % 3)));
                            const uint32_t bt = waddr + uint32_t(tap * NT * X);
#pragma unroll
                            for (int k16 = 0; k16 < PL / 2; ++k16) {
                                mma(d, at + k16 * kA, bt + k16 * kB, acc);
                                acc = 1u;
                            }
                        }
                    }
                } else {
                    // 1x1 conv, or the fused 1x1 skip conv of a 3x3 block (centre tap of the window)
                    const uint32_t shift = is_skip ? uint32_t(p.pad * P + p.pad) : 0u;
                    for (int mb = mw; mb < p.MB; mb += MMA_WARPS) {
                        const uint32_t d = d0 + uint32_t(mb * NT * XA);
                        const uint32_t at = aaddr + uint32_t(mb * 128) + shift;
#pragma unroll
                        for (int k16 = 0; k16 < PL / 2; ++k16)
                            mma(d, at + k16 * kA, waddr + k16 * kB, (kc > 0 || k16 > 0) ? 1u : 0u);
                    }
                }
                umma_commit_elect(empty + stage);
                if (++stage == NS) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            umma_commit_elect(acc_full + buf);
            if (mw == 0 && lane == 0 && it == it_begin) trace(kTraceMma0);
            if (mw == 0 && lane == 0) tl(2, it - it_begin, 1);
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // =========================== TMA: activation windows + weights ============================
        if (warp == WARP_TMA && lane == 0 && it_begin < it_end) {
            int stage = 0;
            uint32_t phase = 0;
            const int planes_main = p.Cin / 8, planes_skip = (p.S0 + p.S1) / 8;
            const uint32_t a_bytes = uint32_t(X * PL) * uint32_t(NQ) * 16u;
            for (int it = it_begin; it < it_end; ++it) {
                const Item I = decode_item(p, it);
                tl(0, it - it_begin, 0);
                for (int kc = 0; kc < n_chunks; ++kc) {
                    if (kc == n_chunks - 1) tl(0, it - it_begin, 1);
                    const bool is_skip = kc >= p.n_main;
                    const int cbase = is_skip ? (kc - p.n_main) * KC : kc * KC;
                    const int CA = is_skip ? p.S0 : p.C0;
                    const bool first = cbase < CA;
                    const CUtensorMap *map = &P_.map[(is_skip ? 2 : 0) + (first ? 0 : 1)];
                    const int g0 = ((first ? cbase : cbase - CA) >> 3) * X;  // first plane of the chunk in the tensor
                    uint32_t bytes = a_bytes;
                    const __nv_bfloat16 *wsrc = nullptr;
                    uint32_t wbytes = 0;
                    if (!p.resident) {
                        wbytes = uint32_t(PL * (is_skip ? 1 : p.taps) * NT * X) * 16;
                        wsrc = is_skip ? p.skip_w + (size_t(I.cc) * planes_skip + size_t(kc - p.n_main) * PL) * NT * 8 * X
                                       : p.weight + (size_t(I.cc) * planes_main + size_t(kc) * PL) * p.taps * NT * 8 * X;
                        bytes += wbytes;
                    }
                    mbar_wait<128>(empty + stage, phase ^ 1u);
                    if (p.stride2) {
                        // four boxes, one per (row, column) parity, each traversing the input with element
                        // stride 2: {1 position, P sub-columns, RW sub-rows, PL planes, 1 sample}
                        mbar_expect_tx(raw_full + stage, 4 * bytes - 3 * wbytes);
#pragma unroll
                        for (int par = 0; par < 4; ++par)
                            tma_load_5d(sA + size_t(stage) * p.a_stage + size_t(par) * p.blk16 * 16, map, 0, 2 * (I.x0 - 1) + (par & 1),
                                        2 * (I.y0 - 1) + (par >> 1), g0, I.b, raw_full + stage);
                    } else {
                        mbar_expect_tx(raw_full + stage, bytes);
                        // box {2P x 8-byte elements, RW rows, PL planes, 1 sample}; x is counted in 8-byte units
                        tma_load_4d(sA + size_t(stage) * p.a_stage, map, 2 * (I.x0 - p.pad), I.y0 - p.pad, g0, I.b, raw_full + stage);
                    }
                    if (wbytes) bulk_g2s(sW + size_t(stage) * p.w_stage, wsrc, wbytes, raw_full + stage);
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, uint32_t(p.tmem_cols));
    }
    if (tid == 0) trace(kTraceEnd), cta_stamp(1);
}

// ---- host-side configuration --------------------------------------------------------------------
struct TmCfg {
    int PL, R, RW, NQ, Wt, P, MB, NT, n_cc, NS, resident, acc2, tmem_cols, tiles_x, tiles, n_main, n_skip, n_items, grid, ips, slots;
    uint32_t a_stage, w_stage, w_main_bytes, w_skip_bytes, magicP;
    size_t smem;
};

constexpr size_t kTmSmemBudget = 224 * 1024;  // of the 227 KB a CTA may opt in to
constexpr size_t kTmResidentMax = 80 * 1024;  // weights kept in smem for the whole launch when they fit
// one CTA per SM: the persistent grid and the statistics-slot layout follow the device the library runs on (148 on a
// B200; the no-GPU planning paths -- dry-run engines, host tests -- assume that)
int tm_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else n = 148;
        (void)cudaGetLastError();
    }
    return n;
}

int tm_nt(int Cout, int taps, int x3) { return tc_nt(Cout, taps, x3); }

// x3: fp16x2 operands -- a stage holds 2*PL planes and the weights are twice as many rows (conv_tma_kernel<..., X3>)
bool tm_configure_pl(int B, int H, int W, int C0, int C1, int S0, int S1, int Cout, int ksize, int stride, int up, int x3, int force_pl,
                     size_t res_max, TmCfg &best);
bool tm_configure_b(int B, int H, int W, int C0, int C1, int S0, int S1, int Cout, int ksize, int stride, int up, int x3, TmCfg &best);

// tile_batch > 0: choose the tile for that batch size, then lay the REAL batch out on it (items, grid, statistics slots)
bool tm_configure(int B, int tile_batch, int H, int W, int C0, int C1, int S0, int S1, int Cout, int ksize, int stride, int up, int x3,
                  TmCfg &best) {
    if (tile_batch <= 0 || tile_batch == B) return tm_configure_b(B, H, W, C0, C1, S0, S1, Cout, ksize, stride, up, x3, best);
    if (!tm_configure_b(tile_batch, H, W, C0, C1, S0, S1, Cout, ksize, stride, up, x3, best)) return false;
    const long long items = (long long)B * best.tiles * best.n_cc;
    const int sms = tm_num_sms();
    best.n_items = int(items);
    best.grid = int(items < sms ? items : sms);
    best.ips = best.tiles * best.n_cc;
    best.slots = 1;
    for (int b = 0; b < B; ++b) {
        const int c_first = int((((long long)b * best.ips + 1) * best.grid - 1) / best.n_items);
        const int c_last = int(((long long)(b + 1) * best.ips * best.grid - 1) / best.n_items);
        if (c_last - c_first + 1 > best.slots) best.slots = c_last - c_first + 1;
    }
    return true;
}

// Weights stay in shared memory for the whole launch when they are small (<= 80 KB).  CCDM_TMA_RESMAX=<KB> raises the
// ceiling for layers that still fit next to >= 3 activation stages (A/B runs).  MEASURED (round 2, 176 KB: the 64-channel
// convs of the 32x32 level and 96->32 @64x64 become resident, at the price of 3-4-row tiles): LIDC exact 4.12 vs 4.07 ms per
// reverse step, Cityscapes 4.97 vs 4.94 -- the re-streamed weights come out of L2 and were not what bounds those launches
// (the tensor pipe is: 60 % busy fetching A operands), the smaller tiles cost more halo.  Off by default.
bool tm_configure_b(int B, int H, int W, int C0, int C1, int S0, int S1, int Cout, int ksize, int stride, int up, int x3, TmCfg &best) {
    static const size_t env_res = getenv("CCDM_TMA_RESMAX") ? size_t(atoi(getenv("CCDM_TMA_RESMAX"))) * 1024 : kTmResidentMax;
    TmCfg big;
    if (env_res > kTmResidentMax && tm_configure_pl(B, H, W, C0, C1, S0, S1, Cout, ksize, stride, up, x3, 0, env_res, big) && big.resident &&
        big.NS >= 3 && big.w_main_bytes + big.w_skip_bytes > kTmResidentMax) {
        best = big;
        return true;
    }
    if (tm_configure_pl(B, H, W, C0, C1, S0, S1, Cout, ksize, stride, up, x3, 0, kTmResidentMax, best)) return true;
    // the preferred K-chunk width does not fit (wide layers whose weight stage alone is > 100 KB): halve it
    return !x3 && tm_configure_pl(B, H, W, C0, C1, S0, S1, Cout, ksize, stride, up, x3, 2, kTmResidentMax, best);
}

bool tm_configure_pl(int B, int H, int W, int C0, int C1, int S0, int S1, int Cout, int ksize, int stride, int up, int x3, int force_pl,
                     size_t res_max, TmCfg &best) {
    const int X = x3 ? 2 : 1;
    const int kTmNumSMs = tm_num_sms();
    const int Cin = C0 + C1, Sk = S0 + S1;
    if (Cin <= 0 || (C0 % 16) || (C1 % 16) || (S0 % 16) || (S1 % 16)) return false;
    const int pad = ksize / 2, taps = up ? 16 : ksize * ksize;  // up: H, W are the LOW-resolution (tile space) size
    const int nsub = up ? 4 : 1;                                 // accumulator sets per M block (output parities)
    const int CoutP = (Cout + 15) / 16 * 16;
    TmCfg c{};
    c.NT = tm_nt(Cout, taps, x3);
    c.n_cc = CoutP / c.NT;
    const bool all32 = !(C0 % 32) && !(C1 % 32) && !(S0 % 32) && !(S1 % 32);
    c.PL = all32 ? 4 : 2;
    // tuning overrides (A/B runs): CCDM_TMA_PL = 2|4 planes per K chunk, CCDM_TMA_R = rows per tile (large images only)
    static const int env_pl = getenv("CCDM_TMA_PL") ? atoi(getenv("CCDM_TMA_PL")) : 0;
    static const int env_r = getenv("CCDM_TMA_R") ? atoi(getenv("CCDM_TMA_R")) : 0;
    if (env_pl == 2 && W >= 64) c.PL = 2;
    // a 32-channel input would be ONE K chunk per item at PL = 4: two chunks of 16 let the TMA, the transform and
    // the MMAs of one item overlap (measured 73.7 -> 63.5 us on 32->32 @128x128, B = 64)
    if (env_pl == 0 && Cin == 32 && W >= 64) c.PL = 2;
    // wider inputs on the mid-size levels (B*H*W <= 512k pixels, e.g. 64x64 at B = 64 or 128x256 at B = 8) also prefer
    // chunks of 16 channels: smaller stages allow taller tiles (less halo) -- measured 33.0 -> 27.6 us (64->32 @128x256, B = 8),
    // 43.4 -> 39.7 us (96->32 @64x64, B = 64); on the full-resolution level it is 1-5 % slower, so not there
    if (env_pl == 0 && W >= 64 && (long long)B * H * W <= 524288) c.PL = 2;
    if (x3) c.PL = 2;  // fp16x2: 16 channels = 4 planes per stage, the only instantiation
    if (force_pl) c.PL = force_pl;
    const int KC = 8 * c.PL;
    c.n_main = Cin / KC;
    c.n_skip = Sk / KC;
    const int n_chunks = c.n_main + c.n_skip;
    const bool s2 = stride == 2;  // H, W: OUTPUT size; the stage holds 4 parity sub-images of (R+1) x (Wt+1) positions
    // Tile width.  bf16: 64 columns (or the whole row).  fp16x2: 64 or 32 -- its accumulators (2 NT columns per M block) cap a
    // tile at 4 M blocks, and 15 x 32 pixels carry less halo than 7 x 64 (1.20 vs 1.33: every role's work scales with it;
    // measured 4.00 vs 4.09 ms per LIDC step); the cost model below picks.  CCDM_TMA_WT overrides (A/B runs).
    static const int env_wt = getenv("CCDM_TMA_WT") ? atoi(getenv("CCDM_TMA_WT")) : 0;
    int widths[2] = {W > 64 ? 64 : W, 0};
    int n_widths = 1;
    if (env_wt >= 16 && env_wt <= 64 && W >= 64 && W % env_wt == 0) widths[0] = env_wt;
    else if (x3 && W >= 64 && W % 32 == 0) widths[n_widths++] = 32;
    double best_cost = 1e300;
    bool found = false;
    for (int wi = 0; wi < n_widths; ++wi) {
    c.Wt = widths[wi];
    c.P = s2 ? c.Wt + 1 : c.Wt + 2 * pad;
    if (2 * c.P > 256) continue;  // TMA box limit (8-byte elements)
    c.magicP = uint32_t(((1u << 20) + c.P - 1) / c.P);
    c.tiles_x = (W + c.Wt - 1) / c.Wt;
    c.w_main_bytes = uint32_t(size_t(Cin) * taps * c.NT * 2 * X);
    c.w_skip_bytes = uint32_t(size_t(Sk) * c.NT * 2 * X);
    const size_t w_total = size_t(c.w_main_bytes) + c.w_skip_bytes;
    c.resident = (c.n_cc == 1 && w_total <= res_max) ? 1 : 0;
    c.w_stage = c.resident ? 0u : uint32_t(c.PL * taps * c.NT * 16 * X);
    const size_t slack = (size_t(128 + 2 * pad * c.P + 2 * pad) * 16 + 127) & ~size_t(127);
    const size_t fixed = sizeof(float) * (2 * size_t(Cin) + c.NT + size_t(TM_EPI_WARPS) * CoutP * 2 * ((x3 && up) ? 2 : 1)) + (3 * MAX_STAGES + 5) * 8 + 64 + 16 * size_t(Cin) + 16 + slack +
                         (c.resident ? w_total : 0) + 1024;
    for (int R = 1; R <= H && 2 * (R + 2 * pad) <= 256; ++R) {
        const int MB = (R * c.P + 127) / 128;
        const int XA = (x3 && !up) ? 2 : 1;  // accumulator columns per (M block, parity) in units of NT
        if (nsub * MB * c.NT * XA > 512) break;
        const int RW = s2 ? R + 1 : R + 2 * pad, NQ = RW * c.P;
        bool magic_ok = true;
        for (int q = 0; q < NQ + 256; ++q)
            if (int((uint32_t(q) * c.magicP) >> 20) != q / c.P) magic_ok = false;
        if (!magic_ok) continue;
        const size_t blk = (size_t(X * c.PL) * NQ * 16 + 127) & ~size_t(127);  // one TMA box (128-byte aligned destination)
        const size_t a_stage = (s2 ? 4 : 1) * blk;
        if (fixed + 2 * (a_stage + c.w_stage) > kTmSmemBudget) break;
        int NS = int((kTmSmemBudget - fixed) / (a_stage + c.w_stage));
        NS = NS > MAX_STAGES ? MAX_STAGES : NS;
        static const int env_maxns = getenv("CCDM_TMA_MAXNS") ? atoi(getenv("CCDM_TMA_MAXNS")) : 0;  // tuning override
        if (env_maxns >= 2 && NS > env_maxns && H * W <= 1024) NS = env_maxns;  // small maps only: leave room for a co-resident CTA
        const int tiles = ((H + R - 1) / R) * c.tiles_x;
        const long long items = (long long)B * tiles * c.n_cc;
        const int grid = int(items < kTmNumSMs ? items : kTmNumSMs);
        const long long per_cta = (items + grid - 1) / grid;
        // transform work per item (window positions x channels, skip chunks are only copied) + M-block padding
        // + a fixed per-item hand-off cost
        const double item_cost = double(NQ) * (Cin + 0.25 * Sk) + double(MB * 128) * (0.15 * X * (Cin + Sk)) + 3000.0;
        double cost = double(per_cta) * item_cost;
        if (NS < 3) cost *= 1.5;
        else if (NS < 4) cost *= 1.1;
        // single accumulator buffer: the epilogue is not overlapped with the next item's MMAs -- costly for the upsampling convs,
        // whose epilogue drains four parity accumulators per M block (measured: ~5.7 us per item of serialised hand-offs)
        if (2 * nsub * MB * c.NT * XA > 512) cost *= (up && x3) ? 1.6 : 1.15;
        // fp16x2, full-resolution level, items of two K chunks and no skip chunks (32->32, 32->K): measured per op class, the
        // 64-wide tile wins there (LIDC 130 -> 115 us, 99 -> 90 us; Cityscapes 111 -> 108) although it loses on every other class
        // of that level (+9 .. +18 us) and on two-chunk items of the smaller levels (36 -> 43 us at 128x256, B = 8)
        if (x3 && !up && !s2 && n_chunks == 2 && Sk == 0 && (long long)B * H * W >= (1ll << 20) && c.Wt == 32) cost *= 1.3;
        if (env_r > 0 && W >= 64 && H >= 64) cost = (R == env_r) ? 0.0 : 1e290;
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
            best.R = R; best.RW = RW; best.NQ = NQ; best.MB = MB; best.NS = NS; best.a_stage = uint32_t(a_stage);
            best.acc2 = 2 * nsub * MB * c.NT * XA <= 512;
            best.tiles = tiles; best.n_items = int(items); best.grid = grid;
            int cols = 32;
            while (cols < (best.acc2 ? 2 : 1) * nsub * MB * c.NT * XA) cols *= 2;
            best.tmem_cols = cols;
            best.smem = fixed + size_t(NS) * (a_stage + c.w_stage);
            found = true;
        }
    }
    }  // tile widths
    if (found) {
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

bool tm_configure_op(const ccdm_op &op, TmCfg &c) {
    const int x3 = op.dtype == CCDM_DT_F16X2;
    if (op.upsample) return tm_configure(op.B, op.tile_batch, op.Hin, op.Win, op.C0, op.C1, op.S0, op.S1, op.Cout, op.ksize, op.stride, 1, x3, c);
    return tm_configure(op.B, op.tile_batch, op.Hout, op.Wout, op.C0, op.C1, op.S0, op.S1, op.Cout, op.ksize, op.stride, 0, x3, c);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// Tensor map of a plane-major bf16 activation [B][C/8][H][W][8], viewed as 8-byte elements so that a
// window row is ONE contiguous run of the innermost dimension: dims {2W, H, C/8, B}, box {2P, RW, PL, 1}.
// (fp16x2 tensors have two planes per 8-channel group: the callers pass C = 2 x channels and PL = 2 x groups per chunk)
int make_map(CUtensorMap *m, const void *base, int B, int C, int H, int W, int P, int RW, int PL) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) CCDM_FAIL(-5, "conv_tma: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {cuuint64_t(2) * W, cuuint64_t(H), cuuint64_t(C / 8), cuuint64_t(B)};
    const cuuint64_t strides[3] = {cuuint64_t(W) * 16, cuuint64_t(H) * W * 16, cuuint64_t(C / 8) * H * W * 16};
    const cuuint32_t box[4] = {cuuint32_t(2 * P), cuuint32_t(RW), cuuint32_t(PL), 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CCDM_FAIL(-5, "conv_tma: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box %dx%dx%d", int(r), B, C, H, W, P, RW, PL);
    return 0;
}

// Stride-2 variant: rank 5 {8-byte halves of a position, W, H, C/8, B} with element strides {1, 2, 2, 1, 1}: one box
// gathers every other column of every other row, i.e. one (row, column) parity sub-image of the window.
int make_map_s2(CUtensorMap *m, const void *base, int B, int C, int H, int W, int P, int RW, int PL) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) CCDM_FAIL(-5, "conv_tma: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[5] = {2, cuuint64_t(W), cuuint64_t(H), cuuint64_t(C / 8), cuuint64_t(B)};
    const cuuint64_t strides[4] = {16, cuuint64_t(W) * 16, cuuint64_t(H) * W * 16, cuuint64_t(C / 8) * H * W * 16};
    const cuuint32_t box[5] = {2u, cuuint32_t(2 * P), cuuint32_t(2 * RW), cuuint32_t(PL), 1u};
    const cuuint32_t estr[5] = {1u, 2u, 2u, 1u, 1u};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CCDM_FAIL(-5, "conv_tma: cuTensorMapEncodeTiled (stride 2) failed (%d) for [%d,%d,%d,%d] box %dx%dx%d", int(r), B, C, H, W, P, RW, PL);
    return 0;
}

}  // namespace

int conv_tma_read_trace(unsigned long long *out, int n) {
    unsigned long long tmp[16 + 64 + 320];
    if (cudaMemcpyFromSymbol(tmp, g_trace, 16 * sizeof(unsigned long long)) != cudaSuccess) return 0;
    if (cudaMemcpyFromSymbol(tmp + 16, g_tl, 64 * sizeof(unsigned long long)) != cudaSuccess) return 0;
    if (cudaMemcpyFromSymbol(tmp + 80, g_cta, 320 * sizeof(unsigned long long)) != cudaSuccess) return 0;
    const int m = n < 400 ? n : 400;  // [0,16): milestones; [16,80): timeline [role][item][begin/end]; [80,400): CTA [start|end][160]
    for (int i = 0; i < m; ++i) out[i] = tmp[i];
    return m;
}

int conv_tc_nt(int Cout, int taps, int x3) { return tm_nt(Cout, taps, x3); }

bool conv_tma_supported(const ccdm_op &op) {
    if ((op.dtype != CCDM_DT_BF16 && op.dtype != CCDM_DT_F16X2) || op.src_kind != 0) return false;
    const bool x3 = op.dtype == CCDM_DT_F16X2;
    if (x3 && (op.res || (op.out_dtype != CCDM_DT_F16X2 && op.out_dtype != CCDM_DT_F32))) return false;
    if (!x3 && op.out_dtype == CCDM_DT_F16X2) return false;
    if (op.ksize != 1 && op.ksize != 3) return false;
    if (op.upsample) {  // Upsample (unet.py:106-116): nearest x2 + 3x3, no norm, no skip, single source
        if (op.ksize != 3 || op.stride != 1 || op.gn || op.silu || op.S0 || op.C1 || op.res) return false;
        if (op.Hout != 2 * op.Hin || op.Wout != 2 * op.Win || (op.Cout % 16) || op.out_dtype != op.dtype) return false;
        TmCfg cu;
        return tm_configure_op(op, cu);
    }
    if ((op.C0 % 16) || (op.C1 % 16) || (op.S0 % 16) || (op.S1 % 16)) return false;
    if (op.out_dtype != CCDM_DT_F32 && (op.Cout % 16)) return false;
    if (op.stride == 2) {  // Downsample (unet.py:136-139): 3x3, no norm, no skip, single source
        if (op.ksize != 3 || op.gn || op.silu || op.S0 || op.C1) return false;
        if (op.Hout != (op.Hin + 1) / 2 || op.Wout != (op.Win + 1) / 2) return false;
    } else if (op.stride != 1 || op.Hin != op.Hout || op.Win != op.Wout) return false;
    TmCfg c;
    return tm_configure_op(op, c);
}

// {PL, R, Wt, MB, NQ, NT, n_cc, NS, resident, acc2, tmem_cols, tiles, n_items, grid, smem bytes, n_chunks}
int conv_tma_config(const ccdm_op &op, int32_t *out) {
    TmCfg c;
    if (!conv_tma_supported(op) || !tm_configure_op(op, c)) return -1;
    const int32_t v[16] = {c.PL, c.R, c.Wt, c.MB, c.NQ, c.NT, c.n_cc, c.NS, c.resident, c.acc2, c.tmem_cols, c.tiles, c.n_items, c.grid,
                           int32_t(c.smem), c.n_main + c.n_skip};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    return 0;
}

// {slots, items per sample, items, grid, row length} of the partial statistics this op's epilogue writes
int conv_tma_stat_layout(const ccdm_op &op, int32_t *out5) {
    TmCfg c;
    if (!conv_tma_supported(op) || !tm_configure_op(op, c)) return -1;
    out5[0] = c.slots; out5[1] = c.ips; out5[2] = c.n_items; out5[3] = c.grid; out5[4] = (op.Cout + 15) / 16 * 16;
    return 0;
}

size_t conv_tma_part_floats(const ccdm_op &op) {
    TmCfg c;
    if (!tm_configure_op(op, c)) return 0;
    return size_t(op.B) * c.slots * ((op.Cout + 15) / 16 * 16) * 2 * (op.dtype == CCDM_DT_F16X2 ? 2 : 1);  // fp16x2: rows of doubles
}

// fp16x2 instruction descriptor / bf16: cute::UMMA::InstrDescriptor -- D = f32 (bit 4), A and B formats at bits 7 and 10
// (0 = f16, 1 = bf16), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
static uint32_t tm_idesc(int NT, bool f16) { return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | (uint32_t(NT >> 3) << 17) | (uint32_t(128 >> 4) << 24); }

int launch_conv_tma(const ccdm_op &op, cudaStream_t s) {
    TmCfg c;
    if (!conv_tma_supported(op) || !tm_configure_op(op, c)) CCDM_FAIL(-3, "conv_tma: unsupported configuration");
    TmP P{};
    WsP &p = P.w;
    p.src0 = (const __nv_bfloat16 *)op.src0; p.src1 = (const __nv_bfloat16 *)op.src1;
    p.stat0 = (const double *)op.stat0; p.stat1 = (const double *)op.stat1;
    p.gamma = (const float *)op.gamma; p.beta = (const float *)op.beta;
    p.weight = (const __nv_bfloat16 *)op.weight; p.bias = (const float *)op.bias; p.emb = (const float *)op.emb;
    p.skip0 = (const __nv_bfloat16 *)op.skip0; p.skip1 = (const __nv_bfloat16 *)op.skip1;
    p.skip_w = (const __nv_bfloat16 *)op.skip_w; p.res = (const __nv_bfloat16 *)op.res;  // (fp16 data in fp16x2 mode: the kernel only moves 16-byte rows)
    p.out = (void *)op.out; p.ostat = (double *)op.ostat; p.part = (float *)op.part; p.ticket = (unsigned int *)op.ticket;
    p.steps = (const ccdm_step_entry *)op.steps; p.step_ptr = (const int *)op.step_ptr;
    p.B = op.B; p.Hin = op.Hin; p.Win = op.Win;
    p.H = op.upsample ? op.Hin : op.Hout; p.W = op.upsample ? op.Win : op.Wout;  // tile space (low resolution when upsampling)
    p.C0 = op.C0; p.C1 = op.C1; p.Cin = op.C0 + op.C1; p.Cout = op.Cout; p.CoutP = (op.Cout + 15) / 16 * 16;
    p.NT = c.NT; p.n_cc = c.n_cc;
#ifdef CCDM_SILU_EXPERIMENTS
    static const int env_silu = getenv("CCDM_SILU_MODE") ? atoi(getenv("CCDM_SILU_MODE")) : 1;
#else
    static const int env_silu = 1;
#endif
    p.upsample = op.upsample; p.nsub = op.upsample ? 4 : 1; p.gn = op.gn; p.silu = op.silu ? (env_silu == 2 || env_silu == 3 ? env_silu : 1) : 0; p.S0 = op.S0; p.S1 = op.S1;
    p.emb_off = op.emb_off; p.emb_cols = op.emb_cols; p.emb_bstride = op.emb_bstride;
    p.out_f32 = op.out_dtype == CCDM_DT_F32;
    if (p.out_f32 && (reinterpret_cast<uintptr_t>(op.out) & 15)) CCDM_FAIL(-2, "conv: fp32 output must be 16-byte aligned");
    p.R = c.R; p.Wt = c.Wt; p.P = c.P; p.MB = c.MB; p.WN = c.NQ; p.tiles_x = c.tiles_x; p.tiles = c.tiles;
    p.taps = op.upsample ? 16 : op.ksize * op.ksize; p.pad = op.ksize / 2;
    p.n_main = c.n_main; p.n_skip = c.n_skip; p.NS = c.NS; p.resident = c.resident; p.acc2 = c.acc2;
    p.tmem_cols = c.tmem_cols; p.n_items = c.n_items; p.ips = c.ips; p.slots = c.slots;
    p.RW = c.RW; p.NQ = c.NQ; p.xf = (op.gn || op.silu) ? 1 : 0; p.stride2 = op.stride == 2;
    for (int i = 0; i < 2; ++i) {
        p.st_slots[i] = op.st_slots[i]; p.st_ips[i] = op.st_ips[i]; p.st_items[i] = op.st_items[i];
        p.st_grid[i] = op.st_grid[i]; p.st_rows[i] = op.st_rows[i];
    }
    p.blk16 = uint32_t(((size_t((op.dtype == CCDM_DT_F16X2 ? 2 : 1) * c.PL) * c.NQ * 16 + 127) & ~size_t(127)) >> 4);
    p.a_stage = c.a_stage; p.w_stage = c.w_stage; p.w_main_bytes = c.w_main_bytes; p.w_skip_bytes = c.w_skip_bytes;
    p.magicP = c.magicP;
    const bool x3 = op.dtype == CCDM_DT_F16X2;
    const int X = x3 ? 2 : 1;
    p.idesc = tm_idesc(c.NT, x3);
    p.idesc2 = tm_idesc(2 * c.NT, x3);
#ifdef CCDM_ABLATE
    static const int env_dbg = getenv("CCDM_ABLATE") ? atoi(getenv("CCDM_ABLATE")) : 0;
    p.dbg = env_dbg;
#endif
    p.x3 = x3 ? 1 : 0;
    p.descale = x3 ? float(ldexp(1.0, -op.acc_shift)) : 1.0f;
    if (x3 && (op.acc_shift < CCDM_F16X2_SCALE_LOG2 || op.acc_shift > 40)) CCDM_FAIL(-2, "conv_tma: fp16x2 op without a valid acc_shift (%d)", op.acc_shift);

    if (op.gn && (!op.stat0 || (op.C1 && !op.stat1) || !op.gamma || !op.beta)) CCDM_FAIL(-2, "conv_tma: gn without stats/affine");
    if (op.gn && op.gn_cpg <= 0 && (p.Cin % kGnGroups)) CCDM_FAIL(-2, "conv_tma: GroupNorm needs Cin %% 32 == 0");
    if (op.gn_cpg < 0 || op.gn_off < 0) CCDM_FAIL(-2, "conv_tma: gn_cpg / gn_off must not be negative");
    p.gn_cpg = op.gn_cpg; p.gn_off = op.gn_off;
    if (op.ostat && (!op.part || !op.ticket)) CCDM_FAIL(-2, "conv_tma: ostat without scratch");
    for (int i = 0; i < 2; ++i)
        if (op.st_slots[i] > 0 && (op.st_ips[i] <= 0 || op.st_items[i] <= 0 || op.st_grid[i] <= 0 || op.st_rows[i] <= 0 ||
                                   (double(op.B) * op.st_ips[i] + 1.0) * op.st_grid[i] >= 2147483648.0))
            CCDM_FAIL(-2, "conv_tma: incomplete deferred-fold layout of stat%d", i);
    if (op.emb && (!op.steps || !op.step_ptr || op.emb_off < 0)) CCDM_FAIL(-2, "conv_tma: emb without step table");
    if (op.S0 > 0 && (!op.skip0 || !op.skip_w)) CCDM_FAIL(-2, "conv_tma: bad skip configuration");
    if (!op.src0 || (op.C1 && !op.src1) || (op.S1 && !op.skip1)) CCDM_FAIL(-2, "conv_tma: missing source tensor");

    const int mapH = op.upsample ? op.Hin : op.Hout, mapW = op.upsample ? op.Win : op.Wout;
    int rc = op.stride == 2 ? make_map_s2(&P.map[0], (const void *)op.src0, op.B, X * op.C0, op.Hin, op.Win, c.P, c.RW, X * c.PL)
                            : make_map(&P.map[0], (const void *)op.src0, op.B, X * op.C0, mapH, mapW, c.P, c.RW, X * c.PL);
    if (rc == 0 && op.C1) rc = make_map(&P.map[1], (const void *)op.src1, op.B, X * op.C1, op.Hout, op.Wout, c.P, c.RW, X * c.PL);
    if (rc == 0 && op.S0) rc = make_map(&P.map[2], (const void *)op.skip0, op.B, X * op.S0, op.Hout, op.Wout, c.P, c.RW, X * c.PL);
    if (rc == 0 && op.S1) rc = make_map(&P.map[3], (const void *)op.skip1, op.B, X * op.S1, op.Hout, op.Wout, c.P, c.RW, X * c.PL);
    if (rc != 0) return rc;

    static bool attr_done = false;
    if (!attr_done) {
        auto opt_in = [](auto kernel) { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTmSmemBudget)); };
        CCDM_CUDA(opt_in(conv_tma_kernel<4, false, true, 0>)); CCDM_CUDA(opt_in(conv_tma_kernel<2, false, true, 0>));
        CCDM_CUDA(opt_in(conv_tma_kernel<4, false, true, 1>)); CCDM_CUDA(opt_in(conv_tma_kernel<2, false, true, 1>));
        CCDM_CUDA(opt_in(conv_tma_kernel<4, false, true, 2>)); CCDM_CUDA(opt_in(conv_tma_kernel<2, false, true, 2>));
        CCDM_CUDA(opt_in(conv_tma_kernel<4, false, false, 3>)); CCDM_CUDA(opt_in(conv_tma_kernel<2, false, false, 3>));
        CCDM_CUDA(opt_in(conv_tma_kernel<4, true, true, 0>)); CCDM_CUDA(opt_in(conv_tma_kernel<2, true, true, 0>));
        CCDM_CUDA(opt_in(conv_tma_kernel<2, false, true, 0, true>)); CCDM_CUDA(opt_in(conv_tma_kernel<2, false, true, 1, true>));
        CCDM_CUDA(opt_in(conv_tma_kernel<2, false, true, 2, true>)); CCDM_CUDA(opt_in(conv_tma_kernel<2, false, false, 3, true>));
        CCDM_CUDA(opt_in(conv_tma_kernel<2, true, true, 0, true>));
        attr_done = true;
    }
    const bool lean = op.out_dtype != CCDM_DT_F32 && !op.res;  // no fp32 store, no epilogue residual: the lean epilogue
    const int xf_kind = !(op.gn || op.silu) ? 0 : (op.silu ? 1 : 2);
    const bool pl4 = c.PL == 4;
    void (*kern)(TmP) = nullptr;
    if (x3) {
        if (c.PL != 2) CCDM_FAIL(-3, "conv_tma: fp16x2 kernels are instantiated for 16-channel K chunks only");
        if (op.upsample) kern = conv_tma_kernel<2, true, true, 0, true>;
        else if (!lean) kern = conv_tma_kernel<2, false, false, 3, true>;
        else if (xf_kind == 0) kern = conv_tma_kernel<2, false, true, 0, true>;
        else if (xf_kind == 1) kern = conv_tma_kernel<2, false, true, 1, true>;
        else kern = conv_tma_kernel<2, false, true, 2, true>;
    } else if (op.upsample) kern = pl4 ? conv_tma_kernel<4, true, true, 0> : conv_tma_kernel<2, true, true, 0>;
    else if (!lean) kern = pl4 ? conv_tma_kernel<4, false, false, 3> : conv_tma_kernel<2, false, false, 3>;
    else if (xf_kind == 0) kern = pl4 ? conv_tma_kernel<4, false, true, 0> : conv_tma_kernel<2, false, true, 0>;
    else if (xf_kind == 1) kern = pl4 ? conv_tma_kernel<4, false, true, 1> : conv_tma_kernel<2, false, true, 1>;
    else kern = pl4 ? conv_tma_kernel<4, false, true, 2> : conv_tma_kernel<2, false, true, 2>;
    CCDM_CUDA(launch_pdl(kern, dim3(c.grid), dim3(TM_THREADS), c.smem, s, P));
    CCDM_LAUNCH_CHECK("conv_tma_kernel");
    return 0;
}

}  // namespace ccdm
