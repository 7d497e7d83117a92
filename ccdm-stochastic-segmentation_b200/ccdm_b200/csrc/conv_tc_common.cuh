// Pieces of the tcgen05 conv kernel (conv_tma.cu) that do not depend on how its operands are staged: the launch
// parameter block, the work-item decoding, the GroupNorm fold and the epilogue role (TMEM -> registers -> +bias
// +embedding +residual -> plane-major bf16 / fp16x2 / NHWC fp32 store + GroupNorm statistics of the output).
#pragma once

#include <type_traits>

#include "tc_common.cuh"

namespace ccdm {
namespace {

constexpr int EPI_WARPS = 4;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int MAX_STAGES = 8;
constexpr int CGW = 16;  // accumulator columns per tcgen05.ld

// Output channels per work item (the N of the MMAs) = the packing unit of the weights, a function of (Cout, taps) only so
// that one packed tensor serves every shape.  Up to 128 output channels, and every 3x3 conv: the largest divisor <= 64
// (narrow items: more CTAs busy on small maps, and a weight stage of 9 or 16 taps stays small).  Wider 1x1 layers (the
// q/k/v convs, Cout = 3C): the largest divisor <= 192, so a sample is 2 items instead of 6 and the input is normalised
// twice instead of six times.
// fp16x2 mode: N <= 128, because an MMA of that mode covers 2 NT accumulator columns (hi and lo rows of B at once).
inline int tc_nt(int Cout, int taps, int x3) {
    const int CoutP = (Cout + 15) / 16 * 16;
    for (int nt = (CoutP > 128 && taps == 1) ? (x3 ? 128 : 192) : 64; nt >= 16; nt -= 16)
        if (CoutP % nt == 0) return nt;
    return 16;
}

struct WsP {
    const __nv_bfloat16 *src0, *src1;
    const double *stat0, *stat1;
    const float *gamma, *beta;
    const __nv_bfloat16 *weight;  // [cc][Cin/8][tap][NT][8]
    const float *bias, *emb;
    const __nv_bfloat16 *skip0, *skip1;
    const __nv_bfloat16 *skip_w;  // [cc][S/8][NT][8]
    const __nv_bfloat16 *res;
    void *out;
    double *ostat;
    float *part;
    unsigned int *ticket;
    const ccdm_step_entry *steps;
    const int *step_ptr;
    int B, Hin, Win, H, W;  // H, W: conv-input == output space (after the optional x2)
    int C0, C1, Cin, Cout, CoutP, NT, n_cc;
    int upsample, gn, silu, S0, S1, emb_off, emb_cols, emb_bstride, out_f32;
    int R, Wt, P, MB, WN, tiles_x, tiles, taps, pad;
    int n_main, n_skip, NS, resident, acc2, tmem_cols, n_items;
    int ips, slots;  // items per sample; statistics slots per sample (CTAs whose item range can touch one sample)
    int st_slots[2], st_ips[2], st_items[2], st_grid[2], st_rows[2];  // layout of stat0 / stat1 (ccdm_op::st_*)
    int nsub;        // accumulator sets per M block: 1, or 4 output parities of a fused nearest-x2 + conv3x3 (conv_tma)
    int stride2;     // conv_tma: 3x3 stride-2 conv (Downsample): the stage holds the 4 (row, column) parity sub-images
    uint32_t blk16;  // conv_tma, stride 2: size of one parity sub-image block in 16-byte rows (128-byte aligned for TMA)
    int RW, NQ, xf;  // conv_tma: window rows, window positions (= plane stride in 16-byte rows), 1 if chunks are transformed in smem
    uint32_t a_stage, w_stage, w_main_bytes, w_skip_bytes, magicP;
    uint32_t idesc;
    uint32_t idesc2; // x3: the same with N = 2 NT
    int x3;          // fp16x2 storage (CCDM_DT_F16X2): operands are (hi, lo) plane pairs, three MMAs per product
    float descale;   // x3: 2^-acc_shift, applied to the accumulator in the epilogue
    int gn_cpg, gn_off;  // GroupNorm group size / channel offset of channel 0 inside the normalised concatenation (ccdm_op::gn_cpg, gn_off)
    int dbg;         // CCDM_ABLATE builds only: bit 0 skip the MMAs, bit 1 skip the transform maths, bit 2 skip the epilogue's stores
};

struct Item {
    int b, tile, cc, y0, x0, co0;
};
__device__ __forceinline__ Item decode_item(const WsP &p, int it) {
    Item r;
    r.cc = it % p.n_cc;
    const int t = it / p.n_cc;
    r.tile = t % p.tiles;
    r.b = t / p.tiles;
    const int ty = r.tile / p.tiles_x, tx = r.tile - ty * p.tiles_x;
    r.y0 = ty * p.R;
    r.x0 = tx * p.Wt;
    r.co0 = r.cc * p.NT;
    return r;
}


// Epilogue role: warps 0..NEW-1 of the CTA, NEW = 4 or 8.  Warp w may only touch TMEM lanes
// 32*(w%4)..+31 (= accumulator rows 32*(w%4)+lane of every M block); with 8 warps the two warps of a lane
// quarter split the 16-column groups of the accumulator between them.  Named barrier 2 is private to
// these NEW*32 threads.
// GroupNorm scale / shift of every input channel of sample b into sAff[0..Cin) / sAff[Cin..2Cin) (the SiLU's
// 1/2 folded in), computed by `nthreads` threads (thread index `t`) that share named barrier `bar_id`.  The
// per-channel {sum, sum of squares} of source i are either folded already (double2 [B, C], st_slots[i] == 0) or
// are the fp32 partial rows its producer's epilogue wrote, one per CTA that touched the sample: folded here in row
// order, in double.  This runs at the head of every launch with nothing to overlap it, so it is arranged as ONE
// round of independent L2 loads per four partial rows: thread c sums the rows of channel c (sSum, [Cin] double2
// scratch), and after a barrier every thread adds up the cpg channels of its group from shared memory.  (The first
// version walked cpg x rows elements per thread, one L2 round trip per channel of the group: 3 us at cpg = 4.)
__device__ __forceinline__ void gn_build_affine(const WsP &p, int b, float *sAff, double2 *sSum, int t, int nthreads, int bar_id) {
    const int cpg = p.gn_cpg > 0 ? p.gn_cpg : p.Cin / kGnGroups;
    const double inv_n = 1.0 / (double(cpg) * double(p.Hin) * double(p.Win));
    const float half = (p.silu == 1 || p.silu == 2) ? 0.5f : 1.0f;  // tanh forms of SiLU work on x/2
    int n_rows[2] = {0, 0};  // partial rows of sample b per source (32-bit: the host checked (B*ips+1)*grid < 2^31)
#pragma unroll
    for (int i = 0; i < 2; ++i)
        if (p.st_slots[i] > 0) {
            const unsigned ips = unsigned(p.st_ips[i]), grid = unsigned(p.st_grid[i]), items = unsigned(p.st_items[i]);
            const unsigned c_first = ((unsigned(b) * ips + 1u) * grid - 1u) / items;
            const unsigned c_last = ((unsigned(b) + 1u) * ips * grid - 1u) / items;
            n_rows[i] = int(c_last - c_first) + 1;
        }
    for (int c = t; c < p.Cin; c += nthreads) {
        const int i = c < p.C0 ? 0 : 1;
        const int cs = i ? c - p.C0 : c;
        const double *st = i ? p.stat1 : p.stat0;
        const float g = p.gamma[c], be = p.beta[c];
        double s = 0.0, q = 0.0;
        if (p.st_slots[i] == 0) {
            const double2 e = __ldcg(reinterpret_cast<const double2 *>(st + (size_t(b) * (i ? p.C1 : p.C0) + cs) * 2));
            s = e.x;
            q = e.y;
        } else {
            const int rows = p.st_rows[i], n = n_rows[i];
            if (p.x3) {
                // fp16x2 producers accumulate their statistics in double from the first addition (see the epilogue): the
                // rows are double2 and their sum does not depend on how the producer's items were grouped
                const double2 *pp = reinterpret_cast<const double2 *>(st) + size_t(b) * p.st_slots[i] * rows + cs;
                for (int r0 = 0; r0 < n; r0 += 4) {
                    double2 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r0 + u < n ? r0 + u : r0;
                        v[u] = __ldcg(pp + size_t(r) * rows);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (r0 + u < n) {
                            s += v[u].x;
                            q += v[u].y;
                        }
                }
            } else {
            const float2 *pp = reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(st) + (size_t(b) * p.st_slots[i] * rows + cs) * 2);
            for (int r0 = 0; r0 < n; r0 += 4) {
                float2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + u < n ? r0 + u : r0;  // clamped: the load is unconditional, the value is masked
                    v[u] = __ldcg(pp + size_t(r) * rows);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (r0 + u < n) {
                        s += double(v[u].x);
                        q += double(v[u].y);
                    }
            }
            }
        }
        sSum[c] = make_double2(s, q);
        sAff[c] = g;
        sAff[p.Cin + c] = be;
    }
    named_bar_sync(bar_id, nthreads);
    for (int c = t; c < p.Cin; c += nthreads) {
        // channels of c's group that this op sees (a group cut by the op's channel range is incomplete: its channels carry
        // zero weights by construction, ccdm_op::gn_off)
        const int g0 = ((c + p.gn_off) / cpg) * cpg - p.gn_off;
        const int j0 = g0 < 0 ? 0 : g0, j1 = g0 + cpg < p.Cin ? g0 + cpg : p.Cin;
        double s = 0.0, q = 0.0;
        for (int j = j0; j < j1; ++j) {
            const double2 e = sSum[j];
            s += e.x;
            q += e.y;
        }
        const double mean = s * inv_n;
        double var = q * inv_n - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        if (p.x3) {
            // exact mode: fp32-grade 1/sqrt; the stored operand is 2^4 x, so only the scale carries the 2^-4
            const float rstd = float(1.0 / sqrt(var + double(kGnEps)));
            const float a = sAff[c] * rstd;
            sAff[c] = a * (1.0f / float(1 << CCDM_F16X2_SCALE_LOG2));
            sAff[p.Cin + c] = sAff[p.Cin + c] - float(mean) * a;
            continue;
        }
        const float rstd = rsqrtf(float(var) + kGnEps);
        const float a = sAff[c] * rstd;
        sAff[c] = half * a;
        sAff[p.Cin + c] = half * (sAff[p.Cin + c] - float(mean) * a);
    }
}

// optional milestone hook (conv_tma.cu defines CCDM_EPI_TRACE before including this header)
#ifndef CCDM_EPI_TRACE
#define CCDM_EPI_TRACE(slot)
#endif
#ifndef CCDM_EPI_TL
#define CCDM_EPI_TL(item, edge)
#endif

// LEAN: bf16 plane-major output without an epilogue residual (every conv of the bf16 chain but the output conv): the fp32
// NHWC store variants and the residual read are compiled out -- the role's code is half of the kernel, and its size costs
// (instruction fetch) even where it is not executed.
template <int NEW, int NSUB = 1, bool LEAN = false, bool X3 = false>
__device__ __forceinline__ void conv_epilogue_role(const WsP &p, float *sAdd, float *sRed, int *s_last, uint64_t *acc_full,
                                                   uint64_t *acc_empty, uint32_t tmem_base, int it_begin, int it_end) {
    constexpr int NTHR = NEW * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3, half = warp >> 2;
    constexpr int NHALF = NEW / 4;
    const int NT = p.NT, P = p.P;
    // GroupNorm statistics of the output: per-thread sums over an item -> warp transpose-reduce ->
    // per-warp running sums in shared memory (sAcc), flushed to global ONCE per (CTA, sample): a CTA's
    // items are contiguous, so this is one or two partial rows per CTA instead of one per item.
    // fp16x2 ("exact" mode): an item's partial sums (per thread over its rows, then across the warp) are fp32 in a fixed order
    // -- a function of the tile only -- and everything ABOVE the item is double: the per-warp running sums over the CTA's items,
    // the cross-warp sum, the partial rows, the consumer's fold.  Sums of the same fp32 item partials in double are exact to
    // 1e-16 whatever their grouping, so with a batch-independent tiling (ccdm_op::tile_batch) a sample's statistics -- and its
    // result -- do not depend on the batch split or the number of GPUs.  (Double from the first addition was measured: +30 % per
    // step -- FP64 issue rate.)  bf16 mode keeps fp32 throughout.
    using ST = typename std::conditional<X3, double, float>::type;
    // running sums: one row per warp (bf16, fp32) or per TMEM lane quarter (fp16x2, double: the two warps of a quarter own
    // different column groups, so they never touch the same entry -- and four rows of doubles are the bytes of eight of floats)
    // (upsampling convs: the two warps of a quarter split the output PARITIES of the same channels, so they keep a row each)
    constexpr int ACC_ROWS = (X3 && NSUB == 1) ? 4 : NEW;
    ST *sAcc = reinterpret_cast<ST *>(sRed);  // [ACC_ROWS][CoutP][2]
    ST *part = reinterpret_cast<ST *>(p.part);
    const int CoutP = p.CoutP;
    const int acc_row = ACC_ROWS == 4 ? (warp & 3) : warp;
    const bool want_stats = p.part != nullptr;  // with ostat: folded here (ticket); without: deferred to the consumer
    if (want_stats && warp < ACC_ROWS) {  // (the first use is behind the named barrier of the first item's bias load)
#pragma unroll 1
        for (int e = lane; e < CoutP * 2; e += 32) sAcc[warp * CoutP * 2 + e] = ST(0);
        __syncwarp();
    }
    auto flush_stats = [&](int b, int n_done) {
        named_bar_sync(2, NTHR);
        const int c_first = int((((long long)b * p.ips + 1) * gridDim.x - 1) / p.n_items);
        const int slot = int(blockIdx.x) - c_first;
#pragma unroll 1
        for (int e = tid; e < CoutP * 2; e += NTHR) {
            ST s = ST(0);
#pragma unroll
            for (int r = 0; r < ACC_ROWS; ++r) {
                s += sAcc[r * CoutP * 2 + e];
                sAcc[r * CoutP * 2 + e] = ST(0);
            }
            part[(size_t(b) * p.slots + slot) * CoutP * 2 + e] = s;
        }
        if (p.ostat == nullptr) {  // deferred fold: the partial row IS the result; the kernel boundary publishes it
            named_bar_sync(2, NTHR);
            return;
        }
        __threadfence();
        named_bar_sync(2, NTHR);
        if (tid == 0) {
            const unsigned int prev = atomicAdd(p.ticket + b, unsigned(n_done));
            *s_last = (prev + unsigned(n_done) == unsigned(p.ips));
        }
        named_bar_sync(2, NTHR);
        if (*s_last) {
            // every item of this sample is done somewhere on the chip: fold the per-CTA partial rows in
            // slot order, in double (the consumer's GroupNorm reads these sums) -- a fixed order, so the
            // result does not depend on which CTA finishes last
            __threadfence();
            const int c_last = int(((long long)(b + 1) * p.ips * gridDim.x - 1) / p.n_items);
            const int n_slots = c_last - c_first + 1;
#pragma unroll 1
            for (int e = tid; e < p.Cout * 2; e += NTHR) {
                const ST *pp = part + size_t(b) * p.slots * CoutP * 2 + e;
                double s = 0.0;
#pragma unroll 2
                for (int t = 0; t < n_slots; ++t) s += double(__ldcg(pp + size_t(t) * CoutP * 2));
                p.ostat[size_t(b) * p.Cout * 2 + e] = s;
            }
            if (tid == 0) p.ticket[b] = 0u;  // self-reset for the next launch
        }
        named_bar_sync(2, NTHR);  // s_last is reused by the next flush
    };

    // accumulator row of this thread inside an M block, as (window row, window column); M blocks advance
    // it by 128 positions = (d128r, d128c) with one carry
    const int j0 = quarter * 32 + lane;
    const int o0 = int((uint32_t(j0) * p.magicP) >> 20), c0 = j0 - o0 * P;
    const int d128r = int((128u * p.magicP) >> 20), d128c = 128 - d128r * P;
    // upsampling conv: the tile space (p.H x p.W) is the low-resolution grid, every accumulator row owns the 2x2
    // output pixels (2y+py, 2x+px), one accumulator set per parity
    constexpr int ups = NSUB == 4 ? 2 : 1, nsub = NSUB;  // compile-time: the common (NSUB = 1) path pays nothing for it
    const int Wo = p.W * ups;
    const size_t hw = size_t(p.H) * p.W * ups * ups;
    const int n_cg = NT / CGW;

    int acc_it = 0, cur_b = -1, cur_cc = -1, n_pending = 0;
    // (one trip past the end: the last sample's statistics are flushed by the same call site as the others -- the
    // flush is ~500 instructions, and this role's code size is what every launch pays for at its start)
    for (int it = it_begin; it <= it_end; ++it, ++acc_it) {
        const bool past_end = it == it_end;
        Item I;
        if (!past_end) I = decode_item(p, it);
        if ((past_end || I.b != cur_b) && n_pending > 0) {
            if (want_stats) flush_stats(cur_b, n_pending);
            n_pending = 0;
        }
        if (past_end) break;
        if (I.b != cur_b || I.cc != cur_cc) {
            named_bar_sync(2, NTHR);
            for (int c = tid; c < NT; c += NTHR) {
                float v = p.bias[I.co0 + c];
                if (p.emb != nullptr && I.co0 + c < p.Cout) {
                    const ccdm_step_entry &se = p.steps[*p.step_ptr];
                    v += p.emb[(size_t(se.emb_row) + size_t(I.b) * p.emb_bstride) * p.emb_cols + p.emb_off + I.co0 + c];
                }
                sAdd[c] = v;
            }
            named_bar_sync(2, NTHR);
            cur_b = I.b;
            cur_cc = I.cc;
        }
        const int buf = p.acc2 ? (acc_it & 1) : 0;
        const uint32_t aph = p.acc2 ? uint32_t((acc_it >> 1) & 1) : uint32_t(acc_it & 1);
        // fp16x2: 2 NT columns per (M block, parity): [0, NT) the hi*hi products, [NT, 2 NT) the small hi*lo + lo*hi terms
        // (upsampling convs keep ONE accumulator per parity -- four parities x two halves would not leave room to double-buffer)
        constexpr int XA = (X3 && NSUB == 1) ? 2 : 1;
        const uint32_t tbase = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * p.MB * NT * nsub * XA);
        const int ylim = min(p.R, p.H - I.y0), xlim = min(p.Wt, p.W - I.x0);  // rows / columns of real outputs
        bool waited = false;
        for (int cgs = half; cgs < n_cg * nsub; cgs += NHALF) {
            const int cg = cgs / nsub, sub = cgs - cg * nsub;  // channel group, output parity (py, px) = (sub >> 1, sub & 1)
            const size_t pix0 = size_t(I.y0 * ups + (sub >> 1)) * Wo + I.x0 * ups + (sub & 1);
            const int cobase = I.co0 + cg * CGW;
            float add[CGW];
#pragma unroll
            for (int i = 0; i < CGW; i += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(sAdd + cg * CGW + i);
                add[i] = t.x; add[i + 1] = t.y; add[i + 2] = t.z; add[i + 3] = t.w;
            }
            // plane-major bases of this thread's two 8-channel planes (bf16 in/out) or the NHWC fp32 row
            // (fp16x2: twice the planes -- the hi and lo plane of an 8-channel group are adjacent)
            const size_t plane0 = (size_t(I.b) * (p.Cout >> 3) + (cobase >> 3)) * (X3 ? 2 : 1) * hw + pix0;
            const __nv_bfloat16 *resb = p.res != nullptr ? p.res + plane0 * 8 : nullptr;
            __nv_bfloat16 *outb = reinterpret_cast<__nv_bfloat16 *>(p.out) + plane0 * 8;
            float *outf = reinterpret_cast<float *>(p.out) + (size_t(I.b) * hw + pix0) * p.Cout + cobase;
            float s1[CGW], s2[CGW];
#pragma unroll
            for (int i = 0; i < CGW; ++i) s1[i] = 0.f, s2[i] = 0.f;
            int o = o0, c = c0;
            if (!waited) {
                mbar_wait<X3 ? 512 : 256>(acc_full + buf, aph);  // the next accumulator is microseconds away
                tc_fence_after();
                waited = true;
                if (tid == 0) CCDM_EPI_TL(it - it_begin, 0);
            }
            // accumulator rows stream out of TMEM double-buffered: the load of row block mb+1 is in flight while
            // block mb is processed (tcgen05.ld is asynchronous until tcgen05.wait::ld)
            auto process = [&](const uint32_t (&raw)[CGW], int mb) {
                (void)mb;
                const bool valid = o < ylim && c < xlim;
                const int off = (o * Wo + c) * ups;
                c += d128c;  // next row block: 128 positions further in the flattened window
                o += d128r;
                if (c >= P) {
                    c -= P;
                    ++o;
                }
                float v[CGW];
#pragma unroll
                for (int i = 0; i < CGW; ++i) v[i] = __uint_as_float(raw[i]);
#ifdef CCDM_ABLATE
                if (p.dbg & 4) return;
#endif
                if (valid) {
#pragma unroll
                    for (int i = 0; i < CGW; ++i) v[i] = X3 ? fmaf(v[i], p.descale, add[i]) : v[i] + add[i];
                    if (!X3 && !LEAN && resb != nullptr) {  // identity residual read here (the engine's bf16 path folds it into the MMA instead)
#pragma unroll
                        for (int h2 = 0; h2 < CGW / 8; ++h2) {
                            const uint4 rr = ldg_nc16(resb + (size_t(h2) * hw + off) * 8);
                            const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float2 f = unpack_bf16(w4[i]);
                                v[h2 * 8 + 2 * i] += f.x;
                                v[h2 * 8 + 2 * i + 1] += f.y;
                            }
                        }
                    }
                    if (!LEAN && p.out_f32) {  // fp32 logits stay NHWC: Cout floats per pixel, written with the widest aligned store
                        float *op = outf + size_t(off) * p.Cout;
                        if ((p.Cout & 3) == 0) {  // K = 20: a pixel's row is 16-byte aligned
#pragma unroll
                            for (int i = 0; i < CGW; i += 4)
                                if (cobase + i < p.Cout) *reinterpret_cast<float4 *>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        } else if ((p.Cout & 1) == 0) {  // K = 2: one 8-byte store per pixel, contiguous across the warp
#pragma unroll
                            for (int i = 0; i < CGW; i += 2)
                                if (cobase + i < p.Cout) *reinterpret_cast<float2 *>(op + i) = make_float2(v[i], v[i + 1]);
                        } else {
#pragma unroll
                            for (int i = 0; i < CGW; ++i)
                                if (cobase + i < p.Cout) op[i] = v[i];
                        }
                    } else if (X3) {
#pragma unroll
                        for (int h2 = 0; h2 < CGW / 8; ++h2) {
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split_f16x2(v[h2 * 8 + 2 * i], v[h2 * 8 + 2 * i + 1], ph[i], pl[i]);
                            *reinterpret_cast<uint4 *>(outb + (size_t(2 * h2) * hw + off) * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4 *>(outb + (size_t(2 * h2 + 1) * hw + off) * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                    } else {
#pragma unroll
                        for (int h2 = 0; h2 < CGW / 8; ++h2) {
                            uint32_t pk[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) pk[i] = pack_bf16(v[h2 * 8 + 2 * i], v[h2 * 8 + 2 * i + 1]);
                            *reinterpret_cast<uint4 *>(outb + (size_t(h2) * hw + off) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    // statistics of the fp32 values (the bf16 rounding error is zero-mean: over >= 64 pixels
                    // per channel it changes the sums far below the GroupNorm epsilon)
#pragma unroll
                    for (int i = 0; i < CGW; ++i) {
                        s1[i] += v[i];
                        s2[i] = fmaf(v[i], v[i], s2[i]);
                    }
                }
            };
            const uint32_t tcol = tbase + uint32_t(sub * NT * XA + cg * CGW), tstep = uint32_t(nsub * NT * XA);
            uint32_t ra[CGW], rb[CGW];
            if constexpr (X3 && NSUB == 1) {
                // main + small accumulator columns of a row block, summed here in fp32 (round to nearest); the loads of block
                // mb+1 are in flight while block mb is processed
                tmem_ld16_issue(tcol, ra);
                tmem_ld16_issue(tcol + uint32_t(NT), rb);
                for (int mb = 0; mb < p.MB; ++mb) {
                    tmem_ld_wait16(ra);
                    tmem_ld_wait16(rb);
                    uint32_t sum[CGW];
#pragma unroll
                    for (int i = 0; i < CGW; ++i) sum[i] = __float_as_uint(__uint_as_float(ra[i]) + __uint_as_float(rb[i]));
                    if (mb + 1 < p.MB) {
                        tmem_ld16_issue(tcol + uint32_t(mb + 1) * tstep, ra);
                        tmem_ld16_issue(tcol + uint32_t(mb + 1) * tstep + uint32_t(NT), rb);
                    }
                    process(sum, mb);
                }
            } else {
            tmem_ld16_issue(tcol, ra);
            for (int mb = 0; mb < p.MB; mb += 2) {
                tmem_ld_wait16(ra);
                if (mb + 1 < p.MB) tmem_ld16_issue(tcol + uint32_t(mb + 1) * tstep, rb);
                process(ra, mb);
                if (mb + 1 < p.MB) {
                    tmem_ld_wait16(rb);
                    if (mb + 2 < p.MB) tmem_ld16_issue(tcol + uint32_t(mb + 2) * tstep, ra);
                    process(rb, mb + 1);
                }
            }
            }
            if (want_stats) {
                const float r1 = warp_transpose_reduce16<float>(s1, lane);
                const float r2 = warp_transpose_reduce16<float>(s2, lane);
                if ((lane & 1) == 0) {
                    const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    ST *a = sAcc + (acc_row * CoutP + cobase + ch) * 2;
                    a[0] += ST(r1);
                    a[1] += ST(r2);
                }
            }
        }
        if (!waited) {  // a warp without a column group of its own still has to observe the phase
            mbar_wait<X3 ? 512 : 256>(acc_full + buf, aph);  // the next accumulator is microseconds away
            tc_fence_after();
        }
        // accumulator buffer drained: hand it back to the MMA warps
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + buf);
        if (tid == 0 && it == it_begin) CCDM_EPI_TRACE(6);
        if (tid == 0) CCDM_EPI_TL(it - it_begin, 1);
        ++n_pending;
    }
    if (tid == 0) CCDM_EPI_TRACE(7);
}

}  // namespace
}  // namespace ccdm
