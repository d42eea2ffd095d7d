// Stage B on the tensor cores: impression scoring for H <= 64 history rows (sm_100a, tcgen05 + TMEM).
//
// Same arithmetic as score.cu (see its header for the algebra and the reference lines), restructured so
// that the 400-wide gate sigmoid is no longer evaluated per (candidate, history row):
//
//   o_h(a) = v_h * (1 - (1 - a) * sigmoid(a * W_g v_h + b_g))       depends on the candidate only through
//                                                                     the scalar attention weight a = a[c][h].
//   Over the <= 42 candidates of a work unit, a[.][h] spans an interval [mid_h - w_h, mid_h + w_h].  o_h is
//   analytic in a, so it is evaluated EXACTLY at the 4 Chebyshev nodes a_j of that interval and every
//   candidate interpolates:  o_h(a[c][h]) = sum_j L_j(t) o_h(a_j),  t = (a[c][h] - mid_h) / w_h.
//   The reductions a pair needs are linear in o (dots with the candidate's 3 folded vectors, sum o) or are
//   scalar functions of a (sum o^2), so they interpolate the same way:
//       D_k[c][h] = sum_j L_j(t) * ( O[4h+j][:] . w_k[c][:] ),    O = [o_h(a_j)]  (4H x 400),  k = 1..3
//   and O . W^T  (4H x 400) x (400 x 3C) is ONE GEMM per work unit -> tcgen05.mma.
//   Interpolation error of the gate for 4 Chebyshev nodes:  <= w^4 (0.125 g^4 + 0.5 |g|^3) / 192  with
//   g = max_d |W_g v_h|  (4th derivative of (1-a) sigmoid(g a + b)); a unit where this exceeds the tolerance
//   (default 1e-6, i.e. below the 2^-22 of the ex2.approx the exact kernel uses) is appended to a
//   device-side list and re-scored by the exact kernel.  fp32 fidelity of the dots: both GEMM operands are
//   split x = hi + lo into two fp16 (11 + 11 significant bits) and D += Ahi Bhi + Ahi Blo + Alo Bhi with fp32
//   accumulation in TMEM (relative error ~2^-21 per product, the level of an fp32 FMA chain; a bf16 pair,
//   2^-17, measurably is not enough: 6e-5 on the logits).  Units holding a value beyond the fp16 range are
//   flagged for the exact kernel as well.
//
// CTA = 8 compute warps + 1 MMA-issuer warp, persistent, TWO per SM (110 KB of shared memory and 256 TMEM
// columns each) so that one CTA's latency-bound phases overlap the other's math; work units as in score.cu.
// Per unit:  phase 0 metadata / bucketize / L2 prefetch of the unit's cache rows -> phase 1 topic attention
//   a[c][h]: head logits gathered from the (candidate topic, history topic) table, softmaxes in registers ->
//   nodes -> 13 K-chunks of 32 dims: compute warps write the swizzled fp16 hi/lo operand tiles of
//   O (256 rows) and W (<=112 rows); the two 32-dim halves of the 64-dim SWIZZLE_128B tile act as a 2-stage
//   ring (mbarrier full/free) so the issuer warp runs chunk k's MMAs while chunk k+1 is produced ->
//   epilogue: TMEM -> registers, Lagrange combination over the 4 lanes of a quad, LayerNorm folding ->
//   phase 3 pooling softmax + GraphSAGE mean + lifetime weight.
#include "score_common.cuh"
#include <cuda_fp16.h>

#include "tc05.cuh"

namespace lime {
namespace {

constexpr int kD = LIME_D;
constexpr int kWarps = 8;                       // compute warps; warp 8 issues the MMAs
constexpr int kCompute = kWarps * 32;
constexpr int kThreads = kCompute + 32;
constexpr int kRows = LIME_TC_MAX_HISTORY;      // history rows per unit -> 4 * 64 = 256 operand rows
constexpr int kTile = LIME_TC_TILE_C;           // candidates per unit -> 3 * 37 = 111 <= 112 MMA columns
constexpr int kNMax = 112;
constexpr int kChunks = 13;                     // 32-wide K chunks over D = 400 (the last holds 16 dims)
constexpr int kABytes = 256 * 128;              // one 64-dim operand image of O (hi or lo)
constexpr int kBBytes = kNMax * 128;            // one 64-dim operand image of W (hi or lo)
constexpr int kTileBytes = 2 * kABytes + 2 * kBBytes;
constexpr int kTabLd = LIME_TOPIC_TAB_LD;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kHalfSafe = 32768.0f;           // operands beyond this are not split into fp16 pairs

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_AHI = 0, OFF_ALO = kABytes, OFF_BHI = 2 * kABytes, OFF_BLO = 2 * kABytes + kBBytes;
constexpr int OFF_LG = 0;                                     // alias (after the MMAs): lg / y / z [37][64]
constexpr int OFF_Y = OFF_LG + kTile * kRows * 4;
constexpr int OFF_Z = OFF_Y + kTile * kRows * 4;
constexpr int OFF_A = kTileBytes;                             // a[c][h]  [37][64]
constexpr int OFF_BIAS = OFF_A + kTile * kRows * 4;           // gate bias' [400]
constexpr int OFF_S01 = OFF_BIAS + kD * 4;                    // node sums [64][4][2]
constexpr int OFF_MID = OFF_S01 + kRows * 8 * 4;
constexpr int OFF_WINV = OFF_MID + kRows * 4;
constexpr int OFF_WHALF = OFF_WINV + kRows * 4;
constexpr int OFF_GMAX = OFF_WHALF + kRows * 4;
constexpr int OFF_CSCAL = OFF_GMAX + kRows * 4;               // [37][8]
constexpr int OFF_CW = OFF_CSCAL + kTile * 8 * 4;
constexpr int OFF_CNEWS = OFF_CW + 160;
constexpr int OFF_CTAB = OFF_CNEWS + 160;
constexpr int OFF_CP = OFF_CTAB + 160;
constexpr int OFF_CTOPIC = OFF_CP + 160;
constexpr int OFF_HNEWS = OFF_CTOPIC + 160;
constexpr int OFF_HTAB = OFF_HNEWS + kRows * 4;
constexpr int OFF_HMASK = OFF_HTAB + kRows * 4;
constexpr int OFF_HTOPIC = OFF_HMASK + kRows * 4;
constexpr int OFF_BARS = OFF_HTOPIC + kRows * 4;              // full[2] free[2] accum
constexpr int OFF_MISC = OFF_BARS + 64;                       // tmem slot, unit broadcast, flag
constexpr int kSmemBytes = OFF_MISC + 64 + 1024;
static_assert(OFF_Z + kTile * kRows * 4 <= 2 * kABytes, "epilogue alias overflows the O operand images");
static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");
static_assert(3 * kTile * 4 <= 2 * kCompute, "candidate operand: at most 2 staging tasks per compute thread");
static_assert(OFF_BARS % 8 == 0 && OFF_A % 16 == 0 && OFF_BIAS % 16 == 0 && OFF_BLO % 1024 == 0, "alignment");

// Chebyshev nodes on [-1, 1]: 4-node and 2-node sets
constexpr float kX0 = -0.92387953251128674f, kX1 = -0.38268343236508977f;
constexpr float kX2 = 0.38268343236508977f, kX3 = 0.92387953251128674f;
constexpr float kY0 = -0.70710678118654752f, kY1 = 0.70710678118654752f;

template <int NODES> __device__ __forceinline__ float node_x(int j) {
    if (NODES == 2) return j == 0 ? kY0 : kY1;
    return j == 0 ? kX0 : j == 1 ? kX1 : j == 2 ? kX2 : kX3;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 8 fp32 -> 8 fp16 hi + 8 fp16 lo (x = hi + lo to 2^-22: 11 + 11 significant bits), as two 16-byte
// chunks; returns max |x| so the caller can flag values outside the fp16 range
__device__ __forceinline__ float split8(const float (&x)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
    float mx = 0.0f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __half2 hh = __floats2half2_rn(x[2 * p], x[2 * p + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(x[2 * p] - hf.x, x[2 * p + 1] - hf.y);
        h[p] = *reinterpret_cast<const uint32_t *>(&hh);
        l[p] = *reinterpret_cast<const uint32_t *>(&ll);
        mx = fmaxf(mx, fmaxf(fabsf(x[2 * p]), fabsf(x[2 * p + 1])));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
    return mx;
}

// Operand production for one work unit: NODES rows of O per history row (fp16 hi / lo images, K chunks of
// 32 dims through the two halves of the 64-dim tile) and the folded vectors of the candidates; also the node
// sums (sum o, sum o^2).  Executed by the 256 compute threads; unit_iter (units done by this CTA) gives the
// mbarrier phases: the two half-tiles are staged 7 and 6 times per unit.
template <int NODES>
__device__ __forceinline__ void produce_operands(unsigned char *base, const LimeNewsCache &C, int H, int cnt, int tid,
                                                 const int *hnews, const int *htab, const int *cnews, const int *ctab,
                                                 const float *mid_s, const float *whalf_s, const float *bias_s,
                                                 float *s01_s, int *flag_s, uint64_t *bar_full, uint64_t *bar_free,
                                                 uint32_t unit_iter) {
    const int hr = tid >> 2, q = tid & 3;
    const bool row_ok = hr < H;
    const float mid = row_ok ? mid_s[hr] : 0.0f, wh = row_ok ? whalf_s[hr] : 0.0f;
    float aj[NODES];
#pragma unroll
    for (int j = 0; j < NODES; ++j) aj[j] = fmaf(wh, node_x<NODES>(j), mid);
    const float *hrow = C.hist_rows + (size_t)(row_ok ? hnews[hr] : 0) * LIME_HIST_LD;
    const float *trow = C.hist_tab + (size_t)(row_ok ? htab[hr] : 0) * LIME_HTAB_LD;
    float ps[2 * NODES];
#pragma unroll
    for (int i = 0; i < 2 * NODES; ++i) ps[i] = 0.0f;
    float xmax = 0.0f;
    const int btasks = 3 * cnt * 4;
    for (int kc = 0; kc < kChunks; ++kc) {
        const int s = kc & 1;
        const uint32_t u = unit_iter * (s ? 6u : 7u) + (uint32_t)(kc >> 1);   // uses of this half so far
        if (u >= 1) tc::mbar_wait(bar_free + s, (u - 1) & 1u);
        const int d0 = 32 * kc + 8 * q;
        const int lq = 4 * s + q;                  // 16-byte chunk inside the 128-byte tile row
        if (row_ok && d0 < kD) {
            float v[8], gg[8];
            {
                const float4 a0 = ldg4(hrow + LIME_HIST_VC + d0), a1 = ldg4(hrow + LIME_HIST_VC + d0 + 4);
                const float4 b0 = ldg4(trow + d0), b1 = ldg4(trow + d0 + 4);
                const float4 c0 = ldg4(hrow + LIME_HIST_GW + d0), c1 = ldg4(hrow + LIME_HIST_GW + d0 + 4);
                const float4 e0 = ldg4(trow + kD + d0), e1 = ldg4(trow + kD + d0 + 4);
                v[0] = a0.x + b0.x; v[1] = a0.y + b0.y; v[2] = a0.z + b0.z; v[3] = a0.w + b0.w;
                v[4] = a1.x + b1.x; v[5] = a1.y + b1.y; v[6] = a1.z + b1.z; v[7] = a1.w + b1.w;
                gg[0] = c0.x + e0.x; gg[1] = c0.y + e0.y; gg[2] = c0.z + e0.z; gg[3] = c0.w + e0.w;
                gg[4] = c1.x + e1.x; gg[5] = c1.y + e1.y; gg[6] = c1.z + e1.z; gg[7] = c1.w + e1.w;
            }
            const float4 bb0 = *reinterpret_cast<const float4 *>(bias_s + d0);
            const float4 bb1 = *reinterpret_cast<const float4 *>(bias_s + d0 + 4);
            const float bb[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) xmax = fmaxf(xmax, fabsf(v[e]));   // |o| <= |v|: one range check per element
#pragma unroll
            for (int j = 0; j < NODES; ++j) {
                // o = v (1 - (1 - a_j) sigmoid(a_j W_g v + b_g)),  sigmoid(z) = 1 / (1 + 2^z'),  z' = -log2(e) z
                const float a = aj[j], oma = 1.0f - a;
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float den = ex2_approx(fmaf(a, gg[e], bb[e])) + 1.0f;
                    o[e] = fmaf(-(v[e] * oma), rcp_approx(den), v[e]);
                    ps[2 * j] += o[e];
                    ps[2 * j + 1] = fmaf(o[e], o[e], ps[2 * j + 1]);
                }
                uint4 hi, lo;
                split8(o, hi, lo);
                const uint32_t off = tc::sw128_offset(NODES * hr + j, lq);
                *reinterpret_cast<uint4 *>(base + OFF_AHI + off) = hi;
                *reinterpret_cast<uint4 *>(base + OFF_ALO + off) = lo;
            }
        }
        for (int task = tid; task < btasks; task += kCompute) {
            const int n = task >> 2, qq = task & 3;
            const int d1 = 32 * kc + 8 * qq;
            if (d1 >= kD) continue;
            const int c = n / 3, k = n - 3 * c;
            const float *cr = C.cand_rows + (size_t)cnews[c] * LIME_CAND_LD + k * kD + d1;
            const float *ct = C.cand_tab + (size_t)ctab[c] * LIME_CTAB_LD + k * kD + d1;
            const float4 a0 = ldg4(cr), a1 = ldg4(cr + 4), b0 = ldg4(ct), b1 = ldg4(ct + 4);
            const float x[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w,
                                a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
            uint4 hi, lo;
            xmax = fmaxf(xmax, split8(x, hi, lo));
            const uint32_t off = tc::sw128_offset(n, 4 * s + qq);
            *reinterpret_cast<uint4 *>(base + OFF_BHI + off) = hi;
            *reinterpret_cast<uint4 *>(base + OFF_BLO + off) = lo;
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(bar_full + s);
    }
    // node sums of the row: sum o, sum o^2 per node, over the 4 lanes that share the row
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
#pragma unroll
        for (int i = 0; i < 2 * NODES; ++i) ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
    }
    if (!(xmax <= kHalfSafe)) atomicOr(flag_s, 4);   // outside the fp16 operand range (or NaN): exact kernel
    if (row_ok && q == 0) {
#pragma unroll
        for (int i = 0; i < 2 * NODES; ++i) s01_s[hr * 8 + i] = ps[i];
    }
}

// Epilogue of one work unit: every compute thread owns one accumulator row (TMEM lane) = one (history row,
// node); the NODES lanes of a row combine their dots with the Lagrange weights of each candidate, lane j = 0
// folds the LayerNorm and writes lg / y / z.
template <int NODES>
__device__ __forceinline__ void epilogue(uint32_t tmem, int warp, int lane, int H, int cnt, int mtiles, float ln_eps,
                                         const float *a_s, const float *mid_s, const float *winv_s, const float *s01_s,
                                         const float *cscal, float *lg_s, float *y_s, float *z_s) {
    const int qd = warp & 3;
    const int mt = NODES == 2 ? 0 : warp >> 2;               // 2 nodes: one 128-row tile, warps 4-7 take the odd
    const int cg0 = NODES == 2 ? warp >> 2 : 0, cgs = NODES == 2 ? 2 : 1;   // groups of 16 candidates
    if (mt < mtiles) {
        const int r = 128 * mt + 32 * qd + lane;
        const int hr = r / NODES, j = r % NODES;
        const bool row_ok = hr < H;
        const int hc = row_ok ? hr : 0;
        const float mid = mid_s[hc], winv = winv_s[hc];
        const float s0n = row_ok ? s01_s[hc * 8 + 2 * j] : 0.0f, s1n = row_ok ? s01_s[hc * 8 + 2 * j + 1] : 0.0f;
        // L_j(t) = prod_{m != j} (t - x_m) / (x_j - x_m); with 2 nodes there is a single factor
        float xa, xb = 0.0f, xc = 0.0f, invden;
        if (NODES == 2) {
            xa = j == 0 ? kY1 : kY0;
            invden = 1.0f / ((j == 0 ? kY0 : kY1) - xa);
        } else {
            const float xj = j == 0 ? kX0 : j == 1 ? kX1 : j == 2 ? kX2 : kX3;
            xa = j == 0 ? kX1 : kX0, xb = j <= 1 ? kX2 : kX1, xc = j == 3 ? kX2 : kX3;
            invden = 1.0f / ((xj - xa) * (xj - xb) * (xj - xc));
        }
        const uint32_t taddr = tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(128 * mt);
        for (int cg = cg0; 16 * cg < cnt; cg += cgs) {
            float v[48];
            tc::tmem_ld16(taddr + 48 * cg, *reinterpret_cast<float(*)[16]>(&v[0]));
            if (cg < 2) {
                tc::tmem_ld16(taddr + 48 * cg + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
                tc::tmem_ld16(taddr + 48 * cg + 32, *reinterpret_cast<float(*)[16]>(&v[32]));
            } else {        // candidates 32..36 live in columns 96..110 of the 112-column tile
#pragma unroll
                for (int i = 16; i < 48; ++i) v[i] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = 16 * cg + i;
                if (c < cnt) {
                    float t = (a_s[c * kRows + hc] - mid) * winv;
                    t = fminf(fmaxf(t, -1.0f), 1.0f);
                    const float L = NODES == 2 ? (t - xa) * invden : (t - xa) * (t - xb) * (t - xc) * invden;
                    float p0 = L * v[3 * i], p1 = L * v[3 * i + 1], p2 = L * v[3 * i + 2];
                    float p3 = L * s0n, p4 = L * s1n;
#pragma unroll
                    for (int o = 1; o < NODES; o <<= 1) {
                        p0 += __shfl_xor_sync(0xffffffffu, p0, o);
                        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                        p2 += __shfl_xor_sync(0xffffffffu, p2, o);
                        p3 += __shfl_xor_sync(0xffffffffu, p3, o);
                        p4 += __shfl_xor_sync(0xffffffffu, p4, o);
                    }
                    if (j == 0 && row_ok) {
                        const float *cs = cscal + c * 8;
                        const float mu = p3 * (1.0f / kD);
                        const float var = fmaxf(fmaf(-mu, mu, p4 * (1.0f / kD)), 0.0f);
                        const float rstd = rsqrtf(var + ln_eps);
                        lg_s[c * kRows + hr] = fmaf(rstd, fmaf(-mu, cs[0], p0), cs[3]);
                        y_s[c * kRows + hr] = fmaf(rstd, fmaf(-mu, cs[1], p1), cs[4]);
                        z_s[c * kRows + hr] = fmaf(rstd, fmaf(-mu, cs[2], p2), cs[5]);
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads, 2) score_tc_kernel(const ScoreArgs args) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float *lg_s = reinterpret_cast<float *>(base + OFF_LG);
    float *y_s = reinterpret_cast<float *>(base + OFF_Y);
    float *z_s = reinterpret_cast<float *>(base + OFF_Z);
    float *a_s = reinterpret_cast<float *>(base + OFF_A);
    float *bias_s = reinterpret_cast<float *>(base + OFF_BIAS);
    float *s01_s = reinterpret_cast<float *>(base + OFF_S01);
    float *mid_s = reinterpret_cast<float *>(base + OFF_MID);
    float *winv_s = reinterpret_cast<float *>(base + OFF_WINV);
    float *whalf_s = reinterpret_cast<float *>(base + OFF_WHALF);
    float *cscal = reinterpret_cast<float *>(base + OFF_CSCAL);
    float *cw = reinterpret_cast<float *>(base + OFF_CW);
    int *cnews = reinterpret_cast<int *>(base + OFF_CNEWS);
    int *ctab = reinterpret_cast<int *>(base + OFF_CTAB);
    int *cP = reinterpret_cast<int *>(base + OFF_CP);
    int *ctopic = reinterpret_cast<int *>(base + OFF_CTOPIC);
    int *hnews = reinterpret_cast<int *>(base + OFF_HNEWS);
    int *htab = reinterpret_cast<int *>(base + OFF_HTAB);
    int *hmask = reinterpret_cast<int *>(base + OFF_HMASK);
    int *htopic = reinterpret_cast<int *>(base + OFF_HTOPIC);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(base + OFF_BARS);
    uint64_t *bar_free = bar_full + 2;
    uint64_t *bar_accum = bar_full + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(base + OFF_MISC);
    int *unit_bcast = reinterpret_cast<int *>(base + OFF_MISC + 8);
    int *flag_s = reinterpret_cast<int *>(base + OFF_MISC + 16);

    const LimeNewsCache &C = args.cache;
    const LimeImpressions &I = args.imp;
    const int H = I.max_history;
    const int nb = C.num_buckets;
    const int T = C.num_topics;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int d = tid; d < kD; d += kThreads) bias_s[d] = C.gate_bias[d];
    if (tid == 0) {
        tc::mbar_init(bar_full + 0, kCompute);
        tc::mbar_init(bar_full + 1, kCompute);
        tc::mbar_init(bar_free + 0, 1);
        tc::mbar_init(bar_free + 1, 1);
        tc::mbar_init(bar_accum, 1);
        tc::mbar_fence_init();
    }
    if (warp == kWarps) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    uint32_t unit_iter = 0;       // units processed so far (phase of bar_accum)

    for (;;) {
        __syncthreads();
        if (tid == 0) unit_bcast[0] = atomicAdd(args.work_counter, 1);
        __syncthreads();
        const int unit = unit_bcast[0];
        if (unit >= I.num_units) break;
        const int imp = I.unit_imp[unit];
        const int pair0 = I.unit_pair0[unit];
        const int cnt = I.unit_count[unit];

        // ---------------- phase 0: unit metadata, L2 prefetch of the rows the unit will read ------
        for (int h = tid; h < H; h += kThreads) {
            const long long o = (long long)imp * H + h;
            int n = I.hist_news[o];
            n = (n < 0 || n >= C.news_num) ? 0 : n;
            hnews[h] = n;
            hmask[h] = I.hist_mask[o];
            const int bf = bucketize_seconds(I.hist_fresh[o], args.bucket_scale, nb);
            const int bl = bucketize_seconds(I.hist_life[o], args.bucket_scale, nb);
            htab[h] = bf * nb + bl;
            const float *hrow = C.hist_rows + (size_t)n * LIME_HIST_LD;
            int tp = __float_as_int(__ldg(hrow + LIME_HIST_TOPIC_ID));
            htopic[h] = (tp < 0 || tp >= T) ? 0 : tp;
        }
        for (int c = tid; c < cnt; c += kThreads) {
            const long long p = (long long)pair0 + c;
            int n = I.cand_news[p];
            n = (n < 0 || n >= C.news_num) ? 0 : n;
            cnews[c] = n;
            const float fr = I.cand_fresh[p], lf = I.cand_life[p];
            ctab[c] = bucketize_seconds(fr, args.bucket_scale, nb) * nb + bucketize_seconds(lf, args.bucket_scale, nb);
            cw[c] = lifetime_weight(I.cand_remaining ? I.cand_remaining[p] : __fsub_rn(lf, fr), C);
            cP[c] = (args.pair_index_base + p >= args.tail_start) ? args.prefix_tail : args.prefix_main;
            const float *crow = C.cand_rows + (size_t)n * LIME_CAND_LD;
            int tp = __float_as_int(__ldg(crow + LIME_CAND_TOPIC_ID));
            ctopic[c] = (tp < 0 || tp >= T) ? 0 : tp;
        }
        if (tid == 0) flag_s[0] = 0;
        __syncthreads();
        // history rows: vc | gw = 3200 B = 25 lines; candidate rows: w1 w2 w3 = 4800 B = 38 lines (+1 unaligned)
        for (int idx = tid; idx < H * 26; idx += kThreads) {
            const int h = idx / 26, l = idx - h * 26;
            prefetch_l2(reinterpret_cast<const char *>(C.hist_rows + (size_t)hnews[h] * LIME_HIST_LD) + l * 128);
        }
        for (int idx = tid; idx < cnt * 39; idx += kThreads) {
            const int c = idx / 39, l = idx - c * 39;
            prefetch_l2(reinterpret_cast<const char *>(C.cand_rows + (size_t)cnews[c] * LIME_CAND_LD) + l * 128);
        }

        // ---------------- phase 1: candidate-aware attention weights a[c][h] (layers.py:66-81) ----
        if (warp < kWarps) {
            const bool v0 = lane < H, v1 = lane + 32 < H;
            const int t0 = htopic[v0 ? lane : 0], t1 = htopic[v1 ? lane + 32 : 0];
            const bool k0 = v0 && hmask[v0 ? lane : 0] != 0, k1 = v1 && hmask[v1 ? lane + 32 : 0] != 0;
            for (int c = warp; c < cnt; c += kWarps) {
                if (lane < 8) {
                    const float *crow = C.cand_rows + (size_t)cnews[c] * LIME_CAND_LD;
                    const float tabv = C.cand_tab[(size_t)ctab[c] * LIME_CTAB_LD + LIME_CAND_SCAL + lane];
                    cscal[c * 8 + lane] = crow[LIME_CAND_SCAL + lane] + tabv;
                }
                const float *trow = C.topic_table + (size_t)ctopic[c] * T * kTabLd;
                float sc[2][LIME_CA_HEADS];
                {
                    const float *r0 = trow + (size_t)t0 * kTabLd, *r1 = trow + (size_t)t1 * kTabLd;
                    const float4 a0 = ldg4(r0), a1 = ldg4(r0 + 4), a2 = ldg4(r0 + 8);
                    const float4 b0 = ldg4(r1), b1 = ldg4(r1 + 4), b2 = ldg4(r1 + 8);
                    const float x0[LIME_CA_HEADS] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y};
                    const float x1[LIME_CA_HEADS] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y};
#pragma unroll
                    for (int hd = 0; hd < LIME_CA_HEADS; ++hd) {   // masked_fill(mask == 0, -1e9), layers.py:72
                        sc[0][hd] = v0 ? (k0 ? x0[hd] : -1e9f) : -INFINITY;
                        sc[1][hd] = v1 ? (k1 ? x1[hd] : -1e9f) : -INFINITY;
                    }
                }
                float agg[2] = {0.0f, 0.0f};
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) {
                    const float m = warp_max(fmaxf(sc[0][hd], sc[1][hd]));
                    const float e0 = __expf(sc[0][hd] - m), e1 = __expf(sc[1][hd] - m);
                    const float inv = __fdividef(1.0f, warp_sum(e0 + e1));
                    agg[0] = fmaf(e0, inv, agg[0]);
                    agg[1] = fmaf(e1, inv, agg[1]);
                }
                // second, unmasked softmax over the history (layers.py:81)
                const float m2 = warp_max(fmaxf(v0 ? agg[0] : -INFINITY, v1 ? agg[1] : -INFINITY));
                agg[0] = v0 ? __expf(agg[0] - m2) : 0.0f;
                agg[1] = v1 ? __expf(agg[1] - m2) : 0.0f;
                const float inv2 = __fdividef(1.0f, warp_sum(agg[0] + agg[1]));
                if (v0) a_s[c * kRows + lane] = agg[0] * inv2;
                if (v1) a_s[c * kRows + lane + 32] = agg[1] * inv2;
            }
        }
        __syncthreads();

        // ---------------- interpolation nodes per history row ------------------------------------
        if (tid < H) {
            float lo = a_s[tid], hi = lo;
            for (int c = 1; c < cnt; ++c) {
                const float a = a_s[c * kRows + tid];
                lo = fminf(lo, a);
                hi = fmaxf(hi, a);
            }
            const float wh = fmaxf(0.5f * (hi - lo), 1e-7f);
            mid_s[tid] = 0.5f * (hi + lo);
            whalf_s[tid] = wh;
            winv_s[tid] = 1.0f / wh;
            // Interpolation error of f(a) = (1 - a) sigmoid(g a + b) on [mid - w, mid + w] with n Chebyshev nodes:
            // max|d^n f| w^n / (n! 2^(n-1));  |d^2 f| <= 0.0962 g^2 + 0.5 |g|,  |d^4 f| <= 0.125 g^4 + 0.5 |g|^3.
            // g is bounded by the cached max |W_g vc| of the news plus the max over the bucket-pair table.
            const float gabs = (__ldg(C.hist_rows + (size_t)hnews[tid] * LIME_HIST_LD + LIME_HIST_GW_ABSMAX) + C.tab_gw_absmax) *
                               (1.0f / kLog2e);
            const float w2 = wh * wh, g2 = gabs * gabs;
            const float err2 = w2 * (0.0962f * g2 + 0.5f * gabs) * 0.25f;
            const float err4 = w2 * w2 * (0.125f * g2 * g2 + 0.5f * g2 * gabs) * (1.0f / 192.0f);
            if (!(err2 <= args.interp_tol)) atomicOr(flag_s, 1);          // 2 nodes are not enough
            if (!(err4 <= args.interp_tol)) atomicOr(flag_s, 2);          // 4 nodes are not enough: exact kernel
        }
        __syncthreads();
        const int flags0 = flag_s[0];
        const int nodes = (flags0 & 1) ? 4 : 2;

        const int n_cols = (3 * cnt + 15) & ~15;
        const int mtiles = (nodes * H + 127) >> 7;

        if (warp < kWarps) {
            if (nodes == 2) produce_operands<2>(base, C, H, cnt, tid, hnews, htab, cnews, ctab, mid_s, whalf_s, bias_s, s01_s,
                                                flag_s, bar_full, bar_free, unit_iter);
            else            produce_operands<4>(base, C, H, cnt, tid, hnews, htab, cnews, ctab, mid_s, whalf_s, bias_s, s01_s,
                                                flag_s, bar_full, bar_free, unit_iter);
        } else {
            // ---------------- MMA issuer ---------------------------------------------------------
            const uint32_t idesc = tc::idesc_f16_f32(128, n_cols);
            const uint32_t sb = tc::smem_u32(base);
            for (int kc = 0; kc < kChunks; ++kc) {
                const int s = kc & 1;
                const uint32_t u = unit_iter * (s ? 6u : 7u) + (uint32_t)(kc >> 1);
                tc::mbar_wait(bar_full + s, u & 1u);
                tc::fence_after_sync();
                if (lane == 0) {
                    const int ksteps = kc < kChunks - 1 ? 2 : (kD - 32 * (kChunks - 1)) / 16;
                    const uint64_t bhi = tc::smem_desc_sw128(sb + OFF_BHI), blo = tc::smem_desc_sw128(sb + OFF_BLO);
                    for (int mt = 0; mt < mtiles; ++mt) {
                        const uint64_t ahi = tc::smem_desc_sw128(sb + OFF_AHI + mt * 16384);
                        const uint64_t alo = tc::smem_desc_sw128(sb + OFF_ALO + mt * 16384);
                        const uint32_t td = tmem + (uint32_t)(mt * 128);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t k2 = (uint64_t)(2 * (2 * s + ks));   // 32 bytes per K step of 16
                            tc::mma_f16(td, ahi + k2, bhi + k2, idesc, (kc | ks) != 0);
                            tc::mma_f16(td, ahi + k2, blo + k2, idesc, true);
                            tc::mma_f16(td, alo + k2, bhi + k2, idesc, true);
                        }
                    }
                    tc::mma_commit(bar_free + s);
                    if (kc == kChunks - 1) tc::mma_commit(bar_accum);
                }
                __syncwarp();
            }
        }
        __syncthreads();   // node sums visible

        if (warp < kWarps) {
            // ---------------- epilogue: TMEM -> Lagrange combination -> LayerNorm folding ---------
            tc::mbar_wait(bar_accum, unit_iter & 1u);
            tc::fence_after_sync();
            if (nodes == 2) epilogue<2>(tmem, warp, lane, H, cnt, mtiles, args.ln_eps, a_s, mid_s, winv_s, s01_s, cscal, lg_s, y_s, z_s);
            else            epilogue<4>(tmem, warp, lane, H, cnt, mtiles, args.ln_eps, a_s, mid_s, winv_s, s01_s, cscal, lg_s, y_s, z_s);
        }
        tc::fence_before_sync();
        __syncthreads();
        ++unit_iter;

        if (tid == 0 && (flag_s[0] & 6) != 0) args.fallback_list[atomicAdd(args.fallback_count, 1)] = unit;

        // ---------------- phase 3: candidate-query pooling + lifetime-weighted dot ---------------
        if (warp < kWarps) {
            for (int c = warp; c < cnt; c += kWarps) {
                const int P = cP[c];
                const int pz = P < H ? P : H;
                float m = -INFINITY;
                for (int hh = lane; hh < H; hh += 32) m = fmaxf(m, lg_s[c * kRows + hh]);
                m = warp_max(m);
                float l = 0.f, acc = 0.f, ms = 0.f;
                for (int hh = lane; hh < H; hh += 32) {
                    const float e = __expf(lg_s[c * kRows + hh] - m);
                    l += e;
                    acc = fmaf(e, y_s[c * kRows + hh], acc);
                    if (hh < pz) ms += z_s[c * kRows + hh];
                }
                l = warp_sum(l);
                acc = warp_sum(acc);
                ms = warp_sum(ms);
                float un = 0.f;
                if (P > H) {   // user-node rows take part in the GraphSAGE mean (userEncoders.py:121,153)
                    int jn = P - H - 1;
                    jn = jn < C.user_nodes ? jn : C.user_nodes - 1;
                    const float *uu = C.un_prefix + (size_t)jn * kD;
                    const float *hr2 = C.hist_rows + (size_t)cnews[c] * LIME_HIST_LD + LIME_HIST_VC;
                    const float *tr2 = C.hist_tab + (size_t)ctab[c] * LIME_HTAB_LD;
                    for (int d = lane; d < kD; d += 32) un = fmaf(hr2[d] + tr2[d], uu[d], un);
                    un = warp_sum(un);
                }
                if (lane == 0) {
                    const float bs = (ms + un) / (float)P + cscal[c * 8 + 6] + acc / l;
                    args.scores[(long long)pair0 + c] = bs * cw[c];
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kWarps) tc::tmem_dealloc(tmem, 256);
}

// out[(tc * T + th) * 12 + head] = sum_k tq[tc][k * 10 + head] * topics[th][k] + tq[tc][500 + head]
// (same accumulation order as phase 1 of the exact kernel, so both kernels see identical logits)
__global__ void topic_pair_table_kernel(const float *__restrict__ topics, int64_t ldt, const float *__restrict__ tq,
                                        int64_t ldq, int T, float *__restrict__ out) {
    __shared__ float q_s[LIME_TOPIC * LIME_CA_HEADS + LIME_CA_HEADS];
    const int tcand = blockIdx.x;
    for (int i = threadIdx.x; i < LIME_TOPIC * LIME_CA_HEADS + LIME_CA_HEADS; i += blockDim.x) q_s[i] = tq[(size_t)tcand * ldq + i];
    __syncthreads();
    for (int th = threadIdx.x; th < T; th += blockDim.x) {
        float acc[LIME_CA_HEADS];
#pragma unroll
        for (int hd = 0; hd < LIME_CA_HEADS; ++hd) acc[hd] = q_s[LIME_TOPIC * LIME_CA_HEADS + hd];
        for (int k = 0; k < LIME_TOPIC; ++k) {
            const float tv = topics[(size_t)th * ldt + k];
#pragma unroll
            for (int hd = 0; hd < LIME_CA_HEADS; ++hd) acc[hd] = fmaf(q_s[k * LIME_CA_HEADS + hd], tv, acc[hd]);
        }
        float *o = out + ((size_t)tcand * T + th) * kTabLd;
#pragma unroll
        for (int hd = 0; hd < LIME_CA_HEADS; ++hd) o[hd] = acc[hd];
        o[10] = 0.0f;
        o[11] = 0.0f;
    }
}

}  // namespace

int launch_score_tc(const ScoreArgs &a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        LIME_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    LIME_CUDA(cudaMemsetAsync(a.work_counter, 0, 2 * sizeof(int32_t), st));   // work counter + fallback count
    int grid = 2 * num_sms();
    if (grid > a.imp.num_units) grid = a.imp.num_units;
    score_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(a);
    LIME_LAUNCH_CHECK("score_tc_kernel");
    return 0;
}

}  // namespace lime

extern "C" int lime_topic_pair_table(const float *topics, int64_t ldt, const float *tq, int64_t ldq, int32_t T,
                                     float *out, void *stream) {
    LIME_CHECK_ARG(topics && tq && out, "lime_topic_pair_table: null argument");
    LIME_CHECK_ARG(T >= 1 && T <= LIME_TC_MAX_TOPICS, "lime_topic_pair_table: T=%d not in [1, %d]", T, LIME_TC_MAX_TOPICS);
    lime::topic_pair_table_kernel<<<T, 128, 0, lime::as_stream(stream)>>>(topics, ldt, tq, ldq, T, out);
    LIME_LAUNCH_CHECK("topic_pair_table_kernel");
    return 0;
}
