// Stage B on the tensor cores: impression scoring for H <= 56 history rows (sm_100a, tcgen05 + TMEM).
//
// Same arithmetic as score.cu (see its header for the algebra and the reference lines), restructured so
// that the 400-wide gate sigmoid is no longer evaluated per (candidate, history row):
//
//   o_h(a) = v_h * (1 - (1 - a) * sigmoid(a * W_g v_h + b_g))       depends on the candidate only through
//                                                                     the scalar attention weight a = a[c][h].
//   Over the <= 42 candidates of a work unit, a[.][h] spans an interval [mid_h - w_h, mid_h + w_h].  o_h is
//   analytic in a, so it is evaluated EXACTLY at 2 (or 4) Chebyshev nodes a_j of that interval and every
//   candidate interpolates:  o_h(a[c][h]) = sum_j L_j(t) o_h(a_j),  t = (a[c][h] - mid_h) / w_h.
//   The reductions a pair needs are linear in o (dots with the candidate's 3 folded vectors, sum o) or are
//   scalar functions of a (sum o^2), so they interpolate the same way:
//       D_k[c][h] = sum_j L_j(t) * ( w_k[c][:] . O[nodes*h + j][:] ),   O = [o_h(a_j)],  k = 1..3
//   and  W . O^T  (3C x 400) x (400 x nodes*H)  is ONE GEMM per work unit -> tcgen05.mma.
//   Interpolation error of the gate:  2 nodes  w^2 (0.0962 g^2 + 0.5 |g|) / 4,   4 nodes
//   w^4 (0.125 g^4 + 0.5 |g|^3) / 192,  g = max_d |W_g v_h|; a unit where even 4 nodes exceed the tolerance
//   (default 1e-6, below the 2^-22 of the ex2.approx the exact kernel uses) is appended to a device-side
//   list and re-scored by the exact kernel.  fp32 fidelity of the dots: both GEMM operands are fp16 hi + lo
//   pairs (11 + 11 significant bits), D += Whi Ohi + Whi Olo + Wlo Ohi with fp32 accumulation in TMEM.
//
// What changed against the first tensor-core version (round 1c, 19.2 ms per bench step):
//   * history rows are DEDUPLICATED per unit: slots with the same (news, bucket pair, mask) -- in practice
//     the padding of a short history, dataset.py:123-128 -- are one operand row with a multiplicity that
//     enters the two softmaxes, the pooling and the GraphSAGE mean.  H = 50 slots -> ~26 rows on average.
//   * the candidate side is the M operand (TMEM lanes) and is NOT produced by compute threads any more: the
//     cache holds every folded candidate vector already split into fp16 hi / lo (cand16, ctab16), a loader
//     warp streams the rows with cp.async straight into the swizzled operand tile; the bucket-pair part
//     is a second K range (K = 400 news + 400 table) instead of a register add + split.
//   * the history side is the N operand: N = nodes * rows <= 112, so the MMA work follows the deduplicated
//     row count; the epilogue thread owns a (candidate, k) lane and walks the columns -- the Lagrange
//     combination is in-thread (no shuffles), lg / y / z lanes are warp-uniform.
//   * candidate-aware attention: 8 lanes per candidate, head logits pre-scaled by log2(e), no max pass
//     (the table's |logit| bound is checked by the host), multiplicity-weighted sums.
//
// CTA = 8 compute warps + 1 MMA-issuer warp + 1 loader warp, persistent, TWO per SM (113 KB of shared
// memory and 256 TMEM columns each).  Per unit: metadata + dedup -> attention a[c][u] -> nodes ->
// 13 K-stages of 32 dims (the two 32-dim halves of the 64-dim SWIZZLE_128B tile are a 2-stage ring with
// full/free mbarriers) -> epilogue TMEM -> lg / y / z -> pooling softmax + GraphSAGE mean + lifetime weight.
#include "score_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc05.cuh"

namespace lime {
namespace {

constexpr int kD = LIME_D;
constexpr int kCWarps = 7;                      // compute warps (8 warps per CTA -> 4 per SM sub-partition at two CTAs per SM: 128 registers, no spills)
constexpr int kCompute = kCWarps * 32;
constexpr int kMmaWarp = kCWarps;               // warp 7 issues the MMAs
constexpr int kRPT = kCompute / 8;              // operand rows produced per pass by one task slot (8 threads per row): 28
constexpr int kThreads = kCompute + 32;
constexpr int kH = LIME_TC_MAX_HISTORY;         // 56 history slots at most
constexpr int kTile = LIME_TC_TILE_C;           // 42 candidates per unit -> 126 of the 128 M rows
constexpr int kQ3 = kTile - 32;                 // candidates living in TMEM quadrant 3
constexpr int kBRows = 112;                     // N rows of the O operand: nodes * unique rows of a pass
constexpr int kStages = 13;                     // 32-wide K stages over D = 400 (the last holds 16 dims)
constexpr int kAImg = 128 * 128;                // one 64-dim image of the candidate operand (hi or lo)
constexpr int kBImg = kBRows * 128;
constexpr int kTabLd = LIME_TOPIC_TAB_LD;
constexpr int kAS = kTile;                      // row stride of a_s / lg / y / z  ([u][c])
constexpr float kLog2e = 1.4426950408889634f;
// Both operands are scaled by a power of two before the fp16 hi / lo split, so that the lo halves of typical values
// (|w| ~ 1e-2, |o| ~ 1e-1) stay in the normal fp16 range (>= 6.1e-5) instead of losing bits as subnormals; the
// accumulators are un-scaled in the epilogue (exact: powers of two).
constexpr float kWScale = LIME_CAND16_SCALE;    // cand16 / ctab16 hold w * 1024
constexpr float kOScale = 256.0f;               // O operand holds o * 256
constexpr float kUnscale = 1.0f / (kWScale * kOScale);
constexpr float kS1Max = 1073741824.0f;         // sum (256 o)^2 <= 2^30  =>  every |256 o| <= 32768 (fp16 operand range)
constexpr float kWAbsMax = 32768.0f / kWScale;  // |w| beyond this leaves the fp16 operand range -> exact kernel
constexpr int kC16 = LIME_CAND16_LD;            // fp16 elements per cand16 / ctab16 row: [k][hi | lo][400]

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_A = 0;                                      // 4 images: news hi, news lo, table hi, table lo
constexpr int OFF_BHI = 4 * kAImg, OFF_BLO = OFF_BHI + kBImg;
constexpr int OFF_OUT = OFF_BHI;                              // alias (after the MMAs): lg / y / z [52][42]
constexpr int OFF_AS = OFF_BLO + kBImg;                       // a[u][c]
constexpr int OFF_S01 = OFF_AS + kH * kAS * 4;                // node sums [52][8]
constexpr int OFF_MID = OFF_S01 + kH * 8 * 4;
constexpr int OFF_WINV = OFF_MID + kH * 4;
constexpr int OFF_WHALF = OFF_WINV + kH * 4;
constexpr int OFF_POOL = OFF_WHALF + kH * 4;                  // [42][4]  running m, l, acc, ms
// Per-unit arrays written by the front end (unit metadata + dedup): DOUBLE BUFFERED, the MMA-issuer warp prepares unit
// i + 1 while the compute warps are in the epilogue / pooling of unit i
constexpr int kCArr = 48 * 4, kUArr = kH * 4;
constexpr int UB_CSCAL = 0;                                   // [42][4]  B1 B2 B3 cb
constexpr int UB_CW = UB_CSCAL + kTile * 16;
constexpr int UB_CNEWS = UB_CW + kCArr;
constexpr int UB_CTAB = UB_CNEWS + kCArr;
constexpr int UB_CP = UB_CTAB + kCArr;
constexpr int UB_CTOPIC = UB_CP + kCArr;
constexpr int UB_UNEWS = UB_CTOPIC + kCArr;
constexpr int UB_UTAB = UB_UNEWS + kUArr;
constexpr int UB_UMASK = UB_UTAB + kUArr;
constexpr int UB_UTOPIC = UB_UMASK + kUArr;
constexpr int UB_UMULT = UB_UTOPIC + kUArr;                   // float multiplicity
constexpr int UB_UMP0 = UB_UMULT + kUArr;                     // float multiplicity inside the GraphSAGE prefix (main)
constexpr int UB_UMP1 = UB_UMP0 + kUArr;                      // ... (tail batch)
constexpr int UB_UGABS = UB_UMP1 + kUArr;
constexpr int UB_INFO = UB_UGABS + kUArr;                     // ints: unit, impression, first pair, count, U, unmasked slots, flags
constexpr int kUnitBuf = UB_INFO + 32;
constexpr int OFF_UB = OFF_POOL + kTile * 16;
// front-end scratch (one warp): keys, topic ids, gate bounds of the H history slots
constexpr int OFF_HKN = OFF_UB + 2 * kUnitBuf, OFF_HKT = OFF_HKN + kUArr, OFF_HTP = OFF_HKT + kUArr, OFF_HGA = OFF_HTP + kUArr;
constexpr int OFF_BARS = OFF_HGA + kUArr;                     // full[2] free[2] accum
constexpr int OFF_MISC = OFF_BARS + 64;                       // tmem slot, nodes, passes, N
constexpr int OFF_PROF = OFF_MISC + 64;                       // phase clocks of thread 0 (diagnostic)
constexpr int kSmemBytes = OFF_PROF + 128 + 1024;
static_assert(3 * kH * kAS * 4 <= 2 * kBImg, "epilogue alias overflows the O operand images");
static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");
static_assert(3 * kTile <= 128 && kTile <= 48 && kTile >= 32, "candidate rows: 32 per k in TMEM quadrants 0-2, the rest in quadrant 3");
static_assert(3 * kQ3 <= 32, "quadrant 3 holds (tile - 32) candidates x 3");
static_assert(OFF_BARS % 8 == 0 && OFF_AS % 16 == 0 && OFF_S01 % 16 == 0 && OFF_UB % 16 == 0 && kUnitBuf % 16 == 0 &&
              OFF_BHI % 1024 == 0 && OFF_BLO % 1024 == 0, "alignment");
static_assert(2 * kH <= kBRows && kBRows % 16 == 0 && kBRows <= 128 && kH % 4 == 0 && kH <= 64 && 4 * kH <= 2 * kBRows, "N operand");
static_assert(kH <= 2 * kRPT && kH <= kCompute / 4 && kBRows / 4 <= kRPT, "row mappings of the compute threads");

// misc ints
enum { M_TMEM = 0, M_NODES = 2, M_NPASS = 3, M_NPAD = 4 };
// unit info ints
enum { UI_UNIT = 0, UI_IMP = 1, UI_PAIR0 = 2, UI_CNT = 3, UI_U = 4, UI_NUN = 5, UI_FLAGS = 6 };

// Chebyshev nodes on [-1, 1]: 4-node and 2-node sets
constexpr float kX0 = -0.92387953251128674f, kX1 = -0.38268343236508977f;
constexpr float kX2 = 0.38268343236508977f, kX3 = 0.92387953251128674f;
constexpr float kY0 = -0.70710678118654752f, kY1 = 0.70710678118654752f;

// Phase timing (diagnostic): thread 0 of every CTA accumulates clock64() deltas per phase; lime_score_phase_clocks reads
// and clears the totals.  Slots: 0 metadata, 1 dedup, 2 attention, 3 nodes, 4 operand production (incl. ring waits),
// 5 wait for the last MMA, 6 epilogue, 7 pooling, 8 final score, 9 units, 10 barrier at the end of production.
// Compiled in only with -DLIME_TC_PHASE_CLOCKS (make PHASE_CLOCKS=1): the extra live registers cost spills.
__device__ unsigned long long g_phase_clocks[16];
#ifdef LIME_TC_PHASE_CLOCKS
#define LIME_TICK(slot)                                                  \
    do {                                                                 \
        if (tid == 0) {                                                  \
            const long long now__ = clock64();                           \
            prof[slot] += (unsigned long long)(now__ - t_last);          \
            t_last = now__;                                              \
        }                                                                \
    } while (0)
#else
#define LIME_TICK(slot) do { } while (0)
#endif

template <int NODES> __device__ __forceinline__ float node_x(int j) {
    if (NODES == 2) return j == 0 ? kY0 : kY1;
    return j == 0 ? kX0 : j == 1 ? kX1 : j == 2 ? kX2 : kX3;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
#if defined(LIME_TC_NO_ROW_PREFETCH)
    (void)p; (void)bytes;
#elif defined(LIME_TC_LINE_PREFETCH)
    for (uint32_t o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p) + o));
#else
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on the mbarrier once every cp.async issued so far by this thread has completed (counts as one arrival)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory"); }

// (candidate, folded vector k) of M row 32 * quadrant + lane (= TMEM lane): k = 0..2 fill quadrants 0..2 for
// c < 32, the remaining candidates share quadrant 3; c = kTile marks an unused row
__device__ __forceinline__ void m_row_owner(int quadrant, int lane, int &c, int &k) {
    if (quadrant < 3) {
        c = lane;
        k = quadrant;
    } else {
        k = lane / kQ3;
        c = 32 + lane - k * kQ3;
        if (k > 2) {
            k = 2;
            c = kTile;
        }
    }
}

// 4 fp32 -> 4 fp16 hi + 4 fp16 lo (x = hi + lo to 2^-22: 11 + 11 significant bits)
__device__ __forceinline__ void split4(const float (&x)[4], uint2 &hi, uint2 &lo) {
    const __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(x[0] - f0.x, x[1] - f0.y), l1 = __floats2half2_rn(x[2] - f1.x, x[3] - f1.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
}

// Candidate operand (M side) of a unit: warp w streams the row groups w, w + 7, w + 14 of every stage with cp.async --
// lane = (row in group, 16-byte chunk), one instruction moves the 64 contiguous bytes of a stage for 8 operand rows.
// cp.async.mbarrier.arrive.noinc publishes a stage when this thread's copies have landed (no thread waits for the data).
struct ACopy {
    uint32_t on[3], ot[3];       // element offsets of this thread's 3 operand rows inside cand16 / ctab16 (0xffffffff: unused)
    uint32_t dst;
    int rsub, ch;
    __device__ __forceinline__ void init(unsigned char *base, int tid, const int *cnews, const int *ctab, int cnt) {
        rsub = (tid >> 2) & 7;
        ch = tid & 3;
        const int g0 = tid >> 5;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int g = g0 + kCWarps * i;
            int c = kTile, k = 0;
            if (g < 16) m_row_owner(g >> 2, 8 * (g & 3) + rsub, c, k);
            const bool rok = c < cnt;
            const int cc = rok ? c : 0;
            on[i] = rok ? (uint32_t)cnews[cc] * (uint32_t)kC16 + (uint32_t)(k * 2 * kD) : 0xffffffffu;
            ot[i] = (uint32_t)ctab[cc] * (uint32_t)kC16 + (uint32_t)(k * 2 * kD);
        }
        dst = tc::smem_u32(base) + OFF_A + (uint32_t)g0 * 1024u + (uint32_t)rsub * 128u;
    }
    // stage kc -> ring slot kc & 1 (waits until the MMAs of the slot's previous use have drained it)
    __device__ __forceinline__ void issue(int kc, const __half *cand16, const __half *ctab16, uint64_t *bar_full,
                                          uint64_t *bar_free, uint32_t &use0, uint32_t &use1) const {
        const int s = kc & 1;
        const uint32_t uses = s ? use1 : use0;
        if (uses >= 1) tc::mbar_wait(bar_free + s, (uses - 1) & 1u);
        if (kc < kStages - 1 || ch < (kD - 32 * (kStages - 1)) / 8) {
            const uint32_t d = dst + (uint32_t)(((4 * s + ch) ^ rsub) << 4);
            const int eo = 32 * kc + 8 * ch;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (on[i] != 0xffffffffu) {
                    const uint32_t di = d + (uint32_t)(kCWarps * i) * 1024u;
                    cp_async16(di, cand16 + on[i] + eo);
                    cp_async16(di + kAImg, cand16 + on[i] + kD + eo);
                    cp_async16(di + 2 * kAImg, ctab16 + ot[i] + eo);
                    cp_async16(di + 3 * kAImg, ctab16 + ot[i] + kD + eo);
                }
            }
        }
        cp_async_mbar_arrive_noinc(bar_full + s);
        if (s) ++use1; else ++use0;
    }
};

// Operand production of one pass: NODES rows of O per unique history row u0 <= u < u0 + nrows (fp16 hi / lo
// images, K stages of 32 dims through the two halves of the 64-dim tile) and the node sums (sum o, sum o^2).
// Executed by the 256 compute threads; a thread owns 4 dims of a row per stage (8 threads per row) and, when
// the pass holds more than 28 rows, the same dims of row + 28.  use0/use1 count the uses of the two ring
// halves over the kernel's lifetime (mbarrier phases).
template <int NODES, int TPT>
__device__ __forceinline__ void produce_operands(unsigned char *base, const LimeNewsCache &C, int u0, int nrows, int tid,
                                                 const int *unews, const int *utab, const float *mid_s,
                                                 const float *whalf_s, float *s01_s, int *flag_s, const int *cnews,
                                                 const int *ctab, int cnt, const __half *cand16, const __half *ctab16,
                                                 uint64_t *bar_full, uint64_t *bar_free, uint32_t &use0, uint32_t &use1,
                                                 int pre_issued) {
    const int sub = tid & 7;
    ACopy ac;
    ac.init(base, tid, cnews, ctab, cnt);
    for (int kc = pre_issued; kc < 1; ++kc) ac.issue(kc, cand16, ctab16, bar_full, bar_free, use0, use1);
    bool ok[TPT];
    const float *hrow[TPT], *trow[TPT];
    float aj[TPT][NODES], ps[TPT][2 * NODES];
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        const int ul = (tid >> 3) + kRPT * t;
        ok[t] = ul < nrows;
        const int u = u0 + (ok[t] ? ul : 0);
        const float mid = mid_s[u], wh = whalf_s[u];
#pragma unroll
        for (int j = 0; j < NODES; ++j) aj[t][j] = fmaf(wh, node_x<NODES>(j), mid);
        hrow[t] = C.hist_rows + (size_t)unews[u] * LIME_HIST_LD;
        trow[t] = C.hist_tab + (size_t)utab[u] * LIME_HTAB_LD;
#pragma unroll
        for (int i = 0; i < 2 * NODES; ++i) ps[t][i] = 0.0f;
    }
    // the first row of a thread: stage kc + 1's global loads are in flight while stage kc is evaluated; the second row
    // (passes with more than 28 rows): loaded at the start of its stage, consumed after the first row's arithmetic
    float4 nx[4], nx1[4];
    float4 nbias = ldg4(C.gate_bias + 4 * sub);      // gate bias' of the next stage's 4 dims (prefetched like the rows)
    nx[0] = ldg4(hrow[0] + LIME_HIST_VC + 4 * sub);
    nx[1] = ldg4(trow[0] + 4 * sub);
    nx[2] = ldg4(hrow[0] + LIME_HIST_GW + 4 * sub);
    nx[3] = ldg4(trow[0] + kD + 4 * sub);
    for (int kc = 0; kc < kStages; ++kc) {
        const int s = kc & 1;
        const int d0 = 32 * kc + 4 * sub;
        float v[4], gg[4];
        v[0] = (nx[0].x + nx[1].x) * kOScale; v[1] = (nx[0].y + nx[1].y) * kOScale;
        v[2] = (nx[0].z + nx[1].z) * kOScale; v[3] = (nx[0].w + nx[1].w) * kOScale;
        gg[0] = nx[2].x + nx[3].x; gg[1] = nx[2].y + nx[3].y;
        gg[2] = nx[2].z + nx[3].z; gg[3] = nx[2].w + nx[3].w;
        const float4 bias_now = nbias;
        if (TPT == 2 && d0 < kD) {
            nx1[0] = ldg4(hrow[TPT - 1] + LIME_HIST_VC + d0);
            nx1[1] = ldg4(trow[TPT - 1] + d0);
            nx1[2] = ldg4(hrow[TPT - 1] + LIME_HIST_GW + d0);
            nx1[3] = ldg4(trow[TPT - 1] + kD + d0);
        }
        if (d0 + 32 < kD) {
            nbias = ldg4(C.gate_bias + d0 + 32);
            nx[0] = ldg4(hrow[0] + LIME_HIST_VC + d0 + 32);
            nx[1] = ldg4(trow[0] + d0 + 32);
            nx[2] = ldg4(hrow[0] + LIME_HIST_GW + d0 + 32);
            nx[3] = ldg4(trow[0] + kD + d0 + 32);
        }
        if (d0 < kD) {      // the slot is free: this thread waited for it when it issued the stage's copies
            const float bb[4] = {bias_now.x, bias_now.y, bias_now.z, bias_now.w};
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                if (t == 1) {
                    v[0] = (nx1[0].x + nx1[1].x) * kOScale; v[1] = (nx1[0].y + nx1[1].y) * kOScale;
                    v[2] = (nx1[0].z + nx1[1].z) * kOScale; v[3] = (nx1[0].w + nx1[1].w) * kOScale;
                    gg[0] = nx1[2].x + nx1[3].x; gg[1] = nx1[2].y + nx1[3].y;
                    gg[2] = nx1[2].z + nx1[3].z; gg[3] = nx1[2].w + nx1[3].w;
                }
                if (ok[t]) {
                    const int nrow0 = NODES * ((tid >> 3) + kRPT * t);
#pragma unroll
                    for (int j = 0; j < NODES; ++j) {
                        // o = v (1 - (1 - a_j) sigmoid(a_j W_g v + b_g)),  sigmoid(z) = 1 / (1 + 2^z'),  z' = -log2(e) z
                        const float a = aj[t][j], oma = 1.0f - a;
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float den = ex2_approx(fmaf(a, gg[e], bb[e])) + 1.0f;
                            o[e] = fmaf(-(v[e] * oma), rcp_approx(den), v[e]);
                            ps[t][2 * j] += o[e];
                            ps[t][2 * j + 1] = fmaf(o[e], o[e], ps[t][2 * j + 1]);
                        }
                        uint2 hi, lo;
                        split4(o, hi, lo);
                        const uint32_t off = tc::sw128_offset(nrow0 + j, 4 * s + (sub >> 1)) + 8u * (sub & 1);
                        *reinterpret_cast<uint2 *>(base + OFF_BHI + off) = hi;
                        *reinterpret_cast<uint2 *>(base + OFF_BLO + off) = lo;
                    }
                }
            }
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(bar_full + s);
        if (kc + 1 < kStages && kc + 1 >= pre_issued) ac.issue(kc + 1, cand16, ctab16, bar_full, bar_free, use0, use1);
    }
    // node sums of the row: sum o, sum o^2 per node, over the 8 lanes that share the row
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
            for (int i = 0; i < 2 * NODES; ++i) ps[t][i] += __shfl_xor_sync(0xffffffffu, ps[t][i], o);
        }
        if (ok[t] && sub == 0) {
            const int u = u0 + (tid >> 3) + kRPT * t;
            bool in_range = true;
#pragma unroll
            for (int j = 0; j < NODES; ++j) {
                s01_s[u * 8 + 2 * j] = ps[t][2 * j] * (1.0f / kOScale);
                s01_s[u * 8 + 2 * j + 1] = ps[t][2 * j + 1] * (1.0f / (kOScale * kOScale));
                in_range = in_range && (ps[t][2 * j + 1] <= kS1Max);
            }
            if (!in_range) atomicOr(flag_s, 4);   // outside the fp16 operand range (or NaN): exact kernel
        }
    }
}

// Epilogue of one pass: a compute thread owns one accumulator row (TMEM lane) = one (candidate, k) and walks
// the columns (unique history row, node) of the pass; warps w and w + 4 share a TMEM quadrant and alternate
// over the 16-column blocks (quadrant 3, the candidates beyond 32, has warp 3 alone).  Writes out[k][u][c]  (k = 0: pooling logit, 1: pooled value, 2: GraphSAGE term).
template <int NODES>
__device__ __forceinline__ void epilogue(uint32_t tmem, int warp, int lane, int u0, int nrows, int cnt, float ln_eps,
                                         const float *a_s, const float *mid_s, const float *winv_s, const float *s01_s,
                                         const float *cscal, float *out_s) {
    constexpr int UPB = 16 / NODES;               // unique rows per 16-column block
    const int qd = warp & 3;
    int c, k;
    m_row_owner(qd, lane, c, k);
    const bool valid = c < cnt;
    const int cc = valid ? c : 0;
    const float bk = cscal[cc * 4 + k];
    float *outk = out_s + k * (kH * kAS);
    const uint32_t taddr = tmem + ((uint32_t)(32 * qd) << 16);
    const int nblocks = (NODES * nrows + 15) >> 4;
    const int bstep = qd + 4 < kCWarps ? 2 : 1;      // quadrants whose partner warp w + 4 is the MMA issuer are walked by one warp
    for (int b = bstep == 2 ? warp >> 2 : 0; b < nblocks; b += bstep) {
      {
        // the two accumulators of the block (news part, table part): both loads are issued before the single wait
        uint32_t rn[16], rt[16];
        tc::tmem_ld16_nowait(taddr + 16 * b, rn);
        tc::tmem_ld16_nowait(taddr + 128 + 16 * b, rt);
        tc::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rn[i]) + __uint_as_float(rt[i]);
#pragma unroll
        for (int i = 0; i < UPB; ++i) {
            const int ul = UPB * b + i;
            if (ul < nrows) {
                const int u = u0 + ul;
                float t = (a_s[u * kAS + cc] - mid_s[u]) * winv_s[u];
                t = fminf(fmaxf(t, -1.0f), 1.0f);
                float p, s0, s1;
                if (NODES == 2) {
                    const float4 ss = *reinterpret_cast<const float4 *>(s01_s + u * 8);
                    const float l1 = fmaf(t, 0.5f / kY1, 0.5f), l0 = 1.0f - l1;      // (t - y0) / (y1 - y0)
                    p = fmaf(l1, v[2 * i + 1], l0 * v[2 * i]);
                    s0 = fmaf(l1, ss.z, l0 * ss.x);
                    s1 = fmaf(l1, ss.w, l0 * ss.y);
                } else {
                    const float4 sa = *reinterpret_cast<const float4 *>(s01_s + u * 8);
                    const float4 sb = *reinterpret_cast<const float4 *>(s01_s + u * 8 + 4);
                    const float t0 = t - kX0, t1 = t - kX1, t2 = t - kX2, t3 = t - kX3;
                    const float l0 = t1 * t2 * t3 * (1.0f / ((kX0 - kX1) * (kX0 - kX2) * (kX0 - kX3)));
                    const float l1 = t0 * t2 * t3 * (1.0f / ((kX1 - kX0) * (kX1 - kX2) * (kX1 - kX3)));
                    const float l2 = t0 * t1 * t3 * (1.0f / ((kX2 - kX0) * (kX2 - kX1) * (kX2 - kX3)));
                    const float l3 = t0 * t1 * t2 * (1.0f / ((kX3 - kX0) * (kX3 - kX1) * (kX3 - kX2)));
                    p = l0 * v[4 * i] + l1 * v[4 * i + 1] + l2 * v[4 * i + 2] + l3 * v[4 * i + 3];
                    s0 = l0 * sa.x + l1 * sa.z + l2 * sb.x + l3 * sb.z;
                    s1 = l0 * sa.y + l1 * sa.w + l2 * sb.y + l3 * sb.w;
                }
                // LayerNorm folded into the dot: the candidate vectors are mean-centred, so x.w = rstd * (o.w)
                const float mu = s0 * (1.0f / kD);
                const float var = fmaxf(fmaf(-mu, mu, s1 * (1.0f / kD)), 0.0f);
                const float rstd = rsqrtf(var + ln_eps) * kUnscale;
                if (valid) outk[u * kAS + c] = fmaf(rstd, p, bk);
            }
        }
      }
    }
}

// Candidate-aware attention weights a[u][c] (layers.py:66-81) of one work unit.  LPC lanes share a candidate, a lane owns
// the unique rows u = l + LPC * i (i < 4): all 12 table loads of a lane are issued before the first exponential, the 40
// exponentials stay in registers for both softmaxes.  Head logits come pre-scaled by log2(e) from the topic-pair table
// (no max pass: the host checks the table's |logit| bound); masked slots (mask == 0 -> -1e9, layers.py:72) contribute
// exactly 0 unless every slot is masked, in which case both softmaxes are uniform over the H slots.
template <int LPC>
__device__ __forceinline__ void attention(const LimeNewsCache &C, int T, int U, int cnt, int nun, int warp, int lane,
                                          const int *ctopic, const int *utopic, const int *umask, const float *umult,
                                          float *a_s) {
    constexpr int CPW = 32 / LPC;                 // candidates per warp and round
    const int l = lane & (LPC - 1), g = lane / LPC;
    float w[4], mu[4];
    int tp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int u = l + LPC * i;
        const bool in = u < U;
        mu[i] = in ? umult[u] : 0.0f;
        w[i] = (in && umask[u] != 0) ? mu[i] : 0.0f;
        tp[i] = in ? utopic[u] : 0;
    }
    for (int c0 = CPW * warp; c0 < cnt; c0 += CPW * kCWarps) {
        const int c = c0 + g;
        const bool cvalid = c < cnt;
        const int cc = cvalid ? c : cnt - 1;
        float e2[4];
        float s2 = 0.0f;
        if (nun > 0) {
            const float *trow = C.topic_table + (size_t)ctopic[cc] * T * kTabLd;
            float4 x[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float *r0 = trow + (size_t)tp[i] * kTabLd;
                x[i][0] = ldg4(r0);
                x[i][1] = ldg4(r0 + 4);
                x[i][2] = ldg4(r0 + 8);
            }
            float e[4][LIME_CA_HEADS], sum[LIME_CA_HEADS];
#pragma unroll
            for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xs[LIME_CA_HEADS] = {x[i][0].x, x[i][0].y, x[i][0].z, x[i][0].w, x[i][1].x,
                                                 x[i][1].y, x[i][1].z, x[i][1].w, x[i][2].x, x[i][2].y};
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) {
                    e[i][hd] = ex2_approx(xs[hd]);
                    sum[hd] = fmaf(w[i], e[i][hd], sum[hd]);
                }
            }
#pragma unroll
            for (int o = 1; o < LPC; o <<= 1) {
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] += __shfl_xor_sync(0xffffffffu, sum[hd], o);
            }
#pragma unroll
            for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = __fdividef(1.0f, sum[hd]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float agg = 0.0f;
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) agg = fmaf(e[i][hd], sum[hd], agg);
                agg = w[i] > 0.0f ? agg : 0.0f;
                // second, unmasked softmax over the history (layers.py:81); agg in [0, 10]: no max needed
                e2[i] = ex2_approx(agg * kLog2e);
                s2 = fmaf(mu[i], e2[i], s2);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                e2[i] = 1.0f;
                s2 += mu[i];
            }
        }
#pragma unroll
        for (int o = 1; o < LPC; o <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        const float inv2 = __fdividef(1.0f, s2);
        if (cvalid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int u = l + LPC * i;
                if (u < U) a_s[u * kAS + c] = e2[i] * inv2;
            }
        }
    }
}

// Front end of one work unit, executed by ONE warp (the MMA issuer) while the compute warps work on the previous unit.
// Part 1, during its attention phase: the history slots (keys, bucket pairs, topic ids, gate bounds) into the front-end
// scratch.  Parts 2 and 3, during its epilogue: dedup of the slots, then the candidates (cache rows, lifetime weights,
// folded scalars).  All results land in the unit buffer `ub`.
__device__ __forceinline__ void front_hist(const ScoreArgs &args, unsigned char *ub, int unit, int lane, int *hkn, int *hkt,
                                           int *htp, float *hga) {
    const LimeNewsCache &C = args.cache;
    const LimeImpressions &I = args.imp;
    const int H = I.max_history, nb = C.num_buckets, T = C.num_topics;
    int *info = reinterpret_cast<int *>(ub + UB_INFO);
    if (unit >= I.num_units) {
        if (lane == 0) info[UI_UNIT] = unit;
        __syncwarp();
        return;
    }
    const int imp = I.unit_imp[unit], pair0 = I.unit_pair0[unit], cnt = I.unit_count[unit];
    for (int h = lane; h < H; h += 32) {
        const long long o = (long long)imp * H + h;
        int n = I.hist_news[o];
        n = (n < 0 || n >= C.news_num) ? 0 : n;
        const float2 mt = __ldg(reinterpret_cast<const float2 *>(C.news_meta + (size_t)n * LIME_META_LD));
        const int tp = __float_as_int(mt.x);
        const int mk = I.hist_mask[o] != 0 ? 1 : 0;
        const int bf = bucketize_seconds(I.hist_fresh[o], args.bucket_scale, nb);
        const int bl = bucketize_seconds(I.hist_life[o], args.bucket_scale, nb);
        hkn[h] = n;
        hkt[h] = 2 * (bf * nb + bl) + mk;          // second key word: bucket pair and mask
        htp[h] = (tp < 0 || tp >= T) ? 0 : tp;
        hga[h] = mt.y;
    }
    // the candidate arrays of the unit: L2 by the time part 3 reads them
    for (int j = lane; 32 * j < cnt; j += 32) {
        prefetch_l2(I.cand_news + pair0 + 32 * j);
        prefetch_l2(I.cand_fresh + pair0 + 32 * j);
        prefetch_l2(I.cand_life + pair0 + 32 * j);
        if (I.cand_remaining) prefetch_l2(I.cand_remaining + pair0 + 32 * j);
    }
    if (lane == 0) {
        info[UI_UNIT] = unit;
        info[UI_IMP] = imp;
        info[UI_PAIR0] = pair0;
        info[UI_CNT] = cnt;
    }
    __syncwarp();
}

__device__ __forceinline__ void front_cand(const ScoreArgs &args, unsigned char *ub, int lane) {
    const LimeNewsCache &C = args.cache;
    const LimeImpressions &I = args.imp;
    const int nb = C.num_buckets, T = C.num_topics;
    int *info = reinterpret_cast<int *>(ub + UB_INFO);
    if (info[UI_UNIT] >= I.num_units) return;
    float *cscal = reinterpret_cast<float *>(ub + UB_CSCAL);
    float *cw = reinterpret_cast<float *>(ub + UB_CW);
    int *cnews = reinterpret_cast<int *>(ub + UB_CNEWS);
    int *ctab = reinterpret_cast<int *>(ub + UB_CTAB);
    int *cP = reinterpret_cast<int *>(ub + UB_CP);
    int *ctopic = reinterpret_cast<int *>(ub + UB_CTOPIC);
    const int pair0 = info[UI_PAIR0], cnt = info[UI_CNT];
    int flags = 0;
    for (int c = lane; c < cnt; c += 32) {
        const long long p = (long long)pair0 + c;
        int n = I.cand_news[p];
        n = (n < 0 || n >= C.news_num) ? 0 : n;
        const float fr = I.cand_fresh[p], lf = I.cand_life[p];
        const float4 m0 = ldg4(C.news_meta + (size_t)n * LIME_META_LD), m1 = ldg4(C.news_meta + (size_t)n * LIME_META_LD + 4);
        const int tb = bucketize_seconds(fr, args.bucket_scale, nb) * nb + bucketize_seconds(lf, args.bucket_scale, nb);
        const float *ctr = C.cand_tab + (size_t)tb * LIME_CTAB_LD;
        const int tp = __float_as_int(m0.x);
        cnews[c] = n;
        ctab[c] = tb;
        cw[c] = lifetime_weight(I.cand_remaining ? I.cand_remaining[p] : __fsub_rn(lf, fr), C);
        cP[c] = (args.pair_index_base + p >= args.tail_start) ? args.prefix_tail : args.prefix_main;
        ctopic[c] = (tp < 0 || tp >= T) ? 0 : tp;
        cscal[c * 4 + 0] = m0.w + __ldg(ctr + LIME_CAND_SCAL + 3);
        cscal[c * 4 + 1] = m1.x + __ldg(ctr + LIME_CAND_SCAL + 4);
        cscal[c * 4 + 2] = m1.y + __ldg(ctr + LIME_CAND_SCAL + 5);
        cscal[c * 4 + 3] = m1.z + __ldg(ctr + LIME_CAND_SCAL + 6);
        if (!(m0.z <= kWAbsMax)) flags = 4;        // beyond the fp16 operand range: exact kernel
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) info[UI_FLAGS] = flags;
    __syncwarp();
}

// Part 2, during the previous unit's epilogue: deduplication of the history slots into unique operand rows.  Slots with
// equal (news, bucket pair, mask) are one row.  A lane owns the slots lane and lane + 32 and scans all H keys (broadcast
// reads, no cross-lane dependency): lowest equal slot = the unique row, number of equal slots = its multiplicity (overall
// and inside the two GraphSAGE prefixes); one ballot pair then compacts the unique rows in slot order.
__device__ __forceinline__ void front_dedup(const ScoreArgs &args, unsigned char *ub, int lane, const int *hkn, const int *hkt,
                                            const int *htp, const float *hga) {
    int *info = reinterpret_cast<int *>(ub + UB_INFO);
    if (info[UI_UNIT] >= args.imp.num_units) return;
    const int H = args.imp.max_history;
    int *unews = reinterpret_cast<int *>(ub + UB_UNEWS);
    int *utab = reinterpret_cast<int *>(ub + UB_UTAB);
    int *umask = reinterpret_cast<int *>(ub + UB_UMASK);
    int *utopic = reinterpret_cast<int *>(ub + UB_UTOPIC);
    float *umult = reinterpret_cast<float *>(ub + UB_UMULT);
    float *ump0 = reinterpret_cast<float *>(ub + UB_UMP0);
    float *ump1 = reinterpret_cast<float *>(ub + UB_UMP1);
    float *ugabs = reinterpret_cast<float *>(ub + UB_UGABS);
    const int pz0 = args.prefix_main < H ? args.prefix_main : H, pz1 = args.prefix_tail < H ? args.prefix_tail : H;
    const int ha = lane, hb = lane + 32;
    const int k1a = ha < H ? hkn[ha] : -1 - ha, k2a = ha < H ? hkt[ha] : -1;
    const int k1b = hb < H ? hkn[hb] : -1 - hb, k2b = hb < H ? hkt[hb] : -1;
    int fa = 99, fb = 99, na = 0, nb_ = 0, na0 = 0, nb0 = 0, na1 = 0, nb1 = 0;
#pragma unroll 4
    for (int j = H - 1; j >= 0; --j) {          // downwards: the last hit is the lowest equal slot
        const int n = hkn[j], t = hkt[j];
        const bool ea = k1a == n && k2a == t, eb = k1b == n && k2b == t;
        fa = ea ? j : fa;
        fb = eb ? j : fb;
        na += ea;
        nb_ += eb;
        na0 += ea && j < pz0;
        nb0 += eb && j < pz0;
        na1 += ea && j < pz1;
        nb1 += eb && j < pz1;
    }
    const bool isfa = ha < H && fa == ha, isfb = hb < H && fb == hb;
    const unsigned b0 = __ballot_sync(0xffffffffu, isfa), b1 = __ballot_sync(0xffffffffu, isfb);
    const unsigned lt = (1u << lane) - 1u;
    if (isfa) {
        const int u = __popc(b0 & lt);
        unews[u] = k1a;
        utab[u] = k2a >> 1;
        umask[u] = k2a & 1;
        utopic[u] = htp[ha];
        ugabs[u] = hga[ha];
        umult[u] = (float)na;
        ump0[u] = (float)na0;
        ump1[u] = (float)na1;
    }
    if (isfb) {
        const int u = __popc(b0) + __popc(b1 & lt);
        unews[u] = k1b;
        utab[u] = k2b >> 1;
        umask[u] = k2b & 1;
        utopic[u] = htp[hb];
        ugabs[u] = hga[hb];
        umult[u] = (float)nb_;
        ump0[u] = (float)nb0;
        ump1[u] = (float)nb1;
    }
    int nun = (ha < H ? k2a & 1 : 0) + (hb < H ? k2b & 1 : 0);      // unmasked history slots
    nun = __reduce_add_sync(0xffffffffu, nun);
    if (lane == 0) {
        info[UI_U] = __popc(b0) + __popc(b1);
        info[UI_NUN] = nun;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kThreads, 2) score_tc_kernel(const ScoreArgs args) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw;
    float *out_s = reinterpret_cast<float *>(base + OFF_OUT);
    float *a_s = reinterpret_cast<float *>(base + OFF_AS);
    float *s01_s = reinterpret_cast<float *>(base + OFF_S01);
    float *mid_s = reinterpret_cast<float *>(base + OFF_MID);
    float *winv_s = reinterpret_cast<float *>(base + OFF_WINV);
    float *whalf_s = reinterpret_cast<float *>(base + OFF_WHALF);
    float *pool_s = reinterpret_cast<float *>(base + OFF_POOL);
    int *hkn = reinterpret_cast<int *>(base + OFF_HKN);
    int *hkt = reinterpret_cast<int *>(base + OFF_HKT);
    int *htp = reinterpret_cast<int *>(base + OFF_HTP);
    float *hga = reinterpret_cast<float *>(base + OFF_HGA);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(base + OFF_BARS);
    uint64_t *bar_free = bar_full + 2;
    uint64_t *bar_accum = bar_full + 4;
    volatile int *misc = reinterpret_cast<volatile int *>(base + OFF_MISC);

    const LimeNewsCache &C = args.cache;
    const LimeImpressions &I = args.imp;
    const int H = I.max_history;
    const int T = C.num_topics;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __half *cand16 = reinterpret_cast<const __half *>(C.cand16);
    const __half *ctab16 = reinterpret_cast<const __half *>(C.ctab16);

    if (tid == 0) {
        tc::mbar_init(bar_full + 0, 2 * kCompute);    // per compute thread: its cp.async copies + its O rows
        tc::mbar_init(bar_full + 1, 2 * kCompute);
        tc::mbar_init(bar_free + 0, 1);
        tc::mbar_init(bar_free + 1, 1);
        tc::mbar_init(bar_accum, 1);
        tc::mbar_fence_init();
    }
    int next_unit = 0;               // issuer warp: the unit whose front end it runs next
    if (warp == kMmaWarp) {
        tc::tmem_alloc(reinterpret_cast<uint32_t *>(base + OFF_MISC) + M_TMEM, 256);
        int u0 = 0;
        if (lane == 0) u0 = atomicAdd(args.work_counter, 1);
        u0 = __shfl_sync(0xffffffffu, u0, 0);
        front_hist(args, base + OFF_UB, u0, lane, hkn, hkt, htp, hga);
        front_dedup(args, base + OFF_UB, lane, hkn, hkt, htp, hga);
        front_cand(args, base + OFF_UB, lane);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *(reinterpret_cast<volatile uint32_t *>(base + OFF_MISC) + M_TMEM);

#ifdef LIME_TC_PHASE_CLOCKS
    unsigned long long *prof = reinterpret_cast<unsigned long long *>(base + OFF_PROF);
    if (tid == 0) {
        for (int i = 0; i < 16; ++i) prof[i] = 0;
    }
    long long t_last = clock64();
#endif
    uint32_t use0 = 0, use1 = 0;     // uses of the two ring halves so far (every role counts the same sequence)
    uint32_t pass_iter = 0;          // passes processed so far (phase of bar_accum)
    int ubi = 0;                     // unit buffer of the current unit

    for (;; ubi ^= 1) {
        unsigned char *ub = base + OFF_UB + ubi * kUnitBuf;
        volatile int *info = reinterpret_cast<volatile int *>(ub + UB_INFO);
        const float *cscal = reinterpret_cast<const float *>(ub + UB_CSCAL);
        const float *cw = reinterpret_cast<const float *>(ub + UB_CW);
        const int *cnews = reinterpret_cast<const int *>(ub + UB_CNEWS);
        const int *ctab = reinterpret_cast<const int *>(ub + UB_CTAB);
        const int *cP = reinterpret_cast<const int *>(ub + UB_CP);
        const int *ctopic = reinterpret_cast<const int *>(ub + UB_CTOPIC);
        const int *unews = reinterpret_cast<const int *>(ub + UB_UNEWS);
        const int *utab = reinterpret_cast<const int *>(ub + UB_UTAB);
        const int *umask = reinterpret_cast<const int *>(ub + UB_UMASK);
        const int *utopic = reinterpret_cast<const int *>(ub + UB_UTOPIC);
        const float *umult = reinterpret_cast<const float *>(ub + UB_UMULT);
        const float *ump0 = reinterpret_cast<const float *>(ub + UB_UMP0);
        const float *ump1 = reinterpret_cast<const float *>(ub + UB_UMP1);
        const float *ugabs = reinterpret_cast<const float *>(ub + UB_UGABS);
        int *flag_s = reinterpret_cast<int *>(ub + UB_INFO) + UI_FLAGS;

        const int unit = info[UI_UNIT];
        if (unit >= I.num_units) break;
#ifdef LIME_TC_PHASE_CLOCKS
        if (tid == 0) { t_last = clock64(); ++prof[9]; }
#endif
        const int pair0 = info[UI_PAIR0], cnt = info[UI_CNT], U = info[UI_U];
        for (int c = tid; c < cnt; c += kThreads) {
            pool_s[c * 4 + 0] = -INFINITY;
            pool_s[c * 4 + 1] = 0.0f;
            pool_s[c * 4 + 2] = 0.0f;
            pool_s[c * 4 + 3] = 0.0f;
        }

        // ================= roles ======================================================================
        if (warp < kCWarps) {
            // the candidate operand of the first two stages goes out now (both ring slots are free since the previous
            // unit's epilogue): it lands during the attention phase
            {
                ACopy ac;
                ac.init(base, tid, cnews, ctab, cnt);
                ac.issue(0, cand16, ctab16, bar_full, bar_free, use0, use1);
                ac.issue(1, cand16, ctab16, bar_full, bar_free, use0, use1);
            }
            // ---------------- phase 1: candidate-aware attention weights a[u][c] (layers.py:66-81) ------
            // lanes per candidate = the smallest power of two that covers the U rows with 4 rows per lane: short (deduplicated)
            // histories put more candidates into a round (7 warps x 32 / LPC), and a round costs one table-load latency
            if (U <= 8)       attention<2>(C, T, U, cnt, info[UI_NUN], warp, lane, ctopic, utopic, umask, umult, a_s);
            else if (U <= 16) attention<4>(C, T, U, cnt, info[UI_NUN], warp, lane, ctopic, utopic, umask, umult, a_s);
            else if (U <= 32) attention<8>(C, T, U, cnt, info[UI_NUN], warp, lane, ctopic, utopic, umask, umult, a_s);
            else              attention<16>(C, T, U, cnt, info[UI_NUN], warp, lane, ctopic, utopic, umask, umult, a_s);
            bar_compute();
            LIME_TICK(2);

            // ---------------- interpolation nodes per unique row (4 lanes per row) --------------------
            {
                const int u = tid >> 2, l4 = tid & 3;
                const int uc = u < U ? u : U - 1;
                float lo = INFINITY, hi = -INFINITY;
                for (int c = l4; c < cnt; c += 4) {
                    const float a = a_s[uc * kAS + c];
                    lo = fminf(lo, a);
                    hi = fmaxf(hi, a);
                }
#pragma unroll
                for (int o = 1; o < 4; o <<= 1) {
                    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                if (u < U && l4 == 0) {
                    const float wh = fmaxf(0.5f * (hi - lo), 1e-7f);
                    mid_s[u] = 0.5f * (hi + lo);
                    whalf_s[u] = wh;
                    winv_s[u] = 1.0f / wh;
                    // Interpolation error of f(a) = (1 - a) sigmoid(g a + b) on [mid - w, mid + w] with n Chebyshev
                    // nodes: max|d^n f| w^n / (n! 2^(n-1));  |d^2 f| <= 0.0962 g^2 + 0.5 |g|,
                    // |d^4 f| <= 0.125 g^4 + 0.5 |g|^3.  g is bounded by the cached max |W_g vc| of the news plus
                    // the max over the bucket-pair table.
                    const float gabs = (ugabs[u] + C.tab_gw_absmax) * (1.0f / kLog2e);
                    const float w2 = wh * wh, g2 = gabs * gabs;
                    const float err2 = w2 * (0.0962f * g2 + 0.5f * gabs) * 0.25f;
                    const float err4 = w2 * w2 * (0.125f * g2 * g2 + 0.5f * g2 * gabs) * (1.0f / 192.0f);
                    if (!(err2 <= args.interp_tol)) atomicOr(flag_s, 1);          // 2 nodes are not enough
                    if (!(err4 <= args.interp_tol)) atomicOr(flag_s, 2);          // 4 nodes are not enough: exact kernel
                }
            }
            bar_compute();
            if (tid == 0) {
                const int nd = (flag_s[0] & 1) ? 4 : 2;
                misc[M_NODES] = nd;
                misc[M_NPASS] = (nd * U + kBRows - 1) / kBRows;
            }
            bar_compute();
            LIME_TICK(3);
        } else {
            // issuer warp: claim the next work unit and run parts 1 and 2 of its front end while the compute warps are in
            // phase 1 (part 3 runs after this unit's last MMA)
            if (lane == 0) next_unit = atomicAdd(args.work_counter, 1);
            next_unit = __shfl_sync(0xffffffffu, next_unit, 0);
            front_hist(args, base + OFF_UB + (ubi ^ 1) * kUnitBuf, next_unit, lane, hkn, hkt, htp, hga);
            front_dedup(args, base + OFF_UB + (ubi ^ 1) * kUnitBuf, lane, hkn, hkt, htp, hga);
        }
        // the number of passes is known to the issuer at the first CTA barrier below; pass 0 always exists
        int npass = 1;
        for (int pass = 0; pass < npass; ++pass) {
            if (warp < kCWarps) {
                const int nodes = misc[M_NODES];
                const int G = kBRows / nodes;
                const int u0 = pass * G;
                const int nrows = min(G, U - u0);
                if (tid == 0) misc[M_NPAD] = (nodes * nrows + 15) & ~15;      // published by the first full-barrier arrive
                if (nodes == 2) {
                    if (nrows > kRPT) produce_operands<2, 2>(base, C, u0, nrows, tid, unews, utab, mid_s, whalf_s, s01_s, flag_s, cnews, ctab, cnt, cand16, ctab16, bar_full, bar_free, use0, use1, pass == 0 ? 2 : 0);
                    else              produce_operands<2, 1>(base, C, u0, nrows, tid, unews, utab, mid_s, whalf_s, s01_s, flag_s, cnews, ctab, cnt, cand16, ctab16, bar_full, bar_free, use0, use1, pass == 0 ? 2 : 0);
                } else {
                    produce_operands<4, 1>(base, C, u0, nrows, tid, unews, utab, mid_s, whalf_s, s01_s, flag_s, cnews, ctab, cnt, cand16, ctab16, bar_full, bar_free, use0, use1, pass == 0 ? 2 : 0);
                }
            } else {
                // ---------------- MMA issuer -------------------------------------------------------------
                const uint32_t sb = tc::smem_u32(base);
                uint32_t idesc = 0;
                for (int kc = 0; kc < kStages; ++kc) {
                    const int s = kc & 1;
                    const uint32_t uses = s ? use1 : use0;
                    tc::mbar_wait(bar_full + s, uses & 1u);
                    tc::fence_after_sync();
                    if (kc == 0) idesc = tc::idesc_f16_f32(128, misc[M_NPAD]);
                    if (lane == 0) {
                        const int ksteps = kc < kStages - 1 ? 2 : (kD - 32 * (kStages - 1)) / 16;
                        const uint64_t bhi = tc::smem_desc_sw128(sb + OFF_BHI), blo = tc::smem_desc_sw128(sb + OFF_BLO);
                        const uint64_t a_nh = tc::smem_desc_sw128(sb + OFF_A), a_nl = tc::smem_desc_sw128(sb + OFF_A + kAImg);
                        const uint64_t a_th = tc::smem_desc_sw128(sb + OFF_A + 2 * kAImg), a_tl = tc::smem_desc_sw128(sb + OFF_A + 3 * kAImg);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t k2 = (uint64_t)(2 * (2 * s + ks));   // 32 bytes per K step of 16
                            // two accumulators (news part: columns 0.., table part: columns 128..): half as many
                            // accumulation steps each; the epilogue adds them in fp32
                            tc::mma_f16(tmem, a_nl + k2, bhi + k2, idesc, (kc | ks) != 0);
                            tc::mma_f16(tmem, a_nh + k2, blo + k2, idesc, true);
                            tc::mma_f16(tmem, a_nh + k2, bhi + k2, idesc, true);
                            tc::mma_f16(tmem + 128, a_tl + k2, bhi + k2, idesc, (kc | ks) != 0);
                            tc::mma_f16(tmem + 128, a_th + k2, blo + k2, idesc, true);
                            tc::mma_f16(tmem + 128, a_th + k2, bhi + k2, idesc, true);
                        }
                        tc::mma_commit(bar_free + s);
                        if (kc == kStages - 1) tc::mma_commit(bar_accum);
                    }
                    __syncwarp();
                    if (s) ++use1; else ++use0;
                }
            }
            LIME_TICK(4);
            __syncthreads();   // node sums, flags and the pass count are visible to every role
            LIME_TICK(10);
            npass = misc[M_NPASS];

            if (warp < kCWarps) {
                // ---------------- epilogue: TMEM -> Lagrange combination -> LayerNorm folding ---------
                const int nodes = misc[M_NODES];
                const int G = kBRows / nodes;
                const int u0 = pass * G;
                const int nrows = min(G, U - u0);
                tc::mbar_wait(bar_accum, pass_iter & 1u);
                tc::fence_after_sync();
                LIME_TICK(5);
                if (nodes == 2) epilogue<2>(tmem, warp, lane, u0, nrows, cnt, args.ln_eps, a_s, mid_s, winv_s, s01_s, cscal, out_s);
                else            epilogue<4>(tmem, warp, lane, u0, nrows, cnt, args.ln_eps, a_s, mid_s, winv_s, s01_s, cscal, out_s);
                tc::fence_before_sync();
                bar_compute();
                LIME_TICK(6);

                // ---------------- candidate-query pooling over the rows of this pass (online softmax) -----
                // 8 lanes per candidate; multiplicities weight the softmax sum, the pooled value and the mean
                const float *lg_s = out_s, *y_s = out_s + kH * kAS, *z_s = out_s + 2 * kH * kAS;
                const int l8 = lane & 7, g = lane >> 3;
                for (int c0 = 4 * warp; c0 < cnt; c0 += 4 * kCWarps) {
                    const int c = c0 + g;
                    const bool cvalid = c < cnt;
                    const int cc = cvalid ? c : cnt - 1;
                    const float *mp = cP[cc] == args.prefix_main ? ump0 : ump1;
                    float m = -INFINITY;
                    for (int ul = l8; ul < nrows; ul += 8) m = fmaxf(m, lg_s[(u0 + ul) * kAS + cc]);
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                    float l = 0.f, acc = 0.f, ms = 0.f;
                    for (int ul = l8; ul < nrows; ul += 8) {
                        const int u = u0 + ul;
                        const float e = umult[u] * ex2_approx((lg_s[u * kAS + cc] - m) * kLog2e);
                        l += e;
                        acc = fmaf(e, y_s[u * kAS + cc], acc);
                        ms = fmaf(mp[u], z_s[u * kAS + cc], ms);
                    }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        l += __shfl_xor_sync(0xffffffffu, l, o);
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        ms += __shfl_xor_sync(0xffffffffu, ms, o);
                    }
                    if (cvalid && l8 == 0) {
                        const float m_old = pool_s[c * 4 + 0];
                        const float m_new = fmaxf(m_old, m);
                        const float f_old = ex2_approx((m_old - m_new) * kLog2e);    // 0 on the first pass (m_old = -inf)
                        const float f_new = ex2_approx((m - m_new) * kLog2e);
                        const float l_new = fmaf(pool_s[c * 4 + 1], f_old, l * f_new);
                        const float acc_new = fmaf(pool_s[c * 4 + 2], f_old, acc * f_new);
                        const float ms_new = pool_s[c * 4 + 3] + ms;
                        pool_s[c * 4 + 0] = m_new;
                        pool_s[c * 4 + 1] = l_new;
                        pool_s[c * 4 + 2] = acc_new;
                        pool_s[c * 4 + 3] = ms_new;
                        // lifetime-weighted click score (util.py:23-49) once the last pass is in; the rare P > H case
                        // (user-node rows in the GraphSAGE mean) is finished after the pass loop
                        const int P = cP[c];
                        if (pass == npass - 1 && P <= H)
                            args.scores[(long long)pair0 + c] = (ms_new / (float)P + cscal[c * 4 + 3] + acc_new / l_new) * cw[c];
                    }
                }
            } else if (pass == npass - 1) {
                // ---------------- issuer warp: front end of the NEXT unit, in the shadow of this epilogue --------
                front_cand(args, base + OFF_UB + (ubi ^ 1) * kUnitBuf, lane);
            }
            ++pass_iter;
            __syncthreads();   // out_s (aliasing the O operand) is free again; TMEM may be overwritten; next unit buffer ready
            LIME_TICK(7);
        }

        if (tid == 0) {
            if ((flag_s[0] & 6) != 0) args.fallback_list[atomicAdd(args.fallback_count, 1)] = unit;
            if (flag_s[0] & 1) atomicAdd(args.fallback_count + 2, 1);   // statistics: 4-node units
        }

        // ---------------- P > H: user-node rows take part in the GraphSAGE mean (userEncoders.py:121,153) -------
        if (args.prefix_main > H || args.prefix_tail > H) {
            if (warp < kCWarps) {
                for (int c = warp; c < cnt; c += kCWarps) {
                    const int P = cP[c];
                    if (P > H) {
                        int jn = P - H - 1;
                        jn = jn < C.user_nodes ? jn : C.user_nodes - 1;
                        const float *uu = C.un_prefix + (size_t)jn * kD;
                        const float *hr2 = C.hist_rows + (size_t)cnews[c] * LIME_HIST_LD + LIME_HIST_VC;
                        const float *tr2 = C.hist_tab + (size_t)ctab[c] * LIME_HTAB_LD;
                        float un = 0.f;
                        for (int d = lane; d < kD; d += 32) un = fmaf(hr2[d] + tr2[d], uu[d], un);
                        un = warp_sum(un);
                        if (lane == 0) {
                            const float bs = (pool_s[c * 4 + 3] + un) / (float)P + cscal[c * 4 + 3] + pool_s[c * 4 + 2] / pool_s[c * 4 + 1];
                            args.scores[(long long)pair0 + c] = bs * cw[c];
                        }
                    }
                }
            }
            __syncthreads();   // pool_s is re-initialised at the top of the next unit
        }
        LIME_TICK(8);
    }
#ifdef LIME_TC_PHASE_CLOCKS
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < 16; ++i) atomicAdd(&g_phase_clocks[i], prof[i]);
    }
#endif
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem, 256);
}

// out[(tc * T + th) * 12 + head] = log2(e) * ( sum_k tq[tc][k * 10 + head] * topics[th][k] + tq[tc][500 + head] )
// (pre-scaled so that the scoring kernel's softmax is a bare ex2)
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
        for (int hd = 0; hd < LIME_CA_HEADS; ++hd) o[hd] = acc[hd] * kLog2e;
        o[10] = 0.0f;
        o[11] = 0.0f;
    }
}

// src [rows, lds] fp32, `blocks` blocks of 400 columns -> dst [rows, blocks * 800] fp16: per block the 400 hi
// halves followed by the 400 lo halves (scale * x = hi + lo to 2^-22); absmax[row * ldo] = max |x| of the row
__global__ void split_f16_pairs_kernel(const float *__restrict__ src, int64_t lds, int blocks, float scale,
                                       __half *__restrict__ dst, float *__restrict__ absmax, int64_t ldo) {
    __shared__ float red[4];
    const int64_t row = blockIdx.x;
    const float *s = src + row * lds;
    __half *d = dst + row * (int64_t)blocks * 2 * kD;
    float mx = 0.0f;
    for (int e = threadIdx.x; e < blocks * kD; e += blockDim.x) {
        const int k = e / kD, dd = e - k * kD;
        const float x = s[e];
        const float xs = x * scale;
        const __half h = __float2half_rn(xs);
        const __half l = __float2half_rn(xs - __half2float(h));
        d[k * 2 * kD + dd] = h;
        d[k * 2 * kD + kD + dd] = l;
        mx = fmaxf(mx, fabsf(x));
        if (x != x) mx = INFINITY;
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0 && absmax != nullptr) absmax[row * ldo] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}

}  // namespace

int launch_score_tc(const ScoreArgs &a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        LIME_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    LIME_CUDA(cudaMemsetAsync(a.work_counter, 0, 4 * sizeof(int32_t), st));   // work counter, fallback count, exact counter, stats
    int grid = 2 * num_sms();
    if (const char *e = getenv("LIME_TC_ONE_CTA_PER_SM")) {      // experiment knob: concurrency vs shared resources
        if (e[0] == '1') grid = num_sms();
    }
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

extern "C" int lime_score_phase_clocks(uint64_t *out16) {
    LIME_CHECK_ARG(out16, "lime_score_phase_clocks: null argument");
    unsigned long long zero[16] = {0};
    LIME_CUDA(cudaDeviceSynchronize());
    LIME_CUDA(cudaMemcpyFromSymbol(out16, lime::g_phase_clocks, sizeof(zero)));
    LIME_CUDA(cudaMemcpyToSymbol(lime::g_phase_clocks, zero, sizeof(zero)));
    return 0;
}

extern "C" int lime_split_f16_pairs(const float *src, int64_t lds, int64_t rows, int32_t blocks, float scale, void *dst,
                                    float *absmax, int64_t ldo, void *stream) {
    LIME_CHECK_ARG(src && dst, "lime_split_f16_pairs: null argument");
    LIME_CHECK_ARG(blocks >= 1 && lds >= (int64_t)blocks * LIME_D, "lime_split_f16_pairs: blocks=%d lds=%lld", blocks, (long long)lds);
    if (rows <= 0) return 0;
    lime::split_f16_pairs_kernel<<<(unsigned)rows, 128, 0, lime::as_stream(stream)>>>(src, lds, blocks, scale, reinterpret_cast<__half *>(dst), absmax, ldo);
    LIME_LAUNCH_CHECK("split_f16_pairs_kernel");
    return 0;
}
