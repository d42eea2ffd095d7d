// Stage B on the tensor cores: impression scoring for H <= 56 history rows (sm_100a, tcgen05 + TMEM).  Round 2.
//
// Same arithmetic as score.cu (see its header for the algebra and the reference lines:
// layers.py:52-93, userEncoders.py:121-171, util.py:23-49), restructured so that the 400-wide gate sigmoid is
// evaluated ONCE per unique history row of a work unit instead of once per (candidate, history row):
//
//   o_h(a) = v_h * (1 - f(a)),  f(a) = (1 - a) * sigmoid(a * W_g v_h + b_g)     depends on the candidate only through
//                                                                                 the scalar attention weight a = a[c][h].
//   Over the candidates of a work unit a[.][h] spans [mid_h - w_h, mid_h + w_h].  o_h is analytic in a, so with
//   t = (a[c][h] - mid_h) / w_h in [-1, 1]
//       o_h(a) = c0_h + t * c1_h + R,      c0 = v (1 - f - (w^2/4) f''),   c1 = -w v f'      (all at a = mid_h)
//   i.e. the second-order Taylor polynomial with t^2 replaced by its best constant 1/2 on [-1, 1] (Chebyshev
//   economisation): |R| <= |v| (w^2 max|f''| / 4 + w^3 max|f'''| / 6), the error of 2-node Chebyshev interpolation,
//   from ONE sigmoid per (row, dim).  |f''| <= 0.0962 g^2 + 0.5 |g|, |f'''| <= 0.125 |g|^3 + 0.2887 g^2,
//   g = max_d |W_g v_h|.  A unit whose bound exceeds the tolerance (default 1e-6, below the 2^-22 of ex2.approx) is
//   appended to a device-side list and re-scored by the exact kernel.
//   Everything a pair needs is linear in o (dots with the candidate's 3 folded vectors, sum o) or quadratic
//   (sum o^2 = sum c0^2 + 2 t sum c0 c1 + t^2 sum c1^2), so per unit the dots are ONE GEMM
//       W [3 C x 400] . [c0 ; c1]^T        -> tcgen05.mma, fp32 accumulation in TMEM.
//
// Precision (fp32 fidelity on fp16 tensor cores): W = Whi + Wlo and c0 = Ohi + Olo are fp16 pairs (11 + 11 bits);
// the derivative operand c1 is |t w f'/(1 - f)| ~ 1e-3 of c0, so ONE fp16 product carries it (its rounding error,
// 2^-12 relative to c1, is 2^-22 relative to the result).  Per K step of 16:
//       D[:, 0 .. 3Up)    += Whi . [Ohi ; Olo ; O1hi]^T      (B operand = the three images stacked along N)
//       D[:, 3Up .. 4Up)  += Wlo . [Ohi]^T
//   = 2 MMAs instead of the 6 (3 split products x news / table halves of K) of round 1; the shared-memory operand
//   bytes per K step drop from 36 KB to <= 12 KB, which was the round-1 bottleneck (tensor pipe operand reads).
//
// The bucket-pair (freshness, lifetime) part of a candidate vector, w(c) = w_news(c) + w_tab(bp_c), is no longer
// a second K range: the DISTINCT bucket pairs of a unit (typically 4-6) are extra M rows of the same MMA
// (rows = 3 x (candidates + distinct pairs) <= 120) and the epilogue adds row bp_c to row c through shared memory.
//
// What else changed against round 1:
//   * the pooling softmax / GraphSAGE mean run inside the epilogue (online softmax per TMEM lane triple, shuffles
//     between the 3 lanes of a candidate): the lg / y / z round trip through shared memory and its phase are gone;
//   * operand production is a dynamic task list (8 rows x 32 dims per warp task, round-robin over 7 warps), so short
//     (deduplicated) histories keep every warp busy; the candidate-operand copies (cp.async, 64 contiguous bytes
//     per row per instruction) are tasks of the same list, two stages ahead, over a ring of four 32-dim slots;
//   * history rows are read as fp32 [v | W_g v] straight from the cache (no per-node re-evaluation).
//
// CTA = 7 compute warps + 1 MMA-issuer / front-end warp, persistent, TWO per SM (<= 113 KB of shared memory and
// 256 TMEM columns each).  Per unit: front end (metadata, dedup of history slots and of candidate bucket pairs;
// issuer warp, one unit ahead) -> attention a[u][c] -> centres mid / w, t = (a - mid) / w -> 13 K stages of
// 32 dims -> epilogue + pooling -> score.
#include "score_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <mutex>

#include "tc05.cuh"

namespace lime {
namespace {

constexpr int kD = LIME_D;
constexpr int kCWarps = 7;                      // compute warps (8 warps per CTA -> 4 per SM sub-partition at two CTAs per SM: 128 registers)
constexpr int kCompute = kCWarps * 32;
constexpr int kMmaWarp = kCWarps;               // warp 7 issues the MMAs and runs the front end of the next unit
constexpr int kThreads = kCompute + 32;
constexpr int kH = LIME_TC_MAX_HISTORY;         // 56 history slots at most
constexpr int kTile = LIME_TC_TILE_C;           // candidates per unit handed out by the host
constexpr int kTriples = 40;                    // M rows = 3 x (candidates + distinct bucket pairs): 10 triples per TMEM quadrant
constexpr int kMaxBp = 20;                      // distinct bucket pairs of a unit that fit beside its candidates (50 lifetime buckets: about one pair per candidate)
constexpr int kAS = kTriples;                   // row stride of a_s / t_s  ([u][c])
constexpr int kStages = 13;                     // 32-wide candidate-operand stages over D = 400 (the last holds 16 dims)
constexpr int kOStages = 7;                     // 64-wide O stages
#ifndef LIME_TC_PF_STAGES
#define LIME_TC_PF_STAGES 1
#endif
constexpr int kPfStages = LIME_TC_PF_STAGES;    // L2 prefetch distance of the history rows, in O stages (0 = none)
constexpr int kWTile = 32768, kWImg = 16384;    // one 64-dim candidate tile = hi image + lo image of 128 rows x 128 B
constexpr int kGroups = 7;                      // row groups of 8 unique history rows (56 / 8)
constexpr int kTabLd = LIME_TOPIC_TAB_LD;
constexpr int kTabStride = 64;                  // tab_s[u][3 * bp + k] (float2), 3 * kMaxBp <= 64
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// Both operands are scaled by a power of two before the fp16 hi / lo split, so that the lo halves of typical values
// (|w| ~ 1e-2, |o| ~ 1e-1) stay in the normal fp16 range (>= 6.1e-5); the accumulators are un-scaled in the epilogue.
constexpr float kWScale = LIME_CAND16_SCALE;    // cand16 / ctab16 hold w * 1024
constexpr float kOScale = 256.0f;               // O operand holds o * 256
constexpr float kUnscale = 1.0f / (kWScale * kOScale);
constexpr float kS1Max = 1073741824.0f;         // sum (256 o)^2 <= 2^30  =>  every |256 o| <= 32768 (fp16 operand range)
constexpr float kWAbsMax = 32768.0f / kWScale;  // |w| beyond this leaves the fp16 operand range -> exact kernel
constexpr int kC16 = LIME_CAND16_LD;            // fp16 elements per cand16 / ctab16 row: [hi | lo][k][400]

// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_W = 0;                                      // 2 tiles x (hi, lo) x 128 rows x 128 B: four 32-dim ring slots
constexpr int OFF_O = 2 * kWTile;                             // [Ohi ; Olo ; O1hi] x Up rows x 128 B (two 32-dim half slots)
constexpr int kOBytes = 3 * 64 * 128;
constexpr int OFF_TAB = OFF_W;                                // alias (after the MMAs; the candidate tiles are refilled by the next unit only): table-row results [u][64] float2
constexpr int OFF_T = OFF_O + kOBytes;                        // a[u][c], then t[u][c]
constexpr int OFF_SUM = OFF_T + kH * kAS * 4;                 // [u][8]: sum c0, c1, c0^2, c0 c1, c1^2
constexpr int OFF_MID = OFF_SUM + kH * 8 * 4;
constexpr int OFF_WINV = OFF_MID + kH * 4;
constexpr int OFF_WHALF = OFF_WINV + kH * 4;
constexpr int OFF_PART = OFF_WHALF + kH * 4;                  // [2][40][4] partial pooling state of the two warps of a quadrant
constexpr int OFF_POOL = OFF_PART;                            // alias: [40][4]  merged m, l, acc, ms of the unit (P > H path)
// Per-unit arrays written by the front end: DOUBLE BUFFERED, the issuer warp prepares unit i + 1 while the compute
// warps work on unit i
constexpr int kCArr = kTriples * 4, kUArr = kH * 4;
constexpr int UB_CSCAL = 0;                                   // [40][4]  B1 B2 B3 cb
constexpr int UB_CW = UB_CSCAL + kTriples * 16;
constexpr int UB_CNEWS = UB_CW + kCArr;
constexpr int UB_CTAB = UB_CNEWS + kCArr;                     // bucket-pair id of the candidate
constexpr int UB_CBIDX = UB_CTAB + kCArr;                     // ... and its index among the unit's distinct pairs
constexpr int UB_CP = UB_CBIDX + kCArr;
constexpr int UB_CTOPIC = UB_CP + kCArr;
constexpr int UB_BTAB = UB_CTOPIC + kCArr;                    // [24] distinct bucket pairs
constexpr int UB_UNEWS = UB_BTAB + 96;
constexpr int UB_UTAB = UB_UNEWS + kUArr;
constexpr int UB_UMASK = UB_UTAB + kUArr;
constexpr int UB_UTOPIC = UB_UMASK + kUArr;
constexpr int UB_UMULT = UB_UTOPIC + kUArr;                   // float multiplicity
constexpr int UB_UMP0 = UB_UMULT + kUArr;                     // float multiplicity inside the GraphSAGE prefix (main)
constexpr int UB_UMP1 = UB_UMP0 + kUArr;                      // ... (tail batch)
constexpr int UB_UGABS = UB_UMP1 + kUArr;
constexpr int UB_USLOT = UB_UGABS + kUArr;                    // first history slot of the unique row (long-history mode: index into the attention matrix)
constexpr int UB_WROW = UB_USLOT + kUArr;                     // [40] uint2: per operand-row triple, the row of the cand16 tensor map (hi rows; lo rows = + 3) and the destination byte offset inside an image
constexpr int UB_INFO = UB_WROW + kTriples * 8;           // ints: unit, impression, first pair, count, U, unmasked slots, flags, bucket pairs
constexpr int kUnitBuf = UB_INFO + 32;
constexpr int OFF_UB = OFF_PART + 2 * kTriples * 16;
// front-end scratch (one warp): keys, topic ids, gate bounds of the H history slots
constexpr int OFF_HKN = OFF_UB + 2 * kUnitBuf, OFF_HKT = OFF_HKN + kUArr, OFF_HTP = OFF_HKT + kUArr, OFF_HGA = OFF_HTP + kUArr;
constexpr int OFF_BARS = OFF_HGA + kUArr;                     // wfull[2] ... ofull ofree accum
constexpr int OFF_MISC = OFF_BARS + 128;                      // tmem slot
constexpr int OFF_PROF = OFF_MISC + 64;                       // phase clocks of thread 0 (diagnostic)
constexpr int OFF_BIAS = OFF_PROF + 128;                      // gate bias [400]: read by every lane at every stage
constexpr int kSmemBytes = OFF_BIAS + kD * 4;
static_assert(kH * kTabStride * 8 <= 2 * kWTile, "table-row alias overflows the candidate tiles");
static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");
static_assert(OFF_BIAS % 16 == 0, "alignment");
static_assert(kTile + 1 <= kTriples && 3 * kMaxBp <= kTabStride, "unit capacity");
static_assert(OFF_BARS % 8 == 0 && OFF_T % 16 == 0 && OFF_SUM % 16 == 0 && OFF_UB % 16 == 0 && kUnitBuf % 16 == 0 &&
              OFF_O % 1024 == 0 && UB_WROW % 8 == 0 && OFF_PART % 16 == 0, "alignment");
static_assert(kH <= 8 * kGroups && kH % 8 == 0 && kH <= 64, "row groups");

enum { M_TMEM = 0 };
// unit info ints
enum { UI_UNIT = 0, UI_IMP = 1, UI_PAIR0 = 2, UI_CNT = 3, UI_U = 4, UI_NUN = 5, UI_FLAGS = 6, UI_NBP = 7 };
// barrier indices (uint64 each)
enum { B_WFULL = 0, B_OFULL = 8, B_OFREE = 10, B_ACCUM = 12 };

// Phase timing (diagnostic): thread 0 of every CTA accumulates clock64() deltas per phase; lime_score_phase_clocks reads
// and clears the totals.  Slots: 2 attention, 3 centres, 4 operand production (incl. ring waits), 5 wait for the last
// MMA, 6 epilogue + pooling, 7 merge + score, 8 tail, 9 units.
// Compiled in only with -DLIME_TC_PHASE_CLOCKS (make PHASE_CLOCKS=1).
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

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// one instruction brings a whole cache row into L2 (the front end runs a unit ahead of the warps that read the row)
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
#ifndef LIME_TC_NO_ROW_PREFETCH
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory"); }
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

// M row (= TMEM lane) of operand row k of triple j: 10 triples per quadrant (lanes 30, 31 of a quadrant stay unused),
// so the 3 rows of a candidate live in one warp of the epilogue
__device__ __forceinline__ int m_row(int j, int k) { return 32 * (j / 10) + 3 * (j % 10) + k; }

// fills of candidate tile t per unit: O stages 0 2 4 6 use tile 0, stages 1 3 5 tile 1
__device__ __forceinline__ uint32_t w_uses(int t) { return t == 0 ? 4u : 3u; }

// pack 2 fp32 -> fp16x2 (round to nearest), returns also the rounded values as fp32
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// ---- candidate operand (M side): TMA ---------------------------------------------------------------------------------
// cand16 is a 2-D tensor [6 x rows][400] of fp16 (row pitch 800 B): per cache row the hi images of w1 w2 w3, then the lo
// images.  One box = 64 dims x 3 rows = one operand-row triple of one image, written by the TMA unit straight into the
// K-major SWIZZLE_128B tile (the swizzle follows the shared-memory address bits, so a box may start at any 128-byte row
// of the tile); the last O stage reads dims 384..447, of which 400.. are out of bounds and arrive as zeros.  The 2 x (candidates
// + bucket pairs) <= 80 boxes of a tile are issued by the compute warps, at most one per lane (12 lanes per warp); the tile's
// mbarrier completes on the byte count (no thread waits for the data, no register or generic-proxy store is involved).
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap *map, int col, int row, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(row), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// compute warp `warp`: its share (every kCWarps-th box, one box per lane) of the candidate tile of O stage kb -> tile kb & 1.
// Warp 0 also posts the byte count (a complete_tx that overtakes it only drives the transaction count negative for a while).
__device__ __forceinline__ void issue_w_tile(unsigned char *base, uint64_t *bars, const CUtensorMap *map, const uint2 *wrow, int nt,
                                             int kb, int warp, int lane) {
    uint64_t *bar = bars + B_WFULL + (kb & 1);
#ifdef LIME_TC_DIAG_NOCOPY       // timing diagnostic only (wrong results): no candidate-operand copies
    if (warp == 0 && lane == 0) tc::mbar_arrive(bar);
    return;
#endif
    if (warp == 0 && lane == 0) mbar_expect_tx(bar, (uint32_t)nt * 2u * 3u * 128u);
    const uint32_t wt = tc::smem_u32(base) + OFF_W + (uint32_t)(kb & 1) * kWTile;
    const int b = warp + kCWarps * lane;
    if (b < 2 * nt) {
        const int lo = b >= nt ? 1 : 0;
        const uint2 rw = wrow[b - (lo ? nt : 0)];
        tma_load_box(wt + (lo ? kWImg : 0) + rw.y, map, 64 * kb, (int)rw.x + 3 * lo, bar);
    }
}
// Ring protocol.  The O operand is ONE 64-dim tile, produced and consumed once per O stage (7 per unit); the candidate
// operand keeps four 32-dim slots (two tiles), two per O stage.  EVERY compute warp waits for and arrives on every O
// stage, whether or not it holds rows -- an mbarrier parity wait is only meaningful for a waiter that has observed
// every earlier phase, and a phase cannot run ahead of a warp whose arrival it needs.  The completion of O stage kb's
// MMAs (O_FREE) also frees its two candidate slots: there is no separate barrier for them.
// (Measured: a stage costs one memory + barrier round trip of 2-3 thousand clocks almost independently of its size,
// so 7 fat stages beat 13 thin ones.)
__device__ __forceinline__ void o_free_wait(uint64_t *bars, int kb, uint32_t pass_iter) {
    const uint32_t fill = pass_iter * (uint32_t)kOStages + (uint32_t)kb;
    if (fill >= 1) tc::mbar_wait(bars + B_OFREE, (fill - 1) & 1u, 200 + kb);
}

// ---- history operand (N side) -------------------------------------------------------------------------------------
// Fixed ownership: lane = (row slot j = lane >> 3, quad q = lane & 7 of the stage's 32 dims); "virtual warp" vw = warp
// (+ 7 in the second pass, histories with more than 26 unique rows) owns the rows 8 (vw >> 1) + P[4 (vw & 1) + j],
// P = [0 4 1 5 2 6 3 7]: the two rows of a half warp differ in bit 2 of the row number, which the 128-byte swizzle turns
// into opposite 64-byte halves of the bank space -- the 8-byte stores of a warp are conflict-free.  Everything that
// depends on the row only (cache pointers, expansion point, half width) is hoisted out of the stage loop, the five row
// sums stay in registers for the whole unit, and the global loads run one stage ahead of the arithmetic.
struct Quad {                // 4 dims of a cache row: per-news part (DRAM) and bucket-pair part (L2); the gate bias is
    float4 vn, gn, vt, gt;   // read from shared memory when the quad is evaluated
};
// v and g of 4 dims with ONE 256-bit load (hist_vg / htab_vg interleave them): two global loads per quad instead of four,
// so the loads of the next stage and of the current one never share a scoreboard slot
__device__ __forceinline__ void ldg8(const float *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
struct RowCtx {
    const float *hrow, *trow, *bias;
    float a, oml, nwv, nso, kq;
    bool ok;
};
__device__ __forceinline__ void quad_load(Quad &q, const RowCtx &c, int d) {
#ifdef LIME_TC_DIAG_NOLOAD       // timing diagnostic only (wrong results): no global loads in the production loop
    q.vn = q.gn = q.vt = q.gt = make_float4(0.01f * d, 0.02f, 0.03f, 0.04f);
    return;
#endif
    if (c.ok && d < kD) {
#if defined(LIME_TC_DIAG_SKIP) && LIME_TC_DIAG_SKIP == 1      // timing diagnostics only (wrong results)
        q.vn = q.gn = make_float4(0.01f, 0.02f, 0.03f, 0.04f);
#else
        ldg8(c.hrow + 2 * d, q.vn, q.gn);
#endif
#if defined(LIME_TC_DIAG_SKIP) && LIME_TC_DIAG_SKIP == 2
        q.vt = q.gt = make_float4(0.01f, 0.02f, 0.03f, 0.04f);
#else
        ldg8(c.trow + 2 * d, q.vt, q.gt);
#endif
    }
    // the same quad kPfStages stages on: the news row slice comes from DRAM, and a whole unit of look-ahead for every CTA
    // does not fit the L2 (measured: the unit-ahead row prefetch of the front end changes nothing)
    if (kPfStages > 0 && c.ok && d + 64 * kPfStages < kD) prefetch_l2(c.hrow + 2 * (d + 64 * kPfStages));
}
__device__ __forceinline__ void row_ctx_init(RowCtx &rc, const LimeNewsCache &C, const float *bias_s, const float *htab, int u,
                                             int U, const int *unews, const int *utab, const float *mid_s, const float *whalf_s) {
    rc.ok = u < U;
    rc.bias = bias_s;
    rc.hrow = rc.trow = nullptr;
    rc.a = rc.oml = rc.nwv = rc.nso = rc.kq = 0.0f;
    if (rc.ok) {
        rc.hrow = C.hist_vg + (size_t)unews[u] * (2 * kD);
        rc.trow = htab + (size_t)utab[u] * (2 * kD);
        rc.a = mid_s[u];
        const float w = whalf_s[u];
        const float om = 1.0f - rc.a;
        rc.oml = -om * kLn2;                              // d/da of the gate argument in natural units: g = -ln2 * g'
        rc.nwv = -kOScale * w;                            // c1 = -256 w v f'
        rc.nso = -kOScale * om;                           // 256 (1 - om s)
        rc.kq = kOScale * 0.25f * w * w * kLn2;           // + 256 (w^2/4) ln2 r' n  ( = -256 (w^2/4) f'' )
    }
}
// 4 elements: s = sigmoid(x) = 1 / (1 + 2^z'),  z' = a g' + b'  (g', b' pre-scaled by -log2 e);  with g = -ln2 g':
//   f = om s,  f' = -s + om g s (1 - s),  f'' = g s (1 - s) (om g (1 - 2 s) - 2)
//   c0 = 256 v (1 - f - (w^2/4) f''),  c1 = -256 w v f'
__device__ __forceinline__ void quad_eval(const Quad &q, const RowCtx &c, int d, float (&ps)[5], uint32_t &hi0, uint32_t &hi1,
                                          uint32_t &lo0, uint32_t &lo1, uint32_t &d0, uint32_t &d1) {
    const float vv[4] = {q.vn.x + q.vt.x, q.vn.y + q.vt.y, q.vn.z + q.vt.z, q.vn.w + q.vt.w};
    const float gg[4] = {q.gn.x + q.gt.x, q.gn.y + q.gt.y, q.gn.z + q.gt.z, q.gn.w + q.gt.w};
    const float4 bb = *reinterpret_cast<const float4 *>(c.bias + d);      // shared memory
    const float bs[4] = {bb.x, bb.y, bb.z, bb.w};
    float c0[4], c1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float s = rcp_approx(ex2_approx(fmaf(c.a, gg[e], bs[e])) + 1.0f);
        const float rp = fmaf(-s, s, s) * gg[e];               // s (1 - s) g'
        const float f1 = fmaf(c.oml, rp, -s);                  // f'
        const float n2 = fmaf(c.oml * gg[e], fmaf(-2.0f, s, 1.0f), -2.0f);
        const float tm = fmaf(c.kq, rp * n2, fmaf(c.nso, s, kOScale));
        c0[e] = vv[e] * tm;
        c1[e] = (vv[e] * c.nwv) * f1;
        ps[0] += c0[e];
        ps[1] += c1[e];
        ps[2] = fmaf(c0[e], c0[e], ps[2]);
        ps[3] = fmaf(c0[e], c1[e], ps[3]);
        ps[4] = fmaf(c1[e], c1[e], ps[4]);
    }
    split2(c0[0], c0[1], hi0, lo0);
    split2(c0[2], c0[3], hi1, lo1);
    d0 = pack_h2(c1[0], c1[1]);
    d1 = pack_h2(c1[2], c1[3]);
}

// Candidate-aware attention weights a[u][c] (layers.py:66-81) of one work unit.  LPC = 2^lpc_log2 lanes share a candidate,
// a lane owns the unique rows u = l + LPC * i (i < 4): all 12 table loads of a lane are issued before the first use.  The
// topic-pair table holds the EXPONENTIALS of the head logits (no max pass: the host checks the table's |logit| bound), so
// the first softmax costs no MUFU at all; masked slots (mask == 0 -> -1e9, layers.py:72) contribute exactly 0 unless every
// slot is masked, in which case both softmaxes are uniform over the H slots.  One copy of the code for every LPC (the
// cross-lane sums are run-time loops): the kernel has to stay inside the instruction cache.
__device__ __forceinline__ void attention(const float *__restrict__ table, int T, int U, int cnt, int nun, int warp, int lane,
                                       int lpc_log2, const int *ctopic, const int *utopic, const int *umask, const float *umult,
                                       float *a_s) {
    const int LPC = 1 << lpc_log2, CPW = 32 >> lpc_log2;      // candidates per warp and round
    const int l = lane & (LPC - 1), g = lane >> lpc_log2;
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
            const float *trow = table + (size_t)ctopic[cc] * T * kTabLd;
            float4 x[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float *r0 = trow + (size_t)tp[i] * kTabLd;
                x[i][0] = ldg4(r0);
                x[i][1] = ldg4(r0 + 4);
                x[i][2] = ldg4(r0 + 8);
            }
            float sum[LIME_CA_HEADS];
#pragma unroll
            for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xs[LIME_CA_HEADS] = {x[i][0].x, x[i][0].y, x[i][0].z, x[i][0].w, x[i][1].x,
                                                 x[i][1].y, x[i][1].z, x[i][1].w, x[i][2].x, x[i][2].y};
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = fmaf(w[i], xs[hd], sum[hd]);
            }
#pragma unroll 1
            for (int o = 1; o < LPC; o <<= 1) {
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] += __shfl_xor_sync(0xffffffffu, sum[hd], o);
            }
#pragma unroll
            for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = __fdividef(1.0f, sum[hd]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xs[LIME_CA_HEADS] = {x[i][0].x, x[i][0].y, x[i][0].z, x[i][0].w, x[i][1].x,
                                                 x[i][1].y, x[i][1].z, x[i][1].w, x[i][2].x, x[i][2].y};
                float agg = 0.0f;
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) agg = fmaf(xs[hd], sum[hd], agg);
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
#pragma unroll 1
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
// front_history: the history slots (news, bucket pair, mask, topic id, gate bound), two slots per lane, deduplicated into
// unique operand rows.  front_cand: the candidates (cache rows, lifetime weights, folded scalars), the distinct bucket
// pairs among them and the operand-triple table of the TMA copies.  All results land in the unit buffer `ub`.
//
// Deduplication: slots with equal (news, bucket pair, mask) -- in practice the zero padding of a short history,
// dataset.py:123-128 -- are ONE operand row whose multiplicity enters the two attention softmaxes, the pooling softmax
// and the GraphSAGE prefix means.  Equal keys are found with match.any inside each half (slots 0..31 / 32..63); across the
// halves only the group of slot 31 is merged (a padding run that starts in the first half covers the whole second half;
// any other duplicate pair split by the halves stays two rows, which is merely not deduplicated).
// one copy of the bit-exact (accurate logf) bucketisation in the kernel image
__device__ __noinline__ int bucket_pair(float fresh, float life, float scale, int nb) {
    return bucketize_seconds(fresh, scale, nb) * nb + bucketize_seconds(life, scale, nb);
}
__device__ __forceinline__ uint32_t prefix_bits(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : (1u << n) - 1u); }
__device__ __forceinline__ void front_history(const ScoreArgs &args, unsigned char *ub, int unit, int lane) {
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
    // raw slot arrays of both slots first, then the dependent news_meta sectors (topic id | max |W_g vc|)
    const bool va = lane < H, vb = lane + 32 < H;
    const long long oa = (long long)imp * H + lane, ob = oa + 32;
    int na = va ? I.hist_news[oa] : 0, nbn = vb ? I.hist_news[ob] : 0;
    const int mka = va && I.hist_mask[oa] != 0 ? 1 : 0, mkb = vb && I.hist_mask[ob] != 0 ? 1 : 0;
    const float fra = va ? I.hist_fresh[oa] : 1.0f, frb = vb ? I.hist_fresh[ob] : 1.0f;
    const float lfa = va ? I.hist_life[oa] : 1.0f, lfb = vb ? I.hist_life[ob] : 1.0f;
    na = (na < 0 || na >= C.news_num) ? 0 : na;
    nbn = (nbn < 0 || nbn >= C.news_num) ? 0 : nbn;
    const float2 mta = __ldg(reinterpret_cast<const float2 *>(C.news_meta + (size_t)na * LIME_META_LD));
    const float2 mtb = __ldg(reinterpret_cast<const float2 *>(C.news_meta + (size_t)nbn * LIME_META_LD));
    // the candidate arrays of the unit: L2 by the time front_cand reads them
    for (int j = lane; 32 * j < cnt; j += 32) {
        prefetch_l2(I.cand_news + pair0 + 32 * j);
        prefetch_l2(I.cand_fresh + pair0 + 32 * j);
        prefetch_l2(I.cand_life + pair0 + 32 * j);
        if (I.cand_remaining) prefetch_l2(I.cand_remaining + pair0 + 32 * j);
    }
    const int bpa = bucket_pair(fra, lfa, args.bucket_scale, nb);
    const int bpb = bucket_pair(frb, lfb, args.bucket_scale, nb);
    // keys: (news, 2 * bucket pair + mask); slots beyond H get keys that match nothing
    const unsigned long long ka = va ? ((unsigned long long)(unsigned)na << 32) | (unsigned)(2 * bpa + mka)
                                     : 0xffffffff00000000ull | (unsigned)lane;
    const unsigned long long kb = vb ? ((unsigned long long)(unsigned)nbn << 32) | (unsigned)(2 * bpb + mkb)
                                     : 0xfffffffe00000000ull | (unsigned)lane;
    const unsigned maa = __match_any_sync(0xffffffffu, ka), mbb = __match_any_sync(0xffffffffu, kb);
    const unsigned long long k31 = __shfl_sync(0xffffffffu, ka, 31);
    const bool dupb = vb && kb == k31;                       // second-half slot that belongs to the group of slot 31
    const unsigned bdup = __ballot_sync(0xffffffffu, dupb);
    const bool in31 = va && ka == k31;
    const bool isfa = va && (__ffs(maa) - 1 == lane);
    const bool isfb = vb && !dupb && (__ffs(mbb) - 1 == lane);
    const unsigned b0 = __ballot_sync(0xffffffffu, isfa), b1 = __ballot_sync(0xffffffffu, isfb);
    const unsigned lt = (1u << lane) - 1u;
    // GraphSAGE prefix inside this history (chunk h0.. of a long history: the part of the prefix that falls into the chunk)
    const int h0 = args.a_matrix != nullptr ? (imp / args.chunk_impressions) * H : 0;
    const int pz0 = min(max(args.prefix_main - h0, 0), H), pz1 = min(max(args.prefix_tail - h0, 0), H);
    const unsigned pa0 = prefix_bits(pz0), pb0 = prefix_bits(pz0 - 32), pa1 = prefix_bits(pz1), pb1 = prefix_bits(pz1 - 32);
    int *unews = reinterpret_cast<int *>(ub + UB_UNEWS);
    int *utab = reinterpret_cast<int *>(ub + UB_UTAB);
    int *umask = reinterpret_cast<int *>(ub + UB_UMASK);
    int *utopic = reinterpret_cast<int *>(ub + UB_UTOPIC);
    float *umult = reinterpret_cast<float *>(ub + UB_UMULT);
    float *ump0 = reinterpret_cast<float *>(ub + UB_UMP0);
    float *ump1 = reinterpret_cast<float *>(ub + UB_UMP1);
    float *ugabs = reinterpret_cast<float *>(ub + UB_UGABS);
    int *uslot = reinterpret_cast<int *>(ub + UB_USLOT);
    if (isfa) {
        const int u = __popc(b0 & lt);
        uslot[u] = lane;
        const int tp = __float_as_int(mta.x);
        unews[u] = na;
        utab[u] = bpa;
        umask[u] = mka;
        utopic[u] = (tp < 0 || tp >= T) ? 0 : tp;
        ugabs[u] = mta.y;
        umult[u] = (float)(__popc(maa) + (in31 ? __popc(bdup) : 0));
        ump0[u] = (float)(__popc(maa & pa0) + (in31 ? __popc(bdup & pb0) : 0));
        ump1[u] = (float)(__popc(maa & pa1) + (in31 ? __popc(bdup & pb1) : 0));
    }
    if (isfb) {
        const int u = __popc(b0) + __popc(b1 & lt);
        uslot[u] = lane + 32;
        const int tp = __float_as_int(mtb.x);
        unews[u] = nbn;
        utab[u] = bpb;
        umask[u] = mkb;
        utopic[u] = (tp < 0 || tp >= T) ? 0 : tp;
        ugabs[u] = mtb.y;
        umult[u] = (float)__popc(mbb);
        ump0[u] = (float)__popc(mbb & pb0);
        ump1[u] = (float)__popc(mbb & pb1);
    }
    const int nun = __reduce_add_sync(0xffffffffu, mka + mkb);      // unmasked history slots
    if (lane == 0) {
        info[UI_UNIT] = unit;
        info[UI_IMP] = imp;
        info[UI_PAIR0] = pair0;
        info[UI_CNT] = cnt;
        info[UI_U] = __popc(b0) + __popc(b1);
        info[UI_NUN] = nun;
    }
    __syncwarp();
}

__device__ __forceinline__ void front_cand(const ScoreArgs &args, unsigned char *ub, int lane) {
    const LimeNewsCache &C = args.cache;
    const uint32_t tab_row0 = (blockIdx.x % (unsigned)(C.tab_replicas > 0 ? C.tab_replicas : 1)) * (uint32_t)(C.num_buckets * C.num_buckets);
    const LimeImpressions &I = args.imp;
    const int nb = C.num_buckets, T = C.num_topics;
    int *info = reinterpret_cast<int *>(ub + UB_INFO);
    if (info[UI_UNIT] >= I.num_units) return;
    float *cscal = reinterpret_cast<float *>(ub + UB_CSCAL);
    float *cw = reinterpret_cast<float *>(ub + UB_CW);
    int *cnews = reinterpret_cast<int *>(ub + UB_CNEWS);
    int *ctab = reinterpret_cast<int *>(ub + UB_CTAB);
    int *cbidx = reinterpret_cast<int *>(ub + UB_CBIDX);
    int *cP = reinterpret_cast<int *>(ub + UB_CP);
    int *ctopic = reinterpret_cast<int *>(ub + UB_CTOPIC);
    int *btab = reinterpret_cast<int *>(ub + UB_BTAB);
    uint2 *wrow = reinterpret_cast<uint2 *>(ub + UB_WROW);
    const int pair0 = info[UI_PAIR0];
    int cnt = info[UI_CNT];
    int flags = 0;
    if (cnt > kTriples - 1) {          // the host built the units for a larger tile: exact kernel
        cnt = kTriples - 1;
        flags = 4;
    }
    // two candidates per lane (c = lane, lane + 32): raw arrays first, then the dependent cache sectors
    int key[2], nn[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int c = lane + 32 * r;
        key[r] = -1 - c;
        nn[r] = 0;
        if (c < cnt) {
            const long long p = (long long)pair0 + c;
            int n = I.cand_news[p];
            n = (n < 0 || n >= C.news_num) ? 0 : n;
            const float fr = I.cand_fresh[p], lf = I.cand_life[p];
            const float rem = I.cand_remaining ? I.cand_remaining[p] : __fsub_rn(lf, fr);
            const float4 m0 = ldg4(C.news_meta + (size_t)n * LIME_META_LD), m1 = ldg4(C.news_meta + (size_t)n * LIME_META_LD + 4);
            prefetch_l2_bulk(reinterpret_cast<const unsigned char *>(C.cand16) + (size_t)n * (2 * kC16), 2 * kC16);
            const int tb = bucket_pair(fr, lf, args.bucket_scale, nb);
            const float *ctr = C.cand_tab + (size_t)tb * LIME_CTAB_LD + LIME_CAND_SCAL + 3;
            const float4 ct = make_float4(__ldg(ctr), __ldg(ctr + 1), __ldg(ctr + 2), __ldg(ctr + 3));
            const int tp = __float_as_int(m0.x);
            key[r] = tb;
            nn[r] = n;
            cnews[c] = n;
            ctab[c] = tb;
            cw[c] = lifetime_weight(rem, C);
            const long long p_real = args.a_matrix != nullptr ? p % args.chunk_pairs : p;       // pairs of a chunk: k * chunk_pairs + p
            cP[c] = (args.pair_index_base + p_real >= args.tail_start) ? args.prefix_tail : args.prefix_main;
            ctopic[c] = (tp < 0 || tp >= T) ? 0 : tp;
            *reinterpret_cast<float4 *>(cscal + c * 4) = make_float4(m0.w + ct.x, m1.x + ct.y, m1.y + ct.z, m1.z + ct.w);
            if (!(m0.z <= kWAbsMax)) flags |= 4;       // beyond the fp16 operand range: exact kernel
        }
    }
    // distinct bucket pairs of the unit's candidates: match.any inside each half, the (<= 7) candidates of the second half
    // are looked up in the first half one by one
    const unsigned ma = __match_any_sync(0xffffffffu, key[0]), mb = __match_any_sync(0xffffffffu, key[1]);
    unsigned cross = 0;
    for (int b = 0; 32 + b < cnt; ++b) {
        const int kv = __shfl_sync(0xffffffffu, key[1], b);
        const unsigned eq = __ballot_sync(0xffffffffu, key[0] == kv);
        cross = lane == b ? eq : cross;
    }
    const bool va = lane < cnt, vb = lane + 32 < cnt;
    const bool isfa = va && (__ffs(ma) - 1 == lane);
    const bool isfb = vb && cross == 0u && (__ffs(mb) - 1 == lane);
    const unsigned b0 = __ballot_sync(0xffffffffu, isfa), b1 = __ballot_sync(0xffffffffu, isfb);
    const int nb0 = __popc(b0);
    int nbp = nb0 + __popc(b1);
    if (va) {
        const int ia = __popc(b0 & ((1u << (__ffs(ma) - 1)) - 1u));
        cbidx[lane] = min(ia, kMaxBp - 1);
        if (isfa && ia < kMaxBp) btab[ia] = key[0];
    }
    if (vb) {
        const int ib = cross != 0u ? __popc(b0 & ((1u << (__ffs(cross) - 1)) - 1u)) : nb0 + __popc(b1 & ((1u << (__ffs(mb) - 1)) - 1u));
        cbidx[lane + 32] = min(ib, kMaxBp - 1);
        if (isfb && ib < kMaxBp) btab[ib] = key[1];
    }
    if (nbp > kMaxBp || cnt + nbp > kTriples) {     // more bucket pairs than spare operand rows: exact kernel
        flags |= 4;
        nbp = min(min(nbp, kMaxBp), kTriples - cnt);
    }
    __syncwarp();
    // operand-row triples: row of the cand16 tensor map (6 rows per cache row; the bucket-pair rows follow the news rows)
    // and destination byte offset inside an image (3 consecutive 128-byte rows)
    const int nt = cnt + nbp;
    for (int j = lane; j < nt; j += 32) {
#ifdef LIME_TC_DIAG_WROW0        // timing diagnostic only (wrong results): every candidate copies news 1..8 -> L2 hits, no DRAM
        const uint32_t row = j < cnt ? (uint32_t)(1 + (j & 7)) : (uint32_t)C.news_num + tab_row0 + (uint32_t)btab[j - cnt];
#else
        const uint32_t row = j < cnt ? (uint32_t)cnews[j] : (uint32_t)C.news_num + tab_row0 + (uint32_t)btab[j - cnt];
#endif
        wrow[j] = make_uint2(6u * row, (uint32_t)m_row(j, 0) * 128u);
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) {
        info[UI_FLAGS] = flags;
        info[UI_NBP] = nbp;
        info[UI_CNT] = cnt;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kThreads, 2) score_tc_kernel(const __grid_constant__ ScoreArgs args, const __grid_constant__ CUtensorMap wmap) {
    // Every shared-memory pointer below is derived from this array by pointer arithmetic only (no integer round trip), so
    // the compiler keeps the shared state space and emits LDS / STS instead of generic loads; the operand tiles need
    // the 1024-byte alignment of the 128-byte swizzle, which the declaration requests and the first thread verifies.
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw;
    if (threadIdx.x == 0 && (tc::smem_u32(smem_raw) & 1023u) != 0) __trap();
    float2 *tab_s = reinterpret_cast<float2 *>(base + OFF_TAB);
    float *t_s = reinterpret_cast<float *>(base + OFF_T);
    float *sum_s = reinterpret_cast<float *>(base + OFF_SUM);
    float *mid_s = reinterpret_cast<float *>(base + OFF_MID);
    float *winv_s = reinterpret_cast<float *>(base + OFF_WINV);
    float *whalf_s = reinterpret_cast<float *>(base + OFF_WHALF);
    float *part_s = reinterpret_cast<float *>(base + OFF_PART);
    float *pool_s = reinterpret_cast<float *>(base + OFF_POOL);
    int *hkn = reinterpret_cast<int *>(base + OFF_HKN);
    int *hkt = reinterpret_cast<int *>(base + OFF_HKT);
    int *htp = reinterpret_cast<int *>(base + OFF_HTP);
    float *hga = reinterpret_cast<float *>(base + OFF_HGA);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + OFF_BARS);

    const LimeNewsCache &C = args.cache;
    const LimeImpressions &I = args.imp;
    const int H = I.max_history;
    const int T = C.num_topics;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // this CTA's copy of the bucket-pair tables (history role; the candidate role's copy sits in the tail of cand16)
    const int nb2 = C.num_buckets * C.num_buckets;
    const int rep = (int)(blockIdx.x % (unsigned)(C.tab_replicas > 0 ? C.tab_replicas : 1));
    const float *htab = C.htab_vg + (size_t)rep * nb2 * (2 * kD);
    float *bias_s = reinterpret_cast<float *>(base + OFF_BIAS);
    for (int i = tid; i < kD; i += kThreads) bias_s[i] = C.gate_bias[i];
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) tc::mbar_init(bars + B_WFULL + s, 1);      // one arrive.expect_tx per fill; the TMA unit completes the bytes
        tc::mbar_init(bars + B_OFULL, kCWarps);                 // one arrival per compute warp
        tc::mbar_init(bars + B_OFREE, 1);
        tc::mbar_init(bars + B_ACCUM, 1);
        tc::mbar_fence_init();
    }
    int next_unit = 0;               // issuer warp: the unit whose front end it runs next
    if (warp == kMmaWarp) {
        tc::tmem_alloc(reinterpret_cast<uint32_t *>(base + OFF_MISC) + M_TMEM, 256);
        int u0 = 0;
        if (lane == 0) u0 = atomicAdd(args.work_counter, 1);
        u0 = __shfl_sync(0xffffffffu, u0, 0);
        front_history(args, base + OFF_UB, u0, lane);
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
    uint32_t pass_iter = 0;          // passes (= units) processed so far: every role derives the barrier phases from it
    int ubi = 0;                     // unit buffer of the current unit

    for (;; ubi ^= 1) {
        unsigned char *ub = base + OFF_UB + ubi * kUnitBuf;
        volatile int *info = reinterpret_cast<volatile int *>(ub + UB_INFO);
        const float *cscal = reinterpret_cast<const float *>(ub + UB_CSCAL);
        const float *cw = reinterpret_cast<const float *>(ub + UB_CW);
        const int *cnews = reinterpret_cast<const int *>(ub + UB_CNEWS);
        const int *ctab = reinterpret_cast<const int *>(ub + UB_CTAB);
        const int *cbidx = reinterpret_cast<const int *>(ub + UB_CBIDX);
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
        const uint2 *wrow = reinterpret_cast<const uint2 *>(ub + UB_WROW);
        int *flag_s = reinterpret_cast<int *>(ub + UB_INFO) + UI_FLAGS;

        const int unit = info[UI_UNIT];
        if (unit >= I.num_units) break;
#ifdef LIME_TC_PHASE_CLOCKS
        if (tid == 0) { t_last = clock64(); ++prof[9]; }
#endif
        const int pair0 = info[UI_PAIR0], cnt = info[UI_CNT], U = info[UI_U], nbp = info[UI_NBP];
        const int Up = (U + 15) & ~15;
        const int ngroups = (U + 7) >> 3;

        // ================= roles ======================================================================
        if (warp < kCWarps) {
            // the candidate operand of the first two O stages goes out now (both tiles are free since the previous unit's
            // last MMA): it lands during the attention phase
            issue_w_tile(base, bars, &wmap, wrow, cnt + nbp, 0, warp, lane);
            issue_w_tile(base, bars, &wmap, wrow, cnt + nbp, 1, warp, lane);
            // ---------------- phase 1: candidate-aware attention weights a[u][c] (layers.py:66-81) ------
            // lanes per candidate = the smallest power of two that covers the U rows with 4 rows per lane
            if (args.a_matrix != nullptr) {
                // long-history mode: the weights over the FULL history were computed by the pre-pass; a unique row takes the
                // weight of its first slot (duplicates share topic and mask, hence the weight)
                const int *uslot = reinterpret_cast<const int *>(ub + UB_USLOT);
                for (int idx = tid; idx < cnt * U; idx += kCompute) {
                    const int c = idx / U, u = idx - c * U;
                    t_s[u * kAS + c] = __ldg(args.a_matrix + ((long long)pair0 + c) * H + uslot[u]);
                }
            } else {
                attention(C.topic_table, T, U, cnt, info[UI_NUN], warp, lane, U <= 8 ? 1 : (U <= 16 ? 2 : (U <= 32 ? 3 : 4)), ctopic, utopic,
                          umask, umult, t_s);
            }
            bar_compute();
            LIME_TICK(2);

            // ---------------- centres: expansion point and half width per unique row (4 lanes per row) ----
            {
                const int u = tid >> 2, l4 = tid & 3;
                const int uc = u < U ? u : U - 1;
                float lo = INFINITY, hi = -INFINITY;
                for (int c = l4; c < cnt; c += 4) {
                    const float a = t_s[uc * kAS + c];
                    lo = fminf(lo, a);
                    hi = fmaxf(hi, a);
                }
#pragma unroll
                for (int o = 1; o < 4; o <<= 1) {
                    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                const float wh = fmaxf(0.5f * (hi - lo), 1e-7f);
                const float mid = 0.5f * (hi + lo), winv = 1.0f / wh;
                if (u < U) {
                    // t = (a - mid) / w in [-1, 1], in place
                    for (int c = l4; c < cnt; c += 4) {
                        const float t = (t_s[u * kAS + c] - mid) * winv;
                        t_s[u * kAS + c] = fminf(fmaxf(t, -1.0f), 1.0f);
                    }
                    if (l4 == 0) {
                        mid_s[u] = mid;
                        whalf_s[u] = wh;
                        winv_s[u] = winv;
                        // Remainder of the economised expansion of f(a) = (1 - a) sigmoid(g a + b) on [mid - w, mid + w]:
                        // w^2 max|f''| / 4 + w^3 max|f'''| / 6; g is bounded by the cached max |W_g vc| of the news plus the
                        // max over the bucket-pair table.
                        const float gabs = (ugabs[u] + C.tab_gw_absmax) * (1.0f / kLog2e);
                        const float w2 = wh * wh, g2 = gabs * gabs;
                        const float err = w2 * (0.0962f * g2 + 0.5f * gabs) * 0.25f +
                                          w2 * wh * (0.125f * g2 * gabs + 0.2887f * g2) * (1.0f / 6.0f);
                        if (!(err <= args.interp_tol)) atomicOr(flag_s, 2);      // exact kernel
                    }
                }
            }
            bar_compute();
            LIME_TICK(3);

            // ---------------- operand production + candidate copies -----------------------------------------
            // per O stage: 64 dims of this warp's rows, then its share of the candidate operand of the next O stage
            {
                const int j = lane >> 3, q = lane & 7;
                const int npass = (8 * ((warp + kCWarps) >> 1) + (((warp + kCWarps) & 1) ? 2 : 0)) < U ? 2 : 1;   // warp-uniform
                RowCtx rc[2];
                int urow[2];
                float ps[2][5];
                unsigned char *orow[2];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int vw = warp + kCWarps * p, x = 4 * (vw & 1) + j;
                    urow[p] = 8 * (vw >> 1) + (x >> 1) + 4 * (x & 1);
                    row_ctx_init(rc[p], C, bias_s, htab, p < npass ? urow[p] : U, U, unews, utab, mid_s, whalf_s);
#pragma unroll
                    for (int i = 0; i < 5; ++i) ps[p][i] = 0.0f;
                    orow[p] = base + OFF_O + (rc[p].ok ? (urow[p] >> 3) * 1024 + (urow[p] & 7) * 128 : 0) + 8 * (q & 1);
                }
                // one quad of each 32-dim half of the stage per lane and row: dims 64 kb + 4 q and 64 kb + 32 + 4 q
                auto store_quad = [&](int p, int h, uint2 hi, uint2 lo, uint2 d1) {
                    // Up is a multiple of 8: the three images share the swizzle phase of the row
                    unsigned char *ob = orow[p] + (((uint32_t)(4 * h + (q >> 1)) ^ (uint32_t)(urow[p] & 7)) << 4);
                    *reinterpret_cast<uint2 *>(ob) = hi;
                    *reinterpret_cast<uint2 *>(ob + (size_t)Up * 128) = lo;
                    *reinterpret_cast<uint2 *>(ob + (size_t)Up * 256) = d1;
                };
                // steps (stage, row pass) in order (0,0) [(0,1)] (1,0) ...: the loads of a step are issued one step ahead
                Quad na, nb;
                quad_load(na, rc[0], 4 * q);
                quad_load(nb, rc[0], 32 + 4 * q);
                for (int kb = 0; kb < kOStages; ++kb) {
                    const int dA = 64 * kb + 4 * q, dB = dA + 32;
                    const bool okA = dA < kD, okB = dB < kD;
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        if (p < npass) {
                            const Quad ca = na, cb = nb;
                            if (p + 1 < npass) {                      // next step: the lane's second row, same stage
                                quad_load(na, rc[1], dA);
                                quad_load(nb, rc[1], dB);
                            } else {                                  // next step: the first row, next stage
                                quad_load(na, rc[0], dA + 64);
                                quad_load(nb, rc[0], dB + 64);
                            }
                            uint2 hi, lo, d1;
                            const bool a0 = rc[p].ok && okA, b0 = rc[p].ok && okB;
                            if (a0) quad_eval(ca, rc[p], dA, ps[p], hi.x, hi.y, lo.x, lo.y, d1.x, d1.y);
                            if (p == 0) {
                                LIME_TICK(12);
                                o_free_wait(bars, kb, pass_iter);     // the tile is free once the MMAs of the previous stage have drained it
                                LIME_TICK(13);
                            }
                            if (a0) store_quad(p, 0, hi, lo, d1);
                            if (b0) {
                                quad_eval(cb, rc[p], dB, ps[p], hi.x, hi.y, lo.x, lo.y, d1.x, d1.y);
                                store_quad(p, 1, hi, lo, d1);
                            }
                        }
                    }
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(bars + B_OFULL);
                    LIME_TICK(14);
                    // the candidate tile of the previous O stage is free (its MMAs completed: observed above): the next O
                    // stage's candidate operand goes into it, one production stage ahead of its MMAs
                    if (kb >= 1 && kb + 1 < kOStages) issue_w_tile(base, bars, &wmap, wrow, cnt + nbp, kb + 1, warp, lane);
                    LIME_TICK(11);
                }
                // row sums over the 8 quad lanes of a row (fixed order: bit-reproducible)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (p < npass) {
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            ps[p][i] += __shfl_xor_sync(0xffffffffu, ps[p][i], 1);
                            ps[p][i] += __shfl_xor_sync(0xffffffffu, ps[p][i], 2);
                            ps[p][i] += __shfl_xor_sync(0xffffffffu, ps[p][i], 4);
                        }
                        if (q == 0 && rc[p].ok) {
                            *reinterpret_cast<float4 *>(sum_s + urow[p] * 8) = make_float4(ps[p][0], ps[p][1], ps[p][2], ps[p][3]);
                            sum_s[urow[p] * 8 + 4] = ps[p][4];
                        }
                    }
                }
            }
            LIME_TICK(4);
        } else {
            // ---------------- issuer warp --------------------------------------------------------------
            // claim the next work unit and run parts 1 and 2 of its front end while the compute warps are in phase 1
#ifdef LIME_TC_PHASE_CLOCKS
            const long long tf0 = clock64();
#endif
            if (lane == 0) next_unit = atomicAdd(args.work_counter, 1);
            next_unit = __shfl_sync(0xffffffffu, next_unit, 0);
            front_history(args, base + OFF_UB + (ubi ^ 1) * kUnitBuf, next_unit, lane);
#ifdef LIME_TC_PHASE_CLOCKS
            if (lane == 0) atomicAdd(&g_phase_clocks[10], (unsigned long long)(clock64() - tf0));
#endif
            const uint32_t sb = tc::smem_u32(base);
            const uint32_t idesc1 = tc::idesc_f16_f32(128, 3 * Up), idesc2 = tc::idesc_f16_f32(128, Up);
            const uint64_t bdesc = tc::smem_desc_sw128(sb + OFF_O);
            for (int kb = 0; kb < kOStages; ++kb) {
#ifdef LIME_TC_PHASE_CLOCKS
                const long long tw0 = clock64();
#endif
                tc::mbar_wait(bars + B_OFULL, (pass_iter * (uint32_t)kOStages + (uint32_t)kb) & 1u, 400 + kb);
#ifdef LIME_TC_PHASE_CLOCKS
                if (lane == 0) atomicAdd(&g_phase_clocks[1], (unsigned long long)(clock64() - tw0));
#endif
                {
                    const int tl = kb & 1;
#ifdef LIME_TC_PHASE_CLOCKS
                    const long long tw1 = clock64();
#endif
                    tc::mbar_wait(bars + B_WFULL + tl, (pass_iter * w_uses(tl) + (uint32_t)(kb >> 1)) & 1u, 300 + kb);
#ifdef LIME_TC_PHASE_CLOCKS
                    if (lane == 0) atomicAdd(&g_phase_clocks[0], (unsigned long long)(clock64() - tw1));
#endif
                    tc::fence_after_sync();
                    if (lane == 0) {
                        const int ksteps = kb < kOStages - 1 ? 4 : (kD - 64 * (kOStages - 1)) / 16;
                        const uint32_t wt = sb + OFF_W + (uint32_t)tl * kWTile;
                        const uint64_t ahi = tc::smem_desc_sw128(wt), alo = tc::smem_desc_sw128(wt + kWImg);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t k2 = (uint64_t)(2 * ks);             // 32 bytes per K step of 16
                            tc::mma_f16(tmem, ahi + k2, bdesc + k2, idesc1, (kb | ks) != 0);
                            tc::mma_f16(tmem + 3 * Up, alo + k2, bdesc + k2, idesc2, (kb | ks) != 0);
                        }
                        tc::mma_commit(bars + B_OFREE);
                        if (kb == kOStages - 1) tc::mma_commit(bars + B_ACCUM);
                    }
                    __syncwarp();
                }
            }
            // part 3 of the next unit's front end, in the shadow of this unit's epilogue
#ifdef LIME_TC_PHASE_CLOCKS
            const long long tf1 = clock64();
#endif
            front_cand(args, base + OFF_UB + (ubi ^ 1) * kUnitBuf, lane);
#ifdef LIME_TC_PHASE_CLOCKS
            if (lane == 0) atomicAdd(&g_phase_clocks[15], (unsigned long long)(clock64() - tf1));
#endif
        }

        if (warp < kCWarps) {
            // ---------------- epilogue: TMEM -> table rows -> Taylor combination -> LayerNorm folding -> pooling ----
            tc::mbar_wait(bars + B_ACCUM, pass_iter & 1u, 500);
            tc::fence_after_sync();
            LIME_TICK(5);
            const int qd = warp & 3;
            const bool paired = qd + 4 < kCWarps;            // quadrants 0..2 are walked by warps qd and qd + 4 in turn
            const int hb = warp >> 2;
            const int ti = lane / 3, k = lane - 3 * ti;
            const int j = 10 * qd + ti;
            const bool is_c = ti < 10 && j < cnt;
            const bool is_t = ti < 10 && j >= cnt && j < cnt + nbp;
            const int cc = is_c ? j : 0;
            const uint32_t taddr = tmem + ((uint32_t)(32 * qd) << 16);
            const int nblocks = (U + 7) >> 3;
            const int b0 = paired ? hb : 0, bstep = paired ? 2 : 1;
            // pass A: the bucket-pair rows of this quadrant publish (p0, p1) per unique row
            if (10 * qd + 10 > cnt && 10 * qd < cnt + nbp) {
                const int jt = 3 * (j - cnt) + k;
                for (int b = b0; b < nblocks; b += bstep) {
                    uint32_t x0[8], x1[8], x2[8], x3[8];
                    tmem_ld8_nowait(taddr + 8 * b, x0);
                    tmem_ld8_nowait(taddr + Up + 8 * b, x1);
                    tmem_ld8_nowait(taddr + 2 * Up + 8 * b, x2);
                    tmem_ld8_nowait(taddr + 3 * Up + 8 * b, x3);
                    tc::tmem_ld_wait();
                    if (is_t) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int u = 8 * b + i;
                            if (u < U)
                                tab_s[u * kTabStride + jt] = make_float2(__uint_as_float(x0[i]) + __uint_as_float(x1[i]) + __uint_as_float(x3[i]),
                                                                         __uint_as_float(x2[i]));
                        }
                    }
                }
            }
            bar_compute();
            // pass B: candidates.  Lane (candidate, k): k = 0 pooling logit, 1 pooled value, 2 GraphSAGE term; the softmax
            // state (m, l) is kept redundantly by the 3 lanes of a candidate, acc / ms are meaningful on lanes k = 1 / 2.
            const float bk = cscal[cc * 4 + k];
            const int tj = 3 * cbidx[cc] + k;
            const float *mp = cP[cc] == args.prefix_main ? ump0 : ump1;
            float m_run = -INFINITY, l_run = 0.0f, acc = 0.0f, ms = 0.0f;
            for (int b = b0; b < nblocks; b += bstep) {
                uint32_t x0[8], x1[8], x2[8], x3[8];
                tmem_ld8_nowait(taddr + 8 * b, x0);
                tmem_ld8_nowait(taddr + Up + 8 * b, x1);
                tmem_ld8_nowait(taddr + 2 * Up + 8 * b, x2);
                tmem_ld8_nowait(taddr + 3 * Up + 8 * b, x3);
                tc::tmem_ld_wait();
                float val[8], lg[8];
                float bm = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int u = min(8 * b + i, U - 1);
                    const float2 tb = tab_s[u * kTabStride + tj];
                    const float t = t_s[u * kAS + cc];
                    const float4 sa = *reinterpret_cast<const float4 *>(sum_s + u * 8);
                    const float q2 = sum_s[u * 8 + 4];
                    const float p0 = __uint_as_float(x0[i]) + __uint_as_float(x1[i]) + __uint_as_float(x3[i]) + tb.x;
                    const float p = fmaf(t, __uint_as_float(x2[i]) + tb.y, p0);
                    const float s0 = fmaf(t, sa.y, sa.x) * (1.0f / kOScale);
                    const float s1 = fmaf(t, fmaf(t, q2, 2.0f * sa.w), sa.z) * (1.0f / (kOScale * kOScale));
                    // LayerNorm folded into the dot: the candidate vectors are mean-centred, so x.w = rstd * (o.w)
                    const float mu = s0 * (1.0f / kD);
                    const float var = fmaxf(fmaf(-mu, mu, s1 * (1.0f / kD)), 0.0f);
                    const float rstd = rsqrtf(var + args.ln_eps) * kUnscale;
                    val[i] = fmaf(rstd, p, bk);
                    lg[i] = __shfl_sync(0xffffffffu, val[i], lane - k);
                    if (8 * b + i < U) bm = fmaxf(bm, lg[i]);
                }
                const float m_new = fmaxf(m_run, bm);
                const float sc = ex2_approx((m_run - m_new) * kLog2e);       // 0 on the first block (m_run = -inf)
                l_run *= sc;
                acc *= sc;
                m_run = m_new;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int u = 8 * b + i;
                    if (u < U) {
                        const float e = umult[u] * ex2_approx((lg[i] - m_new) * kLog2e);
                        l_run += e;
                        acc = fmaf(e, val[i], acc);
                        ms = fmaf(mp[u], val[i], ms);
                    }
                }
            }
            if (is_c) {
                float *pp = part_s + (hb * kTriples + j) * 4;
                if (k == 0) { pp[0] = m_run; pp[1] = l_run; }
                if (k == 1) pp[2] = acc;
                if (k == 2) pp[3] = ms;
            }
            // fp16 operand range of the O rows (or NaN): exact kernel
            if (tid < U && !(sum_s[tid * 8 + 2] <= kS1Max)) atomicOr(flag_s, 4);
            tc::fence_before_sync();
            bar_compute();
            LIME_TICK(6);
            // merge the two halves of a quadrant, lifetime-weighted click score (util.py:23-49); the rare P > H case
            // (user-node rows in the GraphSAGE mean) is finished below
            if (tid < cnt) {
                const int c = tid;
                const float *p0 = part_s + c * 4, *p1 = part_s + (kTriples + c) * 4;
                float m = p0[0], l = p0[1], a2 = p0[2], s2 = p0[3];
                if (c < 30 && nblocks > 1) {
                    const float m1 = p1[0];
                    const float mn = fmaxf(m, m1);
                    const float f0 = ex2_approx((m - mn) * kLog2e), f1 = ex2_approx((m1 - mn) * kLog2e);
                    l = fmaf(l, f0, p1[1] * f1);
                    a2 = fmaf(a2, f0, p1[2] * f1);
                    s2 += p1[3];
                    m = mn;
                }
                pool_s[c * 4 + 0] = m;
                pool_s[c * 4 + 1] = l;
                pool_s[c * 4 + 2] = a2;
                pool_s[c * 4 + 3] = s2;
                const int P = cP[c];
                if (args.a_matrix != nullptr) {
                    // long-history mode: the chunk's partial pooling state; lime_score_impressions_long merges the chunks
                    args.partial_out[(long long)pair0 + c] = make_float4(m, l, a2, s2);
                    if ((long long)pair0 < args.chunk_pairs) args.cbw_out[(long long)pair0 + c] = make_float2(cscal[c * 4 + 3], cw[c]);
                } else if (P <= H) {
                    args.scores[(long long)pair0 + c] = (s2 / (float)P + cscal[c * 4 + 3] + a2 / l) * cw[c];
                }
            }
            LIME_TICK(7);
        }
        ++pass_iter;
        __syncthreads();   // the candidate tiles (aliased by tab_s), the O tile and TMEM may be overwritten; the next unit buffer is ready

        if (tid == 0) {
            if ((flag_s[0] & 6) != 0) args.fallback_list[atomicAdd(args.fallback_count, 1)] = unit;
        }

        // ---------------- P > H: user-node rows take part in the GraphSAGE mean (userEncoders.py:121,153) -------
        const int Hfull = args.a_matrix != nullptr ? args.full_history : H;
        if ((args.prefix_main > Hfull || args.prefix_tail > Hfull) && (args.a_matrix == nullptr || (long long)pair0 < args.chunk_pairs)) {
            if (warp < kCWarps) {
                for (int c = warp; c < cnt; c += kCWarps) {
                    const int P = cP[c];
                    if (P > Hfull) {
                        int jn = P - Hfull - 1;
                        jn = jn < C.user_nodes ? jn : C.user_nodes - 1;
                        const float *uu = C.un_prefix + (size_t)jn * kD;
                        const float *hr2 = C.hist_rows + (size_t)cnews[c] * LIME_HIST_LD + LIME_HIST_VC;
                        const float *tr2 = C.hist_tab + (size_t)ctab[c] * LIME_HTAB_LD;
                        float un = 0.f;
                        for (int d = lane; d < kD; d += 32) un = fmaf(hr2[d] + tr2[d], uu[d], un);
                        un = warp_sum(un);
                        if (lane == 0) {
                            if (args.a_matrix != nullptr) {      // chunk 0 carries the user-node term of the prefix sum
                                args.partial_out[(long long)pair0 + c].w = pool_s[c * 4 + 3] + un;
                            } else {
                                const float bs = (pool_s[c * 4 + 3] + un) / (float)P + cscal[c * 4 + 3] + pool_s[c * 4 + 2] / pool_s[c * 4 + 1];
                                args.scores[(long long)pair0 + c] = bs * cw[c];
                            }
                        }
                    }
                }
            }
            __syncthreads();   // pool_s is rewritten by the next unit
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

// out[(tc * T + th) * 12 + head] = exp( sum_k tq[tc][k * 10 + head] * topics[th][k] + tq[tc][500 + head] ), the numerator of
// the head softmax (ex2.approx of the log2(e)-scaled logit, as the scoring kernel used to evaluate it per pair)
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
        for (int hd = 0; hd < LIME_CA_HEADS; ++hd) o[hd] = ex2_approx(acc[hd] * kLog2e);
        o[10] = 1.0f;
        o[11] = 1.0f;
    }
}

// src [rows, lds] fp32, `blocks` blocks of 400 columns -> dst [rows, 2 * blocks * 400] fp16: the hi halves of all blocks,
// then the lo halves (scale * x = hi + lo to 2^-22), i.e. [hi | lo][block][400]: one operand-row triple of the scoring
// kernel is 3 consecutive 800-byte rows; absmax[row * ldo] = max |x| of the row
__global__ void split_f16_pairs_kernel(const float *__restrict__ src, int64_t lds, int blocks, float scale,
                                       __half *__restrict__ dst, float *__restrict__ absmax, int64_t ldo) {
    __shared__ float red[4];
    const int64_t row = blockIdx.x;
    const float *s = src + row * lds;
    __half *d = dst + row * (int64_t)blocks * 2 * kD;
    float mx = 0.0f;
    for (int e = threadIdx.x; e < blocks * kD; e += blockDim.x) {
        const float x = s[e];
        const float xs = x * scale;
        const __half h = __float2half_rn(xs);
        const __half l = __float2half_rn(xs - __half2float(h));
        d[e] = h;
        d[blocks * kD + e] = l;
        mx = fmaxf(mx, fabsf(x));
        if (x != x) mx = INFINITY;
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0 && absmax != nullptr) absmax[row * ldo] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}

}  // namespace

// ---- long histories: attention pre-pass and chunk merge --------------------------------------------------------------
// Candidate-aware attention weights (layers.py:66-81) over the FULL history of H <= 224 slots, one warp per (impression,
// candidate) pair, lanes over the slots (7 per lane).  Same arithmetic as attention(): numerators of the head softmaxes
// from the exponential table, masked slots contribute 0, second softmax unmasked; all slots masked -> uniform.  Output in
// the chunk-major layout the scoring kernel reads: a[(k * total_pairs + p) * chunk_slots + (h - k * chunk_slots)].
__global__ void __launch_bounds__(256) attention_long_kernel(const ScoreArgs args, float *__restrict__ a_out, int chunks, int chunk_slots,
                                                             long long total_pairs) {
    const LimeNewsCache &C = args.cache;
    const LimeImpressions &I = args.imp;
    const int H = I.max_history, T = C.num_topics;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = blockIdx.x;
    const int imp = I.unit_imp[unit], pair0 = I.unit_pair0[unit], cnt = I.unit_count[unit];
    constexpr int SPL = 7;                       // slots per lane (H <= 224)
    int tp[SPL];
    float w[SPL];
    int nun = 0;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int h = lane + 32 * i;
        tp[i] = 0;
        w[i] = 0.0f;
        if (h < H) {
            int n = I.hist_news[(long long)imp * H + h];
            n = (n < 0 || n >= C.news_num) ? 0 : n;
            const int t = __float_as_int(__ldg(C.news_meta + (size_t)n * LIME_META_LD));
            tp[i] = (t < 0 || t >= T) ? 0 : t;
            w[i] = I.hist_mask[(long long)imp * H + h] != 0 ? 1.0f : 0.0f;
            nun += w[i] > 0.0f;
        }
    }
    nun = __reduce_add_sync(0xffffffffu, nun);
    for (int c = warp; c < cnt; c += 8) {
        const long long p = (long long)pair0 + c;
        int n = I.cand_news[p];
        n = (n < 0 || n >= C.news_num) ? 0 : n;
        int tc_ = __float_as_int(__ldg(C.news_meta + (size_t)n * LIME_META_LD));
        tc_ = (tc_ < 0 || tc_ >= T) ? 0 : tc_;
        const float *trow = C.topic_table + (size_t)tc_ * T * kTabLd;
        float e2[SPL], s2 = 0.0f;
        if (nun > 0) {
            float sum[LIME_CA_HEADS];
#pragma unroll
            for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = 0.0f;
            float4 x[SPL][3];
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const float *r0 = trow + (size_t)tp[i] * kTabLd;
                x[i][0] = ldg4(r0);
                x[i][1] = ldg4(r0 + 4);
                x[i][2] = ldg4(r0 + 8);
                const float xs[LIME_CA_HEADS] = {x[i][0].x, x[i][0].y, x[i][0].z, x[i][0].w, x[i][1].x,
                                                 x[i][1].y, x[i][1].z, x[i][1].w, x[i][2].x, x[i][2].y};
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = fmaf(w[i], xs[hd], sum[hd]);
            }
#pragma unroll
            for (int hd = 0; hd < LIME_CA_HEADS; ++hd) sum[hd] = __fdividef(1.0f, warp_sum(sum[hd]));
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const float xs[LIME_CA_HEADS] = {x[i][0].x, x[i][0].y, x[i][0].z, x[i][0].w, x[i][1].x,
                                                 x[i][1].y, x[i][1].z, x[i][1].w, x[i][2].x, x[i][2].y};
                float agg = 0.0f;
#pragma unroll
                for (int hd = 0; hd < LIME_CA_HEADS; ++hd) agg = fmaf(xs[hd], sum[hd], agg);
                agg = w[i] > 0.0f ? agg : 0.0f;
                e2[i] = lane + 32 * i < H ? ex2_approx(agg * kLog2e) : 0.0f;
                s2 += e2[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                e2[i] = lane + 32 * i < H ? 1.0f : 0.0f;
                s2 += e2[i];
            }
        }
        const float inv2 = __fdividef(1.0f, warp_sum(s2));
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            const int h = lane + 32 * i;
            if (h < H) {
                const int k = h / chunk_slots;
                a_out[((long long)k * total_pairs + p) * chunk_slots + (h - k * chunk_slots)] = e2[i] * inv2;
            }
        }
    }
}

// merge of the chunks' partial pooling states (online softmax) + lifetime-weighted click score (util.py:23-49)
__global__ void __launch_bounds__(256) merge_long_kernel(const ScoreArgs args, const float4 *__restrict__ partial, const float2 *__restrict__ cbw,
                                                         int chunks, long long total_pairs, float *__restrict__ scores) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total_pairs) return;
    float m = -INFINITY, l = 0.0f, acc = 0.0f, s2 = 0.0f;
    for (int k = 0; k < chunks; ++k) {
        const float4 q = partial[(long long)k * total_pairs + p];
        const float mn = fmaxf(m, q.x);
        const float f0 = ex2_approx((m - mn) * kLog2e), f1 = ex2_approx((q.x - mn) * kLog2e);
        l = fmaf(l, f0, q.y * f1);
        acc = fmaf(acc, f0, q.z * f1);
        s2 += q.w;
        m = mn;
    }
    const float2 cw = cbw[p];
    const int P = (args.pair_index_base + p >= args.tail_start) ? args.prefix_tail : args.prefix_main;
    scores[p] = (s2 / (float)P + cw.x + acc / l) * cw.y;
}

// The cand16 tensor map (TMA descriptor) of the candidate operand: encoded on the host by the driver's
// cuTensorMapEncodeTiled (resolved through the runtime, no link-time dependency on libcuda), cached per (pointer, rows).
static int cand16_tensor_map(const void *cand16, uint64_t rows, CUtensorMap *out) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::mutex mu;
    static EncodeFn encode = nullptr;
    static const void *c_ptr = nullptr;
    static uint64_t c_rows = 0;
    static CUtensorMap c_map;
    std::lock_guard<std::mutex> lock(mu);
    if (encode == nullptr) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LIME_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        LIME_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    if (c_ptr != cand16 || c_rows != rows) {
        const cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)(6 * rows)};          // innermost first
        const cuuint64_t strides[1] = {(cuuint64_t)(kD * 2)};                          // bytes between rows
        const cuuint32_t box[2] = {64, 3}, estr[2] = {1, 1};
        const CUresult r = encode(&c_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(cand16), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        LIME_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(cand16, %llu rows) failed: CUresult %d", (unsigned long long)rows, (int)r);
        c_ptr = cand16;
        c_rows = rows;
    }
    *out = c_map;
    return 0;
}

int launch_score_tc(const ScoreArgs &a, cudaStream_t st) {
    LIME_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));   // per device, cheap
    LIME_CUDA(cudaMemsetAsync(a.work_counter, 0, 4 * sizeof(int32_t), st));   // work counter, fallback count, exact counter, stats
    const int reps = a.cache.tab_replicas > 0 ? a.cache.tab_replicas : 1;
    CUtensorMap wmap;
    if (int rc = cand16_tensor_map(a.cache.cand16, (uint64_t)a.cache.news_num + (uint64_t)reps * a.cache.num_buckets * a.cache.num_buckets, &wmap))
        return rc;
    int grid = 2 * num_sms();
    if (const char *e = getenv("LIME_TC_ONE_CTA_PER_SM")) {      // experiment knob: concurrency vs shared resources
        if (e[0] == '1') grid = num_sms();
    }
    if (grid > a.imp.num_units) grid = a.imp.num_units;
    score_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(a, wmap);
    LIME_LAUNCH_CHECK("score_tc_kernel");
    return 0;
}

int launch_attention_long(const ScoreArgs &orig, float *a_matrix, int chunks, int chunk_slots, long long total_pairs, cudaStream_t st) {
    attention_long_kernel<<<orig.imp.num_units, 256, 0, st>>>(orig, a_matrix, chunks, chunk_slots, total_pairs);
    LIME_LAUNCH_CHECK("attention_long_kernel");
    return 0;
}

int launch_merge_long(const ScoreArgs &a, const float4 *partial, const float2 *cbw, int chunks, long long total_pairs, float *scores,
                      cudaStream_t st) {
    merge_long_kernel<<<(unsigned)((total_pairs + 255) / 256), 256, 0, st>>>(a, partial, cbw, chunks, total_pairs, scores);
    LIME_LAUNCH_CHECK("merge_long_kernel");
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
