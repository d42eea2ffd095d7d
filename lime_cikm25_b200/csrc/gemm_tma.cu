// Dense layer of the news encoder on bf16 ACTIVATIONS (Stage A, "bf16 mode"), sm_100a: TMA + tcgen05 + TMEM.
//   C = act(A . W^T + bias) + residual      A [m, k] bf16, W [n, k] bf16 (k padded to a multiple of 64 with zeros),
//                                           bias / residual fp32, C bf16 or fp32, fp32 accumulation in TMEM.
// Replaces, for newsEncoders.py:244-247 (the four GEMMs of nn.TransformerEncoderLayer, 87 % of the encoder's FLOPs), the
// fp32-in / fp32-out kernel of gemm_bf16.cu whose producers converted every operand tile from fp32 inside the kernel.
//
// The shapes are skinny (k = 320 or 512, n <= 900, m = news x tokens = millions), so the kernel is organised around what
// is re-used:  ONE persistent CTA per SM owns one N tile (bn <= 256 columns) for its whole life and keeps that slice of
// W RESIDENT in shared memory (bn x k bf16 <= 160 KB, loaded once by TMA); it then streams 128-row tiles of A through a
// 2-deep TMA ring (16 KB per 64-wide K block; measured: 2 stages + one staging tile per warp + a 256-column W slice beat 3 + 2 + 128).  Per output tile the tensor core runs k/16 MMAs (M = 128, N = bn) into
// one of TWO TMEM accumulators (2 x 256 columns), so the epilogue of tile i (TMEM -> registers -> bias / act / residual
// -> global) overlaps the MMAs of tile i + 1.  The CTAs that share an M tile (one per N tile) run at the same time on
// neighbouring SMs: A comes from HBM once and from L2 otherwise.
//
// Warp roles (320 threads): warps 0-7 epilogue (warp w owns TMEM lanes 32 (w & 3).. = rows of the tile and every second
// 32-column chunk), warp 8 TMA producer (one elected lane), warp 9 TMEM allocation + MMA issue (one elected lane).
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <mutex>

#include "tc05.cuh"

namespace lime {
namespace {

#ifndef LIME_GT_STAGES
#define LIME_GT_STAGES 2
#endif
constexpr int GT_M = 128, GT_STAGES = LIME_GT_STAGES, GT_A_BYTES = GT_M * 128;
#ifndef LIME_GT_STAGE_BUFS
#define LIME_GT_STAGE_BUFS 1
#endif
constexpr int GT_STAGE_BUFS = LIME_GT_STAGE_BUFS;   // epilogue staging tiles per warp
constexpr int GT_STAGE_BYTES = 4096;           // epilogue staging tile of one warp: 32 rows x 128 B
constexpr int GT_W_MAX = (224 - 16 * GT_STAGES - 32 * GT_STAGE_BUFS) * 1024;   // resident W slice: what the A ring and the staging tiles leave
#ifndef LIME_GT_EPI_WARPS
#define LIME_GT_EPI_WARPS 8
#endif
constexpr int GT_EPI_WARPS = LIME_GT_EPI_WARPS;   // two epilogue warps per TMEM lane quadrant, alternating 32-column chunks (measured: 4 -> 8 warps takes the bf16-output GEMMs from 0.48 to 0.32 ms per 262k rows; staging the tile through shared memory for row-contiguous stores was slower)
constexpr int GT_THREADS = 32 * (GT_EPI_WARPS + 2);
constexpr int GT_SMEM = GT_W_MAX + GT_STAGES * GT_A_BYTES + GT_STAGE_BUFS * GT_EPI_WARPS * GT_STAGE_BYTES + 1024 /* bias slice */ + 256 /* barriers */;
static_assert(GT_SMEM + 1024 <= 232448, "shared memory");

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float act_apply(int act, float x) {
    if (act == 1) return fmaxf(x, 0.0f);
    if (act == 2) return tanhf(x);
    return x;
}

// barriers (uint64 each)
enum { GB_WFULL = 0, GB_AFULL = 1, GB_AEMPTY = 1 + GT_STAGES, GB_ACCFULL = 1 + 2 * GT_STAGES, GB_ACCEMPTY = 3 + 2 * GT_STAGES,
       GB_RES = 5 + 2 * GT_STAGES /* one per epilogue warp */, GB_COUNT = GB_RES + GT_EPI_WARPS };
static_assert(GB_COUNT * 8 + 8 <= 256, "barrier area");

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap,
                const __grid_constant__ CUtensorMap amap2, const __grid_constant__ CUtensorMap wmap2, int x3,
                const __grid_constant__ CUtensorMap cmap, const __grid_constant__ CUtensorMap cmap2, float out_scale,
                const __grid_constant__ CUtensorMap rmap, int tma_epilogue, int ncov,
                const float *__restrict__ bias, const float *__restrict__ residual, int64_t ldr, void *__restrict__ Cout,
                int64_t ldc, int c_bf16, int64_t m, int n, int nkb, int bn, int n_tiles, int act, float alpha, int ab_fp16) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw;
    if (threadIdx.x == 0 && (tc::smem_u32(smem_raw) & 1023u) != 0) __trap();
    unsigned char *w_s = base;                                     // [nkb][bn rows x 128 B]
    unsigned char *a_s = base + GT_W_MAX;                          // [GT_STAGES][128 rows x 128 B]
    unsigned char *stage_s = base + GT_W_MAX + GT_STAGES * GT_A_BYTES;          // [GT_EPI_WARPS][2][4096], 1024-byte aligned
    float *bias_s = reinterpret_cast<float *>(stage_s + GT_STAGE_BUFS * GT_EPI_WARPS * GT_STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(bias_s) + 1024);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + GB_COUNT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool res_pre = act >= 16;                                // LIME_ACT_RES_FIRST: the residual joins the sum BEFORE the activation
    act &= 15;
    const int nt = (int)(blockIdx.x % (unsigned)n_tiles);          // this CTA's N tile, for its whole life
    const int col0 = nt * bn;
    const int64_t m_tiles = (m + GT_M - 1) / GT_M;
    const int64_t mt0 = blockIdx.x / (unsigned)n_tiles, mt_step = gridDim.x / (unsigned)n_tiles;
    // fp32x3 mode (x3 != 0): A = amap (hi) + amap2 (lo), W = wmap (hi) + wmap2 (lo), both W images resident (slots 0..nkb-1 hi,
    // nkb.. lo).  Per 64-wide K block the ring carries A_lo then A_hi (2 nkb steps per tile): A_lo feeds small += A_lo.W_hi,
    // A_hi feeds small += A_hi.W_lo AND big += A_hi.W_hi -- every A image crosses the L2 -> shared-memory path once (with
    // bn <= 128 the 8 CTAs of an M tile already read it at the L2's per-SM rate).  TWO accumulators per tile (bn <= 128: big at
    // column 0, small at column 128 of the tile's 256-column half), summed by the epilogue: the tensor core aligns every addend to
    // the accumulator's exponent and truncates, so small products added onto a running hi.hi sum lose their low bits
    // (measured with one accumulator, hi.hi first: 8.6e-6 instead of 3e-6 of the row scale).
    const int nsteps = x3 ? 2 * nkb : nkb;

    if (tid == 0) {
        tc::mbar_init(bars + GB_WFULL, 1);
        for (int s = 0; s < GT_STAGES; ++s) {
            tc::mbar_init(bars + GB_AFULL + s, 1);
            tc::mbar_init(bars + GB_AEMPTY + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(bars + GB_ACCFULL + a, 1);
            tc::mbar_init(bars + GB_ACCEMPTY + a, GT_EPI_WARPS);   // one arrival per epilogue warp
        }
        for (int w = 0; w < GT_EPI_WARPS; ++w) tc::mbar_init(bars + GB_RES + w, 1);
        tc::mbar_fence_init();
    }
    for (int i = tid; i < bn; i += GT_THREADS) bias_s[i] = (bias != nullptr && col0 + i < n) ? bias[col0 + i] : 0.0f;
    if (warp == GT_EPI_WARPS + 1) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == GT_EPI_WARPS) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            mbar_expect_tx(bars + GB_WFULL, (uint32_t)((x3 ? 2 : 1) * nkb * bn * 128));
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(tc::smem_u32(w_s) + (uint32_t)(kb * bn * 128), &wmap, 64 * kb, col0, bars + GB_WFULL);
            if (x3)
                for (int kb = 0; kb < nkb; ++kb) tma_load_2d(tc::smem_u32(w_s) + (uint32_t)((nkb + kb) * bn * 128), &wmap2, 64 * kb, col0, bars + GB_WFULL);
            uint32_t it = 0;
            for (int64_t mt = mt0; mt < m_tiles; mt += mt_step) {
                for (int i = 0; i < nsteps; ++i, ++it) {
                    const int kb = x3 ? i >> 1 : i;
                    const bool lo_img = x3 && (i & 1) == 0;
                    const int s = (int)(it % GT_STAGES);
                    const uint32_t ph = (it / GT_STAGES) & 1u;
                    tc::mbar_wait(bars + GB_AEMPTY + s, ph ^ 1u);
                    mbar_expect_tx(bars + GB_AFULL + s, GT_A_BYTES);
                    tma_load_2d(tc::smem_u32(a_s) + (uint32_t)(s * GT_A_BYTES), lo_img ? &amap2 : &amap, 64 * kb, (int)(mt * GT_M), bars + GB_AFULL + s);
                }
            }
        }
    } else if (warp == GT_EPI_WARPS + 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = ab_fp16 ? tc::idesc_f16_f32(GT_M, bn) : tc::idesc_bf16_f32(GT_M, bn);
            tc::mbar_wait(bars + GB_WFULL, 0);
            tc::fence_after_sync();
            uint32_t it = 0, tile = 0;
            for (int64_t mt = mt0; mt < m_tiles; mt += mt_step, ++tile) {
                const uint32_t acc = tile & 1u;
                tc::mbar_wait(bars + GB_ACCEMPTY + acc, ((tile >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
                tc::fence_after_sync();
                for (int i = 0; i < nsteps; ++i, ++it) {
                    const int kb = x3 ? i >> 1 : i;
                    const bool lo_img = x3 && (i & 1) == 0;
                    const int s = (int)(it % GT_STAGES);
                    tc::mbar_wait(bars + GB_AFULL + s, (it / GT_STAGES) & 1u);
                    tc::fence_after_sync();
                    const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(a_s) + (uint32_t)(s * GT_A_BYTES));
                    const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(w_s) + (uint32_t)(kb * bn * 128));
                    if (x3) {
                        const uint64_t dbl = tc::smem_desc_sw128(tc::smem_u32(w_s) + (uint32_t)((nkb + kb) * bn * 128));
                        if (lo_img) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc::mma_bf16(tmem + acc * 256u + 128u, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (kb | ks) != 0);
                        } else {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                tc::mma_bf16(tmem + acc * 256u + 128u, da + (uint64_t)(2 * ks), dbl + (uint64_t)(2 * ks), idesc, true);
                                tc::mma_bf16(tmem + acc * 256u, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (kb | ks) != 0);
                            }
                        }
                        tc::mma_commit(bars + GB_AEMPTY + s);
                        continue;
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc::mma_bf16(tmem + acc * 256u, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (i | ks) != 0);
                    tc::mma_commit(bars + GB_AEMPTY + s);
                }
                tc::mma_commit(bars + GB_ACCFULL + acc);
            }
        }
    } else {
        // ---------------- epilogue (warps 0-7) ----------------
        const int quad = warp & 3, half = warp >> 2;
        uint32_t tile = 0;
        if (tma_epilogue) {
            // ---- TMA epilogue: the output leaves as 32 x 32 boxes (rows x columns) by cp.async.bulk.tensor stores ----
            // TMEM hands every lane one ROW of the tile, and a plain store from that layout touches 32 rows per instruction:
            // measured, the four GEMMs then run at 1.6 TB/s of output traffic, 2.5x slower than their main loops (0.126 ms vs
            // 0.315 ms for the in_proj GEMM of 262k rows).  Here a lane writes its row into a swizzled 32-row staging tile in
            // shared memory (128-byte rows = 32 fp32 or 64 bf16 columns, SWIZZLE_128B: conflict-free 16-byte stores) and one lane
            // hands the tile to the TMA unit, which writes whole lines and clips the box at the tensor's edge (no tail code;
            // the zero padding columns of a bf16 output are part of the box).  An fp32 residual comes the same way: TMA-loaded
            // into the staging tile, added in place.  Two staging tiles per warp alternate.
            const int ncols_t = min(bn, ncov - col0);              // columns of this N tile inside the stored width
            unsigned char *my_stage = stage_s + warp * GT_STAGE_BUFS * GT_STAGE_BYTES;
            uint64_t *rbar = bars + GB_RES + warp;
            uint32_t chunk_it = 0, res_it = 0;
            for (int64_t mt = mt0; mt < m_tiles; mt += mt_step, ++tile) {
                const uint32_t acc = tile & 1u;
                const int row0 = (int)(mt * GT_M) + quad * 32;
                const uint32_t tlane = tmem + acc * 256u + ((uint32_t)(quad * 32) << 16);
                bool waited = false;
                const int cw = c_bf16 ? 64 : 32;                   // columns per box: 128-byte rows either way
                for (int c0 = cw * half; c0 < ncols_t; c0 += cw * (GT_EPI_WARPS / 4), ++chunk_it) {
                    unsigned char *buf = my_stage + (chunk_it % GT_STAGE_BUFS) * GT_STAGE_BYTES;
                    if (lane == 0) {
                        bulk_wait_read<GT_STAGE_BUFS - 1>();       // the store that last read this tile is done with it
                        if (residual != nullptr) {
                            mbar_expect_tx(rbar, 32 * 32 * 4);
                            tma_load_2d(tc::smem_u32(buf), &rmap, col0 + c0, row0, rbar);
                        }
                    }
                    __syncwarp();
                    if (!waited) {                                 // (the residual of the first chunk is in flight while the MMAs finish)
                        tc::mbar_wait(bars + GB_ACCFULL + acc, (tile >> 1) & 1u);
                        tc::fence_after_sync();
                        waited = true;
                    }
                    // 128-byte rows, SWIZZLE_128B: 16-byte unit u of row r at position u ^ (r & 7)
                    const CUtensorMap *smap = &cmap;
                    if (c_bf16 == 2) {
                        // fp16 PAIR output (the next fp32x3 layer's operand): out_scale * y = hi + lo, hi through cmap, lo through cmap2.
                        // One staging tile: the hi image is handed to the TMA unit, then the tile is refilled with the lo image
                        // (TMEM is simply read again).
#pragma unroll 1
                        for (int img = 0; img < 2; ++img) {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                uint32_t v[32];
                                tmem_ld32(tlane + (uint32_t)(c0 + 32 * hh), v);
                                if (x3) {
                                    uint32_t v2[32];
                                    tmem_ld32(tlane + 128u + (uint32_t)(c0 + 32 * hh), v2);
#pragma unroll
                                    for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    uint32_t o[4];
#pragma unroll
                                    for (int e2 = 0; e2 < 4; ++e2) {
                                        const int c = 8 * u + 2 * e2;
                                        const float ya = act_apply(act, fmaf(alpha, __uint_as_float(v[c]), bias_s[c0 + 32 * hh + c])) * out_scale;
                                        const float yb = act_apply(act, fmaf(alpha, __uint_as_float(v[c + 1]), bias_s[c0 + 32 * hh + c + 1])) * out_scale;
                                        const __half2 h2 = __floats2half2_rn(ya, yb);
                                        const float2 hf = __half22float2(h2);
                                        const __half2 l2 = __floats2half2_rn(ya - hf.x, yb - hf.y);
                                        o[e2] = img == 0 ? *reinterpret_cast<const uint32_t *>(&h2) : *reinterpret_cast<const uint32_t *>(&l2);
                                    }
                                    *reinterpret_cast<uint4 *>(buf + lane * 128 + 16 * ((4 * hh + u) ^ (lane & 7))) = make_uint4(o[0], o[1], o[2], o[3]);
                                }
                            }
                            if (img == 0) {
                                tc::fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_2d(&cmap, col0 + c0, row0, tc::smem_u32(buf));
                                    bulk_commit();
                                    bulk_wait_read<0>();           // the tile is refilled with the lo image
                                }
                                __syncwarp();
                            }
                        }
                        smap = &cmap2;
                    } else if (c_bf16) {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {           // bn is a multiple of 64 here: both halves are columns of this tile
                            uint32_t v[32];
                            tmem_ld32(tlane + (uint32_t)(c0 + 32 * hh), v);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float4 b0 = *reinterpret_cast<const float4 *>(bias_s + c0 + 32 * hh + 8 * u);
                                const float4 b1 = *reinterpret_cast<const float4 *>(bias_s + c0 + 32 * hh + 8 * u + 4);
                                uint4 o;
                                o.x = tc::pack_bf16(act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u]), b0.x)), act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u + 1]), b0.y)));
                                o.y = tc::pack_bf16(act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u + 2]), b0.z)), act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u + 3]), b0.w)));
                                o.z = tc::pack_bf16(act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u + 4]), b1.x)), act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u + 5]), b1.y)));
                                o.w = tc::pack_bf16(act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u + 6]), b1.z)), act_apply(act, fmaf(alpha, __uint_as_float(v[8 * u + 7]), b1.w)));
                                *reinterpret_cast<uint4 *>(buf + lane * 128 + 16 * ((4 * hh + u) ^ (lane & 7))) = o;
                            }
                        }
                    } else {
                        uint32_t v[32];
                        tmem_ld32(tlane + (uint32_t)c0, v);
                        if (x3) {                                  // big + small accumulator
                            uint32_t v2[32];
                            tmem_ld32(tlane + 128u + (uint32_t)c0, v2);
#pragma unroll
                            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
                        }
                        if (residual != nullptr) {
                            tc::mbar_wait(rbar, res_it & 1u);
                            ++res_it;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 bb = *reinterpret_cast<const float4 *>(bias_s + c0 + 4 * u);
                            float4 *p = reinterpret_cast<float4 *>(buf + lane * 128 + 16 * (u ^ (lane & 7)));
                            float4 y = make_float4(fmaf(alpha, __uint_as_float(v[4 * u]), bb.x), fmaf(alpha, __uint_as_float(v[4 * u + 1]), bb.y),
                                                   fmaf(alpha, __uint_as_float(v[4 * u + 2]), bb.z), fmaf(alpha, __uint_as_float(v[4 * u + 3]), bb.w));
                            float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (residual != nullptr) rv = *p;
                            if (res_pre) { y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w; }
                            y = make_float4(act_apply(act, y.x), act_apply(act, y.y), act_apply(act, y.z), act_apply(act, y.w));
                            if (!res_pre) { y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w; }
                            *p = y;
                        }
                    }
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(smap, col0 + c0, row0, tc::smem_u32(buf));
                        bulk_commit();
                    }
                }
                if (!waited) {                                     // a warp without a chunk in this N tile still takes part in the hand-over
                    tc::mbar_wait(bars + GB_ACCFULL + acc, (tile >> 1) & 1u);
                    tc::fence_after_sync();
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bars + GB_ACCEMPTY + acc);
            }
            if (lane == 0) bulk_wait_read<0>();                    // shared memory must outlive the last stores' reads
            __syncwarp();
        } else {
        const int ncols = min(bn, n - col0);                       // real columns of this N tile
        const bool pad = c_bf16 && nt == n_tiles - 1;              // bf16 output: the padding columns [n, ldc) are written as zeros
        for (int64_t mt = mt0; mt < m_tiles; mt += mt_step, ++tile) {
            const uint32_t acc = tile & 1u;
            tc::mbar_wait(bars + GB_ACCFULL + acc, (tile >> 1) & 1u);
            tc::fence_after_sync();
            const int64_t r = mt * GT_M + quad * 32 + lane;
            const uint32_t tlane = tmem + acc * 256u + ((uint32_t)(quad * 32) << 16);
            for (int c0 = 32 * half; c0 < ncols; c0 += 8 * GT_EPI_WARPS) {
                uint32_t v[32];
                tmem_ld32(tlane + (uint32_t)c0, v);
                if (x3) {                                          // big + small accumulator
                    uint32_t v2[32];
                    tmem_ld32(tlane + 128u + (uint32_t)c0, v2);
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
                }
#ifdef LIME_GT_NOSTORE           // timing diagnostic only (wrong results): the epilogue drains TMEM and stores nothing
                if (v[0] == 0x7fc12345u && r < m) reinterpret_cast<float *>(Cout)[0] = 1.0f;
                continue;
#endif
                if (r < m) {
                    const int c = col0 + c0;
                    const int lim = min(32, ncols - c0);
                    float x[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bb = *reinterpret_cast<const float4 *>(bias_s + c0 + 4 * q);
                        x[4 * q] = fmaf(alpha, __uint_as_float(v[4 * q]), bb.x);
                        x[4 * q + 1] = fmaf(alpha, __uint_as_float(v[4 * q + 1]), bb.y);
                        x[4 * q + 2] = fmaf(alpha, __uint_as_float(v[4 * q + 2]), bb.z);
                        x[4 * q + 3] = fmaf(alpha, __uint_as_float(v[4 * q + 3]), bb.w);
                    }
                    if (!res_pre) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) x[e] = act_apply(act, x[e]);
                    }
                    if (residual != nullptr) {
                        const float *rp = residual + r * ldr + c;
                        if (lim == 32 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 rr = __ldg(reinterpret_cast<const float4 *>(rp) + q);
                                x[4 * q] += rr.x; x[4 * q + 1] += rr.y; x[4 * q + 2] += rr.z; x[4 * q + 3] += rr.w;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (e < lim) x[e] += rp[e];
                        }
                    }
                    if (res_pre) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) x[e] = act_apply(act, x[e]);
                    }
                    if (c_bf16) {
                        __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(Cout) + r * ldc + c;
                        if (lim == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 o;
                                o.x = tc::pack_bf16(x[8 * q], x[8 * q + 1]);
                                o.y = tc::pack_bf16(x[8 * q + 2], x[8 * q + 3]);
                                o.z = tc::pack_bf16(x[8 * q + 4], x[8 * q + 5]);
                                o.w = tc::pack_bf16(x[8 * q + 6], x[8 * q + 7]);
                                reinterpret_cast<uint4 *>(op)[q] = o;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (e < lim) op[e] = __float2bfloat16_rn(x[e]);
                        }
                    } else {
                        float *op = reinterpret_cast<float *>(Cout) + r * ldc + c;
                        if (lim == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) reinterpret_cast<float4 *>(op)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (e < lim) op[e] = x[e];
                        }
                    }
                }
            }
            if (pad && half == 0 && r < m) {
                __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(Cout) + r * ldc;
                for (int64_t c = n; c < ldc; ++c) op[c] = __float2bfloat16_rn(0.0f);
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bars + GB_ACCEMPTY + acc);
        }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == GT_EPI_WARPS + 1) tc::tmem_dealloc(tmem, 512);
}


// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient of an nn.Linear in bf16 mode, dW = dZ^T . X (trainer.py:131-146 differentiates newsEncoders.py:244-247):
//   C[mo, no] (+)= alpha * sum_{r < k} A[r, i] * B[r, j]        A [k, lda], B [k, ldb] bf16, row-major over the TOKENS r
// Both operands are MN-major for the tensor core (contiguous along the output dimension), so nothing is transposed: a TMA
// box of 64 columns x 64 token rows with SWIZZLE_128B lands exactly as one 64-wide MN block of the canonical MN-major
// layout (8 K-rows x 128 B atoms, 16-byte chunk c of K-row r at c ^ (r & 7)); the instruction descriptor's major bits
// select it.  One CTA = one 128 x bn output tile (bn <= 256, whole 64-column blocks) over one K split; 4-stage TMA ring
// (<= 48 KB per stage), warp 0 = producer, warp 1 = MMA issue, warps 2-5 = epilogue (atomicAdd when K is split).
// Replaces gemm_bf16_general_kernel<0, 0> whose producers converted fp32 tiles through registers (50 TFLOP/s).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TN_STAGES = 4, TN_K = 64, TN_M = 128, TN_NMAX = 256;
constexpr int TN_A_BYTES = TN_M * TN_K * 2, TN_B_BYTES = TN_NMAX * TN_K * 2, TN_STAGE = TN_A_BYTES + TN_B_BYTES;
constexpr int TN_SMEM = TN_STAGES * TN_STAGE + 256;
constexpr int TN_THREADS = 192;
static_assert(TN_SMEM + 1024 <= 232448, "shared memory");

__device__ __forceinline__ uint64_t smem_desc_mn_sw128_tn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(8192 >> 4) << 16;     // leading byte offset: next 64-wide MN block
    d |= (uint64_t)(1024 >> 4) << 32;     // stride byte offset: next group of 8 K-rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;               // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tn_tma_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, float *__restrict__ C,
                   int64_t ldc, int mo, int no, int64_t k, int64_t k_per_split, int bn, int n_tiles, float alpha, int accumulate,
                   uint32_t idesc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw;
    if (threadIdx.x == 0 && (tc::smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + TN_STAGES * TN_STAGE);
    uint64_t *full = bars, *empty = bars + TN_STAGES, *accum = bars + 2 * TN_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TN_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = (int)(blockIdx.x / (unsigned)n_tiles) * TN_M;
    const int col0 = (int)(blockIdx.x % (unsigned)n_tiles) * bn;
    const int64_t kbeg = (int64_t)blockIdx.y * k_per_split;
    const int64_t kend = kbeg + k_per_split < k ? kbeg + k_per_split : k;
    const int kchunks = (int)((kend - kbeg + TN_K - 1) / TN_K);
    const int nblk = bn / 64;

    if (tid == 0) {
        for (int s = 0; s < TN_STAGES; ++s) {
            tc::mbar_init(full + s, 1);
            tc::mbar_init(empty + s, 1);
        }
        tc::mbar_init(accum, 1);
        tc::mbar_fence_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kc = 0; kc < kchunks; ++kc) {
                const int s = kc % TN_STAGES;
                tc::mbar_wait(empty + s, (((uint32_t)(kc / TN_STAGES)) & 1u) ^ 1u);
                const uint32_t st = tc::smem_u32(base + s * TN_STAGE);
                const int kr = (int)(kbeg + (int64_t)kc * TN_K);
                // rows beyond k (and columns beyond the image) arrive as zeros; a box always counts its full size
                mbar_expect_tx(full + s, (uint32_t)((2 + nblk) * 8192));
                tma_load_2d(st, &amap, row0, kr, full + s);
                tma_load_2d(st + 8192, &amap, row0 + 64, kr, full + s);
                for (int b = 0; b < nblk; ++b) tma_load_2d(st + TN_A_BYTES + 8192 * b, &bmap, col0 + 64 * b, kr, full + s);
            }
        }
    } else if (warp == 1) {
        for (int kc = 0; kc < kchunks; ++kc) {
            const int s = kc % TN_STAGES;
            tc::mbar_wait(full + s, ((uint32_t)(kc / TN_STAGES)) & 1u);
            tc::fence_after_sync();
            if (lane == 0) {
                const uint32_t st = tc::smem_u32(base + s * TN_STAGE);
                const uint64_t da = smem_desc_mn_sw128_tn(st), db = smem_desc_mn_sw128_tn(st + TN_A_BYTES);
                // rows past kend inside the last chunk belong to the next split (or are zeros past k): whole K steps only
                const int64_t rem = kend - (kbeg + (int64_t)kc * TN_K);
                const int ksteps = rem >= TN_K ? TN_K / 16 : (int)((rem + 15) / 16);
                for (int ks = 0; ks < ksteps; ++ks)     // K step of 16 token rows = 2 groups of 8 K-rows = 2048 B
                    tc::mma_bf16(tmem, da + (uint64_t)(128 * ks), db + (uint64_t)(128 * ks), idesc, (kc | ks) != 0);
                tc::mma_commit(empty + s);
                if (kc == kchunks - 1) tc::mma_commit(accum);
            }
            __syncwarp();
        }
    } else {
        tc::mbar_wait(accum, 0);
        tc::fence_after_sync();
        const int q = warp & 3;                              // the TMEM lane quadrant this warp may read
        const int r = row0 + 32 * q + lane;
        const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
        const bool atomic = gridDim.y > 1;
        for (int c0 = 0; c0 < bn; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tlane + (uint32_t)c0, v);
            const int c = col0 + c0;
            if (r >= mo || c >= no) continue;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                if (c + e < no) {
                    float *o = C + (int64_t)r * ldc + c + e;
                    const float x = alpha * v[e];
                    if (atomic) atomicAdd(o, x);
                    else *o = accumulate ? *o + x : x;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int tensor_map_2d(CUtensorMap *out, CUtensorMapDataType dt, int esize, const void *ptr, uint64_t cols, uint64_t rows, uint64_t ld,
                  uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
    static std::mutex mu;
    static EncodeFn encode = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (encode == nullptr) {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            LIME_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            LIME_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
            encode = reinterpret_cast<EncodeFn>(fn);
        }
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)(ld * esize)};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    const CUresult r = encode(out, dt, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LIME_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: CUresult %d (cols %llu rows %llu ld %llu)", (int)r,
                   (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld);
    return 0;
}
int tensor_map_bf16_2d(CUtensorMap *out, const void *ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    return tensor_map_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, cols, rows, ld, 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace
}  // namespace lime

// A2 / W2 != NULL: the fp32x3 form (lime_linear_x3_tma), lo images of A and W beside the hi images
// c_is_bf16 == 2: fp16 pair output (C = hi image, C2 = lo image, out_scale), the x3 form only
static int linear_tma_launch(const void *A, const void *A2, int64_t lda, const void *W, const void *W2, int64_t ldw, const float *bias,
                             const float *residual, int64_t ldr, void *C, int64_t ldc, int32_t c_is_bf16,
                             int64_t m, int32_t n, int32_t k, int32_t act, float alpha, int32_t ab_is_fp16, void *stream,
                             void *C2 = nullptr, float out_scale = 1.0f) {
    using namespace lime;
    const int x3 = A2 != nullptr;
    LIME_CHECK_ARG(A && W && C && (A2 != nullptr) == (W2 != nullptr), "lime_linear_bf16_tma: null argument");
    LIME_CHECK_ARG(!x3 || ((((uintptr_t)A2 | (uintptr_t)W2) & 15) == 0 && c_is_bf16 != 1), "lime_linear_x3_tma: lo operands must be 16-byte aligned, output fp32");
    LIME_CHECK_ARG(c_is_bf16 != 2 || (x3 && C2 != nullptr && residual == nullptr && ((uintptr_t)C2 & 15) == 0), "lime_linear_x3_pairs_tma: needs the x3 form, both images 16-byte aligned, no residual");
    LIME_CHECK_ARG(k >= 64 && k % 64 == 0 && k <= 512, "lime_linear_bf16_tma: k=%d must be a multiple of 64 in [64, 512] (pad with zeros)", k);
    LIME_CHECK_ARG(n >= 1 && lda >= k && ldw >= k && lda % 8 == 0 && ldw % 8 == 0 && ldc >= n,
                   "lime_linear_bf16_tma: bad leading dimensions (lda %lld ldw %lld ldc %lld, n %d k %d)", (long long)lda,
                   (long long)ldw, (long long)ldc, n, k);
    LIME_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "lime_linear_bf16_tma: operands must be 16-byte aligned");
    LIME_CHECK_ARG((act & 15) >= 0 && (act & 15) <= 2 && (act >> 4) <= 1, "lime_linear_bf16_tma: act=%d", act);
    if (m <= 0) return 0;
    const int nkb = k / 64;
    // TMA epilogue (stores, and the residual loads): needs 16-byte aligned rows; bf16 output without residual, or fp32 output with an
    // optional fp32 residual (the combinations Stage A uses); anything else takes the plain-store epilogue
    const int esz = c_is_bf16 ? 2 : 4;
    const bool tma_epi = ((uintptr_t)C & 15) == 0 && (ldc * esz) % 16 == 0 && m < (int64_t)1 << 31 && !(c_is_bf16 && residual != nullptr) &&
                         (residual == nullptr || (((uintptr_t)residual & 15) == 0 && (ldr * 4) % 16 == 0 && ldr >= n));
    const int ncov = tma_epi && c_is_bf16 ? (int)ldc : n;          // stored width: a bf16 output includes its zero padding columns
    LIME_CHECK_ARG(ncov - n < 64, "lime_linear_bf16_tma: ldc %lld leaves more than 63 padding columns after n %d", (long long)ldc, n);
    // N tile: the widest multiple of 32 (<= 256) whose W slice fits the resident area, then balanced over the tiles
    int bn_max = GT_W_MAX / ((x3 ? 2 : 1) * nkb * 128);
    bn_max = bn_max > (x3 ? 128 : 256) ? (x3 ? 128 : 256) : (bn_max / 32) * 32;      // x3: two accumulators share a 256-column half of TMEM
    const int gran = tma_epi && c_is_bf16 ? 64 : 32;              // bf16 boxes are 64 columns wide
    bn_max = bn_max / gran * gran;
    const int n_tiles = (ncov + bn_max - 1) / bn_max;
    int bn = (((ncov + n_tiles - 1) / n_tiles) + gran - 1) / gran * gran;
    LIME_CHECK_ARG(bn <= bn_max && n_tiles <= 64, "lime_linear_bf16_tma: n=%d does not tile", n);
    LIME_CHECK_ARG(c_is_bf16 != 2 || tma_epi, "lime_linear_x3_pairs_tma: the pair output needs 16-byte aligned rows (ld16 a multiple of 8)");
    CUtensorMap amap, wmap, amap2, wmap2, cmap, cmap2, rmap;
    if (int rc = tensor_map_bf16_2d(&amap, A, (uint64_t)k, (uint64_t)m, (uint64_t)lda, GT_M)) return rc;
    if (int rc = tensor_map_bf16_2d(&wmap, W, (uint64_t)k, (uint64_t)n, (uint64_t)ldw, (uint32_t)bn)) return rc;
    amap2 = amap;
    wmap2 = wmap;
    if (x3) {
        if (int rc = tensor_map_bf16_2d(&amap2, A2, (uint64_t)k, (uint64_t)m, (uint64_t)lda, GT_M)) return rc;
        if (int rc = tensor_map_bf16_2d(&wmap2, W2, (uint64_t)k, (uint64_t)n, (uint64_t)ldw, (uint32_t)bn)) return rc;
    }
    cmap = amap;
    cmap2 = amap;
    rmap = amap;
    if (tma_epi) {
        if (int rc = tensor_map_2d(&cmap, c_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, esz, C, (uint64_t)ncov,
                                   (uint64_t)m, (uint64_t)ldc, c_is_bf16 ? 64 : 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))
            return rc;
        cmap2 = cmap;
        if (c_is_bf16 == 2)
            if (int rc = tensor_map_2d(&cmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C2, (uint64_t)ncov, (uint64_t)m, (uint64_t)ldc, 64, 32,
                                       CU_TENSOR_MAP_SWIZZLE_128B))
                return rc;
        if (residual != nullptr)
            if (int rc = tensor_map_2d(&rmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, residual, (uint64_t)n, (uint64_t)m, (uint64_t)ldr, 32, 32,
                                       CU_TENSOR_MAP_SWIZZLE_128B))
                return rc;
    }
    LIME_CUDA(cudaFuncSetAttribute(gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM));
    const int64_t m_tiles = (m + GT_M - 1) / GT_M;
    int groups = num_sms() / n_tiles;
    if (groups < 1) groups = 1;
    if (groups > m_tiles) groups = (int)m_tiles;
    gemm_tma_kernel<<<groups * n_tiles, GT_THREADS, GT_SMEM, as_stream(stream)>>>(amap, wmap, amap2, wmap2, x3, cmap, cmap2, out_scale, rmap, tma_epi ? 1 : 0, ncov, bias, residual, ldr,
                                                                                    C, ldc, c_is_bf16, m, n, nkb, bn, n_tiles, act, alpha, ab_is_fp16);
    LIME_LAUNCH_CHECK("gemm_tma_kernel");
    return 0;
}

extern "C" int lime_linear_bf16_tma(const void *A, int64_t lda, const void *W, int64_t ldw, const float *bias,
                                    const float *residual, int64_t ldr, void *C, int64_t ldc, int32_t c_is_bf16,
                                    int64_t m, int32_t n, int32_t k, int32_t act, float alpha, int32_t ab_is_fp16, void *stream) {
    return linear_tma_launch(A, nullptr, lda, W, nullptr, ldw, bias, residual, ldr, C, ldc, c_is_bf16, m, n, k, act, alpha, ab_is_fp16, stream);
}

// fp32x3 dense layer in ONE launch: C = act(alpha ((Alo Whi^T + Ahi Wlo^T) + Ahi Whi^T) + bias [+ residual]) -- the small products and
// the hi.hi product accumulate in two TMEM tiles that the epilogue sums, so the fp32 output is written once instead of being
// re-read and re-written by two more accumulating passes, and every A image is staged once per K block.
extern "C" int lime_linear_x3_tma(const void *Ahi, const void *Alo, int64_t lda, const void *Whi, const void *Wlo, int64_t ldw,
                                  const float *bias, const float *residual, int64_t ldr, float *C, int64_t ldc,
                                  int64_t m, int32_t n, int32_t k, int32_t act, float alpha, int32_t ab_is_fp16, void *stream) {
    LIME_CHECK_ARG(Alo && Wlo, "lime_linear_x3_tma: null argument");
    return linear_tma_launch(Ahi, Alo, lda, Whi, Wlo, ldw, bias, residual, ldr, C, ldc, 0, m, n, k, act, alpha, ab_is_fp16, stream);
}

// ... with the result leaving as the NEXT x3 layer's operand pair: out_scale * act(alpha (...) + bias) = hi + lo (fp16, each [m, ld16],
// columns n..ld16-1 zero) instead of an fp32 matrix that a lime_split_bf16_pairs pass would re-read (the FFN hidden layer).
extern "C" int lime_linear_x3_pairs_tma(const void *Ahi, const void *Alo, int64_t lda, const void *Whi, const void *Wlo, int64_t ldw,
                                        const float *bias, void *Chi, void *Clo, int64_t ld16, float out_scale,
                                        int64_t m, int32_t n, int32_t k, int32_t act, float alpha, int32_t ab_is_fp16, void *stream) {
    LIME_CHECK_ARG(Alo && Wlo && Chi && Clo, "lime_linear_x3_pairs_tma: null argument");
    return linear_tma_launch(Ahi, Alo, lda, Whi, Wlo, ldw, bias, nullptr, 0, Chi, ld16, 2, m, n, k, act, alpha, ab_is_fp16, stream, Clo, out_scale);
}

// dW = dZ^T . X on bf16 images (see gemm_tn_tma_kernel): C[m, n] (+)= alpha * sum_{r < k} A[r, i] * B[r, j]
extern "C" int lime_gemm_bf16_tn_tma(const void *A, int64_t lda, const void *B, int64_t ldb, float *C, int64_t ldc, int32_t m,
                                     int32_t n, int64_t k, float alpha, int32_t accumulate, void *stream) {
    using namespace lime;
    LIME_CHECK_ARG(A && B && C && m > 0 && n > 0 && k > 0, "lime_gemm_bf16_tn_tma: bad argument");
    LIME_CHECK_ARG(lda >= m && ldb >= n && lda % 8 == 0 && ldb % 8 == 0 && ldc >= n, "lime_gemm_bf16_tn_tma: bad leading dimensions (lda %lld ldb %lld ldc %lld)",
                   (long long)lda, (long long)ldb, (long long)ldc);
    LIME_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "lime_gemm_bf16_tn_tma: operands must be 16-byte aligned");
    LIME_CHECK_ARG(k < ((int64_t)1 << 31), "lime_gemm_bf16_tn_tma: k too large");
    cudaStream_t st = as_stream(stream);
    const int ntiles0 = (n + TN_NMAX - 1) / TN_NMAX;
    int bn = (((n + ntiles0 - 1) / ntiles0) + 63) / 64 * 64;
    if (bn > TN_NMAX) bn = TN_NMAX;
    const int n_tiles = (n + bn - 1) / bn;
    const int m_tiles = (m + TN_M - 1) / TN_M;
    const int64_t tiles = (int64_t)m_tiles * n_tiles;
    // one CTA per SM: split K until the grid covers the GPU about twice (tail balance), at least 512 token rows per split
    int64_t splits = 1;
    const int64_t target = 2LL * num_sms();
    if (tiles < target && k >= 1024) {
        splits = (target + tiles - 1) / tiles;
        const int64_t max_splits = k / 512;
        if (splits > max_splits) splits = max_splits;
        if (splits > 65535) splits = 65535;
        if (splits < 1) splits = 1;
    }
    int64_t kps = (k + splits - 1) / splits;
    kps = (kps + TN_K - 1) / TN_K * TN_K;
    splits = (k + kps - 1) / kps;
    if (splits > 1 && !accumulate) {
        if (ldc == n) {
            LIME_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)m * n, st));
        } else {
            LIME_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * n, (size_t)m, st));
        }
    }
    CUtensorMap amap, bmap;
    if (int rc = tensor_map_2d(&amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, (uint64_t)lda, (uint64_t)k, (uint64_t)lda, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = tensor_map_2d(&bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, (uint64_t)ldb, (uint64_t)k, (uint64_t)ldb, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const uint32_t idesc = tc::idesc_bf16_f32(TN_M, bn) | (1u << 15) | (1u << 16);       // both operands MN-major
    LIME_CUDA(cudaFuncSetAttribute(gemm_tn_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
    dim3 grid((unsigned)tiles, (unsigned)splits);
    gemm_tn_tma_kernel<<<grid, TN_THREADS, TN_SMEM, st>>>(amap, bmap, C, ldc, m, n, k, kps, bn, n_tiles, alpha, accumulate, idesc);
    LIME_LAUNCH_CHECK("gemm_tn_tma_kernel");
    return 0;
}
