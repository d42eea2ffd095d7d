// "bf16 mode" dense layer on the 5th-generation tensor cores (sm_100a):
//   C = act(bf16(A) . bf16(W)^T + bias) + residual,  fp32 accumulation in TMEM, fp32 in / fp32 out.
// Same contract as lime_linear (gemm.cu); used for the four transformer GEMMs of the news encoder
// (newsEncoders.py:244-247) when NewsEncoderEngine.bf16 is set.  Parity bar: metrics within 1e-3.
//
// One CTA = one 128 x bn output tile (bn <= 256, multiple of 16).  Warp roles:
//   warps 0-3  producers: fp32 global -> bf16 -> shared memory in the canonical K-major SWIZZLE_128B
//              layout (tc05.cuh), 64-wide K chunks, 2-stage ring guarded by full/empty mbarriers;
//              afterwards the same four warps are the epilogue (warp w owns TMEM lanes 32w..32w+31).
//   warp 4     allocates TMEM; lane 0 issues tcgen05.mma (M=128, N=bn, K=16) and commits to mbarriers.
// Two CTAs fit on an SM (96 KB + 256 TMEM columns each), so one CTA's epilogue overlaps the other's
// main loop.
#include "common.cuh"
#include "tc05.cuh"

namespace lime {

constexpr int GB_M = 128, GB_K = 64, GB_STAGES = 2, GB_NMAX = 256;
constexpr int GB_A_BYTES = GB_M * 128, GB_W_BYTES = GB_NMAX * 128;
constexpr int GB_STAGE_BYTES = GB_A_BYTES + GB_W_BYTES;
constexpr int GB_SMEM = GB_STAGES * GB_STAGE_BYTES + 1024 /* alignment slack */ + 64 /* barriers */;
constexpr int GB_THREADS = 160;

__device__ __forceinline__ float act_apply(int act, float x) {
    if (act == 1) return fmaxf(x, 0.0f);
    if (act == 2) return tanhf(x);
    return x;
}

// rows x 64 fp32 (row stride ld, valid rows < rows_valid, valid k < k_total) -> bf16 swizzled tile
__device__ __forceinline__ void stage_tile(unsigned char *tile, const float *__restrict__ src, int64_t ld,
                                           int64_t row0, int64_t rows_valid, int rows, int k0, int k_total, int tid) {
    for (int t = tid; t < rows * 8; t += 128) {
        const int row = t >> 3, chunk = t & 7;
        const int kk = k0 + chunk * 8;
        const int64_t r = row0 + row;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (r < rows_valid) {
            const float *p = src + r * ld + kk;
            if (kk + 4 <= k_total) a = *reinterpret_cast<const float4 *>(p);
            if (kk + 8 <= k_total) b = *reinterpret_cast<const float4 *>(p + 4);
        }
        uint4 o;
        o.x = tc::pack_bf16(a.x, a.y);
        o.y = tc::pack_bf16(a.z, a.w);
        o.z = tc::pack_bf16(b.x, b.y);
        o.w = tc::pack_bf16(b.z, b.w);
        *reinterpret_cast<uint4 *>(tile + tc::sw128_offset(row, chunk)) = o;
    }
}

__global__ void __launch_bounds__(GB_THREADS, 2)
linear_bf16_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ W, int64_t ldw,
                   const float *__restrict__ bias, const float *__restrict__ residual, int64_t ldr,
                   float *__restrict__ C, int64_t ldc, int64_t m, int n, int k, int bn, int act, uint32_t idesc) {
    // pointer arithmetic on the declared array only (an integer round trip would lose the shared state space and turn
    // every access below into a generic load / store); the declaration requests the swizzle's 1024-byte alignment
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw;
    if (threadIdx.x == 0 && (tc::smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + GB_STAGES * GB_STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + GB_STAGES, *accum = bars + 2 * GB_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * GB_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (n + bn - 1) / bn;      // N tiles of one M tile are adjacent CTAs: the A tile is an L2 hit
    const int64_t row0 = (int64_t)(blockIdx.x / ntiles) * GB_M;
    const int col0 = (int)(blockIdx.x % ntiles) * bn;
    const int kchunks = (k + GB_K - 1) / GB_K;

    if (tid == 0) {
        for (int s = 0; s < GB_STAGES; ++s) {
            tc::mbar_init(full + s, 128);
            tc::mbar_init(empty + s, 1);
        }
        tc::mbar_init(accum, 1);
        tc::mbar_fence_init();
    }
    if (warp == 4) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        // ---------------- producers ----------------
        for (int kc = 0; kc < kchunks; ++kc) {
            const int s = kc % GB_STAGES;
            const uint32_t ph = (uint32_t)(kc / GB_STAGES) & 1u;
            tc::mbar_wait(empty + s, ph ^ 1u);
            unsigned char *st = base + s * GB_STAGE_BYTES;
            stage_tile(st, A, lda, row0, m, GB_M, kc * GB_K, k, tid);
            stage_tile(st + GB_A_BYTES, W, ldw, col0, n, bn, kc * GB_K, k, tid);
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(full + s);
        }
        // ---------------- epilogue ----------------
        tc::mbar_wait(accum, 0);
        tc::fence_after_sync();
        const int64_t r = row0 + warp * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
        const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                         (residual == nullptr || (((ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(residual) & 15) == 0)));
        for (int c0 = 0; c0 < bn; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tlane + (uint32_t)c0, v);
            const int c = col0 + c0;
            if (r >= m || c >= n) continue;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float x = v[e];
                if (bias != nullptr && c + e < n) x += bias[c + e];
                v[e] = act_apply(act, x);
            }
            if (vec && c + 16 <= n) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    if (residual != nullptr) {
                        const float4 rr = *reinterpret_cast<const float4 *>(residual + r * ldr + c + 4 * q);
                        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                    }
                    *reinterpret_cast<float4 *>(C + r * ldc + c + 4 * q) = o;
                }
            } else {
                for (int e = 0; e < 16 && c + e < n; ++e) {
                    float x = v[e];
                    if (residual != nullptr) x += residual[r * ldr + c + e];
                    C[r * ldc + c + e] = x;
                }
            }
        }
    } else {
        // ---------------- MMA issuer ----------------
        for (int kc = 0; kc < kchunks; ++kc) {
            const int s = kc % GB_STAGES;
            const uint32_t ph = (uint32_t)(kc / GB_STAGES) & 1u;
            tc::mbar_wait(full + s, ph);
            tc::fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = tc::smem_u32(base + s * GB_STAGE_BYTES);
                const uint64_t da = tc::smem_desc_sw128(a_addr), db = tc::smem_desc_sw128(a_addr + GB_A_BYTES);
                const int rem = k - kc * GB_K;
                const int ksteps = rem >= GB_K ? GB_K / 16 : (rem + 15) / 16;
                for (int ks = 0; ks < ksteps; ++ks)
                    tc::mma_bf16(tmem, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (kc | ks) != 0);
                tc::mma_commit(empty + s);
                if (kc == kchunks - 1) tc::mma_commit(accum);
            }
            __syncwarp();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 256);
}


// ---------------------------------------------------------------------------------------------------
// General bf16 tensor-core GEMM for the backward passes in "bf16 mode":  C (+)= alpha * op(A) . op(B),
// same operand conventions as lime_gemm (train_kernels.cu).  A K-major source (contiguous along the
// contraction) is staged exactly like above; an MN-major source (contiguous along M or N, e.g. dY and X in
// dW = dY^T X, or W in dX = dY W) is staged WITHOUT transposition into the canonical MN-major SWIZZLE_128B
// layout and the instruction descriptor's major bits tell tcgen05.mma:
//   atom = 64 MN-elements (128 B) x 8 K-rows (1024 B), 16-byte chunk c of K-row r at position c ^ (r & 7);
//   atoms of one 64-wide MN block follow each other along K (stride byte offset 1024 B),
//   MN blocks are 8192 B apart (leading byte offset) inside a 64-deep K stage.
// Split-K (gridDim.y) with atomicAdd accumulation for the weight gradients (K = all tokens).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(8192 >> 4) << 16;     // leading byte offset: next 64-wide MN block
    d |= (uint64_t)(1024 >> 4) << 32;     // stride byte offset: next group of 8 K-rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// 64 K-rows x `width` MN-elements (width multiple of 64) of an MN-major fp32 source -> bf16 MN-major tile.
// element (mn, kk) at src[(k0 + kk) * ld + mn0 + mn]
__device__ __forceinline__ void stage_tile_mn(unsigned char *tile, const float *__restrict__ src, int64_t ld, int64_t mn0,
                                              int64_t mn_lim, int width, int64_t k0, int64_t k_lim, int tid) {
    const int chunks = width >> 3;                       // 16-byte chunks per K-row
    for (int t = tid; t < 64 * chunks; t += 128) {
        const int kk = t / chunks, c = t - kk * chunks;
        const int64_t mn = mn0 + 8 * c;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (k0 + kk < k_lim && mn < mn_lim) {
            const float *p = src + (k0 + kk) * ld + mn;
            if (mn + 8 <= mn_lim && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (mn + e < mn_lim) v[e] = p[e];
            }
        }
        uint4 o;
        o.x = tc::pack_bf16(v[0], v[1]);
        o.y = tc::pack_bf16(v[2], v[3]);
        o.z = tc::pack_bf16(v[4], v[5]);
        o.w = tc::pack_bf16(v[6], v[7]);
        const uint32_t off = (uint32_t)(c >> 3) * 8192u + (uint32_t)(kk >> 3) * 1024u + (uint32_t)(kk & 7) * 128u +
                             (uint32_t)(((c & 7) ^ (kk & 7)) << 4);
        *reinterpret_cast<uint4 *>(tile + off) = o;
    }
}

// K-major source with 64-bit row / k offsets (the contraction may be all tokens)
__device__ __forceinline__ void stage_tile_k(unsigned char *tile, const float *__restrict__ src, int64_t ld, int64_t row0,
                                             int64_t rows_valid, int rows, int64_t k0, int64_t k_lim, int tid) {
    for (int t = tid; t < rows * 8; t += 128) {
        const int row = t >> 3, chunk = t & 7;
        const int64_t kk = k0 + chunk * 8;
        const int64_t r = row0 + row;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < rows_valid && kk < k_lim) {
            const float *p = src + r * ld + kk;
            if (kk + 8 <= k_lim && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (kk + e < k_lim) v[e] = p[e];
            }
        }
        uint4 o;
        o.x = tc::pack_bf16(v[0], v[1]);
        o.y = tc::pack_bf16(v[2], v[3]);
        o.z = tc::pack_bf16(v[4], v[5]);
        o.w = tc::pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4 *>(tile + tc::sw128_offset(row, chunk)) = o;
    }
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(GB_THREADS, 2)
gemm_bf16_general_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb,
                         float *__restrict__ C, int64_t ldc, int64_t m, int n, int64_t k, int64_t k_per_split, int bn,
                         float alpha, int accumulate, uint32_t idesc) {
    // pointer arithmetic on the declared array only (an integer round trip would lose the shared state space and turn
    // every access below into a generic load / store); the declaration requests the swizzle's 1024-byte alignment
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *base = smem_raw;
    if (threadIdx.x == 0 && (tc::smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + GB_STAGES * GB_STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + GB_STAGES, *accum = bars + 2 * GB_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * GB_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (n + bn - 1) / bn;
    const int64_t row0 = (int64_t)(blockIdx.x / ntiles) * GB_M;
    const int col0 = (int)(blockIdx.x % ntiles) * bn;
    const int64_t kbeg = (int64_t)blockIdx.y * k_per_split;
    const int64_t kend = kbeg + k_per_split < k ? kbeg + k_per_split : k;
    const int kchunks = (int)((kend - kbeg + GB_K - 1) / GB_K);

    if (tid == 0) {
        for (int s = 0; s < GB_STAGES; ++s) {
            tc::mbar_init(full + s, 128);
            tc::mbar_init(empty + s, 1);
        }
        tc::mbar_init(accum, 1);
        tc::mbar_fence_init();
    }
    if (warp == 4) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        for (int kc = 0; kc < kchunks; ++kc) {
            const int s = kc % GB_STAGES;
            const uint32_t ph = (uint32_t)(kc / GB_STAGES) & 1u;
            tc::mbar_wait(empty + s, ph ^ 1u);
            unsigned char *st = base + s * GB_STAGE_BYTES;
            const int64_t k0 = kbeg + (int64_t)kc * GB_K;
            if (A_KMAJOR) stage_tile_k(st, A, lda, row0, m, GB_M, k0, kend, tid);
            else stage_tile_mn(st, A, lda, row0, m, GB_M, k0, kend, tid);
            if (B_KMAJOR) stage_tile_k(st + GB_A_BYTES, B, ldb, col0, n, bn, k0, kend, tid);
            else stage_tile_mn(st + GB_A_BYTES, B, ldb, col0, n, bn, k0, kend, tid);
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(full + s);
        }
        tc::mbar_wait(accum, 0);
        tc::fence_after_sync();
        const int64_t r = row0 + warp * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
        const bool atomic = gridDim.y > 1;
        for (int c0 = 0; c0 < bn; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tlane + (uint32_t)c0, v);
            const int c = col0 + c0;
            if (r >= m || c >= n) continue;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                if (c + e < n) {
                    float *o = C + r * ldc + c + e;
                    const float x = alpha * v[e];
                    if (atomic) atomicAdd(o, x);
                    else *o = accumulate ? *o + x : x;
                }
            }
        }
    } else {
        for (int kc = 0; kc < kchunks; ++kc) {
            const int s = kc % GB_STAGES;
            const uint32_t ph = (uint32_t)(kc / GB_STAGES) & 1u;
            tc::mbar_wait(full + s, ph);
            tc::fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = tc::smem_u32(base + s * GB_STAGE_BYTES);
                const uint64_t da = A_KMAJOR ? tc::smem_desc_sw128(a_addr) : smem_desc_mn_sw128(a_addr);
                const uint64_t db = B_KMAJOR ? tc::smem_desc_sw128(a_addr + GB_A_BYTES) : smem_desc_mn_sw128(a_addr + GB_A_BYTES);
                const int64_t rem = kend - (kbeg + (int64_t)kc * GB_K);
                const int ksteps = rem >= GB_K ? GB_K / 16 : (int)((rem + 15) / 16);
                for (int ks = 0; ks < ksteps; ++ks) {
                    // K step of 16: +32 B inside the 128-byte row (K-major) or +2 groups of 8 K-rows = 2048 B (MN-major)
                    const uint64_t ka = A_KMAJOR ? (uint64_t)(2 * ks) : (uint64_t)(128 * ks);
                    const uint64_t kb = B_KMAJOR ? (uint64_t)(2 * ks) : (uint64_t)(128 * ks);
                    tc::mma_bf16(tmem, da + ka, db + kb, idesc, (kc | ks) != 0);
                }
                tc::mma_commit(empty + s);
                if (kc == kchunks - 1) tc::mma_commit(accum);
            }
            __syncwarp();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 256);
}

}  // namespace lime

using namespace lime;

extern "C" int lime_linear_bf16(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias,
                                const float *residual, int64_t ldr, float *C, int64_t ldc, int64_t m,
                                int n, int k, int act, void *stream) {
    LIME_CHECK_ARG(A && W && C, "lime_linear_bf16: null pointer");
    LIME_CHECK_ARG(m >= 0 && n > 0 && k > 0, "lime_linear_bf16: bad shape m=%lld n=%d k=%d", (long long)m, n, k);
    LIME_CHECK_ARG((k & 3) == 0 && (lda & 3) == 0 && (ldw & 3) == 0, "lime_linear_bf16: k, lda, ldw must be multiples of 4");
    LIME_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "lime_linear_bf16: A and W must be 16-byte aligned");
    LIME_CHECK_ARG(act >= 0 && act <= 2, "lime_linear_bf16: act %d", act);
    if (m == 0) return 0;
    // balanced N tiles: fewest tiles of at most 256 columns, each a multiple of 16
    const int ntiles = (n + GB_NMAX - 1) / GB_NMAX;
    const int bn = (((n + ntiles - 1) / ntiles) + 15) & ~15;
    const int64_t mtiles = (m + GB_M - 1) / GB_M;
    const int nt = (n + bn - 1) / bn;
    LIME_CHECK_ARG(mtiles * nt < (int64_t)1 << 31, "lime_linear_bf16: m too large for one launch (%lld rows)", (long long)m);
    // per device and cheap: set on every call (a process-wide "done" flag would miss a second GPU or a second thread)
    LIME_CUDA(cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM));
    linear_bf16_kernel<<<(unsigned)(mtiles * nt), GB_THREADS, GB_SMEM, as_stream(stream)>>>(
        A, lda, W, ldw, bias, residual, ldr, C, ldc, m, n, k, bn, act, tc::idesc_bf16_f32(GB_M, bn));
    LIME_LAUNCH_CHECK("linear_bf16_kernel");
    return 0;
}

// bf16 tensor-core counterpart of lime_gemm (same operand conventions): C (+)= alpha * op(A) . op(B)
extern "C" int lime_gemm_bf16(const float *A, int64_t lda, int a_kmajor, const float *B, int64_t ldb, int b_kmajor, float *C,
                              int64_t ldc, int64_t m, int n, int64_t k, float alpha, int accumulate, void *stream) {
    LIME_CHECK_ARG(A && B && C && m > 0 && n > 0 && k > 0, "lime_gemm_bf16: bad argument");
    cudaStream_t st = as_stream(stream);
    // N tiles: K-major B may use any multiple of 16 columns, an MN-major B whole 64-column blocks
    const int gran = b_kmajor ? 16 : 64;
    const int ntiles = (n + GB_NMAX - 1) / GB_NMAX;
    int bn = (((n + ntiles - 1) / ntiles) + gran - 1) / gran * gran;
    if (bn > GB_NMAX) bn = GB_NMAX;
    const int nt = (n + bn - 1) / bn;
    const int64_t mtiles = (m + GB_M - 1) / GB_M;
    const int64_t tiles = mtiles * nt;
    LIME_CHECK_ARG(tiles < (int64_t)1 << 31, "lime_gemm_bf16: too many tiles");
    int64_t splits = 1;
    const int64_t target = 2LL * num_sms();
    if (tiles < target && k >= 2048) {
        splits = (target + tiles - 1) / tiles;
        const int64_t max_splits = k / 512;
        if (splits > max_splits) splits = max_splits;
        if (splits > 65535) splits = 65535;
        if (splits < 1) splits = 1;
    }
    int64_t kps = (k + splits - 1) / splits;
    kps = (kps + GB_K - 1) / GB_K * GB_K;
    splits = (k + kps - 1) / kps;
    if (splits > 1 && !accumulate) {
        if (ldc == n) {
            LIME_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)m * n, st));
        } else {
            LIME_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * n, (size_t)m, st));
        }
    }
    const uint32_t idesc = tc::idesc_bf16_f32(GB_M, bn) | (a_kmajor ? 0u : (1u << 15)) | (b_kmajor ? 0u : (1u << 16));
    dim3 grid((unsigned)tiles, (unsigned)splits);
#define LIME_LAUNCH_GB(AK, BK)                                                                                         \
    do {                                                                                                               \
        LIME_CUDA(cudaFuncSetAttribute(gemm_bf16_general_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM)); \
        gemm_bf16_general_kernel<AK, BK><<<grid, GB_THREADS, GB_SMEM, st>>>(A, lda, B, ldb, C, ldc, m, n, k, kps, bn, alpha,   \
                                                                          accumulate, idesc);                          \
    } while (0)
    if (a_kmajor && b_kmajor) LIME_LAUNCH_GB(true, true);
    else if (a_kmajor) LIME_LAUNCH_GB(true, false);
    else if (b_kmajor) LIME_LAUNCH_GB(false, true);
    else LIME_LAUNCH_GB(false, false);
#undef LIME_LAUNCH_GB
    LIME_LAUNCH_CHECK("gemm_bf16_general_kernel");
    return 0;
}
