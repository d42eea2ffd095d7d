// Stage A glue kernels of the LIME/CROWN news encoder (sm_100a): everything in
// newsEncoders.CROWN.forward (newsEncoders.py:302-373) and FreshnessEncoder (:53-83) that is not a
// dense layer.  The dense layers are lime_linear / lime_linear_bf16.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace lime {

// ---- FreshnessEncoder.bucketize ---------------------------------------------------------------
__global__ void bucketize_kernel(const float *__restrict__ x, int64_t n, float scale, int nb,
                                 int32_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bucketize_seconds(x[i], scale, nb);
}

// ---- word embedding gather + positional encoding (newsEncoders.py:311-315, 822-828) ------------
// one thread per 128-bit piece of an output row; the embedding row (d*4 bytes, 16-byte multiple for
// d = 300) is read with one LDG.128 per thread, coalesced across the row.
__global__ void __launch_bounds__(256)
embed_pe_kernel(const float *__restrict__ E, int64_t vocab, const int32_t *__restrict__ ids,
                int64_t rows, int T, int d4, const float *__restrict__ pe, float *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * d4) return;
    const int64_t r = idx / d4;
    const int q = (int)(idx - r * d4);
    int64_t id = ids[r];
    id = (id < 0 || id >= vocab) ? 0 : id;
    const float4 e = reinterpret_cast<const float4 *>(E)[id * d4 + q];
    const float4 p = reinterpret_cast<const float4 *>(pe)[(r % T) * d4 + q];
    reinterpret_cast<float4 *>(out)[idx] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
}

// the same with a second, bf16 image of the row (the A operand of the TMA GEMM): [rows, ld16], columns d.. zero
__global__ void __launch_bounds__(256)
embed_pe_bf16_kernel(const float *__restrict__ E, int64_t vocab, const int32_t *__restrict__ ids, int64_t rows, int T, int d4,
                     const float *__restrict__ pe, float *__restrict__ out, __nv_bfloat16 *__restrict__ out16, int ld16) {
    const int q4 = ld16 / 4;                         // 4-element pieces per bf16 row (incl. padding)
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * q4) return;
    const int64_t r = idx / q4;
    const int q = (int)(idx - r * q4);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < d4) {
        int64_t id = ids[r];
        id = (id < 0 || id >= vocab) ? 0 : id;
        const float4 e = reinterpret_cast<const float4 *>(E)[id * d4 + q];
        const float4 p = reinterpret_cast<const float4 *>(pe)[(r % T) * d4 + q];
        o = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
        reinterpret_cast<float4 *>(out)[r * d4 + q] = o;
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t *>(&lo);
    pk.y = *reinterpret_cast<uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(out16 + r * ld16 + 4 * q) = pk;
}

// the same with the fp16 operand PAIR of the fp32x3 mode beside the fp32 row: scale * x = hi + lo, each [rows, ld16], columns d.. zero
__global__ void __launch_bounds__(256)
embed_pe_pairs_kernel(const float *__restrict__ E, int64_t vocab, const int32_t *__restrict__ ids, int64_t rows, int T, int d4,
                      const float *__restrict__ pe, float *__restrict__ out, __half *__restrict__ hi16, __half *__restrict__ lo16, int ld16,
                      float scale) {
    const int q4 = ld16 / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * q4) return;
    const int64_t r = idx / q4;
    const int q = (int)(idx - r * q4);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < d4) {
        int64_t id = ids[r];
        id = (id < 0 || id >= vocab) ? 0 : id;
        const float4 e = reinterpret_cast<const float4 *>(E)[id * d4 + q];
        const float4 p = reinterpret_cast<const float4 *>(pe)[(r % T) * d4 + q];
        o = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
        reinterpret_cast<float4 *>(out)[r * d4 + q] = o;
    }
    const float v[4] = {o.x * scale, o.y * scale, o.z * scale, o.w * scale};
    uint32_t h[2], l[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        h[e] = *reinterpret_cast<const uint32_t *>(&hh);
        l[e] = *reinterpret_cast<const uint32_t *>(&ll);
    }
    *reinterpret_cast<uint2 *>(hi16 + r * ld16 + 4 * q) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2 *>(lo16 + r * ld16 + 4 * q) = make_uint2(l[0], l[1]);
}

// ---- multi-head self-attention core, no mask (nn.MultiheadAttention inside the encoder layer) ---
// 128 threads: one query row per thread, G = 128/T heads of one news per block.  K and V of the
// block's heads sit in shared memory (rows padded to 32 floats); every thread walks the same key j
// at the same time, so all K/V reads are warp broadcasts.  Online softmax in chunks of 8 keys.
__device__ __forceinline__ float ldf(const float *p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float *p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// QT / OT = float (fp32 mode and the training path) or bf16 (Stage A in bf16 mode: qkv [rows, ldq] from the TMA GEMM, ctx
// [rows, ldo] is the next GEMM's A operand, K-padded with zeros to ldo columns)
template <int T, typename QT, typename OT>
__global__ void __launch_bounds__(128)
mha_kernel(const QT *__restrict__ qkv, OT *__restrict__ ctx, int64_t ld, int64_t ldo, int d, int nhead, int hd, float scale,
           float p_drop, uint64_t seed, int64_t news0) {
    constexpr int G = 128 / T;
    constexpr int HP = 32;
    __shared__ __align__(16) float Ks[G][T][HP];
    __shared__ __align__(16) float Vs[G][T][HP];
    const int64_t news = blockIdx.y;
    const int head0 = blockIdx.x * G;
    const int tid = threadIdx.x;
    const QT *base = qkv + news * T * ld;

    for (int idx = tid; idx < G * T * HP; idx += 128) {
        const int e = idx % HP;
        const int j = (idx / HP) % T;
        const int g = idx / (HP * T);
        const int head = head0 + g;
        float kv = 0.f, vv = 0.f;
        if (head < nhead && e < hd) {
            kv = ldf(base + j * ld + d + head * hd + e);
            vv = ldf(base + j * ld + 2 * d + head * hd + e);
        }
        Ks[g][j][e] = kv;
        Vs[g][j][e] = vv;
    }
    __syncthreads();

    const int g = tid / T;
    const int i = tid - g * T;
    const int head = head0 + g;
    if (head >= nhead) return;
    float q[HP], acc[HP];
#pragma unroll
    for (int e = 0; e < HP; ++e) {
        q[e] = (e < hd) ? ldf(base + i * ld + head * hd + e) * scale : 0.0f;
        acc[e] = 0.0f;
    }
    float m = -INFINITY, l = 0.0f;
    // dropout on the attention weights (nn.MultiheadAttention(dropout=p) inside nn.TransformerEncoderLayer,
    // newsEncoders.py:244-247): P~ = P * keep / (1 - p) with the stateless mask of (seed, news, head, i, j); the
    // softmax normaliser l keeps every term, only the value accumulation is masked
    const uint64_t drop0 = (((uint64_t)(news0 + news) * nhead + head) * T + i) * T;
    for (int j0 = 0; j0 < T; j0 += 8) {
        float s[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float4 *kr = reinterpret_cast<const float4 *>(&Ks[g][j0 + jj][0]);
            float a = 0.0f;
#pragma unroll
            for (int e4 = 0; e4 < HP / 4; ++e4) {
                const float4 kk = kr[e4];
                a = fmaf(q[4 * e4 + 0], kk.x, a);
                a = fmaf(q[4 * e4 + 1], kk.y, a);
                a = fmaf(q[4 * e4 + 2], kk.z, a);
                a = fmaf(q[4 * e4 + 3], kk.w, a);
            }
            s[jj] = a;
        }
        float cm = s[0];
#pragma unroll
        for (int jj = 1; jj < 8; ++jj) cm = fmaxf(cm, s[jj]);
        const float mn = fmaxf(m, cm);
        const float corr = __expf(m - mn);
        l *= corr;
#pragma unroll
        for (int e = 0; e < HP; ++e) acc[e] *= corr;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            float p = __expf(s[jj] - mn);
            l += p;
            if (p_drop > 0.0f) p *= drop_scale(seed, drop0 + j0 + jj, p_drop);
            const float4 *vr = reinterpret_cast<const float4 *>(&Vs[g][j0 + jj][0]);
#pragma unroll
            for (int e4 = 0; e4 < HP / 4; ++e4) {
                const float4 vv = vr[e4];
                acc[4 * e4 + 0] = fmaf(p, vv.x, acc[4 * e4 + 0]);
                acc[4 * e4 + 1] = fmaf(p, vv.y, acc[4 * e4 + 1]);
                acc[4 * e4 + 2] = fmaf(p, vv.z, acc[4 * e4 + 2]);
                acc[4 * e4 + 3] = fmaf(p, vv.w, acc[4 * e4 + 3]);
            }
        }
        m = mn;
    }
    const float inv = 1.0f / l;
    OT *o = ctx + (news * T + i) * ldo + head * hd;
#pragma unroll
    for (int e = 0; e < HP; ++e)
        if (e < hd) stf(o + e, acc[e] * inv);
    if (head == nhead - 1) {                    // K padding of the next GEMM's operand
        for (int64_t c = d; c < ldo; ++c) stf(ctx + (news * T + i) * ldo + c, 0.0f);
    }
}

// ---- the same attention core on the tensor cores, bf16 activations (Stage A in bf16 mode) ---------------------------
// One CTA = one news, 8 warps; a warp owns 16 query rows of one head, so T / 16 warps cover a head and 8 / (T / 16)
// heads are in flight per pass (T = 128: one head per pass, T = 32: four).  Q, K, V head slices sit in shared memory as
// [T][40] bf16 (30 dims, zero padded; the 80-byte pitch keeps ldmatrix conflict-free).  S = Q K^T by mma.m16n8k16 with
// the whole key range in registers (T / 8 accumulator tiles: no online softmax needed), softmax on the accumulator
// fragments (row max / sum across the 4 lanes of a quad), P re-used in place as the A fragments of O = P V (V through
// ldmatrix.trans).  tcgen05 would need M = 128 query rows per instruction and a TMEM round trip for the softmax of a
// 30-dim head; the warp-level MMA keeps the 9 % of the encoder's FLOPs that live here off the FP32 pipe.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void *p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void *p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

// Head-padded layout (``hp`` = 32): qkv [rows, 3 * nhead * 32] holds q | k | v with every head slice padded to 32
// columns (the in_proj weight rows are permuted and zero-padded on the host, so the padding columns are exact zeros):
// a head slice of a row is 64 contiguous, 16-byte aligned bytes and goes to shared memory by cp.async, one head pass
// ahead of the arithmetic (two tile sets).
__device__ __forceinline__ void cp_async_16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int T>
__global__ void __launch_bounds__(256)
mha_tc_kernel(const __nv_bfloat16 *__restrict__ qkv, __nv_bfloat16 *__restrict__ ctx, int64_t ld, int64_t ldo, int d, int nhead,
              int hd, float scale_log2e) {
    constexpr int WPH = T / 16;                 // warps per head
    constexpr int HPC = 8 / WPH;                // heads per pass
    constexpr int P = 40;                       // shared-memory row pitch (bf16 elements): 80 bytes, ldmatrix conflict-free
    constexpr int NT = T / 8;                   // key tiles of S
    constexpr int HP = 32;                      // padded head width in qkv
    extern __shared__ __align__(16) unsigned char mha_smem[];
    typedef __nv_bfloat16 Tile[HPC][T][P];
    Tile *tiles = reinterpret_cast<Tile *>(mha_smem);          // [2 sets][3 matrices]
    const int64_t news = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const __nv_bfloat16 *base = qkv + news * T * ld;
    const int npass = (nhead + HPC - 1) / HPC;
    auto issue = [&](int pass) {
        Tile *set = tiles + 3 * (pass & 1);
        for (int i = tid; i < HPC * 3 * T * 4; i += 256) {
            const int c = i & 3, r = (i >> 2) % T, m = (i / (4 * T)) % 3, gg = i / (4 * T * 3);
            const int head = pass * HPC + gg;
            if (head < nhead) cp_async_16(&set[m][gg][r][8 * c], base + r * ld + (m * nhead + head) * HP + 8 * c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int g = warp / WPH, r0 = 16 * (warp % WPH);
    issue(0);
    for (int pass = 0; pass < npass; ++pass) {
        if (pass + 1 < npass) {
            issue(pass + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        Tile *set = tiles + 3 * (pass & 1);
        const int head = pass * HPC + g;
        if (head < nhead) {
            // ---- S = Q K^T (16 x T per warp) ----
            uint32_t qa[2][4];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) ldsm_x4(qa[ks], &set[0][g][r0 + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * ks + 8 * (lane >> 4)]);
            float sacc[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
                uint32_t kb[4];
                ldsm_x4(kb, &set[1][g][8 * j + (lane & 7)][8 * (lane >> 3)]);
                mma_bf16_16816(sacc[j], qa[0], kb[0], kb[1]);
                mma_bf16_16816(sacc[j], qa[1], kb[2], kb[3]);
            }
            // ---- softmax over the keys: rows lane / 4 (c0, c1) and lane / 4 + 8 (c2, c3) ----
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                m0 = fmaxf(m0, fmaxf(sacc[j][0], sacc[j][1]));
                m1 = fmaxf(m1, fmaxf(sacc[j][2], sacc[j][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                sacc[j][0] = exp2f((sacc[j][0] - m0) * scale_log2e);
                sacc[j][1] = exp2f((sacc[j][1] - m0) * scale_log2e);
                sacc[j][2] = exp2f((sacc[j][2] - m1) * scale_log2e);
                sacc[j][3] = exp2f((sacc[j][3] - m1) * scale_log2e);
                l0 += sacc[j][0] + sacc[j][1];
                l1 += sacc[j][2] + sacc[j][3];
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            // ---- O = P V (16 x 32 per warp) ----
            float oacc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.0f;
#pragma unroll
            for (int kk = 0; kk < T / 16; ++kk) {
                uint32_t pa[4];
                pa[0] = pack2_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);
                pa[1] = pack2_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
                pa[2] = pack2_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
                pa[3] = pack2_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    uint32_t vb[4];
                    ldsm_x4_trans(vb, &set[2][g][16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * jp + 8 * (lane >> 4)]);
                    mma_bf16_16816(oacc[2 * jp], pa, vb[0], vb[1]);
                    mma_bf16_16816(oacc[2 * jp + 1], pa, vb[2], vb[3]);
                }
            }
            const float i0 = 1.0f / l0, i1 = 1.0f / l1;
            __nv_bfloat16 *o0 = ctx + (news * T + r0 + (lane >> 2)) * ldo + head * hd, *o1 = o0 + 8 * ldo;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 8 * j + 2 * (lane & 3);
                if (c < hd) {
                    *reinterpret_cast<uint32_t *>(o0 + c) = pack2_bf16(oacc[j][0] * i0, oacc[j][1] * i0);
                    *reinterpret_cast<uint32_t *>(o1 + c) = pack2_bf16(oacc[j][2] * i1, oacc[j][3] * i1);
                }
            }
        }
        __syncthreads();                        // the set is refilled two passes on
    }
    // K padding of the next GEMM's operand
    for (int64_t i = tid; i < (int64_t)T * (ldo - d); i += 256) {
        const int64_t r = i / (ldo - d), c = d + i % (ldo - d);
        ctx[(news * T + r) * ldo + c] = __float2bfloat16_rn(0.0f);
    }
}


// ---- backward of the attention core on the tensor cores (bf16 mode of the training step, trainer.py:131-146 through
// newsEncoders.py:244-247) ----------------------------------------------------------------------------------------------
// One CTA = one (news, head), T / 32 warps; fp32 q | k | v and dO head slices are rounded to bf16 tiles [T][40] on the way
// into shared memory.  Phase 1, warp w = query rows 32 w .. 32 w + 31 in two 16-row blocks: S = Q K^T and dP = dO V^T by
// mma.m16n8k16 with the whole key range in registers, softmax and rowdot = sum_j P dP on the accumulator fragments, the
// attention-weight dropout mask of the forward re-evaluated per element (O = (P * M) V: dP <- M * dP, dV uses P * M),
// dS = P (dP - rowdot) re-used in place as the A fragments of dQ = scale dS K; P * M and dS also go to shared memory as
// bf16 [T][T + 8].  Phase 2, warp w = key rows 32 w ..: dK = scale dS^T Q and dV = (P M)^T dO with the A fragments read
// through ldmatrix.trans.  Replaces mha_bwd_kernel<T> (FFMA, 7.5 ms of a 25 ms training step) in bf16 mode.
template <int T>
__global__ void __launch_bounds__(T)
mha_bwd_tc_kernel(const float *__restrict__ qkv, const float *__restrict__ dctx, float *__restrict__ dqkv, int d, int nhead,
                  int hd, float scale, float p_drop, uint64_t seed, int64_t news0) {
    constexpr int P = 40;                       // pitch of the Q K V dO tiles (bf16 elements): 80 bytes, ldmatrix conflict-free
    constexpr int PS = T + 8;                   // pitch of the P / dS tiles: (T + 8) * 2 bytes = an odd number of 16-byte chunks
    constexpr int NT = T / 8;
    extern __shared__ __align__(16) unsigned char mhab_smem[];
    typedef __nv_bfloat16 Row[P];
    typedef __nv_bfloat16 RowS[PS];
    Row *Qs = reinterpret_cast<Row *>(mhab_smem), *Ks = Qs + T, *Vs = Ks + T, *Os = Vs + T;
    RowS *Ps = reinterpret_cast<RowS *>(Os + T), *Ds = Ps + T;
    const int64_t news = blockIdx.y;
    const int head = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t2 = 2 * (lane & 3);
    const int64_t ld = 3 * (int64_t)d;
    const float *base = qkv + news * T * ld;
    // T * 16 column pairs per matrix, 16 per thread, loaded 8 at a time (32 x 8 bytes in flight per thread) before the first
    // conversion: the prologue's load latency is what bounds these one-(news, head)-per-CTA kernels (ncu on mha_x3_kernel)
    const bool vec2 = (hd & 1) == 0 && (d & 1) == 0 && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dctx)) & 7) == 0;
    const float *dbase = dctx + news * T * (int64_t)d;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float2 x[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + (8 * b + i) * T;
            const int r = idx >> 4, e = 2 * (idx & 15);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float *src = m < 3 ? base + r * ld + m * d + head * hd + e : dbase + r * (int64_t)d + head * hd + e;
                if (vec2 && e + 1 < hd) x[i][m] = __ldg(reinterpret_cast<const float2 *>(src));
                else x[i][m] = make_float2(e < hd ? __ldg(src) : 0.0f, e + 1 < hd ? __ldg(src + 1) : 0.0f);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + (8 * b + i) * T;
            const int r = idx >> 4, e = 2 * (idx & 15);
            *reinterpret_cast<uint32_t *>(&Qs[r][e]) = pack2_bf16(x[i][0].x, x[i][0].y);
            *reinterpret_cast<uint32_t *>(&Ks[r][e]) = pack2_bf16(x[i][1].x, x[i][1].y);
            *reinterpret_cast<uint32_t *>(&Vs[r][e]) = pack2_bf16(x[i][2].x, x[i][2].y);
            *reinterpret_cast<uint32_t *>(&Os[r][e]) = pack2_bf16(x[i][3].x, x[i][3].y);
        }
    }
    __syncthreads();
    const uint64_t drop_nh = ((uint64_t)(news0 + news) * nhead + head) * T;
    const float sl2 = scale * 1.4426950408889634f;
    float *out = dqkv + news * T * ld + head * hd;

    // ---------------- phase 1: query rows ----------------
#pragma unroll 1
    for (int mb = 0; mb < 2; ++mb) {
        const int r0 = 32 * warp + 16 * mb;
        uint32_t qa[2][4], oa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            ldsm_x4(qa[ks], &Qs[r0 + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * ks + 8 * (lane >> 4)]);
            ldsm_x4(oa[ks], &Os[r0 + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * ks + 8 * (lane >> 4)]);
        }
        float sacc[NT][4], pacc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
            pacc[j][0] = pacc[j][1] = pacc[j][2] = pacc[j][3] = 0.0f;
            uint32_t kb[4], vb[4];
            ldsm_x4(kb, &Ks[8 * j + (lane & 7)][8 * (lane >> 3)]);
            ldsm_x4(vb, &Vs[8 * j + (lane & 7)][8 * (lane >> 3)]);
            mma_bf16_16816(sacc[j], qa[0], kb[0], kb[1]);
            mma_bf16_16816(sacc[j], qa[1], kb[2], kb[3]);
            mma_bf16_16816(pacc[j], oa[0], vb[0], vb[1]);
            mma_bf16_16816(pacc[j], oa[1], vb[2], vb[3]);
        }
        // softmax over the keys: rows g (c0, c1) and g + 8 (c2, c3)
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            m0 = fmaxf(m0, fmaxf(sacc[j][0], sacc[j][1]));
            m1 = fmaxf(m1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            sacc[j][0] = exp2f((sacc[j][0] - m0) * sl2);
            sacc[j][1] = exp2f((sacc[j][1] - m0) * sl2);
            sacc[j][2] = exp2f((sacc[j][2] - m1) * sl2);
            sacc[j][3] = exp2f((sacc[j][3] - m1) * sl2);
            l0 += sacc[j][0] + sacc[j][1];
            l1 += sacc[j][2] + sacc[j][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const int ra = r0 + g, rb = ra + 8;
        float rd0 = 0.0f, rd1 = 0.0f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c = 8 * j + t2;
            float ms[4] = {1.0f, 1.0f, 1.0f, 1.0f};
            if (p_drop > 0.0f) {
                ms[0] = drop_scale(seed, (drop_nh + ra) * T + c, p_drop);
                ms[1] = drop_scale(seed, (drop_nh + ra) * T + c + 1, p_drop);
                ms[2] = drop_scale(seed, (drop_nh + rb) * T + c, p_drop);
                ms[3] = drop_scale(seed, (drop_nh + rb) * T + c + 1, p_drop);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sacc[j][e] *= e < 2 ? i0 : i1;           // P
                pacc[j][e] *= ms[e];                     // M * dP
            }
            rd0 = fmaf(sacc[j][0], pacc[j][0], fmaf(sacc[j][1], pacc[j][1], rd0));
            rd1 = fmaf(sacc[j][2], pacc[j][2], fmaf(sacc[j][3], pacc[j][3], rd1));
            *reinterpret_cast<uint32_t *>(&Ps[ra][c]) = pack2_bf16(sacc[j][0] * ms[0], sacc[j][1] * ms[1]);
            *reinterpret_cast<uint32_t *>(&Ps[rb][c]) = pack2_bf16(sacc[j][2] * ms[2], sacc[j][3] * ms[3]);
        }
        rd0 += __shfl_xor_sync(0xffffffffu, rd0, 1);
        rd0 += __shfl_xor_sync(0xffffffffu, rd0, 2);
        rd1 += __shfl_xor_sync(0xffffffffu, rd1, 1);
        rd1 += __shfl_xor_sync(0xffffffffu, rd1, 2);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c = 8 * j + t2;
            sacc[j][0] *= pacc[j][0] - rd0;              // dS = P (M dP - rowdot)
            sacc[j][1] *= pacc[j][1] - rd0;
            sacc[j][2] *= pacc[j][2] - rd1;
            sacc[j][3] *= pacc[j][3] - rd1;
            *reinterpret_cast<uint32_t *>(&Ds[ra][c]) = pack2_bf16(sacc[j][0], sacc[j][1]);
            *reinterpret_cast<uint32_t *>(&Ds[rb][c]) = pack2_bf16(sacc[j][2], sacc[j][3]);
        }
        // dQ = scale dS K  (16 x 32 per block): dS accumulator fragments as A, K through ldmatrix.trans
        float qacc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) qacc[j][0] = qacc[j][1] = qacc[j][2] = qacc[j][3] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < T / 16; ++kk) {
            uint32_t da[4];
            da[0] = pack2_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);
            da[1] = pack2_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
            da[2] = pack2_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
            da[3] = pack2_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                uint32_t kb[4];
                ldsm_x4_trans(kb, &Ks[16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * jp + 8 * (lane >> 4)]);
                mma_bf16_16816(qacc[2 * jp], da, kb[0], kb[1]);
                mma_bf16_16816(qacc[2 * jp + 1], da, kb[2], kb[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 8 * j + t2;
            if (c < hd) {
                out[ra * ld + c] = qacc[j][0] * scale;
                out[rb * ld + c] = qacc[j][2] * scale;
            }
            if (c + 1 < hd) {
                out[ra * ld + c + 1] = qacc[j][1] * scale;
                out[rb * ld + c + 1] = qacc[j][3] * scale;
            }
        }
    }
    __syncthreads();

    // ---------------- phase 2: key rows ----------------
#pragma unroll 1
    for (int mb = 0; mb < 2; ++mb) {
        const int j0 = 32 * warp + 16 * mb;
        float kacc[4][4], vacc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            kacc[j][0] = kacc[j][1] = kacc[j][2] = kacc[j][3] = 0.0f;
            vacc[j][0] = vacc[j][1] = vacc[j][2] = vacc[j][3] = 0.0f;
        }
#pragma unroll 2
        for (int kk = 0; kk < T / 16; ++kk) {
            // A (m = key j, k = query i) = the transpose of the stored [i][j] tiles
            uint32_t dsa[4], pma[4];
            ldsm_x4_trans(dsa, &Ds[16 * kk + (lane & 7) + 8 * (lane >> 4)][j0 + 8 * ((lane >> 3) & 1)]);
            ldsm_x4_trans(pma, &Ps[16 * kk + (lane & 7) + 8 * (lane >> 4)][j0 + 8 * ((lane >> 3) & 1)]);
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                uint32_t qb[4], ob[4];
                ldsm_x4_trans(qb, &Qs[16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * jp + 8 * (lane >> 4)]);
                ldsm_x4_trans(ob, &Os[16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * jp + 8 * (lane >> 4)]);
                mma_bf16_16816(kacc[2 * jp], dsa, qb[0], qb[1]);
                mma_bf16_16816(kacc[2 * jp + 1], dsa, qb[2], qb[3]);
                mma_bf16_16816(vacc[2 * jp], pma, ob[0], ob[1]);
                mma_bf16_16816(vacc[2 * jp + 1], pma, ob[2], ob[3]);
            }
        }
        const int ra = j0 + g, rb = ra + 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 8 * j + t2;
            if (c < hd) {
                out[ra * ld + d + c] = kacc[j][0] * scale;
                out[rb * ld + d + c] = kacc[j][2] * scale;
                out[ra * ld + 2 * d + c] = vacc[j][0];
                out[rb * ld + 2 * d + c] = vacc[j][2];
            }
            if (c + 1 < hd) {
                out[ra * ld + d + c + 1] = kacc[j][1] * scale;
                out[rb * ld + d + c + 1] = kacc[j][3] * scale;
                out[ra * ld + 2 * d + c + 1] = vacc[j][1];
                out[rb * ld + 2 * d + c + 1] = vacc[j][3];
            }
        }
    }
}


// Forward of the same core for the TRAINING layout (fp32 q | k | v rows of width 3 d, fp32 ctx, attention-weight dropout):
// one CTA per (news, head), warp w = query rows 32 w .., S = Q K^T, softmax, P * M, O = (P M) V on mma.m16n8k16.
// bf16 mode of lime_mha (mha_kernel<T> stays the fp32 parity path).
template <int T>
__global__ void __launch_bounds__(T)
mha_fwd_tc_kernel(const float *__restrict__ qkv, float *__restrict__ ctx, int d, int nhead, int hd, float scale, float p_drop,
                  uint64_t seed, int64_t news0) {
    constexpr int P = 40;
    constexpr int NT = T / 8;
    __shared__ __align__(16) __nv_bfloat16 Qs[T][P], Ks[T][P], Vs[T][P];
    const int64_t news = blockIdx.y;
    const int head = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t2 = 2 * (lane & 3);
    const int64_t ld = 3 * (int64_t)d;
    const float *base = qkv + news * T * ld;
    // T * 16 column pairs per matrix, 16 per thread, loaded 8 at a time before the first conversion (see mha_bwd_tc_kernel)
    const bool vec2 = (hd & 1) == 0 && (d & 1) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 7) == 0;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float2 x[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + (8 * b + i) * T;
            const int r = idx >> 4, e = 2 * (idx & 15);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const float *src = base + r * ld + m * d + head * hd + e;
                if (vec2 && e + 1 < hd) x[i][m] = __ldg(reinterpret_cast<const float2 *>(src));
                else x[i][m] = make_float2(e < hd ? __ldg(src) : 0.0f, e + 1 < hd ? __ldg(src + 1) : 0.0f);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + (8 * b + i) * T;
            const int r = idx >> 4, e = 2 * (idx & 15);
            *reinterpret_cast<uint32_t *>(&Qs[r][e]) = pack2_bf16(x[i][0].x, x[i][0].y);
            *reinterpret_cast<uint32_t *>(&Ks[r][e]) = pack2_bf16(x[i][1].x, x[i][1].y);
            *reinterpret_cast<uint32_t *>(&Vs[r][e]) = pack2_bf16(x[i][2].x, x[i][2].y);
        }
    }
    __syncthreads();
    const uint64_t drop_nh = ((uint64_t)(news0 + news) * nhead + head) * T;
    const float sl2 = scale * 1.4426950408889634f;
#pragma unroll 1
    for (int mb = 0; mb < 2; ++mb) {
        const int r0 = 32 * warp + 16 * mb;
        uint32_t qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ldsm_x4(qa[ks], &Qs[r0 + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * ks + 8 * (lane >> 4)]);
        float sacc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
            uint32_t kb[4];
            ldsm_x4(kb, &Ks[8 * j + (lane & 7)][8 * (lane >> 3)]);
            mma_bf16_16816(sacc[j], qa[0], kb[0], kb[1]);
            mma_bf16_16816(sacc[j], qa[1], kb[2], kb[3]);
        }
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            m0 = fmaxf(m0, fmaxf(sacc[j][0], sacc[j][1]));
            m1 = fmaxf(m1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.0f, l1 = 0.0f;
        const int ra = r0 + g, rb = ra + 8;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            sacc[j][0] = exp2f((sacc[j][0] - m0) * sl2);
            sacc[j][1] = exp2f((sacc[j][1] - m0) * sl2);
            sacc[j][2] = exp2f((sacc[j][2] - m1) * sl2);
            sacc[j][3] = exp2f((sacc[j][3] - m1) * sl2);
            l0 += sacc[j][0] + sacc[j][1];
            l1 += sacc[j][2] + sacc[j][3];
            if (p_drop > 0.0f) {                             // the sums are taken BEFORE the mask: O = (softmax * M) V
                const int c = 8 * j + t2;
                sacc[j][0] *= drop_scale(seed, (drop_nh + ra) * T + c, p_drop);
                sacc[j][1] *= drop_scale(seed, (drop_nh + ra) * T + c + 1, p_drop);
                sacc[j][2] *= drop_scale(seed, (drop_nh + rb) * T + c, p_drop);
                sacc[j][3] *= drop_scale(seed, (drop_nh + rb) * T + c + 1, p_drop);
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float oacc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < T / 16; ++kk) {
            uint32_t pa[4];
            pa[0] = pack2_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);
            pa[1] = pack2_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
            pa[2] = pack2_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
            pa[3] = pack2_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                uint32_t vb[4];
                ldsm_x4_trans(vb, &Vs[16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * jp + 8 * (lane >> 4)]);
                mma_bf16_16816(oacc[2 * jp], pa, vb[0], vb[1]);
                mma_bf16_16816(oacc[2 * jp + 1], pa, vb[2], vb[3]);
            }
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        float *o0 = ctx + (news * T + ra) * (int64_t)d + head * hd, *o1 = o0 + 8 * (int64_t)d;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 8 * j + t2;
            if (c < hd) {
                o0[c] = oacc[j][0] * i0;
                o1[c] = oacc[j][2] * i1;
            }
            if (c + 1 < hd) {
                o0[c + 1] = oacc[j][1] * i0;
                o1[c + 1] = oacc[j][3] * i1;
            }
        }
    }
}

// fp32x3 mode of lime_mha (Stage A, fp32 q | k | v rows and fp32 ctx as mha_kernel<T>): the same fragment layout as
// mha_fwd_tc_kernel, every operand an fp16 hi / lo pair (x = hi + lo, 22 significant bits; activations of a LayerNormed
// transformer sit well inside the fp16 range) and every product three MMAs: hi*hi + lo*hi + hi*lo, fp32 accumulation --
// the dropped lo*lo term is 2^-22 of the product.  q is scaled by 1/sqrt(hd) in fp32 BEFORE the split, like mha_kernel.
// The softmax weights p in [0, 1] are split the same way for O = P V.  Replaces the FFMA mha_kernel<T> in that mode
// (22 % of its cache build).  Shared memory: four [T][40] fp16 tiles (K, V hi / lo: 40,960 B at T = 128) and 128 registers, so
// that FOUR CTAs of 4 warps share an SM (the kernel is latency bound: ldmatrix -> mma chains at 12 warps per SM before).
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

template <int T>
__global__ void __launch_bounds__(T, T == 128 ? 4 : 1)
mha_x3_kernel(const float *__restrict__ qkv, float *__restrict__ ctx, __half *__restrict__ ctx_hi, __half *__restrict__ ctx_lo,
              int ld16, float out_scale, int d, int nhead, int hd, float scale, float p_drop, uint64_t seed, int64_t news0) {
    constexpr int P = 40;
    constexpr int NT = T / 8;
    typedef __half Tile[T][P];
    extern __shared__ __align__(16) unsigned char mhax_smem[];
    Tile *tl = reinterpret_cast<Tile *>(mhax_smem);             // Khi Klo Vhi Vlo (the query fragments come straight from global memory)
    const int64_t news = blockIdx.y;
    const int head = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t2 = 2 * (lane & 3);
    const int64_t ld = 3 * (int64_t)d;
    const float *base = qkv + news * T * ld;
    // a head slice of a row is hd contiguous floats; with hd even and d * 4 a multiple of 8 every pair is an aligned float2
    const bool vec2 = (hd & 1) == 0 && (d & 1) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 7) == 0;
    // T * 16 column pairs per matrix, 16 per thread: the loads of 8 pairs (K and V: 16 x 8 bytes in flight per thread) are issued
    // before the first split / store -- the kernel is bound by the latency of this prologue (ncu: 57 % of the stall samples on its loads)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float2 x[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + (8 * b + i) * T;
            const int r = idx >> 4, e = 2 * (idx & 15);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float *src = base + r * ld + (m + 1) * d + head * hd + e;
                if (vec2 && e + 1 < hd) x[i][m] = __ldg(reinterpret_cast<const float2 *>(src));
                else x[i][m] = make_float2(e < hd ? __ldg(src) : 0.0f, e + 1 < hd ? __ldg(src + 1) : 0.0f);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + (8 * b + i) * T;
            const int r = idx >> 4, e = 2 * (idx & 15);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                uint32_t hi, lo;
                split2_f16(x[i][m].x, x[i][m].y, hi, lo);
                *reinterpret_cast<uint32_t *>(&tl[2 * m][r][e]) = hi;
                *reinterpret_cast<uint32_t *>(&tl[2 * m + 1][r][e]) = lo;
            }
        }
    }
    __syncthreads();
    const uint64_t drop_nh = ((uint64_t)(news0 + news) * nhead + head) * T;
    constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll 1
    for (int mb = 0; mb < 2; ++mb) {
        const int r0 = 32 * warp + 16 * mb;
        // A fragments of m16n8k16: a0 = (row g, cols t2..), a1 = (row g + 8, same), a2 / a3 = the same rows, cols + 8
        uint32_t qh[2][4], ql[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int row = r0 + g + 8 * (f & 1), col = 16 * ks + t2 + 8 * (f >> 1);
                const float *src = base + row * ld + head * hd + col;
                float2 q;
                if (vec2 && col + 1 < hd) q = __ldg(reinterpret_cast<const float2 *>(src));
                else q = make_float2(col < hd ? __ldg(src) : 0.0f, col + 1 < hd ? __ldg(src + 1) : 0.0f);
                split2_f16(q.x * scale, q.y * scale, qh[ks][f], ql[ks][f]);
            }
        }
        float sacc[NT][4];
        uint32_t kh[2][4], kl[2][4];                                       // the key fragments one tile ahead of their MMAs
        ldsm_x4(kh[0], &tl[0][lane & 7][8 * (lane >> 3)]);
        ldsm_x4(kl[0], &tl[1][lane & 7][8 * (lane >> 3)]);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
            if (j + 1 < NT) {
                ldsm_x4(kh[(j + 1) & 1], &tl[0][8 * (j + 1) + (lane & 7)][8 * (lane >> 3)]);
                ldsm_x4(kl[(j + 1) & 1], &tl[1][8 * (j + 1) + (lane & 7)][8 * (lane >> 3)]);
            }
            const uint32_t(&a)[4] = kh[j & 1], (&b)[4] = kl[j & 1];
            mma_f16_16816(sacc[j], ql[0], a[0], a[1]);                     // the small terms first
            mma_f16_16816(sacc[j], ql[1], a[2], a[3]);
            mma_f16_16816(sacc[j], qh[0], b[0], b[1]);
            mma_f16_16816(sacc[j], qh[1], b[2], b[3]);
            mma_f16_16816(sacc[j], qh[0], a[0], a[1]);
            mma_f16_16816(sacc[j], qh[1], a[2], a[3]);
        }
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            m0 = fmaxf(m0, fmaxf(sacc[j][0], sacc[j][1]));
            m1 = fmaxf(m1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.0f, l1 = 0.0f;
        const int ra = r0 + g, rb = ra + 8;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            sacc[j][0] = exp2f((sacc[j][0] - m0) * kLog2e);
            sacc[j][1] = exp2f((sacc[j][1] - m0) * kLog2e);
            sacc[j][2] = exp2f((sacc[j][2] - m1) * kLog2e);
            sacc[j][3] = exp2f((sacc[j][3] - m1) * kLog2e);
            l0 += sacc[j][0] + sacc[j][1];
            l1 += sacc[j][2] + sacc[j][3];
            if (p_drop > 0.0f) {                             // the sums are taken BEFORE the mask: O = (softmax * M) V
                const int c = 8 * j + t2;
                sacc[j][0] *= drop_scale(seed, (drop_nh + ra) * T + c, p_drop);
                sacc[j][1] *= drop_scale(seed, (drop_nh + ra) * T + c + 1, p_drop);
                sacc[j][2] *= drop_scale(seed, (drop_nh + rb) * T + c, p_drop);
                sacc[j][3] *= drop_scale(seed, (drop_nh + rb) * T + c + 1, p_drop);
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float oacc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.0f;
        uint32_t vh[2][4], vl[2][4];                                       // the value fragments one step (kk, jp) ahead
        ldsm_x4_trans(vh[0], &tl[2][(lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]);
        ldsm_x4_trans(vl[0], &tl[3][(lane & 7) + 8 * ((lane >> 3) & 1)][8 * (lane >> 4)]);
#pragma unroll
        for (int kk = 0; kk < T / 16; ++kk) {
            uint32_t ph[4], pl[4];
            split2_f16(sacc[2 * kk][0], sacc[2 * kk][1], ph[0], pl[0]);
            split2_f16(sacc[2 * kk][2], sacc[2 * kk][3], ph[1], pl[1]);
            split2_f16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1], ph[2], pl[2]);
            split2_f16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                const int st = 2 * kk + jp;                                // step; the next one is (kk, 1) or (kk + 1, 0)
                if (st + 1 < T / 8) {
                    const int nk = (st + 1) >> 1, nj = (st + 1) & 1;
                    ldsm_x4_trans(vh[(st + 1) & 1], &tl[2][16 * nk + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * nj + 8 * (lane >> 4)]);
                    ldsm_x4_trans(vl[(st + 1) & 1], &tl[3][16 * nk + (lane & 7) + 8 * ((lane >> 3) & 1)][16 * nj + 8 * (lane >> 4)]);
                }
                const uint32_t(&a)[4] = vh[st & 1], (&b)[4] = vl[st & 1];
                mma_f16_16816(oacc[2 * jp], pl, a[0], a[1]);
                mma_f16_16816(oacc[2 * jp + 1], pl, a[2], a[3]);
                mma_f16_16816(oacc[2 * jp], ph, b[0], b[1]);
                mma_f16_16816(oacc[2 * jp + 1], ph, b[2], b[3]);
                mma_f16_16816(oacc[2 * jp], ph, a[0], a[1]);
                mma_f16_16816(oacc[2 * jp + 1], ph, a[2], a[3]);
            }
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        if (ctx_hi != nullptr) {
            // the context leaves as the next GEMM's operand pair: out_scale * ctx = hi + lo (fp16), no fp32 copy, no split pass
            const int64_t p0 = (news * T + ra) * (int64_t)ld16 + head * hd, p1 = p0 + 8 * (int64_t)ld16;
            const float s0 = i0 * out_scale, s1 = i1 * out_scale;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 8 * j + t2;
                if (c + 1 < hd && ((ld16 | hd) & 1) == 0) {
                    uint32_t hi, lo;
                    split2_f16(oacc[j][0] * s0, oacc[j][1] * s0, hi, lo);
                    *reinterpret_cast<uint32_t *>(ctx_hi + p0 + c) = hi;
                    *reinterpret_cast<uint32_t *>(ctx_lo + p0 + c) = lo;
                    split2_f16(oacc[j][2] * s1, oacc[j][3] * s1, hi, lo);
                    *reinterpret_cast<uint32_t *>(ctx_hi + p1 + c) = hi;
                    *reinterpret_cast<uint32_t *>(ctx_lo + p1 + c) = lo;
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (c + e < hd) {
                            const float xa = oacc[j][e] * s0, xb = oacc[j][2 + e] * s1;
                            const __half ha = __float2half_rn(xa), hb = __float2half_rn(xb);
                            ctx_hi[p0 + c + e] = ha;
                            ctx_lo[p0 + c + e] = __float2half_rn(xa - __half2float(ha));
                            ctx_hi[p1 + c + e] = hb;
                            ctx_lo[p1 + c + e] = __float2half_rn(xb - __half2float(hb));
                        }
                    }
                }
            }
            if (head == nhead - 1) {                             // zero K padding of the operand (columns d .. ld16 - 1)
                for (int c = d + (lane & 3); c < ld16; c += 4) {
                    const int64_t q0 = (news * T + ra) * (int64_t)ld16 + c, q1 = q0 + 8 * (int64_t)ld16;
                    ctx_hi[q0] = ctx_lo[q0] = ctx_hi[q1] = ctx_lo[q1] = __float2half_rn(0.0f);
                }
            }
            continue;
        }
        float *o0 = ctx + (news * T + ra) * (int64_t)d + head * hd, *o1 = o0 + 8 * (int64_t)d;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 8 * j + t2;
            if (c + 1 < hd && vec2 && (reinterpret_cast<uintptr_t>(ctx) & 7) == 0) {          // one 8-byte store per row
                *reinterpret_cast<float2 *>(o0 + c) = make_float2(oacc[j][0] * i0, oacc[j][1] * i0);
                *reinterpret_cast<float2 *>(o1 + c) = make_float2(oacc[j][2] * i1, oacc[j][3] * i1);
                continue;
            }
            if (c < hd) {
                o0[c] = oacc[j][0] * i0;
                o1[c] = oacc[j][2] * i1;
            }
            if (c + 1 < hd) {
                o0[c + 1] = oacc[j][1] * i0;
                o1[c + 1] = oacc[j][3] * i1;
            }
        }
    }
}

// ---- LayerNorm (two-pass, like ATen) ------------------------------------------------------------
constexpr int kLnMaxPerLane = 16;   // d <= 512

__device__ __forceinline__ void ln_row(const float *__restrict__ x, int d, int lane, float eps,
                                       const float *__restrict__ gamma, const float *__restrict__ beta,
                                       float (&y)[kLnMaxPerLane]) {
    float v[kLnMaxPerLane];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        v[i] = (c < d) ? x[c] : 0.0f;
        s += v[i];
    }
    const float mu = warp_sum(s) / (float)d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        const float t = (c < d) ? (v[i] - mu) : 0.0f;
        q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        y[i] = (c < d) ? fmaf((v[i] - mu) * rstd, gamma[c], beta[c]) : 0.0f;
    }
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ gamma,
                 const float *__restrict__ beta, float *__restrict__ y, int64_t ldy, int64_t rows,
                 int d, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    float o[kLnMaxPerLane];
    ln_row(x + r * ldx, d, lane, eps, gamma, beta, o);
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < d) y[r * ldy + c] = o[i];
    }
}

// the same with a second, bf16 image of the row (the A operand of the TMA GEMM): [rows, ld16], columns d.. zero
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ gamma, const float *__restrict__ beta,
                      float *__restrict__ y, int64_t ldy, __nv_bfloat16 *__restrict__ y16, int ld16, int64_t rows, int d, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    float o[kLnMaxPerLane];
    ln_row(x + r * ldx, d, lane, eps, gamma, beta, o);
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < d) y[r * ldy + c] = o[i];
        if (c < ld16) y16[r * ld16 + c] = __float2bfloat16_rn(c < d ? o[i] : 0.0f);
    }
}

// ... and with the fp16 operand pair of the fp32x3 mode: scale * y = hi + lo, each [rows, ld16], columns d.. zero
__global__ void __launch_bounds__(256)
layernorm_pairs_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ gamma, const float *__restrict__ beta,
                       float *__restrict__ y, int64_t ldy, __half *__restrict__ hi16, __half *__restrict__ lo16, int ld16, float scale,
                       int64_t rows, int d, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    float o[kLnMaxPerLane];
    ln_row(x + r * ldx, d, lane, eps, gamma, beta, o);
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < d) y[r * ldy + c] = o[i];
        if (c < ld16) {
            const float xs = c < d ? o[i] * scale : 0.0f;
            const __half h = __float2half_rn(xs);
            hi16[r * ld16 + c] = h;
            lo16[r * ld16 + c] = __float2half_rn(xs - __half2float(h));
        }
    }
}

// LayerNorm of every token + unmasked mean over the T tokens (newsEncoders.py:317,321)
__global__ void __launch_bounds__(256)
layernorm_meanpool_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                          const float *__restrict__ beta, float *__restrict__ out, int64_t ldo, int T,
                          int d, float eps) {
    __shared__ float part[8][kLnMaxPerLane * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t news = blockIdx.x;
    float acc[kLnMaxPerLane];
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) acc[i] = 0.0f;
    for (int t = warp; t < T; t += 8) {
        float o[kLnMaxPerLane];
        ln_row(x + (news * T + t) * (int64_t)d, d, lane, eps, gamma, beta, o);
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i) acc[i] += o[i];
    }
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) part[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w][c];
        out[news * ldo + c] = s / (float)T;
    }
}

// ---- topic representation: category_affine(cat_emb[c] || sub_emb[s]) --------------------------
__global__ void __launch_bounds__(64)
topic_rep_kernel(const float *__restrict__ cat_emb, const float *__restrict__ sub_emb,
                 const float *__restrict__ W, const float *__restrict__ b,
                 const int32_t *__restrict__ cat, const int32_t *__restrict__ sub, int64_t n,
                 float *__restrict__ out, int64_t ldo, int width) {
    __shared__ float in[2 * LIME_TOPIC];
    const int64_t i = blockIdx.x;
    const int t = threadIdx.x;
    if (t < LIME_TOPIC) {
        in[t] = cat_emb[(int64_t)cat[i] * LIME_TOPIC + t];
        in[LIME_TOPIC + t] = sub_emb[(int64_t)sub[i] * LIME_TOPIC + t];
    }
    __syncthreads();
    if (t < LIME_TOPIC) {
        float a = b[t];
        const float *w = W + t * 2 * LIME_TOPIC;
#pragma unroll 10
        for (int k = 0; k < 2 * LIME_TOPIC; ++k) a = fmaf(w[k], in[k], a);
        out[i * ldo + t] = a;
    } else if (t < width) {
        out[i * ldo + t] = 0.0f;
    }
}

// ---- layers.Attention over the k intents (layers.py:285-300) -----------------------------------
__global__ void __launch_bounds__(128)
intent_pool_kernel(const float *__restrict__ pre, const float *__restrict__ e,
                   const float *__restrict__ w2, float *__restrict__ out, int64_t ldo, int64_t n, int k,
                   int D) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    float sc[8];
    float m = -INFINITY;
    for (int kk = 0; kk < k; ++kk) {
        const float *p = pre + (i * k + kk) * (int64_t)D;
        float a = 0.0f;
        for (int c = lane; c < D; c += 32) a = fmaf(w2[c], tanhf(p[c]), a);
        sc[kk] = warp_sum(a);
        m = fmaxf(m, sc[kk]);
    }
    float l = 0.0f;
    for (int kk = 0; kk < k; ++kk) {
        sc[kk] = expf(sc[kk] - m);
        l += sc[kk];
    }
    for (int c = lane; c < D; c += 32) {
        float a = 0.0f;
        for (int kk = 0; kk < k; ++kk) a = fmaf(sc[kk] / l, e[(i * k + kk) * (int64_t)D + c], a);
        out[i * ldo + c] = a;
    }
}

// ---- cosine gate + concat + feature fusion (newsEncoders.py:297-300, 367-371, 221-225) ----------
__global__ void __launch_bounds__(128)
content_fuse_kernel(const float *__restrict__ title, const float *__restrict__ body,
                    const float *__restrict__ cat_emb, const float *__restrict__ sub_emb,
                    const int32_t *__restrict__ cat, const int32_t *__restrict__ sub, int64_t n, int D,
                    int cat_dim, int sub_dim, float *__restrict__ content, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float *t = title + i * (int64_t)D;
    const float *b = body + i * (int64_t)D;
    float tb = 0.f, tt = 0.f, bb = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float x = t[c], y = b[c];
        tb = fmaf(x, y, tb);
        tt = fmaf(x, x, tt);
        bb = fmaf(y, y, bb);
    }
    tb = warp_sum(tb);
    tt = warp_sum(tt);
    bb = warp_sum(bb);
    const float eps = 1e-8f;   // F.cosine_similarity default
    const float cosv = tb / (fmaxf(sqrtf(tt), eps) * fmaxf(sqrtf(bb), eps));
    const float sim = (cosv + 1.0f) / 2.0f;
    float *o = content + i * ldo;
    for (int c = lane; c < D; c += 32) {
        o[c] = t[c];
        o[D + c] = sim * b[c];
    }
    const float *ce = cat_emb + (int64_t)cat[i] * cat_dim;
    const float *se = sub_emb + (int64_t)sub[i] * sub_dim;
    for (int c = lane; c < cat_dim; c += 32) o[2 * D + c] = ce[c];
    for (int c = lane; c < sub_dim; c += 32) o[2 * D + cat_dim + c] = se[c];
}

__global__ void bucket_pairs_kernel(const float *__restrict__ Ef, const float *__restrict__ El, int nb,
                                    int dim, float *__restrict__ out) {
    const int pair = blockIdx.x;
    const int bf = pair / nb, bl = pair - bf * nb;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        out[(int64_t)pair * 2 * dim + c] = Ef[(int64_t)bf * dim + c];
        out[(int64_t)pair * 2 * dim + dim + c] = El[(int64_t)bl * dim + c];
    }
}

__global__ void scale_rows_kernel(float *M, int64_t ld, const float *__restrict__ rs, float alpha, int rows,
                                  int cols) {
    const int r = blockIdx.x;
    const float s = alpha * (rs ? rs[r] : 1.0f);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) M[(int64_t)r * ld + c] *= s;
}

// one warp per row: out[row] = max |M[row][:]|
__global__ void row_absmax_kernel(const float *__restrict__ M, int64_t ld, int64_t rows, int cols, float *__restrict__ out,
                                  int64_t ldo) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float m = 0.0f;
    for (int j = threadIdx.x & 31; j < cols; j += 32) m = fmaxf(m, fabsf(M[row * ld + j]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) out[row * ldo] = m;
}

__global__ void prefix_rows_kernel(float *M, int64_t ld, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float run = 0.0f;
    for (int r = 0; r < rows; ++r) {
        run += M[(int64_t)r * ld + c];
        M[(int64_t)r * ld + c] = run;
    }
}

}  // namespace lime

using namespace lime;

extern "C" int lime_bucketize(const float *seconds, int64_t n, int num_buckets, int32_t *buckets,
                              void *stream) {
    LIME_CHECK_ARG(seconds && buckets && num_buckets >= 1, "lime_bucketize: bad argument");
    if (n <= 0) return 0;
    const float scale = (float)((double)num_buckets / 7.0);
    bucketize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(seconds, n, scale, num_buckets, buckets);
    LIME_LAUNCH_CHECK("bucketize_kernel");
    return 0;
}

extern "C" int lime_embed_pe(const float *E, int64_t vocab, const int32_t *ids, int64_t rows, int T,
                             int d, const float *pe, float *out, void *stream) {
    LIME_CHECK_ARG(E && ids && pe && out, "lime_embed_pe: null argument");
    LIME_CHECK_ARG((d & 3) == 0 && T > 0 && vocab > 0, "lime_embed_pe: d=%d must be a multiple of 4", d);
    if (rows <= 0) return 0;
    const int64_t total = rows * (d / 4);
    embed_pe_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(E, vocab, ids, rows, T, d / 4, pe, out);
    LIME_LAUNCH_CHECK("embed_pe_kernel");
    return 0;
}

extern "C" int lime_embed_pe_bf16(const float *E, int64_t vocab, const int32_t *ids, int64_t rows, int T, int d, const float *pe,
                                  float *out, void *out16, int32_t ld16, void *stream) {
    LIME_CHECK_ARG(E && ids && pe && out && out16, "lime_embed_pe_bf16: null argument");
    LIME_CHECK_ARG((d & 3) == 0 && T > 0 && vocab > 0 && ld16 >= d && (ld16 & 7) == 0, "lime_embed_pe_bf16: d=%d ld16=%d", d, ld16);
    if (rows <= 0) return 0;
    const int64_t total = rows * (ld16 / 4);
    embed_pe_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(E, vocab, ids, rows, T, d / 4, pe, out,
                                                                                       reinterpret_cast<__nv_bfloat16 *>(out16), ld16);
    LIME_LAUNCH_CHECK("embed_pe_bf16_kernel");
    return 0;
}

extern "C" int lime_embed_pe_pairs(const float *E, int64_t vocab, const int32_t *ids, int64_t rows, int T, int d, const float *pe,
                                   float *out, void *hi16, void *lo16, int32_t ld16, float scale, void *stream) {
    LIME_CHECK_ARG(E && ids && pe && out && hi16 && lo16, "lime_embed_pe_pairs: null argument");
    LIME_CHECK_ARG((d & 3) == 0 && T > 0 && vocab > 0 && ld16 >= d && (ld16 & 7) == 0, "lime_embed_pe_pairs: d=%d ld16=%d", d, ld16);
    if (rows <= 0) return 0;
    const int64_t total = rows * (ld16 / 4);
    embed_pe_pairs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        E, vocab, ids, rows, T, d / 4, pe, out, reinterpret_cast<__half *>(hi16), reinterpret_cast<__half *>(lo16), ld16, scale);
    LIME_LAUNCH_CHECK("embed_pe_pairs_kernel");
    return 0;
}

// fp32 rows -> two 16-bit images hi = r16(s x), lo = r16(s x - hi) of [rows, ld16] (columns d.. zero): the operands of the
// three-pass dense layer (x . w ~ xh . wh + xl . wh + xh . wl) of the fp32-accurate tensor-core mode.  fp16 pairs carry
// 11 + 11 bits (2^-22 relative; s = a power of two that keeps the lo halves of typical values out of the subnormals), bf16
// pairs 8 + 8 bits (2^-17).
template <bool F16>
__global__ void __launch_bounds__(256)
split16_kernel(const float *__restrict__ x, int64_t ldx, int64_t rows, int d, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo,
               int ld16, float scale) {
    const int q4 = ld16 / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * q4) return;
    const int64_t r = idx / q4;
    const int c = 4 * (int)(idx - r * q4);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c + 4 <= d && ((reinterpret_cast<uintptr_t>(x + r * ldx + c) & 15) == 0)) {
        const float4 t = *reinterpret_cast<const float4 *>(x + r * ldx + c);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        for (int e = 0; e < 4; ++e)
            if (c + e < d) v[e] = x[r * ldx + c + e];
    }
    uint16_t h[4], l[4];
    for (int e = 0; e < 4; ++e) {
        const float xs = v[e] * scale;
        if (F16) {
            const __half hh = __float2half_rn(xs);
            const __half ll = __float2half_rn(xs - __half2float(hh));
            h[e] = *reinterpret_cast<const uint16_t *>(&hh);
            l[e] = *reinterpret_cast<const uint16_t *>(&ll);
        } else {
            const __nv_bfloat16 hh = __float2bfloat16_rn(xs);
            const __nv_bfloat16 ll = __float2bfloat16_rn(xs - __bfloat162float(hh));
            h[e] = *reinterpret_cast<const uint16_t *>(&hh);
            l[e] = *reinterpret_cast<const uint16_t *>(&ll);
        }
    }
    *reinterpret_cast<uint2 *>(hi + r * ld16 + c) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    if (lo != nullptr)       // lo == NULL: a plain rounded cast with zero padding (the bf16 training GEMMs)
        *reinterpret_cast<uint2 *>(lo + r * ld16 + c) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}


// fp32 rows -> bf16 image (zero padded to ld16 columns) AND the column sums of the fp32 rows in the same pass: the operand
// cast of dZ for dX / dW and the bias gradient db = sum_r dZ[r, :] of an nn.Linear backward (bf16 training mode) read dZ once.
// Block = 256 rows; thread = (column quad, row phase); colsum is ACCUMULATED (one 128-bit reduction per thread).
__global__ void __launch_bounds__(256)
cast_bf16_colsum_kernel(const float *__restrict__ x, int64_t ldx, int64_t rows, int d, uint16_t *__restrict__ out16, int ld16,
                        float *__restrict__ colsum) {
    const int q4 = ld16 >> 2;                          // column quads per row (<= 256)
    const int rpp = 256 / q4;                          // rows per pass
    const int c4 = threadIdx.x % q4, rsub = threadIdx.x / q4;
    if (rsub >= rpp) return;
    const int c = 4 * c4;
    const int64_t r0 = (int64_t)blockIdx.x * 256, r1 = r0 + 256 < rows ? r0 + 256 : rows;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + rsub; r < r1; r += rpp) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d) v = *reinterpret_cast<const float4 *>(x + r * ldx + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2 *>(out16 + r * ld16 + c) = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
    }
    if (c < d)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
}

extern "C" int lime_split_bf16_pairs(const float *x, int64_t ldx, int64_t rows, int32_t d, void *hi, void *lo, int32_t ld16, float scale,
                                     int32_t as_fp16, void *stream) {
    LIME_CHECK_ARG(x && hi, "lime_split_bf16_pairs: null argument");
    LIME_CHECK_ARG(d >= 1 && ld16 >= d && (ld16 & 7) == 0 && ldx >= d, "lime_split_bf16_pairs: d=%d ld16=%d ldx=%lld", d, ld16, (long long)ldx);
    if (rows <= 0) return 0;
    const int64_t total = rows * (ld16 / 4);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (as_fp16)
        split16_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, ldx, rows, d, reinterpret_cast<uint16_t *>(hi), reinterpret_cast<uint16_t *>(lo), ld16, scale);
    else
        split16_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, ldx, rows, d, reinterpret_cast<uint16_t *>(hi), reinterpret_cast<uint16_t *>(lo), ld16, scale);
    LIME_LAUNCH_CHECK("split16_kernel");
    return 0;
}

extern "C" int lime_mha(const float *qkv, float *ctx, int64_t n_news, int T, int d, int nhead, float p_drop, uint64_t seed,
                        int64_t news0, void *stream) {
    LIME_CHECK_ARG(qkv && ctx, "lime_mha: null argument");
    LIME_CHECK_ARG(nhead > 0 && d % nhead == 0 && d / nhead <= 32, "lime_mha: head dim %d unsupported (<= 32)", nhead ? d / nhead : -1);
    LIME_CHECK_ARG(T == 32 || T == 128, "lime_mha: T=%d unsupported (32 or 128)", T);
    LIME_CHECK_ARG(n_news <= 65535, "lime_mha: at most 65535 news per call (got %lld)", (long long)n_news);
    if (n_news <= 0) return 0;
    const int hd = d / nhead;
    const float scale = 1.0f / sqrtf((float)hd);
    if (T == 32) {
        dim3 grid((nhead + 3) / 4, (unsigned)n_news);
        mha_kernel<32, float, float><<<grid, 128, 0, as_stream(stream)>>>(qkv, ctx, 3 * (int64_t)d, d, d, nhead, hd, scale, p_drop, seed, news0);
    } else {
        dim3 grid(nhead, (unsigned)n_news);
        mha_kernel<128, float, float><<<grid, 128, 0, as_stream(stream)>>>(qkv, ctx, 3 * (int64_t)d, d, d, nhead, hd, scale, p_drop, seed, news0);
    }
    LIME_LAUNCH_CHECK("mha_kernel");
    return 0;
}

// fp32x3 mode of lime_mha: same arguments, fp16 hi / lo operand pairs on the tensor cores (mma.sync m16n8k16, 3 MMAs per product)
extern "C" int lime_mha_x3(const float *qkv, float *ctx, void *ctx_hi, void *ctx_lo, int32_t ld16, float out_scale, int64_t n_news, int T,
                           int d, int nhead, float p_drop, uint64_t seed, int64_t news0, void *stream) {
    LIME_CHECK_ARG(qkv && (ctx || (ctx_hi && ctx_lo)), "lime_mha_x3: null argument");
    LIME_CHECK_ARG(!ctx_hi || (ctx_lo && ld16 >= d && ((((uintptr_t)ctx_hi | (uintptr_t)ctx_lo) & 3) == 0)), "lime_mha_x3: pair output needs both images, ld16 >= d, 4-byte alignment");
    __half *ch = reinterpret_cast<__half *>(ctx_hi), *cl = reinterpret_cast<__half *>(ctx_lo);
    LIME_CHECK_ARG((T == 32 || T == 128) && nhead > 0 && d % nhead == 0 && d / nhead <= 32, "lime_mha_x3: unsupported shape T=%d d=%d heads=%d", T, d, nhead);
    LIME_CHECK_ARG(n_news <= 65535, "lime_mha_x3: at most 65535 news per call (got %lld)", (long long)n_news);
    if (n_news <= 0) return 0;
    const int hd = d / nhead;
    const float scale = 1.0f / sqrtf((float)hd);
    dim3 grid(nhead, (unsigned)n_news);
    const int smem = 4 * T * 40 * (int)sizeof(__half);
    if (T == 32) {
        mha_x3_kernel<32><<<grid, 32, smem, as_stream(stream)>>>(qkv, ctx, ch, cl, ld16, out_scale, d, nhead, hd, scale, p_drop, seed, news0);
    } else {
        LIME_CUDA(cudaFuncSetAttribute(mha_x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mha_x3_kernel<128><<<grid, 128, smem, as_stream(stream)>>>(qkv, ctx, ch, cl, ld16, out_scale, d, nhead, hd, scale, p_drop, seed, news0);
    }
    LIME_LAUNCH_CHECK("mha_x3_kernel");
    return 0;
}

extern "C" int lime_mha_bf16(const void *qkv, int64_t ldq, void *ctx, int64_t ldo, int64_t n_news, int T, int d, int nhead,
                             void *stream) {
    LIME_CHECK_ARG(qkv && ctx, "lime_mha_bf16: null argument");
    LIME_CHECK_ARG(nhead > 0 && d % nhead == 0 && d / nhead <= 32 && (d / nhead) % 2 == 0, "lime_mha_bf16: head dim %d unsupported (even, <= 32)", nhead ? d / nhead : -1);
    LIME_CHECK_ARG(T == 32 || T == 128, "lime_mha_bf16: T=%d unsupported (32 or 128)", T);
    LIME_CHECK_ARG(ldq >= 3 * nhead * 32 && ldq % 8 == 0 && ldo >= d && ldo % 2 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)ctx & 3) == 0,
                   "lime_mha_bf16: qkv must be head-padded [rows, >= %d] with 16-byte aligned rows", 3 * nhead * 32);
    LIME_CHECK_ARG(n_news <= 0x7fffffff, "lime_mha_bf16: too many news");
    if (n_news <= 0) return 0;
    const int hd = d / nhead;
    const float sl = 1.4426950408889634f / sqrtf((float)hd);
    const __nv_bfloat16 *q = reinterpret_cast<const __nv_bfloat16 *>(qkv);
    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(ctx);
    const int smem = 2 * 3 * 128 * 40 * 2;       // two sets of Q, K, V tiles (HPC * T = 128 rows either way)
    if (T == 32) {
        LIME_CUDA(cudaFuncSetAttribute(mha_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mha_tc_kernel<32><<<(unsigned)n_news, 256, smem, as_stream(stream)>>>(q, o, ldq, ldo, d, nhead, hd, sl);
    } else {
        LIME_CUDA(cudaFuncSetAttribute(mha_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mha_tc_kernel<128><<<(unsigned)n_news, 256, smem, as_stream(stream)>>>(q, o, ldq, ldo, d, nhead, hd, sl);
    }
    LIME_LAUNCH_CHECK("mha_tc_kernel");
    return 0;
}

extern "C" int lime_layernorm(const float *x, int64_t ldx, const float *gamma, const float *beta, float *y,
                              int64_t ldy, int64_t rows, int d, float eps, void *stream) {
    LIME_CHECK_ARG(x && gamma && beta && y, "lime_layernorm: null argument");
    LIME_CHECK_ARG(d > 0 && d <= 32 * kLnMaxPerLane, "lime_layernorm: d=%d unsupported (<= 512)", d);
    if (rows <= 0) return 0;
    layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, ldx, gamma, beta, y, ldy, rows, d, eps);
    LIME_LAUNCH_CHECK("layernorm_kernel");
    return 0;
}

extern "C" int lime_layernorm_bf16(const float *x, int64_t ldx, const float *gamma, const float *beta, float *y, int64_t ldy,
                                   void *y16, int32_t ld16, int64_t rows, int d, float eps, void *stream) {
    LIME_CHECK_ARG(x && gamma && beta && y && y16, "lime_layernorm_bf16: null argument");
    LIME_CHECK_ARG(d > 0 && d <= 32 * kLnMaxPerLane && ld16 >= d && ld16 <= 32 * kLnMaxPerLane, "lime_layernorm_bf16: d=%d ld16=%d unsupported (<= 512)", d, ld16);
    if (rows <= 0) return 0;
    layernorm_bf16_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, ldx, gamma, beta, y, ldy,
                                                                                   reinterpret_cast<__nv_bfloat16 *>(y16), ld16, rows, d, eps);
    LIME_LAUNCH_CHECK("layernorm_bf16_kernel");
    return 0;
}

extern "C" int lime_layernorm_pairs(const float *x, int64_t ldx, const float *gamma, const float *beta, float *y, int64_t ldy,
                                    void *hi16, void *lo16, int32_t ld16, float scale, int64_t rows, int d, float eps, void *stream) {
    LIME_CHECK_ARG(x && gamma && beta && y && hi16 && lo16, "lime_layernorm_pairs: null argument");
    LIME_CHECK_ARG(d > 0 && d <= 32 * kLnMaxPerLane && ld16 >= d && ld16 <= 32 * kLnMaxPerLane, "lime_layernorm_pairs: d=%d ld16=%d unsupported (<= 512)", d, ld16);
    if (rows <= 0) return 0;
    layernorm_pairs_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(
        x, ldx, gamma, beta, y, ldy, reinterpret_cast<__half *>(hi16), reinterpret_cast<__half *>(lo16), ld16, scale, rows, d, eps);
    LIME_LAUNCH_CHECK("layernorm_pairs_kernel");
    return 0;
}

extern "C" int lime_layernorm_meanpool(const float *x, const float *gamma, const float *beta, float *out,
                                       int64_t ldo, int64_t n_news, int T, int d, float eps, void *stream) {
    LIME_CHECK_ARG(x && gamma && beta && out, "lime_layernorm_meanpool: null argument");
    LIME_CHECK_ARG(d > 0 && d <= 32 * kLnMaxPerLane && T > 0, "lime_layernorm_meanpool: d=%d unsupported", d);
    if (n_news <= 0) return 0;
    layernorm_meanpool_kernel<<<(unsigned)n_news, 256, 0, as_stream(stream)>>>(x, gamma, beta, out, ldo, T, d, eps);
    LIME_LAUNCH_CHECK("layernorm_meanpool_kernel");
    return 0;
}

extern "C" int lime_topic_rep(const float *cat_emb, const float *sub_emb, const float *W, const float *b,
                              const int32_t *cat, const int32_t *sub, int64_t n, float *out, int64_t ldo,
                              int width, void *stream) {
    LIME_CHECK_ARG(cat_emb && sub_emb && W && b && cat && sub && out, "lime_topic_rep: null argument");
    LIME_CHECK_ARG(width >= LIME_TOPIC && width <= 64, "lime_topic_rep: width %d not in [50,64]", width);
    if (n <= 0) return 0;
    topic_rep_kernel<<<(unsigned)n, 64, 0, as_stream(stream)>>>(cat_emb, sub_emb, W, b, cat, sub, n, out, ldo, width);
    LIME_LAUNCH_CHECK("topic_rep_kernel");
    return 0;
}

extern "C" int lime_intent_pool(const float *pre, const float *e, const float *w2, float *out, int64_t ldo,
                                int64_t n, int k, int D, void *stream) {
    LIME_CHECK_ARG(pre && e && w2 && out, "lime_intent_pool: null argument");
    LIME_CHECK_ARG(k >= 1 && k <= 8, "lime_intent_pool: k=%d unsupported (1..8)", k);
    if (n <= 0) return 0;
    intent_pool_kernel<<<(unsigned)((n + 3) / 4), 128, 0, as_stream(stream)>>>(pre, e, w2, out, ldo, n, k, D);
    LIME_LAUNCH_CHECK("intent_pool_kernel");
    return 0;
}

extern "C" int lime_content_fuse(const float *title, const float *body, const float *cat_emb,
                                 const float *sub_emb, const int32_t *cat, const int32_t *sub, int64_t n,
                                 int D, int cat_dim, int sub_dim, float *content, int64_t ldo, void *stream) {
    LIME_CHECK_ARG(title && body && cat_emb && sub_emb && cat && sub && content, "lime_content_fuse: null argument");
    if (n <= 0) return 0;
    content_fuse_kernel<<<(unsigned)((n + 3) / 4), 128, 0, as_stream(stream)>>>(title, body, cat_emb, sub_emb, cat, sub, n, D, cat_dim, sub_dim, content, ldo);
    LIME_LAUNCH_CHECK("content_fuse_kernel");
    return 0;
}

extern "C" int lime_bucket_pairs(const float *Ef, const float *El, int num_buckets, int dim, float *out,
                                 void *stream) {
    LIME_CHECK_ARG(Ef && El && out && num_buckets >= 1 && dim >= 1, "lime_bucket_pairs: bad argument");
    bucket_pairs_kernel<<<num_buckets * num_buckets, 256, 0, as_stream(stream)>>>(Ef, El, num_buckets, dim, out);
    LIME_LAUNCH_CHECK("bucket_pairs_kernel");
    return 0;
}

extern "C" int lime_scale_rows(float *M, int64_t ld, const float *row_scale, float alpha, int rows, int cols,
                               void *stream) {
    LIME_CHECK_ARG(M && rows > 0 && cols > 0, "lime_scale_rows: bad argument");
    scale_rows_kernel<<<rows, 128, 0, as_stream(stream)>>>(M, ld, row_scale, alpha, rows, cols);
    LIME_LAUNCH_CHECK("scale_rows_kernel");
    return 0;
}

extern "C" int lime_prefix_rows(float *M, int64_t ld, int rows, int cols, void *stream) {
    LIME_CHECK_ARG(M && rows > 0 && cols > 0, "lime_prefix_rows: bad argument");
    prefix_rows_kernel<<<(cols + 127) / 128, 128, 0, as_stream(stream)>>>(M, ld, rows, cols);
    LIME_LAUNCH_CHECK("prefix_rows_kernel");
    return 0;
}

extern "C" int lime_row_absmax(const float *M, int64_t ld, int64_t rows, int cols, float *out, int64_t ldo, void *stream) {
    LIME_CHECK_ARG(M && out && cols > 0, "lime_row_absmax: bad argument");
    if (rows <= 0) return 0;
    row_absmax_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, as_stream(stream)>>>(M, ld, rows, cols, out, ldo);
    LIME_LAUNCH_CHECK("row_absmax_kernel");
    return 0;
}

// bf16-mode backward of lime_mha on the tensor cores (same arguments as lime_mha_bwd, train_kernels.cu)
extern "C" int lime_mha_bwd_bf16(const float *qkv, const float *dctx, float *dqkv, int64_t n_news, int T, int d, int nhead,
                                 float p_drop, uint64_t seed, int64_t news0, void *stream) {
    LIME_CHECK_ARG(qkv && dctx && dqkv, "lime_mha_bwd_bf16: null argument");
    LIME_CHECK_ARG((T == 32 || T == 128) && d % nhead == 0 && d / nhead <= 32, "lime_mha_bwd_bf16: unsupported shape T=%d d=%d heads=%d", T, d, nhead);
    if (n_news <= 0) return 0;
    LIME_CHECK_ARG(n_news <= 65535, "lime_mha_bwd_bf16: at most 65535 news per call");
    const int hd = d / nhead;
    const float scale = 1.0f / sqrtf((float)hd);
    dim3 grid(nhead, (unsigned)n_news);
    const size_t smem = 2 * (4 * (size_t)T * 40 + 2 * (size_t)T * (T + 8));
    if (T == 32) {
        mha_bwd_tc_kernel<32><<<grid, 32, smem, as_stream(stream)>>>(qkv, dctx, dqkv, d, nhead, hd, scale, p_drop, seed, news0);
    } else {
        LIME_CUDA(cudaFuncSetAttribute(mha_bwd_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mha_bwd_tc_kernel<128><<<grid, 128, smem, as_stream(stream)>>>(qkv, dctx, dqkv, d, nhead, hd, scale, p_drop, seed, news0);
    }
    LIME_LAUNCH_CHECK("mha_bwd_tc_kernel");
    return 0;
}

// bf16-mode forward of lime_mha for the training layout (fp32 in / out, dropout on the attention weights)
extern "C" int lime_mha_fwd_bf16(const float *qkv, float *ctx, int64_t n_news, int T, int d, int nhead, float p_drop, uint64_t seed,
                                 int64_t news0, void *stream) {
    LIME_CHECK_ARG(qkv && ctx, "lime_mha_fwd_bf16: null argument");
    LIME_CHECK_ARG((T == 32 || T == 128) && d % nhead == 0 && d / nhead <= 32, "lime_mha_fwd_bf16: unsupported shape T=%d d=%d heads=%d", T, d, nhead);
    if (n_news <= 0) return 0;
    LIME_CHECK_ARG(n_news <= 65535, "lime_mha_fwd_bf16: at most 65535 news per call");
    const int hd = d / nhead;
    const float scale = 1.0f / sqrtf((float)hd);
    dim3 grid(nhead, (unsigned)n_news);
    if (T == 32) mha_fwd_tc_kernel<32><<<grid, 32, 0, as_stream(stream)>>>(qkv, ctx, d, nhead, hd, scale, p_drop, seed, news0);
    else mha_fwd_tc_kernel<128><<<grid, 128, 0, as_stream(stream)>>>(qkv, ctx, d, nhead, hd, scale, p_drop, seed, news0);
    LIME_LAUNCH_CHECK("mha_fwd_tc_kernel");
    return 0;
}

extern "C" int lime_cast_bf16_colsum(const float *x, int64_t ldx, int64_t rows, int32_t d, void *out16, int32_t ld16, float *colsum,
                                     void *stream) {
    LIME_CHECK_ARG(x && out16 && colsum, "lime_cast_bf16_colsum: null argument");
    LIME_CHECK_ARG(d >= 4 && (d & 3) == 0 && ld16 >= d && (ld16 & 7) == 0 && ld16 <= 1024 && (ldx & 3) == 0 && ldx >= d &&
                       (((uintptr_t)x | (uintptr_t)colsum) & 15) == 0 && ((uintptr_t)out16 & 7) == 0,
                   "lime_cast_bf16_colsum: d=%d ld16=%d ldx=%lld (multiples of 4 / 8, ld16 <= 1024, 16-byte aligned)", d, ld16, (long long)ldx);
    if (rows <= 0) return 0;
    cast_bf16_colsum_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, as_stream(stream)>>>(x, ldx, rows, d, reinterpret_cast<uint16_t *>(out16), ld16, colsum);
    LIME_LAUNCH_CHECK("cast_bf16_colsum_kernel");
    return 0;
}
