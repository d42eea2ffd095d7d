// Backward (and training-only forward) kernels of the LIME scoring path (sm_100a): everything
// trainer.py:131-146 differentiates through that is not already a forward kernel of news_kernels.cu.
// All fp32 (the reference trains in fp32, config.py:218-219).  Gradients of parameters are ACCUMULATED
// into caller-zeroed buffers (atomics / split-K), matching autograd's accumulate semantics.
#include "common.cuh"

namespace lime {

// =================================================================================================
// General fp32 GEMM for the backward passes of nn.Linear:  C (+)= alpha * op(A) . op(B)
//   A_KMAJOR: element (i, kk) of op(A) at A[i * lda + kk], else at A[kk * lda + i]
//   B_KMAJOR: element (kk, j) of op(B) at B[j * ldb + kk], else at B[kk * ldb + j]
// 128 x 128 x 16 tiles, 8 x 8 per thread; split-K over gridDim.z with atomicAdd accumulation.
// =================================================================================================
constexpr int TBM = 128, TBN = 128, TBK = 16;

template <bool KMAJOR>
__device__ __forceinline__ void load_tile(float (&dst)[TBK][TBM + 4], const float *__restrict__ src, int64_t ld,
                                          int64_t mn0, int64_t mn_lim, int k0, int k_lim, int tid) {
    if (KMAJOR) {
        // element (i, kk) at src[(mn0 + i) * ld + k0 + kk]: float4 along kk
        const int lr = tid >> 2, lk = (tid & 3) * 4;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = lr + 64 * q;
            const int64_t r = mn0 + i;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (r < mn_lim) {
                const float *p = src + r * ld + k0 + lk;
                if (k0 + lk + 3 < k_lim && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                    const float4 t = *reinterpret_cast<const float4 *>(p);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (k0 + lk + e < k_lim) v[e] = p[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[lk + e][i] = v[e];
        }
    } else {
        // element (i, kk) at src[(k0 + kk) * ld + mn0 + i]: float4 along i
        const int kk = tid >> 5, i4 = (tid & 31) * 4;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = kk + 8 * q;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (k0 + k < k_lim) {
                const float *p = src + (int64_t)(k0 + k) * ld + mn0 + i4;
                if (mn0 + i4 + 3 < mn_lim && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                    const float4 t = *reinterpret_cast<const float4 *>(p);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (mn0 + i4 + e < mn_lim) v[e] = p[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[k][i4 + e] = v[e];
        }
    }
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256, 2)
gemm_general_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb,
                    float *__restrict__ C, int64_t ldc, int64_t m, int n, int64_t k, int64_t k_per_split, float alpha,
                    int accumulate) {
    __shared__ __align__(16) float As[TBK][TBM + 4];
    __shared__ __align__(16) float Bs[TBK][TBN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.y * TBM;
    const int col0 = blockIdx.x * TBN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = (kbeg + k_per_split < k) ? kbeg + k_per_split : k;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    // operands are addressed relative to the split's first k so that int k indices stay small
    const float *Ab = A_KMAJOR ? A + kbeg : A + kbeg * lda;
    const float *Bb = B_KMAJOR ? B + kbeg : B + kbeg * ldb;
    const int klen = (int)(kend - kbeg);
    for (int k0 = 0; k0 < klen; k0 += TBK) {
        load_tile<A_KMAJOR>(As, Ab, lda, row0, m, k0, klen, tid);
        load_tile<B_KMAJOR>(Bs, Bb, ldb, col0, n, k0, klen, tid);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool atomic = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (c >= n) continue;
            float *o = C + r * ldc + c;
            const float v = alpha * acc[i][j];
            if (atomic) atomicAdd(o, v);
            else *o = accumulate ? *o + v : v;
        }
    }
}

// dx = dy * act'(y) from the OUTPUT y of the fused activation (1 relu, 2 tanh); in place allowed
__global__ void act_bwd_kernel(const float *__restrict__ dy, int64_t lddy, const float *__restrict__ y, int64_t ldy,
                               float *__restrict__ dx, int64_t lddx, int64_t rows, int cols, int act) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t r = idx / cols;
    const int c = (int)(idx - r * cols);
    const float g = dy[r * lddy + c], o = y[r * ldy + c];
    dx[r * lddx + c] = act == 1 ? (o > 0.0f ? g : 0.0f) : g * (1.0f - o * o);
}

// out[c] += sum_r M[r][c]   (bias gradients)
__global__ void __launch_bounds__(256) col_sum_kernel(const float *__restrict__ M, int64_t ld, int64_t rows, int cols,
                                                      int64_t rows_per_block, float *__restrict__ out) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float s = 0.0f;
    for (int64_t r = r0; r < r1; ++r) s += M[r * ld + c];
    atomicAdd(out + c, s);
}

// =================================================================================================
// LayerNorm backward (nn.LayerNorm, biased variance).  x: the layer input.  bcast_T > 0: the upstream
// gradient is per NEWS (row r uses dy[(r / bcast_T)] / bcast_T) - the backward of the unmasked token
// mean that follows the second norm of the encoder layer (newsEncoders.py:317,321).
// One warp walks rows r = w, w + W, ...; per-lane column accumulators give dgamma / dbeta with one
// atomicAdd per (warp, column).
// =================================================================================================
constexpr int kLnCols = 16;   // d <= 512

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ gamma,
                     const float *__restrict__ dy, int64_t lddy, int bcast_T, float *__restrict__ dx, int64_t lddx,
                     float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t rows, int d, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), W = (int64_t)gridDim.x * 8;
    float dg[kLnCols], db[kLnCols], gm[kLnCols];
#pragma unroll
    for (int i = 0; i < kLnCols; ++i) {
        const int c = lane + 32 * i;
        dg[i] = 0.0f;
        db[i] = 0.0f;
        gm[i] = c < d ? gamma[c] : 0.0f;
    }
    const float inv_d = 1.0f / (float)d;
    const float gscale = bcast_T > 0 ? 1.0f / (float)bcast_T : 1.0f;
    for (int64_t r = w0; r < rows; r += W) {
        const float *xr = x + r * ldx;
        const float *gr = dy + (bcast_T > 0 ? (r / bcast_T) : r) * lddy;
        float v[kLnCols], g[kLnCols];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < kLnCols; ++i) {
            const int c = lane + 32 * i;
            v[i] = c < d ? xr[c] : 0.0f;
            g[i] = c < d ? gr[c] * gscale : 0.0f;
            s += v[i];
        }
        const float mu = warp_sum(s) * inv_d;
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < kLnCols; ++i) {
            const int c = lane + 32 * i;
            v[i] = c < d ? v[i] - mu : 0.0f;
            q = fmaf(v[i], v[i], q);
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int i = 0; i < kLnCols; ++i) {
            v[i] *= rstd;                       // xhat
            dg[i] = fmaf(g[i], v[i], dg[i]);
            db[i] += g[i];
            g[i] *= gm[i];                      // dxhat
            s1 += g[i];
            s2 = fmaf(g[i], v[i], s2);
        }
        s1 = warp_sum(s1) * inv_d;
        s2 = warp_sum(s2) * inv_d;
#pragma unroll
        for (int i = 0; i < kLnCols; ++i) {
            const int c = lane + 32 * i;
            if (c < d) dx[r * lddx + c] = rstd * (g[i] - s1 - v[i] * s2);
        }
    }
#pragma unroll
    for (int i = 0; i < kLnCols; ++i) {
        const int c = lane + 32 * i;
        if (c < d) {
            atomicAdd(dgamma + c, dg[i]);
            atomicAdd(dbeta + c, db[i]);
        }
    }
}

// =================================================================================================
// Embedding-style gathers / scatters: out[r, :d] = table[ids[r], :d];  dtable[ids[r], :d] += src[r, :d]
// (word embedding + positional encoding backward: the encoding is additive, so src = dx0;
//  category / freshness / lifetime tables likewise.)
// =================================================================================================
__global__ void gather_rows_kernel(const float *__restrict__ table, int64_t ldt, int64_t nrows_table,
                                   const int32_t *__restrict__ ids, int64_t n, int d, float *__restrict__ out, int64_t ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const int64_t r = idx / d;
    const int c = (int)(idx - r * d);
    int64_t id = ids[r];
    id = (id < 0 || id >= nrows_table) ? 0 : id;
    out[r * ldo + c] = table[id * ldt + c];
}

__global__ void scatter_add_rows_kernel(const float *__restrict__ src, int64_t lds, const int32_t *__restrict__ ids,
                                        int64_t n, int d, float *__restrict__ dtable, int64_t ldt, int64_t nrows_table) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const int64_t r = idx / d;
    const int c = (int)(idx - r * d);
    int64_t id = ids[r];
    id = (id < 0 || id >= nrows_table) ? 0 : id;
    atomicAdd(dtable + id * ldt + c, src[r * lds + c]);
}

// 4 columns per thread and ONE 128-bit reduction (red.global.add.v4.f32, sm_90+): a quarter of the atomic traffic of the
// word-embedding gradient (281,600 token rows of 300 columns per training step).  d, lds, ldt multiples of 4, 16-byte
// aligned bases.
__global__ void scatter_add_rows4_kernel(const float *__restrict__ src, int64_t lds, const int32_t *__restrict__ ids,
                                         int64_t n, int d4, float *__restrict__ dtable, int64_t ldt, int64_t nrows_table) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d4) return;
    const int64_t r = idx / d4;
    const int c = 4 * (int)(idx - r * d4);
    int64_t id = ids[r];
    id = (id < 0 || id >= nrows_table) ? 0 : id;
    const float4 v = *reinterpret_cast<const float4 *>(src + r * lds + c);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dtable + id * ldt + c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}


// The same scatter for LARGE id lists with hot rows (the word-embedding gradient: 281,600 token rows per step over a
// Zipf-distributed vocabulary plus the pad id -- unsorted, the reductions of a hot row serialise in L2).  The caller sorts the
// ids (sorted_ids, perm = the source row of every sorted position); a warp walks 32 consecutive sorted positions, keeps the
// running sum of a run of equal ids in registers (lanes = 128-bit column groups) and issues one red.v4 per run and chunk.
__global__ void __launch_bounds__(256)
scatter_add_rows_sorted_kernel(const float *__restrict__ src, int64_t lds, const int32_t *__restrict__ sorted_ids,
                               const int64_t *__restrict__ perm, int64_t n, int d4, float *__restrict__ dtable, int64_t ldt,
                               int64_t nrows_table) {
    const int lane = threadIdx.x & 31;
    const int64_t k0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (k0 >= n) return;
    const int cnt = (int)((n - k0) < 32 ? (n - k0) : 32);
    int64_t my_id = 0, my_row = 0;
    if (lane < cnt) {
        my_id = sorted_ids[k0 + lane];
        my_id = (my_id < 0 || my_id >= nrows_table) ? 0 : my_id;
        my_row = perm[k0 + lane];
    }
    constexpr int G = 4;                           // column groups per lane: d <= 512
    float4 acc[G];
#pragma unroll
    for (int gI = 0; gI < G; ++gI) acc[gI] = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t cur = __shfl_sync(0xffffffffu, my_id, 0);
    auto flush = [&](int64_t id) {
#pragma unroll
        for (int gI = 0; gI < G; ++gI) {
            const int c4 = lane + 32 * gI;
            if (c4 < d4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dtable + id * ldt + 4 * c4), "f"(acc[gI].x),
                             "f"(acc[gI].y), "f"(acc[gI].z), "f"(acc[gI].w)
                             : "memory");
            acc[gI] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
        const int64_t id = __shfl_sync(0xffffffffu, my_id, i);
        const int64_t row = __shfl_sync(0xffffffffu, my_row, i);
        if (id != cur) {                            // warp-uniform
            flush(cur);
            cur = id;
        }
#pragma unroll
        for (int gI = 0; gI < G; ++gI) {
            const int c4 = lane + 32 * gI;
            if (c4 < d4) {
                const float4 v = *reinterpret_cast<const float4 *>(src + row * lds + 4 * c4);
                acc[gI].x += v.x; acc[gI].y += v.y; acc[gI].z += v.z; acc[gI].w += v.w;
            }
        }
    }
    flush(cur);
}

// =================================================================================================
// Self-attention core backward (nn.MultiheadAttention without mask): one CTA per (news, head), one thread per
// token.  Nothing of size T x T is stored: thread i first derives the softmax statistics of query row i
// (max, 1/sum, rowdot_i = sum_j P_ij dP_ij) and dQ_i; then, as key/value token j, it walks all query rows and
// rebuilds P_ij and dS_ij = P_ij (dP_ij - rowdot_i) from those statistics to accumulate dK_j and dV_j.
// Shared memory is only Q, K, V, dO (4 x T x 33 floats = 67 KB at T = 128), so three CTAs share an SM.
// =================================================================================================
template <int T>
__global__ void __launch_bounds__(T)
mha_bwd_kernel(const float *__restrict__ qkv, const float *__restrict__ dctx, float *__restrict__ dqkv, int d, int nhead,
               int hd, float scale, float p_drop, uint64_t seed, int64_t news0) {
    constexpr int HP = 33;                    // head dim (<= 32) padded: conflict-free row reads
    extern __shared__ float sm[];
    float *Qs = sm, *Ks = Qs + T * HP, *Vs = Ks + T * HP, *Os = Vs + T * HP, *st = Os + T * HP;   // st: [3][T]
    const int64_t news = blockIdx.y;
    const int head = blockIdx.x;
    const int i = threadIdx.x;
    const int64_t ld = 3 * (int64_t)d;
    const float *base = qkv + news * T * ld;
    for (int idx = i; idx < T * 32; idx += T) {
        const int r = idx >> 5, e = idx & 31;
        const bool ok = e < hd;
        Qs[r * HP + e] = ok ? base[r * ld + head * hd + e] : 0.0f;
        Ks[r * HP + e] = ok ? base[r * ld + d + head * hd + e] : 0.0f;
        Vs[r * HP + e] = ok ? base[r * ld + 2 * d + head * hd + e] : 0.0f;
        Os[r * HP + e] = ok ? dctx[(news * T + r) * (int64_t)d + head * hd + e] : 0.0f;
    }
    __syncthreads();
    float q[32], o[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        q[e] = Qs[i * HP + e] * scale;
        o[e] = Os[i * HP + e];
    }
    // attention-weight dropout of the forward (same stateless mask): O = (P * M) V, so dP = M * (dO V^T) and dV uses P * M
    const uint64_t drop_nh = ((uint64_t)(news0 + news) * nhead + head) * T;
    // pass 1: row maximum
    float m = -INFINITY;
    for (int j = 0; j < T; ++j) {
        float a = 0.0f;
#pragma unroll
        for (int e = 0; e < 32; ++e) a = fmaf(q[e], Ks[j * HP + e], a);
        m = fmaxf(m, a);
    }
    // pass 2: sum, rowdot and the un-normalised dQ pieces:  dQ_i = s * sum_j P_ij (dP_ij - rowdot) K_j
    float l = 0.0f, pd = 0.0f;
    float a1[32], a2[32];                     // sum_j p~ dP K_j   and   sum_j p~ K_j   (p~ = exp(s - m))
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        a1[e] = 0.0f;
        a2[e] = 0.0f;
    }
    for (int j = 0; j < T; ++j) {
        float a = 0.0f, dp = 0.0f;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            a = fmaf(q[e], Ks[j * HP + e], a);
            dp = fmaf(o[e], Vs[j * HP + e], dp);
        }
        const float p = expf(a - m);
        l += p;
        if (p_drop > 0.0f) dp *= drop_scale(seed, (drop_nh + i) * T + j, p_drop);
        pd = fmaf(p, dp, pd);
        const float pdp = p * dp;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const float k = Ks[j * HP + e];
            a1[e] = fmaf(pdp, k, a1[e]);
            a2[e] = fmaf(p, k, a2[e]);
        }
    }
    const float inv = 1.0f / l;
    const float rowdot = pd * inv;
    st[i] = m;
    st[T + i] = inv;
    st[2 * T + i] = rowdot;
    {
        float *oq = dqkv + (news * T + i) * ld + head * hd;
#pragma unroll
        for (int e = 0; e < 32; ++e)
            if (e < hd) oq[e] = scale * inv * (a1[e] - rowdot * a2[e]);
    }
    __syncthreads();
    // pass 3: this thread as key/value token j = i
    float kj[32], vj[32], dk[32], dv[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        kj[e] = Ks[i * HP + e] * scale;
        vj[e] = Vs[i * HP + e];
        dk[e] = 0.0f;
        dv[e] = 0.0f;
    }
    for (int r = 0; r < T; ++r) {
        float a = 0.0f, dp = 0.0f;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            a = fmaf(Qs[r * HP + e], kj[e], a);
            dp = fmaf(Os[r * HP + e], vj[e], dp);
        }
        const float p = expf(a - st[r]) * st[T + r];
        const float mm = p_drop > 0.0f ? drop_scale(seed, (drop_nh + r) * T + i, p_drop) : 1.0f;
        const float ds = p * (dp * mm - st[2 * T + r]) * scale;
        const float pm = p * mm;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            dv[e] = fmaf(pm, Os[r * HP + e], dv[e]);
            dk[e] = fmaf(ds, Qs[r * HP + e], dk[e]);
        }
    }
    float *ok = dqkv + (news * T + i) * ld + d + head * hd;
    float *ov = dqkv + (news * T + i) * ld + 2 * d + head * hd;
#pragma unroll
    for (int e = 0; e < 32; ++e)
        if (e < hd) {
            ok[e] = dk[e];
            ov[e] = dv[e];
        }
}

// =================================================================================================
// layers.Attention over the k intents, backward (layers.py:285-300).  One warp per news.
//   score_k = w2 . tanh(pre_k),  alpha = softmax_k(score),  out = sum_k alpha_k e_k
// =================================================================================================
__global__ void __launch_bounds__(128)
intent_pool_bwd_kernel(const float *__restrict__ pre, const float *__restrict__ e, const float *__restrict__ w2,
                       const float *__restrict__ dout, int64_t lddo, float *__restrict__ dpre, float *__restrict__ de,
                       float *__restrict__ dw2, int64_t n, int k, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    float sc[8], da[8];
    float m = -INFINITY;
    for (int kk = 0; kk < k; ++kk) {
        const float *p = pre + (i * k + kk) * (int64_t)D;
        const float *ee = e + (i * k + kk) * (int64_t)D;
        float a = 0.0f, b = 0.0f;
        for (int c = lane; c < D; c += 32) {
            a = fmaf(w2[c], tanhf(p[c]), a);
            b = fmaf(dout[i * lddo + c], ee[c], b);
        }
        sc[kk] = warp_sum(a);
        da[kk] = warp_sum(b);
        m = fmaxf(m, sc[kk]);
    }
    float l = 0.0f;
    for (int kk = 0; kk < k; ++kk) {
        sc[kk] = expf(sc[kk] - m);
        l += sc[kk];
    }
    float dot = 0.0f;
    for (int kk = 0; kk < k; ++kk) {
        sc[kk] /= l;                      // alpha_k
        dot = fmaf(sc[kk], da[kk], dot);
    }
    for (int kk = 0; kk < k; ++kk) {
        const float ds = sc[kk] * (da[kk] - dot);      // d score_k
        const float *p = pre + (i * k + kk) * (int64_t)D;
        for (int c = lane; c < D; c += 32) {
            const float t = tanhf(p[c]);
            de[(i * k + kk) * (int64_t)D + c] = sc[kk] * dout[i * lddo + c];
            dpre[(i * k + kk) * (int64_t)D + c] = ds * w2[c] * (1.0f - t * t);
            atomicAdd(dw2 + c, ds * t);
        }
    }
}

// =================================================================================================
// cosine gate + concat, backward (newsEncoders.py:297-300, 367-371): content = [t | sim b | cat | sub],
// sim = (cos(t, b) + 1) / 2.  dcat_rows gets the gradient of the category slice (scattered by the caller).
// =================================================================================================
__global__ void __launch_bounds__(128)
content_fuse_bwd_kernel(const float *__restrict__ title, const float *__restrict__ body, const float *__restrict__ dcontent,
                        int64_t ldc, int64_t n, int D, float *__restrict__ dtitle, float *__restrict__ dbody) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float *t = title + i * (int64_t)D, *b = body + i * (int64_t)D;
    const float *g = dcontent + i * ldc;
    float tb = 0.f, tt = 0.f, bb = 0.f, gs = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float x = t[c], y = b[c];
        tb = fmaf(x, y, tb);
        tt = fmaf(x, x, tt);
        bb = fmaf(y, y, bb);
        gs = fmaf(g[D + c], y, gs);            // d sim = sum_c dcontent_body[c] * body[c]
    }
    tb = warp_sum(tb);
    tt = warp_sum(tt);
    bb = warp_sum(bb);
    gs = warp_sum(gs);
    const float eps = 1e-8f;
    const float nt = sqrtf(tt), nbv = sqrtf(bb);
    const float ct = fmaxf(nt, eps), cb = fmaxf(nbv, eps);
    const float cosv = tb / (ct * cb);
    const float sim = (cosv + 1.0f) * 0.5f;
    const float dcos = 0.5f * gs;
    // d cos / d t = b / (ct cb) - cos * t / nt^2 (when nt > eps), symmetric for b
    const float it = nt > eps ? 1.0f / (nt * nt) : 0.0f, ib = nbv > eps ? 1.0f / (nbv * nbv) : 0.0f;
    const float inv = 1.0f / (ct * cb);
    for (int c = lane; c < D; c += 32) {
        const float x = t[c], y = b[c];
        dtitle[i * (int64_t)D + c] = g[c] + dcos * (y * inv - cosv * x * it);
        dbody[i * (int64_t)D + c] = sim * g[D + c] + dcos * (x * inv - cosv * y * ib);
    }
}

// =================================================================================================
// Inverted dropout with a stateless counter-based generator (forward and backward apply the same
// mask: y = x * keep / (1 - p)).  element index -> 32 random bits via two rounds of a 64-bit mix.
// =================================================================================================
__global__ void dropout_kernel(const float *__restrict__ x, int64_t ldx, float *__restrict__ y, int64_t ldy, int64_t rows,
                               int cols, float p, uint64_t seed) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t r = idx / cols;
    const int c = (int)(idx - r * cols);
    const float u = (float)mix_bits(seed, (uint64_t)idx) * (1.0f / 4294967296.0f);
    y[r * ldy + c] = u >= p ? x[r * ldx + c] * (1.0f / (1.0f - p)) : 0.0f;
}

// The same mask (element index = r * cols + c), 4 consecutive columns per thread: 128-bit loads / stores and one row
// division per quad (the scalar kernel spends its time in the 64-bit division, not in memory).
__global__ void dropout4_kernel(const float *__restrict__ x, int64_t ldx, float *__restrict__ y, int64_t ldy, int64_t rows,
                                int cols4, float p, uint64_t seed) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols4) return;
    const int64_t r = idx / cols4;
    const int c = 4 * (int)(idx - r * cols4);
    const float4 v = *reinterpret_cast<const float4 *>(x + r * ldx + c);
    const uint64_t e0 = (uint64_t)(r * (4 * (int64_t)cols4) + c);
    const float keep = 1.0f / (1.0f - p), k32 = 1.0f / 4294967296.0f;
    float4 o;
    o.x = (float)mix_bits(seed, e0) * k32 >= p ? v.x * keep : 0.0f;
    o.y = (float)mix_bits(seed, e0 + 1) * k32 >= p ? v.y * keep : 0.0f;
    o.z = (float)mix_bits(seed, e0 + 2) * k32 >= p ? v.z * keep : 0.0f;
    o.w = (float)mix_bits(seed, e0 + 3) * k32 >= p ? v.w * keep : 0.0f;
    *reinterpret_cast<float4 *>(y + r * ldy + c) = o;
}


// Fused forms of the dropout sites of a transformer branch (bf16 and fp32 training alike, same masks as dropout_kernel):
//   y = x * m(seed) [* m(seed2)] [+ res]          4 columns per thread
// "+ res": dropout(out_proj(.)) + x and dropout(linear2(.)) + x1 (newsEncoders.py:244-247, post-LN residuals) in one pass;
// two masks: the backward of the embedding path below.
__global__ void dropout_fused4_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ res, int64_t ldr,
                                      float *__restrict__ y, int64_t ldy, int64_t rows, int cols4, float p, uint64_t seed,
                                      uint64_t seed2, int two) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols4) return;
    const int64_t r = idx / cols4;
    const int c = 4 * (int)(idx - r * cols4);
    const float4 v = *reinterpret_cast<const float4 *>(x + r * ldx + c);
    const uint64_t e0 = (uint64_t)(r * (4 * (int64_t)cols4) + c);
    float m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        m[e] = drop_scale(seed, e0 + e, p);
        if (two) m[e] *= drop_scale(seed2, e0 + e, p);
    }
    float4 o = make_float4(v.x * m[0], v.y * m[1], v.z * m[2], v.w * m[3]);
    if (res != nullptr) {
        const float4 q = *reinterpret_cast<const float4 *>(res + r * ldr + c);
        o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
    }
    *reinterpret_cast<float4 *>(y + r * ldy + c) = o;
}

// word_embedding(ids) -> dropout(seed_w) -> + positional encoding -> dropout(seed_x) in one pass (newsEncoders.py:311-315,
// :828 in training mode; the unfused path took a gather, two dropouts, a repeat and an add over the [tokens, 300] rows):
//   out[r, c] = m_x(r d + c) * (m_w(r d + c) * E[ids[r], c] + pe[r % T, c])
__global__ void embed_pe_dropout_kernel(const float *__restrict__ E, int64_t vocab, const int32_t *__restrict__ ids, int64_t rows,
                                        int T, int d4, const float *__restrict__ pe, float p, uint64_t seed_w, uint64_t seed_x,
                                        float *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * d4) return;
    const int64_t r = idx / d4;
    const int c = 4 * (int)(idx - r * d4);
    const int d = 4 * d4;
    int64_t id = ids[r];
    id = (id < 0 || id >= vocab) ? 0 : id;
    const float4 w = *reinterpret_cast<const float4 *>(E + id * d + c);
    const float4 q = *reinterpret_cast<const float4 *>(pe + (int64_t)(r % T) * d + c);
    const uint64_t e0 = (uint64_t)(r * d + c);
    float4 o;
    o.x = drop_scale(seed_x, e0, p) * fmaf(drop_scale(seed_w, e0, p), w.x, q.x);
    o.y = drop_scale(seed_x, e0 + 1, p) * fmaf(drop_scale(seed_w, e0 + 1, p), w.y, q.y);
    o.z = drop_scale(seed_x, e0 + 2, p) * fmaf(drop_scale(seed_w, e0 + 2, p), w.z, q.z);
    o.w = drop_scale(seed_x, e0 + 3, p) * fmaf(drop_scale(seed_w, e0 + 3, p), w.w, q.w);
    *reinterpret_cast<float4 *>(out + r * d + c) = o;
}

}  // namespace lime

using namespace lime;

// C (+)= alpha * op(A) . op(B);  a_kmajor / b_kmajor as documented above;  accumulate: add to C.
extern "C" int lime_gemm(const float *A, int64_t lda, int a_kmajor, const float *B, int64_t ldb, int b_kmajor, float *C,
                         int64_t ldc, int64_t m, int n, int64_t k, float alpha, int accumulate, void *stream) {
    LIME_CHECK_ARG(A && B && C && m > 0 && n > 0 && k > 0, "lime_gemm: bad argument");
    cudaStream_t st = as_stream(stream);
    const int64_t gy = (m + TBM - 1) / TBM;
    const int gx = (n + TBN - 1) / TBN;
    LIME_CHECK_ARG(gy <= 65535, "lime_gemm: m too large (%lld rows)", (long long)m);
    // split K when the output grid cannot fill the GPU (weight gradients: k = all tokens)
    int64_t splits = 1;
    const int64_t ctas = gy * gx, target = 2LL * num_sms();
    if (ctas < target && k >= 4096) {
        splits = (target + ctas - 1) / ctas;
        const int64_t max_splits = k / 1024;
        if (splits > max_splits) splits = max_splits;
        if (splits > 65535) splits = 65535;
        if (splits < 1) splits = 1;
    }
    int64_t kps = (k + splits - 1) / splits;
    kps = (kps + TBK - 1) / TBK * TBK;
    splits = (k + kps - 1) / kps;
    if (splits > 1 && !accumulate) {
        if (ldc == n) {
            LIME_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)m * n, st));
        } else {
            LIME_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * n, (size_t)m, st));
        }
    }
    dim3 grid(gx, (unsigned)gy, (unsigned)splits);
    if (a_kmajor && b_kmajor) gemm_general_kernel<true, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, m, n, k, kps, alpha, accumulate);
    else if (a_kmajor) gemm_general_kernel<true, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, m, n, k, kps, alpha, accumulate);
    else if (b_kmajor) gemm_general_kernel<false, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, m, n, k, kps, alpha, accumulate);
    else gemm_general_kernel<false, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, m, n, k, kps, alpha, accumulate);
    LIME_LAUNCH_CHECK("gemm_general_kernel");
    return 0;
}

extern "C" int lime_act_bwd(const float *dy, int64_t lddy, const float *y, int64_t ldy, float *dx, int64_t lddx,
                            int64_t rows, int cols, int act, void *stream) {
    LIME_CHECK_ARG(dy && y && dx && (act == 1 || act == 2), "lime_act_bwd: bad argument");
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t total = rows * cols;
    act_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(dy, lddy, y, ldy, dx, lddx, rows, cols, act);
    LIME_LAUNCH_CHECK("act_bwd_kernel");
    return 0;
}

extern "C" int lime_col_sum(const float *M, int64_t ld, int64_t rows, int cols, float *out, void *stream) {
    LIME_CHECK_ARG(M && out && cols > 0, "lime_col_sum: bad argument");
    if (rows <= 0) return 0;
    const int64_t rpb = 512;
    dim3 grid((cols + 255) / 256, (unsigned)((rows + rpb - 1) / rpb));
    LIME_CHECK_ARG(grid.y <= 65535, "lime_col_sum: too many rows");
    col_sum_kernel<<<grid, 256, 0, as_stream(stream)>>>(M, ld, rows, cols, rpb, out);
    LIME_LAUNCH_CHECK("col_sum_kernel");
    return 0;
}

extern "C" int lime_layernorm_bwd(const float *x, int64_t ldx, const float *gamma, const float *dy, int64_t lddy,
                                  int bcast_T, float *dx, int64_t lddx, float *dgamma, float *dbeta, int64_t rows, int d,
                                  float eps, void *stream) {
    LIME_CHECK_ARG(x && gamma && dy && dx && dgamma && dbeta, "lime_layernorm_bwd: null argument");
    LIME_CHECK_ARG(d >= 1 && d <= 32 * kLnCols, "lime_layernorm_bwd: d=%d unsupported", d);
    if (rows <= 0) return 0;
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = 8LL * num_sms();
    if (blocks > cap) blocks = cap;
    layernorm_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, ldx, gamma, dy, lddy, bcast_T, dx, lddx, dgamma,
                                                                          dbeta, rows, d, eps);
    LIME_LAUNCH_CHECK("layernorm_bwd_kernel");
    return 0;
}

extern "C" int lime_gather_rows(const float *table, int64_t ldt, int64_t table_rows, const int32_t *ids, int64_t n, int d,
                                float *out, int64_t ldo, void *stream) {
    LIME_CHECK_ARG(table && ids && out && d > 0 && table_rows > 0, "lime_gather_rows: bad argument");
    if (n <= 0) return 0;
    const int64_t total = n * d;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(table, ldt, table_rows, ids, n, d, out, ldo);
    LIME_LAUNCH_CHECK("gather_rows_kernel");
    return 0;
}

extern "C" int lime_scatter_add_rows(const float *src, int64_t lds, const int32_t *ids, int64_t n, int d, float *dtable,
                                     int64_t ldt, int64_t table_rows, void *stream) {
    LIME_CHECK_ARG(src && ids && dtable && d > 0 && table_rows > 0, "lime_scatter_add_rows: bad argument");
    if (n <= 0) return 0;
    const int64_t total = n * d;
    LIME_CHECK_ARG((total + 255) / 256 < (1LL << 31), "lime_scatter_add_rows: too large");
    if ((d & 3) == 0 && (lds & 3) == 0 && (ldt & 3) == 0 && (((uintptr_t)src | (uintptr_t)dtable) & 15) == 0) {
        scatter_add_rows4_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, as_stream(stream)>>>(src, lds, ids, n, d / 4, dtable,
                                                                                                  ldt, table_rows);
        LIME_LAUNCH_CHECK("scatter_add_rows4_kernel");
        return 0;
    }
    scatter_add_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(src, lds, ids, n, d, dtable, ldt,
                                                                                         table_rows);
    LIME_LAUNCH_CHECK("scatter_add_rows_kernel");
    return 0;
}


extern "C" int lime_scatter_add_rows_sorted(const float *src, int64_t lds, const int32_t *sorted_ids, const int64_t *perm, int64_t n,
                                            int d, float *dtable, int64_t ldt, int64_t table_rows, void *stream) {
    LIME_CHECK_ARG(src && sorted_ids && perm && dtable && d > 0 && table_rows > 0, "lime_scatter_add_rows_sorted: bad argument");
    LIME_CHECK_ARG((d & 3) == 0 && d <= 512 && (lds & 3) == 0 && (ldt & 3) == 0 && (((uintptr_t)src | (uintptr_t)dtable) & 15) == 0,
                   "lime_scatter_add_rows_sorted: d=%d, lds=%lld, ldt=%lld must be multiples of 4 (d <= 512), 16-byte aligned bases", d,
                   (long long)lds, (long long)ldt);
    if (n <= 0) return 0;
    const int64_t warps = (n + 31) / 32;
    scatter_add_rows_sorted_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, as_stream(stream)>>>(src, lds, sorted_ids, perm, n, d / 4, dtable,
                                                                                             ldt, table_rows);
    LIME_LAUNCH_CHECK("scatter_add_rows_sorted_kernel");
    return 0;
}

extern "C" int lime_mha_bwd(const float *qkv, const float *dctx, float *dqkv, int64_t n_news, int T, int d, int nhead,
                            float p_drop, uint64_t seed, int64_t news0, void *stream) {
    LIME_CHECK_ARG(qkv && dctx && dqkv, "lime_mha_bwd: null argument");
    LIME_CHECK_ARG((T == 32 || T == 128) && d % nhead == 0 && d / nhead <= 32, "lime_mha_bwd: unsupported shape T=%d d=%d heads=%d", T, d, nhead);
    if (n_news <= 0) return 0;
    LIME_CHECK_ARG(n_news <= 65535, "lime_mha_bwd: at most 65535 news per call");
    const int hd = d / nhead;
    const float scale = 1.0f / sqrtf((float)hd);
    dim3 grid(nhead, (unsigned)n_news);
    const size_t smem = sizeof(float) * (4 * (size_t)T * 33 + 3 * (size_t)T);
    if (T == 32) {
        mha_bwd_kernel<32><<<grid, 32, smem, as_stream(stream)>>>(qkv, dctx, dqkv, d, nhead, hd, scale, p_drop, seed, news0);
    } else {
        LIME_CUDA(cudaFuncSetAttribute(mha_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mha_bwd_kernel<128><<<grid, 128, smem, as_stream(stream)>>>(qkv, dctx, dqkv, d, nhead, hd, scale, p_drop, seed, news0);
    }
    LIME_LAUNCH_CHECK("mha_bwd_kernel");
    return 0;
}

extern "C" int lime_intent_pool_bwd(const float *pre, const float *e, const float *w2, const float *dout, int64_t lddo,
                                    float *dpre, float *de, float *dw2, int64_t n, int k, int D, void *stream) {
    LIME_CHECK_ARG(pre && e && w2 && dout && dpre && de && dw2, "lime_intent_pool_bwd: null argument");
    LIME_CHECK_ARG(k >= 1 && k <= 8, "lime_intent_pool_bwd: k=%d unsupported (1..8)", k);
    if (n <= 0) return 0;
    intent_pool_bwd_kernel<<<(unsigned)((n + 3) / 4), 128, 0, as_stream(stream)>>>(pre, e, w2, dout, lddo, dpre, de, dw2, n, k, D);
    LIME_LAUNCH_CHECK("intent_pool_bwd_kernel");
    return 0;
}

extern "C" int lime_content_fuse_bwd(const float *title, const float *body, const float *dcontent, int64_t ldc, int64_t n,
                                     int D, float *dtitle, float *dbody, void *stream) {
    LIME_CHECK_ARG(title && body && dcontent && dtitle && dbody, "lime_content_fuse_bwd: null argument");
    if (n <= 0) return 0;
    content_fuse_bwd_kernel<<<(unsigned)((n + 3) / 4), 128, 0, as_stream(stream)>>>(title, body, dcontent, ldc, n, D, dtitle, dbody);
    LIME_LAUNCH_CHECK("content_fuse_bwd_kernel");
    return 0;
}

extern "C" int lime_dropout(const float *x, int64_t ldx, float *y, int64_t ldy, int64_t rows, int cols, float p,
                            uint64_t seed, void *stream) {
    LIME_CHECK_ARG(x && y && p >= 0.0f && p < 1.0f, "lime_dropout: bad argument");
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t total = rows * cols;
    if ((cols & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
        dropout4_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, as_stream(stream)>>>(x, ldx, y, ldy, rows, cols / 4, p, seed);
        LIME_LAUNCH_CHECK("dropout4_kernel");
        return 0;
    }
    dropout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, ldx, y, ldy, rows, cols, p, seed);
    LIME_LAUNCH_CHECK("dropout_kernel");
    return 0;
}

extern "C" int lime_dropout_fused(const float *x, int64_t ldx, const float *res, int64_t ldr, float *y, int64_t ldy, int64_t rows,
                                  int cols, float p, uint64_t seed, uint64_t seed2, int two_masks, void *stream) {
    LIME_CHECK_ARG(x && y && p >= 0.0f && p < 1.0f, "lime_dropout_fused: bad argument");
    LIME_CHECK_ARG((cols & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0 &&
                       (res == nullptr || ((ldr & 3) == 0 && ((uintptr_t)res & 15) == 0)),
                   "lime_dropout_fused: cols and leading dimensions must be multiples of 4, bases 16-byte aligned");
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t total = rows * (cols / 4);
    dropout_fused4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, ldx, res, ldr, y, ldy, rows, cols / 4, p, seed,
                                                                                        seed2, two_masks);
    LIME_LAUNCH_CHECK("dropout_fused4_kernel");
    return 0;
}

extern "C" int lime_embed_pe_dropout(const float *E, int64_t vocab, const int32_t *ids, int64_t rows, int T, int d, const float *pe,
                                     float p, uint64_t seed_w, uint64_t seed_x, float *out, void *stream) {
    LIME_CHECK_ARG(E && ids && pe && out && vocab > 0 && T > 0 && p >= 0.0f && p < 1.0f, "lime_embed_pe_dropout: bad argument");
    LIME_CHECK_ARG((d & 3) == 0 && d > 0 && (((uintptr_t)E | (uintptr_t)pe | (uintptr_t)out) & 15) == 0,
                   "lime_embed_pe_dropout: d must be a multiple of 4, bases 16-byte aligned");
    if (rows <= 0) return 0;
    const int64_t total = rows * (d / 4);
    embed_pe_dropout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(E, vocab, ids, rows, T, d / 4, pe, p, seed_w,
                                                                                          seed_x, out);
    LIME_LAUNCH_CHECK("embed_pe_dropout_kernel");
    return 0;
}
