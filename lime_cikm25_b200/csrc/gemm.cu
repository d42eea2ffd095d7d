// fp32 dense layers for the news encoder (sm_100a, CUDA cores: the reference's GEMMs are true fp32,
// config.py:218-219 never enables TF32, and the fp32-mode parity bar is 1e-4 relative on logits).
//
//   lime_linear        C = act(A . W^T + bias) + residual     128x128x16 tiles, 8x8 per thread,
//                      register-prefetch double buffering, 128-bit global and shared accesses.
//   lime_gemm_strided  small any-stride GEMM used once per checkpoint to fold weights.
#include "common.cuh"

namespace lime {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int LDS_A = BM + 4, LDS_B = BN + 4;

template <int ACT>
__device__ __forceinline__ float activate(float x) {
    if (ACT == 1) return fmaxf(x, 0.0f);
    if (ACT == 2) return tanhf(x);
    return x;
}

template <int ACT>
__global__ void __launch_bounds__(256, 2)
linear_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ W, int64_t ldw,
              const float *__restrict__ bias, const float *__restrict__ residual, int64_t ldr,
              float *__restrict__ C, int64_t ldc, int64_t m, int n, int k) {
    __shared__ __align__(16) float As[2][BK][LDS_A];
    __shared__ __align__(16) float Bs[2][BK][LDS_B];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.y * BM;
    const int col0 = blockIdx.x * BN;

    // global->register staging: 2 float4 of A and 2 of W per thread per k-tile
    const int lr = tid >> 2;          // 0..63 (+64 for the second)
    const int lk = (tid & 3) * 4;     // 0,4,8,12
    float4 ra[2], rb[2];
    auto load_tiles = [&](int kt) {
        const int kk = kt * BK + lk;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t r = row0 + lr + 64 * q;
            ra[q] = (r < m && kk < k) ? *reinterpret_cast<const float4 *>(A + r * lda + kk)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            const int c = col0 + lr + 64 * q;
            rb[q] = (c < n && kk < k) ? *reinterpret_cast<const float4 *>(W + (int64_t)c * ldw + kk)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int r = lr + 64 * q;
            As[buf][lk + 0][r] = ra[q].x;
            As[buf][lk + 1][r] = ra[q].y;
            As[buf][lk + 2][r] = ra[q].z;
            As[buf][lk + 3][r] = ra[q].w;
            Bs[buf][lk + 0][r] = rb[q].x;
            Bs[buf][lk + 1][r] = rb[q].y;
            Bs[buf][lk + 2][r] = rb[q].z;
            Bs[buf][lk + 3][r] = rb[q].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int ktiles = (k + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < ktiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) load_tiles(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                        (residual == nullptr ||
                         (((ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(residual) & 15) == 0)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= m) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int c = col0 + jh * 64 + tx * 4;
            if (c >= n) continue;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float x = acc[i][jh * 4 + e];
                if (bias != nullptr && c + e < n) x += bias[c + e];
                o[e] = activate<ACT>(x);
            }
            if (vec_ok && c + 3 < n) {
                if (residual != nullptr) {
                    const float4 rr = *reinterpret_cast<const float4 *>(residual + r * ldr + c);
                    o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
                }
                *reinterpret_cast<float4 *>(C + r * ldc + c) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (c + e < n) {
                        float x = o[e];
                        if (residual != nullptr) x += residual[r * ldr + c + e];
                        C[r * ldc + c + e] = x;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
gemm_strided_kernel(const float *__restrict__ A, int64_t sam, int64_t sak, const float *__restrict__ B,
                    int64_t sbk, int64_t sbn, float *__restrict__ C, int64_t ldc, int m, int n, int k,
                    float alpha) {
    __shared__ float As[16][17], Bs[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
    float acc = 0.0f;
    for (int k0 = 0; k0 < k; k0 += 16) {
        As[ty][tx] = (i < m && k0 + tx < k) ? A[(int64_t)i * sam + (int64_t)(k0 + tx) * sak] : 0.0f;
        Bs[ty][tx] = (k0 + ty < k && j < n) ? B[(int64_t)(k0 + ty) * sbk + (int64_t)j * sbn] : 0.0f;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc = fmaf(As[ty][kk], Bs[kk][tx], acc);
        __syncthreads();
    }
    if (i < m && j < n) C[(int64_t)i * ldc + j] = alpha * acc;
}

}  // namespace lime

using namespace lime;

extern "C" int lime_linear(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias,
                           const float *residual, int64_t ldr, float *C, int64_t ldc, int64_t m,
                           int n, int k, int act, void *stream) {
    LIME_CHECK_ARG(A && W && C, "lime_linear: null operand");
    LIME_CHECK_ARG(m >= 0 && n > 0 && k > 0, "lime_linear: bad shape m=%lld n=%d k=%d", (long long)m, n, k);
    LIME_CHECK_ARG((k & 3) == 0 && (lda & 3) == 0 && (ldw & 3) == 0,
                   "lime_linear: k=%d, lda=%lld, ldw=%lld must be multiples of 4", k, (long long)lda, (long long)ldw);
    LIME_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                   "lime_linear: A and W must be 16-byte aligned");
    LIME_CHECK_ARG(act >= 0 && act <= 2, "lime_linear: unknown activation %d", act);
    if (m == 0) return 0;
    const int64_t gy = (m + BM - 1) / BM;
    LIME_CHECK_ARG(gy <= 65535 * 1024LL, "lime_linear: m too large");
    // grid.y is limited to 65535: split very tall problems into row slabs
    const int64_t max_rows = 65535LL * BM;
    cudaStream_t st = as_stream(stream);
    for (int64_t r0 = 0; r0 < m; r0 += max_rows) {
        const int64_t mm = (m - r0 < max_rows) ? (m - r0) : max_rows;
        dim3 grid((n + BN - 1) / BN, (unsigned)((mm + BM - 1) / BM));
        const float *Ap = A + r0 * lda;
        const float *Rp = residual ? residual + r0 * ldr : nullptr;
        float *Cp = C + r0 * ldc;
        if (act == 0) linear_kernel<0><<<grid, 256, 0, st>>>(Ap, lda, W, ldw, bias, Rp, ldr, Cp, ldc, mm, n, k);
        else if (act == 1) linear_kernel<1><<<grid, 256, 0, st>>>(Ap, lda, W, ldw, bias, Rp, ldr, Cp, ldc, mm, n, k);
        else linear_kernel<2><<<grid, 256, 0, st>>>(Ap, lda, W, ldw, bias, Rp, ldr, Cp, ldc, mm, n, k);
        LIME_LAUNCH_CHECK("linear_kernel");
    }
    return 0;
}

extern "C" int lime_gemm_strided(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk,
                                 int64_t sbn, float *C, int64_t ldc, int m, int n, int k, float alpha,
                                 void *stream) {
    LIME_CHECK_ARG(A && B && C && m > 0 && n > 0 && k > 0, "lime_gemm_strided: bad argument");
    dim3 grid((n + 15) / 16, (m + 15) / 16);
    gemm_strided_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, sam, sak, B, sbk, sbn, C, ldc, m, n, k, alpha);
    LIME_LAUNCH_CHECK("gemm_strided_kernel");
    return 0;
}
