// Micro-benchmark (GPU box): cost of many small TMA loads.  nvcc -arch=sm_100a -o tma_bench tma_bench.cu && ./tma_bench
// Every CTA keeps R rounds in flight; a round = L lanes of warp 0 each issuing one box {64 fp16, ROWS rows} (or a gather4 of
// 4 rows) from random rows of a [rows][400] fp16 table into shared memory; reports clocks per op (issue side) and per round.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mb_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mb_try(uint64_t *b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap *m, int c, int r, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(m), "r"(c), "r"(r), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_g4(uint32_t dst, const CUtensorMap *m, int c, int r0, int r1, int r2, int r3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(m), "r"(c), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(s32(bar)) : "memory");
}
constexpr int R = 4;
template <int ROWS, bool G4>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, int nrows, int L, int W, int rounds, unsigned long long *out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
    unsigned char *buf = sm + 1024;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rb = (uint32_t)(ROWS * 128);
    if (tid == 0) { for (int i = 0; i < R; ++i) mb_init(bars + i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    unsigned long long t_issue = 0;
    uint32_t rng = blockIdx.x * 7919u + tid * 104729u + 1u;
    long long t0 = clock64();
    if (warp < W) {
        for (int r = 0; r < rounds + R - 1; ++r) {
            if (r < rounds) {
                const int s = r % R;
                if (warp == 0 && lane == 0) mb_expect(bars + s, (uint32_t)(W * L) * rb);
                __syncwarp();
                long long a = clock64();
                if (lane < L) {
                    rng = rng * 1664525u + 1013904223u;
                    const int row = (int)((rng >> 8) % (uint32_t)(nrows - 8));
                    const int col = 64 * (int)((rng >> 4) % 6u);
                    const uint32_t dst = s32(buf) + (uint32_t)s * (uint32_t)(W * L) * rb + (uint32_t)(warp * L + lane) * rb;
                    if (G4) tma_g4(dst, &map, col, row, row + 3, row + 1, row + 5, bars + s);
                    else tma2d(dst, &map, col, row, bars + s);
                }
                __syncwarp();
                if (lane == 0 && warp == 0) t_issue += (unsigned long long)(clock64() - a);
            }
            if (r >= R - 1) {
                const int w = r - (R - 1), s = w % R;
                while (!mb_try(bars + s, (uint32_t)(w / R) & 1u)) { }
            }
            // rounds are consumed in order by every issuing warp: keeps the ring consistent across warps
            asm volatile("bar.sync 1, %0;" ::"r"(W * 32) : "memory");
        }
    }
    __syncthreads();
    if (tid == 0) { atomicAdd(out, (unsigned long long)(clock64() - t0)); atomicAdd(out + 1, t_issue); }
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int ROWS, bool G4>
void run(EncodeFn enc, void *d, uint64_t nrows, int L, int W, int ctas_per_sm) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {400, nrows}; const cuuint64_t str[1] = {800};
    const cuuint32_t box[2] = {64, (cuuint32_t)(G4 ? 1 : ROWS)}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    unsigned long long *out; CK(cudaMalloc(&out, 16)); CK(cudaMemset(out, 0, 16));
    const int smem = 1024 + R * W * L * ROWS * 128;
    if (smem > 110000 && ctas_per_sm > 1) { printf("rows %d skip (smem)\n", ROWS); return; }
    CK(cudaFuncSetAttribute(k<ROWS, G4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = 148 * ctas_per_sm, rounds = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ROWS, G4><<<grid, 128, smem>>>(m, (int)nrows, L, W, 20, out);
    CK(cudaMemset(out, 0, 16));
    cudaEventRecord(e0);
    k<ROWS, G4><<<grid, 128, smem>>>(m, (int)nrows, L, W, rounds, out);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
    const double ops = (double)grid * rounds * W * L, bytes = ops * ROWS * 128;
    printf("%s rows/op %2d  lanes %2d warps %d ctas/sm %d: %.3f ms  %.1f GB/s  clocks/round %.0f  issue clocks/round(warp0) %.0f  -> %.1f clk per op per SM\n", G4 ? "gather4" : "tiled  ", ROWS, L, W, ctas_per_sm, ms, bytes / ms / 1e6,
           (double)h[0] / grid / rounds, (double)h[1] / grid / rounds, (double)h[0] / grid / rounds / (W * L) / ctas_per_sm);
    cudaFree(out);
}
int main() {
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    const uint64_t nrows = 6 * 65538;      // 315 MB: DRAM resident
    void *d; CK(cudaMalloc(&d, nrows * 800)); CK(cudaMemset(d, 1, nrows * 800));
    for (int cps = 1; cps <= 2; ++cps) {
        run<3, false>(enc, d, nrows, 12, 1, cps);
        run<3, false>(enc, d, nrows, 32, 1, cps);
        run<3, false>(enc, d, nrows, 12, 4, cps);
        run<3, false>(enc, d, nrows, 1, 4, cps);
        run<4, true>(enc, d, nrows, 32, 1, cps);
        run<4, true>(enc, d, nrows, 12, 4, cps);
        run<6, false>(enc, d, nrows, 32, 1, cps);
        run<6, false>(enc, d, nrows, 12, 4, cps);
        run<1, false>(enc, d, nrows, 32, 1, cps);
        run<8, false>(enc, d, nrows, 16, 1, cps);
        run<16, false>(enc, d, nrows, 8, 1, cps);
    }
    // L2-resident table
    const uint64_t small = 6 * 4096;
    printf("L2-resident table (%.1f MB)\n", small * 800 / 1e6);
    run<3, false>(enc, d, small, 32, 1, 2);
    run<3, false>(enc, d, small, 12, 4, 2);
    run<4, true>(enc, d, small, 32, 1, 2);
    run<6, false>(enc, d, small, 32, 1, 2);
    run<16, false>(enc, d, small, 8, 1, 2);
    return 0;
}
