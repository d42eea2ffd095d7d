// Definitions shared by the two scoring kernels: score.cu (exact per-pair evaluation, any H) and
// score_tc.cu (gate interpolated in the attention weight + tcgen05 dots, H <= 64).
#pragma once

#include "common.cuh"

namespace lime {

struct ScoreArgs {
    LimeNewsCache cache;
    LimeImpressions imp;
    long long pair_index_base;
    long long tail_start;
    int prefix_main;
    int prefix_tail;
    float bucket_scale;
    float ln_eps;
    float *scores;
    int *work_counter;
    const int *unit_list;        // exact kernel as fallback pass: unit ids flagged by score_tc_kernel ...
    const int *unit_list_count;  // ... and how many (device-side); NULL = all units 0..num_units-1
    int *fallback_list;          // score_tc_kernel: where flagged units are appended
    int *fallback_count;
    float interp_tol;            // score_tc_kernel: admissible interpolation error of the gate (<= 0: flag every unit)
    // Long histories (H > LIME_TC_MAX_HISTORY) on the tensor-core kernel: the history is cut into chunks of `max_history` slots,
    // every (impression, chunk) is a pseudo-impression (chunk-major: pseudo-impression k * chunk_impressions + i, pairs k *
    // chunk_pairs + p), the attention weights over the FULL history come from a pre-pass (a_matrix) and the kernel writes the
    // partial pooling state of its chunk instead of the score.  a_matrix == NULL: ordinary call.
    const float *a_matrix;       // [chunks * chunk_pairs, max_history]
    float4 *partial_out;         // [chunks * chunk_pairs]: m, l, acc of the pooling softmax, GraphSAGE prefix sum
    float2 *cbw_out;             // [chunk_pairs]: cb, lifetime weight (written by chunk 0)
    int chunk_impressions;       // real impressions
    int full_history;            // H of the whole history
    long long chunk_pairs;       // real pairs
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// RemainingLifetimeWeighting weight (util.py:39-46), IEEE fp32 like torch's CUDA sigmoid.
__device__ __forceinline__ float lifetime_weight(float r, const LimeNewsCache &c) {
    if (!c.use_lifetime_weighting) return 1.0f;
    if (c.use_expired_penalty) {
        float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-__fmul_rn(c.sigmoid_alpha, r))));
        float pos = (r >= 0.0f) ? 1.0f : 0.0f;
        float neg = (r < 0.0f) ? 1.0f : 0.0f;
        return __fadd_rn(__fmul_rn(pos, s), __fmul_rn(__fmul_rn(neg, c.penalty_beta), s));
    }
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-__fmul_rn(c.sigmoid_alpha, fabsf(r)))));
}


// host-side launchers (each in its own translation unit)
int launch_score_exact(const ScoreArgs &a, int grid_limit, cudaStream_t st);
int launch_score_tc(const ScoreArgs &a, cudaStream_t st);
int launch_attention_long(const ScoreArgs &orig, float *a_matrix, int chunks, int chunk_slots, long long total_pairs, cudaStream_t st);
int launch_merge_long(const ScoreArgs &a, const float4 *partial, const float2 *cbw, int chunks, long long total_pairs, float *scores, cudaStream_t st);
int64_t score_exact_smem(int H, int TC);

}  // namespace lime
