// rl_kernels.cu -- rollout-storage kernels (include/fortattack_rollout.h).
//
// rl_gae: one thread per (agent, env) column walks the T steps backwards; a warp's 32 lanes are 32
// consecutive envs of one agent, so every load/store is a coalesced 128-byte request.  HBM-bound streaming:
// 4 floats read (reward, V[t+1] is carried in a register, V[t], mask) + 1 written per (t, agent, env).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fortattack_rollout.h"

int fa_internal_fail(int code, const char *fmt, ...);

namespace rl {

__global__ void __launch_bounds__(128) gae_kernel(const float *__restrict__ rew, float *__restrict__ val,
                                                  const float *__restrict__ next_value, const float *__restrict__ msk,
                                                  const uint8_t *__restrict__ ends, float *__restrict__ ret, int T, int A,
                                                  int E, float g, float gt) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (e >= E) return;
    const size_t plane = (size_t)A * E, col = (size_t)a * E + e;
    float v1 = next_value[col];                    // value_preds[T] = next_value (storage.py:61)
    val[(size_t)T * plane + col] = v1;
    float m1 = msk[(size_t)T * plane + col];
    bool end1 = ends[(size_t)T * E + e] != 0;
    float gae = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
        const size_t k = (size_t)t * plane + col;
        const float r = rew[k], v0 = val[k], m0 = msk[k];
        const bool end0 = ends[(size_t)t * E + e] != 0;
        if (end1) gae = 0.0f;                                               // a segment ends right after step t
        // delta = r + gamma * V[t+1] * m[t+1] - V[t];  gae = delta + gamma * tau * m[t+1] * gae  (no FMA contraction)
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, v1), m1)), v0);
        gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gt, m1), gae));
        if (!end0) ret[k] = __fadd_rn(gae, v0);                             // index `end` itself is skipped
        else gae = 0.0f;
        v1 = v0; m1 = m0; end1 = end0;
    }
}

// ---- minibatch gather: one thread per output row (k, j); rows of one agent are consecutive threads ---------------
struct GatherParams {
    const int64_t *idx;
    int mb, T, A, E, a0, n, o0, m;
    const float *obs, *value_preds, *returns, *masks, *old_logp, *adv;
    const int64_t *actions;
    float *obs_own, *obs_opp, *o_value_preds, *o_returns, *o_masks, *o_old_logp, *o_adv, *alive, *alive_sum;
    int64_t *o_actions;
};

__global__ void __launch_bounds__(256) gather_kernel(const GatherParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;     // sample j of team slot k
    float alive = 0.0f;
    if (j < p.mb) {
        const int64_t id = p.idx[j];
        const int t = (int)(id / p.E), e = (int)(id - (int64_t)t * p.E);
        const size_t plane = (size_t)p.A * p.E;
        if (k < p.n) {
            const size_t src = (size_t)t * plane + (size_t)(p.a0 + k) * p.E + e, dst = (size_t)k * p.mb + j;
            const float2 *o = reinterpret_cast<const float2 *>(p.obs + src * 6);
            float2 *d = reinterpret_cast<float2 *>(p.obs_own + dst * 6);
            const float2 o0 = o[0];
            d[0] = o0; d[1] = o[1]; d[2] = o[2];
            alive = o0.x;
            p.alive[dst] = alive;
            p.o_actions[dst] = p.actions[src];
            p.o_value_preds[dst] = p.value_preds[src];
            p.o_returns[dst] = p.returns[src];
            p.o_masks[dst] = p.masks[src];
            p.o_old_logp[dst] = p.old_logp[src];
            p.o_adv[dst] = p.adv[((size_t)k * p.T + t) * p.E + e];
        } else {
            const int ko = k - p.n;
            const size_t src = (size_t)t * plane + (size_t)(p.o0 + ko) * p.E + e, dst = (size_t)ko * p.mb + j;
            const float2 *o = reinterpret_cast<const float2 *>(p.obs + src * 6);
            float2 *d = reinterpret_cast<float2 *>(p.obs_opp + dst * 6);
            d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
        }
    }
    // alive count of the block -> one atomic (the loss normaliser mask.sum())
    __shared__ float red[8];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) alive += __shfl_xor_sync(0xffffffffu, alive, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = alive;
    __syncthreads();
    if (threadIdx.x == 0 && k < p.n) {
        float t = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        if (t != 0.0f) atomicAdd(p.alive_sum, t);
    }
}

// ---- masked clipped-PPO loss + gradient -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ppo_loss_kernel(const float *__restrict__ values, const float *__restrict__ logp,
                                                       const float *__restrict__ entropy, const float *__restrict__ old_values,
                                                       const float *__restrict__ returns, const float *__restrict__ old_logp,
                                                       const float *__restrict__ adv, const float *__restrict__ mask,
                                                       const float *__restrict__ norm, int N, float clip, float vcoef, float ecoef,
                                                       float *out, float *gvalues, float *glogp, float *gentropy) {
    const float inv = 1.0f / norm[0];
    float sv = 0.0f, sa = 0.0f, se = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float m = mask[i], v = values[i], ov = old_values[i], ret = returns[i], a = adv[i];
        // entropy
        se += entropy[i] * m;
        gentropy[i] = -ecoef * m * inv;
        // surrogate
        const float ratio = m * expf(logp[i] - old_logp[i]);
        const float clamped = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip);
        const float s1 = ratio * a, s2 = clamped * a;
        sa += m * -fminf(s1, s2);
        const float d1 = ratio * a, d2 = (ratio >= 1.0f - clip && ratio <= 1.0f + clip) ? ratio * a : 0.0f;
        const float dmin = s1 < s2 ? d1 : (s1 > s2 ? d2 : 0.5f * (d1 + d2));
        glogp[i] = -m * dmin * inv;
        // clipped value loss
        const float dv = v - ov;
        const float vc = ov + fminf(fmaxf(dv, -clip), clip);
        const float e1 = v - ret, e2 = vc - ret;
        const float A1 = e1 * e1, A2 = e2 * e2;
        sv += 0.5f * fmaxf(A1, A2) * m;
        const float g1 = 2.0f * e1, g2 = (dv >= -clip && dv <= clip) ? 2.0f * e2 : 0.0f;
        const float dmax = A1 > A2 ? g1 : (A1 < A2 ? g2 : 0.5f * (g1 + g2));
        gvalues[i] = vcoef * 0.5f * dmax * m * inv;
    }
    __shared__ float red[3][8];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, s);
        sa += __shfl_xor_sync(0xffffffffu, sa, s);
        se += __shfl_xor_sync(0xffffffffu, se, s);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = sa; red[2][threadIdx.x >> 5] = se; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tv = 0.0f, ta = 0.0f, te = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tv += red[0][w]; ta += red[1][w]; te += red[2][w]; }
        tv *= inv; ta *= inv; te *= inv;
        atomicAdd(out + 0, tv);
        atomicAdd(out + 1, ta);
        atomicAdd(out + 2, te);
        atomicAdd(out + 3, tv * vcoef + ta - te * ecoef);
    }
}

}  // namespace rl

extern "C" int rl_gather_minibatch(const int64_t *d_idx, int mb, int T, int A, int E, int a0, int n, int o0, int m,
                                   const float *d_obs, const int64_t *d_actions, const float *d_value_preds,
                                   const float *d_returns, const float *d_masks, const float *d_old_logp, const float *d_adv,
                                   float *obs_own, float *obs_opp, int64_t *actions, float *value_preds, float *returns,
                                   float *masks, float *old_logp, float *adv, float *alive, float *d_alive_sum, void *stream) {
    if (!d_idx || !d_obs || !d_actions || !d_value_preds || !d_returns || !d_masks || !d_old_logp || !d_adv || !obs_own ||
        !obs_opp || !actions || !value_preds || !returns || !masks || !old_logp || !adv || !alive || !d_alive_sum)
        return fa_internal_fail(-1, "rl_gather_minibatch: NULL pointer");
    if (mb < 1 || T < 1 || E < 1 || n < 1 || m < 1 || a0 < 0 || o0 < 0 || a0 + n > A || o0 + m > A)
        return fa_internal_fail(-1, "rl_gather_minibatch: bad sizes");
    rl::GatherParams p;
    p.idx = d_idx; p.mb = mb; p.T = T; p.A = A; p.E = E; p.a0 = a0; p.n = n; p.o0 = o0; p.m = m;
    p.obs = d_obs; p.actions = d_actions; p.value_preds = d_value_preds; p.returns = d_returns; p.masks = d_masks;
    p.old_logp = d_old_logp; p.adv = d_adv; p.obs_own = obs_own; p.obs_opp = obs_opp; p.o_actions = actions;
    p.o_value_preds = value_preds; p.o_returns = returns; p.o_masks = masks; p.o_old_logp = old_logp; p.o_adv = adv;
    p.alive = alive; p.alive_sum = d_alive_sum;
    const dim3 grid((unsigned)((mb + 255) / 256), (unsigned)(n + m));
    rl::gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_gather_minibatch: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_ppo_loss(const float *d_values, const float *d_logp, const float *d_entropy, const float *d_old_values,
                           const float *d_returns, const float *d_old_logp, const float *d_adv, const float *d_mask,
                           const float *d_norm, int N, float clip, float vcoef, float ecoef, float *d_out, float *d_gvalues,
                           float *d_glogp, float *d_gentropy, void *stream) {
    if (!d_values || !d_logp || !d_entropy || !d_old_values || !d_returns || !d_old_logp || !d_adv || !d_mask || !d_norm ||
        !d_out || !d_gvalues || !d_glogp || !d_gentropy)
        return fa_internal_fail(-1, "rl_ppo_loss: NULL pointer");
    if (N < 1) return fa_internal_fail(-1, "rl_ppo_loss: N must be >= 1");
    int blocks = (N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rl::ppo_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_values, d_logp, d_entropy, d_old_values, d_returns,
                                                                  d_old_logp, d_adv, d_mask, d_norm, N, clip, vcoef, ecoef,
                                                                  d_out, d_gvalues, d_glogp, d_gentropy);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_ppo_loss: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_gae(const float *d_rewards, float *d_value_preds, const float *d_next_value, const float *d_masks,
                      const uint8_t *d_ends, float *d_returns, int T, int A, int E, double gamma, double tau, void *stream) {
    if (!d_rewards || !d_value_preds || !d_next_value || !d_masks || !d_ends || !d_returns)
        return fa_internal_fail(-1, "rl_gae: NULL pointer");
    if (T < 1 || A < 1 || E < 1 || A > 65535) return fa_internal_fail(-1, "rl_gae: bad sizes T=%d A=%d E=%d", T, A, E);
    const dim3 grid((unsigned)((E + 127) / 128), (unsigned)A);
    rl::gae_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_rewards, d_value_preds, d_next_value, d_masks, d_ends, d_returns,
                                                           T, A, E, (float)gamma, (float)(gamma * tau));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_gae: launch: %s", cudaGetErrorString(e));
    return 0;
}
