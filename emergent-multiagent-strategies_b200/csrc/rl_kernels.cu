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

}  // namespace rl

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
