/*
 * fortattack_rollout.h -- C ABI of the rollout-storage kernels (libfortattack_b200.so).
 *
 * Replaces the per-agent, per-segment Python loops of the reference's return computation
 *     RolloutStorage.compute_returns (use_gae=True)        rlcore/storage.py:59-66
 *     Learner.wrap_horizon                                 learner.py:191-211
 *     Neo.wrap_horizon                                     rlagent.py:41-42
 * by ONE launch over every (agent, env) column of the shared rollout blocks.  Conventions as in
 * fortattack.h (plain C types, caller-owned device memory, explicit stream, 0 / negative error code).
 */
#ifndef FORTATTACK_ROLLOUT_B200_H
#define FORTATTACK_ROLLOUT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Segment-wise GAE for A agents x E envs whose episodes end at different steps.
 *   d_rewards      float [T][A][E]
 *   d_value_preds  float [T+1][A][E]   slot T receives d_next_value (storage.py:61)
 *   d_next_value   float [A][E]        V(obs[T])
 *   d_masks        float [T+1][A][E]
 *   d_ends         uint8 [T+1][E]      ends[t][e] != 0: t is an end point of env e, exactly the reference's
 *                                      `end_pts` list (train_fortattack.py:97-109: the step count right after a
 *                                      done, and always T)
 *   d_returns      float [T+1][A][E]   written for every t < T that is not an end point (learner.py:205 starts the
 *                                      next segment at end + 1, so index `end` keeps what it held)
 * Per column this is, for each segment [start, end):  gae = 0;  for t = end-1 .. start:
 *     delta = r[t] + gamma * V[t+1] * m[t+1] - V[t];  gae = delta + gamma * tau * m[t+1] * gae;  ret[t] = gae + V[t]
 * evaluated in float32 with the reference's operation order (bit-equal to the torch expression). */
int rl_gae(const float *d_rewards, float *d_value_preds, const float *d_next_value, const float *d_masks,
           const uint8_t *d_ends, float *d_returns, int T, int A, int E, double gamma, double tau, void *stream);

#ifdef __cplusplus
}
#endif
#endif
