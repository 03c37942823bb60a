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

/* Per-step bookkeeping of the rollout loop (train_fortattack.py:53,97-104; RolloutStorage.insert, storage.py:41) for every
 * agent and env in one launch: with alive_before = obs_t[a][e][0] and fin = done[e] != 0,
 *   masks_t1[a][e] = fin ? obs_t1[a][e][0] (the new episode's alive flags) : alive_before
 *   ends_t1[e]     = fin                       (the end point train_fortattack.py:98 appends)
 *   episode_rewards[a][e] += reward[a][e] * alive_before
 * obs_t / obs_t1 float [A][E][6]; done uint8 [E]; reward, masks_t1, episode_rewards float [A][E]; ends_t1 uint8 [E]. */
int rl_rollout_bookkeeping(const float *d_obs_t, const float *d_obs_t1, const uint8_t *d_done, const float *d_reward,
                           float *d_masks_t1, uint8_t *d_ends_t1, float *d_episode_rewards, int A, int E, void *stream);

/* Team minibatch gather (replaces magent_feed_forward_generator's per-agent index + cat,
 * rlcore/algo/ppo.py:207-246): for mb sample indices idx[j] = t * E + e and the team's agents a0 .. a0+n-1 (opponents
 * o0 .. o0+m-1), rows are emitted agent-major (row = k * mb + j), exactly the order torch.cat([x[i][idx] ...]) gives.
 *   inputs  (shared rollout blocks, SharedRollouts): d_obs float [T+1][A][E][6]; d_actions int64 [T][A][E];
 *           d_value_preds / d_returns / d_masks float [T+1][A][E]; d_old_logp float [T][A][E]; d_adv float [n][T][E]
 *   outputs: obs_own float [n*mb][6], obs_opp float [m*mb][6], actions int64 [n*mb], value_preds, returns, masks,
 *            old_logp, adv, alive float [n*mb] (alive = obs_own[:, 0], the loss mask of ppo.py:224), and
 *            d_alive_sum float [1] += sum(alive)  (zero it first; it is the loss normaliser) */
int rl_gather_minibatch(const int64_t *d_idx, int mb, int T, int A, int E, int a0, int n, int o0, int m,
                        const float *d_obs, const int64_t *d_actions, const float *d_value_preds, const float *d_returns,
                        const float *d_masks, const float *d_old_logp, const float *d_adv, float *obs_own, float *obs_opp,
                        int64_t *actions, float *value_preds, float *returns, float *masks, float *old_logp, float *adv,
                        float *alive, float *d_alive_sum, void *stream);

/* Masked clipped-PPO loss, forward and gradient in one pass (rlcore/algo/ppo.py:150-187 with
 * use_clipped_value_loss=True):  with S = *d_norm (sum of the alive mask, or N if that is 0; made global by the caller
 * in multi-rank runs),
 *   entropy     = sum(ent * mask) / S
 *   action_loss = sum(mask * -min(r * adv, clamp(r, 1-c, 1+c) * adv)) / S,   r = mask * exp(logp - old_logp)
 *   value_loss  = sum(mask * 0.5 * max((v - ret)^2, (old_v + clamp(v - old_v, -c, c) - ret)^2)) / S
 *   total       = value_loss * vcoef + action_loss - entropy * ecoef
 * d_out float [4] += {value_loss, action_loss, entropy, total} (zero it first); d_gvalues / d_glogp / d_gentropy
 * float [N] receive d total / d input (torch's tie rule for min/max: the gradient is split evenly). */
int rl_ppo_loss(const float *d_values, const float *d_logp, const float *d_entropy, const float *d_old_values,
                const float *d_returns, const float *d_old_logp, const float *d_adv, const float *d_mask,
                const float *d_norm, int N, float clip, float vcoef, float ecoef, float *d_out, float *d_gvalues,
                float *d_glogp, float *d_gentropy, void *stream);

/* The same loss taken from the action head's LOGITS [N][8] and the actions int64 [N]: the Categorical's log-prob and
 * entropy (rlcore/distributions.py:9-17 as used by mpnn.py:199-200) and their gradients are evaluated inside, so the ~50
 * elementwise launches of log_softmax / gather / entropy and their autograd backward disappear from the optimizer step:
 *   logp_i = l_{i,a_i} - logsumexp_j l_ij,  H_i = -sum_j p_ij log p_ij,
 *   d total / d l_ij = glogp_i ([j == a_i] - p_ij) - gentropy_i p_ij (log p_ij + H_i).
 * d_out float [4] is OVERWRITTEN with {value_loss, action_loss, entropy, total}: per-block partial sums are added in block
 * order by the last block (d_scratch: rl_ppo_loss_logits_scratch_floats() floats, its first word zero before the first
 * call; the kernel leaves it zero), so the statistics are bit-reproducible.  d_logp / d_entropy (optional) receive
 * the per-row values.  d_out == NULL: evaluation only (d_logp and / or d_entropy), nothing else is read. */
size_t rl_ppo_loss_logits_scratch_floats(void);
int rl_ppo_loss_logits(const float *d_values, const float *d_logits, const int64_t *d_actions, const float *d_old_values,
                       const float *d_returns, const float *d_old_logp, const float *d_adv, const float *d_mask,
                       const float *d_norm, int N, int n_actions, float clip, float vcoef, float ecoef, float *d_out,
                       float *d_gvalues, float *d_glogits, float *d_logp, float *d_entropy, float *d_scratch, void *stream);

/* Single-head attention over a handful of agents, forward and backward, for the TRAINING forward of the MPNN
 * (mpnn.py:249-331 MultiHeadAttention without self-messages, mpnn.py:376-443 MultiHeadOppAttention): replaces
 * bmm -> mask -> softmax -> bmm (and their four backward bmm's) over [batch, <=5, <=5] matrices by one warp per
 * environment.  For every batch element b:   s_ij = norm * <A_i, B_j>  (s_ii = -inf if mask_diag),
 * p_i. = softmax_j s_ij (all-masked rows give p = 0),  out_i = sum_j p_ij V_j.
 * Operands are addressed as  X[b][i][c] = X + b * batch_stride + i * row_stride + c  (element strides), so the
 * agent-major activations of the reference ([n * batch, d] rows) and packed Q|K|V products are read in place.
 *   n rows of A / out, m rows of B / V, feature size k in {32, 64, 96, 128}; d_attn float [batch][n][m]. */
typedef struct RlAttnOperand {
    float *ptr;              /* first element */
    int64_t batch_stride;    /* elements between consecutive batch entries */
    int64_t row_stride;      /* elements between consecutive agents */
} RlAttnOperand;

int rl_attn_forward(const RlAttnOperand *A, const RlAttnOperand *B, const RlAttnOperand *V, const RlAttnOperand *out,
                    float *d_attn, int batch, int n, int m, int k, float norm, int mask_diag, void *stream);

/* Gradients of the same op: given d(out), the saved attention matrix and the forward operands, writes dA, dB, dV
 * (each with its own strides; they may alias slices of one packed gradient tensor). */
int rl_attn_backward(const RlAttnOperand *dout, const RlAttnOperand *A, const RlAttnOperand *B, const RlAttnOperand *V,
                     const float *d_attn, const RlAttnOperand *dA, const RlAttnOperand *dB, const RlAttnOperand *dV,
                     int batch, int n, int m, int k, float norm, void *stream);

/* The same attention with keys == values == X (the "folded" training forward: with G = H (W_query W_key^T) the
 * scores of mpnn.py:289-295 are <G_i, H_j>, and sum_j p_ij (H_j W_val) W_out = (sum_j p_ij H_j) (W_val W_out), so the
 * kernel mixes the hidden rows themselves and the projections collapse into two [d, d] products on the weights).
 *   out_i = sum_j softmax_j(norm <A_i, X_j>) X_j;  x_copy (may be NULL) receives a copy of the X rows -- with `out`
 *   pointing at the other half of the same [rows, 2d] buffer this builds the input cat(h, m) of update (mpnn.py:159)
 *   without a cat kernel.
 * Backward: dA as rl_attn_backward; dX = dB + dV (+ dx_add, may be NULL: a gradient that reaches X by another path). */
int rl_attn_mix_forward(const RlAttnOperand *A, const RlAttnOperand *X, const RlAttnOperand *out, const RlAttnOperand *x_copy,
                        float *d_attn, int batch, int n, int m, int k, float norm, int mask_diag, void *stream);
int rl_attn_mix_backward(const RlAttnOperand *dout, const RlAttnOperand *A, const RlAttnOperand *X, const float *d_attn,
                         const RlAttnOperand *dA, const RlAttnOperand *dX, const RlAttnOperand *dx_add, int batch, int n,
                         int m, int k, float norm, void *stream);

/* Backward of y = relu(x W^T + b) up to the GEMMs (every Linear + ReLU of mpnn.py:37-58 in JointPPO.update's backward,
 * ppo.py:189-190): d_dpre = d_dout * [d_out > 0] and the bias gradient in the same pass.  Row-major [rows, cols] floats,
 * cols in {32, 64, 128, 256}; d_partial float [rl_relu_bwd_colsum_blocks(rows, cols)][cols] receives one partial column
 * sum per block in a fixed order (the bias gradient is their sum over dim 0 -- reproducible, no atomics).  Replaces
 * threshold_backward + a column reduction that ran at 12 % of the HBM bandwidth. */
int rl_relu_bwd_colsum_blocks(long long rows, int cols);     /* number of partial rows, or -1 for unsupported sizes */
int rl_relu_bwd_colsum(const float *d_dout, const float *d_out, float *d_dpre, float *d_partial, long long rows, int cols,
                       void *stream);
/* Up to RL_SMALL_MATMUL_MAX small fp32 products in ONE launch:  C (+)= op(A) op(B),  op(X) = X or X^T, every dimension <= 256,
 * row-major with row strides.  The folds of the attention projections into [d, d] weights (W_query W_key^T, W_val W_out U2^T:
 * rlcore/fused.py fold_weights, mpnn.py:286-327 folded) and their backward are six products of ~2 MFLOP per optimizer step;
 * the items of one call must not depend on each other.  Plain fp32 FMAs in a fixed order (bit-reproducible). */
#define RL_SMALL_MATMUL_MAX 8
typedef struct RlSmallMatmul {
    const float *A, *B;
    float *C;
    int32_t M, N, K;                  /* C is M x N, the reduction runs over K */
    int32_t lda, ldb, ldc;            /* row strides (floats) of A, B, C as stored */
    int32_t trans_a, trans_b;         /* op(A) = A^T (A stored K x M) / op(B) = B^T (B stored N x K) */
    int32_t accumulate;               /* C += instead of C = */
} RlSmallMatmul;
int rl_small_matmul(const RlSmallMatmul *items, int n_items, void *stream);

/* out[c] = sum_r x[r][c] for a tall row-major matrix (cols a power of two <= 256, row stride ld): the bias gradients of the
 * dense layers (rows of dpre, or relu_bwd_colsum's per-block partials) in one launch, partial sums added in a fixed order
 * (bit-reproducible).  d_scratch: 4 + rl_colsum_blocks(rows, cols) * cols floats, first word zero before the first call. */
int rl_colsum_blocks(long long rows, int cols);
int rl_colsum(const float *d_x, long long rows, int cols, int ld, float *d_out, float *d_scratch, void *stream);

/* The same on row-major operands with row strides (floats, multiples of 4): column slices of wider matrices (the encoder half
 * of the [h0 | eOpp] feature block) are masked in place of a contiguous copy. */
int rl_relu_bwd_colsum_ld(const float *d_dout, int ldd, const float *d_out, int ldo, float *d_dpre, int ldp, float *d_partial,
                          long long rows, int cols, void *stream);

#ifdef __cplusplus
}
#endif
#endif
