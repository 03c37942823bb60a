/*
 * fortattack_policy.h -- C ABI of the fused MPNN rollout forward (libfortattack_b200.so).
 *
 * Replaces, for rollouts (no autograd), the reference's team policy forward
 *     MPNN._fwd / act / get_value                   mpnn.py:117-205
 *     MultiHeadOppAttention.forward                 mpnn.py:376-443
 *     MultiHeadAttention.forward                    mpnn.py:249-331
 *     Categorical / FixedCategorical                rlcore/distributions.py:9-32
 * as called by Learner.act and Learner.wrap_horizon (learner.py:143-172,191-211): one launch per team
 * computes value, sampled (or arg-max) action and its log-probability for every (agent, env) row.
 *
 * The network is the reference's with hidden_dim = 128, input_size = 6, 8 actions, one attention head,
 * policy_layers = 1, three message rounds (mpnn.py:25-84; learner.py:57-69).  Dense layers run on the
 * 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulation in tensor memory); the
 * two input encoders, both attention soft-maxes, the value / action heads and the sampling run in fp32
 * on the CUDA cores of the same kernel.  The attention projections are folded on the host (Wq Wk^T,
 * Wv Wout U2^T: see csrc/mp_policy.cu), which changes rounding but not the function computed.
 * Weights come as ONE packed blob that the host side builds from MPNN.state_dict() (emergent-multiagent-strategies_b200/policy_kernel.py: pack_mpnn()).
 *
 * Conventions as in fortattack.h: plain C types, caller-owned device memory, explicit stream, 0 or a
 * negative error code, message through fa_last_error().
 */
#ifndef FORTATTACK_POLICY_B200_H
#define FORTATTACK_POLICY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_HIDDEN 128
#define MP_OBS_DIM 6
#define MP_ACTIONS 8
#define MP_MAX_TEAM 5
#define MP_MAX_ENSEMBLE 8

/* packed weight blob: MP_BLOB_F16_BYTES of fp16 GEMM operands (UMMA canonical K-major core-matrix
 * order, in the order the kernel consumes them) followed by MP_BLOB_CONST_FLOATS fp32 values
 * (encoders, biases, value / action head).  Layout: csrc/mp_policy.cu "blob layout". */
#define MP_BLOB_F16_BYTES 180224
#define MP_BLOB_CONST_FLOATS 2576
#define MP_BLOB_BYTES (MP_BLOB_F16_BYTES + 4 * MP_BLOB_CONST_FLOATS)

enum {
    MP_MODE_SAMPLE = 0,   /* action ~ Categorical(logits)      (mpnn.py:186, dist.sample())     */
    MP_MODE_ARGMAX = 1,   /* action = argmax                   (mpnn.py:184, dist.mode())       */
    MP_MODE_EVAL = 2      /* action given; log-prob and entropy of it (mpnn.py:190-196)          */
};

/* One team forward over E environments.
 *   d_blob      packed weights, 16-byte aligned, MP_BLOB_BYTES
 *   d_obs_own   float [n_own][E][6]  observations of the team's agents (agent-major planes, the step kernel's layout)
 *   d_obs_opp   float [n_opp][E][6]  observations of the opposing team
 *   mode        MP_MODE_*
 *   seed, offset  Philox key / call counter for MP_MODE_SAMPLE (row r of call `offset` always draws the same number)
 *   d_counter   optional uint64[2] on the device {call counter, 0}: when given it replaces `offset` and the kernel
 *               advances it by one per launch, so launches replayed from a CUDA graph keep drawing fresh numbers
 *   env_id0     global id of env 0 of this shard (keeps samples identical under any sharding)
 *   d_action_in int64 [n_own][E], MP_MODE_EVAL only
 * Outputs (any may be NULL): d_value float [n_own][E]; d_action int64 [n_own][E]; d_action_i32 int32 [n_own][E]
 * (what fa_step reads); d_logp float [n_own][E]; d_entropy float [n_own][E]; d_logits float [n_own][E][8].
 * d_env_sel (optional int32 [E]) / sel_value: outputs are written only for environments with d_env_sel[e] == sel_value.
 * This is how an ensemble of K frozen opponent checkpoints is played with a per-environment, per-episode choice
 * (Learner.sample_attacker / select_attacker, learner.py:119-140; train_fortattack_v2.py:34-35,110-111): one launch per
 * checkpoint over all environments, each writing only the rows of the environments currently assigned to it.
 * d_env_order (int32 [E]) / d_env_offsets (int32 [K+1]), both on the device and optional: the compacted form of the same
 * thing -- the launch covers only the environments d_env_order[d_env_offsets[sel_value] .. d_env_offsets[sel_value+1]),
 * e.g. argsort of the per-environment checkpoint index and the running sum of its histogram.  The kernel reads the list
 * bounds itself, so the host needs no synchronisation and the launches can be replayed from a CUDA graph.
 * d_status: uint32 on the device, set non-zero if the kernel's internal pipeline timed out (a bug, never expected). */
int mp_forward(const void *d_blob, const float *d_obs_own, const float *d_obs_opp, int n_own, int n_opp, int n_envs,
               int mode, uint64_t seed, uint64_t offset, uint64_t *d_counter, uint64_t env_id0, const int64_t *d_action_in, float *d_value,
               int64_t *d_action, int32_t *d_action_i32, float *d_logp, float *d_entropy, float *d_logits,
               const int32_t *d_env_sel, int32_t sel_value, const int32_t *d_env_order, const int32_t *d_env_offsets,
               uint32_t *d_status, void *stream);

/* The same forward for an ENSEMBLE of n_ckpt (<= MP_MAX_ENSEMBLE) frozen checkpoints in ONE launch: CTA c keeps the weights
 * of checkpoint c % n_ckpt resident and works through the environments d_env_order[d_env_offsets[k] .. d_env_offsets[k+1])
 * of its checkpoint k (Learner.sample_attacker, learner.py:119-130: every environment plays the checkpoint it drew at its
 * episode start).  d_blobs: host array of n_ckpt device pointers.  MP_MODE_SAMPLE or MP_MODE_ARGMAX. */
int mp_forward_ensemble(const void *const *d_blobs, int n_ckpt, const float *d_obs_own, const float *d_obs_opp, int n_own,
                        int n_opp, int n_envs, int mode, uint64_t seed, uint64_t offset, uint64_t *d_counter,
                        uint64_t env_id0, float *d_value, int64_t *d_action, int32_t *d_action_i32, float *d_logp,
                        const int32_t *d_env_order, const int32_t *d_env_offsets, uint32_t *d_status, void *stream);

/* Static facts about the policy kernel: registers/thread, threads/block, dynamic shared memory bytes,
 * resident blocks per SM, environments per 128-row tile for this team shape. */
int mp_kernel_info(int n_own, int n_opp, int32_t *regs, int32_t *block, int32_t *smem, int32_t *blocks_per_sm,
                   int32_t *envs_per_tile);

#ifdef __cplusplus
}
#endif
#endif
