/*
 * fortattack.h -- C ABI of the B200-native batched FortAttack simulator (libfortattack_b200.so).
 *
 * The reference (Ankur-Deka/Emergent-Multiagent-Strategies) has no FFI: its "plugin" surface for
 * the step path is the Python seam
 *     make_fortattack_env / FortAttackGlobalEnv.reset / .step      gym_fortattack/fortattack.py:17-27,127-186
 *     World.step                                                   gym_fortattack/core.py:191-218
 *     scenario callbacks reset_world / reward / observation        gym_fortattack/envs/fortattack_env_v1.py:47-238
 * Each entry point below names the reference interface it replaces.  The Python side binds these
 * with ctypes (emergent-multiagent-strategies_b200/_capi.py; the stub a reference maintainer would
 * add is shown in INTEGRATION.md).
 *
 * Conventions
 *  - plain C types only; every `d_*` pointer is a DEVICE pointer that the caller owns (a PyTorch
 *    tensor's data_ptr()) and that must stay alive until the stream has run the call; `h_*`
 *    pointers are HOST pointers (page-locked memory gives asynchronous copies).
 *  - nothing is allocated or freed on the device by this library: the caller supplies one
 *    workspace of fa_workspace_bytes() bytes at fa_create().
 *  - every function returns 0 on success or a negative FA_E* code; fa_last_error() returns the
 *    calling thread's message for the last failure.  Work is enqueued on `stream` (a cudaStream_t
 *    passed as void*; NULL = legacy default stream) and is asynchronous with respect to the host,
 *    except fa_step_host which returns after its results are in host memory.
 *  - one host thread per handle at a time; different handles are independent.
 *
 * Layouts (E = n_envs, A = n_guards + n_attackers, guards first; "Real" = float or double as
 * chosen by FaConfig.scalar):
 *    actions  int32 [A][E]        0 none, 1 +x, 2 -x, 3 +y, 4 -y, 5 +rot, 6 -rot, 7 shoot
 *                                 (fortattack.py:253-263); values outside 0..7 act as 0
 *    obs      Real  [A][E][6]     alive, x, y, ang, vx, vy   (fortattack_env_v1.py:238)
 *    reward   Real  [A][E]
 *    done     uint8 [E]
 *    result   uint8 [E]           0 running, 1 all attackers dead, 2 time limit, 3 attacker reached the
 *                                 fort (world.gameResult[0], [1], [2]; fortattack.py:202-225)
 * Agent-major planes are what the reference's callers index (`obs[i]`, `reward[i]`, learner.py:239-243,
 * rlcore/storage.py:33-43) and what makes every per-env access of the step kernel coalesced.
 */
#ifndef FORTATTACK_B200_H
#define FORTATTACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_ABI_VERSION 1

enum {
    FA_OK = 0,
    FA_EINVAL = -1,      /* bad argument (NULL pointer, size out of range, unsupported team size) */
    FA_ECUDA = -2,       /* a CUDA runtime call failed; see fa_last_error() */
    FA_ENODEVICE = -3,   /* no CUDA device / wrong architecture (the library only carries sm_100a code) */
    FA_EALIGN = -4       /* a pointer is not aligned as required (16 bytes for obs and the workspace) */
};

enum { FA_F32 = 0, FA_F64 = 1 };

/* Thread mapping of the step kernel: one thread per env (large batches, register-resident agent block), one thread per
 * agent with one warp per agent index (shared-memory agent block, block barriers), or one sub-warp group of 2/4/8/16
 * lanes per env (small batches: agent block staged per warp, hit / alive / fort-reached resolution by warp ballots, no
 * block barrier); AUTO picks by batch size (GROUP, then AGENT, then ENV as the batch grows). */
enum { FA_MAP_AUTO = 0, FA_MAP_ENV = 1, FA_MAP_AGENT = 2, FA_MAP_GROUP = 3 };

#define FA_MAX_TEAM 5    /* kernels are instantiated for 1..5 guards x 1..5 attackers */

typedef struct FaHandle FaHandle;

/* Scenario constants are the reference's (core.py:32,100-101,115-128; fortattack_env_v1.py:16-35)
 * and are compiled in; what the reference hard-codes but a batched engine must vary is here. */
typedef struct FaConfig {
    int32_t n_envs;        /* E >= 1 */
    int32_t n_guards;      /* reference: 5 (fortattack_env_v1.py:18) */
    int32_t n_attackers;   /* reference: 5 (fortattack_env_v1.py:19) */
    int32_t max_steps;     /* world.max_time_steps, episode cap (fortattack.py:21) */
    int32_t scalar;        /* FA_F32 (production) or FA_F64 (parity mode, same kernels in double) */
    int32_t device;        /* CUDA device ordinal */
    int32_t mapping;       /* FA_MAP_AUTO / FA_MAP_ENV / FA_MAP_AGENT / FA_MAP_GROUP */
    int32_t reserved;      /* 0 */
    uint64_t seed;         /* Philox key of the reset streams */
    uint64_t env_id0;      /* global id of env 0 of this shard: resets are keyed by (seed, env_id0+e, episode) */
} FaConfig;

/* Canonical, layout-independent view of the full simulator state for fa_get_state / fa_set_state
 * (teacher-forced parity tests, checkpointing).  Env-major, float64, same as the oracle's:
 *   st_f double [E][A][6]  x, y, vx, vy, ang, prevDist (NaN = None, core.py:104)
 *   st_i uint8  [E][A][6]  alive, justDied, hit, wasHit, numHit, numWasHit (core.py:90-96,103)
 *   time_step int32 [E]    world.time_step
 *   episode  uint32 [E]    number of resets so far (selects the reset stream)
 * All four are device pointers. */
typedef struct FaState {
    double *d_st_f;
    uint8_t *d_st_i;
    int32_t *d_time_step;
    uint32_t *d_episode;
} FaState;

/* ABI version of the loaded library (== FA_ABI_VERSION of the header it was built from). */
int fa_abi_version(void);

/* Message for the last failure on the calling thread ("" if none). */
const char *fa_last_error(void);

/* Bytes of device workspace fa_create needs for this configuration. */
int fa_workspace_bytes(const FaConfig *cfg, size_t *out_bytes);

/* Replaces: make_fortattack_env(num_steps) + FortAttackEnvV1.__init__ + World.__init__
 * (fortattack.py:17-27, fortattack_env_v1.py:10-45, core.py:109-128) for E envs at once.
 * d_workspace: fa_workspace_bytes() bytes, 256-byte aligned, owned by the caller.  Every env starts
 * un-reset (agents at the origin, alive, prevDist = None as in core.py:104, episode 0): call fa_reset before
 * the first step, as the reference's callers do (train_fortattack.py:29). */
int fa_create(const FaConfig *cfg, void *d_workspace, FaHandle **out);
int fa_destroy(FaHandle *h);

/* Replaces: FortAttackGlobalEnv.reset -> FortAttackEnvV1.reset_world (fortattack.py:175-186,
 * fortattack_env_v1.py:47-75).  d_env_mask: uint8 [E], only envs with mask != 0 are reset (NULL = all).
 * d_obs (may be NULL): Real [A][E][6], receives the current observation of EVERY env. */
int fa_reset(FaHandle *h, const uint8_t *d_env_mask, void *d_obs, void *stream);

/* Replaces: FortAttackGlobalEnv.step (fortattack.py:127-173): _set_action for every agent,
 * World.step (laser, action/contact/wall force, integrate), observation + reward per agent,
 * _get_done, time_step += 1 -- one fused kernel launch.
 * auto_reset != 0 folds the caller's `if done: obs = env.reset()` (train_fortattack.py:97-104) into
 * the same launch: finished envs are reset in place and their obs rows hold the first observation
 * of the new episode; reward/done/result always describe the step just taken.
 * d_done / d_result may be NULL. */
int fa_step(FaHandle *h, const int32_t *d_actions, void *d_obs, void *d_reward, uint8_t *d_done,
            uint8_t *d_result, int auto_reset, void *stream);

/* T consecutive steps in ONE persistent launch (state stays in registers; auto-reset always on).
 * d_actions int32 [T][A][E]; outputs are streams obs Real [T][A][E][6], reward Real [T][A][E],
 * done uint8 [T][E], result uint8 [T][E]; any output may be NULL to skip storing it. */
int fa_step_many(FaHandle *h, int T, const int32_t *d_actions, void *d_obs, void *d_reward,
                 uint8_t *d_done, uint8_t *d_result, void *stream);

/* Same as fa_step with HOST buffers; returns when the results are in host memory.  This is the call the
 * numpy-facing env.step() of the Python facade makes.
 *  - all buffers page-locked (cudaHostAlloc / cudaHostRegister / pinned torch tensors): the kernel reads the
 *    actions and writes its results directly through the mapped host addresses (one launch, no copies);
 *  - otherwise: actions are copied host->device, results device->host, through staging buffers in the
 *    workspace; buffers laid out as fa_host_layout() describes come back in a single copy.
 * Environment FA_HOST_PATH=staged|mapped forces one path (read at fa_create). */
int fa_step_host(FaHandle *h, const int32_t *h_actions, void *h_obs, void *h_reward, uint8_t *h_done,
                 uint8_t *h_result, int auto_reset, void *stream);

/* Byte offsets of reward / done / result relative to obs, and the total size, of the packed host result
 * block that fa_step_host's staged path returns in one copy. */
int fa_host_layout(const FaHandle *h, size_t *off_reward, size_t *off_done, size_t *off_result, size_t *total);

/* T consecutive steps with HOST streams: h_actions int32 [T][A][E] in, h_obs Real [T][A][E][6], h_reward Real [T][A][E],
 * h_done / h_result uint8 [T][E] out (any output may be NULL); auto-reset always on; returns when the results are in
 * host memory.  This is the reference's `for step in range(T): obs, reward, done, info = env.step(actions[step])`
 * (train_fortattack.py:51-104 with a pre-drawn action stream, the synthetic-action workload of BASELINE configs[1])
 * for all E envs, with the host<->device traffic of step t overlapped with the arithmetic of its neighbours:
 *  - d_stage != NULL: a device staging buffer of stage_bytes (256-byte aligned, caller-owned; size it with
 *    fa_host_stage_bytes for the chunk length you want, any size >= one-step chunks works and the chunk length is
 *    derived from it).  Chunks of c steps run as a three-stage pipeline on two internal copy streams and `stream`:
 *    actions of chunk k+1 host->device | one persistent fa_step_many launch for chunk k | results of chunk k-1
 *    device->host.  Page-locked host buffers make the copies asynchronous; pageable ones still give correct results.
 *  - d_stage == NULL: all host buffers must be page-locked; ONE persistent launch reads the action stream and writes
 *    the result streams through the mapped host addresses (no copy calls).
 * Results are bit-identical to T calls of fa_step_host (and to fa_step_many on device buffers). */
int fa_step_many_host(FaHandle *h, int T, const int32_t *h_actions, void *h_obs, void *h_reward, uint8_t *h_done,
                      uint8_t *h_result, void *d_stage, size_t stage_bytes, void *stream);

/* Bytes of staging buffer for fa_step_many_host chunks of chunk_steps steps (two chunk buffers). */
int fa_host_stage_bytes(const FaHandle *h, int chunk_steps, size_t *out_bytes);

/* Full state exchange in the canonical layout (device pointers).  fa_set_state accepts any state
 * the reference can be in; FA_F32 handles store ang as (ang mod 2pi, turn count < 65536). */
int fa_get_state(FaHandle *h, const FaState *out, void *stream);
int fa_set_state(FaHandle *h, const FaState *in, void *stream);

/* world.numAliveGuards / world.numAliveAttackers (core.py:113-114): int32 [2][E], device. */
int fa_alive_counts(FaHandle *h, int32_t *d_counts, void *stream);

/* Change world.max_time_steps for subsequent steps (the reference's scripts assign it directly). */
int fa_set_max_steps(FaHandle *h, int32_t max_steps);

/* Optional extra output of every subsequent step: d_alive_end uint8 [E] (fa_step) or [T][E] (fa_step_many) receives
 * numAliveGuards | numAliveAttackers << 4 as the step leaves them, i.e. before an auto-reset -- what the reference's
 * evaluation script reads off the world when an episode ends (test_fortattack_v2.py:98-99, core.py:113-114).
 * NULL (the default) switches it off. */
int fa_set_alive_end_buffer(FaHandle *h, uint8_t *d_alive_end);

/* Optional rollout bookkeeping written by every subsequent fa_step / fa_step_many (device calls; the host-buffer calls
 * ignore it), replacing the per-step glue of the reference's collection loop (train_fortattack.py:53,97-104;
 * RolloutStorage.insert, rlcore/storage.py:41).  Each pointer may be NULL; all NULL (the default) switches it off.
 *   d_mask_next float [A][E] ([T][A][E] for fa_step_many): done ? alive flag of the new (reset) observation
 *                                                                : alive flag before the step
 *   d_end_next  uint8 [E]    ([T][E]):                     done
 *   d_ep_reward float [A][E] (accumulated in place):       += reward * alive flag before the step */
int fa_set_rollout_outputs(FaHandle *h, float *d_mask_next, uint8_t *d_end_next, float *d_ep_reward);

/* Number of kernel launches this handle has enqueued so far (bench.py's gpu_launches). */
int fa_launch_count(const FaHandle *h, uint64_t *out);

/* Static facts about the step kernel chosen for this handle (for DESIGN.md / bench.py):
 * registers per thread, threads per block, blocks per launch, static shared memory bytes, and the mapping in
 * use (FA_MAP_ENV, FA_MAP_AGENT or FA_MAP_GROUP). */
int fa_kernel_info(const FaHandle *h, int32_t *regs, int32_t *block, int32_t *grid, int32_t *smem, int32_t *mapping);

#ifdef __cplusplus
}
#endif
#endif /* FORTATTACK_B200_H */
