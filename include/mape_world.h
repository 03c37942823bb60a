/*
 * mape_world.h -- C ABI of the batched GENERIC particle world (libfortattack_b200.so).
 *
 * Replaces World.step of the reference's generic multi-agent particle environment
 *     apply_action_force, apply_environment_force / get_collision_force, apply_wall_collision_force /
 *     get_wall_collision_force, integrate_state                        multiagent/core.py:118-225
 * for E independent worlds that share one entity configuration (agents first, then landmarks: per-entity size,
 * mass, max_speed, collide, movable), e.g. the worlds of multiagent/scenarios/  (simple_spread: 3 agents + 3
 * non-colliding landmarks; simple_tag: 4 agents + 2 obstacles).  Scenario callbacks (reward, observation, reset) stay
 * with the caller: they are ordinary tensor expressions over the position / velocity planes this call updates in place.
 * Conventions as in fortattack.h.
 */
#ifndef MAPE_WORLD_B200_H
#define MAPE_WORLD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MW_MAX_ENTITIES 12

typedef struct MwConfig {
    int32_t n_envs;
    int32_t n_agents;                      /* the first n_agents entities take actions and feel the walls */
    int32_t n_entities;                    /* agents + landmarks, <= MW_MAX_ENTITIES */
    int32_t scalar;                        /* FA_F32 = 0 / FA_F64 = 1 */
    double dt, damping, contact_force, contact_margin;   /* World.__init__: 0.1, 0.25, 1e2, 1e-10 (core.py:95-105) */
    double wall[4];                        /* xmin, xmax, ymin, ymax */
    double size[MW_MAX_ENTITIES], mass[MW_MAX_ENTITIES];
    double max_speed[MW_MAX_ENTITIES];     /* < 0: None */
    uint8_t collide[MW_MAX_ENTITIES], movable[MW_MAX_ENTITIES];
} MwConfig;

/* One World.step for every env.  Entity-major planes (coalesced per-env access):
 *   d_pos, d_vel  Real [n_entities][E][2]   updated in place
 *   d_u           Real [n_agents][E][2]     agent.action.u, already scaled by the caller (environment.py _set_action) */
int mw_step(const MwConfig *cfg, void *d_pos, void *d_vel, const void *d_u, void *stream);

/* Scenario callbacks of two of the reference's scenarios, evaluated for every agent of every world the way
 * MultiAgentEnv.step does after world.step() (multiagent/environment.py:97-108):
 *   MW_SIMPLE_SPREAD  observation = [vel, pos, landmarks - pos, other agents - pos, other agents' comm (zeros, dim_c 2)],
 *                     reward = -sum_l min_a |a - l| - #{a : |a - i| < size_a + size_i} (the agent itself included), then
 *                     shared: every agent receives the sum over agents (world.collaborative)
 *                                                                        multiagent/scenarios/simple_spread.py:72-101
 *   MW_SIMPLE_TAG     the first n_adversaries agents chase the others: observation = [vel, pos, landmarks - pos, other
 *                     agents - pos, velocities of the other good agents]; adversaries get +10 per (good, adversary) pair in
 *                     contact, a good agent -10 per adversary touching it and the screen-exit penalty bound(|x|), bound(|y|)
 *                                                                        multiagent/scenarios/simple_tag.py:83-179
 *   d_obs     Real [n_agents][E][obs_stride] (rows zero-padded to obs_stride >= mw_scenario_obs_dim)
 *   d_reward  Real [n_agents][E] */
enum { MW_SIMPLE_SPREAD = 0, MW_SIMPLE_TAG = 1 };
int mw_scenario_obs_dim(int scenario, int n_agents, int n_entities, int n_adversaries);
int mw_scenario_callbacks(const MwConfig *cfg, int scenario, int n_adversaries, const void *d_pos, const void *d_vel,
                          void *d_obs, int obs_stride, void *d_reward, void *stream);

/* reset_world of those scenarios (simple_spread.py:32-45, simple_tag.py:43-58) for the worlds with d_mask[e] != 0 (all when
 * d_mask is NULL): agent positions uniform in [lo_agents, hi_agents)^2, landmark positions uniform in
 * [lo_landmarks, hi_landmarks)^2, velocities zero.  Philox4x32-10 keyed by (seed; env, entity, episode) -- the reference
 * draws from numpy's global MT19937 stream, which a batched engine cannot reproduce; parity tests set the state. */
int mw_scenario_reset(const MwConfig *cfg, double lo_agents, double hi_agents, double lo_landmarks, double hi_landmarks,
                      uint64_t seed, uint32_t episode, const uint8_t *d_mask, void *d_pos, void *d_vel, void *stream);

#ifdef __cplusplus
}
#endif
#endif
