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

#ifdef __cplusplus
}
#endif
#endif
