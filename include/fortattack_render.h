/*
 * fortattack_render.h -- C ABI of the device-state rasteriser (libfortattack_b200.so).
 *
 * Replaces the scene the reference draws with pyglet/OpenGL in FortAttackGlobalEnv.render
 *     gym_fortattack/fortattack.py:368-596   (geometry, colours, paint order)
 *     gym_fortattack/rendering.py:37-361     (700 x 700 viewer over [-1, 1]^2, white clear colour, SRC_ALPHA blending)
 * for a SELECTION of the E environments of a batch, straight from the observation planes the step kernel writes
 * (no device -> host copy of the state, no GL context): qualitative checks of what the batched simulator is doing,
 * SURVEY.md 8(f) N4.  Text labels and the shot sound are not reproduced; circles are exact discs (the reference
 * approximates them by 30-gons, rendering.py:260-270).
 *
 * Paint order, as the reference appends its geoms (later covers earlier):
 *   white clear colour
 *   black world rectangle   wall_pos = [-1, 1] x [-0.8, 0.8]                        fortattack.py:408-417
 *   cyan fort disc          radius fortDim = 0.15 at doorLoc = (0, 0.8)             fortattack.py:433-439
 *   yellow attention halos  radius size * (1 + w), alpha 0.9 (alive) / 0.3 (dead)   fortattack.py:450-466   (optional)
 *   dead agents             head + body in colour * 0.5 (core.py:297)               fortattack.py:469-490   (optional)
 *   alive agents, in index order: head disc (0.5 size, 0.8 size ahead), laser triangle of a shooting agent
 *                           (core.py:373-382; agent colour, alpha 0.3), body disc   fortattack.py:493-529
 *   grey strips             |y| > 0.8, colour 0.5                                   fortattack.py:549-563
 * Agent colours: guards (0, 1, 0), attackers (1, 0, 0) (fortattack_env_v1.py:57); size 0.05 (core.py:32).
 * Pixel (px, py) samples the world point (-1 + (px + 0.5) * 2 / W, 1 - (py + 0.5) * 2 / H); a shape covers the pixels
 * whose sample lies inside it.  All predicates and blends are single float operations without contraction, so
 * the numpy restatement (oracle/render_oracle.py) reproduces the image bit for bit.
 */
#ifndef FORTATTACK_RENDER_B200_H
#define FORTATTACK_RENDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct FrConfig {
    int32_t n_envs;        /* E of the planes below */
    int32_t n_guards;      /* agents 0 .. n_guards-1 are guards (green), the rest attackers (red) */
    int32_t n_attackers;
    int32_t width;         /* image size in pixels; the reference's viewer is 700 x 700 (fortattack.py:392) */
    int32_t height;
    int32_t draw_dead;     /* world.vizDead (fortattack_env_v1.py:42: False) */
    int32_t reserved0, reserved1;
} FrConfig;

/* d_obs      float [A][E][6]  alive, x, y, ang, vx, vy -- the observation planes of fa_step / fa_reset
 * d_actions  int32 [A][E] or NULL: agents with action 7 (shoot) get their laser triangle
 * d_halo     float [A][E] or NULL: attention weight w >= 0 per agent (negative = no halo for this agent)
 * d_env_ids  int32 [n_img]: which environments to draw
 * d_rgb      uint8 [n_img][height][width][3], row 0 = top (y = +1)
 * One thread per pixel; the agents' geometry is prepared once per block in shared memory. */
int fr_render(const FrConfig *cfg, const float *d_obs, const int32_t *d_actions, const float *d_halo, const int32_t *d_env_ids,
              int n_img, uint8_t *d_rgb, void *stream);

#ifdef __cplusplus
}
#endif
#endif
