"""TEST INFRASTRUCTURE (not product code): numpy restatement of the device-state rasteriser (include/fortattack_render.h,
csrc/fr_render.cu), i.e. of the scene FortAttackGlobalEnv.render draws (gym_fortattack/fortattack.py:368-596: world
rectangle, fort disc, attention halos, dead / alive agents with head, laser triangle and body, grey strips; colours from
fortattack_env_v1.py:57 and core.py:297; 700 x 700 viewer over [-1, 1]^2, rendering.py:90).

Parity status: pinned to the reference's own pixels as far as they exist -- its renderer needs an OpenGL display (pyglet)
and approximates circles by 30-gons, so no golden image can be reproduced bit for bit, but two frames of its recording
out_files/1.gif are committed (tests/golden/render_ref_frames.npz) and tests/test_render_cpu.py compares against them: the
strip rows, the fort disc (IoU 0.99) and the agent blobs of the reset frame (position fitted, body + head IoU > 0.9, area within
5 %); plus the geometry constants and the blob area over further frames (945..981 px; pi * 17.5^2 = 962).  The CUDA kernel is compared with THIS restatement
bit for bit: every operation below is a single float32 operation in the kernel's order, trigonometry in float64 rounded
once to float32."""
import numpy as np

F = np.float32
SIZE = F(0.05)


def _disc(wx, wy, cx, cy, r2):
    dx, dy = wx - cx, wy - cy
    return dx * dx + dy * dy <= r2


def _edge(ax, ay, bx, by, px, py):
    return (bx - ax) * (py - ay) - (by - ay) * (px - ax)


def _blend(c, rgb, a, m):
    om = F(1) - F(a)
    for k in range(3):
        c[k] = np.where(m, c[k] * om + F(rgb[k]) * F(a), c[k])


def _paint(c, rgb, m):
    for k in range(3):
        c[k] = np.where(m, F(rgb[k]), c[k])


def _agent(c, wx, wy, x, y, ang, guard, shoot, scale):
    rgb = [F(v) * F(scale) for v in ((0, 1, 0) if guard else (1, 0, 0))]
    ang = np.float64(ang)
    cs, sn = np.cos(ang), np.sin(ang)
    hx, hy = F(np.float64(x) + 0.8 * 0.05 * cs), F(np.float64(y) + 0.8 * 0.05 * sn)
    half = F(0.5) * SIZE
    _paint(c, rgb, _disc(wx, wy, hx, hy, half * half))
    if shoot:
        p1x, p1y = np.float64(x) + np.float64(SIZE) * cs, np.float64(y) + np.float64(SIZE) * sn
        hw = 0.39269908169872414
        lx = [F(p1x), F(p1x + 0.8 * np.cos(ang + hw)), F(p1x + 0.8 * np.cos(ang - hw))]
        ly = [F(p1y), F(p1y + 0.8 * np.sin(ang + hw)), F(p1y + 0.8 * np.sin(ang - hw))]
        e = [_edge(lx[i], ly[i], lx[(i + 1) % 3], ly[(i + 1) % 3], wx, wy) for i in range(3)]
        inside = ((e[0] >= 0) & (e[1] >= 0) & (e[2] >= 0)) | ((e[0] <= 0) & (e[1] <= 0) & (e[2] <= 0))
        _blend(c, rgb, 0.3, inside)
    _paint(c, rgb, _disc(wx, wy, F(x), F(y), SIZE * SIZE))


def render(obs, n_guards, actions=None, halo=None, width=700, height=700, draw_dead=False):
    """obs float32 [A, 6] of ONE env (alive, x, y, ang, vx, vy); actions int [A] or None; halo float [A] or None.
    Returns uint8 [height, width, 3]."""
    obs = np.asarray(obs, dtype=np.float32)
    A = obs.shape[0]
    sx, sy = F(2) / F(width), F(2) / F(height)
    px, py = np.meshgrid(np.arange(width, dtype=np.float32), np.arange(height, dtype=np.float32))
    wx = (px + F(0.5)) * sx - F(1)
    wy = F(1) - (py + F(0.5)) * sy
    c = [np.ones((height, width), np.float32) for _ in range(3)]
    _paint(c, (0, 0, 0), (wx >= F(-1)) & (wx <= F(1)) & (wy >= F(-0.8)) & (wy <= F(0.8)))
    _paint(c, (0, 1, 1), _disc(wx, wy, F(0), F(0.8), F(0.15) * F(0.15)))
    alive = obs[:, 0] != 0
    if halo is not None:
        for i in range(A):
            w = F(halo[i])
            if w >= 0 and (alive[i] or draw_dead):
                r = SIZE * (F(1) + w)
                _blend(c, (1, 1, 0), 0.9 if alive[i] else 0.3, _disc(wx, wy, obs[i, 1], obs[i, 2], r * r))
    if draw_dead:
        for i in range(A):
            if not alive[i]:
                _agent(c, wx, wy, obs[i, 1], obs[i, 2], obs[i, 3], i < n_guards, False, 0.5)
    for i in range(A):
        if alive[i]:
            _agent(c, wx, wy, obs[i, 1], obs[i, 2], obs[i, 3], i < n_guards, actions is not None and int(actions[i]) == 7, 1.0)
    _paint(c, (0.5, 0.5, 0.5), (wy > F(0.8)) | (wy < F(-0.8)))
    return np.stack([np.floor(ch * F(255) + F(0.5)).astype(np.uint8) for ch in c], axis=-1)
