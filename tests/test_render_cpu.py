"""CPU: known answers of the rasteriser's numpy restatement (oracle/render_oracle.py) derived from the reference's scene
description (gym_fortattack/fortattack.py:368-596) and from its recorded out_files/1.gif."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import render_oracle as ro

GREEN, RED, BLACK, GREY, CYAN = (0, 255, 0), (255, 0, 0), (0, 0, 0), (128, 128, 128), (0, 255, 255)


def count(img, rgb):
    return int((img == np.array(rgb, np.uint8)).all(-1).sum())


def one_guard(x=0.3, y=0.1, ang=1.0, alive=1.0):
    obs = np.zeros((2, 6), np.float32)
    obs[0] = [alive, x, y, ang, 0, 0]
    obs[1] = [0.0, -0.5, -0.5, 0, 0, 0]          # a dead attacker: not drawn unless draw_dead
    return obs


def test_static_scene():
    img = ro.render(one_guard(alive=0.0), 1)
    assert img.shape == (700, 700, 3)
    # grey strips |y| > 0.8 <-> rows 0..69 and 630..699 (fortattack.py:549-563); black world in between
    assert (img[:70] == GREY).all() and (img[630:] == GREY).all() and count(img[70:630], GREY) == 0
    # lower half of the fort disc (radius 0.15 = 52.5 px at doorLoc (0, 0.8)) shows below the strip: pi r^2 / 2
    assert abs(count(img, CYAN) - math.pi * 52.5 ** 2 / 2) < 60
    assert count(img, CYAN) + count(img, BLACK) == 560 * 700 and count(img, GREEN) == 0 == count(img, RED)


def test_agent_blob_area_matches_the_reference_recording():
    """Body disc radius 0.05 = 17.5 px, head disc 8.75 px centred 14 px ahead.  Isolated agent blobs in the reference's own
    recording out_files/1.gif (frames 1, 7, 14; green, colour-thresholded on its dithered palette) measure 945..981 px;
    the reference draws 30-gons (rendering.py:260-270), whose body area is 955 px."""
    img = ro.render(one_guard(), 1)
    R, r, d = 17.5, 8.75, 14.0
    lens = (r * r * math.acos((d * d + r * r - R * R) / (2 * d * r)) + R * R * math.acos((d * d + R * R - r * r) / (2 * d * R))
            - 0.5 * math.sqrt((-d + r + R) * (d + r - R) * (d - r + R) * (d + r + R)))
    union = math.pi * (R * R + r * r) - lens
    assert abs(count(img, GREEN) - union) < 0.015 * union
    for gif_blob in (964, 971, 974, 981):                       # measured from out_files/1.gif
        assert 0.90 < gif_blob / count(img, GREEN) < 1.0        # the threshold drops the dithered rim and part of the head


def test_dead_laser_halo_and_order():
    obs = one_guard(x=-0.5, y=0.0, ang=0.0)
    base = ro.render(obs, 1)
    # dead agents appear only with draw_dead, in half colour (core.py:297)
    dead = ro.render(obs, 1, draw_dead=True)
    assert count(base, (128, 0, 0)) == 0 and count(dead, (128, 0, 0)) > 900
    # laser triangle: area 0.5 * 0.8^2 * sin(pi/4) world units = 0.2263 * 350^2 px, blended 0.3 * green over black = (0, 77, 0)
    shot = ro.render(obs, 1, actions=[7, 7])                    # the dead attacker's action is ignored (core.py:268)
    tri = 0.5 * 0.8 * 0.8 * math.sin(math.pi / 4) * 350 * 350
    assert abs(count(shot, (0, 77, 0)) - tri) < 0.04 * tri and count(shot, (77, 0, 0)) == 0
    assert count(shot, GREEN) == count(base, GREEN)             # body and head cover the laser (paint order)
    # halo: yellow disc of radius size * (1 + w) under the agent, alpha 0.9
    halo = ro.render(obs, 1, halo=[1.0, -1.0])
    ring = math.pi * (35.0 ** 2) - count(base, GREEN)
    assert abs(count(halo, (230, 230, 0)) - ring) < 0.06 * ring
    # a later agent covers an earlier one
    two = np.zeros((2, 6), np.float32)
    two[0] = [1, 0.0, 0.0, 0.0, 0, 0]
    two[1] = [1, 0.02, 0.0, 0.0, 0, 0]
    img = ro.render(two, 1)
    assert count(img, RED) > count(img, GREEN) > 0


def test_attention_halos_pick_the_reference_agent_like_the_reference():
    """render.attention_halos (pure tensor code) against the selection loop of fortattack.py:441-466, per env."""
    import torch
    from importlib import import_module
    rd = import_module("emergent-multiagent-strategies_b200.render")
    g = torch.Generator().manual_seed(0)
    ng, na, E = 3, 4, 64
    obs = torch.zeros(ng + na, E, 6)
    obs[:, :, 0] = (torch.rand(ng + na, E, generator=g) > 0.4).float()
    team, opp = torch.rand(E, ng, ng, generator=g), torch.rand(E, ng, na, generator=g)
    halo = rd.attention_halos(obs, ng, team, opp)
    for e in range(E):
        k = next(i for i in range(ng + na) if obs[i, e, 0] != 0 or i == ng - 1)       # :441-446
        assert k < ng
        for i in range(ng + na):
            want = -1.0 if i == k else float(team[e, k, i] if i < ng else opp[e, k, i - ng])     # :452-458
            assert float(halo[i, e]) == want


# ---- against the reference's own pixels: two frames decoded from out_files/1.gif (tests/golden/make_render_golden.py) ----------
FRAMES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_ref_frames.npz")


def ref_frame(k):
    g = np.load(FRAMES)
    return g["frame%d/palette" % k][g["frame%d/index" % k]].astype(int)            # [700, 700, 3]


def near(img, rgb, tol=60):
    return np.abs(img.astype(int) - np.array(rgb)).sum(-1) < tol


def iou(a, b):
    return (a & b).sum() / max(1, (a | b).sum())


def test_static_scene_matches_reference_frames():
    """World rectangle, strips and fort disc of the reference's recorded frames against the rasteriser: the strip rows are the
    same rows, and the fort disc of frame 8 (nothing in front of it) covers the same pixels (IoU > 0.97; the reference draws
    a 30-gon, rendering.py:260-270)."""
    ours = ro.render(one_guard(alive=0.0), 1)
    for k in (0, 8):
        ref = ref_frame(k)
        nonblack = ref.sum(-1) > 150                               # the recording's edges are one blurred row wide: half intensity decides
        col = nonblack[:, 40]                                      # a column no agent crosses in either frame
        assert col[:70].all() and not col[70:630].any() and col[630:].all()
        assert ((ours[:, 40] != 0).any(-1) == col).all()
    from scipy import ndimage
    ref = ref_frame(8)
    # the recording is dithered over a palette with four blue levels: cyan-ish pixels, closed over the dither pattern, largest blob
    m = (ref[..., 1] >= 144) & (ref[..., 2] >= 85) & (ref[..., 0] <= 110)
    m[:70] = False
    lab, n = ndimage.label(ndimage.binary_fill_holes(ndimage.binary_closing(m, np.ones((3, 3)))))
    fort_ref = lab == 1 + int(np.argmax(ndimage.sum(np.ones_like(lab), lab, range(1, n + 1))))
    fort_ours = near(ours, CYAN, 1)
    assert iou(fort_ref, fort_ours) > 0.97, iou(fort_ref, fort_ours)


def test_agent_blobs_match_reference_frame():
    """Frame 0 of the recording is a reset state: attackers head up (pi/2), no lasers, no halos.  For the two attackers that are
    isolated and fully inside the world, the position is fitted from the blob and the rasteriser's agent (body disc + head) must
    cover the same pixels as the reference's (IoU > 0.9, area within 5 %; the black digits the reference prints on the
    agents are filled first)."""
    from scipy import ndimage
    ref = ref_frame(0)
    red = ndimage.binary_fill_holes(near(ref, (252, 0, 0), 120))
    lab, n = ndimage.label(red)
    checked = 0
    for b in range(1, n + 1):
        m = lab == b
        ys, xs = np.nonzero(m)
        if m.sum() < 900 or m.sum() > 1300 or xs.min() < 5 or ys.max() > 625:       # merged, clipped by the frame or by the strip
            continue
        # fit: place the agent so that the rasterised blob's centroid coincides with the recorded blob's
        x, y = (xs.mean() + 0.5) / 350 - 1, 1 - (ys.mean() + 0.5) / 350
        for _ in range(3):
            obs = np.zeros((2, 6), np.float32)
            obs[0] = [0.0, 0, 0, 0, 0, 0]
            obs[1] = [1.0, x, y, math.pi / 2, 0, 0]
            mine = near(ro.render(obs, 1), RED, 1)
            my, mx = np.nonzero(mine)
            x += (xs.mean() - mx.mean()) / 350
            y -= (ys.mean() - my.mean()) / 350
        assert iou(m, mine) > 0.9, (b, iou(m, mine))
        assert abs(int(mine.sum()) - int(m.sum())) < 0.05 * m.sum(), (int(mine.sum()), int(m.sum()))
        checked += 1
    assert checked >= 2
