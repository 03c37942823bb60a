"""GPU: the device-state rasteriser (fr_render through the C ABI) against its numpy restatement, bit for bit, on states
the step kernel produced; and the gym facade's render()."""
import os
import sys
from importlib import import_module

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import render_oracle

from gpu_util import make

pytestmark = pytest.mark.gpu
PKG = "emergent-multiagent-strategies_b200"
rd = import_module(PKG + ".render")


@pytest.mark.parametrize("ng,na,w,h", [(3, 3, 700, 700), (5, 5, 333, 257), (1, 2, 64, 48)])
def test_render_matches_numpy_restatement(ng, na, w, h):
    E, A = 200, ng + na
    env = make(E, ng, na, torch.float32, max_steps=60, seed=4)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    p = torch.tensor([.1] * 7 + [.3], device="cuda")
    for _ in range(25):                                         # shoot-heavy stream: some agents are dead by now
        acts = torch.multinomial(p, A * E, replacement=True, generator=g).view(A, E).to(torch.int32)
        obs, _, _, _ = env.step(acts, auto_reset=False)
    assert (obs[:, :, 0] == 0).any() and (acts == 7).any()
    halo = torch.rand(A, E, device="cuda", generator=g) * 1.5 - 0.3        # some negative: no halo
    ids = [0, 7, 199, 42]
    for draw_dead in (False, True):
        imgs = rd.render_batch(obs, ng, actions=acts, env_ids=ids, halo=halo, width=w, height=h, draw_dead=draw_dead)
        assert imgs.shape == (len(ids), h, w, 3) and imgs.dtype == torch.uint8
        for k, e in enumerate(ids):
            ref = render_oracle.render(obs[:, e].cpu().numpy(), ng, actions=acts[:, e].cpu().numpy(), halo=halo[:, e].cpu().numpy(),
                                       width=w, height=h, draw_dead=draw_dead)
            bad = int((imgs[k].cpu().numpy() != ref).any(-1).sum())
            # float64 cos/sin of the device and of libm may differ in their last bit before the single rounding to float32
            assert bad <= 3, (e, draw_dead, bad)
    plain = rd.render_batch(obs, ng, env_ids=[3], width=w, height=h)       # no actions, no halos
    ref = render_oracle.render(obs[:, 3].cpu().numpy(), ng, width=w, height=h)
    assert int((plain[0].cpu().numpy() != ref).any(-1).sum()) <= 3


def test_render_arguments_are_checked():
    obs = torch.zeros(6, 10, 6, device="cuda")
    with pytest.raises(ValueError):
        rd.render_batch(obs, 3, env_ids=[10])
    with pytest.raises(ValueError):
        rd.render_batch(obs.double(), 3)
    with pytest.raises(ValueError):
        rd.render_batch(obs, 3, actions=torch.zeros(6, 10, device="cuda"))
    _capi = import_module(PKG + "._capi")
    cfg = _capi.FrConfig(10, 6, 3, 64, 64, 0, 0, 0)
    ids = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.zeros(64 * 64 * 3, dtype=torch.uint8, device="cuda")
    assert _capi.lib().fr_render(cfg, obs.data_ptr(), None, None, ids.data_ptr(), 1, out.data_ptr(), None) == -1
    assert _capi.lib().fr_render(cfg, None, None, None, ids.data_ptr(), 1, out.data_ptr(), None) == -1


def test_gym_facade_render():
    fa = import_module(PKG + ".gym_fortattack.fortattack")
    env = fa.make_fortattack_env(20, n_guards=3, n_attackers=3, seed=1)
    obs = env.reset()
    obs, _, _, _ = env.step([7, 0, 0, 1, 2, 7])
    attn = [[np.full((3, 3), 0.5), np.full((3, 3), 0.25)], [np.full((3, 3), 0.5), np.full((3, 3), 0.25)]]
    frame = env.render(attn, mode="rgb_array")[0]
    halo = np.array([-1, 0.5, 0.5, 0.25, 0.25, 0.25], np.float32)          # reference agent k = 0 (fortattack.py:441-446)
    ref = render_oracle.render(obs.astype(np.float32), 3, actions=[7, 0, 0, 1, 2, 7], halo=halo)
    assert frame.shape == (700, 700, 3) and int((frame != ref).any(-1).sum()) <= 3
    assert env.render() == [True] and env.last_frame.shape == (700, 700, 3)
    env.close()
