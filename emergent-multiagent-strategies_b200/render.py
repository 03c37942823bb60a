"""Device-state renderer: the scene of FortAttackGlobalEnv.render (gym_fortattack/fortattack.py:368-596) for selected
environments of a batch, rasterised on the GPU from the observation planes (fr_render, include/fortattack_render.h).

    frames = render_batch(obs, n_guards=3, actions=acts, env_ids=[0, 17, 4095])      # uint8 [3, 700, 700, 3] on the device

`attention_halos` turns the attention matrices MPNN exposes (mpnn.py:140,166; the `attn_list` argument of the reference's
render) into the per-agent halo weights the reference draws (fortattack.py:441-466)."""
import torch

from . import _capi


def render_batch(obs, n_guards, actions=None, env_ids=None, halo=None, width=700, height=700, draw_dead=False, out=None):
    """obs float32 [A, E, 6] (device); actions int32 [A, E] or None; halo float32 [A, E] or None (negative = no halo);
    env_ids: iterable / tensor of env indices (default: all, only sensible for small E).  Returns uint8
    [n, height, width, 3] on the device, row 0 at the top."""
    if not obs.is_cuda or obs.dtype != torch.float32 or obs.dim() != 3 or obs.shape[2] != 6 or not obs.is_contiguous():
        raise ValueError("obs must be a contiguous float32 CUDA tensor [A, E, 6]")
    A, E, _ = obs.shape
    dev = obs.device
    ids = torch.arange(E, dtype=torch.int32, device=dev) if env_ids is None else \
        torch.as_tensor(env_ids, dtype=torch.int32).to(dev).contiguous()
    if ids.dim() != 1 or ids.numel() < 1 or int(ids.min()) < 0 or int(ids.max()) >= E:
        raise ValueError("env_ids must be a non-empty 1-D selection of 0..%d" % (E - 1))

    def plane(t, dtype, name):
        if t is None:
            return None
        if tuple(t.shape) != (A, E) or t.dtype != dtype or t.device != dev or not t.is_contiguous():
            raise ValueError("%s must be a contiguous %s tensor [%d, %d] on %s" % (name, dtype, A, E, dev))
        return t.data_ptr()

    n = ids.numel()
    if out is None:
        out = torch.empty(n, height, width, 3, dtype=torch.uint8, device=dev)
    elif tuple(out.shape) != (n, height, width, 3) or out.dtype != torch.uint8 or out.device != dev or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 tensor [%d, %d, %d, 3] on %s" % (n, height, width, dev))
    cfg = _capi.FrConfig(E, int(n_guards), A - int(n_guards), int(width), int(height), int(bool(draw_dead)), 0, 0)
    _capi.check(_capi.lib().fr_render(cfg, obs.data_ptr(), plane(actions, torch.int32, "actions"),
                                      plane(halo, torch.float32, "halo"), ids.data_ptr(), n, out.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream))
    return out


def attention_halos(obs, n_guards, team_attn, opp_attn):
    """Per-agent halo weights [A, E] as the reference picks them (fortattack.py:441-466): the reference agent k is the first
    alive agent (or the last guard if no earlier one is alive); teammate i of the guards gets team_attn[e, k, i], attacker j
    gets opp_attn[e, k, j]; k itself gets none (-1).  team_attn [E, n_guards, n_guards], opp_attn [E, n_guards, n_att]
    are the guards' attention matrices."""
    A, E, _ = obs.shape
    alive = obs[:, :, 0] != 0                                                    # [A, E]
    first = torch.where(alive[:n_guards].any(0), alive[:n_guards].float().argmax(0), torch.full((E,), n_guards - 1, device=obs.device))
    e = torch.arange(E, device=obs.device)
    halo = torch.cat((team_attn[e, first].t(), opp_attn[e, first].t()), 0).to(torch.float32).contiguous()    # [A, E]
    halo[first, e] = -1.0
    return halo
