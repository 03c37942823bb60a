"""Checker for the fused policy kernel: decodes the packed weight blob (layout documented in
csrc/mp_policy.cu) and evaluates the network in plain torch, optionally rounding activations to fp16 at
exactly the points where the kernel does.  quantize=False validates pack_mpnn() against the module;
quantize=True is the kernel's arithmetic up to fp32 summation order."""
import importlib

import torch

pk = importlib.import_module("emergent-multiagent-strategies_b200.policy_kernel")


def _uncanon(flat, n, k):
    return flat.reshape(n // 8, k // 8, 8, 8).permute(0, 2, 1, 3).reshape(n, k)


def decode_blob(blob):
    f16 = blob[:pk.BLOB_F16_BYTES].view(torch.float16).double()
    c = blob[pk.BLOB_F16_BYTES:].view(torch.float32).double()
    w, pos = {}, 0

    def take(n, k):
        nonlocal pos
        t = _uncanon(f16[pos:pos + n * k], n, k)
        pos += n * k
        return t
    w["og"], w["oz"] = take(64, 64), take(64, 64)                # (Wkey Wquery^T)^T, (Wval Wout)^T
    w["g"], w["wz"], w["u1"] = take(128, 128), take(128, 128), take(128, 128)
    w["v0"] = torch.cat((take(64, 128), take(64, 128)))
    w["p0"] = torch.cat((take(64, 128), take(64, 128)))
    assert pos * 2 == pk.BLOB_F16_BYTES
    enc, oenc = c[0:512].reshape(64, 8), c[512:1024].reshape(64, 8)
    w.update(enc_w=enc[:, :6], enc_b=enc[:, 6], oenc_w=oenc[:, :6], oenc_b=oenc[:, 6], ub=c[1024:1152], vb=c[1152:1280],
             vw=c[1280:1408], pb=c[1408:1536], dw=c[1536:2560].reshape(128, 8), db=c[2560:2568], vb2=c[2568])
    return w


def emulate(blob, own, opp, quantize):
    """own [n,E,6], opp [m,E,6] -> (logits [n,E,8], value [n,E]) in float64."""
    w = decode_blob(blob.cpu())
    own, opp = own.detach().cpu().double(), opp.detach().cpu().double()
    q16 = (lambda t: t.to(torch.float16).double()) if quantize else (lambda t: t)
    relu = torch.relu
    h0 = q16(relu(own @ w["enc_w"].t() + w["enc_b"]))            # [n,E,64]
    ho = q16(relu(opp @ w["oenc_w"].t() + w["oenc_b"]))          # [m,E,64]
    T = h0 @ w["og"].t()                                         # fp32 accumulators stay unrounded
    Z = q16(ho @ w["oz"].t())
    p = torch.softmax(torch.einsum("nek,mek->enm", T, ho) * 0.125, dim=-1)
    h = torch.cat((h0, q16(torch.einsum("enm,mek->nek", p, Z))), dim=-1)     # [n,E,128]
    n = own.shape[0]
    for _ in range(3):
        T, Z, Y = h @ w["g"].t(), q16(h @ w["wz"].t()), h @ w["u1"].t()
        s = torch.einsum("aek,bek->eab", T, h) / 128 ** 0.5
        if n > 1:
            p = torch.softmax(s.masked_fill(torch.eye(n, dtype=torch.bool), float("-inf")), dim=-1)
        else:
            p = torch.zeros_like(s)
        h = q16(relu(Y + torch.einsum("eab,bek->aek", p, Z) + w["ub"]))
    value = relu(h @ w["v0"].t() + w["vb"]) @ w["vw"] + w["vb2"]
    logits = relu(h @ w["p0"].t() + w["pb"]) @ w["dw"] + w["db"]
    return logits, value


def module_forward(m, own, opp):
    """The torch module on the same inputs -> (logits [n,E,8], value [n,E])."""
    n, E = own.shape[0], own.shape[1]
    with torch.no_grad():
        x = m._fwd(own.reshape(-1, 6), opp.reshape(-1, 6), None)
        value = m._value(x).view(n, E)
        logits = m.dist.linear(m._policy(x)).view(n, E, 8)
    return logits, value


def random_obs(n, E, gen, device="cpu"):
    """Observation-like rows [alive, x, y, ang, vx, vy] (fortattack_env_v1.py:238)."""
    o = torch.empty(n, E, 6)
    o[..., 0] = (torch.rand(n, E, generator=gen) > 0.2).float()
    o[..., 1] = torch.rand(n, E, generator=gen) * 2 - 1
    o[..., 2] = torch.rand(n, E, generator=gen) * 1.6 - 0.8
    o[..., 3] = torch.rand(n, E, generator=gen) * 40
    o[..., 4:] = torch.randn(n, E, 2, generator=gen)
    return o.to(device)


def trained_state_dict():
    """The guards' policy of the reference's shipped checkpoint marlsave/tmp_1/ep2520.pt (5v5), as stored in
    tests/golden/rl_mpnn.npz by tests/golden/make_rl_golden.py (ckpt/param/*)."""
    import os
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rl_mpnn.npz"))
    return {k[len("ckpt/param/"):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("ckpt/param/")}, d


def trained_net(n=5, m=5, device="cpu"):
    """MPNN with the shipped ep2520 guard weights (no parameter depends on the team sizes: SURVEY 3.5)."""
    mp = importlib.import_module("emergent-multiagent-strategies_b200.mpnn")

    class Shape(object):
        shape = (8,)
    net = mp.MPNN(action_space=Shape(), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6, hidden_dim=128, pos_index=2)
    net.load_state_dict(trained_state_dict()[0])
    return net.to(device)
