"""CPU: the weight blob pack_mpnn() builds for the fused policy kernel decodes back to the module's
function (including the folded output projection), and the blob has the size the header declares."""
from importlib import import_module

import pytest
import torch

import policy_util as pu

PKG = "emergent-multiagent-strategies_b200"
MPNN = import_module(PKG + ".mpnn").MPNN
pk = import_module(PKG + ".policy_kernel")


class Shape(object):
    def __init__(self, *shape):
        self.shape = shape


def make(n, m, seed=0, bias=True):
    torch.manual_seed(seed)
    net = MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6, hidden_dim=128, pos_index=2)
    if bias:                       # trained checkpoints have non-zero biases; zero init would hide offset bugs
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 1:
                    p.uniform_(-0.3, 0.3)
    return net


@pytest.mark.parametrize("n,m", [(3, 3), (5, 5), (1, 1), (2, 4), (4, 1)])
def test_blob_decodes_to_module(n, m):
    net = make(n, m, seed=n * 10 + m)
    blob = pk.pack_mpnn(net)
    assert blob.dtype == torch.uint8 and blob.numel() == pk.BLOB_F16_BYTES + 4 * pk.BLOB_CONST_FLOATS
    gen = torch.Generator().manual_seed(1)
    own, opp = pu.random_obs(n, 37, gen), pu.random_obs(m, 37, gen)
    lg_ref, v_ref = pu.module_forward(net, own, opp)
    lg, v = pu.emulate(blob, own, opp, quantize=False)
    # the only difference is the fp16 rounding of the GEMM weights
    assert (lg - lg_ref.double()).abs().max() < 2e-2 * max(1.0, float(lg_ref.abs().max()))
    assert (v - v_ref.double()).abs().max() < 2e-2 * max(1.0, float(v_ref.abs().max()))
    lgq, vq = pu.emulate(blob, own, opp, quantize=True)
    assert (lgq - lg).abs().max() < 5e-2 * max(1.0, float(lg.abs().max()))


def test_blob_exact_with_fp16_representable_weights():
    """With weights that fp16 holds exactly the decoded network equals the module to fp32 round-off."""
    net = make(3, 3, seed=5)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(p.to(torch.float16).float())
        # the folded projection W' = U2 W_out^T is rounded once more; make it exact by a sparse W_out
        net.messages.W_out.zero_()
        net.messages.W_out[0].fill_diagonal_(0.5)
    blob = pk.pack_mpnn(net)
    gen = torch.Generator().manual_seed(2)
    own, opp = pu.random_obs(3, 64, gen), pu.random_obs(3, 64, gen)
    lg_ref, v_ref = pu.module_forward(net, own, opp)
    lg, v = pu.emulate(blob, own, opp, quantize=False)
    assert (lg - lg_ref.double()).abs().max() < 1e-4 * max(1.0, float(lg_ref.abs().max()))
    assert (v - v_ref.double()).abs().max() < 1e-4 * max(1.0, float(v_ref.abs().max()))


def test_unsupported_shapes_rejected():
    net = MPNN(action_space=Shape(8), num_agents=3, num_opp_agents=3, input_size=6, hidden_dim=64)
    with pytest.raises(Exception):
        pk.pack_mpnn(net)
