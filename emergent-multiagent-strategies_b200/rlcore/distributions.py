"""Categorical action head of the reference (rlcore/distributions.py:9-32).

`FixedCategorical` is torch.distributions.Categorical with column-vector conventions: sample() and
mode() return [N,1], log_probs(actions[N,1]) returns [N,1].  Implemented as a subclass (the reference
monkey-patches the torch class globally, distributions.py:9-17)."""
import torch
import torch.nn as nn


class FixedCategorical(torch.distributions.Categorical):
    def sample(self, sample_shape=torch.Size()):
        return super().sample(sample_shape).unsqueeze(-1)

    def log_probs(self, actions):
        return super().log_prob(actions.squeeze(-1)).unsqueeze(-1)

    def mode(self):
        return self.probs.argmax(dim=1, keepdim=True)


class Categorical(nn.Module):
    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.linear = nn.Linear(num_inputs, num_outputs)
        nn.init.orthogonal_(self.linear.weight.data, gain=0.01)       # distributions.py:23-28
        nn.init.constant_(self.linear.bias.data, 0)

    def forward(self, x):
        x = self.linear(x)
        # argument validation synchronises the device (constraint checks end in .all()): skip it while a CUDA graph is
        # being captured (JointPPO(graph_update=True))
        capturing = x.is_cuda and torch.cuda.is_current_stream_capturing()
        return FixedCategorical(logits=x, validate_args=False if capturing else None)
