"""Shared fusion heads (reference models/fusion_modules.py:16-35).

Under --gs_flag only `fc_out` is used, applied to ONE modality's feature at a time
(main.py:432,444,456,636-639); `forward` (concatenation) is kept for API completeness.
`fc_out` is a plain nn.Linear so that `named_parameters()` yields `weight`/`bias` exactly as
in the reference (this is what makes the published GS hook a no-op). Calling it on
a CUDA tensor under no_grad (evaluation) runs the native head kernel; with autograd enabled
it stays a regular differentiable Linear. The fused forward+backward used by train_epoch is
`head_turn`.
"""
import torch
import torch.nn as nn

from . import ops


class SharedLinear(nn.Linear):
    """nn.Linear whose no-grad CUDA forward is the native head kernel (logits only)."""

    def forward(self, x):
        if x.is_cuda and not torch.is_grad_enabled() and x.dim() == 2 and x.dtype == torch.float32 \
                and x.shape[1] % 4 == 0:
            label = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device)
            o = ops.head_ce(x.contiguous(), self.weight.detach(), self.bias.detach(), label, need_grad=False)
            return o["logits"]
        return super().forward(x)


def head_turn(fc_out, feat, label, grad_scale=None, out=None):
    """One modality turn of the head, fused (main.py:432-435): sets fc_out.weight.grad and
    fc_out.bias.grad, returns dict(logits, loss, dfeat, feat_sum, dW, db)."""
    o = ops.head_ce(feat, fc_out.weight.detach(), fc_out.bias.detach(), label, grad_scale=grad_scale, out=out)
    fc_out.weight.grad = o["dW"]
    fc_out.bias.grad = o["db"]
    return o


class ConcatFusion(nn.Module):
    def __init__(self, input_dim=512, output_dim=100):
        super().__init__()
        self.fc_out = SharedLinear(input_dim, output_dim)

    def forward(self, x, y):
        return x, y, self.fc_out(torch.cat((x, y), dim=1))


class ConcatFusion3(nn.Module):
    def __init__(self, input_dim=512, output_dim=100):
        super().__init__()
        self.fc_out = SharedLinear(input_dim, output_dim)

    def forward(self, x, y, z):
        return x, y, z, self.fc_out(torch.cat((x, y, z), dim=1))
