"""End-to-end ENCODER-GRADIENT parity of the ResNet-18 path and the autograd contract of the
encoder engine.

Gradient criterion. TF32-class arithmetic flips ReLU masks, which puts a ~10 % Frobenius error on the encoder gradients of
ANY 10-bit-operand implementation (torch's own cuDNN-TF32 path vs torch fp32: median 1e-1). A flat 1e-3 on dW is
therefore unattainable even for the reference run on a GPU. The test is calibrated against the reference's own arithmetic
on the same GPU: for EVERY parameter, err(ours, fp32) <= 1.25 * err(torch cuDNN-TF32, fp32) + 2e-4, in every arithmetic
mode the engine offers (2-byte operands: fp16 forward and power-of-two-scaled fp16 backward; TF32 operands) — 1.25 for
the convolution weights, 1.6 for the BatchNorm affine vectors (64-512 elements: the ratio of two noise realisations scatters
more) — and the MEDIAN ratio over all 120 encoder parameters must not exceed 1.15 (measured: 1.09 with 2-byte operands, 1.08 with TF32
operands — the two arithmetic modes are indistinguishable at the gradient level; profiles/r2_encoder_grad_modes.txt).
Quantities no ReLU mask sits in front of — the head's projected gradient — are held to the flat rel 1e-3 of north_star.
"""
import argparse
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import mla_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _args():
    return argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=True, lorb="base", modal3=False, clip=False)


def _model():
    import mla_b200
    mla_b200.setup_seed(0)
    net = mla_b200.AVClassifier(_args()).apply(mla_b200.weight_init)
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net.cuda(), state


def _torch_grads(state, spec, image, wa, wv, tf32):
    """Parameter gradients of L = <a, wa> + <v, wv> through the oracle's torch restatement on the GPU: cuDNN fp32
    (tf32=False: the referee) or cuDNN TF32 (tf32=True: what the reference itself runs under torch's defaults)."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        sd = {k: v.cuda().clone() for k, v in state.items()}
        names = [k for k in sd if sd[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))
                 and not k.startswith("fusion_module")]
        for k in names:
            sd[k].requires_grad_(True)
        a, v = orc.av_forward(sd, spec.unsqueeze(1), image, training=True)
        ((a * wa).sum() + (v * wv).sum()).backward()
        return {k: sd[k].grad.detach().double().cpu().numpy() for k in names}, a.detach(), v.detach()
    finally:
        torch.backends.cudnn.allow_tf32 = prev


_MODE_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import test_gpu_encoder_grad as T
from oracle import mla_oracle as orc
net, state = T._model()
spec, image, _ = orc.synthetic_av_batch(8, 33)
spec, image = spec.cuda(), image.cuda()
g = torch.Generator().manual_seed(7)
wa, wv = (torch.randn(8, 512, generator=g) / 8).cuda(), (torch.randn(8, 512, generator=g) / 8).cuda()
net.train()
a, v = net(spec.unsqueeze(1), image)
((a * wa).sum() + (v * wv).sum()).backward()
torch.cuda.synchronize()
out = {k: p.grad.detach().double().cpu().numpy() for k, p in net.named_parameters() if p.grad is not None}
out["__a"], out["__v"] = a.detach().cpu().numpy(), v.detach().cpu().numpy()
np.savez(sys.argv[1], **out)
print("worker ok", os.environ.get("MLA_F16"))
"""


def _ours_in_mode(tmp_path, f16):
    """Our encoder gradients in one arithmetic mode; the mode is fixed at import time (MLA_F16), hence a subprocess."""
    script = tmp_path / ("worker_%s.py" % f16)
    script.write_text(_MODE_WORKER % {"root": ROOT})
    out = str(tmp_path / ("grads_%s.npz" % f16))
    env = dict(os.environ, MLA_F16=f16, MLA_GRAPHS="0")
    r = subprocess.run([sys.executable, str(script), out], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return dict(np.load(out))


def test_encoder_gradients_three_arithmetic_modes(built_lib, tmp_path):
    """ONE full-size ResNet-18 pair backward (B = 8, 1x257x188 / 2x3x224x224) in ours-TF32, ours-2-byte and torch
    cuDNN-TF32, per-parameter rel-F of dW against torch fp32."""
    _, state = _model()
    spec, image, _ = orc.synthetic_av_batch(8, 33)
    spec, image = spec.cuda(), image.cuda()
    g = torch.Generator().manual_seed(7)
    wa, wv = (torch.randn(8, 512, generator=g) / 8).cuda(), (torch.randn(8, 512, generator=g) / 8).cuda()
    ref, ra, rv = _torch_grads(state, spec, image, wa, wv, tf32=False)
    cud, ca, cv = _torch_grads(state, spec, image, wa, wv, tf32=True)
    modes = {"2-byte (fp16 fwd, scaled fp16 bwd)": _ours_in_mode(tmp_path, "1"), "tf32": _ours_in_mode(tmp_path, "0")}
    failures = []
    for mode, ours in modes.items():
        ea, ev = relf(ours["__a"], ra.cpu().numpy()), relf(ours["__v"], rv.cpu().numpy())
        print("%s: feature rel-F %.2e %.2e (cuDNN-TF32: %.2e %.2e)" % (mode, ea, ev, relf(ca.cpu().numpy(), ra.cpu().numpy()),
                                                                      relf(cv.cpu().numpy(), rv.cpu().numpy())))
        if not (ea < 1e-3 and ev < 1e-3):
            failures.append("%s: features %.2e %.2e" % (mode, ea, ev))
        ratios = []
        for k in ref:
            e_ours, e_cud = relf(ours[k], ref[k]), relf(cud[k], ref[k])
            ratio = e_ours / max(e_cud, 1e-12)
            ratios.append(ratio)
            line = "  %-44s ours %.3e  cuDNN-TF32 %.3e  ratio %.2f" % (k, e_ours, e_cud, ratio)
            print(line)
            # weight tensors (thousands to millions of elements): 1.25 x; BatchNorm affine vectors (64-512 elements, so
            # the ratio of two noise realisations scatters more): 1.6 x
            lim = 1.25 if ref[k].ndim == 4 else 1.6
            if not e_ours <= lim * e_cud + 2e-4:
                failures.append(mode + line)
        ratios = np.sort(np.array(ratios))
        print("%s: error ratio ours / cuDNN-TF32 over %d parameters: min %.2f median %.2f p90 %.2f max %.2f"
              % (mode, len(ratios), ratios[0], np.median(ratios), ratios[int(0.9 * len(ratios))], ratios[-1]))
        if not np.median(ratios) <= 1.15:
            failures.append("%s: median error ratio %.3f > 1.15" % (mode, np.median(ratios)))
    assert not failures, "\n".join(failures)


def test_gradients_accumulate_through_the_autograd_path(built_lib):
    """the native backward overwrites its buffers; the autograd node adds back gradients that were already
    there, so two backward passes without zero_grad give the sum (torch semantics)."""
    net, _ = _model()
    spec, image, _ = orc.synthetic_av_batch(2, 5, spec_hw=(65, 48), image_hw=(64, 64))
    spec, image = spec.cuda(), image.cuda()
    net.train()

    def run():
        a, v = net(spec.unsqueeze(1), image)
        (a.sum() + 2 * v.sum()).backward()

    run()
    g1 = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    run()
    for k, p in net.named_parameters():
        if k in g1:
            assert relf(p.grad.cpu(), (2 * g1[k]).cpu()) < 1e-6, k


def test_backward_through_overwritten_activations_raises(built_lib):
    """a plan owns one set of activation buffers per input shape; backward through an EARLIER forward of the same
    shape must raise instead of differentiating through the later forward's activations."""
    net, _ = _model()
    spec, image, _ = orc.synthetic_av_batch(2, 5, spec_hw=(65, 48), image_hw=(64, 64))
    spec, image = spec.cuda(), image.cuda()
    net.train()
    a1, _ = net(spec.unsqueeze(1), image)
    a2, _ = net(spec.unsqueeze(1) * 2, image)
    with pytest.raises(RuntimeError, match="overwritten"):
        a1.sum().backward()
    a2.sum().backward()                                             # the latest forward is fine


def test_plan_is_rebuilt_when_parameters_are_rehomed(built_lib):
    """plans cache raw parameter addresses; re-homing a parameter (p.data = ...) must invalidate them."""
    net, _ = _model()
    spec, image, _ = orc.synthetic_av_batch(2, 5, spec_hw=(65, 48), image_hw=(64, 64))
    spec, image = spec.cuda(), image.cuda()
    net.eval()
    with torch.no_grad():
        a0, _ = net(spec.unsqueeze(1), image)
        w = net.audio_net.layer1[0].conv1.weight
        w.data = (w.data * 0.5).clone()                             # new memory, new values
        a1, _ = net(spec.unsqueeze(1), image)
        w.data = (w.data * 2.0).clone()
        a2, _ = net(spec.unsqueeze(1), image)
    assert relf(a1.cpu(), a0.cpu()) > 1e-3                          # the new weights were really used
    assert relf(a2.cpu(), a0.cpu()) < 1e-6
