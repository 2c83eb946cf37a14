"""How far does torch's OWN cuDNN-TF32 path sit from its fp32 path on the encoder gradients?
(Calibrates the gradient tolerance: ReLU-mask flips make the Frobenius error of gradients ~sqrt(P(flip))
per layer for ANY TF32 implementation.)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mla_b200  # noqa: E402
from oracle import mla_oracle as orc  # noqa: E402


def relf(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def grads(B, hw, img, tf32, seed=3):
    torch.backends.cudnn.allow_tf32 = tf32
    dev = torch.device("cuda:0")
    args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=True, lorb="base", modal3=False, clip=False)
    mla_b200.setup_seed(0)
    net = mla_b200.AVClassifier(args).apply(mla_b200.weight_init)
    sd = {k: v.detach().clone().to(dev) for k, v in net.state_dict().items()}
    for k, v in sd.items():
        if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    spec, image, _ = orc.synthetic_av_batch(B, seed, spec_hw=hw, image_hw=(img, img))
    g = torch.Generator().manual_seed(seed + 1)
    da, dv = torch.randn(B, 512, generator=g).to(dev) / B, torch.randn(B, 512, generator=g).to(dev) / B
    a, v = orc.av_forward(sd, spec.to(dev).unsqueeze(1), image.to(dev), training=True)
    a.backward(da)
    v.backward(dv)
    return a.detach(), v.detach(), {k: t.grad for k, t in sd.items() if t.grad is not None}


for B, hw, img in [(2, (65, 48), 64), (8, (257, 188), 224)]:
    a0, v0, g0 = grads(B, hw, img, False)
    a1, v1, g1 = grads(B, hw, img, True)
    errs = sorted(relf(g1[k], g0[k]) for k in g0 if not k.startswith("fusion"))
    cos = sorted(float(torch.nn.functional.cosine_similarity(g1[k].flatten().double(), g0[k].flatten().double(), dim=0))
                 for k in g0 if not k.startswith("fusion"))
    print("cuDNN-TF32 vs fp32, B=%d: feat a %.2e v %.2e | grad rel-F median %.2e max %.2e | cosine min %.4f median %.4f" % (
        B, relf(a1, a0), relf(v1, v0), errs[len(errs) // 2], errs[-1], cos[0], cos[len(cos) // 2]))
