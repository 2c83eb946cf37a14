"""Native ResNet-18 encoder (forward + backward) vs the oracle's torch-functional restatement run on
the same GPU in full fp32 (cuDNN TF32 off): Frobenius-relative errors of features, BN running
stats and every parameter gradient."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mla_b200  # noqa: E402
from oracle import mla_oracle as orc  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def relf(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def use_exact_convs():
    """Replace the plan's three conv launchers by torch fp32 convolutions (TEST ONLY): isolates the
    orchestration (BN backward, residual accumulation, buffer reuse, stem) from TF32 rounding, whose
    ReLU-mask flips alone put a ~sqrt(P(flip)) ~ 3e-2 per layer Frobenius error on gradients."""
    import torch.nn.functional as F
    from mla_b200 import encoder_engine as ee

    def nchw(t, N, H, W, C):
        return t.view(N, H, W, C).permute(0, 3, 1, 2)

    def wt(w, Cout, R, Cin):
        # 4-D parameters are logically OIHW whatever their strides; 2-D [Cout][K] buffers are KRSC memory
        return w.detach() if w.dim() == 4 else w.detach().view(Cout, R, R, Cin).permute(0, 3, 1, 2)

    def conv(self, x, w, y, N, H, W, Cin, Cout, R, stride, pad, st, k_alg=None):  # noqa: D401
        y.copy_(F.conv2d(nchw(x, N, H, W, Cin), wt(w, Cout, R, Cin), None, stride, pad).permute(0, 2, 3, 1))

    def dgrad(self, dy, w, dx, N, H, W, Cin, Cout, R, stride, pad, acc, st):
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
        g = torch.nn.grad.conv2d_input((N, Cin, H, W), wt(w, Cout, R, Cin), nchw(dy, N, OH, OW, Cout), stride, pad)
        g = g.permute(0, 2, 3, 1)
        dx.view(N, H, W, Cin).copy_(dx.view(N, H, W, Cin) + g if acc else g)

    def wgrad(self, x, dy, dw, N, H, W, Cin, Cout, R, stride, pad, st, k_alg=None):
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
        g = torch.nn.grad.conv2d_weight(nchw(x, N, H, W, Cin), (Cout, Cin, R, R), nchw(dy, N, OH, OW, Cout), stride, pad)
        if dw.dim() == 4:
            dw.copy_(g)
        else:
            dw.view(Cout, R, R, Cin).copy_(g.permute(0, 2, 3, 1))
    def conv_bn(self, x, w, y, N, H, W, Cin, Cout, R, stride, pad, b, training, st, k_alg=None):
        conv(self, x, w, y, N, H, W, Cin, Cout, R, stride, pad, st)
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
        self._bn_coeffs(y, N * OH * OW, b, training, st)
    ee.USE_GRAPHS = False          # torch convolutions inside the launch sequence are not graph-captured here
    ee.USE_F16 = False             # the patched launchers are the fp32-operand ones
    ee.ResNetPlan._conv, ee.ResNetPlan._dgrad, ee.ResNetPlan._wgrad = conv, dgrad, wgrad
    ee.ResNetPlan._conv_bn = conv_bn


def run(B, hw, img, seed=3, verbose=False):
    dev = torch.device("cuda:0")
    args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=True, lorb="base", modal3=False, clip=False)
    mla_b200.setup_seed(0)
    net = mla_b200.AVClassifier(args).apply(mla_b200.weight_init)
    sd = {k: v.detach().clone().to(dev) for k, v in net.state_dict().items()}
    net = net.to(dev)
    for k, v in sd.items():
        if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    spec, image, _ = orc.synthetic_av_batch(B, seed, spec_hw=hw, image_hw=(img, img))
    spec, image = spec.to(dev), image.to(dev)
    g = torch.Generator().manual_seed(seed + 1)
    da, dv = torch.randn(B, 512, generator=g).to(dev) / B, torch.randn(B, 512, generator=g).to(dev) / B
    # oracle (torch fp32 on GPU)
    ra, rv = orc.av_forward(sd, spec.unsqueeze(1), image, training=True)
    ra.backward(da)
    rv.backward(dv)
    # native
    net.train()
    a, v = net(spec.unsqueeze(1), image)
    a.backward(da)
    v.backward(dv)
    torch.cuda.synchronize()
    print("B=%d spec=%s img=%d  feat a %.2e v %.2e" % (B, hw, img, relf(a, ra), relf(v, rv)))
    worst = []
    for name, p in net.named_parameters():
        if name.startswith("fusion_module"):
            continue
        e = relf(p.grad, sd[name].grad)
        worst.append((e, name))
        if verbose:
            print("   %-45s %.2e" % (name, e))
    worst.sort(reverse=True)
    print("   worst grads:", ", ".join("%s %.1e" % (n, e) for e, n in worst[:6]))
    print("   median grad err %.2e" % worst[len(worst) // 2][0])
    for k in ("audio_net.bn1.running_mean", "audio_net.bn1.running_var", "visual_net.layer4.1.bn2.running_var",
              "visual_net.layer2.0.downsample.1.running_mean"):
        print("   %-45s %.2e" % (k, relf(net.state_dict()[k], sd[k])))
    net.eval()
    with torch.no_grad():
        a, v = net(spec.unsqueeze(1), image)
        ra, rv = orc.av_forward(sd, spec.unsqueeze(1), image, training=False)
    print("   eval feat a %.2e v %.2e" % (relf(a, ra), relf(v, rv)))
    return net, spec, image, da, dv


if __name__ == "__main__":
    verbose = "-v" in sys.argv
    if "--exact-conv" in sys.argv:
        use_exact_convs()
    run(2, (65, 48), 64, verbose=verbose)
    run(4, (97, 64), 96)
    net, spec, image, da, dv = run(8, (257, 188), 224)
    if "-t" in sys.argv:
        B = 64
        spec, image, _ = orc.synthetic_av_batch(B, 5)
        spec, image = spec.cuda(), image.cuda()
        da = torch.randn(B, 512, device="cuda") / B
        net.train()
        for it in range(5):
            if it == 2:
                torch.cuda.synchronize(); t0 = time.time()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
            a, v = net(spec.unsqueeze(1), image)
            if it == 4:
                e1.record()
            a.backward(da)
            v.backward(da)
        e2.record()
        torch.cuda.synchronize()
        print("B=64 full size: 3 iters fwd+bwd %.2f ms/iter (wall %.2f); last bwd %.2f ms" % (
            e0.elapsed_time(e2) / 3, (time.time() - t0) * 1e3 / 3, e1.elapsed_time(e2)))
