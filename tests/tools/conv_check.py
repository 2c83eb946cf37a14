"""Bring-up check of the tcgen05 implicit-GEMM convolutions against torch (fp32, TF32 off).
Prints per case the Frobenius-relative error of fprop / dgrad / wgrad and a timing."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def relf(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def case(N, H, W, Cin, Cout, R, stride, which="fdw", time_it=False):
    pad = R // 2
    g = torch.Generator(device="cuda").manual_seed(N + H + Cin + Cout + R)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, R, R, device="cuda", generator=g) * (1.0 / (Cin * R * R) ** 0.5)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wn = w.permute(0, 2, 3, 1).contiguous()
    x.requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    y = F.conv2d(x, wr, None, stride, pad)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    res = {}
    if "f" in which:
        yn = ops.conv2d_fprop(xn, wn, stride, pad)
        torch.cuda.synchronize()
        res["fprop"] = relf(yn.permute(0, 3, 1, 2), y.detach())
    if "d" in which:
        dxn = ops.conv2d_dgrad(dyn, wn, xn.shape, stride, pad)
        torch.cuda.synchronize()
        res["dgrad"] = relf(dxn.permute(0, 3, 1, 2), x.grad)
        acc = torch.ones_like(dxn)
        ops.conv2d_dgrad(dyn, wn, xn.shape, stride, pad, out=acc, accumulate=True)
        res["dgrad_acc"] = relf((acc - 1).permute(0, 3, 1, 2), x.grad)
    if "w" in which:
        dwn = ops.conv2d_wgrad(xn, dyn, wn.shape, stride, pad)
        torch.cuda.synchronize()
        res["wgrad"] = relf(dwn.permute(0, 3, 1, 2), wr.grad)
    msg = "N%d %dx%d Cin%d Cout%d k%d s%d: " % (N, H, W, Cin, Cout, R, stride) + " ".join(
        "%s=%.2e" % kv for kv in res.items())
    if time_it:
        flops = 2.0 * y.numel() * Cin * R * R
        for name, fn in (("fprop", lambda: ops.conv2d_fprop(xn, wn, stride, pad)),
                         ("dgrad", lambda: ops.conv2d_dgrad(dyn, wn, xn.shape, stride, pad)),
                         ("wgrad", lambda: ops.conv2d_wgrad(xn, dyn, wn.shape, stride, pad)),
                         ("cudnn_fprop_tf32", None)):
            if fn is None:
                torch.backends.cudnn.allow_tf32 = True
                xc = x.detach().contiguous(memory_format=torch.channels_last)
                wc = w.contiguous(memory_format=torch.channels_last)
                fn = lambda: F.conv2d(xc, wc, None, stride, pad)   # noqa: E731
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 10
            msg += " | %s %.3f ms %.0f TF" % (name, t, flops / t / 1e9)
        torch.backends.cudnn.allow_tf32 = False
    print(msg, flush=True)
    return res


if __name__ == "__main__":
    t0 = time.time()
    quick = [(2, 8, 8, 64, 64, 1, 1), (2, 8, 8, 64, 64, 3, 1), (1, 16, 16, 64, 128, 3, 2), (2, 9, 6, 128, 128, 3, 1),
             (3, 14, 14, 128, 256, 1, 2), (2, 7, 7, 256, 512, 3, 2), (1, 17, 12, 512, 512, 3, 1)]
    which = sys.argv[1] if len(sys.argv) > 1 else "fdw"
    for c in quick:
        case(*c, which=which)
    if "t" in which:
        for c in [(128, 56, 56, 64, 64, 3, 1), (128, 56, 56, 64, 128, 3, 2), (128, 28, 28, 128, 128, 3, 1),
                  (128, 14, 14, 256, 256, 3, 1), (128, 7, 7, 512, 512, 3, 1), (64, 65, 47, 64, 64, 3, 1),
                  (128, 56, 56, 64, 128, 1, 2)]:
            case(*c, which=which.replace("t", ""), time_it=True)
    print("done in %.1fs" % (time.time() - t0))
