"""Each memory-bound encoder kernel vs torch (fp32) on the GPU."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import _lib  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
L = _lib.lib()
st = lambda: _lib.stream_ptr()   # noqa: E731
P = lambda t: None if t is None else t.data_ptr()   # noqa: E731


def relf(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def check_bn(N, H, W, C, relu_mask=True, res=True):
    dev = "cuda"
    M = N * H * W
    y = torch.randn(M, C, device=dev) * 2 + 0.5
    gamma = torch.rand(C, device=dev) + 0.5
    beta = torch.randn(C, device=dev) * 0.1
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    rm2, rv2 = rm.clone(), rv.clone()
    mean, invstd, scale, shift = (torch.empty(C, device=dev) for _ in range(4))
    ws = torch.zeros(L.mla_bn_workspace_bytes(M, C), dtype=torch.uint8, device=dev)
    rc = L.mla_bn_train_stats(P(y), M, C, P(gamma), P(beta), P(rm), P(rv), 0.1, 1e-5, P(mean), P(invstd), P(scale), P(shift),
                              P(ws), ws.numel(), st())
    assert rc == 0, rc
    yt = y.clone().requires_grad_(True)
    gt, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    idn = torch.randn(M, C, device=dev)
    o = F.batch_norm(yt.view(N, H, W, C).permute(0, 3, 1, 2), rm2, rv2, gt, bt, True, 0.1, 1e-5).permute(0, 2, 3, 1).reshape(M, C)
    z = o + idn if res else o
    z = F.relu(z) if relu_mask else z
    out = torch.empty(M, C, device=dev)
    rc = L.mla_bn_apply(P(y), P(scale), P(shift), P(idn) if res else None, None, None, 1 if relu_mask else 0, P(out), M, C, st())
    assert rc == 0, rc
    dz = torch.randn(M, C, device=dev)
    z.backward(dz)
    dy, g = torch.empty(M, C, device=dev), torch.empty(M, C, device=dev)
    dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
    rc = L.mla_bn_backward(P(dz), P(out) if relu_mask else None, P(y), P(mean), P(invstd), P(gamma), M, C, P(dg), P(db), P(dy),
                           P(g), P(ws), ws.numel(), st())
    assert rc == 0, rc
    torch.cuda.synchronize()
    print("bn N%d %dx%d C%d: mean %.1e rm %.1e rv %.1e out %.1e dgamma %.1e dbeta %.1e dy %.1e g %.1e" % (
        N, H, W, C, relf(mean, y.mean(0)), relf(rm, rm2), relf(rv, rv2), relf(out, z.detach()), relf(dg, gt.grad),
        relf(db, bt.grad), relf(dy, yt.grad), relf(g, dz * (z.detach() > 0) if relu_mask else dz)))


def check_stem(B, T, Cin, H, W):
    dev = "cuda"
    x = torch.randn(B, Cin, T, H, W, device=dev) if T > 1 else torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(64, Cin, 7, 7, device=dev) * 0.1
    wk = w.permute(0, 2, 3, 1).contiguous()
    N = B * T
    OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    Kp = {1: 64, 3: 192}[Cin]
    col = torch.empty(N * OH * OW, Kp, device=dev)
    HW = H * W
    sB, sT, sC = (Cin * T * HW, HW, T * HW) if T > 1 else (Cin * HW, 0, HW)
    assert L.mla_stem_im2col(P(x), P(col), N, T, sB, sT, sC, Cin, H, W, 7, 7, 2, 3, Kp, st()) == 0
    wpad = torch.empty(64, Kp, device=dev)
    assert L.mla_pad_rows(P(wk), P(wpad), 64, 49 * Cin, Kp, 0, st()) == 0
    y = torch.empty(N, OH, OW, 64, device=dev)
    assert L.mla_conv2d_fprop(P(col), P(wpad), P(y), N, OH, OW, Kp, 64, 1, 1, 1, 0, st()) == 0
    xr = x.permute(0, 2, 1, 3, 4).reshape(N, Cin, H, W) if T > 1 else x
    ref = F.conv2d(xr, w, None, 2, 3)
    # fused bn+relu+maxpool
    scale, shift = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.2
    PH, PW = (OH + 2 - 3) // 2 + 1, (OW + 2 - 3) // 2 + 1
    p = torch.empty(N, PH, PW, 64, device=dev)
    idx = torch.empty(N, PH, PW, 64, dtype=torch.uint8, device=dev)
    assert L.mla_bn_relu_maxpool(P(y), P(scale), P(shift), P(p), P(idx), N, OH, OW, 64, st()) == 0
    yr = y.permute(0, 3, 1, 2).clone().requires_grad_(True)
    a = F.relu(yr * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    pr = F.max_pool2d(a, 3, 2, 1)
    dp = torch.randn_like(pr)
    pr.backward(dp)
    g = torch.empty(N, OH, OW, 64, device=dev)
    dpn = dp.permute(0, 2, 3, 1).contiguous()
    assert L.mla_maxpool_relu_backward(P(dpn), P(p), P(idx), P(g), N, OH, OW, 64, st()) == 0
    torch.cuda.synchronize()
    # reference grad wrt a-relu input: yr.grad = scale * g  -> g = yr.grad / scale
    gref = (yr.grad / scale.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    print("stem B%d T%d Cin%d %dx%d: conv %.1e pool %.1e poolbwd %.1e" % (
        B, T, Cin, H, W, relf(y.permute(0, 3, 1, 2), ref), relf(p.permute(0, 3, 1, 2), pr.detach()), relf(g, gref)))


def check_avgpool(B, rows, C):
    dev = "cuda"
    fm = torch.randn(B * rows, C, device=dev)
    feat = torch.empty(B, C, device=dev)
    assert L.mla_avgpool_forward(P(fm), P(feat), B, rows, C, st()) == 0
    df = torch.randn(B, C, device=dev)
    dfm = torch.empty(B * rows, C, device=dev)
    assert L.mla_avgpool_backward(P(df), P(dfm), B, rows, C, st()) == 0
    torch.cuda.synchronize()
    ref = fm.view(B, rows, C).mean(1)
    dref = (df / rows).view(B, 1, C).expand(B, rows, C).reshape(B * rows, C)
    print("avgpool B%d rows%d C%d: fwd %.1e bwd %.1e" % (B, rows, C, relf(feat, ref), relf(dfm, dref)))


if __name__ == "__main__":
    check_bn(2, 5, 3, 64)
    check_bn(4, 9, 6, 512)
    check_bn(16, 28, 28, 128, relu_mask=True, res=False)
    check_bn(8, 56, 56, 64, relu_mask=False, res=False)
    check_bn(3, 7, 7, 256)
    check_stem(2, 1, 1, 65, 48)
    check_stem(2, 2, 3, 64, 64)
    check_stem(1, 1, 1, 257, 188)
    check_avgpool(4, 54, 512)
    check_avgpool(3, 98, 512)
