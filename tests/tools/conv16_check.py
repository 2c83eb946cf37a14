"""Bring-up check of the 2-byte operand convolutions (fprop16: fp16 x fp16, dgrad16: bf16 dy x transposed fp16 filter)
against torch fp32 convolutions of the SAME rounded operands, plus timings."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import _lib  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
L = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream   # noqa: E731


def relf(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def case(N, H, W, Cin, Cout, R, stride, time_it=False):
    pad = R // 2
    g = torch.Generator(device="cuda").manual_seed(N + H + Cin + Cout + R)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, R, R, device="cuda", generator=g) * (1.0 / (Cin * R * R) ** 0.5)
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(N, Cout, OH, OW, device="cuda", generator=g)
    x16 = x.permute(0, 2, 3, 1).contiguous().half()
    w_krsc = w.permute(0, 2, 3, 1).contiguous()
    w16 = torch.empty(Cout, R, R, Cin, dtype=torch.float16, device="cuda")
    assert L.mla_cast16(w_krsc.data_ptr(), w16.data_ptr(), w_krsc.numel(), 0, st()) == 0
    wt16 = torch.empty(Cin, R, R, Cout, dtype=torch.bfloat16, device="cuda")
    assert L.mla_filter_transpose16(w_krsc.data_ptr(), wt16.data_ptr(), Cout, R * R, Cin, 1, st()) == 0
    dy16 = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
    y = torch.empty(N, OH, OW, Cout, device="cuda")
    dx = torch.empty(N, H, W, Cin, device="cuda")
    rc = L.mla_conv2d_fprop16(x16.data_ptr(), w16.data_ptr(), y.data_ptr(), N, H, W, Cin, Cout, R, R, stride, pad, None, st())
    assert rc == 0, rc
    rc = L.mla_conv2d_dgrad16(dy16.data_ptr(), wt16.data_ptr(), dx.data_ptr(), N, H, W, Cin, Cout, R, R, stride, pad, 0, st())
    assert rc == 0, rc
    torch.cuda.synchronize()
    # references on the rounded operands (fp32 math)
    xr = x16.float().permute(0, 3, 1, 2)
    wr = w16.float().permute(0, 3, 1, 2)
    yr = F.conv2d(xr, wr, None, stride, pad)
    wbr = wt16.float().permute(3, 0, 1, 2)            # [Cout, Cin, R, R] from the bf16 transposed filter
    dxr = torch.nn.grad.conv2d_input((N, Cin, H, W), wbr, dy16.float().permute(0, 3, 1, 2), stride, pad)
    wt_ok = torch.equal(wt16, w_krsc.permute(3, 1, 2, 0).contiguous().bfloat16())
    acc = torch.ones_like(dx)
    L.mla_conv2d_dgrad16(dy16.data_ptr(), wt16.data_ptr(), acc.data_ptr(), N, H, W, Cin, Cout, R, R, stride, pad, 1, st())
    torch.cuda.synchronize()
    msg = "N%d %dx%d Cin%d Cout%d k%d s%d: fprop16=%.2e dgrad16=%.2e dgrad16_acc=%.2e transpose_ok=%s" % (
        N, H, W, Cin, Cout, R, stride, relf(y.permute(0, 3, 1, 2), yr), relf(dx.permute(0, 3, 1, 2), dxr),
        relf((acc - 1).permute(0, 3, 1, 2), dxr), wt_ok)
    xb = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    dw = torch.empty(Cout, R, R, Cin, device="cuda")
    nb = L.mla_conv2d_wgrad16_workspace_bytes(N, H, W, Cin, Cout, R, R, stride, pad)
    wsb = torch.empty(nb, dtype=torch.uint8, device="cuda")
    rc = L.mla_conv2d_wgrad16(xb.data_ptr(), dy16.data_ptr(), dw.data_ptr(), N, H, W, Cin, Cout, R, R, stride, pad, wsb.data_ptr(),
                              nb, st())
    assert rc == 0, rc
    torch.cuda.synchronize()
    dwr = torch.nn.grad.conv2d_weight(xb.float().permute(0, 3, 1, 2), (Cout, Cin, R, R), dy16.float().permute(0, 3, 1, 2),
                                      stride, pad)
    msg += " wgrad16=%.2e" % relf(dw.permute(0, 3, 1, 2), dwr)
    if time_it:
        flops = 2.0 * y.numel() * Cin * R * R
        for name, fn in (("fprop16", lambda: L.mla_conv2d_fprop16(x16.data_ptr(), w16.data_ptr(), y.data_ptr(), N, H, W, Cin, Cout,
                                                                    R, R, stride, pad, None, st())),
                         ("dgrad16", lambda: L.mla_conv2d_dgrad16(dy16.data_ptr(), wt16.data_ptr(), dx.data_ptr(), N, H, W, Cin,
                                                                    Cout, R, R, stride, pad, 0, st())),
                         ("wgrad16", lambda: L.mla_conv2d_wgrad16(xb.data_ptr(), dy16.data_ptr(), dw.data_ptr(), N, H, W, Cin,
                                                                    Cout, R, R, stride, pad, wsb.data_ptr(), nb, st()))):
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
    print(msg, flush=True)


if __name__ == "__main__":
    for c in [(2, 8, 8, 64, 64, 1, 1), (2, 8, 8, 64, 64, 3, 1), (1, 16, 16, 64, 128, 3, 2), (2, 9, 6, 128, 128, 3, 1),
              (3, 14, 14, 128, 256, 1, 2), (2, 7, 7, 256, 512, 3, 2), (1, 17, 12, 512, 512, 3, 1)]:
        case(*c)
    for c in [(128, 56, 56, 64, 64, 3, 1), (128, 56, 56, 64, 128, 3, 2), (128, 28, 28, 128, 128, 3, 1),
              (128, 14, 14, 256, 256, 3, 1), (128, 7, 7, 512, 512, 3, 1), (64, 65, 47, 64, 64, 3, 1)]:
        case(*c, time_it=True)
