"""The 2-byte convolution kernels on two ResNet-18 layers at the bench batch (128 visual frames): the launch set to put
under `ncu --set full -k regex:conv16_persistent|conv_gemm`.   python tests/tools/profile_conv16.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import _lib  # noqa: E402

L = _lib.lib()


def st():
    return torch.cuda.current_stream().cuda_stream


def layer(N, H, W, Cin, Cout, reps=3):
    """The entry points the engine really calls (encoder_engine.py:_conv_bn16 / _dgrad16 / _wgrad16): fp16 forward with the
    BatchNorm partial sums, accumulating dgrad and wgrad on the power-of-two-scaled fp16 gradient."""
    g = torch.Generator(device="cuda").manual_seed(H + Cin)
    x16 = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
    w = torch.randn(Cout, 3, 3, Cin, device="cuda", generator=g) * (1.0 / (Cin * 9) ** 0.5)
    w16 = w.half()
    wt16 = w.permute(3, 1, 2, 0).contiguous().half()                  # [Cin][R][S][Cout]
    dy16 = (torch.randn(N, H, W, Cout, device="cuda", generator=g) * 2048).half()
    scale = torch.tensor([2048.0, 1.0 / 2048.0], device="cuda")
    inv = scale.data_ptr() + 4
    y = torch.empty(N, H, W, Cout, device="cuda")
    dx = torch.zeros(N, H, W, Cin, device="cuda")
    dw = torch.empty(Cout, 3, 3, Cin, device="cuda")
    tiles = L.mla_conv2d_fprop16_stat_tiles(N, H, W, Cin, Cout, 3, 3, 1, 1)
    part = torch.empty(max(tiles, 1) * 2 * Cout, device="cuda")
    nb = L.mla_conv2d_wgrad16_workspace_bytes(N, H, W, Cin, Cout, 3, 3, 1, 1)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for r in range(reps):
        ev[0].record()
        assert L.mla_conv2d_fprop16(x16.data_ptr(), w16.data_ptr(), y.data_ptr(), N, H, W, Cin, Cout, 3, 3, 1, 1, part.data_ptr(), st()) == 0
        ev[1].record()
        assert L.mla_conv2d_dgrad16_f16(dy16.data_ptr(), wt16.data_ptr(), inv, dx.data_ptr(), N, H, W, Cin, Cout, 3, 3, 1, 1, 1, st()) == 0
        ev[2].record()
        assert L.mla_conv2d_wgrad16_f16(x16.data_ptr(), dy16.data_ptr(), inv, dw.data_ptr(), N, H, W, Cin, Cout, 3, 3, 1, 1,
                                        ws.data_ptr(), nb, st()) == 0
        ev[3].record()
    torch.cuda.synchronize()
    fl = 2.0 * N * H * W * Cin * Cout * 9
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    print("N%d %dx%d %d->%d: fprop16+stats %.3f ms %.0f TF | dgrad16(+=) %.3f ms %.0f TF | wgrad16 %.3f ms %.0f TF" % (
        N, H, W, Cin, Cout, t[0], fl / t[0] / 1e9, t[1], fl / t[1] / 1e9, t[2], fl / t[2] / 1e9))


layer(128, 56, 56, 64, 64)
layer(128, 28, 28, 128, 128)
layer(128, 14, 14, 256, 256)
