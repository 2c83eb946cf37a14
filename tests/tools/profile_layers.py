"""Per-layer times of the 2-byte convolution kernels on the ResNet-18 shapes of the bench batch (B = 64: 128 visual frames,
64 spectrograms), L2 flushed between launches, CUDA events, median of 5.   python tests/tools/profile_layers.py [visual|audio]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import _lib  # noqa: E402

L = _lib.lib()
FLUSH = None


def st():
    return torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    ts = []
    for _ in range(reps):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def layer(N, H, W, Cin, Cout, R, stride):
    pad = R // 2
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(H + Cin)
    x16 = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
    w = torch.randn(Cout, R, R, Cin, device="cuda", generator=g) * (1.0 / (Cin * R * R) ** 0.5)
    w16 = w.half()
    wt16 = torch.empty(Cin, R, R, Cout, dtype=torch.float16, device="cuda")
    assert L.mla_filter_transpose16(w.data_ptr(), wt16.data_ptr(), Cout, R * R, Cin, 0, st()) == 0
    dy16 = torch.randn(N, OH, OW, Cout, device="cuda", generator=g).half()
    one = torch.ones(1, device="cuda")
    y = torch.empty(N, OH, OW, Cout, device="cuda")
    dx = torch.zeros(N, H, W, Cin, device="cuda")
    dw = torch.empty(Cout, R, R, Cin, device="cuda")
    nb = L.mla_conv2d_wgrad16_workspace_bytes(N, H, W, Cin, Cout, R, R, stride, pad)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    part = torch.zeros(L.mla_conv2d_fprop16_stat_tiles(N, H, W, Cin, Cout, R, R, stride, pad) * 2 * Cout, device="cuda")

    def chk(rc):
        assert rc == 0, rc
    t = [timed(lambda: chk(L.mla_conv2d_fprop16(x16.data_ptr(), w16.data_ptr(), y.data_ptr(), N, H, W, Cin, Cout, R, R, stride,
                                                pad, part.data_ptr(), st()))),
         timed(lambda: chk(L.mla_conv2d_dgrad16_f16(dy16.data_ptr(), wt16.data_ptr(), one.data_ptr(), dx.data_ptr(), N, H, W, Cin,
                                                    Cout, R, R, stride, pad, 0, st()))),
         timed(lambda: chk(L.mla_conv2d_dgrad16_f16(dy16.data_ptr(), wt16.data_ptr(), one.data_ptr(), dx.data_ptr(), N, H, W, Cin,
                                                    Cout, R, R, stride, pad, 1, st()))),
         timed(lambda: chk(L.mla_conv2d_wgrad16_f16(x16.data_ptr(), dy16.data_ptr(), one.data_ptr(), dw.data_ptr(), N, H, W, Cin,
                                                    Cout, R, R, stride, pad, ws.data_ptr(), nb, st())))]
    fl = 2.0 * N * OH * OW * Cin * Cout * R * R
    print("N%-3d %3dx%-3d %3d->%-3d %dx%d/%d: " % (N, H, W, Cin, Cout, R, R, stride) +
          " | ".join("%s %.3f ms %4.0f TF" % (n, v, fl / v / 1e9) for n, v in zip(("fprop+stats", "dgrad", "dgrad+=", "wgrad"), t)))
    return t


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "visual"
    only = set(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else None      # e.g. "0,6": conv indices to run
    N, h, w = (128, 56, 56) if which == "visual" else (64, 65, 47)
    tot = [0.0] * 4
    cin = 64
    idx = -1
    for li, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2))):
        ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        convs = [(N, h, w, cin, planes, 3, stride, 1), (N, ho, wo, planes, planes, 3, 1, 3)]
        if stride != 1:
            convs.append((N, h, w, cin, planes, 1, stride, 1))
        for (n_, hh, ww, ci, co, r, s_, mult) in convs:
            idx += 1
            if only is not None and idx not in only:
                continue
            t = layer(n_, hh, ww, ci, co, r, s_)
            for i in range(4):
                tot[i] += t[i] * mult
        h, w, cin = ho, wo, planes
    print("%s encoder, all BasicBlock convolutions: fprop %.3f ms, dgrad %.3f ms, dgrad+= %.3f ms, wgrad %.3f ms" % (
        which, *tot))


if __name__ == "__main__":
    main()
