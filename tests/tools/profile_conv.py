"""Launch the tcgen05 convolution kernels once per representative ResNet-18 layer shape (B=64 per GPU:
visual N=128 frames, audio N=64) — the target of `ncu --set full -k regex:conv_gemm` captures — and
print CUDA-event timings (median of 10, L2 flushed) with the achieved TF32 TFLOP/s."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import ops  # noqa: E402

SHAPES = [  # N, H, W, Cin, Cout, R, stride
    (128, 56, 56, 64, 64, 3, 1),     # visual layer1
    (64, 65, 47, 64, 64, 3, 1),      # audio layer1
    (128, 56, 56, 64, 128, 3, 2),    # visual layer2.0.conv1
    (128, 28, 28, 128, 128, 3, 1),   # visual layer2
    (128, 14, 14, 256, 256, 3, 1),   # visual layer3
    (128, 7, 7, 512, 512, 3, 1),     # visual layer4
    (128, 56, 56, 64, 128, 1, 2),    # visual layer2 downsample
]


def main():
    quick = "--once" in sys.argv
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (N, H, W, Cin, Cout, R, stride) in SHAPES:
        pad = R // 2
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
        x = torch.randn(N, H, W, Cin, device=dev)
        w = torch.randn(Cout, R, R, Cin, device=dev)
        dy = torch.randn(N, OH, OW, Cout, device=dev)
        y = torch.empty(N, OH, OW, Cout, device=dev)
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        flops = 2.0 * N * OH * OW * Cout * Cin * R * R
        fns = (("fprop", lambda: ops.conv2d_fprop(x, w, stride, pad, out=y)),
               ("dgrad", lambda: ops.conv2d_dgrad(dy, w, x.shape, stride, pad, out=dx)),
               ("wgrad", lambda: ops.conv2d_wgrad(x, dy, w.shape, stride, pad, out=dw)))
        msg = "N%d %dx%d Cin%d Cout%d k%d s%d:" % (N, H, W, Cin, Cout, R, stride)
        for name, fn in fns:
            if quick:
                fn()
                continue
            ts = []
            for _ in range(2):
                fn()
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = statistics.median(ts)
            msg += "  %s %.3f ms %.0f TF/s" % (name, t, flops / t / 1e9)
        print(msg, flush=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
