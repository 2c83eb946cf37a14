"""The top sweep point of the three small kernels (GS projection, shared head, fusion), two launches each on fresh buffers:
the launch set to put under `ncu --set full -k regex:gs_project|head_|fuse_eval`.   python tests/tools/profile_top.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, D, C = 4096, 2048, 6
    for _ in range(2):
        feat = torch.randn(B, D, device=dev).relu()
        grad = torch.randn(C, D, device=dev)
        P = torch.eye(D, device=dev)
        ops.gs_project(P, grad, 0.05, feat=feat)
    for _ in range(2):
        feat = torch.randn(B, D, device=dev).relu()
        W = torch.randn(C, D, device=dev) * 0.05
        b = torch.zeros(C, device=dev)
        lab = torch.randint(0, C, (B,), device=dev)
        ops.head_ce(feat, W, b, lab, out={})
    for (Bf, Cf, M) in [(4096, 101, 2), (64, 6, 2)]:
        outs = [torch.randn(Bf, Cf, device=dev) for _ in range(M)]
        lab = torch.randint(0, Cf, (Bf,), device=dev)
        hits = torch.zeros(M + 1, Cf, dtype=torch.int64, device=dev)
        num = torch.zeros(Cf, dtype=torch.int64, device=dev)
        ops.fuse_eval(outs, lab, hits=hits, num=num)
    # the tiled large-C head (Food-101 size at the m3ae training batch, and the top of the sweep)
    for (Bh, Dh, Ch) in [(64, 768, 101), (4096, 2048, 101)]:
        feat = torch.randn(Bh, Dh, device=dev).relu()
        ops.head_ce(feat, torch.randn(Ch, Dh, device=dev) * 0.05, torch.zeros(Ch, device=dev),
                    torch.randint(0, Ch, (Bh,), device=dev), out={})
    # the dataset producers: 64 x 2 frames 360 x 480 -> [64, 3, 2, 224, 224]; 64 filterbanks [1024, 128]
    import numpy as np
    Bp, Tp, H, W, S = 64, 2, 360, 480, 224
    desc = np.zeros((Bp * Tp, 14), np.int32)
    for n in range(Bp * Tp):
        desc[n] = (n * H * W * 3, 0, H, W, 0, 0, H, W, 0, n, S, S, 0, 0)
    src = torch.randint(0, 256, (Bp * Tp * H * W * 3,), dtype=torch.uint8, device=dev)
    ops.frames_to_batch(src, torch.from_numpy(desc).to(dev), Bp, Tp, S, H, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    fb = torch.randn(64, 1024, 128, device=dev)
    prm = torch.zeros(64, 6, dtype=torch.int32, device=dev)
    ops.spec_to_batch(fb, prm, torch.zeros(64, device=dev), None, -5.081, 4.4849)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
