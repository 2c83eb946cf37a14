"""Odd small shapes of the kernels rewritten in round 2 (GS projection, fusion, tiled head, frame producer) in one quick
script: a crash / hang / non-finite canary (and a compute-sanitizer target where that tool is available — it is closed on
the shared B200 pool).   python tests/tools/sanitize_new.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import ops  # noqa: E402
from mla_b200.dataset import FrameBatchProducer  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for (B, D, C) in [(64, 512, 6), (7, 132, 101), (300, 2048, 6), (33, 768, 17), (5, 4, 1)]:
        P = torch.eye(D, device=dev)
        g = torch.randn(C, D, device=dev)
        ops.gs_project(P, g, 0.05, feat=torch.randn(B, D, device=dev).relu())
        ops.gs_project(P, g, 0.05, feat_sum=torch.randn(D, device=dev), inv_batch=1.0 / B)
        ops.gs_project(P, None, 0.05, feat=torch.randn(B, D, device=dev))
    for (B, D, C) in [(64, 768, 101), (130, 132, 65), (1, 2048, 64), (700, 64, 129), (2, 4, 20), (64, 512, 6), (9, 12, 16)]:
        o = ops.head_ce(torch.randn(B, D, device=dev), torch.randn(C, D, device=dev) * 0.05, torch.zeros(C, device=dev),
                        torch.randint(0, C, (B,), device=dev))
        ops.head_ce(torch.randn(B, D, device=dev), torch.randn(C, D, device=dev) * 0.05, torch.zeros(C, device=dev),
                    torch.randint(0, C, (B,), device=dev), need_grad=False)
        assert torch.isfinite(o["dW"]).all()
    for (B, C, M) in [(64, 6, 2), (257, 101, 3), (4096, 101, 3), (1, 6, 2), (33, 1, 4), (1000, 37, 2), (70, 1030, 2)]:
        outs = [torch.randn(B, C, device=dev) for _ in range(M)]
        lab = torch.randint(0, C, (B,), device=dev)
        hits = torch.zeros(M + 1, C, dtype=torch.int64, device=dev)
        num = torch.zeros(C, dtype=torch.int64, device=dev)
        ops.fuse_eval(outs, lab, hits=hits, num=num)
        ops.fuse_eval(outs, dynamic=False, fixed_w=[1.0 / M] * M)
        assert int(num.sum()) == B
    rng = np.random.default_rng(0)
    frames = [[rng.integers(0, 256, (int(rng.integers(30, 200)), int(rng.integers(30, 200)), 3), dtype=np.uint8) for _ in range(2)]
              for _ in range(3)]
    torch.manual_seed(0)
    out = FrameBatchProducer(64, "train")(frames)
    out2 = FrameBatchProducer(224, "test")(frames)
    out3 = FrameBatchProducer(96, "center", interpolation="bicubic")(frames)
    assert torch.isfinite(out).all() and torch.isfinite(out2).all() and torch.isfinite(out3).all()
    torch.cuda.synchronize()
    print("sanitize_new ok")


if __name__ == "__main__":
    main()
