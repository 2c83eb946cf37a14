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
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
