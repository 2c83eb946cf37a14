"""Launch gs_project_kernel at the sweep points bench.py reports (target of `ncu --set full -k regex:gs_project`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for (B, D, C) in [(4096, 2048, 6), (64, 512, 6)]:
    feat = torch.randn(B, D, device=dev).relu()
    grad = torch.randn(C, D, device=dev)
    P = torch.eye(D, device=dev)
    for _ in range(2):
        ops.gs_project(P, grad, 0.05, feat=feat)
torch.cuda.synchronize()
print("ok")
