"""BatchNorm backward at the visual layer1 size (M = 128*56*56, C = 64): target of ncu captures of channel_reduce_kernel<1>."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import _lib  # noqa: E402

L = _lib.lib()
P = lambda t: None if t is None else t.data_ptr()   # noqa: E731
st = lambda: torch.cuda.current_stream().cuda_stream   # noqa: E731
M, C = 128 * 56 * 56, 64
dz, y = torch.randn(M, C, device="cuda"), torch.randn(M, C, device="cuda")
mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (M * C // 32,), dtype=torch.int32, device="cuda")
mean, invstd, gamma = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.ones(C, device="cuda")
dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
dy16 = torch.empty(M, C, dtype=torch.bfloat16, device="cuda")
ws = torch.zeros(L.mla_bn_workspace_bytes(M, C), dtype=torch.uint8, device="cuda")
for _ in range(3):
    assert L.mla_bn_backward_ex(P(dz), None, P(mask), P(y), P(mean), P(invstd), P(gamma), M, C, P(dg), P(db), None, P(dy16), None,
                                P(ws), ws.numel(), st()) == 0
torch.cuda.synchronize()
print("ok")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(10):
    flush.zero_()
    e0.record()
    L.mla_bn_backward_ex(P(dz), None, P(mask), P(y), P(mean), P(invstd), P(gamma), M, C, P(dg), P(db), None, P(dy16), None,
                         P(ws), ws.numel(), st())
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print("bn backward (reduce + apply), M=%d C=%d: median %.1f us" % (M, C, ts[len(ts) // 2]))
