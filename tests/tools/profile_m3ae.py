"""One transformer block (m3ae 'base' width, B=32, S=257, key padding) forward + backward through _BlockFn: the launch set
to put under ncu.   python tests/tools/profile_m3ae.py [B=32] [S=257]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import m3ae  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 257
    torch.manual_seed(0)
    blk = m3ae.Block(768, 12).cuda()
    x = torch.randn(B, S, 768, device="cuda", requires_grad=True)
    dy = torch.randn(B, S, 768, device="cuda")
    n_valid = torch.randint(S // 4, S + 1, (B,))
    mask = (torch.arange(S)[None, :] >= n_valid[:, None]).float().cuda()
    for _ in range(2):
        blk.zero_grad()
        y = blk(x, mask)
        y.backward(dy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = blk(x, mask)
    y.backward(dy)
    e1.record()
    torch.cuda.synchronize()
    flops = 3 * 2 * B * S * 768 * (2304 + 768 + 3072 + 3072) + 18 * B * 12 * S * S * 64
    print("block fwd+bwd B=%d S=%d: %.3f ms, %.1f TFLOP/s (Linear x3 + attention 18 B H S^2 Dh)" % (
        B, S, e0.elapsed_time(e1), flops / e0.elapsed_time(e1) / 1e9))


if __name__ == "__main__":
    main()
