"""Forward / backward time of the fused attention core at the benchmark shapes (CUDA events, L2 flushed between launches).

    python tests/tools/attn_time.py            # MLA_ATTN_TC=0 for the mma.sync forward
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mla_b200  # noqa: E402,F401
from mla_b200 import m3ae  # noqa: E402


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (B, S, H, Dh, masked) in [(64, 257, 12, 64, False), (64, 257, 12, 64, True), (64, 512, 12, 64, False)]:
        qkv = (torch.randn(B, S, 3 * H * Dh, device=dev) * 0.8).requires_grad_(True)
        dout = torch.randn(B, S, H * Dh, device=dev)
        mask = None
        if masked:
            n_valid = torch.randint(S // 2, S + 1, (B,))
            mask = (torch.arange(S)[None, :] >= n_valid[:, None]).float().to(dev)
        tf, tb = [], []
        for it in range(8):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            out = m3ae._AttentionFn.apply(qkv, mask, H, Dh ** -0.5)
            e1.record()
            out.backward(dout)
            e2.record()
            torch.cuda.synchronize()
            if it >= 3:
                tf.append(e0.elapsed_time(e1) * 1e3)
                tb.append(e1.elapsed_time(e2) * 1e3)
        f, b = statistics.median(tf), statistics.median(tb)
        fl = 4.0 * B * H * S * S * Dh
        print("B %d S %d H %d Dh %d masked %s: forward (cast + attention) %.1f us = %.0f TF/s; backward %.1f us = %.0f TF/s (10 B H S^2 Dh)"
              % (B, S, H, Dh, masked, f, fl / f / 1e6, b, 2.5 * fl / b / 1e6))


if __name__ == "__main__":
    main()
