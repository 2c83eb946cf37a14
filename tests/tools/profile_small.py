"""Launch the three small kernels at representative sizes (for ncu captures and quick timing)."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mla_b200 import ops  # noqa: E402


def timeit(fn, n=20, flush=None):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"gs": [], "head": [], "fuse": []}
    for D in (512, 768, 1024, 2048):
        for B in (64, 256, 1024, 4096):
            for C in (6, 101):
                feat = torch.randn(B, D, device=dev).relu()
                grad = torch.randn(C, D, device=dev)
                P = torch.eye(D, device=dev)
                fs = feat.sum(0)
                t_raw = timeit(lambda: ops.gs_project(P, grad, 0.05, feat=feat), flush=flush)
                t_sum = timeit(lambda: ops.gs_project(P, grad, 0.05, feat_sum=fs, inv_batch=1.0 / B), flush=flush)
                t_hot = timeit(lambda: ops.gs_project(P, grad, 0.05, feat=feat))
                nbytes = 4 * (B * D + 2 * D * D + 2 * C * D)
                res["gs"].append(dict(B=B, D=D, C=C, us_cold=t_raw, us_sum_cold=t_sum, us_hot=t_hot, bytes=nbytes,
                                      gbs_cold=nbytes / t_raw / 1e3))
    for (B, D, C) in [(64, 512, 6), (64, 768, 101), (64, 768, 4), (4096, 2048, 101)]:
        feat = torch.randn(B, D, device=dev).relu()
        W = torch.randn(C, D, device=dev) * 0.05
        b = torch.zeros(C, device=dev)
        lab = torch.randint(0, C, (B,), device=dev)
        o = {}
        t = timeit(lambda: ops.head_ce(feat, W, b, lab, out=o))
        res["head"].append(dict(B=B, D=D, C=C, us_hot=t, bytes=4 * (2 * B * D + 3 * C * D + 2 * B * C)))
    for (B, C, M) in [(64, 6, 2), (64, 101, 3), (4096, 6, 2), (4096, 101, 3)]:
        outs = [torch.randn(B, C, device=dev) for _ in range(M)]
        lab = torch.randint(0, C, (B,), device=dev)
        hits = torch.zeros(M + 1, C, dtype=torch.int64, device=dev)
        num = torch.zeros(C, dtype=torch.int64, device=dev)
        t = timeit(lambda: ops.fuse_eval(outs, lab, hits=hits, num=num))
        res["fuse"].append(dict(B=B, C=C, M=M, us_hot=t, bytes=4 * (M + 1) * B * C + 8 * B))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
