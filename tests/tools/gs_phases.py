"""Phase timing of gs_project_kernel (CTA 0's %globaltimer stamps, written to the tail of the workspace) + CUDA-event time.

    python tests/tools/gs_phases.py            # the corners of the BASELINE.json configs[4] sweep
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mla_b200  # noqa: E402
from mla_b200 import _lib, ops  # noqa: E402

NAMES = ["feat partials + grad copy (P rows in flight)", "grid.sync", "r = sum of partials + grid.sync + smem",
         "k = P r (smem)", "grid.sync + update + sum sq", "grid.sync + normalise + write P", "projection"]


def main():
    dev = torch.device("cuda")
    L = _lib.lib()
    for (B, D, C) in [(64, 512, 6), (64, 768, 101), (4096, 512, 6), (64, 2048, 6), (4096, 2048, 6), (4096, 2048, 101),
                      (1024, 1024, 101)]:
        # cold inputs, warm code: rotate over argument sets whose total footprint exceeds twice the L2 (as bench.py does)
        nb = 4 * (B * D + 2 * D * D + 2 * C * D)
        nsets = max(2, min(128, -(-(300 << 20) // nb) + 1))
        sets = [(torch.eye(D, device=dev), torch.randn(C, D, device=dev), torch.randn(B, D, device=dev).relu()) for _ in range(nsets)]
        for i in range(min(3, nsets)):
            ops.gs_project(sets[i][0], sets[i][1], 0.05, feat=sets[i][2])
        ts, phases = [], []
        nbytes = L.mla_gs_project_workspace_bytes(B, D, C)
        for i in range(max(12, nsets)):
            P, grad, feat = sets[i % nsets]
            torch.cuda._sleep(120000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gs_project(P, grad, 0.05, feat=feat); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
            ws = ops._ws[("gs", P.device)].buf
            st = ws[nbytes - 4096:nbytes].view(torch.int64).tolist()
            phases.append([(st[i + 1] - st[i]) / 1e3 for i in range(7)])
            g = (D + max(4, -(-D // 148)) - 1) // max(4, -(-D // 148))
            arr = st[16:16 + g]
            t_start, t_end = st[176:176 + g], st[336:336 + g]
            span = "kernel: first CTA start -> last CTA end %.2f us; CTA starts span %.2f us, ends span %.2f us" % (
                (max(t_end) - min(t_start)) / 1e3, (max(t_start) - min(t_start)) / 1e3, (max(t_end) - min(t_end)) / 1e3)
            proj = "projection detail (CTA 0): normalise + store loop %.2f us, barrier %.2f; tile 0 ready +%.2f, multiplied +%.2f, tile 1 ready +%.2f, multiplied +%.2f, last barrier +%.2f, tail +%.2f" % (
                (st[15] - st[9]) / 1e3, (st[6] - st[15]) / 1e3, (st[10] - st[6]) / 1e3, (st[11] - st[10]) / 1e3,
                (st[12] - st[11]) / 1e3, (st[13] - st[12]) / 1e3, (st[14] - st[13]) / 1e3, (st[7] - st[14]) / 1e3)
            extra = "last barrier: CTA arrivals span %.2f us (CTA0 at +%.2f); after barrier +%.2f us, norm +%.2f us, write +%.2f us" % (
                (max(arr) - min(arr)) / 1e3, (st[5] - min(arr)) / 1e3, (st[8] - max(arr)) / 1e3, (st[9] - st[8]) / 1e3,
                (st[6] - st[9]) / 1e3)
        med = [statistics.median(p[i] for p in phases) for i in range(7)]
        t = statistics.median(ts)
        print("B %5d D %5d C %4d: %7.2f us (events), %6.1f GB/s; CTA-0 phases [us]: %s" % (
            B, D, C, t, nb / t / 1e3, "  ".join("%s %.2f" % (n.split(" ")[0], m) for n, m in zip(NAMES, med))))
        print("      " + extra)
        print("      " + proj)
        print("      " + span)


if __name__ == "__main__":
    main()
