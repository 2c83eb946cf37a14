"""Step time of the --lorb m3ae --gs_flag path ('base' encoders, Food-101 shapes) through train_epoch, next to the oracle's
torch restatement on the same GPU (what the reference's own code would cost), plus a per-kernel breakdown.

    python tests/tools/m3ae_time.py [B=32] [steps=4] [profile=0|1] [eager=0|1] [modal3=0|1] [only_default=0|1]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mla_b200  # noqa: E402
from mla_b200 import m3ae  # noqa: E402
from mla_b200.main import SyntheticTextImageLoader  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    profile = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    eager = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    modal3 = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
    args = argparse.Namespace(dataset="IEMOCAP" if modal3 else "Food101", fusion_method="concat", modulation="Normal",
                              gs_flag=True, dynamic=True, lorb="m3ae", modal3=modal3, clip=False)
    mla_b200.setup_seed(0)
    net = mla_b200.Modal3Classifier(args) if modal3 else mla_b200.M3AEClassifier(args)

    def Loader(b, n, seed):
        return SyntheticTextImageLoader(b, n, seed, n_classes=4 if modal3 else 101, audio_len=1024 if modal3 else 0)
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    model = mla_b200.ModuleHolder(net.cuda())
    dev = torch.device("cuda")
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin()

    def run(n):
        loader = Loader(B, n, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses = mla_b200.train_epoch(args, 0, model, dev, loader, opt, sch, gs_plugin=gs, gs_flag=True, av_alpha=0.55)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, losses

    only_default = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
    for fused, bwd in ((True, "fp16"), (True, "tf32"), (False, "tf32"), (False, "bf16"))[:1 if only_default else 4]:
        m3ae.FUSED_BLOCK, m3ae.BLOCK_BACKWARD, m3ae.BACKWARD_BF16 = fused, bwd if fused else "fp16", bwd == "bf16"
        run(2)
        torch.cuda.reset_peak_memory_stats()
        ms, losses = run(steps)
        print("%s base B=%d %s backward=%s: %.2f ms/step, %.1f samples/s, peak mem %.1f GB, losses %s" % (
            "modal3" if modal3 else "m3ae", B, "fused blocks" if fused else "per-module", bwd, ms,
            B * 1000.0 / ms, torch.cuda.max_memory_allocated() / 2**30, tuple(round(x, 4) for x in losses)), flush=True)
    m3ae.FUSED_BLOCK, m3ae.BLOCK_BACKWARD, m3ae.BACKWARD_BF16 = True, "fp16", False
    if profile:
        from torch.profiler import profile as tprof, ProfilerActivity
        loader = Loader(B, 2, 1)
        with tprof(activities=[ProfilerActivity.CUDA]) as prof:
            mla_b200.train_epoch(args, 0, model, dev, loader, opt, sch, gs_plugin=gs, gs_flag=True, av_alpha=0.55)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
    if eager:
        from oracle import mla_oracle as orc
        del model, net, opt
        torch.cuda.empty_cache()
        O = orc.Modal3Oracle if modal3 else orc.M3AEOracle
        o = O({k: v.cuda() for k, v in state.items()}, num_heads=12)
        loader = Loader(B, steps, 1)
        bl = [tuple(t.cuda() for t in b[:-1]) for b in loader.batches]
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            o.train_epoch(bl[:1])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o.train_epoch([bl[i % len(bl)] for i in range(steps)])
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / steps
            print("torch eager restatement (matmul allow_tf32=%s) B=%d: %.2f ms/step, %.1f samples/s" % (
                tf32, B, ms, B * 1000.0 / ms), flush=True)


if __name__ == "__main__":
    main()
