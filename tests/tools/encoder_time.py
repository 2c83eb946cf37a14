"""Time each encoder's forward / backward alone on an otherwise idle GPU (B=64, BASELINE shapes): how long is the
visual chain — the critical path of the overlapped step — by itself?"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mla_b200  # noqa: E402
from mla_b200 import dist as mdist  # noqa: E402
from mla_b200 import encoder_engine as ee  # noqa: E402
from oracle import mla_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True, dynamic=True,
                          lorb="base", modal3=False, clip=False)
mla_b200.setup_seed(0)
net = mla_b200.AVClassifier(args).apply(mla_b200.weight_init).to(dev).train()
spec, image, _ = orc.synthetic_av_batch(64, 5)
spec, image = spec.to(dev).unsqueeze(1), image.to(dev)
dfeat = torch.randn(64, 512, device=dev) / 64
for name, enc, x in (("audio", net.audio_net, spec), ("visual", net.visual_net, image)):
    flat = None
    for it in range(4):
        f = enc.pooled(x)
        plan = f._mla_plan
        if flat is None:
            flat = mdist.FlatGrads(list(enc.parameters()))
        flat.attach()
        plan.backward(dfeat)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for it in range(10):
        e[0].record()
        f = enc.pooled(x)
        e[1].record()
        flat.attach()
        f._mla_plan.backward(dfeat)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print("%s encoder alone: forward %.2f ms, backward %.2f ms (wgrad stream %s, graphs %s)" % (
        name, tf / 10, tb / 10, ee._OVERLAP_WGRAD, ee.USE_GRAPHS))
