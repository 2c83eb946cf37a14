"""The bar to beat on a B200: the reference's step executed by PyTorch eager (cuDNN/cuBLAS/ATen),
restated with the oracle's functional forward on CUDA tensors + torch.optim.SGD + the
reference's per-step .item() syncs (main.py:419-476), and its eval loop with the per-sample
host round trips (main.py:636-676). Test tooling: prints one JSON line; not part of the product.

    python tests/tools/eager_baseline.py [--steps 10] [--batch 64] [--no-tf32]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import mla_oracle as orc  # noqa: E402
import mla_b200  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--batch", type=int, default=64)
    p.add_argument("--no-tf32", action="store_true")
    a = p.parse_args()
    if a.no_tf32:
        torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=True, lorb="base", modal3=False, clip=False)
    mla_b200.setup_seed(0)
    net = mla_b200.AVClassifier(args).apply(mla_b200.weight_init)
    sd = {k: v.detach().to(dev) for k, v in net.state_dict().items()}
    names = [k for k in sd if sd[k].dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))]
    for k in names:
        sd[k].requires_grad_(True)
    opt = torch.optim.SGD([sd[k] for k in names], lr=1e-3, momentum=0.9, weight_decay=1e-4)
    W, b = sd["fusion_module.fc_out.weight"], sd["fusion_module.fc_out.bias"]
    batches = [tuple(t.to(dev) for t in orc.synthetic_av_batch(a.batch, 1 + i)) for i in range(2)]

    def step(i):
        spec, image, label = batches[i % 2]
        opt.zero_grad()
        fa, fv = orc.av_forward(sd, spec.unsqueeze(1).float(), image.float(), training=True)
        tot = []
        for feat in (fa, fv):
            loss = F.cross_entropy(F.linear(feat, W, b), label)
            loss.backward()
            opt.step()
            opt.zero_grad()
            tot.append(loss)
        return (tot[0] * 0.55 + tot[1] * 0.45).item(), tot[0].item(), tot[1].item()   # main.py:472-475

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    t_train = e0.elapsed_time(e1) * 1e-3 / a.steps

    def evaluate(i):
        spec, image, label = batches[i % 2]
        with torch.no_grad():
            fa, fv = orc.av_forward(sd, spec.unsqueeze(1).float(), image.float(), training=False)
            oa, ov = F.linear(fa, W, b), F.linear(fv, W, b)
            ents = []
            for o in (oa, ov):
                pr = F.softmax(o, dim=0)
                ents.append(-torch.sum(pr * torch.log(pr)))
            mx = max(ents[0], ents[1])
            g = [torch.exp(mx - e) for e in ents]
            w = [x / (g[0] + g[1]) for x in g]
            out = oa * w[0] + ov * w[1]
            preds = [F.softmax(x, dim=1) for x in (out, ov, oa)]
            hits = 0
            for j in range(image.shape[0]):                         # main.py:659-676: per-sample host round trips
                m = [np.argmax(p_[j].cpu().data.numpy()) for p_ in preds]
                hits += int(np.asarray(label[j].cpu()) == m[0])
        return hits

    evaluate(0)
    torch.cuda.synchronize()
    e0.record()
    for i in range(max(2, a.steps // 2)):
        evaluate(i)
    e1.record()
    torch.cuda.synchronize()
    t_eval = e0.elapsed_time(e1) * 1e-3 / max(2, a.steps // 2)
    print(json.dumps({"impl": "torch-eager restatement of the reference step on cuda:0", "batch": a.batch,
                      "cudnn_tf32": not a.no_tf32, "train_ms_per_step": 1e3 * t_train,
                      "train_samples_per_s": a.batch / t_train, "eval_ms_per_batch": 1e3 * t_eval,
                      "eval_samples_per_s": a.batch / t_eval,
                      "encoder_tflops": 32.47e9 * a.batch / t_train / 1e12}))


if __name__ == "__main__":
    main()
