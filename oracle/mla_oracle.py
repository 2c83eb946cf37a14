"""CPU ORACLE for the MLA hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module, and only as the checker / the timed CPU baseline. Nothing under the
product package imports it; the product path has no CPU fallback.

It restates, on the CPU, what the reference computes on the path named by BASELINE.json:
  * GSPlugin.before_update            utils/utils.py:24-41          (numpy, fp32 or fp64)
  * calculate_entropy / gating        main.py:65-106                (numpy)
  * eval fusion + accuracy counters   main.py:636-679               (numpy)
  * shared head + CrossEntropyLoss    fusion_modules.py:19, main.py:130,432-435 (numpy)
  * ResNet-18 encoders / AVClassifier models/backbone.py:142-160, basic_model.py:52-77 (torch CPU fp32)
  * the alternating gs train step     main.py:419-476               (torch CPU fp32 + autograd + SGD)
  * valid() gs branch                 main.py:622-679
  * joint training step + OGM / OGM-GE main.py:165-168,269-418; valid() without gs_flag main.py:538-620  (torch CPU fp32)
  * m3ae encoders / M3AEClassifier    models/m3ae.py:65-224,337-370, basic_model.py:184-200 (torch CPU fp32)
  * CAV-MAE audio / Modal3Classifier  models/cav_mae.py:69-151,337-351, basic_model.py:252-275 (torch CPU fp32; the
                                      block's Attention / Mlp come from un-vendored timm==0.4.5: those two UNPINNED)
  * visual dataset transform           dataset/dataset.py:123-161 (numpy integer arithmetic). Its arithmetic lives in
                                      third-party libraries the reference calls — torchvision.transforms (Resize /
                                      RandomResizedCrop / ToTensor / Normalize) over Pillow's Image.resize(BILINEAR)
                                      (libImaging/Resample.c) — which ARE installed here (torchvision 0.26, Pillow 12.2):
                                      pinned by running them live in tests/test_oracle_frames.py plus a committed fixture.

Pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container under torch
2.11: tests/golden/make_golden.py imports /root/reference, runs its functions on
seeded inputs and commits the results as fixtures under tests/golden/; tests/test_oracle_golden.py
checks every function here against them. The floating-point encoder/step restatement uses
torch CPU fp32 ops (the task allows a torch fp32 reference for floating-point kernels).
"""
import math

import numpy as np

# --------------------------------------------------------------------------------------------
# GSPlugin.before_update — utils/utils.py:24-41
# --------------------------------------------------------------------------------------------


def gs_alpha(batch_index, len_dataloader):
    """utils.py:26-27 — Python double arithmetic."""
    lamda = batch_index / len_dataloader + 1
    return 1.0 * 0.1 ** lamda


def gs_before_update(P, feat, grad_w, batch_index, len_dataloader, counter, dtype=np.float32, mode=0):
    """utils.py:24-41 with the parameter already selected (the 'module.weight' name gate is
    host logic). Returns (P_out, grad_out) as new arrays of `dtype`.
    mode 0 = the reference's elementwise denominator; mode 1 = scalar OWM."""
    P = np.asarray(P, dtype=dtype)
    g = None if grad_w is None else np.asarray(grad_w, dtype=dtype)
    if counter == 0:                                             # utils.py:29
        return P.copy(), (None if g is None else g.copy())
    alpha = dtype(gs_alpha(batch_index, len_dataloader))
    feat = np.asarray(feat, dtype=dtype)
    r = feat.mean(axis=0, keepdims=True, dtype=dtype)            # utils.py:34  [1, D]
    k = P @ r.T                                                  # utils.py:35  [D, 1]
    if mode == 0:
        den = alpha + k @ r                                      # [D, D] outer product: elementwise
    else:
        den = alpha + (r @ k)[0, 0]
    P1 = P - (k @ k.T) / den                                     # utils.py:36
    P1 = (P1 / np.sqrt((P1.astype(dtype) ** 2).sum(dtype=dtype))).astype(dtype)   # utils.py:38-40
    g1 = None if g is None else (g @ P1.T).astype(dtype)         # utils.py:41
    return P1, g1


def gs_before_update_sum(P, feat_sum, inv_batch, grad_w, alpha, dtype=np.float32, mode=0):
    """Same update driven by an already reduced sum_b feat (data-parallel form)."""
    P = np.asarray(P, dtype=dtype)
    r = (np.asarray(feat_sum, dtype=dtype) * dtype(inv_batch)).reshape(1, -1)
    k = P @ r.T
    den = dtype(alpha) + (k @ r if mode == 0 else (r @ k)[0, 0])
    P1 = P - (k @ k.T) / den
    P1 = (P1 / np.sqrt((P1 ** 2).sum(dtype=dtype))).astype(dtype)
    g1 = None if grad_w is None else (np.asarray(grad_w, dtype=dtype) @ P1.T).astype(dtype)
    return P1, g1


# --------------------------------------------------------------------------------------------
# Test-time fusion — main.py:65-106, 636-679
# --------------------------------------------------------------------------------------------


def calculate_entropy(output, dtype=np.float32):
    """main.py:65-70 — softmax over dim 0 (the batch axis), summed over everything."""
    x = np.asarray(output, dtype=dtype)
    e = np.exp(x - x.max(axis=0, keepdims=True))
    p = (e / e.sum(axis=0, keepdims=True, dtype=dtype)).astype(dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        return dtype(-(p * np.log(p)).sum(dtype=dtype))


def calculate_gating_weights(*outputs, dtype=np.float32):
    """main.py:72-87 (2 modalities) / 89-106 (3): w = softmax(-H) with Python max()."""
    ents = [calculate_entropy(o, dtype) for o in outputs]
    mx = ents[0]
    for e in ents[1:]:           # Python max(): first argument wins unless a later one is greater
        if e > mx:
            mx = e
    with np.errstate(invalid="ignore"):
        g = [np.exp(dtype(mx - e)) for e in ents]
        s = dtype(0)
        for x in g:
            s = dtype(s + x)
        return tuple(dtype(x / s) for x in g)


def fuse_eval(outputs, label, n_classes, dynamic=True, fixed_w=None, dtype=np.float32):
    """main.py:636-676. Returns dict(fused, w, argmax[(M+1),B], num[C], hits[(M+1),C])."""
    outs = [np.asarray(o, dtype=dtype) for o in outputs]
    M = len(outs)
    if dynamic:
        w = calculate_gating_weights(*outs, dtype=dtype)                       # main.py:640-646
    else:
        w = tuple(dtype(x) for x in fixed_w)                                   # main.py:648-651
    with np.errstate(invalid="ignore"):
        fused = outs[0] * w[0]
        for m in range(1, M):
            fused = (fused + outs[m] * w[m]).astype(dtype)

    def row_argmax(x):
        # np.argmax(softmax(x)) (main.py:653-664): NaN anywhere -> NaN row -> index 0
        out = np.argmax(x, axis=1).astype(np.int32)
        out[np.isnan(x).any(axis=1)] = 0
        return out

    argmax = np.stack([row_argmax(fused)] + [row_argmax(o) for o in outs])
    num = np.zeros(n_classes, np.int64)
    hits = np.zeros((M + 1, n_classes), np.int64)
    if label is not None:
        lab = np.asarray(label)
        for i in range(lab.shape[0]):                                          # main.py:659-676
            num[lab[i]] += 1
            for j in range(M + 1):
                if argmax[j, i] == lab[i]:
                    hits[j, lab[i]] += 1
    return dict(fused=fused, w=np.array(w, dtype=dtype), argmax=argmax, num=num, hits=hits)


# --------------------------------------------------------------------------------------------
# Shared head + CrossEntropyLoss — fusion_modules.py:19, main.py:130, 432-435
# --------------------------------------------------------------------------------------------


def head_ce(feat, W, b, label, grad_scale=None, dtype=np.float64):
    feat = np.asarray(feat, dtype=dtype)
    W = np.asarray(W, dtype=dtype)
    b = np.asarray(b, dtype=dtype)
    lab = np.asarray(label)
    B = feat.shape[0]
    logits = feat @ W.T + b
    m = logits.max(axis=1, keepdims=True)
    lse = np.log(np.exp(logits - m).sum(axis=1, keepdims=True)) + m
    logp = logits - lse
    loss = -logp[np.arange(B), lab].mean()
    gs = (1.0 / B) if grad_scale is None else grad_scale
    dl = np.exp(logp)
    dl[np.arange(B), lab] -= 1.0
    dl *= gs
    return dict(logits=logits, loss=loss, dW=dl.T @ feat, db=dl.sum(0), dfeat=dl @ W, feat_sum=feat.sum(0))


# --------------------------------------------------------------------------------------------
# OGM / OGM-GE coefficients — main.py:315-334 (three modalities), 373-384 (two)
# --------------------------------------------------------------------------------------------
def ogm_coefficients(outs, label, alpha):
    """scores = sum_b softmax(out_m)[b][label_b] added in index order in fp32 (the reference's Python sum of 0-dim
    tensors); coefficients 1 - tanh(alpha * relu(ratio)) of the dominant modality. Returns (scores, coeffs) as fp32
    arrays; modality order a, v(, t)."""
    lab = np.asarray(label)
    scores = []
    for o in outs:
        o = np.asarray(o, np.float32)
        e = np.exp(o - o.max(axis=1, keepdims=True), dtype=np.float32)
        p = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
        acc = np.float32(0)
        for i in range(o.shape[0]):
            acc = np.float32(acc + p[i, lab[i]])
        scores.append(acc)
    alpha = np.float32(alpha)
    f = lambda r: np.float32(1) - np.tanh(alpha * np.maximum(r, np.float32(0)), dtype=np.float32)   # noqa: E731
    coeff = [np.float32(1)] * len(outs)
    if len(outs) == 2:
        ratio_v = np.float32(scores[1] / scores[0])
        ratio_a = np.float32(np.float32(1) / ratio_v)
        if ratio_v > 1:
            coeff[1] = f(ratio_v)
        else:
            coeff[0] = f(ratio_a)
    else:
        ratio_v = np.float32(scores[1] / np.float32(scores[0] + scores[2]))
        ratio_a = np.float32(scores[0] / np.float32(scores[1] + scores[2]))
        ratio_t = np.float32(scores[2] / np.float32(scores[1] + scores[0]))
        if ratio_v > 1:
            coeff[1] = f(ratio_v)
        elif ratio_t > 1:
            coeff[2] = f(ratio_t)
        else:
            coeff[0] = f(ratio_a)
    return np.array(scores, np.float32), np.array(coeff, np.float32)


# --------------------------------------------------------------------------------------------
# ResNet-18 encoders, AVClassifier, the alternating step — torch CPU fp32 restatement
# --------------------------------------------------------------------------------------------

_LAYERS = ((64, 1), (128, 2), (256, 2), (512, 2))   # resnet18: BasicBlock x [2,2,2,2], backbone.py:211-213


def _bn(x, sd, prefix, training, momentum=0.1, eps=1e-5):
    import torch.nn.functional as F
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training, momentum, eps)


def resnet18_forward(sd, prefix, x, modality, training):
    """models/backbone.py:142-160 (ResNet.forward) + :36-52 (BasicBlock.forward), functional.
    `sd` maps the reference's state-dict names to tensors (parameters may require grad;
    BN running statistics are updated in place when training)."""
    import torch.nn.functional as F
    if modality == "visual":
        B, C, T, H, W = x.shape
        x = x.permute(0, 2, 1, 3, 4).contiguous().view(B * T, C, H, W)        # backbone.py:144-147
    x = F.conv2d(x, sd[prefix + "conv1.weight"], None, stride=2, padding=3)    # backbone.py:149
    x = F.relu(_bn(x, sd, prefix + "bn1", training))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)                    # backbone.py:152
    inplanes = 64
    for li, (planes, stride) in enumerate(_LAYERS, start=1):
        for bi in range(2):
            s = stride if bi == 0 else 1
            p = "%slayer%d.%d." % (prefix, li, bi)
            identity = x
            out = F.conv2d(x, sd[p + "conv1.weight"], None, stride=s, padding=1)
            out = F.relu(_bn(out, sd, p + "bn1", training))
            out = F.conv2d(out, sd[p + "conv2.weight"], None, stride=1, padding=1)
            out = _bn(out, sd, p + "bn2", training)
            if bi == 0 and (s != 1 or inplanes != planes):
                identity = F.conv2d(x, sd[p + "downsample.0.weight"], None, stride=s)
                identity = _bn(identity, sd, p + "downsample.1", training)
            x = F.relu(out + identity)
            inplanes = planes
    return x


def av_forward(sd, audio, visual, training):
    """models/basic_model.py:52-77, gs_flag branch: returns (a, v) in [B, 512]."""
    import torch
    import torch.nn.functional as F
    a = resnet18_forward(sd, "audio_net.", audio, "audio", training)
    v = resnet18_forward(sd, "visual_net.", visual, "visual", training)
    _, C, H, W = v.shape
    B = a.shape[0]
    v = v.view(B, -1, C, H, W).permute(0, 2, 1, 3, 4)
    a = torch.flatten(F.adaptive_avg_pool2d(a, 1), 1)
    v = torch.flatten(F.adaptive_avg_pool3d(v, 1), 1)
    return a, v


def _bump_num_batches(sd, prefix):
    for k in sd:
        if k.startswith(prefix) and k.endswith("num_batches_tracked"):
            sd[k] += 1


class SGD:
    """torch.optim.SGD(lr, momentum=0.9, weight_decay=1e-4) restated (main.py:749): parameters
    whose grad is None are skipped, as under torch >= 2.0 zero_grad(set_to_none=True)."""

    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=1e-4):
        self.params, self.lr, self.momentum, self.wd = list(params), lr, momentum, weight_decay
        self.buf = {}

    def step(self):
        import torch
        with torch.no_grad():
            for i, p in enumerate(self.params):
                if p.grad is None:
                    continue
                g = p.grad.add(p, alpha=self.wd)
                if i not in self.buf:
                    self.buf[i] = g.clone()
                else:
                    self.buf[i].mul_(self.momentum).add_(g)
                p.add_(self.buf[i], alpha=-self.lr)

    def zero_grad(self):
        for p in self.params:
            p.grad = None


class AVOracle:
    """State + step functions of the CREMA-D AVClassifier path (configs 1/2 of BASELINE.json).

    `state` is a dict with the reference's state-dict names WITHOUT the 'module.' prefix
    (audio_net.*, visual_net.*, fusion_module.fc_out.{weight,bias}); tensors are torch CPU.
    (Tests may also hand it CUDA tensors: the same torch code then runs the reference's
    arithmetic as the reference itself would on a GPU — cuDNN convolutions, TF32 under torch's
    defaults — which calibrates how far ANY TF32 implementation sits from the fp32 fixtures.)"""

    def __init__(self, state, lr=1e-3, force_projection=False, gs_mode=0):
        import torch
        self.sd = {k: v.detach().clone() for k, v in state.items()}
        self.param_names = [k for k in self.sd if self.sd[k].dtype.is_floating_point
                            and not k.endswith(("running_mean", "running_var"))]
        for k in self.param_names:
            self.sd[k].requires_grad_(True)
        self.opt = SGD([self.sd[k] for k in self.param_names], lr=lr)
        D = self.sd["fusion_module.fc_out.weight"].shape[1]
        self.Pl = np.eye(D, dtype=np.float32)         # utils.py:19-20 (sized from the head)
        self.exp_count = 0
        self.force_projection = force_projection      # False = as published: the hook never fires
        self.gs_mode = gs_mode
        self.torch = torch

    def _turn(self, feat, label, batch_step, len_dl, encoder_prefix):
        torch = self.torch
        import torch.nn.functional as F
        W, b = self.sd["fusion_module.fc_out.weight"], self.sd["fusion_module.fc_out.bias"]
        out = F.linear(feat, W, b)                                     # main.py:432
        loss = F.cross_entropy(out, label)                             # main.py:434
        loss.backward(retain_graph=True)                               # main.py:435
        if self.force_projection:                                      # main.py:437-438 -> utils.py:24-41
            P1, g1 = gs_before_update(self.Pl, feat.detach().cpu().numpy(), W.grad.cpu().numpy(), batch_step, len_dl,
                                      self.exp_count, mode=self.gs_mode)
            self.Pl = P1
            W.grad = torch.from_numpy(np.ascontiguousarray(g1)).to(W.device)
        self.opt.step()                                                # main.py:439
        self.opt.zero_grad()                                           # main.py:440
        self.exp_count += 1                                            # main.py:442
        return float(loss.detach())

    def train_step(self, spec, image, label, batch_step=0, len_dl=1):
        """One iteration of the loop body main.py:143-476 (gs branch). Returns (loss_a, loss_v)."""
        self.opt.zero_grad()                                           # main.py:164
        a, v = av_forward(self.sd, spec.unsqueeze(1).float(), image.float(), training=True)   # main.py:431
        _bump_num_batches(self.sd, "audio_net.")
        _bump_num_batches(self.sd, "visual_net.")
        la = self._turn(a, label, batch_step, len_dl, "audio_net.")
        lv = self._turn(v, label, batch_step, len_dl, "visual_net.")
        return la, lv

    def train_step_dp(self, shards, batch_step=0, len_dl=1):
        """The same iteration under nn.DataParallel (main.py:732) with one shard per replica: every replica runs
        the encoders on ITS shard with its own BatchNorm batch statistics (only replica 0's running statistics
        survive), the features are gathered and the head / loss / GS hook / optimiser see the GLOBAL batch.
        `shards` = [(spec, image, label), ...]. Returns (loss_a, loss_v). This is what the one-process-per-GPU
        runtime of the product must reproduce."""
        torch = self.torch
        self.opt.zero_grad()
        feats_a, feats_v, labels = [], [], []
        stat_keys = [k for k in self.sd if k.endswith(("running_mean", "running_var"))]
        for r, (spec, image, label) in enumerate(shards):
            sd = self.sd
            if r > 0:                      # replicas > 0 update throw-away copies of the running statistics
                sd = dict(self.sd)
                for k in stat_keys:
                    sd[k] = self.sd[k].clone()
            a, v = av_forward(sd, spec.unsqueeze(1).float(), image.float(), training=True)
            feats_a.append(a); feats_v.append(v); labels.append(label)
        _bump_num_batches(self.sd, "audio_net.")
        _bump_num_batches(self.sd, "visual_net.")
        a, v, label = torch.cat(feats_a), torch.cat(feats_v), torch.cat(labels)
        la = self._turn(a, label, batch_step, len_dl, "audio_net.")
        lv = self._turn(v, label, batch_step, len_dl, "visual_net.")
        return la, lv

    def train_epoch(self, batches, av_alpha=0.5):
        """main.py:127-484 (gs branch) over a list of (spec, image, label). Returns (loss, loss_a, loss_v)."""
        tot = tot_a = tot_v = 0.0
        for step, (spec, image, label) in enumerate(batches):
            la, lv = self.train_step(spec, image, label, step, len(batches))
            # main.py:472: (loss_a * av_alpha + loss_v * (1 - av_alpha)).item() in fp32
            tot += float(np.float32(np.float32(la) * np.float32(av_alpha)) +
                         np.float32(np.float32(lv) * np.float32(1 - av_alpha)))
            tot_a += la
            tot_v += lv
        n = len(batches)
        return tot / n, tot_a / n, tot_v / n

    # ---- joint training without --gs_flag: main.py:165-168, 269-310, 312-410, 412-418 ----
    def joint_step(self, spec, image, label, modulation="Normal", alpha=0.3, in_window=True):
        """One iteration of the non-gs loop body for AVClassifier with the concatenated head. Returns
        (loss, loss_a, loss_v); self.last_ogm = (scores, coeffs) when a modulation is active."""
        torch = self.torch
        import torch.nn.functional as F
        self.opt.zero_grad()                                           # main.py:164
        a, v = av_forward(self.sd, spec.unsqueeze(1).float(), image.float(), training=True)
        _bump_num_batches(self.sd, "audio_net.")
        _bump_num_batches(self.sd, "visual_net.")
        W, b = self.sd["fusion_module.fc_out.weight"], self.sd["fusion_module.fc_out.bias"]
        out = F.linear(torch.cat((a, v), dim=1), W, b)                 # fusion_modules.py:21-24
        D = W.shape[1] // 2
        out_v = torch.mm(v, W[:, D:].t()) + b / 2                      # main.py:304-305
        out_a = torch.mm(a, W[:, :D].t()) + b / 2                      # main.py:307-308
        loss = F.cross_entropy(out, label)                             # main.py:309
        loss_a, loss_v = F.cross_entropy(out_a, label), F.cross_entropy(out_v, label)
        loss.backward()                                                # main.py:313
        if modulation in ("OGM", "OGM_GE"):
            scores, coeff = ogm_coefficients([out_a.detach().cpu().numpy(), out_v.detach().cpu().numpy()],
                                             label.cpu().numpy(), alpha)
            self.last_ogm = (scores, coeff)
            if in_window:                                              # main.py:393
                for k in self.param_names:                             # named_parameters() order
                    p = self.sd[k]
                    if p.grad is None or p.grad.dim() != 4:
                        continue
                    for key, c in (("audio", coeff[0]), ("visual", coeff[1])):
                        if key in k.split(".")[0]:
                            if modulation == "OGM_GE":                 # main.py:398-399
                                p.grad = p.grad * float(c) + torch.zeros_like(p.grad).normal_(0, p.grad.std().item() + 1e-8)
                            else:
                                p.grad *= float(c)                     # main.py:401
        self.opt.step()                                                # main.py:412
        return float(loss.detach()), float(loss_a.detach()), float(loss_v.detach())

    def joint_epoch(self, batches, modulation="Normal", alpha=0.3, epoch=0, starts=0, ends=50):
        tot = np.zeros(3)
        for spec, image, label in batches:
            tot += np.array(self.joint_step(spec, image, label, modulation, alpha, starts <= epoch <= ends))
        return tuple(tot / len(batches))

    def joint_valid(self, batches, n_classes=6):
        """main.py:538-620, 653-679 for the concatenated head. Returns (acc, acc_a, acc_v)."""
        torch = self.torch
        import torch.nn.functional as F
        num = np.zeros(n_classes, np.int64)
        hits = np.zeros((3, n_classes), np.int64)
        with torch.no_grad():
            for spec, image, label in batches:
                a, v = av_forward(self.sd, spec.unsqueeze(1).float(), image.float(), training=False)
                W, b = self.sd["fusion_module.fc_out.weight"], self.sd["fusion_module.fc_out.bias"]
                D = W.shape[1] // 2
                outs = [F.linear(torch.cat((a, v), dim=1), W, b), torch.mm(a, W[:, :D].t()) + b / 2,
                        torch.mm(v, W[:, D:].t()) + b / 2]
                lab = label.cpu().numpy()
                for i, o in enumerate(outs):
                    pred = np.argmax(F.softmax(o, dim=1).cpu().numpy(), axis=1)     # main.py:653-663
                    np.add.at(hits[i], lab[pred == lab], 1)
                np.add.at(num, lab, 1)
        tot = float(num.sum())
        return hits[0].sum() / tot, hits[1].sum() / tot, hits[2].sum() / tot

    def eval_logits(self, spec, image):
        torch = self.torch
        import torch.nn.functional as F
        with torch.no_grad():
            a, v = av_forward(self.sd, spec.unsqueeze(1).float(), image.float(), training=False)
            W, b = self.sd["fusion_module.fc_out.weight"], self.sd["fusion_module.fc_out.bias"]
            return F.linear(a, W, b), F.linear(v, W, b)

    def valid(self, batches, n_classes=6, dynamic=True, av_alpha=0.5):
        """main.py:486-679 (gs branch, 2 modalities). Returns (acc, acc_a, acc_v)."""
        num = np.zeros(n_classes, np.int64)
        hits = np.zeros((3, n_classes), np.int64)
        for spec, image, label in batches:
            oa, ov = self.eval_logits(spec, image)
            r = fuse_eval([oa.cpu().numpy(), ov.cpu().numpy()], label.cpu().numpy(), n_classes, dynamic=dynamic,
                          fixed_w=(av_alpha, 1 - av_alpha))
            num += r["num"]
            hits += r["hits"]
        tot = float(num.sum())
        return hits[0].sum() / tot, hits[1].sum() / tot, hits[2].sum() / tot


# --------------------------------------------------------------------------------------------
# m3ae encoders / M3AEClassifier — models/m3ae.py:86-179,181-224,337-370, models/basic_model.py:184-200
# --------------------------------------------------------------------------------------------
def sincos_1d(embed_dim, pos):
    """m3ae.py:181-194: [sin(pos * w_i) | cos(pos * w_i)], w_i = 10000^(-i / (D/2)), float32 throughout."""
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.
    omega = 1. / 10000 ** omega
    out = np.einsum("m,d->md", np.asarray(pos, np.float32).reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim, length):
    """m3ae.py:207-224: first half of the channels from meshgrid(w, h)[0] (the column index), second half from the row."""
    g = int(length ** 0.5)
    assert g * g == length
    grid = np.stack(np.meshgrid(np.arange(g, dtype=np.float32), np.arange(g, dtype=np.float32)), axis=0)
    return np.concatenate([sincos_1d(embed_dim // 2, grid[0]), sincos_1d(embed_dim // 2, grid[1])], axis=1)


def preln_block(x, mask, num_heads, g1, b1, wq, bq, wo, bo, g2, b2, w1, c1, w2, c2):
    """One pre-LN transformer block: m3ae.py:128-154 (Block) with :86-125 (Attention: scores * scale, padded keys FILLED
    with -1e7, softmax) and :65-83 (MLP, exact GELU); also cav_mae.py:86-113 with timm 0.4.5's Attention / Mlp (mask None)."""
    import torch
    import torch.nn.functional as F
    B, S, D = x.shape
    scale = (D // num_heads) ** -0.5
    h = F.layer_norm(x, (D,), g1, b1)
    qkv = F.linear(h, wq, bq).view(B, S, 3, num_heads, D // num_heads).permute(2, 0, 3, 1, 4)
    att = torch.matmul(qkv[0], qkv[1].transpose(-2, -1)) * scale
    if mask is not None:
        att = torch.where(mask[:, None, None, :].expand(att.shape) > 0, torch.tensor(-1e7, device=x.device), att)
    att = F.softmax(att, dim=-1)
    h = torch.matmul(att, qkv[2]).permute(0, 2, 1, 3).reshape(B, S, D)
    x = x + F.linear(h, wo, bo)
    h = F.layer_norm(x, (D,), g2, b2)
    return x + F.linear(F.gelu(F.linear(h, w1, c1)), w2, c2)


def m3ae_representation(sd, prefix, image, text, text_padding_mask, num_heads):
    """MaskedMultimodalAutoencoder.forward_representation (m3ae.py:337-370) + Transformer/Block/Attention/TransformerMLP
    (m3ae.py:65-179), functional over a state dict. DropPath is the identity (as published it returns None and
    the forward raises; the configured rate is 0), dropout rates are 0."""
    import torch
    import torch.nn.functional as F
    cls = sd[prefix + "cls_token"]
    D = cls.shape[-1]
    B = image.shape[0] if image is not None else text.shape[0]
    dev = cls.device
    xs = [cls.expand(B, 1, D)]
    masks = [torch.zeros(B, 1, dtype=torch.float32, device=dev)]
    if image is not None:
        pe = torch.from_numpy(sincos_2d(D, image.shape[1])[None]).to(dev)
        xs.append(F.linear(image, sd[prefix + "image_embedding.weight"], sd[prefix + "image_embedding.bias"]) + pe
                  + sd[prefix + "encoder_image_type_embedding"])
        masks.append(torch.zeros(B, image.shape[1], dtype=torch.float32, device=dev))
    if text is not None:
        pe = torch.from_numpy(sincos_1d(D, np.arange(text.shape[1], dtype=np.float32))[None]).to(dev)
        xs.append(F.embedding(text, sd[prefix + "text_embedding.weight"]) + pe + sd[prefix + "encoder_text_type_embedding"])
        masks.append(text_padding_mask)
    x = torch.cat(xs, dim=1)
    mask = torch.cat(masks, dim=1)
    depth = 1 + max(int(k[len(prefix) + 15:].split(".")[0]) for k in sd if k.startswith(prefix + "encoder.blocks."))
    for i in range(depth):
        p = "%sencoder.blocks.%d." % (prefix, i)
        x = preln_block(x, mask, num_heads, *[sd[p + k] for k in (
            "layer_norm1.weight", "layer_norm1.bias", "attention.qkv_linear.weight", "attention.qkv_linear.bias",
            "attention.fc.weight", "attention.fc.bias", "layer_norm2.weight", "layer_norm2.bias",
            "transformer_mlp.fc1.weight", "transformer_mlp.fc1.bias", "transformer_mlp.fc2.weight", "transformer_mlp.fc2.bias")])
    return F.layer_norm(x, (D,), sd[prefix + "encoder.layer_norm.weight"], sd[prefix + "encoder.layer_norm.bias"])


def m3ae_forward(sd, token, padding_mask, visual, num_heads):
    """M3AEClassifier.forward (basic_model.py:184-200): 16x16 patches 'b c (h p1) (w p2) -> b (h w) (c p1 p2)', text through
    mae_a, patches through mae_v, mean over ALL tokens (CLS and padded ones included)."""
    B, C, H, W = visual.shape
    patches = visual.reshape(B, C, H // 16, 16, W // 16, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 16) * (W // 16), C * 256)
    a = m3ae_representation(sd, "mae_a.", None, token.squeeze(1), padding_mask.squeeze(1), num_heads)
    v = m3ae_representation(sd, "mae_v.", patches, None, None, num_heads)
    return a.mean(dim=1), v.mean(dim=1)


class M3AEOracle(AVOracle):
    """The same alternating step (main.py:419-476) with the m3ae packet layout (token, padding_mask, image, label)."""

    def __init__(self, state, num_heads, **kw):
        super().__init__(state, **kw)
        self.num_heads = num_heads

    def train_step(self, token, padding_mask, image, label, batch_step=0, len_dl=1):
        self.opt.zero_grad()
        a, v = m3ae_forward(self.sd, token, padding_mask, image, self.num_heads)          # main.py:428
        la = self._turn(a, label, batch_step, len_dl, "mae_a.")
        lv = self._turn(v, label, batch_step, len_dl, "mae_v.")
        return la, lv

    def train_epoch(self, batches, av_alpha=0.5):
        tot = tot_a = tot_v = 0.0
        for step, b in enumerate(batches):
            la, lv = self.train_step(b[0], b[1], b[2], b[3], step, len(batches))
            tot += float(np.float32(np.float32(la) * np.float32(av_alpha)) + np.float32(np.float32(lv) * np.float32(1 - av_alpha)))
            tot_a += la
            tot_v += lv
        n = len(batches)
        return tot / n, tot_a / n, tot_v / n

    def eval_logits(self, token, padding_mask, image):
        torch = self.torch
        import torch.nn.functional as F
        with torch.no_grad():
            a, v = m3ae_forward(self.sd, token, padding_mask, image, self.num_heads)
            W, b = self.sd["fusion_module.fc_out.weight"], self.sd["fusion_module.fc_out.bias"]
            return F.linear(a, W, b), F.linear(v, W, b)

    def valid(self, batches, n_classes=101, dynamic=True, av_alpha=0.5):
        num = np.zeros(n_classes, np.int64)
        hits = np.zeros((3, n_classes), np.int64)
        for b in batches:
            oa, ov = self.eval_logits(b[0], b[1], b[2])
            r = fuse_eval([oa.cpu().numpy(), ov.cpu().numpy()], b[3].cpu().numpy(), n_classes, dynamic=dynamic,
                          fixed_w=(av_alpha, 1 - av_alpha))
            num += r["num"]
            hits += r["hits"]
        tot = float(num.sum())
        return hits[0].sum() / tot, hits[1].sum() / tot, hits[2].sum() / tot


# --------------------------------------------------------------------------------------------
# CAV-MAE audio encoder / Modal3Classifier — models/cav_mae.py:69-113,116-151,337-351, basic_model.py:252-275.
# The block's Attention / Mlp classes come from timm==0.4.5 (requirements.txt:55), which is NOT in /root/reference and not
# installed: PARITY OF THOSE TWO CLASSES IS UNPINNED. Their published algorithm (qkv Linear with bias -> heads ->
# softmax(q k^T * head_dim^-0.5) v -> proj Linear; fc1 -> exact GELU -> fc2) is what preln_block computes with mask=None;
# everything around them (patch embedding, embeddings, block wiring, norms, the classifier, the step) is pinned to the
# reference's own code executed with those two classes restated (tests/golden/make_golden.py).
# --------------------------------------------------------------------------------------------
def cav_audio_features(sd, prefix, audio, num_heads):
    """CAVMAEFT.forward_feat(a, None, 'a') (cav_mae.py:337-351): [B, T, 128] -> [B, T*128/256, D]."""
    import torch.nn.functional as F
    a = audio.unsqueeze(1).transpose(2, 3)
    a = F.conv2d(a, sd[prefix + "patch_embed_a.proj.weight"], sd[prefix + "patch_embed_a.proj.bias"], stride=16)
    a = a.flatten(2).transpose(1, 2) + sd[prefix + "pos_embed_a"] + sd[prefix + "modality_a"]
    D = a.shape[-1]

    def run(group, n1, n2, x):
        depth = 1 + max([int(k[len(prefix) + len(group) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + group + ".")],
                        default=-1)
        for i in range(depth):
            p = "%s%s.%d." % (prefix, group, i)
            x = preln_block(x, None, num_heads, *[sd[p + k] for k in (
                n1 + ".weight", n1 + ".bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                n2 + ".weight", n2 + ".bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")])
        return x
    a = run("blocks_a", "norm1", "norm2", a)
    a = run("blocks_u", "norm1_a", "norm2_a", a)
    return F.layer_norm(a, (D,), sd[prefix + "norm_a.weight"], sd[prefix + "norm_a.bias"])


def modal3_forward(sd, token, padding_mask, visual, audio, num_heads):
    """Modal3Classifier.forward (basic_model.py:252-275) -> (a, v, t)."""
    B, C, H, W = visual.shape
    patches = visual.reshape(B, C, H // 16, 16, W // 16, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 16) * (W // 16), C * 256)
    a = cav_audio_features(sd, "mae_a.", audio, num_heads)
    t = m3ae_representation(sd, "mae_t.", None, token.squeeze(1), padding_mask.squeeze(1), num_heads)
    v = m3ae_representation(sd, "mae_v.", patches, None, None, num_heads)
    return a.mean(dim=1), v.mean(dim=1), t.mean(dim=1)


class Modal3Oracle(AVOracle):
    """The alternating step over THREE encoders (main.py:419-466, modal3): a -> v -> t turns on the shared head."""

    def __init__(self, state, num_heads, **kw):
        super().__init__(state, **kw)
        self.num_heads = num_heads

    def train_step(self, token, padding_mask, image, spec, label, batch_step=0, len_dl=1):
        self.opt.zero_grad()
        a, v, t = modal3_forward(self.sd, token, padding_mask, image, spec, self.num_heads)   # main.py:426
        return (self._turn(a, label, batch_step, len_dl, "mae_a."), self._turn(v, label, batch_step, len_dl, "mae_v."),
                self._turn(t, label, batch_step, len_dl, "mae_t."))

    def train_epoch(self, batches, av_alpha=0.5):
        tot = np.zeros(4)
        for step, b in enumerate(batches):
            la, lv, lt = self.train_step(b[0], b[1], b[2], b[3], b[4], step, len(batches))
            mix = float(np.float32(np.float32(la) * np.float32(av_alpha)) + np.float32(np.float32(lv) * np.float32(1 - av_alpha)))
            tot += (mix, la, lv, lt)                       # main.py:472-476: the mixed loss ignores the third modality
        return tuple(tot / len(batches))

    def valid(self, batches, n_classes=4, dynamic=True, alphas=(0.35, 0.25, 0.4)):
        torch = self.torch
        import torch.nn.functional as F
        num = np.zeros(n_classes, np.int64)
        hits = np.zeros((4, n_classes), np.int64)
        W, b = self.sd["fusion_module.fc_out.weight"], self.sd["fusion_module.fc_out.bias"]
        for bt in batches:
            with torch.no_grad():
                feats = modal3_forward(self.sd, bt[0], bt[1], bt[2], bt[3], self.num_heads)
                outs = [F.linear(f, W, b).cpu().numpy() for f in feats]
            r = fuse_eval(outs, bt[4].cpu().numpy(), n_classes, dynamic=dynamic, fixed_w=alphas)
            num += r["num"]
            hits += r["hits"]
        tot = float(num.sum())
        return tuple(hits[i].sum() / tot for i in range(4))


def synthetic_m3ae_batch(batch, seed, text_len=256, image_hw=(256, 256), n_classes=101, vocab=30522):
    """(token [B,1,L] int64, padding_mask [B,1,L] int64 with a random padded tail, image [B,3,H,W], label)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    token = torch.randint(0, vocab, (batch, 1, text_len), generator=g)
    n_valid = torch.randint(max(1, text_len // 4), text_len + 1, (batch,), generator=g)
    padding_mask = (torch.arange(text_len)[None, :] >= n_valid[:, None]).long()[:, None, :]
    image = torch.randn(batch, 3, *image_hw, generator=g)
    label = torch.randint(0, n_classes, (batch,), generator=g)
    return token, padding_mask, image, label


def synthetic_av_batch(batch, seed, spec_hw=(257, 188), frames=2, image_hw=(224, 224), n_classes=6):
    """Seeded CREMA-D-shaped batch: spec ~ N(0,1) [B,257,188], image ~ N(0,1)
    [B,3,T,224,224], label ~ U{0..5}. torch CPU generator => identical on every box."""
    import torch
    g = torch.Generator().manual_seed(seed)
    spec = torch.randn(batch, *spec_hw, generator=g)
    image = torch.randn(batch, 3, frames, *image_hw, generator=g)
    label = torch.randint(0, n_classes, (batch,), generator=g)
    return spec, image, label


def math_isclose(a, b, rel):
    return math.isclose(a, b, rel_tol=rel, abs_tol=rel)


# ------------------------------------------------------------------------------------------------------------------
# Visual dataset transform (dataset/dataset.py:123-161): torchvision Resize / resized_crop on PIL images -> ToTensor ->
# Normalize. Pillow's bilinear resampling (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
# ImagingResampleHorizontal_8bpc / Vertical_8bpc) restated in numpy integer arithmetic.
PIL_PRECISION_BITS = 32 - 8 - 2


def _pil_filter(x, bicubic):
    """Resample.c: bilinear_filter (support 1) / bicubic_filter (support 2, a = -0.5)."""
    x = -x if x < 0.0 else x
    if not bicubic:
        return 1.0 - x if x < 1.0 else 0.0
    a = -0.5
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bilinear_coeffs(in_size, out_size, bicubic=False):
    """(bounds [out, 2] = (xmin, n), fixed-point coefficients [out, ksize] int32) of Pillow's BILINEAR (or BICUBIC) filter
    for a full-box resize in_size -> out_size: support = filter support * max(scale, 1), weights normalised per output
    coordinate in double, then (int)(+-0.5 + k * 2^22)."""
    scale = float(in_size) / out_size
    fscale = max(scale, 1.0)
    support = (2.0 if bicubic else 1.0) * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / fscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            w = _pil_filter((x + xmin - center + 0.5) * ss, bicubic)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            for x in range(xmax):
                kk[xx, x] /= ww
        bounds[xx] = (xmin, xmax)
    ki = np.where(kk < 0, (-0.5 + kk * (1 << PIL_PRECISION_BITS)).astype(np.int64),
                  (0.5 + kk * (1 << PIL_PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    return bounds, ki


def _pil_pass(img, out_size, axis, bicubic=False):
    """One resampling pass along `axis` (1 = horizontal, 0 = vertical) of a uint8 [H, W, C] image."""
    bounds, k = pil_bilinear_coeffs(img.shape[axis], out_size, bicubic)
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.zeros(shape, np.uint8)
    src = img.astype(np.int64)
    for o in range(out_size):
        lo, n = bounds[o]
        win = src[:, lo:lo + n, :] if axis == 1 else src[lo:lo + n, :, :]
        kw = k[o, :n].astype(np.int64)
        acc = (win * (kw[None, :, None] if axis == 1 else kw[:, None, None])).sum(axis=axis) + (1 << (PIL_PRECISION_BITS - 1))
        val = np.clip(acc >> PIL_PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 1:
            out[:, o, :] = val
        else:
            out[o, :, :] = val
    return out


def pil_resize_u8(img, out_h, out_w, bicubic=False):
    """Image.resize((out_w, out_h), BILINEAR | BICUBIC) of a uint8 [H, W, C] array: horizontal pass into an 8-bit
    intermediate, then the vertical pass (Resample.c:ImagingResampleInner; a pass whose size does not change is skipped, as
    Pillow does)."""
    img = np.ascontiguousarray(img)
    if img.shape[1] != out_w:
        img = _pil_pass(img, out_w, 1, bicubic)
    if img.shape[0] != out_h:
        img = _pil_pass(img, out_h, 0, bicubic)
    return img


def resize_center_crop_u8(img, size, bicubic=True):
    """transforms.Resize(size, BICUBIC) + transforms.CenterCrop(size) on a PIL image (dataset/dataset.py:251-253, 415-416):
    shorter side -> size, longer side int(size * long / short), crop origin int(round((extent - size) / 2.0))."""
    H, W = img.shape[:2]
    if W <= H:
        rw, rh = size, int(size * H / W)
    else:
        rh, rw = size, int(size * W / H)
    r = pil_resize_u8(img, rh, rw, bicubic)
    oy, ox = int(round((rh - size) / 2.0)), int(round((rw - size) / 2.0))
    return r[oy:oy + size, ox:ox + size]


def frames_to_tensor(batch_of_frames, params, size, mean, std):
    """dataset/dataset.py:123-161 for a whole batch: per frame crop (top, left, h, w) -> resize to size x size -> hflip if
    flagged -> ToTensor (u8 / 255, fp32) -> Normalize ((x - mean) / std, fp32); frames stacked on dim 1.
    Returns float32 [B, 3, T, size, size]."""
    B, T = len(batch_of_frames), len(batch_of_frames[0])
    out = np.zeros((B, 3, T, size, size), np.float32)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    n = 0
    for b in range(B):
        for t in range(T):
            top, left, h, w, flip = params[n]
            n += 1
            img = pil_resize_u8(batch_of_frames[b][t][top:top + h, left:left + w], size, size)
            if flip:
                img = img[:, ::-1]
            x = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
            out[b, :, t] = (x - m) / s
    return out


def spec_augment(fbank, params, amp, noise, mean, std, skip_norm=False):
    """dataset/dataset.py:281-294 (SpecAugment masks, applied to the raw filterbank) + :312-321 (normalisation, additive
    noise rand * amp / 10, roll along time) for one sample, numpy float32 with torch's roundings.
    params = (f0, f1, t0, t1, shift, add_noise); fbank / noise [T, F]."""
    f0, f1, t0, t1, shift, add_noise = [int(v) for v in params]
    x = np.array(fbank, np.float32, copy=True)
    x[:, f0:f1] = 0.0
    x[t0:t1, :] = 0.0
    if not skip_norm:
        x = (x - np.float32(mean)) / np.float32(std)
    if add_noise and noise is not None:
        x = x + (np.asarray(noise, np.float32) * np.float32(amp)) / np.float32(10)
    return np.roll(x, shift, axis=0)            # the reference only draws a non-zero shift together with the noise
