"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (/root/reference) in the build container.

    python tests/golden/make_golden.py          # needs /root/reference; torch 2.11 CPU

The reference has no tests/fixtures of its own, so these files are the pin for
oracle/mla_oracle.py and, through it, for the CUDA kernels. Nothing here runs on the GPU box
(the reference does not travel); only the small .npz outputs are committed.

Shims (no reference source is edited or copied):
  * stub modules `ml_collections` and `timm` (not installed); timm 0.4.5's Attention / Mlp, which the CAV-MAE block of the
    three-modality path calls, are restated in import_reference()
  * GSPlugin built via __new__ + a CPU `Pl` (its __init__ needs torch.cuda.FloatTensor)
  * nn.DataParallel(model, device_ids=[]) keeps the `.module` indirection on CPU
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("MLA_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.argv = ["main.py", "--ckpt_path", "x"]
    sys.path.insert(0, REF)
    import transformers  # noqa: F401  (must be imported BEFORE the timm stub)

    class ConfigDict(dict):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.__dict__ = self

        def copy_and_resolve_references(self):
            return ConfigDict(self)

    mlc = types.ModuleType("ml_collections")
    mlc.ConfigDict = ConfigDict
    mlc.config_dict = types.ModuleType("ml_collections.config_dict")
    mlc.config_dict.ConfigDict = ConfigDict
    mlc.config_dict.config_dict = mlc.config_dict          # models/m3ae.py:6 imports it by this name
    mlc.config_dict.placeholder = lambda *a, **k: None
    sys.modules["ml_collections"] = mlc
    sys.modules["ml_collections.config_dict"] = mlc.config_dict

    import torch.nn as nn
    timm = types.ModuleType("timm")
    for name in ("timm.models", "timm.models.layers", "timm.models.vision_transformer", "timm.data"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["timm"] = timm
    timm.models = sys.modules["timm.models"]
    timm.data = sys.modules["timm.data"]
    lay = sys.modules["timm.models.layers"]
    lay.to_2tuple = lambda x: (x, x)
    lay.trunc_normal_ = nn.init.trunc_normal_
    lay.DropPath = nn.Identity
    vt = sys.modules["timm.models.vision_transformer"]
    vt.PatchEmbed = vt.Block = nn.Identity              # models/cav_mae.py defines its own and overwrites these
    timm.models.vision_transformer = vt

    # timm==0.4.5 (requirements.txt:55) is not installed and not part of /root/reference: its two classes the CAV-MAE block
    # uses (cav_mae.py:15-16,93-94,101) are RESTATED here from the published 0.4.5 vision_transformer module. Everything
    # else on the modal3 path is the reference's own code. Parity of these two classes is therefore unpinned.
    class Attention(nn.Module):
        def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
            super().__init__()
            self.num_heads = num_heads
            self.scale = qk_scale or (dim // num_heads) ** -0.5
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.attn_drop = nn.Dropout(attn_drop)
            self.proj = nn.Linear(dim, dim)
            self.proj_drop = nn.Dropout(proj_drop)

        def forward(self, x):
            B, N, C = x.shape
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
            attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * self.scale).softmax(dim=-1)
            x = (self.attn_drop(attn) @ qkv[2]).transpose(1, 2).reshape(B, N, C)
            return self.proj_drop(self.proj(x))

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
            super().__init__()
            self.fc1 = nn.Linear(in_features, hidden_features or in_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
            self.drop = nn.Dropout(drop)

        def forward(self, x):
            return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))
    vt.Attention, vt.Mlp = Attention, Mlp
    sys.modules["timm.data"].create_transform = lambda *a, **k: None

    import main as ref_main
    from utils import utils as ref_utils
    return ref_main, ref_utils


def make_gs(ref_utils):
    import torch
    import torch.nn as nn

    class Wrap(nn.Module):           # makes named_parameters() yield "module.weight" 
        def __init__(self, m):
            super().__init__()
            self.module = m

    def new_plugin(D, dtype=torch.float32):
        gs = ref_utils.GSPlugin.__new__(ref_utils.GSPlugin)
        gs.Pl = torch.eye(D, dtype=dtype)
        gs.exp_count = 0
        return gs

    out = {}
    # KAT-1: bare nn.Linear -> the published hook is a no-op for any counter
    fc = nn.Linear(64, 6)
    fc.weight.grad = torch.randn(6, 64)
    g0 = fc.weight.grad.clone()
    gs = new_plugin(64)
    gs.before_update(fc, torch.randn(8, 64), 3, 10, 5)
    out["kat1_noop"] = np.array([bool(torch.equal(gs.Pl, torch.eye(64))), bool(torch.equal(fc.weight.grad, g0))])

    # teacher-forced chains: every step stores the fp32 inputs, the reference's fp32 outputs and
    # the outputs of the SAME reference code run in fp64 on the same inputs
    cases = [("signed_d64", 64, 6, 16, False), ("relu_d64", 64, 6, 16, True), ("signed_d128", 128, 11, 32, False),
             ("relu_d512", 512, 6, 64, True)]
    for name, D, C, B, relu in cases:
        g = torch.Generator().manual_seed(sum(map(ord, name)) + D)
        gs32 = new_plugin(D)
        steps = 1 if D >= 128 else 3
        for step in range(steps + 1):
            feat = torch.randn(B, D, generator=g)
            if relu:
                feat = feat.relu()
            grad = torch.randn(C, D, generator=g)
            bi, L, counter = step, 7, step            # counter 0 on the first call: skipped (utils.py:29)
            P_in = gs32.Pl.clone()
            fc32 = Wrap(nn.Linear(D, C))
            fc32.module.weight.grad = grad.clone()
            bias_grad = torch.randn(C, generator=g)
            fc32.module.bias.grad = bias_grad.clone()
            gs32.before_update(fc32, feat, bi, L, counter)
            assert torch.equal(fc32.module.bias.grad, bias_grad)          # bias never touched
            gs64 = new_plugin(D, torch.float64)
            gs64.Pl = P_in.double()
            fc64 = Wrap(nn.Linear(D, C).double())
            fc64.module.weight.grad = grad.double()
            gs64.before_update(fc64, feat.double(), bi, L, counter)
            k = "%s_s%d_" % (name, step)
            if D >= 512:      # keep the fixture small: inputs regenerate from the seed; store outputs on a row subset
                rows = np.arange(0, D, 16)
                out[k + "rows"] = rows
                out[k + "P_out32"] = gs32.Pl.detach().numpy()[rows]
                out[k + "P_out64"] = gs64.Pl.detach().numpy()[rows]
                out[k + "P_fro32"] = np.array(float(gs32.Pl.detach().norm()))
                out[k + "P_in_is_eye"] = np.array(bool(torch.equal(P_in, torch.eye(D))))
            else:
                out[k + "P_in"] = P_in.numpy()
                out[k + "P_out32"] = gs32.Pl.detach().numpy()
                out[k + "P_out64"] = gs64.Pl.detach().numpy()
            out[k + "feat"] = feat.numpy()
            out[k + "grad_in"] = grad.numpy()
            out[k + "grad_out32"] = fc32.module.weight.grad.numpy()
            out[k + "grad_out64"] = fc64.module.weight.grad.numpy()
            out[k + "meta"] = np.array([bi, L, counter, int(relu)])
            gs32.Pl = gs32.Pl.detach()
    np.savez_compressed(os.path.join(OUT, "gs_plugin.npz"), **out)


def make_fusion(ref_main):
    import torch
    import torch.nn.functional as F
    out = {}
    cases = [("b64c6m2", 64, 6, 2), ("b32c101m3", 32, 101, 3), ("b7c4m3", 7, 4, 3), ("b256c6m2", 256, 6, 2)]
    for name, B, C, M in cases:
        for scale in (0.3, 1.0, 3.0):
            g = torch.Generator().manual_seed(B * 131 + C * 7 + M + int(scale * 10))
            outs = [torch.randn(B, C, generator=g) * scale for _ in range(M)]
            label = torch.randint(0, C, (B,), generator=g)
            if M == 2:
                w = ref_main.calculate_gating_weights(*outs)
            else:
                w = ref_main.calculate_gating_weights3(*outs)
            assert all(x.dim() == 0 for x in w)
            fused = outs[0] * w[0] + outs[1] * w[1]              # main.py:643 / 646
            if M == 3:
                fused = fused + outs[2] * w[2]
            k = "%s_x%d_" % (name, int(scale * 10))
            for m in range(M):
                out[k + "out%d" % m] = outs[m].numpy()
            out[k + "label"] = label.numpy()
            out[k + "entropy"] = np.array([float(ref_main.calculate_entropy(o)) for o in outs], np.float32)
            out[k + "w"] = np.array([float(x) for x in w], np.float32)
            out[k + "fused"] = fused.numpy()
            preds = [F.softmax(fused, dim=1)] + [F.softmax(o, dim=1) for o in outs]       # main.py:653-657
            out[k + "argmax"] = np.stack([np.argmax(p.numpy(), axis=1) for p in preds]).astype(np.int32)
    # KAT-3: one +200 logit -> 0 * log 0 -> NaN weights
    g = torch.Generator().manual_seed(5)
    o1, o2 = torch.randn(64, 6, generator=g), torch.randn(64, 6, generator=g)
    o1[3, 2] = 200.0
    w = ref_main.calculate_gating_weights(o1, o2)
    out["nan_out0"], out["nan_out1"] = o1.numpy(), o2.numpy()
    out["nan_w"] = np.array([float(x) for x in w], np.float32)
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), **out)


def make_head():
    """The head is nn.Linear + nn.CrossEntropyLoss + autograd (main.py:130,432-435)."""
    import torch
    import torch.nn as nn
    out = {}
    for name, B, D, C in [("b16d64c6", 16, 64, 6), ("b8d128c101", 8, 128, 101), ("b32d256c6", 32, 256, 6)]:
        g = torch.Generator().manual_seed(B + D + C)
        fc = nn.Linear(D, C).double()
        with torch.no_grad():
            fc.weight.copy_(torch.randn(C, D, generator=g).double() * 0.05)
            fc.bias.copy_(torch.randn(C, generator=g).double() * 0.1)
        feat = torch.randn(B, D, generator=g).relu().double().requires_grad_(True)
        label = torch.randint(0, C, (B,), generator=g)
        logits = fc(feat)
        loss = nn.CrossEntropyLoss()(logits, label)
        loss.backward()
        k = name + "_"
        out[k + "feat"] = feat.detach().float().numpy()
        out[k + "W"] = fc.weight.detach().float().numpy()
        out[k + "b"] = fc.bias.detach().float().numpy()
        out[k + "label"] = label.numpy()
        out[k + "logits"] = logits.detach().numpy()
        out[k + "loss"] = np.array(float(loss))
        out[k + "dW"] = fc.weight.grad.numpy()
        out[k + "db"] = fc.bias.grad.numpy()
        out[k + "dfeat"] = feat.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "head.npz"), **out)


def make_av(ref_main, ref_utils):
    import torch
    import torch.nn as nn
    from torch.optim import SGD
    from torch.optim.lr_scheduler import StepLR
    out = {}
    args = ref_main.get_arguments()
    args.dataset, args.lorb, args.gs_flag, args.dynamic = "CREMAD", "base", True, True
    args.fusion_method, args.modulation, args.modal3, args.clip = "concat", "Normal", False, False

    def build():
        ref_utils.setup_seed(0)
        model = ref_main.AVClassifier(args)
        model.apply(ref_utils.weight_init)
        return model

    def batches(n, B, seed, hw, img):
        g = torch.Generator().manual_seed(seed)
        res = []
        for _ in range(n):
            spec = torch.randn(B, *hw, generator=g)
            image = torch.randn(B, 3, 2, img, img, generator=g)
            label = torch.randint(0, 6, (B,), generator=g)
            res.append((spec, image, label, torch.zeros(B, 1, dtype=torch.long)))
        return res

    model = build()
    sd = model.state_dict()
    out["n_state"] = np.array(len(sd))
    out["state_names"] = np.array(list(sd.keys()))
    out["state_sum"] = np.array([float(v.double().sum()) for v in sd.values()])
    out["state_abs"] = np.array([float(v.double().abs().sum()) for v in sd.values()])
    out["n_params"] = np.array([sum(p.numel() for p in model.audio_net.parameters()),
                                sum(p.numel() for p in model.visual_net.parameters()),
                                sum(p.numel() for p in model.fusion_module.parameters())])

    # forward, train mode (batch statistics) and eval mode (running statistics), small spatial size
    (spec, image, label, _), = batches(1, 2, 11, (65, 48), 64)
    model.train()
    a, v = model(spec.unsqueeze(1).float(), image.float())
    out["fwd_train_a"], out["fwd_train_v"] = a.detach().numpy(), v.detach().numpy()
    out["fwd_bn1_running_mean"] = model.audio_net.bn1.running_mean.numpy().copy()
    out["fwd_bn1_running_var"] = model.audio_net.bn1.running_var.numpy().copy()
    model.eval()
    with torch.no_grad():
        a, v = model(spec.unsqueeze(1).float(), image.float())
    out["fwd_eval_a"], out["fwd_eval_v"] = a.numpy(), v.numpy()
    # KAT-4 feature-map shapes at the BASELINE.json input size
    with torch.no_grad():
        fa = model.audio_net(torch.zeros(1, 1, 257, 188))
        fv = model.visual_net(torch.zeros(1, 3, 2, 224, 224))
    out["kat4_shapes"] = np.array(list(fa.shape) + list(fv.shape))

    def run_epoch(bl, fire):
        model = build()
        if fire:   # make the published hook fire: named_parameters() must yield "module.weight" 
            class Wrap(nn.Module):
                def __init__(self, m):
                    super().__init__()
                    self.module = m

                def forward(self, x):
                    return self.module(x)
            model.fusion_module.fc_out = Wrap(model.fusion_module.fc_out)
        dp = nn.DataParallel(model, device_ids=[])
        opt = SGD(dp.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
        sch = StepLR(opt, 70, 0.1)
        gs = ref_utils.GSPlugin.__new__(ref_utils.GSPlugin)
        gs.Pl = torch.eye(512)
        gs.exp_count = 0
        losses = ref_main.train_epoch(args, 0, dp, torch.device("cpu"), bl, opt, sch, gs_plugin=gs, gs_flag=True,
                                      av_alpha=0.55)
        accs_dyn = ref_main.valid(args, dp, torch.device("cpu"), bl, gs_flag=True, av_alpha=0.55)
        args.dynamic = False
        accs_fix = ref_main.valid(args, dp, torch.device("cpu"), bl, gs_flag=True, av_alpha=0.55)
        args.dynamic = True
        fc = model.fusion_module.fc_out.module if fire else model.fusion_module.fc_out
        return dict(losses=np.array(losses), accs_dyn=np.array(accs_dyn), accs_fix=np.array(accs_fix),
                    exp_count=np.array(gs.exp_count), Pl_is_eye=np.array(bool(torch.equal(gs.Pl.detach(), torch.eye(512)))),
                    fc_w=fc.weight.detach().numpy().copy(), fc_b=fc.bias.detach().numpy().copy(),
                    a_conv1_sum=np.array(float(model.audio_net.conv1.weight.double().sum())),
                    v_conv1_sum=np.array(float(model.visual_net.conv1.weight.double().sum())),
                    Pl_fro=np.array(float(gs.Pl.detach().norm())))

    small = batches(3, 4, 1, (65, 48), 64)
    for tag, fire in (("small_noop", False), ("small_fire", True)):
        for k, v in run_epoch(small, fire).items():
            out["%s_%s" % (tag, k)] = v
    # single-step epochs: the losses depend on the forward pass and the head update only (no encoder
    # update has happened yet), so they pin forward parity without the chaotic TF32 trajectory
    out["small_step1_losses"] = run_epoch(small[:1], False)["losses"]
    # KAT-6: three B=4 batches at the full BASELINE.json size
    full = batches(3, 4, 1, (257, 188), 224)
    out["full_step1_losses"] = run_epoch(full[:1], False)["losses"]
    for k, v in run_epoch(full, False).items():
        if k in ("losses", "accs_dyn", "accs_fix", "exp_count", "Pl_is_eye"):
            out["full_noop_%s" % k] = v
    np.savez_compressed(os.path.join(OUT, "av_classifier.npz"), **out)


def make_av_joint(ref_main, ref_utils):
    """Joint training without --gs_flag (main.py:165-168, 269-310, 412-418) and its OGM / OGM-GE modulation
    (main.py:312-410), executed by the reference's own train_epoch / valid on small CREMA-D-shaped batches (SURVEY
    section 8 f1 / f2). OGM_GE draws its noise from torch's default CPU generator, re-seeded before the epoch."""
    import torch
    import torch.nn as nn
    from torch.optim import SGD
    from torch.optim.lr_scheduler import StepLR
    out = {}
    args = ref_main.get_arguments()
    args.dataset, args.lorb, args.gs_flag, args.dynamic = "CREMAD", "base", False, False
    args.fusion_method, args.modal3, args.clip, args.use_tensorboard = "concat", False, False, False
    args.alpha, args.modulation_starts, args.modulation_ends = 0.8, 0, 50

    def batches(n, B, seed, hw, img):
        g = torch.Generator().manual_seed(seed)
        res = []
        for _ in range(n):
            spec = torch.randn(B, *hw, generator=g)
            image = torch.randn(B, 3, 2, img, img, generator=g)
            label = torch.randint(0, 6, (B,), generator=g)
            res.append((spec, image, label, torch.zeros(B, 1, dtype=torch.long)))
        return res

    bl = batches(3, 4, 3, (65, 48), 64)
    for modulation, epoch in (("Normal", 0), ("OGM", 0), ("OGM", 51), ("OGM_GE", 0)):
        args.modulation = modulation
        ref_utils.setup_seed(0)
        model = ref_main.AVClassifier(args)
        model.apply(ref_utils.weight_init)
        if modulation == "Normal":
            out["head_shape"] = np.array(model.fusion_module.fc_out.weight.shape)
        dp = nn.DataParallel(model, device_ids=[])
        opt = SGD(dp.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
        sch = StepLR(opt, 70, 0.1)
        torch.manual_seed(1234)                      # the OGM_GE noise stream
        losses = ref_main.train_epoch(args, epoch, dp, torch.device("cpu"), bl, opt, sch)
        accs = ref_main.valid(args, dp, torch.device("cpu"), bl)
        tag = "%s_e%d_" % (modulation, epoch)
        out[tag + "losses"] = np.array(losses)
        out[tag + "accs"] = np.array(accs)
        sd = model.state_dict()
        out[tag + "fc_w"] = sd["fusion_module.fc_out.weight"].numpy().copy()
        # row / column subsets keep the fixture small; the initial values regenerate from the seed
        out[tag + "a_l4_conv"] = sd["audio_net.layer4.1.conv2.weight"].numpy()[::64, ::8].copy()
        out[tag + "v_l1_conv"] = sd["visual_net.layer1.0.conv1.weight"].numpy()[::4].copy()
        out[tag + "a_bn1_w"] = sd["audio_net.bn1.weight"].numpy().copy()
    # the coefficient rule alone, two and three modalities (main.py:373-384, 315-334), from seeded logits
    tanh, relu, softmax = nn.Tanh(), nn.ReLU(inplace=True), nn.Softmax(dim=1)
    for name, B, C, M in (("c2", 16, 6, 2), ("c3", 12, 4, 3), ("c3t", 12, 4, 3), ("c2a", 16, 6, 2), ("c2v", 16, 6, 2),
                           ("c3v", 12, 4, 3)):
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        outs = [torch.randn(B, C, generator=g) for _ in range(M)]
        label = torch.randint(0, C, (B,), generator=g)
        if name == "c3t":
            outs[2][torch.arange(B), label] += 4.0     # text dominates
        if name == "c2a":
            outs[0][torch.arange(B), label] += 3.0     # audio dominates
        if name in ("c2v", "c3v"):
            outs[1][torch.arange(B), label] += 3.5     # visual dominates
        score = [sum([softmax(o)[i][label[i]] for i in range(B)]) for o in outs]
        if M == 2:
            ratio_v = score[1] / score[0]
            ratio_a = 1 / ratio_v
            coeff = [1.0, 1.0]
            if ratio_v > 1:
                coeff[1] = float(1 - tanh(args.alpha * relu(ratio_v)))
            else:
                coeff[0] = float(1 - tanh(args.alpha * relu(ratio_a)))
        else:
            ratio_v = score[1] / (score[0] + score[2])
            ratio_a = score[0] / (score[1] + score[2])
            ratio_t = score[2] / (score[1] + score[0])
            coeff = [1.0, 1.0, 1.0]
            if ratio_v > 1:
                coeff[1] = float(1 - tanh(args.alpha * relu(ratio_v)))
            elif ratio_t > 1:
                coeff[2] = float(1 - tanh(args.alpha * relu(ratio_t)))
            else:
                coeff[0] = float(1 - tanh(args.alpha * relu(ratio_a)))
        for m in range(M):
            out["%s_out%d" % (name, m)] = outs[m].numpy()
        out[name + "_label"] = label.numpy()
        out[name + "_score"] = np.array([float(x) for x in score], np.float32)
        out[name + "_coeff"] = np.array(coeff, np.float32)
    out["alpha"] = np.array(args.alpha)
    np.savez_compressed(os.path.join(OUT, "av_joint.npz"), **out)


M3AE_TINY = dict(model_type=None, emb_dim=64, depth=2, num_heads=2)
M3AE_VOCAB = 512


def m3ae_batches(n, B, seed, L=12, img=32, n_classes=101, vocab=M3AE_VOCAB):
    import torch
    g = torch.Generator().manual_seed(seed)
    res = []
    for _ in range(n):
        token = torch.randint(0, vocab, (B, 1, L), generator=g)
        n_valid = torch.randint(3, L + 1, (B,), generator=g)
        pm = (torch.arange(L)[None, :] >= n_valid[:, None]).long()[:, None, :]
        image = torch.randn(B, 3, img, img, generator=g)
        label = torch.randint(0, n_classes, (B,), generator=g)
        res.append((token, pm, image, label, torch.zeros(B, 1, dtype=torch.long)))
    return res


def make_m3ae(ref_main, ref_utils):
    """--lorb m3ae --gs_flag (BASELINE.json configs[2]) on a tiny encoder configuration (emb 64, depth 2, 2 heads,
    vocabulary 512), executed by the reference's own M3AEClassifier.forward / MaskedMultimodalAutoencoder /
    train_epoch / valid. Shims (generator only): DropPath.forward -> identity (it returns None), tensors
    sent `.to(cuda:0)` stay on the CPU, the classifier object is assembled without its __init__ (which torch.load()s
    hard-coded placeholder checkpoint paths and fixes the head at 768)."""
    import torch
    import torch.nn as nn
    from torch.optim import SGD
    from torch.optim.lr_scheduler import StepLR
    from models import m3ae as ref_m3ae
    from models import basic_model as ref_bm
    from models.fusion_modules import ConcatFusion
    ref_m3ae.DropPath.forward = lambda self, input, deterministic=False: input
    real_to = torch.Tensor.to

    def cpu_to(self, *a, **k):
        a = tuple(torch.device("cpu") if isinstance(x, torch.device) and x.type == "cuda" else x for x in a)
        return real_to(self, *a, **k)
    torch.Tensor.to = cpu_to
    out = {}
    args = ref_main.get_arguments()
    args.dataset, args.lorb, args.gs_flag, args.dynamic = "Food101", "m3ae", True, True
    args.fusion_method, args.modulation, args.modal3, args.clip = "concat", "Normal", False, False
    cfg = sys.modules["ml_collections"].ConfigDict(M3AE_TINY)

    def build():
        ref_utils.setup_seed(0)
        model = ref_bm.M3AEClassifier.__new__(ref_bm.M3AEClassifier)
        nn.Module.__init__(model)
        model.fusion_module = ConcatFusion(input_dim=M3AE_TINY["emb_dim"], output_dim=101)      # basic_model.py:150
        model.mae_a = ref_m3ae.MaskedMultimodalAutoencoder(text_vocab_size=M3AE_VOCAB, config_updates=cfg)
        model.mae_v = ref_m3ae.MaskedMultimodalAutoencoder(text_vocab_size=M3AE_VOCAB, config_updates=cfg)
        model.args = args
        return model

    model = build()
    sd = model.state_dict()
    for k, v in sd.items():
        out["state/" + k] = v.numpy().copy()
    # the positional tables
    out["pos1d_64_12"] = ref_m3ae.get_1d_sincos_pos_embed(64, 12)
    out["pos2d_64_4"] = ref_m3ae.get_2d_sincos_pos_embed(64, 4)
    out["pos2d_768_256"] = ref_m3ae.get_2d_sincos_pos_embed(768, 256)[:, ::17, ::5]
    (token, pm, image, label, _), = m3ae_batches(1, 4, 31)
    a, v = model(token, pm, image)
    out["fwd_a"], out["fwd_v"] = a.detach().numpy(), v.detach().numpy()
    # encoder-output gradients of a fixed scalar (pins the backward of the restatement)
    (a.square().sum() + v.square().sum()).backward()
    for k in ("mae_a.encoder.blocks.0.attention.qkv_linear.weight", "mae_v.image_embedding.weight",
              "mae_v.encoder.blocks.1.transformer_mlp.fc2.weight", "mae_a.cls_token", "mae_v.encoder.layer_norm.weight"):
        out["grad/" + k] = dict(model.named_parameters())[k].grad.numpy().copy()
    out["grad_text_embedding_rows"] = model.mae_a.text_embedding.weight.grad.abs().sum(1).numpy()

    def run_epoch(bl):
        model = build()
        dp = nn.DataParallel(model, device_ids=[])
        opt = SGD(dp.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
        sch = StepLR(opt, 70, 0.1)
        gs = ref_utils.GSPlugin.__new__(ref_utils.GSPlugin)
        gs.Pl = torch.eye(M3AE_TINY["emb_dim"])
        gs.exp_count = 0
        losses = ref_main.train_epoch(args, 0, dp, torch.device("cpu"), bl, opt, sch, gs_plugin=gs, gs_flag=True,
                                      av_alpha=0.55)
        accs_dyn = ref_main.valid(args, dp, torch.device("cpu"), bl, gs_flag=True, av_alpha=0.55)
        args.dynamic = False
        accs_fix = ref_main.valid(args, dp, torch.device("cpu"), bl, gs_flag=True, av_alpha=0.55)
        args.dynamic = True
        fin = model.state_dict()
        return dict(losses=np.array(losses), accs_dyn=np.array(accs_dyn), accs_fix=np.array(accs_fix),
                    exp_count=np.array(gs.exp_count), fc_w=fin["fusion_module.fc_out.weight"].numpy().copy(),
                    qkv0_a=fin["mae_a.encoder.blocks.0.attention.qkv_linear.weight"].numpy().copy(),
                    fc2_v=fin["mae_v.encoder.blocks.1.transformer_mlp.fc2.weight"].numpy().copy())

    bl = m3ae_batches(3, 8, 7)
    for k, v in run_epoch(bl[:1]).items():
        out["step1_" + k] = v
    for k, v in run_epoch(bl).items():
        out["step3_" + k] = v
    torch.Tensor.to = real_to
    np.savez_compressed(os.path.join(OUT, "m3ae.npz"), **out)

    # seeded initialisation of the full-size 'base' encoder: per-tensor sums only (the module must draw the same
    # random numbers in the same order to reproduce them)
    ref_utils.setup_seed(0)
    enc = ref_m3ae.MaskedMultimodalAutoencoder(text_vocab_size=30522,
                                               config_updates=sys.modules["ml_collections"].ConfigDict(dict(model_type="base")))
    sd = enc.state_dict()
    np.savez_compressed(os.path.join(OUT, "m3ae_base_init.npz"), names=np.array(list(sd.keys())),
                        shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
                        sums=np.array([float(v.double().sum()) for v in sd.values()]),
                        abs_sums=np.array([float(v.double().abs().sum()) for v in sd.values()]))


def make_m3ae_dh64(ref_main, ref_utils):
    """A second encoder geometry for the oracle: head width 64 (the 'base' head width; the tiny configuration above has 32),
    emb 128, 2 heads, 1 block, vocabulary 64: text and image representations of the reference's encoder + their gradients."""
    import torch
    from models import m3ae as ref_m3ae
    ref_m3ae.DropPath.forward = lambda self, input, deterministic=False: input
    real_to = torch.Tensor.to

    def cpu_to(self, *a, **k):
        a = tuple(torch.device("cpu") if isinstance(x, torch.device) and x.type == "cuda" else x for x in a)
        return real_to(self, *a, **k)
    torch.Tensor.to = cpu_to
    try:
        ref_utils.setup_seed(3)
        cfg = sys.modules["ml_collections"].ConfigDict(dict(model_type=None, emb_dim=128, depth=1, num_heads=2))
        enc = ref_m3ae.MaskedMultimodalAutoencoder(text_vocab_size=64, config_updates=cfg)
        out = {"state/" + k: v.numpy().copy() for k, v in enc.state_dict().items()}
        g = torch.Generator().manual_seed(17)
        text = torch.randint(0, 64, (3, 70), generator=g)                # 71 tokens with CLS: two 64-key tiles
        pm = (torch.arange(70)[None, :] >= torch.tensor([70, 33, 1])[:, None]).long()
        image = torch.randn(3, 9, 768, generator=g)
        t = enc.forward_representation(None, text, pm)
        v = enc.forward_representation(image, None, None)
        out["rep_text"], out["rep_image"] = t.detach().numpy(), v.detach().numpy()
        # a fixed random linear functional of the outputs (NOT a function of the LayerNorm-ed rows' norms, whose gradient
        # cancels to rounding noise)
        wt, wv = torch.randn(t.shape, generator=g), torch.randn(v.shape, generator=g)
        ((t * wt).sum() + (v * wv).sum()).backward()
        for k in ("encoder.blocks.0.attention.qkv_linear.weight", "encoder.blocks.0.attention.fc.bias", "cls_token",
                  "image_embedding.weight", "encoder.blocks.0.layer_norm2.weight"):
            out["grad/" + k] = dict(enc.named_parameters())[k].grad.numpy().copy()
    finally:
        torch.Tensor.to = real_to
    np.savez_compressed(os.path.join(OUT, "m3ae_dh64.npz"), **out)


CAV_TINY = dict(img_size=32, audio_length=64, embed_dim=64, modality_specific_depth=1, num_heads=2)


def modal3_batches(n, B, seed, L=12, img=32, T=64, n_classes=4, vocab=M3AE_VOCAB):
    import torch
    g = torch.Generator().manual_seed(seed)
    res = []
    for _ in range(n):
        token = torch.randint(0, vocab, (B, 1, L), generator=g)
        n_valid = torch.randint(3, L + 1, (B,), generator=g)
        pm = (torch.arange(L)[None, :] >= n_valid[:, None]).long()[:, None, :]
        image = torch.randn(B, 3, img, img, generator=g)
        spec = torch.randn(B, T, 128, generator=g)
        label = torch.randint(0, n_classes, (B,), generator=g)
        res.append((token, pm, image, spec, label, torch.zeros(B, 1, dtype=torch.long)))
    return res


def make_modal3(ref_main, ref_utils):
    """--lorb m3ae --modal3 --gs_flag (BASELINE.json configs[3]) on tiny encoders: the reference's own CAVMAEFT (PatchEmbed,
    Block wiring, embeddings, initialisation, forward_feat), Modal3Classifier.forward, train_epoch and valid; timm's
    Attention / Mlp restated (import_reference). Same generator-side shims as make_m3ae."""
    import torch
    import torch.nn as nn
    from torch.optim import SGD
    from torch.optim.lr_scheduler import StepLR
    from models import m3ae as ref_m3ae
    from models import cav_mae as ref_cav
    from models import basic_model as ref_bm
    from models.fusion_modules import ConcatFusion3
    ref_m3ae.DropPath.forward = lambda self, input, deterministic=False: input
    real_to = torch.Tensor.to

    def cpu_to(self, *a, **k):
        a = tuple(torch.device("cpu") if isinstance(x, torch.device) and x.type == "cuda" else x for x in a)
        return real_to(self, *a, **k)
    torch.Tensor.to = cpu_to
    out = {}
    args = ref_main.get_arguments()
    args.dataset, args.lorb, args.gs_flag, args.dynamic, args.modal3 = "IEMOCAP", "m3ae", True, True, True
    args.fusion_method, args.modulation, args.clip = "concat", "Normal", False
    cfg = sys.modules["ml_collections"].ConfigDict(M3AE_TINY)

    def build():
        ref_utils.setup_seed(0)
        model = ref_bm.Modal3Classifier.__new__(ref_bm.Modal3Classifier)
        nn.Module.__init__(model)
        model.fusion_module = ConcatFusion3(input_dim=M3AE_TINY["emb_dim"], output_dim=4)        # basic_model.py:218
        model.mae_a = ref_cav.CAVMAEFT(4, **CAV_TINY)
        model.mae_v = ref_m3ae.MaskedMultimodalAutoencoder(text_vocab_size=M3AE_VOCAB, config_updates=cfg)
        model.mae_t = ref_m3ae.MaskedMultimodalAutoencoder(text_vocab_size=M3AE_VOCAB, config_updates=cfg)
        model.args = args
        return model

    model = build()
    for k, v in model.state_dict().items():
        out["state/" + k] = v.numpy().copy()
    out["pos_embed_a_8x64_768"] = ref_cav.get_2d_sincos_pos_embed(768, 8, 64)[::7, ::5].astype(np.float32)
    (token, pm, image, spec, label, _), = modal3_batches(1, 4, 41)
    a, v, t = model(token, pm, image, spec)
    out["fwd_a"], out["fwd_v"], out["fwd_t"] = a.detach().numpy(), v.detach().numpy(), t.detach().numpy()
    a.square().sum().backward()
    named = dict(model.named_parameters())
    for k in ("mae_a.patch_embed_a.proj.weight", "mae_a.pos_embed_a", "mae_a.modality_a", "mae_a.blocks_a.0.attn.qkv.weight",
              "mae_a.blocks_u.0.norm1_a.weight", "mae_a.blocks_u.0.mlp.fc2.weight", "mae_a.norm_a.bias"):
        out["grad/" + k] = named[k].grad.numpy().copy()
    out["grad_none"] = np.array([k for k, p_ in named.items() if k.startswith("mae_a.") and p_.grad is None])

    def run_epoch(bl):
        model = build()
        dp = nn.DataParallel(model, device_ids=[])
        opt = SGD(dp.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
        sch = StepLR(opt, 70, 0.1)
        gs = ref_utils.GSPlugin.__new__(ref_utils.GSPlugin)
        gs.Pl = torch.eye(M3AE_TINY["emb_dim"])
        gs.exp_count = 0
        losses = ref_main.train_epoch(args, 0, dp, torch.device("cpu"), bl, opt, sch, gs_plugin=gs, gs_flag=True,
                                      av_alpha=0.55)
        accs_dyn = ref_main.valid(args, dp, torch.device("cpu"), bl, gs_flag=True, av_alpha=0.55)
        args.dynamic = False
        accs_fix = ref_main.valid(args, dp, torch.device("cpu"), bl, gs_flag=True, av_alpha=0.55)
        args.dynamic = True
        fin = model.state_dict()
        return dict(losses=np.array(losses), accs_dyn=np.array(accs_dyn), accs_fix=np.array(accs_fix),
                    exp_count=np.array(gs.exp_count), fc_w=fin["fusion_module.fc_out.weight"].numpy().copy(),
                    qkv_a=fin["mae_a.blocks_a.0.attn.qkv.weight"].numpy().copy(),
                    patch_a=fin["mae_a.patch_embed_a.proj.weight"].numpy().copy(),
                    unused_v=fin["mae_a.blocks_v.0.attn.qkv.weight"].numpy().copy(),
                    fc2_t=fin["mae_t.encoder.blocks.1.transformer_mlp.fc2.weight"].numpy().copy())

    bl = modal3_batches(3, 8, 9)
    for k, v in run_epoch(bl[:1]).items():
        out["step1_" + k] = v
    for k, v in run_epoch(bl).items():
        out["step3_" + k] = v
    torch.Tensor.to = real_to
    np.savez_compressed(os.path.join(OUT, "modal3.npz"), **out)


if __name__ == "__main__":
    ref_main, ref_utils = import_reference()
    import torch
    torch.set_num_threads(8)
    if os.environ.get("MLA_GOLDEN_ONLY_JOINT"):
        make_av_joint(ref_main, ref_utils)
        sys.exit(0)
    if not any(os.environ.get(k) for k in ("MLA_GOLDEN_ONLY_M3AE", "MLA_GOLDEN_ONLY_MODAL3", "MLA_GOLDEN_ONLY_DH64")):
        make_gs(ref_utils)
        make_fusion(ref_main)
        make_head()
        make_av(ref_main, ref_utils)
        make_av_joint(ref_main, ref_utils)
    if os.environ.get("MLA_GOLDEN_ONLY_DH64"):
        make_m3ae_dh64(ref_main, ref_utils)
    else:
        if not os.environ.get("MLA_GOLDEN_ONLY_MODAL3"):
            make_m3ae(ref_main, ref_utils)
            make_m3ae_dh64(ref_main, ref_utils)
        make_modal3(ref_main, ref_utils)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
