"""The alternating training step and evaluation through the public API (train_epoch / valid /
AVClassifier.forward) against the oracle and the reference fixtures. Tolerance: rel 1e-3 on
features, losses and accuracies (north_star: fp32/TF32, rel 1e-3)."""
import argparse

import numpy as np
import pytest
import torch

from oracle import mla_oracle as orc

pytestmark = pytest.mark.gpu

# Tolerances. Encoder convolutions run with TF32 operands (10-bit mantissa, fp32 accumulate) — the same
# arithmetic torch's cuDNN path uses for the reference on a GPU (torch default cudnn.allow_tf32=True).
#  * FORWARD-ONLY quantities (features, the losses of a single step): rel 1e-3 at the BASELINE.json
#    input size (north_star). The deliberately tiny fixtures (B=2..4, 64x64 frames: BatchNorm over a
#    handful of samples in layer4 amplifies rounding) get 3e-3; torch's own cuDNN-TF32 path sits
#    1.6e-3 from its fp32 path on them (tests/tools/tf32_noise.py, profiles/).
#  * MULTI-STEP trajectories: TF32 rounding flips ReLU masks, which puts a ~10 % Frobenius error on the
#    encoder gradients of ANY TF32 implementation (torch cuDNN-TF32 vs torch fp32: median 1.05e-1,
#    cosine 0.994; ours vs torch fp32: 1.03e-1..1.17e-1). After one SGD step the losses of two correct
#    implementations therefore differ in the third digit. The criterion is CALIBRATED like the GS one
#    (SURVEY F10): our deviation from the reference's fp32 fixture must be within 2x the deviation of
#    the reference's own arithmetic run on this GPU under torch's TF32 default (+1e-3).
FEAT_TOL = 3e-3
STEP1_TOL_FULL = 1e-3
STEP1_TOL_SMALL = 3e-3


def relf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _args(dynamic=True):
    return argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=dynamic, lorb="base", modal3=False, clip=False)


def _model(built_lib):
    import mla_b200
    mla_b200.setup_seed(0)
    net = mla_b200.AVClassifier(_args()).apply(mla_b200.weight_init)
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return mla_b200.ModuleHolder(net.cuda()), state


def _batches(n, B, seed, hw, img):
    gen = torch.Generator().manual_seed(seed)
    res = []
    for _ in range(n):
        spec = torch.randn(B, *hw, generator=gen)
        image = torch.randn(B, 3, 2, img, img, generator=gen)
        label = torch.randint(0, 6, (B,), generator=gen)
        res.append((spec, image, label, torch.zeros(B, 1, dtype=torch.long)))
    return res


def _torch_tf32_epoch(state, batches, fire, av_alpha=0.55):
    """The reference's arithmetic on THIS GPU under torch's defaults (cuDNN convolutions with TF32):
    the oracle's torch restatement handed CUDA tensors. Calibrates the multi-step criterion."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        o = orc.AVOracle({k: v.cuda() for k, v in state.items()}, force_projection=fire)
        return o.train_epoch([(b[0].cuda(), b[1].cuda(), b[2].cuda()) for b in batches], av_alpha=av_alpha), o
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def _calibrated(ours, ref_fp32, torch_tf32, what):
    ours, ref_fp32, torch_tf32 = (np.asarray(x, np.float64) for x in (ours, ref_fp32, torch_tf32))
    d_ours = np.abs(ours - ref_fp32) / np.abs(ref_fp32)
    d_torch = np.abs(torch_tf32 - ref_fp32) / np.abs(ref_fp32)
    print("%s: ours %s | fp32 fixture %s | torch-TF32 %s | dev ours %s torch %s" % (what, ours, ref_fp32, torch_tf32,
                                                                                 d_ours, d_torch))
    assert (d_ours <= 2 * d_torch.max() + 1e-3).all(), (what, d_ours, d_torch)


def test_forward_matches_reference_fixture(built_lib, golden):
    g = golden("av_classifier")
    model, _ = _model(built_lib)
    (spec, image, _, _), = _batches(1, 2, 11, (65, 48), 64)
    model.train()
    a, v = model(spec.unsqueeze(1).cuda(), image.cuda())
    assert a.requires_grad and v.requires_grad and a.shape == (2, 512)
    ea, ev = relf(a.detach().cpu().numpy(), g["fwd_train_a"]), relf(v.detach().cpu().numpy(), g["fwd_train_v"])
    print("feature rel-F error (train):", ea, ev)
    assert ea < FEAT_TOL and ev < FEAT_TOL
    assert np.allclose(model.module.audio_net.bn1.running_mean.cpu().numpy(), g["fwd_bn1_running_mean"], atol=1e-4)
    assert np.allclose(model.module.audio_net.bn1.running_var.cpu().numpy(), g["fwd_bn1_running_var"], rtol=1e-4)
    assert int(model.module.audio_net.bn1.num_batches_tracked) == 1
    model.eval()
    with torch.no_grad():
        a, v = model(spec.unsqueeze(1).cuda(), image.cuda())
    ea, ev = relf(a.cpu().numpy(), g["fwd_eval_a"]), relf(v.cpu().numpy(), g["fwd_eval_v"])
    print("feature rel-F error (eval):", ea, ev)
    assert ea < FEAT_TOL and ev < FEAT_TOL


def test_forward_full_size_vs_oracle(built_lib):
    """Features at the BASELINE.json input size (1x257x188, 2 frames 3x224x224): rel 1e-3 (Frobenius)
    against the oracle's fp32 restatement (torch fp32 on the GPU, TF32 off)."""
    model, state = _model(built_lib)
    spec, image, _ = orc.synthetic_av_batch(8, 21)
    spec, image = spec.cuda(), image.cuda()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd = {k: v.cuda() for k, v in state.items()}
        with torch.no_grad():
            ra, rv = orc.av_forward(sd, spec.unsqueeze(1), image, training=True)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    model.train()
    a, v = model(spec.unsqueeze(1), image)
    ea, ev = relf(a.detach().cpu().numpy(), ra.cpu().numpy()), relf(v.detach().cpu().numpy(), rv.cpu().numpy())
    print("full-size feature rel-F error:", ea, ev)
    assert ea < 1e-3 and ev < 1e-3
    assert bool((a >= 0).all())                                        # KAT-4: post-ReLU pooled features


@pytest.mark.parametrize("case", ["small", "full"])
def test_single_step_losses_match_reference_fixture(built_lib, golden, case):
    """One alternating step (as published: the hook is a no-op): the three returned losses
    depend on the forward pass and the head update only -> forward-level tolerance."""
    import mla_b200
    g = golden("av_classifier")
    model, _ = _model(built_lib)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin()
    batches = _batches(3, 4, 1, (65, 48), 64)[:1] if case == "small" else _batches(3, 4, 1, (257, 188), 224)[:1]
    losses = mla_b200.train_epoch(_args(), 0, model, torch.device("cuda"), batches, opt, sch, gs_plugin=gs,
                                  gs_flag=True, av_alpha=0.55)
    ref = g[case + "_step1_losses"]
    print("single-step losses", losses, "fixture", ref)
    assert np.allclose(losses, ref, rtol=STEP1_TOL_SMALL if case == "small" else STEP1_TOL_FULL), (losses, ref)
    assert gs.exp_count == 2 and bool(torch.equal(gs.Pl, torch.eye(512, device="cuda")))      # KAT-1


@pytest.mark.parametrize("fire", [False, True])
def test_train_epoch_and_valid_match_reference_fixture(built_lib, golden, fire):
    import mla_b200
    g = golden("av_classifier")
    tag = "small_fire_" if fire else "small_noop_"
    model, state = _model(built_lib)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin(force_projection=fire)
    batches = _batches(3, 4, 1, (65, 48), 64)
    dev = torch.device("cuda")
    losses = mla_b200.train_epoch(_args(), 0, model, dev, batches, opt, sch, gs_plugin=gs, gs_flag=True, av_alpha=0.55)
    t_losses, t_orc = _torch_tf32_epoch(state, batches, fire)
    _calibrated(losses, g[tag + "losses"], t_losses, "3-step losses (small, fire=%s)" % fire)
    assert gs.exp_count == 6
    assert bool(torch.equal(gs.Pl, torch.eye(512, device="cuda"))) == bool(g[tag + "Pl_is_eye"])
    assert abs(float(gs.Pl.norm()) - 1) < 1e-4 or not fire
    fcw = model.module.fusion_module.fc_out.weight.detach().cpu().numpy()
    d_ours = relf(fcw, g[tag + "fc_w"])
    d_torch = relf(t_orc.sd["fusion_module.fc_out.weight"].detach().cpu().numpy(), g[tag + "fc_w"])
    print("head weight rel-F deviation from the fp32 fixture: ours %.3e torch-TF32 %.3e" % (d_ours, d_torch))
    assert d_ours <= 2 * d_torch + 1e-3
    accs = mla_b200.valid(_args(True), model, dev, batches, gs_flag=True, av_alpha=0.55)
    accs_fix = mla_b200.valid(_args(False), model, dev, batches, gs_flag=True, av_alpha=0.55)
    # 12 samples: accuracies are multiples of 1/12; allow one borderline sample
    assert np.abs(np.array(accs) - g[tag + "accs_dyn"]).max() <= 1 / 12 + 1e-9
    assert np.abs(np.array(accs_fix) - g[tag + "accs_fix"]).max() <= 1 / 12 + 1e-9


def test_kat6_full_size_three_steps(built_lib, golden):
    """SURVEY KAT-6: three B=4 batches at the BASELINE.json input size -> (1.5840, 1.5922, 1.5740)."""
    import mla_b200
    g = golden("av_classifier")
    model, state = _model(built_lib)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin()
    batches = _batches(3, 4, 1, (257, 188), 224)
    losses = mla_b200.train_epoch(_args(), 0, model, torch.device("cuda"), batches, opt, sch, gs_plugin=gs,
                                  gs_flag=True, av_alpha=0.55)
    t_losses, _ = _torch_tf32_epoch(state, batches, False)
    _calibrated(losses, g["full_noop_losses"], t_losses, "KAT-6 3-step losses")
    assert np.allclose(losses, g["full_noop_losses"], rtol=5e-3)          # and never further than 0.5 %


def test_step_vs_oracle_with_projection(built_lib):
    """Two steps with the projection firing, against the oracle run (a) in fp32 on the CPU and (b) on this
    GPU under torch's TF32 default; same calibrated criterion, and the encoder weights move the same way."""
    import mla_b200
    model, state = _model(built_lib)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin(force_projection=True)
    batches = _batches(2, 8, 5, (97, 64), 96)
    losses = mla_b200.train_epoch(_args(), 0, model, torch.device("cuda"), batches, opt, sch, gs_plugin=gs,
                                  gs_flag=True, av_alpha=0.55)
    o = orc.AVOracle(state, force_projection=True)
    ref = o.train_epoch([b[:3] for b in batches], av_alpha=0.55)
    t_losses, t_orc = _torch_tf32_epoch(state, batches, True)
    _calibrated(losses, ref, t_losses, "2-step losses with projection")
    # the stem weights' UPDATE (w - w0) points the same way as the fp32 oracle's
    w0 = state["audio_net.conv1.weight"].double().flatten()
    dw = model.module.audio_net.conv1.weight.detach().cpu().double().flatten() - w0
    dr = o.sd["audio_net.conv1.weight"].detach().double().flatten() - w0
    dt = t_orc.sd["audio_net.conv1.weight"].detach().cpu().double().flatten() - w0
    cos = float(torch.nn.functional.cosine_similarity(dw, dr, dim=0))
    cos_t = float(torch.nn.functional.cosine_similarity(dt, dr, dim=0))
    print("stem update cosine vs fp32 oracle: ours %.4f torch-TF32 %.4f" % (cos, cos_t))
    assert cos > 0.95 and cos >= cos_t - 0.03
