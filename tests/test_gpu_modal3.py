"""--lorb m3ae --modal3 --gs_flag (BASELINE.json configs[3]) through the public API against
fixtures produced by the reference's own CAVMAEFT / Modal3Classifier / train_epoch / valid (tests/golden/modal3.npz; timm
0.4.5's Attention / Mlp restated in the generator — those two classes are unpinned) and against the oracle.
Tolerance: rel 1e-3 (fp32/TF32 of north_star), as for the m3ae path (tests/test_gpu_m3ae.py)."""
import argparse

import numpy as np
import pytest
import torch

from oracle import mla_oracle as orc

pytestmark = pytest.mark.gpu
TINY = dict(model_type=None, emb_dim=64, depth=2, num_heads=2)
CAV_TINY = dict(img_size=32, audio_length=64, embed_dim=64, modality_specific_depth=1, num_heads=2)


def relf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _args(dynamic=True):
    return argparse.Namespace(dataset="IEMOCAP", fusion_method="concat", modulation="Normal", gs_flag=True, dynamic=dynamic,
                              lorb="m3ae", modal3=True, clip=False)


def _state(g):
    return {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}


def _batches(n, B, seed, L=12, img=32, T=64, n_classes=4, vocab=512):
    gen = torch.Generator().manual_seed(seed)
    res = []
    for _ in range(n):
        token = torch.randint(0, vocab, (B, 1, L), generator=gen)
        n_valid = torch.randint(3, L + 1, (B,), generator=gen)
        pm = (torch.arange(L)[None, :] >= n_valid[:, None]).long()[:, None, :]
        image = torch.randn(B, 3, img, img, generator=gen)
        spec = torch.randn(B, T, 128, generator=gen)
        label = torch.randint(0, n_classes, (B,), generator=gen)
        res.append((token, pm, image, spec, label, torch.zeros(B, 1, dtype=torch.long)))
    return res


def _tiny_model(built_lib, g):
    import mla_b200
    net = mla_b200.Modal3Classifier(_args(), model_config=TINY, text_vocab_size=512, audio_kwargs=CAV_TINY)
    net.load_state_dict(_state(g), strict=True)
    return mla_b200.ModuleHolder(net.cuda())


def test_forward_and_gradients_match_reference_fixture(built_lib, golden):
    g = golden("modal3")
    model = _tiny_model(built_lib, g)
    (token, pm, image, spec, _, _), = _batches(1, 4, 41)
    a, v, t = model(token.cuda(), pm.cuda(), image.cuda(), spec.cuda())
    e = [relf(x.detach().cpu(), g[k]) for x, k in ((a, "fwd_a"), (v, "fwd_v"), (t, "fwd_t"))]
    print("modal3 feature rel-F error vs the reference fixture (a, v, t):", e)
    assert a.shape == (4, 64) and max(e) < 1e-3
    a.square().sum().backward()
    params = dict(model.module.named_parameters())
    for k in g.files:
        if k.startswith("grad/"):
            err = relf(params[k[5:]].grad.cpu(), g[k])
            print("  grad %-45s rel-F %.2e" % (k[5:], err))
            assert err < 2e-3, k
    none = sorted(k for k, p in params.items() if k.startswith("mae_a.") and p.grad is None)
    assert none == sorted(g["grad_none"])


@pytest.mark.parametrize("steps", [1, 3])
def test_train_epoch_and_valid_match_reference_fixture(built_lib, golden, steps):
    import mla_b200
    g = golden("modal3")
    tag = "step%d_" % steps
    model = _tiny_model(built_lib, g)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin()
    batches = _batches(3, 8, 9)[:steps]
    dev = torch.device("cuda")
    losses = mla_b200.train_epoch(_args(), 0, model, dev, batches, opt, sch, gs_plugin=gs, gs_flag=True, av_alpha=0.55)
    print("modal3 %d-step losses" % steps, losses, "fixture", g[tag + "losses"])
    assert len(losses) == 4 and np.allclose(losses, g[tag + "losses"], rtol=1e-3)
    assert gs.exp_count == int(g[tag + "exp_count"]) == 3 * steps
    sd = model.module.state_dict()
    for name, key in (("fusion_module.fc_out.weight", "fc_w"), ("mae_a.blocks_a.0.attn.qkv.weight", "qkv_a"),
                      ("mae_a.patch_embed_a.proj.weight", "patch_a"), ("mae_t.encoder.blocks.1.transformer_mlp.fc2.weight", "fc2_t")):
        w0 = g["state/" + name].astype(np.float64)
        upd, ref = sd[name].cpu().numpy().astype(np.float64) - w0, g[tag + key].astype(np.float64) - w0
        cos = float((upd * ref).sum() / (np.linalg.norm(upd) * np.linalg.norm(ref)))
        e = relf(sd[name].cpu(), g[tag + key])
        print("  %-50s rel-F %.2e, update cosine %.5f" % (name, e, cos))
        assert e < 1e-3 and cos > 0.999
    # CAVMAEFT's visual branch is not on the path: untouched, exactly (the reference's optimiser skips grad None)
    assert torch.equal(sd["mae_a.blocks_v.0.attn.qkv.weight"].cpu(), torch.from_numpy(g[tag + "unused_v"]))
    accs = mla_b200.valid(_args(True), model, dev, batches, gs_flag=True, av_alpha=0.55)
    accs_fix = mla_b200.valid(_args(False), model, dev, batches, gs_flag=True, av_alpha=0.55)
    n = 8.0 * steps
    assert len(accs) == 4
    assert np.abs(np.array(accs) - g[tag + "accs_dyn"]).max() <= 1 / n + 1e-9
    assert np.abs(np.array(accs_fix) - g[tag + "accs_fix"]).max() <= 1 / n + 1e-9


def test_full_size_step_vs_oracle(built_lib):
    """IEMOCAP shapes (spectrogram 1024 x 128 -> 512 audio tokens, 256 text tokens, 256 x 256 image), 'base' encoders, B=2:
    features and one three-turn step against the oracle's fp32 restatement on this GPU."""
    import mla_b200
    mla_b200.setup_seed(0)
    net = mla_b200.Modal3Classifier(_args())
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    model = mla_b200.ModuleHolder(net.cuda())
    token, pm, image, _ = orc.synthetic_m3ae_batch(2, 5, n_classes=4)
    gen = torch.Generator().manual_seed(6)
    spec = torch.randn(2, 1024, 128, generator=gen)
    label = torch.randint(0, 4, (2,), generator=gen)
    prev = torch.backends.cuda.matmul.allow_tf32
    prev_c = torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        o = orc.Modal3Oracle({k: v.cuda() for k, v in state.items()}, num_heads=12)
        with torch.no_grad():
            ref = orc.modal3_forward(o.sd, token.cuda(), pm.cuda(), image.cuda(), spec.cuda(), 12)
        out = model(token.cuda(), pm.cuda(), image.cuda(), spec.cuda())
        e = [relf(x.detach().cpu(), r.cpu()) for x, r in zip(out, ref)]
        print("base modal3 feature rel-F error (a, v, t):", e)
        assert out[0].shape == (2, 768) and max(e) < 1e-3
        opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
        sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
        gs = mla_b200.GSPlugin()
        batch = [(token, pm, image, spec, label, torch.zeros(2, 1, dtype=torch.long))]
        losses = mla_b200.train_epoch(_args(), 0, model, torch.device("cuda"), batch, opt, sch, gs_plugin=gs, gs_flag=True,
                                      av_alpha=0.55)
        ref_l = o.train_epoch([(token.cuda(), pm.cuda(), image.cuda(), spec.cuda(), label.cuda())], av_alpha=0.55)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
        torch.backends.cudnn.allow_tf32 = prev_c
    print("base modal3 step losses", losses, "oracle(fp32, GPU)", ref_l)
    assert np.allclose(losses, ref_l, rtol=1e-3)
    for name in ("mae_a.blocks_a.10.mlp.fc1.weight", "mae_a.patch_embed_a.proj.weight", "mae_a.blocks_u.0.attn.qkv.weight"):
        w0 = state[name].double()
        upd = model.module.state_dict()[name].cpu().double() - w0
        rupd = o.sd[name].detach().cpu().double() - w0
        e = float((upd - rupd).norm() / rupd.norm())
        print("  %-45s weight-update rel-F error %.2e" % (name, e))
        assert e < 5e-3
