"""--lorb m3ae --gs_flag (BASELINE.json configs[2]) through the public API against the
reference fixtures (tests/golden/m3ae.npz, produced by executing the reference) and the oracle.

Tolerance: the encoder GEMMs multiply operands with a 10-bit mantissa (fp16 forward; fp16 backward with an exact
power-of-two gradient scale in the fused blocks, TF32 on the per-module path; fp32 accumulate) where the reference
multiplies in fp32: north_star's fp32/TF32 tolerance, rel 1e-3 (Frobenius), on features, losses and weights after a step;
gradients of a single Linear 1e-3 as well. The optional bf16 backward gets 4e-3 on gradients."""
import argparse

import numpy as np
import pytest
import torch

from oracle import mla_oracle as orc

pytestmark = pytest.mark.gpu
TINY = dict(model_type=None, emb_dim=64, depth=2, num_heads=2)


def relf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _args(dynamic=True):
    return argparse.Namespace(dataset="Food101", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=dynamic, lorb="m3ae", modal3=False, clip=False)


def _state(g):
    return {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}


def _batches(n, B, seed, L=12, img=32, n_classes=101, vocab=512):
    gen = torch.Generator().manual_seed(seed)
    res = []
    for _ in range(n):
        token = torch.randint(0, vocab, (B, 1, L), generator=gen)
        n_valid = torch.randint(3, L + 1, (B,), generator=gen)
        pm = (torch.arange(L)[None, :] >= n_valid[:, None]).long()[:, None, :]
        image = torch.randn(B, 3, img, img, generator=gen)
        label = torch.randint(0, n_classes, (B,), generator=gen)
        res.append((token, pm, image, label, torch.zeros(B, 1, dtype=torch.long)))
    return res


def _tiny_model(built_lib, g):
    import mla_b200
    net = mla_b200.M3AEClassifier(_args(), model_config=TINY, text_vocab_size=512)
    net.load_state_dict(_state(g), strict=True)
    return mla_b200.ModuleHolder(net.cuda())


@pytest.mark.parametrize("M,K,N", [(257 * 3, 768, 2304), (513 * 2, 3072, 768), (48, 64, 192), (4 * 513, 768, 768),
                                   (130, 768, 64), (1, 64, 64)])
@pytest.mark.parametrize("bf16_bwd", [False, True])
def test_native_linear_forward_backward(built_lib, M, K, N, bf16_bwd):
    """One Linear on the tcgen05 GEMM kernels (ragged row counts: the 128-row tiles and the 32/64-row k-blocks of the
    weight gradient end inside a tile) against torch fp64."""
    from mla_b200 import m3ae
    gen = torch.Generator().manual_seed(M + K + N)
    lin = m3ae.NativeLinear(K, N).cuda()
    x = torch.randn(M, K, generator=gen).cuda().requires_grad_(True)
    dy = torch.randn(M, N, generator=gen).cuda()
    prev = m3ae.BACKWARD_BF16
    m3ae.BACKWARD_BF16 = bf16_bwd
    try:
        y = lin(x)
        y.backward(dy)
    finally:
        m3ae.BACKWARD_BF16 = prev
    torch.cuda.synchronize()
    xd, wd, bd, dyd = x.detach().double(), lin.weight.detach().double(), lin.bias.detach().double(), dy.double()
    ry = xd @ wd.t() + bd
    tol = 4e-3 if bf16_bwd else 1e-3
    e = (relf(y.detach().cpu(), ry.cpu()), relf(x.grad.cpu(), (dyd @ wd).cpu()), relf(lin.weight.grad.cpu(), (dyd.t() @ xd).cpu()),
         relf(lin.bias.grad.cpu(), dyd.sum(0).cpu()))
    print("linear M=%d K=%d N=%d bf16_bwd=%s: rel-F y %.2e dx %.2e dW %.2e db %.2e" % ((M, K, N, bf16_bwd) + e))
    assert e[0] < 1e-3 and e[1] < tol and e[2] < tol and e[3] < 1e-5


def test_native_linear_has_no_cpu_path(built_lib):
    from mla_b200 import m3ae
    with pytest.raises(RuntimeError):
        m3ae.NativeLinear(64, 64)(torch.zeros(2, 64))


def test_forward_matches_reference_fixture(built_lib, golden):
    g = golden("m3ae")
    model = _tiny_model(built_lib, g)
    (token, pm, image, _, _), = _batches(1, 4, 31)
    a, v = model(token.cuda(), pm.cuda(), image.cuda())
    assert a.shape == (4, 64) and a.requires_grad and v.requires_grad
    ea, ev = relf(a.detach().cpu(), g["fwd_a"]), relf(v.detach().cpu(), g["fwd_v"])
    print("m3ae feature rel-F error vs the reference fixture:", ea, ev)
    assert ea < 1e-3 and ev < 1e-3
    # gradients of the fixture's scalar through the native backward
    (a.square().sum() + v.square().sum()).backward()
    params = dict(model.module.named_parameters())
    for k in g.files:
        if k.startswith("grad/"):
            e = relf(params[k[5:]].grad.cpu(), g[k])
            print("  grad %-60s rel-F %.2e" % (k[5:], e))
            assert e < 2e-3, k


@pytest.mark.parametrize("steps", [1, 3])
def test_train_epoch_and_valid_match_reference_fixture(built_lib, golden, steps):
    import mla_b200
    g = golden("m3ae")
    tag = "step%d_" % steps
    model = _tiny_model(built_lib, g)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin()
    batches = _batches(3, 8, 7)[:steps]
    dev = torch.device("cuda")
    losses = mla_b200.train_epoch(_args(), 0, model, dev, batches, opt, sch, gs_plugin=gs, gs_flag=True, av_alpha=0.55)
    print("m3ae %d-step losses" % steps, losses, "fixture", g[tag + "losses"])
    assert np.allclose(losses, g[tag + "losses"], rtol=1e-3)
    assert gs.exp_count == int(g[tag + "exp_count"])
    sd = model.module.state_dict()
    for name, key in (("fusion_module.fc_out.weight", "fc_w"), ("mae_a.encoder.blocks.0.attention.qkv_linear.weight", "qkv0_a"),
                      ("mae_v.encoder.blocks.1.transformer_mlp.fc2.weight", "fc2_v")):
        w0 = g["state/" + name].astype(np.float64)
        upd, ref = sd[name].cpu().numpy().astype(np.float64) - w0, g[tag + key].astype(np.float64) - w0
        e = relf(sd[name].cpu(), g[tag + key])
        cos = float((upd * ref).sum() / (np.linalg.norm(upd) * np.linalg.norm(ref)))
        print("  %-55s rel-F %.2e, update cosine %.5f, |update| ratio %.4f" % (name, e, cos,
                                                                               np.linalg.norm(upd) / np.linalg.norm(ref)))
        assert e < 1e-3 and cos > 0.999
    # parameters the forward never reads get no gradient, so SGD (weight decay, momentum) must leave them bit-unchanged
    for name in ("mae_v.text_embedding.weight", "mae_v.encoder_text_type_embedding", "mae_a.image_embedding.weight",
                 "mae_a.image_embedding.bias", "mae_a.encoder_image_type_embedding"):
        assert np.array_equal(sd[name].cpu().numpy(), g["state/" + name]), name
    accs = mla_b200.valid(_args(True), model, dev, batches, gs_flag=True, av_alpha=0.55)
    accs_fix = mla_b200.valid(_args(False), model, dev, batches, gs_flag=True, av_alpha=0.55)
    n = 8.0 * steps
    assert np.abs(np.array(accs) - g[tag + "accs_dyn"]).max() <= 1 / n + 1e-9
    assert np.abs(np.array(accs_fix) - g[tag + "accs_fix"]).max() <= 1 / n + 1e-9


def test_step_with_projection_vs_oracle(built_lib, golden):
    """Two steps with the GS projection firing ("what the paper meant") against the fp32 CPU oracle."""
    import mla_b200
    g = golden("m3ae")
    model = _tiny_model(built_lib, g)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
    gs = mla_b200.GSPlugin(force_projection=True)
    batches = _batches(2, 8, 19)
    losses = mla_b200.train_epoch(_args(), 0, model, torch.device("cuda"), batches, opt, sch, gs_plugin=gs, gs_flag=True,
                                  av_alpha=0.55)
    o = orc.M3AEOracle(_state(g), num_heads=2, force_projection=True)
    ref = o.train_epoch([b[:4] for b in batches], av_alpha=0.55)
    print("m3ae 2-step losses with projection", losses, "oracle", ref)
    assert np.allclose(losses, ref, rtol=1e-3)
    # P itself is ill-conditioned on SIGNED (LayerNorm) features: the elementwise denominator alpha + k_i r_j
    # crosses zero, so fp32 runs of the same update diverge from each other (the kernel-level GS tests carry
    # the calibrated criterion). Here: P stays normalised and finite, and the projected head update agrees.
    e_p = relf(gs.Pl.cpu(), o.Pl)
    e_w = relf(model.module.fusion_module.fc_out.weight.detach().cpu(), o.sd["fusion_module.fc_out.weight"].detach())
    print("  P rel-F deviation %.2e (ill-conditioned, informational), |P|_F %.6f, head weight rel-F %.2e" % (
        e_p, float(gs.Pl.norm()), e_w))
    assert bool(torch.isfinite(gs.Pl).all()) and abs(float(gs.Pl.norm()) - 1) < 1e-4
    assert not bool(torch.equal(gs.Pl, torch.eye(64, device="cuda")))
    assert e_w < 1e-3


def test_base_encoder_full_size_vs_oracle(built_lib):
    """The 'base' encoders (768 wide, 12 blocks, 12 heads) at the Food-101 shapes (256 tokens, 256x256 image, B=2): features
    and one training step against the oracle's restatement run in fp32 on this GPU (TF32 off)."""
    import mla_b200
    mla_b200.setup_seed(0)
    net = mla_b200.M3AEClassifier(_args())
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    model = mla_b200.ModuleHolder(net.cuda())
    token, pm, image, label = orc.synthetic_m3ae_batch(2, 5)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        o = orc.M3AEOracle({k: v.cuda() for k, v in state.items()}, num_heads=12)
        with torch.no_grad():
            ra, rv = orc.m3ae_forward(o.sd, token.cuda(), pm.cuda(), image.cuda(), 12)
        a, v = model(token.cuda(), pm.cuda(), image.cuda())
        ea, ev = relf(a.detach().cpu(), ra.cpu()), relf(v.detach().cpu(), rv.cpu())
        print("base m3ae feature rel-F error:", ea, ev)
        assert a.shape == (2, 768) and ea < 1e-3 and ev < 1e-3
        opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
        sch = torch.optim.lr_scheduler.StepLR(opt, 70, 0.1)
        gs = mla_b200.GSPlugin()
        batch = [(token, pm, image, label, torch.zeros(2, 1, dtype=torch.long))]
        losses = mla_b200.train_epoch(_args(), 0, model, torch.device("cuda"), batch, opt, sch, gs_plugin=gs, gs_flag=True,
                                      av_alpha=0.55)
        ref = o.train_epoch([(token.cuda(), pm.cuda(), image.cuda(), label.cuda())], av_alpha=0.55)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    print("base m3ae step losses", losses, "oracle(fp32, GPU)", ref)
    assert np.allclose(losses, ref, rtol=1e-3)
    for name in ("mae_a.encoder.blocks.11.transformer_mlp.fc1.weight", "mae_v.image_embedding.weight",
                 "mae_a.encoder.blocks.0.attention.qkv_linear.weight"):
        w0 = state[name].double()
        upd = model.module.state_dict()[name].cpu().double() - w0
        rupd = o.sd[name].detach().cpu().double() - w0
        e = float((upd - rupd).norm() / rupd.norm())
        print("  %-55s weight-update rel-F error %.2e" % (name, e))
        assert e < 5e-3


def _attention_ref(qkv, mask, H, scale):
    """The reference's attention core (m3ae.py:103-121) in fp64."""
    B, S, C3 = qkv.shape
    Dh = C3 // (3 * H)
    q, k, v = qkv.view(B, S, 3, H, Dh).permute(2, 0, 3, 1, 4)
    att = torch.matmul(q, k.transpose(-2, -1)) * scale
    if mask is not None:
        att = torch.where(mask[:, None, None, :].expand(att.shape) > 0, torch.tensor(-1e7, dtype=att.dtype, device=att.device),
                          att)
    att = torch.softmax(att, dim=-1)
    return torch.matmul(att, v).permute(0, 2, 1, 3).reshape(B, S, H * Dh)


@pytest.mark.parametrize("B,S,H,Dh,masked", [(2, 513, 12, 64, True), (3, 257, 12, 64, False), (4, 13, 2, 32, True),
                                             (1, 64, 1, 64, True), (2, 65, 3, 32, False), (1, 1, 2, 64, False),
                                             (2, 130, 4, 64, "all"), (2, 71, 2, 64, True), (1, 527, 3, 64, True),
                                             (3, 16, 2, 64, False), (2, 300, 5, 64, True), (2, 15, 1, 64, "all"),
                                             (1, 384, 2, 64, False)])
def test_fused_attention_forward_backward(built_lib, B, S, H, Dh, masked):
    """csrc/attention.cu / attention_tc.cu (tcgen05 forward for head width 64: keys in multiples of 16 on the tensor cores, the
    remainder on the CUDA cores, 1 or 2 UMMA column blocks) against the reference formulation in fp64: ragged sequence lengths, padded keys
    (filled with -1e7, no gradient to their scores), a batch row whose keys are ALL padded (uniform weights, like the
    reference), both head widths."""
    from mla_b200 import m3ae
    gen = torch.Generator().manual_seed(B * 1000 + S)
    qkv = (torch.randn(B, S, 3 * H * Dh, generator=gen) * 0.8).cuda().requires_grad_(True)
    dout = torch.randn(B, S, H * Dh, generator=gen).cuda()
    mask = None
    if masked:
        n_valid = torch.randint(1, S + 1, (B,), generator=gen)
        if masked == "all":
            n_valid[0] = 0
        mask = (torch.arange(S)[None, :] >= n_valid[:, None]).float().cuda()
    scale = Dh ** -0.5
    out = m3ae._AttentionFn.apply(qkv, mask, H, scale)
    out.backward(dout)
    torch.cuda.synchronize()
    q64 = qkv.detach().double().requires_grad_(True)
    ref = _attention_ref(q64, mask.double() if mask is not None else None, H, scale)
    ref.backward(dout.double())
    e_o, e_g = relf(out.detach().cpu(), ref.detach().cpu()), relf(qkv.grad.cpu(), q64.grad.cpu())
    g3 = qkv.grad.view(B, S, 3, H * Dh).cpu().double()
    r3 = q64.grad.view(B, S, 3, H * Dh).cpu()
    # per-part error relative to the largest part (dq = dk = 0 exactly when there is a single key)
    den = max(float(r3[:, :, i].norm()) for i in range(3))
    parts = [float((g3[:, :, i] - r3[:, :, i]).norm()) / den for i in range(3)]
    print("attention B=%d S=%d H=%d Dh=%d masked=%s: rel-F out %.2e dqkv %.2e (dq %.2e dk %.2e dv %.2e)" % (
        B, S, H, Dh, masked, e_o, e_g, parts[0], parts[1], parts[2]))
    assert e_o < 1e-3 and e_g < 1e-3 and max(parts) < 2e-3
    assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(qkv.grad).all())


@pytest.mark.parametrize("M,D", [(1, 64), (37, 384), (1000, 768), (130, 1024), (9, 1280), (513, 32)])
def test_layernorm_kernels(built_lib, M, D):
    """csrc/transformer_ops.cu LayerNorm forward (fp32 / fp16 / TF32-rounded outputs) and backward (dx with the residual
    add, dgamma, dbeta) against torch in fp64."""
    from mla_b200 import m3ae
    gen = torch.Generator().manual_seed(M * 7 + D)
    x = (torch.randn(M, D, generator=gen) * 1.7 + 0.3).cuda()
    w = (torch.rand(D, generator=gen) + 0.5).cuda()
    b = torch.randn(D, generator=gen).cuda()
    dy = torch.randn(M, D, generator=gen).cuda()
    resid = torch.randn(M, D, generator=gen).cuda()
    y, _, _, mean, rstd = m3ae._ln_fwd(x, w, b, 1e-5, want_y=True)
    _, y16, y_r, _, _ = m3ae._ln_fwd(x, w, b, 1e-5)
    dx, dw, db = m3ae._ln_bwd(dy, x, mean, rstd, w, resid)
    dx0, _, _ = m3ae._ln_bwd(dy, x, mean, rstd, w, None)
    torch.cuda.synchronize()
    x64 = x.double().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(x64, (D,), w64, b64, 1e-5)
    ref.backward(dy.double())
    assert relf(y.cpu(), ref.detach().cpu()) < 1e-6
    assert relf(y16.float().cpu(), ref.detach().cpu()) < 6e-4 and relf(y_r.cpu(), ref.detach().cpu()) < 6e-4
    assert bool((y_r.view(torch.int32) & 0x1fff).eq(0).all())                   # TF32: low 13 mantissa bits clear
    assert relf(dx0.cpu(), x64.grad.cpu()) < 1e-5 and relf(dx.cpu(), (x64.grad + resid.double()).cpu()) < 1e-5
    assert relf(dw.cpu(), w64.grad.cpu()) < 1e-5 and relf(db.cpu(), b64.grad.cpu()) < 1e-5


@pytest.mark.parametrize("M,N", [(1, 64), (771, 768), (1026, 3072), (50, 2304), (17, 192)])
def test_elementwise_operand_kernels(built_lib, M, N):
    """cast_round (with and without GELU), round_colsum (with and without the GELU derivative)."""
    from mla_b200 import m3ae
    gen = torch.Generator().manual_seed(M + N)
    u = (torch.randn(M, N, generator=gen) * 1.5).cuda()
    dy = torch.randn(M, N, generator=gen).cuda()
    u16, u_r = m3ae._cast_round(u, False)
    g16, g_r = m3ae._cast_round(u, True)
    d_r, col = m3ae._round_colsum(dy, None)
    du_r, colu = m3ae._round_colsum(dy, u)
    torch.cuda.synchronize()
    u64 = u.double().requires_grad_(True)
    g = torch.nn.functional.gelu(u64)
    g.backward(dy.double())
    assert torch.equal(u16, u.half()) and relf(u_r.cpu(), u.cpu()) < 4e-4
    assert relf(g16.float().cpu(), g.detach().cpu()) < 4e-4 and relf(g_r.cpu(), g.detach().cpu()) < 4e-4
    assert relf(d_r.cpu(), dy.cpu()) < 4e-4 and relf(col.cpu(), dy.double().sum(0).cpu()) < 1e-5
    assert relf(du_r.cpu(), u64.grad.cpu()) < 4e-4 and relf(colu.cpu(), u64.grad.sum(0).cpu()) < 1e-5
    for t in (u_r, g_r, d_r, du_r):
        assert bool((t.view(torch.int32) & 0x1fff).eq(0).all())


@pytest.mark.parametrize("M,K,N", [(771, 768, 2304), (130, 3072, 768), (5, 64, 64)])
def test_linear_epilogue_bias_and_residual(built_lib, M, K, N):
    from mla_b200 import m3ae
    gen = torch.Generator().manual_seed(M + K + N)
    x16 = torch.randn(M, K, generator=gen).cuda().half()
    w = (torch.randn(N, K, generator=gen) * 0.05).cuda()
    bias = torch.randn(N, generator=gen).cuda()
    resid = torch.randn(M, N, generator=gen).cuda()
    ref = x16.double() @ w.half().double().t()
    y0 = m3ae._linear16(x16, w, None, None, M, K, N)
    y1 = m3ae._linear16(x16, w, bias, None, M, K, N)
    y2 = m3ae._linear16(x16, w, bias, resid, M, K, N)
    torch.cuda.synchronize()
    assert relf(y0.cpu(), ref.cpu()) < 1e-5
    assert relf(y1.cpu(), (ref + bias.double()).cpu()) < 1e-5
    assert relf(y2.cpu(), (ref + bias.double() + resid.double()).cpu()) < 1e-5


@pytest.mark.parametrize("M,K,N,scale", [(771, 768, 2304, 3e-6), (1026, 3072, 768, 40.0), (48, 64, 192, 1e-9), (130, 768, 64, 1.0)])
def test_fp16_scaled_linear_backward(built_lib, M, K, N, scale):
    """mla_grad_operand16 + mla_linear_dgrad16 / mla_linear_wgrad16: gradients of any magnitude (here 1e-9 .. 40, with
    three decades of spread inside the tensor) through fp16 operands and an exact power-of-two scale, against fp64."""
    from mla_b200 import m3ae
    gen = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=gen).cuda()
    w = (torch.randn(N, K, generator=gen) * 0.05).cuda()
    dy = (torch.randn(M, N, generator=gen) * scale * torch.logspace(-3, 0, M, dtype=torch.float32)[:, None]).cuda()
    u = (torch.randn(M, N, generator=gen) * 1.5).cuda()
    x16 = x.half()
    for uu in (None, u):
        d16, col, inv = m3ae._grad_operand16(dy, uu)
        dx, dw = m3ae._linear_grads16(x16, d16, inv, w, M, K, N)
        torch.cuda.synchronize()
        v = dy.double()
        if uu is not None:
            u64 = uu.double().requires_grad_(True)
            torch.nn.functional.gelu(u64).backward(dy.double())
            v = u64.grad
        F = 1.0 / float(inv[0])
        amax = float(v.abs().max())
        assert 2048 <= F * amax < 4096.001 and F == 2.0 ** round(np.log2(F))
        e = (relf(d16.float().cpu() / F, v.cpu()), relf(col.cpu(), v.sum(0).cpu()), relf(dx.cpu(), (v @ w.double()).cpu()),
             relf(dw.cpu(), (v.t() @ x16.double()).cpu()))
        print("fp16-scaled backward M=%d K=%d N=%d |dy|~%.0e gelu=%s: operand %.2e colsum %.2e dx %.2e dW %.2e" % (
            (M, K, N, scale, uu is not None) + e))
        assert e[0] < 4e-4 and e[1] < 1e-5 and e[2] < 1e-3 and e[3] < 1e-3


@pytest.mark.parametrize("bwd", ["fp16", "tf32"])
@pytest.mark.parametrize("B,S,D,H,masked", [(2, 513, 768, 12, True), (3, 17, 64, 2, False)])
def test_fused_block_matches_per_module_path(built_lib, B, S, D, H, masked, bwd):
    """_BlockFn (one node, hand-written backward) against the per-module autograd path of the same Block and against the
    oracle's fp32 restatement of the block on this GPU: output, input gradient and every parameter gradient."""
    from mla_b200 import m3ae
    torch.manual_seed(B + S + D)
    blk = m3ae.Block(D, H).cuda()
    for p_ in blk.parameters():                    # non-trivial LayerNorm affine / biases
        if p_.dim() == 1:
            p_.data.add_(torch.randn_like(p_) * 0.1)
    x = torch.randn(B, S, D, device="cuda")
    dy = torch.randn(B, S, D, device="cuda")
    mask = None
    if masked:
        n_valid = torch.randint(1, S + 1, (B,))
        mask = (torch.arange(S)[None, :] >= n_valid[:, None]).float().cuda()

    def run(fused):
        prev = m3ae.FUSED_BLOCK, m3ae.BLOCK_BACKWARD
        m3ae.FUSED_BLOCK, m3ae.BLOCK_BACKWARD = fused, bwd
        try:
            blk.zero_grad()
            xi = x.clone().requires_grad_(True)
            y = blk(xi, mask)
            y.backward(dy)
            torch.cuda.synchronize()
            return y.detach(), xi.grad.clone(), {k: v.grad.clone() for k, v in blk.named_parameters()}
        finally:
            m3ae.FUSED_BLOCK, m3ae.BLOCK_BACKWARD = prev

    yf, dxf, gf = run(True)
    ym, dxm, gm = run(False)
    # fp32 reference: the oracle's block arithmetic via torch ops (matmul TF32 off)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        import torch.nn.functional as F
        ps = {k: v.detach().clone().requires_grad_(True) for k, v in blk.named_parameters()}
        xi = x.clone().requires_grad_(True)
        h = F.layer_norm(xi, (D,), ps["layer_norm1.weight"], ps["layer_norm1.bias"])
        qkv = F.linear(h, ps["attention.qkv_linear.weight"], ps["attention.qkv_linear.bias"])
        a = _attention_ref(qkv, mask, H, (D // H) ** -0.5)
        x1 = xi + F.linear(a, ps["attention.fc.weight"], ps["attention.fc.bias"])
        h = F.layer_norm(x1, (D,), ps["layer_norm2.weight"], ps["layer_norm2.bias"])
        h = F.gelu(F.linear(h, ps["transformer_mlp.fc1.weight"], ps["transformer_mlp.fc1.bias"]))
        yr = x1 + F.linear(h, ps["transformer_mlp.fc2.weight"], ps["transformer_mlp.fc2.bias"])
        yr.backward(dy)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    e = {"y": (relf(yf.cpu(), yr.detach().cpu()), relf(ym.cpu(), yr.detach().cpu())),
         "dx": (relf(dxf.cpu(), xi.grad.cpu()), relf(dxm.cpu(), xi.grad.cpu()))}
    for k in gf:
        e[k] = (relf(gf[k].cpu(), ps[k].grad.cpu()), relf(gm[k].cpu(), ps[k].grad.cpu()))
    for k, (a_, b_) in e.items():
        print("  %-32s fused %.2e  per-module %.2e" % (k, a_, b_))
    assert max(v[0] for v in e.values()) < 1e-3 and max(v[1] for v in e.values()) < 1e-3


def test_encoder_head_width_64_matches_reference_fixture(built_lib, golden):
    """tests/golden/m3ae_dh64.npz: the reference's own encoder at head width 64 (emb 128, 2 heads, 71 text tokens = two key
    tiles, a sample with one unpadded key besides CLS; 10 image tokens). Exercises every Linear dispatch: persistent
    (N = 512), CTA-pair 128-column tiles (N = 384, 128) and the plain kernel (M = 30 < 128 rows)."""
    from mla_b200 import m3ae
    g = golden("m3ae_dh64")
    enc = m3ae.MaskedMultimodalAutoencoder(64, dict(model_type=None, emb_dim=128, depth=1, num_heads=2))
    enc.load_state_dict({k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}, strict=True)
    enc = enc.cuda()
    gen = torch.Generator().manual_seed(17)
    text = torch.randint(0, 64, (3, 70), generator=gen)
    pm = (torch.arange(70)[None, :] >= torch.tensor([70, 33, 1])[:, None]).long()
    image = torch.randn(3, 9, 768, generator=gen)
    t = enc.forward_representation(None, text.cuda(), pm.cuda())
    v = enc.forward_representation(image.cuda(), None, None)
    et, ev = relf(t.detach().cpu(), g["rep_text"]), relf(v.detach().cpu(), g["rep_image"])
    print("head-width-64 encoder rel-F error vs the reference fixture: text %.2e image %.2e" % (et, ev))
    assert t.shape == (3, 71, 128) and et < 1e-3 and ev < 1e-3
    wt, wv = torch.randn(t.shape, generator=gen), torch.randn(v.shape, generator=gen)
    ((t * wt.cuda()).sum() + (v * wv.cuda()).sum()).backward()
    params = dict(enc.named_parameters())
    for k in g.files:
        if k.startswith("grad/"):
            e = relf(params[k[5:]].grad.cpu(), g[k])
            print("  grad %-50s rel-F %.2e" % (k[5:], e))
            assert e < 2e-3, k
