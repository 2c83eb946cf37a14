"""Parity of the CUDA kernels (through the C ABI) with the CPU oracle and with the fixtures
produced by executing the reference. Tolerances (floating point, stated per test):
  GS projection  Frobenius-relative error vs the fp64 run of the reference code must be within
                 4x the reference's own fp32 error (+1e-6) — SURVEY F10 (the recursion is chaotic,
                 so parity is single-step / teacher-forced);
  head           rel 1e-5 vs the fp64 oracle;
  fusion         weights |dw| <= 2e-6 + 2.5e-8 max|H| (see w_tol), fused logits 1e-5, argmax and
                 counters bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import mla_oracle as orc

pytestmark = pytest.mark.gpu


def relf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.fixture(scope="module")
def ops(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from mla_b200 import ops as o
    return o


def w_tol(H):
    """Tolerance on the fusion weights w = softmax(-H): H is an fp32 sum of B*C terms (|H| up to
    C ln B, hundreds), so a few ulps of H (which torch's own unspecified reduction order also
    moves) shift w by w(1-w) dH. Bit-equality with torch is not attainable; this bound is ~0.4 ulp of H."""
    return 2e-6 + 6e-8 * float(np.max(np.abs(H)))


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to("cuda", dtype)


# ------------------------------------------------------------------------------------ GS
@pytest.mark.parametrize("name,steps", [("signed_d64", 3), ("relu_d64", 3), ("signed_d128", 1)])
def test_gs_teacher_forced_vs_reference_fixture(ops, golden, name, steps):
    g = golden("gs_plugin")
    for s in range(1, steps + 1):
        k = "%s_s%d_" % (name, s)
        bi, L, counter, _ = g[k + "meta"]
        P, grad = dev(g[k + "P_in"]), dev(g[k + "grad_in"])
        ops.gs_project(P, grad, orc.gs_alpha(int(bi), int(L)), feat=dev(g[k + "feat"]))
        ref_err = relf(g[k + "P_out32"], g[k + "P_out64"])
        assert relf(P.cpu().numpy(), g[k + "P_out64"]) <= 4 * ref_err + 1e-6
        ref_gerr = relf(g[k + "grad_out32"], g[k + "grad_out64"])
        assert relf(grad.cpu().numpy(), g[k + "grad_out64"]) <= 4 * ref_gerr + 1e-6
        assert abs(float(P.norm()) - 1) < 2e-6


def test_gs_d512_from_identity_fixture(ops, golden):
    g = golden("gs_plugin")
    k = "relu_d512_s1_"
    P, grad = torch.eye(512, device="cuda"), dev(g[k + "grad_in"])
    ops.gs_project(P, grad, orc.gs_alpha(1, 7), feat=dev(g[k + "feat"]))
    rows = g[k + "rows"]
    Pn = P.cpu().numpy()
    assert relf(Pn[rows], g[k + "P_out64"]) <= 4 * relf(g[k + "P_out32"], g[k + "P_out64"]) + 1e-6
    assert np.allclose(Pn, Pn.T, atol=1e-7)                       # symmetric after the first update from I
    assert relf(grad.cpu().numpy(), g[k + "grad_out64"]) < 1e-5
    assert abs(np.linalg.norm(Pn.astype(np.float64)) - 1) < 2e-6  # ||P||_F = 1


GS_SHAPES = [(64, 512, 6), (64, 768, 101), (256, 1024, 6), (1, 64, 1), (7, 68, 3), (64, 2048, 101), (300, 2048, 6)]


@pytest.mark.parametrize("B,D,C", GS_SHAPES)
@pytest.mark.parametrize("mode", [0, 1])
def test_gs_vs_oracle_seeded(ops, B, D, C, mode):
    rng = np.random.default_rng(B * 7 + D + C)
    feat = rng.standard_normal((B, D)).astype(np.float32)
    grad = rng.standard_normal((C, D)).astype(np.float32)
    # P after one warm update from I (SURVEY §8d cfg 5), computed by the oracle
    P0, _ = orc.gs_before_update(np.eye(D, dtype=np.float32), rng.standard_normal((B, D)).astype(np.float32), None,
                                 0, 5, 1, mode=mode)
    P64, g64 = orc.gs_before_update(P0, feat, grad, 2, 5, 3, dtype=np.float64, mode=mode)
    P32, g32 = orc.gs_before_update(P0, feat, grad, 2, 5, 3, dtype=np.float32, mode=mode)
    P, gw = dev(P0), dev(grad)
    ops.gs_project(P, gw, orc.gs_alpha(2, 5), feat=dev(feat), mode=mode)
    assert relf(P.cpu().numpy(), P64) <= 4 * relf(P32, P64) + 1e-6
    assert relf(gw.cpu().numpy(), g64) <= 4 * relf(g32, g64) + 2e-6
    # the pre-reduced (data-parallel) form gives the same result as the raw-feature form
    P2, gw2 = dev(P0), dev(grad)
    ops.gs_project(P2, gw2, orc.gs_alpha(2, 5), feat_sum=dev(feat.sum(0, dtype=np.float32)), inv_batch=1.0 / B,
                   mode=mode)
    assert relf(P2.cpu().numpy(), P64) <= 4 * relf(P32, P64) + 1e-6
    # deterministic: a second launch on the same inputs gives the same bits
    P3, gw3 = dev(P0), dev(grad)
    ops.gs_project(P3, gw3, orc.gs_alpha(2, 5), feat=dev(feat), mode=mode)
    assert torch.equal(P3, P) and torch.equal(gw3, gw)


def test_gs_without_gradient_and_errors(ops):
    P = torch.eye(128, device="cuda")
    ops.gs_project(P, None, 0.05, feat=torch.randn(8, 128, device="cuda"))
    assert abs(float(P.norm()) - 1) < 2e-6
    with pytest.raises(RuntimeError):
        ops.gs_project(P, None, 0.05, feat=torch.randn(8, 64, device="cuda"))          # width mismatch
    with pytest.raises(RuntimeError):
        ops.gs_project(torch.eye(4096, device="cuda"), None, 0.05, feat=torch.randn(8, 4096, device="cuda"))


def test_gsplugin_class_behaviour_on_gpu(golden):
    """KAT-1/KAT-2 through the reference-named class: bare Linear no-op, wrapped Linear fires,
    bias grad untouched, counter 0 skipped, Pl sized from the head (F4)."""
    import mla_b200
    g = golden("gs_plugin")
    k = "signed_d128_s1_"
    fc = torch.nn.Linear(128, 11).cuda()
    fc.weight.grad = dev(g[k + "grad_in"])
    fc.bias.grad = torch.ones(11, device="cuda")
    gs = mla_b200.GSPlugin()
    gs.before_update(fc, dev(g[k + "feat"]), 1, 7, 1)
    assert torch.equal(gs.Pl, torch.eye(512, device="cuda"))                           # published behaviour

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.module = m
    gs.before_update(Wrap(fc), dev(g[k + "feat"]), 1, 7, 0)
    assert torch.equal(gs.Pl, torch.eye(512, device="cuda"))                           # counter 0
    gs.before_update(Wrap(fc), dev(g[k + "feat"]), 1, 7, 1)
    assert gs.Pl.shape == (128, 128)
    assert relf(gs.Pl.cpu().numpy(), g[k + "P_out64"]) <= 4 * relf(g[k + "P_out32"], g[k + "P_out64"]) + 1e-6
    assert relf(fc.weight.grad.cpu().numpy(), g[k + "grad_out64"]) <= \
        4 * relf(g[k + "grad_out32"], g[k + "grad_out64"]) + 1e-6
    assert torch.equal(fc.bias.grad, torch.ones(11, device="cuda"))
    gs2 = mla_b200.GSPlugin(force_projection=True)
    fc.weight.grad = dev(g[k + "grad_in"])
    gs2.before_update(fc, dev(g[k + "feat"]), 1, 7, 1)
    assert torch.equal(gs2.Pl, gs.Pl)


# ---------------------------------------------------------------------------------- head
@pytest.mark.parametrize("name", ["b16d64c6", "b8d128c101", "b32d256c6"])
def test_head_vs_reference_fixture(ops, golden, name):
    g = golden("head")
    k = name + "_"
    o = ops.head_ce(dev(g[k + "feat"]), dev(g[k + "W"]), dev(g[k + "b"]), dev(g[k + "label"], torch.int64))
    for key in ("logits", "dW", "db", "dfeat"):
        assert relf(o[key].cpu().numpy(), g[k + key]) < 1e-5, key
    assert abs(float(o["loss"]) - float(g[k + "loss"])) < 1e-5
    assert relf(o["feat_sum"].cpu().numpy(), g[k + "feat"].astype(np.float64).sum(0)) < 1e-6


@pytest.mark.parametrize("B,D,C", [(64, 512, 6), (64, 768, 101), (64, 768, 4), (3, 2048, 33), (517, 512, 6), (1, 64, 1),
                                   # the tiled large-C path: tile edges (C = 17, 64, 65, 129), the logits GEMM split over D (small
                                   # batches), the weight-gradient GEMM split over the batch (B >= 256), narrow / ragged D
                                   (64, 768, 17), (130, 132, 65), (1, 2048, 64), (4096, 64, 129), (1000, 516, 101), (2, 4, 20),
                                   (333, 1024, 1000)])
def test_head_vs_oracle_seeded(ops, B, D, C):
    rng = np.random.default_rng(B + D + C)
    feat = np.maximum(rng.standard_normal((B, D)), 0).astype(np.float32)
    W = (rng.standard_normal((C, D)) * 0.05).astype(np.float32)
    b = (rng.standard_normal(C) * 0.1).astype(np.float32)
    lab = rng.integers(0, C, B)
    ref = orc.head_ce(feat, W, b, lab, grad_scale=1.0 / (2 * B))
    o = ops.head_ce(dev(feat), dev(W), dev(b), dev(lab, torch.int64), grad_scale=1.0 / (2 * B))
    for key in ("logits", "dW", "db", "dfeat", "feat_sum"):
        assert relf(o[key].cpu().numpy(), ref[key]) < 1e-5, key
    assert abs(float(o["loss"]) - ref["loss"]) < 1e-5 * max(1, abs(ref["loss"]))
    fwd = ops.head_ce(dev(feat), dev(W), dev(b), dev(lab, torch.int64), need_grad=False)
    assert torch.equal(fwd["logits"], o["logits"]) and fwd["dW"] is None
    assert relf(fwd["feat_sum"].cpu().numpy(), ref["feat_sum"]) < 1e-5      # the forward-only call still emits sum_b feat
    again = ops.head_ce(dev(feat), dev(W), dev(b), dev(lab, torch.int64), grad_scale=1.0 / (2 * B))
    assert all(torch.equal(again[k], o[k]) for k in ("logits", "dW", "db", "dfeat", "feat_sum", "loss"))      # deterministic


def test_head_out_of_range_label_poisons_the_loss(ops):
    """nn.CrossEntropyLoss raises for a label outside [0, C); the kernel must not read out of bounds and must
    not return a plausible number: the row's loss and gradient become NaN."""
    rng = np.random.default_rng(5)
    feat = rng.standard_normal((8, 64)).astype(np.float32)
    W = (rng.standard_normal((6, 64)) * 0.05).astype(np.float32)
    b = np.zeros(6, np.float32)
    for bad in (6, -1, 1 << 40):
        lab = rng.integers(0, 6, 8)
        lab[3] = bad
        o = ops.head_ce(dev(feat), dev(W), dev(b), dev(lab, torch.int64))
        assert np.isnan(float(o["loss"]))
        assert torch.isnan(o["dfeat"][3]).all() and not torch.isnan(o["dfeat"][2]).any()
    with pytest.raises(RuntimeError):
        ops.head_ce(dev(feat), dev(W), dev(b), dev(np.zeros(7, np.int64), torch.int64))


# -------------------------------------------------------------------------------- fusion
@pytest.mark.parametrize("name", ["b64c6m2", "b32c101m3", "b7c4m3", "b256c6m2"])
@pytest.mark.parametrize("x", [3, 10, 30])
def test_fusion_vs_reference_fixture(ops, golden, name, x):
    g = golden("fusion")
    k = "%s_x%d_" % (name, x)
    M = 3 if name.endswith("m3") else 2
    outs = [dev(g[k + "out%d" % m]) for m in range(M)]
    B, C = outs[0].shape
    lab = dev(g[k + "label"], torch.int64)
    hits = torch.zeros(M + 1, C, dtype=torch.int64, device="cuda")
    num = torch.zeros(C, dtype=torch.int64, device="cuda")
    fused, w, am, ent = ops.fuse_eval(outs, lab, hits=hits, num=num, want_entropy=True)
    assert np.allclose(ent.cpu().numpy(), g[k + "entropy"], rtol=3e-6)
    assert np.abs(w.cpu().numpy() - g[k + "w"]).max() <= w_tol(g[k + "entropy"])
    # fused = sum_m w_m out_m: a weight shift dw (bounded above) moves it by at most dw * sum_m max|out_m|
    amax = sum(float(np.abs(g[k + "out%d" % m]).max()) for m in range(M))
    assert np.allclose(fused.cpu().numpy(), g[k + "fused"], atol=1e-5 + w_tol(g[k + "entropy"]) * amax)
    assert np.array_equal(am.cpu().numpy(), g[k + "argmax"])                        # bit-exact predictions
    ref = orc.fuse_eval([g[k + "out%d" % m] for m in range(M)], g[k + "label"], C)
    assert np.array_equal(hits.cpu().numpy(), ref["hits"]) and np.array_equal(num.cpu().numpy(), ref["num"])
    # counters accumulate across batches (main.py:659-676)
    ops.fuse_eval(outs, lab, hits=hits, num=num)
    assert np.array_equal(hits.cpu().numpy(), 2 * ref["hits"])


def test_fusion_nan_and_fixed_weights(ops, golden):
    import mla_b200
    g = golden("fusion")
    o0, o1 = dev(g["nan_out0"]), dev(g["nan_out1"])
    w = mla_b200.calculate_gating_weights(o0, o1)                                      # KAT-3: NaN propagates
    assert w[0].dim() == 0 and bool(torch.isnan(w[0])) and bool(torch.isnan(w[1]))
    _, _, am = ops.fuse_eval([o0, o1])
    assert (am[0] == 0).all()
    k = "b64c6m2_x10_"
    a, v = dev(g[k + "out0"]), dev(g[k + "out1"])
    fused, w, am = ops.fuse_eval([a, v], dynamic=False, fixed_w=(0.55, 1 - 0.55))       # main.py:651
    ref = 0.55 * a + (1 - 0.55) * v
    assert torch.equal(fused, ref)                                                     # same roundings
    assert np.array_equal(am[0].cpu().numpy(), ref.argmax(1).cpu().numpy())
    h = mla_b200.calculate_entropy(a)
    assert abs(float(h) - float(g[k + "entropy"][0])) < 1e-4


@pytest.mark.parametrize("B,C,M", [(64, 6, 2), (257, 101, 3)])
def test_fusion_exact_guarantees(ops, B, C, M):
    """What IS bit-exact about the fusion kernel. The weights are fp32 functions of entropies that are
    sums of B*C terms: no two fp32 implementations (torch's own reduction order is unspecified) agree on them to the last
    bit, so they are held to w_tol() elsewhere. Exact guarantees, tested here bit for bit:
      1. determinism: the same inputs give the same weights, fused logits, predictions and counters, run after run;
      2. given the weights the kernel reports, the fused logits are EXACTLY the reference's expression
         (out_0 * w_0 + out_1 * w_1) + out_2 * w_2 with its fp32 roundings (main.py:643 / 646), and the predictions are
         exactly numpy's argmax of softmax of those logits (main.py:653-663: first maximum wins)."""
    rng = np.random.default_rng(B * 3 + C + M)
    outs = [dev(rng.standard_normal((B, C)).astype(np.float32) * 2) for _ in range(M)]
    lab = dev(rng.integers(0, C, B), torch.int64)
    runs = []
    for _ in range(3):
        hits = torch.zeros(M + 1, C, dtype=torch.int64, device="cuda")
        num = torch.zeros(C, dtype=torch.int64, device="cuda")
        fused, w, am = ops.fuse_eval(outs, lab, hits=hits, num=num)
        runs.append((fused.clone(), w.clone(), am.clone(), hits, num))
    for r in runs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(runs[0], r))
    fused, w, am = runs[0][:3]
    ref = outs[0] * w[0]
    for m in range(1, M):
        ref = ref + outs[m] * w[m]
    assert torch.equal(fused, ref)
    pred = np.argmax(torch.softmax(ref, dim=1).cpu().numpy(), axis=1)
    safe = np.ones(B, bool)            # softmax can merge two nearly equal logits into a tie that argmax(logits) does not see
    srt = np.sort(ref.cpu().numpy(), axis=1)
    if C > 1:
        safe = srt[:, -1] - srt[:, -2] > 1e-6
    assert np.array_equal(am[0].cpu().numpy()[safe], pred[safe])


@pytest.mark.parametrize("B,C,M", [(4096, 101, 3), (4096, 6, 2), (1000, 37, 2), (1, 6, 2), (33, 1, 3)])
def test_fusion_large_vs_oracle(ops, B, C, M):
    rng = np.random.default_rng(B + C + M)
    outs = [rng.standard_normal((B, C)).astype(np.float32) for _ in range(M)]
    lab = rng.integers(0, C, B)
    ref = orc.fuse_eval(outs, lab, C, dtype=np.float64)
    hits = torch.zeros(M + 1, C, dtype=torch.int64, device="cuda")
    num = torch.zeros(C, dtype=torch.int64, device="cuda")
    fused, w, am = ops.fuse_eval([dev(o) for o in outs], dev(lab, torch.int64), hits=hits, num=num)
    H = [orc.calculate_entropy(o, np.float64) for o in outs]
    assert np.abs(w.cpu().numpy() - ref["w"]).max() <= w_tol(H) and abs(float(w.sum()) - 1) < 1e-6
    assert np.array_equal(am.cpu().numpy()[1:], ref["argmax"][1:])
    # fused argmax: exact wherever the top-2 margin exceeds the weight tolerance
    f64 = ref["fused"]
    srt = np.sort(f64, axis=1)
    safe = (srt[:, -1] - srt[:, -2] > 1e-4) if C > 1 else np.ones(B, bool)
    assert np.array_equal(am.cpu().numpy()[0][safe], ref["argmax"][0][safe])
    assert int(num.sum()) == B and (hits.cpu().numpy() <= num.cpu().numpy()[None]).all()
