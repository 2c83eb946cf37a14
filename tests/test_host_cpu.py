"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared
symbol, argument errors are reported (not crashed on), the GSPlugin name/counter gates, the
fail-loudly rule, and the data-parallel plumbing under gloo with world_size 2."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    names = built_lib.declared_symbols()
    assert len(names) >= 10 and "mla_gs_project" in names and "mla_fuse_eval" in names
    raw = ctypes.CDLL(built_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libmla_b200.so does not export %s" % n
    assert built_lib.lib().mla_abi_version() == 1
    assert b"workspace" in built_lib.lib().mla_error_string(-3)


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_argument_errors_do_not_need_a_gpu(built_lib):
    L = built_lib.lib()
    # both feat and feat_sum NULL -> BADARG before any CUDA call
    assert L.mla_gs_project(None, None, None, 1.0, 0.1, None, 4, 64, 6, 0, None, 0, None) == -1
    assert L.mla_head_ce(None, None, None, None, 4, 64, 6, None, None, None, None, None, None, 1.0, None, 0, None) == -1
    assert L.mla_head_ce_workspace_bytes(0, 64, 6) == 0
    # transformer entry points: NULL / misaligned pointers and bad shapes are rejected before any CUDA call
    assert L.mla_attention_forward(None, None, None, None, None, 2, 16, 2, 64, 0.125, None) == -1
    assert L.mla_attention_backward(None, None, None, None, None, None, 2, 16, 2, 64, 0.125, None, 0, None) == -1
    assert L.mla_attention_backward_workspace_bytes(2, 16, 2, 48) == 0            # head width must be 32 or 64
    assert L.mla_attention_backward_workspace_bytes(2, 16, 2, 64) >= (2 * 2 * 16 + 4) * 4 + 2 * 16 * 2 * 64 * 2
    assert L.mla_linear_forward16(None, None, None, None, None, 128, 64, 64, None) == -1
    assert L.mla_linear_dgrad16(None, None, None, None, 128, 64, 64, None) == -1
    assert L.mla_linear_wgrad16(None, None, None, None, 128, 64, 64, None, 0, None) == -1
    assert L.mla_linear_dgrad(None, None, None, 128, 64, 64, None) == -1
    assert L.mla_layernorm_forward(None, None, None, 1e-5, 8, 64, None, None, None, None, None, None) == -1
    assert L.mla_layernorm_backward(None, None, None, None, None, None, 8, 64, None, None, None, None, 0, None) == -1
    assert L.mla_cast_round(None, None, None, 64, 0, None) == -1
    assert L.mla_round_colsum(None, None, None, None, 8, 64, None, 0, None) == -1
    assert L.mla_grad_operand16(None, None, None, None, None, 8, 64, None, 0, None) == -1


def test_transformer_host_mirrors_refuse_out_of_scope_configurations():
    import argparse
    import mla_b200
    from mla_b200.main import build_model
    ok = dict(dataset="Food101", fusion_method="concat", modulation="Normal", gs_flag=True, dynamic=True, lorb="m3ae",
              modal3=False, clip=False)
    for bad in (dict(fusion_method="sum"), dict(modulation="QMF"), dict(dataset="AVE")):
        with pytest.raises(NotImplementedError):
            mla_b200.M3AEClassifier(argparse.Namespace(**{**ok, **bad}), model_config=dict(model_type=None, emb_dim=64, depth=1,
                                                                                           num_heads=2), text_vocab_size=16)
    with pytest.raises(NotImplementedError):                                       # Modal3Classifier knows IEMOCAP only
        mla_b200.Modal3Classifier(argparse.Namespace(**{**ok, "modal3": True}))
    for bad in (dict(lorb="large"), dict(clip=True)):                               # CAVClassifier / CLIPClassifier
        with pytest.raises(NotImplementedError):
            build_model(argparse.Namespace(**{**ok, **bad, "ckpt_load_path_train": None}), "cpu")


def test_no_cpu_fallback(built_lib):
    from mla_b200 import ops
    P = torch.eye(64)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gs_project(P, torch.zeros(6, 64), 0.1, feat=torch.zeros(4, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.fuse_eval([torch.zeros(4, 6), torch.zeros(4, 6)])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal-learning-with-alternating-unimodal-adaptation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), "%s mentions the oracle" % f


def test_gsplugin_gates_match_reference():
    """KAT-1 / utils.py:29-32: bare Linear -> no-op; counter 0 -> no-op; alpha formula."""
    import mla_b200
    gs = mla_b200.GSPlugin(device="cpu")
    assert gs.exp_count == 0 and torch.equal(gs.Pl, torch.eye(512))
    fc = torch.nn.Linear(512, 6)
    fc.weight.grad = torch.randn(6, 512)
    g0 = fc.weight.grad.clone()
    gs.before_update(fc, torch.randn(8, 512), 3, 10, 5)                  # bare Linear: names are weight/bias
    assert torch.equal(fc.weight.grad, g0) and torch.equal(gs.Pl, torch.eye(512))

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.module = m
    gs.before_update(Wrap(fc), torch.randn(8, 512), 3, 10, 0)           # counter 0: skipped
    assert torch.equal(fc.weight.grad, g0)
    with pytest.raises(RuntimeError, match="CUDA"):                      # would fire -> needs the kernel
        gs.before_update(Wrap(fc), torch.randn(8, 512), 3, 10, 1)
    assert abs(mla_b200.GSPlugin.alpha(3, 10) - 0.1 ** 1.3) < 1e-15
    gs2 = mla_b200.GSPlugin(device="cpu", force_projection=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        gs2.before_update(fc, torch.randn(8, 512), 3, 10, 1)


def test_flag_surface_matches_reference_defaults():
    import mla_b200
    a = mla_b200.get_arguments(["--ckpt_path", "x"])
    assert (a.batch_size, a.learning_rate, a.lr_decay_step, a.lr_decay_ratio, a.random_seed) == (64, 1e-3, 70, 0.1, 0)
    assert (a.lorb, a.modulation, a.fusion_method, a.av_alpha, a.a_alpha, a.v_alpha, a.t_alpha) == \
        ("m3ae", "Normal", "concat", 0.5, 0.35, 0.25, 0.4)
    assert not (a.gs_flag or a.dynamic or a.modal3 or a.train)
    a = mla_b200.get_arguments(["--ckpt_path", "x", "--lorb", "base", "--gs_flag", "--dynamic", "--modal3",
                                "--modulation", "OGM_GE"])
    assert a.gs_flag and a.dynamic and a.modal3 and a.modulation == "OGM_GE"


def test_out_of_scope_paths_raise():
    import argparse
    import mla_b200
    args = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="Normal", gs_flag=True,
                              dynamic=True, lorb="base", modal3=False, clip=False)
    for bad in (dict(modulation="QMF"), dict(lorb="large"), dict(clip=True), dict(fusion_method="sum")):
        with pytest.raises(NotImplementedError):                  # joint training: QMF / large / clip / sum stay out of scope
            mla_b200.train_epoch(argparse.Namespace(**{**vars(args), **bad}), 0, None, "cpu", [], None, None, gs_flag=False)
    args.fusion_method = "film"
    with pytest.raises(NotImplementedError):
        mla_b200.AVClassifier(args)


def test_joint_mode_head_widths_and_ogm_name_gate():
    """Without --gs_flag the head is the concatenated one (basic_model.py:34,153,221); the OGM name test of main.py:350-362 /
    396-403 selects encoders by the substrings the reference uses (so it never fires for the 2-modality m3ae encoders)."""
    import argparse
    import mla_b200
    from mla_b200.engine import ogm_coeff_index
    a = argparse.Namespace(dataset="CREMAD", fusion_method="concat", modulation="OGM", gs_flag=False, dynamic=False,
                           lorb="base", modal3=False, clip=False)
    assert mla_b200.AVClassifier(a).fusion_module.fc_out.weight.shape == (6, 1024)
    a.gs_flag = True
    assert mla_b200.AVClassifier(a).fusion_module.fc_out.weight.shape == (6, 512)
    tiny = dict(model_type=None, emb_dim=64, depth=1, num_heads=2)
    m = argparse.Namespace(dataset="Food101", fusion_method="concat", modulation="Normal", gs_flag=False, dynamic=False,
                           lorb="m3ae", modal3=False, clip=False)
    assert mla_b200.M3AEClassifier(m, model_config=tiny, text_vocab_size=16).fusion_module.fc_out.weight.shape == (101, 128)
    assert [ogm_coeff_index(n, False) for n in ("audio_net", "visual_net", "mae_a", "mae_v")] == [0, 1, None, None]
    assert [ogm_coeff_index(n, True) for n in ("mae_a", "mae_v", "mae_t", "audio_net")] == [0, 1, 2, None]


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from mla_b200 import dist as mdist
from oracle import mla_oracle as orc
rank, world = mdist.init_from_env("gloo")
assert world == 2
rng = np.random.default_rng(0)
B, D, C = 8, 32, 5
feat = rng.standard_normal((B, D)); W = rng.standard_normal((C, D)) * 0.1; b = rng.standard_normal(C) * 0.1
lab = rng.integers(0, C, B)
full = orc.head_ce(feat, W, b, lab)
lo, hi = rank * B // 2, (rank + 1) * B // 2
loc = orc.head_ce(feat[lo:hi], W, b, lab[lo:hi], grad_scale=1.0 / B)       # scaled by 1/B_GLOBAL
packed = torch.from_numpy(np.concatenate([loc["dW"].ravel(), loc["db"], loc["feat_sum"]]))
mdist.allreduce_sum_(packed)                                                # the small head all-reduce
ref = np.concatenate([full["dW"].ravel(), full["db"], full["feat_sum"]])
assert np.allclose(packed.numpy(), ref, atol=1e-12), "packed head all-reduce != global batch"
# identical inputs -> identical P on every rank (deterministic update): compare checksums
P1, _ = orc.gs_before_update_sum(np.eye(D, dtype=np.float32), packed.numpy()[-D:].astype(np.float32), 1.0 / B,
                                 None, 0.05)
cs = torch.tensor([mdist.params_checksum(torch.from_numpy(P1))])
lst = [torch.zeros_like(cs) for _ in range(world)]
dist.all_gather(lst, cs)
assert lst[0].item() == lst[1].item(), "P differs across ranks"
# flat gradient bucket: views alias the flat buffer, all-reduce sums in place
ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
fg = mdist.FlatGrads(ps); fg.attach()
ps[0].grad += rank + 1; ps[1].grad += 10 * (rank + 1)
mdist.allreduce_sum_(fg.flat)
assert torch.all(ps[0].grad == 3) and torch.all(ps[1].grad == 30)
g = mdist.all_gather_rows(torch.full((2, 3), float(rank)))
assert g.shape == (4, 3) and g[0, 0] == 0 and g[3, 0] == 1
# ragged shards (last batch of a sharded loader without drop_last): rank 0 holds 3 rows, rank 1 holds 1
g = mdist.all_gather_rows(torch.full((3 - 2 * rank, 2), float(rank + 5)))
assert g.shape == (4, 2) and g[:3].eq(5).all() and g[3].eq(6).all()
assert mdist.gather_sizes(3 - 2 * rank, "cpu") == [3, 1]
# bucketed asynchronous all-reduce (tail bucket first, own communicator) == one all-reduce of the whole buffer
flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
bk = mdist.BucketedAllReduce(flat, mdist.aux_group("enc0"))
bk.tail(6); bk.head(6); bk.wait()
assert torch.equal(flat, torch.arange(10, dtype=torch.float32) * 3), flat
assert mdist.aux_group("enc0") is mdist.aux_group("enc0")
# three buckets in the order the backward segments complete them (layer4 | layer3 + 2 | layer1 + stem) next to an
# asynchronous small all-reduce that is awaited later (the head's [dW | db | sum_b feat] packet)
flat = torch.arange(12, dtype=torch.float32) * (rank + 1)
small = torch.full((3,), float(rank + 1))
work = mdist.allreduce_sum_async(small)
bk = mdist.BucketedAllReduce(flat)
bk.span(8, 12); bk.span(3, 8); bk.span(0, 3); bk.wait()
work.wait()
assert torch.equal(flat, torch.arange(12, dtype=torch.float32) * 3) and torch.all(small == 3)
assert mdist.bind_host_to_gpu() is None                     # no CUDA device here: the NUMA binding is a no-op
# loss mean over ranks == global mean
l = torch.tensor([loc["loss"]], dtype=torch.float64); mdist.allreduce_sum_(l)
assert abs(l.item() / world - full["loss"]) < 1e-12
dist.barrier(); print("rank", rank, "ok")
"""


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_flat_gradient_buckets_with_autograd_encoders():
    """The transformer encoders are autograd graphs: gradients ACCUMULATE into the flat per-encoder bucket, which is therefore
    zero-filled when attached; p.grad stays a view of the bucket (the all-reduce and the optimiser read the same memory)."""
    from mla_b200 import dist as mdist
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    fg = mdist.FlatGrads(ps)
    for _ in range(2):                                   # a second step must not see the first step's gradients
        fg.flat.fill_(7.0)
        fg.attach(zero=True)
        (ps[0].sum() * 2 + ps[1].sum() * 3).backward()
        assert torch.equal(fg.flat, torch.tensor([2.0] * 12 + [3.0] * 5))
        assert ps[0].grad.data_ptr() == fg.flat.data_ptr() and ps[1].grad.data_ptr() == fg.flat[12:].data_ptr()
    fg.detach()
    assert ps[0].grad is None and ps[1].grad is None


def test_encoder_param_groups_follow_the_turn_order_and_hot_parameters():
    from mla_b200.engine import encoder_param_groups

    class Enc(torch.nn.Module):
        def __init__(self, hot_only):
            super().__init__()
            self.used, self.unused = torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)
            if hot_only:
                self.hot_parameters = lambda: list(self.used.parameters())

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mae_t, self.mae_a, self.mae_v = Enc(False), Enc(True), Enc(False)      # creation order != turn order

    net = Net()
    groups = encoder_param_groups(net)
    assert len(groups) == 3                                                            # a -> v -> t (main.py:432-466)
    assert [id(p) for p in groups[0]] == [id(p) for p in net.mae_a.used.parameters()]   # only what the forward reads
    assert len(groups[1]) == 4 and groups[1][0] is net.mae_v.used.weight
    assert groups[2][0] is net.mae_t.used.weight
    with pytest.raises(RuntimeError):
        encoder_param_groups(torch.nn.Linear(2, 2))


def test_m3ae_encoder_names_only_the_parameters_its_forward_reads():
    """an image encoder never reads its text embedding table (and vice versa); those parameters must keep grad
    None so that SGD skips them (no weight decay / momentum drift), as in the reference where autograd never touches them."""
    from mla_b200.m3ae import MaskedMultimodalAutoencoder
    enc = MaskedMultimodalAutoencoder(32, dict(model_type=None, emb_dim=64, depth=1, num_heads=2))
    everything = {id(p) for p in enc.parameters()}
    assert {id(p) for p in enc.hot_parameters()} == everything          # before any forward: nothing is known
    enc._used = (True, False)                                           # what forward_representation(image, None, None) records
    hot = {id(p) for p in enc.hot_parameters()}
    cold = everything - hot
    assert cold == {id(enc.text_embedding.weight), id(enc.encoder_text_type_embedding)}
    enc._used = (False, True)
    hot = {id(p) for p in enc.hot_parameters()}
    assert everything - hot == {id(enc.image_embedding.weight), id(enc.image_embedding.bias),
                                id(enc.encoder_image_type_embedding)}
    assert len(hot) == len(enc.hot_parameters())                        # no duplicates
