"""Parity of the im2col-free stem (csrc/stem_s2d.cu: space-to-depth + sliding-window TMA + tcgen05) and of the fused
BN + ReLU + MaxPool forward / backward around it (models/backbone.py:78-83, 149-152) through the C ABI, against plain PyTorch
on the GPU.

Tolerances (Frobenius-relative):
  fprop   y is STORED as fp16: 4e-4 vs torch fp32 on the same fp16-rounded operands (half an fp16 ulp is 2.4e-4 relative at
          worst, ~1.6e-4 rms); the BatchNorm partial sums come from the fp32 accumulators: 2e-5
  wgrad   2e-5 vs torch fp32 on the same rounded operands (only the accumulation order differs); 1e-3 vs fp64 exact operands
  pool / BatchNorm backward: same numbers as the unfused kernels it replaces (dgamma / dbeta 1e-4, dy = fp16 * 2^k: 1e-3)
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from mla_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib.lib()


def st():
    return torch.cuda.current_stream().cuda_stream


def P(t):
    return None if t is None else t.data_ptr()


def relf(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _pack(L, x, B, T, Cin, H, W):
    N = B * T
    HW = H * W
    sB, sT, sC = (Cin * T * HW, HW, T * HW) if T > 1 else (Cin * HW, 0, HW)
    xs = torch.empty(L.mla_stem_s2d_input_elems(N, H, W), dtype=torch.float16, device="cuda")
    assert L.mla_stem_s2d_pack(P(x), P(xs), N, T, sB, sT, sC, Cin, H, W, st()) == 0
    return xs


STEM_CASES = [(2, 1, 1, 65, 48), (2, 2, 3, 64, 64), (1, 1, 1, 257, 188), (3, 2, 3, 224, 224), (1, 1, 1, 7, 9), (5, 1, 3, 33, 31)]


@pytest.mark.parametrize("B,T,Cin,H,W", STEM_CASES)
def test_stem_s2d_fprop_wgrad(L, B, T, Cin, H, W):
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(B * 1000 + H)
    x = (torch.randn(B, Cin, T, H, W, device=dev, generator=gen) if T > 1
         else torch.randn(B, Cin, H, W, device=dev, generator=gen))
    w = torch.randn(64, Cin, 7, 7, device=dev, generator=gen) * 0.1
    wk = w.permute(0, 2, 3, 1).contiguous()                       # [64][7][7][Cin]: the channels_last memory of the parameter
    N = B * T
    OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    xs = _pack(L, x, B, T, Cin, H, W)
    w2 = torch.empty(64 * 256, dtype=torch.float16, device=dev)
    assert L.mla_stem_s2d_weights(P(wk), P(w2), Cin, st()) == 0
    ntiles = L.mla_stem_s2d_tiles(N, H, W)
    assert ntiles == N * ((OH + 7) // 8) * ((OW + 15) // 16)
    part = torch.full((ntiles, 2, 64), float("nan"), device=dev)
    y16 = torch.full((N, OH, OW, 64), float("nan"), dtype=torch.float16, device=dev)
    assert L.mla_stem_s2d_fprop(P(xs), P(w2), P(y16), N, H, W, P(part), st()) == 0
    torch.cuda.synchronize()
    xr = (x.permute(0, 2, 1, 3, 4).reshape(N, Cin, H, W) if T > 1 else x)
    xq, wq = xr.half().float(), w.half().float()                   # what the tensor cores multiplied
    ref = F.conv2d(xq, wq, None, 2, 3)
    assert torch.isfinite(y16.float()).all()
    assert relf(y16.float().permute(0, 3, 1, 2), ref) < 4e-4
    assert relf(y16.float().permute(0, 3, 1, 2), F.conv2d(xr.double(), w.double(), None, 2, 3)) < 1e-3
    sums = part.double().sum(0)
    s1, sabs = ref.double().sum((0, 2, 3)), ref.double().abs().sum((0, 2, 3))
    assert bool(((sums[0] - s1).abs() <= 2e-5 * sabs + 1e-6).all())          # sums cancel: bound against sum |y|
    assert relf(sums[1], (ref.double() ** 2).sum((0, 2, 3))) < 2e-5
    # evaluation form: no statistics
    y16b = torch.empty_like(y16)
    assert L.mla_stem_s2d_fprop(P(xs), P(w2), P(y16b), N, H, W, None, st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(y16b, y16)

    # ---- weight gradient: dy fp16 (already scaled by F), out_scale = 1 / F
    dy = torch.randn(N, OH, OW, 64, device=dev, generator=gen)
    Fs = 64.0
    dy16 = (dy * Fs).half()
    inv = torch.tensor([1.0 / Fs], device=dev)
    dw = torch.full((64, 7, 7, Cin), float("nan"), device=dev)
    nb = L.mla_stem_s2d_wgrad_workspace_bytes()
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    assert L.mla_stem_s2d_wgrad(P(xs), P(dy16), P(inv), P(dw), N, H, W, Cin, P(ws), nb, st()) == 0
    dw2 = torch.empty_like(dw)
    assert L.mla_stem_s2d_wgrad(P(xs), P(dy16), P(inv), P(dw2), N, H, W, Cin, P(ws), nb, st()) == 0
    torch.cuda.synchronize()
    dyq = (dy16.float() / Fs).permute(0, 3, 1, 2)
    dwr = torch.nn.grad.conv2d_weight(xq, (64, Cin, 7, 7), dyq, 2, 3)
    assert relf(dw.permute(0, 3, 1, 2), dwr) < 2e-5
    dw64 = torch.nn.grad.conv2d_weight(xr.double(), (64, Cin, 7, 7), dy.double().permute(0, 3, 1, 2), 2, 3)
    assert relf(dw.permute(0, 3, 1, 2), dw64) < 1e-3
    assert torch.equal(dw, dw2)                                    # deterministic
    assert L.mla_stem_s2d_wgrad(P(xs), P(dy16), P(inv), P(dw), N, H, W, Cin, None, 0, st()) < 0       # workspace missing
    assert L.mla_stem_s2d_fprop(P(xs), P(w2), None, N, H, W, None, st()) < 0


@pytest.mark.parametrize("N,H,W", [(2, 33, 24), (3, 129, 94), (4, 112, 112), (1, 4, 5)])
def test_stem_pool_bn_forward_backward(L, N, H, W):
    """BN + ReLU + MaxPool(3, 2, 1) over the fp16 convolution output, and BatchNorm backward with the pooling / ReLU
    backward folded into its two passes, against autograd on the same fp16-rounded y."""
    dev, C = "cuda", 64
    gen = torch.Generator(device=dev).manual_seed(N * 100 + H)
    y16 = (torch.randn(N, H, W, C, device=dev, generator=gen) * 2 + 0.3).half()
    gamma = torch.rand(C, device=dev, generator=gen) + 0.5
    beta = torch.randn(C, device=dev, generator=gen) * 0.2
    M = N * H * W
    yf = y16.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gt, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(yf, None, None, gt, bt, True, 0.1, 1e-5)
    pr = F.max_pool2d(F.relu(z), 3, 2, 1)
    dp = torch.randn(pr.shape, device=dev, generator=gen) * 1e-3
    pr.backward(dp)
    mean = y16.float().view(M, C).double().mean(0)
    var = y16.float().view(M, C).double().var(0, unbiased=False)
    invstd = (1.0 / torch.sqrt(var + 1e-5)).float()
    mean = mean.float()
    scale = gamma * invstd
    shift = beta - mean * scale
    PH, PW = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    p = torch.empty(N, PH, PW, C, device=dev)
    p16 = torch.empty(N, PH, PW, C, dtype=torch.float16, device=dev)
    idx = torch.empty(N, PH, PW, C, dtype=torch.uint8, device=dev)
    assert L.mla_bn_relu_maxpool16(P(y16), P(scale), P(shift), P(p), P(p16), P(idx), N, H, W, C, st()) == 0
    dpn = dp.permute(0, 2, 3, 1).contiguous()
    dy16 = torch.empty(N, H, W, C, dtype=torch.float16, device=dev)
    dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
    gs = torch.zeros(2, device=dev)
    ws = torch.zeros(L.mla_bn_workspace_bytes(M, C), dtype=torch.uint8, device=dev)
    for _ in range(2):                                             # twice: the workspace tickets must come back to zero
        assert L.mla_pool_bn_backward_f16(P(dpn), P(idx), P(y16), P(mean), P(invstd), P(gamma), N, H, W, C, P(dg), P(db),
                                          P(dy16), P(gs), P(ws), ws.numel(), st()) == 0
    torch.cuda.synchronize()
    assert relf(p.permute(0, 3, 1, 2), pr.detach()) < 1e-5
    assert relf(p16.float().permute(0, 3, 1, 2), pr.detach()) < 1e-3
    dead = (p <= 0)
    assert torch.equal((idx >= 16), dead)
    Fv, inv = float(gs[0]), float(gs[1])
    assert Fv > 0 and abs(Fv * inv - 1) < 1e-6 and abs(torch.log2(gs[0]).item() - round(torch.log2(gs[0]).item())) < 1e-6
    assert relf(dg, gt.grad) < 1e-4 and relf(db, bt.grad) < 1e-4
    assert relf(dy16.float() * inv, yf.grad.permute(0, 2, 3, 1)) < 1e-3
    assert float(dy16.float().abs().max()) >= 64.0                # the scale really lifts the operand into fp16's range
