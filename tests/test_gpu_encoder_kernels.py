"""Parity of the encoder kernels (through the C ABI) with a plain PyTorch fp32 reference of the same op on the GPU:
tcgen05 convolutions (TF32 and 2-byte operands; fprop / dgrad / wgrad, strides 1 and 2, 1x1 and 3x3, ragged sizes),
BatchNorm statistics / apply / backward (incl. the epilogue partial-sum path, ReLU bitmask and 2-byte side outputs),
stem im2col + BN+ReLU+MaxPool, global average pool.

Tolerances (Frobenius-relative):
  TF32 convolutions            2e-3  vs torch fp32 (operands lose 13 mantissa bits: ~8e-4 measured)
  2-byte operand convolutions  2e-5  vs torch fp32 applied to the SAME rounded operands (only the accumulation differs)
  BatchNorm / pooling          1e-5  (fp32 arithmetic, fp64 reductions); outputs that feed convolutions are rounded to
                               TF32 on store, so they are compared at 1e-3
"""
import numpy as np
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


CONV_CASES = [  # N, H, W, Cin, Cout, R, stride
    (2, 8, 8, 64, 64, 1, 1), (2, 8, 8, 64, 64, 3, 1), (1, 16, 16, 64, 128, 3, 2), (2, 9, 6, 128, 128, 3, 1),
    (3, 14, 14, 128, 256, 1, 2), (2, 7, 7, 256, 512, 3, 2), (1, 17, 12, 512, 512, 3, 1), (3, 33, 24, 64, 128, 3, 2),
    (5, 13, 11, 64, 64, 3, 1),
    # large enough for the CTA-pair kernel (256-row tiles): 256-column, 128-column and 64-column pair tiles, ragged last tile
    (8, 14, 14, 256, 256, 3, 1), (12, 14, 14, 256, 512, 3, 2), (9, 17, 12, 128, 128, 3, 1), (11, 13, 11, 64, 64, 3, 1),
    # 64 -> 64 channel 3x3 / 1 geometries that take the halo-strip kernel with resident weights: the two layer1 shapes of the
    # benchmark (56x56, 65x47: odd height -> a one-row last tile), the widest supported row (W + 2 = 64), several tiles / CTA
    (2, 56, 56, 64, 64, 3, 1), (2, 65, 47, 64, 64, 3, 1), (1, 9, 62, 64, 64, 3, 1), (160, 20, 30, 64, 64, 3, 1),
    (4, 17, 12, 64, 64, 3, 1), (8, 16, 16, 64, 64, 3, 1), (3, 33, 24, 64, 64, 3, 1),      # 9-, 7- and 4-row strip tiles
]


def _conv_data(N, H, W, Cin, Cout, R, stride):
    pad = R // 2
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + H + Cin + Cout + R)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, R, R, device="cuda", generator=g) * (1.0 / (Cin * R * R) ** 0.5)
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(N, Cout, OH, OW, device="cuda", generator=g)
    return pad, x, w, dy, OH, OW


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tf32_fprop_dgrad_wgrad(L, case):
    from mla_b200 import ops
    N, H, W, Cin, Cout, R, stride = case
    pad, x, w, dy, OH, OW = _conv_data(*case)
    xn, wn = x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous()
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    y = F.conv2d(x, w, None, stride, pad)
    dx = torch.nn.grad.conv2d_input(x.shape, w, dy, stride, pad)
    dw = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride, pad)
    assert relf(ops.conv2d_fprop(xn, wn, stride, pad).permute(0, 3, 1, 2), y) < 2e-3
    assert relf(ops.conv2d_dgrad(dyn, wn, xn.shape, stride, pad).permute(0, 3, 1, 2), dx) < 2e-3
    acc = torch.ones_like(xn)
    ops.conv2d_dgrad(dyn, wn, xn.shape, stride, pad, out=acc, accumulate=True)          # dx += (residual gradient path)
    assert relf((acc - 1).permute(0, 3, 1, 2), dx) < 2e-3
    assert relf(ops.conv2d_wgrad(xn, dyn, wn.shape, stride, pad).permute(0, 3, 1, 2), dw) < 2e-3
    # BatchNorm partial sums from the fprop epilogue: per-tile (sum, sum of squares) must add up to those of y
    tiles = L.mla_conv2d_fprop_stat_tiles(N, H, W, R, R, stride, pad)
    part = torch.zeros(tiles, 2, Cout, device="cuda")
    yk = torch.empty(N, OH, OW, Cout, device="cuda")
    assert L.mla_conv2d_fprop_bnstats(P(xn), P(wn), P(yk), N, H, W, Cin, Cout, R, R, stride, pad, P(part), st()) == 0
    torch.cuda.synchronize()
    flat = yk.view(-1, Cout).double()
    assert relf(part[:, 0].double().sum(0), flat.sum(0)) < 1e-5
    assert relf(part[:, 1].double().sum(0), (flat * flat).sum(0)) < 1e-5


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_2byte_operands(L, case):
    """fprop16 (fp16 x fp16) and dgrad16 (bf16 dy x transposed bf16 filter) against fp32 math on the rounded operands."""
    N, H, W, Cin, Cout, R, stride = case
    pad, x, w, dy, OH, OW = _conv_data(*case)
    x16 = x.permute(0, 2, 3, 1).contiguous().half()
    wk = w.permute(0, 2, 3, 1).contiguous()
    w16 = torch.empty(Cout, R, R, Cin, dtype=torch.float16, device="cuda")
    assert L.mla_cast16(P(wk), P(w16), wk.numel(), 0, st()) == 0
    wt16 = torch.empty(Cin, R, R, Cout, dtype=torch.bfloat16, device="cuda")
    assert L.mla_filter_transpose16(P(wk), P(wt16), Cout, R * R, Cin, 1, st()) == 0
    dy16 = dy.permute(0, 2, 3, 1).contiguous().bfloat16()
    y = torch.empty(N, OH, OW, Cout, device="cuda")
    dx = torch.empty(N, H, W, Cin, device="cuda")
    assert L.mla_conv2d_fprop16(P(x16), P(w16), P(y), N, H, W, Cin, Cout, R, R, stride, pad, None, st()) == 0
    assert L.mla_conv2d_dgrad16(P(dy16), P(wt16), P(dx), N, H, W, Cin, Cout, R, R, stride, pad, 0, st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(w16, wk.half()) and torch.equal(wt16, wk.permute(3, 1, 2, 0).contiguous().bfloat16())
    yr = F.conv2d(x16.float().permute(0, 3, 1, 2), w16.float().permute(0, 3, 1, 2), None, stride, pad)
    dxr = torch.nn.grad.conv2d_input((N, Cin, H, W), wt16.float().permute(3, 0, 1, 2), dy16.float().permute(0, 3, 1, 2),
                                     stride, pad)
    assert relf(y.permute(0, 3, 1, 2), yr) < 2e-5
    assert relf(dx.permute(0, 3, 1, 2), dxr) < 2e-5
    # the same forward with BatchNorm partial sums from the epilogue: identical y, partials add up to the sums of y
    tiles = L.mla_conv2d_fprop16_stat_tiles(N, H, W, Cin, Cout, R, R, stride, pad)
    part = torch.full((tiles, 2, Cout), float("nan"), device="cuda")
    y2 = torch.empty_like(y)
    assert L.mla_conv2d_fprop16(P(x16), P(w16), P(y2), N, H, W, Cin, Cout, R, R, stride, pad, P(part), st()) == 0
    torch.cuda.synchronize()
    flat = y2.view(-1, Cout).double()
    assert torch.equal(y2, y)
    assert relf(part[:, 0].double().sum(0), flat.sum(0)) < 1e-5 and relf(part[:, 1].double().sum(0), (flat * flat).sum(0)) < 1e-5
    # fp16 operands carry TF32's mantissa: against the unrounded fp32 convolution the error is the TF32 one
    assert relf(y.permute(0, 3, 1, 2), F.conv2d(x, w, None, stride, pad)) < 2e-3
    # wgrad16: bf16 x bf16, both operands MN-major (K = pixels)
    xb = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    dw = torch.empty(Cout, R, R, Cin, device="cuda")
    nb = L.mla_conv2d_wgrad16_workspace_bytes(N, H, W, Cin, Cout, R, R, stride, pad)
    assert nb > 0
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    assert L.mla_conv2d_wgrad16(P(xb), P(dy16), P(dw), N, H, W, Cin, Cout, R, R, stride, pad, P(ws), nb, st()) == 0
    torch.cuda.synchronize()
    dwr = torch.nn.grad.conv2d_weight(xb.float().permute(0, 3, 1, 2), (Cout, Cin, R, R), dy16.float().permute(0, 3, 1, 2),
                                      stride, pad)
    assert relf(dw.permute(0, 3, 1, 2), dwr) < 2e-5


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mag", [3e-7, 1.0, 5e3])
def test_conv_backward_scaled_fp16_operands(L, case, mag):
    """The default backward: BN backward emits dy as fp16 * F (F = a power of two chosen on the device from the data),
    dgrad16_f16 / wgrad16_f16 multiply fp16 x fp16 and undo the scale in the epilogue. Gradients of any magnitude (3e-7 ..
    5e3, three decades of spread inside the tensor) must come out with TF32-class error against fp64, and EXACTLY equal to
    fp32 math on the rounded operands (2e-5)."""
    N, H, W, Cin, Cout, R, stride = case
    pad, x, w, dy, OH, OW = _conv_data(*case)
    gen = torch.Generator(device="cuda").manual_seed(N + H + Cin)
    dy = dy * mag * torch.pow(10.0, -3 * torch.rand(dy.shape, device="cuda", generator=gen))
    M, C = N * OH * OW, Cout
    # a BatchNorm in front of the conv output: dz -> dy through mla_bn_backward_f16
    y = torch.randn(M, C, device="cuda", generator=gen) * 1.5 + 0.2
    gamma = torch.rand(C, device="cuda", generator=gen) + 0.5
    mean, var = y.mean(0), y.var(0, unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    dz = dy.permute(0, 2, 3, 1).reshape(M, C).contiguous()
    xhat = (y - mean) * invstd
    dy_ref = (gamma * invstd * (dz.double() - dz.double().mean(0) - xhat.double() * (dz.double() * xhat.double()).mean(0)))
    ws = torch.zeros(L.mla_bn_workspace_bytes(M, C), dtype=torch.uint8, device="cuda")
    dy16 = torch.empty(M, C, dtype=torch.float16, device="cuda")
    gs = torch.zeros(2, device="cuda")
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    for _ in range(2):                                       # twice: the self-clearing accumulators must leave no state
        assert L.mla_bn_backward_f16(P(dz), None, P(y), P(mean), P(invstd), P(gamma), M, C, P(dg), P(db), P(dy16), None, P(gs),
                                     P(ws), ws.numel(), st()) == 0
    torch.cuda.synchronize()
    Fs, inv = float(gs[0]), float(gs[1])
    assert Fs > 0 and np.log2(Fs) == round(np.log2(Fs)) and Fs * inv == 1.0          # an exact power of two
    amax = float((dy16.float().abs().max()))
    assert 16.0 <= amax <= 65504.0 and not torch.isinf(dy16.float()).any()           # well inside fp16, never saturated here
    assert relf(dy16.double() * inv, dy_ref) < 6e-4                                   # 10-bit mantissa
    assert relf(dg, (dz.double() * xhat.double()).sum(0)) < 1e-4 and relf(db, dz.double().sum(0)) < 1e-4
    # dgrad / wgrad on the scaled operand
    wk = w.permute(0, 2, 3, 1).contiguous()
    wt16 = torch.empty(Cin, R, R, Cout, dtype=torch.float16, device="cuda")
    assert L.mla_filter_transpose16(P(wk), P(wt16), Cout, R * R, Cin, 0, st()) == 0
    x16 = x.permute(0, 2, 3, 1).contiguous().half()
    base = 1.0 if mag == 1.0 else 0.0                     # accumulating form: dx = base + ... (exactly representable sums only)
    dx = torch.full((N, H, W, Cin), base, device="cuda")
    assert L.mla_conv2d_dgrad16_f16(P(dy16), P(wt16), gs.data_ptr() + 4, P(dx), N, H, W, Cin, Cout, R, R, stride, pad, 1,
                                    st()) == 0
    dw = torch.empty(Cout, R, R, Cin, device="cuda")
    nb = L.mla_conv2d_wgrad16_workspace_bytes(N, H, W, Cin, Cout, R, R, stride, pad)
    wsw = torch.empty(nb, dtype=torch.uint8, device="cuda")
    assert L.mla_conv2d_wgrad16_f16(P(x16), P(dy16), gs.data_ptr() + 4, P(dw), N, H, W, Cin, Cout, R, R, stride, pad, P(wsw),
                                    nb, st()) == 0
    torch.cuda.synchronize()
    dyq = (dy16.float() * inv).view(N, OH, OW, Cout).permute(0, 3, 1, 2)             # what the tensor cores multiplied
    dxr = torch.nn.grad.conv2d_input((N, Cin, H, W), wt16.float().permute(3, 0, 1, 2), dyq, stride, pad)
    dwr = torch.nn.grad.conv2d_weight(x16.float().permute(0, 3, 1, 2), (Cout, Cin, R, R), dyq, stride, pad)
    assert relf((dx.double() - base).float().permute(0, 3, 1, 2), dxr) < (2e-5 if base == 0.0 else 2e-4)
    assert relf(dw.permute(0, 3, 1, 2), dwr) < 2e-5
    # against exact math on the unrounded operands: the TF32-class bound
    d64 = dy_ref.view(N, OH, OW, Cout).permute(0, 3, 1, 2)
    dx64 = torch.nn.grad.conv2d_input((N, Cin, H, W), w.double(), d64, stride, pad)
    dw64 = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, R, R), d64, stride, pad)
    assert relf(dw.permute(0, 3, 1, 2), dw64) < 1e-3
    assert relf((dx.double() - base).permute(0, 3, 1, 2), dx64) < 1e-3


def test_filter_transpose_batch(L):
    """Every filter of an encoder transposed to fp16 [Cin][R][S][Cout] by one launch over a segment table."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(64, 9, 64), (128, 9, 64), (128, 1, 64), (256, 9, 128), (96, 9, 40)]
    offs, off = [], 8
    for co, rs, ci in shapes:
        offs.append(off)
        off += (co * rs * ci + 3) // 4 * 4 + 4
    flat = torch.randn(off, device="cuda", generator=gen)
    seg, t0 = [], 0
    for (co, rs, ci), o in zip(shapes, offs):
        seg.append((o, co, rs, ci, t0))
        t0 += rs * ((co + 31) // 32) * ((ci + 31) // 32)
    tab = np.array(seg, dtype=np.dtype([("off", "<i8"), ("co", "<i4"), ("rs", "<i4"), ("ci", "<i4"), ("t0", "<i4")]))
    table = torch.from_numpy(tab.view(np.uint8).copy()).cuda()
    out = torch.zeros(off, dtype=torch.float16, device="cuda")
    assert L.mla_filter_transpose16_batch(P(flat), P(out), P(table), len(seg), t0, 0, st()) == 0
    torch.cuda.synchronize()
    for (co, rs, ci), o in zip(shapes, offs):
        w = flat[o:o + co * rs * ci].view(co, rs, ci)
        assert torch.equal(out[o:o + co * rs * ci].view(ci, rs, co), w.permute(2, 1, 0).contiguous().half())


@pytest.mark.parametrize("N,H,W,C,relu,res", [(2, 5, 3, 64, True, True), (4, 9, 6, 512, True, True),
                                              (16, 28, 28, 128, True, False), (8, 56, 56, 64, False, False),
                                              (3, 7, 7, 256, True, True)])
def test_batchnorm_forward_backward(L, N, H, W, C, relu, res):
    dev = "cuda"
    M = N * H * W
    g0 = torch.Generator(device=dev).manual_seed(M + C)
    y = torch.randn(M, C, device=dev, generator=g0) * 2 + 0.5
    gamma = torch.rand(C, device=dev, generator=g0) + 0.5
    beta = torch.randn(C, device=dev, generator=g0) * 0.1
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    rm2, rv2 = rm.clone(), rv.clone()
    mean, invstd, scale, shift = (torch.empty(C, device=dev) for _ in range(4))
    ws = torch.zeros(L.mla_bn_workspace_bytes(M, C), dtype=torch.uint8, device=dev)
    assert L.mla_bn_train_stats(P(y), M, C, P(gamma), P(beta), P(rm), P(rv), 0.1, 1e-5, P(mean), P(invstd), P(scale), P(shift),
                                P(ws), ws.numel(), st()) == 0
    yt = y.clone().requires_grad_(True)
    gt, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    idn = torch.randn(M, C, device=dev, generator=g0)
    o = F.batch_norm(yt.view(N, H, W, C).permute(0, 3, 1, 2), rm2, rv2, gt, bt, True, 0.1, 1e-5).permute(0, 2, 3, 1).reshape(M, C)
    z = o + idn if res else o
    z = F.relu(z) if relu else z
    out = torch.empty(M, C, device=dev)
    mask = torch.zeros(M * C // 32, dtype=torch.int32, device=dev)
    out16 = torch.empty(M, C, dtype=torch.float16, device=dev)
    out16b = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
    assert L.mla_bn_apply_ex(P(y), P(scale), P(shift), P(idn) if res else None, None, None, 1 if relu else 0, P(out), P(mask),
                             P(out16), P(out16b), M, C, st()) == 0
    dz = torch.randn(M, C, device=dev, generator=g0)
    z.backward(dz)
    dy, g = torch.empty(M, C, device=dev), torch.empty(M, C, device=dev)
    dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
    assert L.mla_bn_backward(P(dz), P(out) if relu else None, P(y), P(mean), P(invstd), P(gamma), M, C, P(dg), P(db), P(dy),
                             P(g), P(ws), ws.numel(), st()) == 0
    torch.cuda.synchronize()
    assert relf(mean, y.mean(0)) < 1e-5 and relf(rm, rm2) < 1e-5 and relf(rv, rv2) < 1e-5     # torch running-stat semantics
    assert relf(out, z.detach()) < 1e-3                                  # stored TF32-rounded
    assert relf(out16.float(), z.detach()) < 1e-3                        # fp16 copy = the same 10-bit mantissa
    assert relf(out16b.float(), z.detach()) < 4e-3                       # bf16 copy: 8-bit mantissa
    assert relf(dg, gt.grad) < 1e-4 and relf(db, bt.grad) < 1e-4
    assert relf(dy, yt.grad) < 1e-3
    assert relf(g, dz * (z.detach() > 0) if relu else dz) < 1e-6
    if relu:
        # the ReLU bitmask written by the forward replaces the activation in the backward: identical results
        bits = (out.view(-1) > 0).view(-1, 32).to(torch.int64)
        words = (bits << torch.arange(32, device=dev)).sum(1)
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
        assert torch.equal(mask, words)
        dy2, g2, dy16 = torch.empty_like(dy), torch.empty_like(g), torch.empty(M, C, dtype=torch.bfloat16, device=dev)
        dg2, db2 = torch.empty_like(dg), torch.empty_like(db)
        assert L.mla_bn_backward_ex(P(dz), None, P(mask), P(y), P(mean), P(invstd), P(gamma), M, C, P(dg2), P(db2), P(dy2),
                                    P(dy16), P(g2), P(ws), ws.numel(), st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(dy2, dy) and torch.equal(g2, g) and torch.equal(dg2, dg) and torch.equal(db2, db)
        assert relf(dy16.float(), dy) < 4e-3                              # bf16: 8-bit mantissa
        dy16b = torch.empty_like(dy16)                                    # 2-byte output only (no fp32 dy)
        assert L.mla_bn_backward_ex(P(dz), None, P(mask), P(y), P(mean), P(invstd), P(gamma), M, C, P(dg2), P(db2), None,
                                    P(dy16b), None, P(ws), ws.numel(), st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(dy16b, dy16)


@pytest.mark.parametrize("B,T,Cin,H,W", [(2, 1, 1, 65, 48), (2, 2, 3, 64, 64), (1, 1, 1, 257, 188)])
def test_stem_im2col_conv_maxpool(L, B, T, Cin, H, W):
    dev = "cuda"
    x = torch.randn(B, Cin, T, H, W, device=dev) if T > 1 else torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(64, Cin, 7, 7, device=dev) * 0.1
    wk = w.permute(0, 2, 3, 1).contiguous()
    N = B * T
    OH, OW = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    Kp = {1: 64, 3: 160}[Cin]
    col = torch.empty(N * OH * OW, Kp, device=dev)
    HW = H * W
    sB, sT, sC = (Cin * T * HW, HW, T * HW) if T > 1 else (Cin * HW, 0, HW)
    assert L.mla_stem_im2col(P(x), P(col), N, T, sB, sT, sC, Cin, H, W, 7, 7, 2, 3, Kp, st()) == 0
    wpad = torch.empty(64, Kp, device=dev)
    assert L.mla_pad_rows(P(wk), P(wpad), 64, 49 * Cin, Kp, 0, st()) == 0
    y = torch.empty(N, OH, OW, 64, device=dev)
    assert L.mla_conv2d_fprop(P(col), P(wpad), P(y), N, OH, OW, Kp, 64, 1, 1, 1, 0, st()) == 0
    xr = x.permute(0, 2, 1, 3, 4).reshape(N, Cin, H, W) if T > 1 else x
    ref = F.conv2d(xr, w, None, 2, 3)
    scale, shift = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.2
    PH, PW = (OH + 2 - 3) // 2 + 1, (OW + 2 - 3) // 2 + 1
    p = torch.empty(N, PH, PW, 64, device=dev)
    p16 = torch.empty(N, PH, PW, 64, dtype=torch.float16, device=dev)
    idx = torch.empty(N, PH, PW, 64, dtype=torch.uint8, device=dev)
    assert L.mla_bn_relu_maxpool_ex(P(y), P(scale), P(shift), P(p), P(p16), None, P(idx), N, OH, OW, 64, st()) == 0
    yr = y.permute(0, 3, 1, 2).clone().requires_grad_(True)
    pr = F.max_pool2d(F.relu(yr * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)), 3, 2, 1)
    dp = torch.randn_like(pr)
    pr.backward(dp)
    g = torch.empty(N, OH, OW, 64, device=dev)
    dpn = dp.permute(0, 2, 3, 1).contiguous()
    assert L.mla_maxpool_relu_backward(P(dpn), P(p), P(idx), P(g), N, OH, OW, 64, st()) == 0
    torch.cuda.synchronize()
    assert relf(y.permute(0, 3, 1, 2), ref) < 2e-3                       # TF32 stem GEMM
    assert relf(p.permute(0, 3, 1, 2), pr.detach()) < 1e-3 and relf(p16.float().permute(0, 3, 1, 2), pr.detach()) < 1e-3
    assert relf(g, (yr.grad / scale.view(1, -1, 1, 1)).permute(0, 2, 3, 1)) < 1e-6


@pytest.mark.parametrize("B,rows,C", [(4, 54, 512), (3, 98, 512), (1, 1, 64)])
def test_global_average_pool(L, B, rows, C):
    fm = torch.randn(B * rows, C, device="cuda")
    feat = torch.empty(B, C, device="cuda")
    assert L.mla_avgpool_forward(P(fm), P(feat), B, rows, C, st()) == 0
    df = torch.randn(B, C, device="cuda")
    dfm = torch.empty(B * rows, C, device="cuda")
    assert L.mla_avgpool_backward(P(df), P(dfm), B, rows, C, st()) == 0
    torch.cuda.synchronize()
    assert relf(feat, fm.view(B, rows, C).mean(1)) < 1e-6
    assert relf(dfm, (df / rows).view(B, 1, C).expand(B, rows, C).reshape(B * rows, C)) < 1e-6


def test_conv_argument_errors(L):
    x = torch.zeros(1, 8, 8, 64, device="cuda")
    w = torch.zeros(64, 3, 3, 64, device="cuda")
    y = torch.zeros(1, 8, 8, 64, device="cuda")
    assert L.mla_conv2d_fprop(None, P(w), P(y), 1, 8, 8, 64, 64, 3, 3, 1, 1, st()) < 0                 # null pointer
    assert L.mla_conv2d_fprop(P(x), P(w), P(y), 1, 8, 8, 48, 64, 3, 3, 1, 1, st()) < 0                 # Cin % 32
    assert L.mla_conv2d_fprop(P(x), P(w), P(y), 1, 8, 8, 64, 64, 3, 3, 3, 1, st()) < 0                 # stride 3
    assert L.mla_conv2d_fprop16(P(x), P(w), P(y), 1, 8, 8, 96, 64, 3, 3, 1, 1, None, st()) < 0         # Cin % 64
    xb, yb = torch.zeros(8, 16, 16, 64, device="cuda"), torch.zeros(8, 16, 16, 64, device="cuda")
    assert L.mla_conv2d_wgrad_workspace_bytes(8, 16, 16, 64, 64, 3, 3, 1, 1) > 256
    assert L.mla_conv2d_wgrad(P(xb), P(yb), P(w), 8, 16, 16, 64, 64, 3, 3, 1, 1, None, 0, st()) < 0    # split-K workspace missing


_PAIR_WORKER = r"""
import sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
import torch
import test_gpu_encoder_kernels as T
from mla_b200 import _lib
L = _lib.lib()
torch.backends.cudnn.allow_tf32 = False
for case in [(8, 14, 14, 256, 256, 3, 1), (12, 14, 14, 256, 512, 3, 2), (9, 17, 12, 128, 128, 3, 1), (11, 13, 11, 64, 64, 3, 1)]:
    T.test_conv_2byte_operands(L, case)
    T.test_conv_backward_scaled_fp16_operands(L, case, 1.0)
print("pair kernel ok")
"""


def test_conv_pair_kernel(L, tmp_path):
    """The opt-in CTA-pair persistent kernel (MLA_CONV_PAIR16=1: cta_group::2, 256 x {64, 128, 256} tiles) through the
    same parity checks as the default kernels: forward with BatchNorm partial sums, plain and accumulating dgrad."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "pair_worker.py"
    script.write_text(_PAIR_WORKER % {"root": root})
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, MLA_CONV_PAIR16="1"))
    assert r.returncode == 0 and "pair kernel ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
