"""m3ae transformer encoder and M3AEClassifier (reference models/m3ae.py:65-179,226-370 and
models/basic_model.py:127-200) for the --lorb m3ae --gs_flag path (BASELINE.json configs[2], Food-101-shaped).

Parameter names, shapes and creation order follow the reference (`text_embedding`, `image_embedding`,
`encoder_{image,text}_type_embedding`, `cls_token`, `encoder.blocks.N.{layer_norm1,attention.qkv_linear,attention.fc,
layer_norm2,transformer_mlp.fc1,transformer_mlp.fc2}`, `encoder.layer_norm`), so seeding reproduces its initial weights
and its state dicts load. What runs where:

  * every Linear of the encoder (patch embedding, qkv, attention output, fc1, fc2 — >99 % of the FLOPs) runs on the
    library's tcgen05 GEMMs: the 1x1 case of the implicit-GEMM convolution kernels, fp16 x fp16 forward, bf16 x bf16 for
    dx = dy W and dW = dy^T x (`NativeLinear`, csrc/conv_gemm.cu);
  * the attention core (scores, -1e7 key-padding fill, softmax, weighted sum; forward and backward) is one fused
    tensor-core kernel per pass (fp16 operands, fp32 softmax) that never writes the S x S matrix (csrc/attention.cu);
  * LayerNorm (forward / backward), exact GELU, the fp16 / TF32 operand copies and the bias gradients are fused
    single-pass kernels (csrc/transformer_ops.cu); bias and residual adds ride in the GEMM epilogues; a whole block is one
    autograd node with a hand-written backward (`_BlockFn`);
  * what is left to ATen: the embedding lookup, the positional / type-embedding adds, the token mean and the optimiser.

Reference behaviours kept: DropPath is the identity (the only configured rate is 0; as published its forward returns
None and the model cannot run); the token mean includes CLS and padded tokens (basic_model.py:193-194).
No CPU fallback: NativeLinear raises on non-CUDA tensors.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .fusion_modules import ConcatFusion

_SIZES = {  # emb_dim, depth, heads (m3ae.py:226-270)
    "small": (384, 12, 6), "base": (768, 12, 12), "large": (1024, 24, 16), "huge": (1280, 32, 16), "debug": (1024, 2, 16),
}


class _Workspace:
    """Grow-only device scratch, one buffer per (device, CUDA stream): the kernels that use it are ordered on their stream,
    and two streams never share a buffer."""
    bufs = {}

    @classmethod
    def get(cls, nbytes, device):
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        buf = cls.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            cls.bufs[key] = buf
        return buf


def _cast16(x, bf16):
    out = torch.empty(x.shape, dtype=torch.bfloat16 if bf16 else torch.float16, device=x.device)
    _lib.check(_lib.lib().mla_cast16(x.data_ptr(), out.data_ptr(), x.numel(), 1 if bf16 else 0, _lib.stream_ptr()), "mla_cast16")
    return out


def _round_tf32(x):
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mla_round_tf32(x.data_ptr(), out.data_ptr(), x.numel(), _lib.stream_ptr()), "mla_round_tf32")
    return out


# Backward operand type. False (default): TF32 kernels on operands rounded to nearest TF32 — forward (fp16 operands, the
# same 10-bit mantissa) and backward then carry TF32 precision throughout, inside BASELINE.json's fp32/TF32 rel-1e-3
# tolerance. True: bf16 x bf16 kernels (twice the tensor-core rate, 8-bit mantissa: ~2e-3 relative error on gradients).
BACKWARD_BF16 = False


class _LinearFn(torch.autograd.Function):
    """y = x W^T (+ b) on the tcgen05 GEMM kernels: a Linear is the 1x1 case of the implicit-GEMM convolution over an
    N=1 image of M x 1 pixels. x [M, K] fp32 contiguous, W [N, K]; K, N % 64 == 0."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        L = _lib.lib()
        M, K = x.shape
        N = weight.shape[0]
        st = _lib.stream_ptr()
        x16 = _cast16(x, False)
        w16 = _cast16(weight, False)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        _lib.check(L.mla_conv2d_fprop16(x16.data_ptr(), w16.data_ptr(), y.data_ptr(), 1, M, 1, K, N, 1, 1, 1, 0, None, st),
                   "mla_conv2d_fprop16")
        if bias is not None:
            y += bias
        saved_x = None
        if weight.requires_grad:
            saved_x = _cast16(x, True) if BACKWARD_BF16 else _round_tf32(x)
        ctx.save_for_backward(saved_x, weight)
        ctx.has_bias = bias is not None
        ctx.need_dx = x.requires_grad
        ctx.bf16 = BACKWARD_BF16
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        xs, weight = ctx.saved_tensors
        dy = dy.contiguous()
        M, N = dy.shape
        K = weight.shape[1]
        st = _lib.stream_ptr()
        dev = dy.device
        dx = dw = db = None
        dyo = _cast16(dy, True) if ctx.bf16 else _round_tf32(dy)
        if ctx.need_dx:
            dx = torch.empty(M, K, dtype=torch.float32, device=dev)
            if ctx.bf16:
                wt = torch.empty(K, N, dtype=torch.bfloat16, device=dev)              # W^T: the K-major operand of dx = dy W
                _lib.check(L.mla_filter_transpose16(weight.data_ptr(), wt.data_ptr(), N, 1, K, 1, st), "mla_filter_transpose16")
                _lib.check(L.mla_conv2d_dgrad16(dyo.data_ptr(), wt.data_ptr(), dx.data_ptr(), 1, M, 1, K, N, 1, 1, 1, 0, 0, st),
                           "mla_conv2d_dgrad16")
            else:
                wr = _round_tf32(weight)
                _lib.check(L.mla_conv2d_dgrad(dyo.data_ptr(), wr.data_ptr(), dx.data_ptr(), 1, M, 1, K, N, 1, 1, 1, 0, 0, st),
                           "mla_conv2d_dgrad")
        if xs is not None:
            dw = torch.empty(N, K, dtype=torch.float32, device=dev)
            if ctx.bf16:
                ws = _Workspace.get(L.mla_conv2d_wgrad16_workspace_bytes(1, M, 1, K, N, 1, 1, 1, 0), dev)
                _lib.check(L.mla_conv2d_wgrad16(xs.data_ptr(), dyo.data_ptr(), dw.data_ptr(), 1, M, 1, K, N, 1, 1, 1, 0,
                                                ws.data_ptr(), ws.numel(), st), "mla_conv2d_wgrad16")
            else:
                ws = _Workspace.get(L.mla_conv2d_wgrad_workspace_bytes(1, M, 1, K, N, 1, 1, 1, 0), dev)
                _lib.check(L.mla_conv2d_wgrad(xs.data_ptr(), dyo.data_ptr(), dw.data_ptr(), 1, M, 1, K, N, 1, 1, 1, 0,
                                              ws.data_ptr(), ws.numel(), st), "mla_conv2d_wgrad")
        if ctx.has_bias:
            db = dy.sum(0)
        return dx, dw, db


class _AttentionFn(torch.autograd.Function):
    """softmax(scale * q k^T, padded keys filled with -1e7) v for every head, fused (csrc/attention.cu). qkv [B, S, 3*H*Dh]
    as produced by qkv_linear, mask [B, S] float (> 0: padded key) or None -> [B, S, H*Dh]."""

    @staticmethod
    def forward(ctx, qkv, mask, num_heads, scale):
        B, S, C3 = qkv.shape
        Dh = C3 // (3 * num_heads)
        qkv = qkv.contiguous()
        dev = qkv.device
        out = torch.empty(B, S, num_heads * Dh, dtype=torch.float32, device=dev)
        stats = torch.empty(B, num_heads, S, 2, dtype=torch.float32, device=dev)
        qkv16 = torch.empty(B, S, C3, dtype=torch.float16, device=dev)
        _lib.check(_lib.lib().mla_attention_forward(qkv.data_ptr(), mask.data_ptr() if mask is not None else None,
                                                    out.data_ptr(), stats.data_ptr(), qkv16.data_ptr(), B, S, num_heads, Dh,
                                                    scale, _lib.stream_ptr()), "mla_attention_forward")
        ctx.save_for_backward(qkv16, mask, out, stats)
        ctx.cfg = (B, S, num_heads, Dh, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv16, mask, out, stats = ctx.saved_tensors
        B, S, H, Dh, scale = ctx.cfg
        L = _lib.lib()
        dout = dout.contiguous()
        dqkv = torch.empty(B, S, 3 * H * Dh, dtype=torch.float32, device=dout.device)
        ws = torch.empty(L.mla_attention_backward_workspace_bytes(B, S, H, Dh), dtype=torch.uint8, device=dout.device)
        _lib.check(L.mla_attention_backward(qkv16.data_ptr(), mask.data_ptr() if mask is not None else None, out.data_ptr(),
                                            dout.data_ptr(), stats.data_ptr(), dqkv.data_ptr(), B, S, H, Dh, scale,
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mla_attention_backward")
        return dqkv, None, None, None


# One autograd node per transformer block (forward and hand-written backward over the C ABI) instead of one per op: every
# activation is produced directly in the formats its consumers need (fp16 GEMM operand, TF32-rounded weight-gradient operand),
# bias / residual adds ride in the GEMM epilogues, LayerNorm / GELU / bias-gradient work is fused into single passes
# (csrc/transformer_ops.cu). False: the per-module path above (same arithmetic up to summation order), which also offers
# BACKWARD_BF16.
FUSED_BLOCK = True
# Backward GEMM operands of the fused block. "fp16": x in fp16 (the copy the forward GEMM already consumed) and dy in fp16
# after an exact power-of-two scale taken from max|dy| (mla_grad_operand16) — TF32's 10-bit mantissa at the full fp16
# tensor-core rate, nothing kept in fp32 for the backward GEMMs. "tf32": TF32-rounded fp32 operands (half the rate).
BLOCK_BACKWARD = "fp16"


def _scratch(nbytes, dev):
    return _Workspace.get(nbytes, dev)


def _ln_fwd(x2, w, b, eps, want_y=False, want_r=True):
    L = _lib.lib()
    M, D = x2.shape
    dev = x2.device
    y = torch.empty(M, D, dtype=torch.float32, device=dev) if want_y else None
    y16 = None if want_y else torch.empty(M, D, dtype=torch.float16, device=dev)
    y_r = None if (want_y or not want_r) else torch.empty(M, D, dtype=torch.float32, device=dev)
    mean = torch.empty(M, dtype=torch.float32, device=dev)
    rstd = torch.empty(M, dtype=torch.float32, device=dev)
    _lib.check(L.mla_layernorm_forward(x2.data_ptr(), w.data_ptr(), b.data_ptr(), eps, M, D,
                                       y.data_ptr() if want_y else None, None if want_y else y16.data_ptr(),
                                       y_r.data_ptr() if y_r is not None else None, mean.data_ptr(), rstd.data_ptr(),
                                       _lib.stream_ptr()), "mla_layernorm_forward")
    return y, y16, y_r, mean, rstd


def _ln_bwd(dy2, x2, mean, rstd, w, resid):
    L = _lib.lib()
    M, D = x2.shape
    dev = x2.device
    dx = torch.empty(M, D, dtype=torch.float32, device=dev)
    dw = torch.empty(D, dtype=torch.float32, device=dev)
    db = torch.empty(D, dtype=torch.float32, device=dev)
    ws = _scratch(L.mla_layernorm_backward_workspace_bytes(M, D), dev)
    _lib.check(L.mla_layernorm_backward(dy2.data_ptr(), x2.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w.data_ptr(),
                                        resid.data_ptr() if resid is not None else None, M, D, dx.data_ptr(), dw.data_ptr(),
                                        db.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mla_layernorm_backward")
    return dx, dw, db


def _linear16(x16, w, bias, resid, M, K, N):
    """y = x16 w^T + bias (+ resid) with the adds in the GEMM epilogue; w fp32 [N, K] is cast here (small)."""
    y = torch.empty(M, N, dtype=torch.float32, device=x16.device)
    w16 = _cast16(w, False)
    _lib.check(_lib.lib().mla_linear_forward16(x16.data_ptr(), w16.data_ptr(), bias.data_ptr() if bias is not None else None,
                                               resid.data_ptr() if resid is not None else None, y.data_ptr(), M, K, N,
                                               _lib.stream_ptr()), "mla_linear_forward16")
    return y


def _cast_round(x2, gelu, want_r=True):
    x16 = torch.empty(x2.shape, dtype=torch.float16, device=x2.device)
    x_r = torch.empty_like(x2) if want_r else None
    _lib.check(_lib.lib().mla_cast_round(x2.data_ptr(), x16.data_ptr(), x_r.data_ptr() if want_r else None, x2.numel(),
                                         1 if gelu else 0, _lib.stream_ptr()), "mla_cast_round")
    return x16, x_r


def _grad_operand16(dy2, u):
    """fp16(F * dy [* gelu'(u)]) with F a power of two from max|.|, the bias gradient, and the device scalar 1/F."""
    L = _lib.lib()
    M, N = dy2.shape
    dev = dy2.device
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    col = torch.empty(N, dtype=torch.float32, device=dev)
    sc = torch.empty(4, dtype=torch.float32, device=dev)
    ws = _scratch(L.mla_round_colsum_workspace_bytes(M, N), dev)
    _lib.check(L.mla_grad_operand16(dy2.data_ptr(), u.data_ptr() if u is not None else None, out.data_ptr(), col.data_ptr(),
                                    sc.data_ptr(), M, N, ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mla_grad_operand16")
    return out, col, sc[1:]


def _linear_grads16(x16, dy16, inv_scale, w, M, K, N):
    """dx [M, K] = dy w, dw [N, K] = dy^T x from fp16 operands (dy16 = F * dy, results multiplied by 1/F in the epilogues)."""
    L = _lib.lib()
    dev = dy16.device
    st = _lib.stream_ptr()
    wt16 = torch.empty(K, N, dtype=torch.float16, device=dev)
    _lib.check(L.mla_filter_transpose16(w.data_ptr(), wt16.data_ptr(), N, 1, K, 0, st), "mla_filter_transpose16")
    dx = torch.empty(M, K, dtype=torch.float32, device=dev)
    _lib.check(L.mla_linear_dgrad16(dy16.data_ptr(), wt16.data_ptr(), inv_scale.data_ptr(), dx.data_ptr(), M, K, N, st),
               "mla_linear_dgrad16")
    dw = torch.empty(N, K, dtype=torch.float32, device=dev)
    ws = _scratch(L.mla_conv2d_wgrad16_workspace_bytes(1, M, 1, K, N, 1, 1, 1, 0), dev)
    _lib.check(L.mla_linear_wgrad16(x16.data_ptr(), dy16.data_ptr(), inv_scale.data_ptr(), dw.data_ptr(), M, K, N,
                                    ws.data_ptr(), ws.numel(), st), "mla_linear_wgrad16")
    return dx, dw


def _round_colsum(dy2, u):
    """TF32-rounded dy (times gelu'(u) when u is given) and its column sums (the bias gradient)."""
    L = _lib.lib()
    M, N = dy2.shape
    out = torch.empty_like(dy2)
    col = torch.empty(N, dtype=torch.float32, device=dy2.device)
    ws = _scratch(L.mla_round_colsum_workspace_bytes(M, N), dy2.device)
    _lib.check(L.mla_round_colsum(dy2.data_ptr(), u.data_ptr() if u is not None else None, out.data_ptr(), col.data_ptr(), M, N,
                                  ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mla_round_colsum")
    return out, col


def _linear_grads(x_r, dy_r, w, M, K, N):
    """dx [M, K] = dy_r w, dw [N, K] = dy_r^T x_r on the TF32 tcgen05 kernels (operands already TF32-rounded)."""
    L = _lib.lib()
    dev = dy_r.device
    st = _lib.stream_ptr()
    w_r = _round_tf32(w)
    dx = torch.empty(M, K, dtype=torch.float32, device=dev)
    _lib.check(L.mla_linear_dgrad(dy_r.data_ptr(), w_r.data_ptr(), dx.data_ptr(), M, K, N, st), "mla_linear_dgrad")
    dw = torch.empty(N, K, dtype=torch.float32, device=dev)
    ws = _scratch(L.mla_conv2d_wgrad_workspace_bytes(1, M, 1, K, N, 1, 1, 1, 0), dev)
    _lib.check(L.mla_conv2d_wgrad(x_r.data_ptr(), dy_r.data_ptr(), dw.data_ptr(), 1, M, 1, K, N, 1, 1, 1, 0, ws.data_ptr(),
                                  ws.numel(), st), "mla_conv2d_wgrad")
    return dx, dw


class _BlockFn(torch.autograd.Function):
    """x + attention(LN1(x)), then + MLP(LN2(.)) — reference Block.forward, m3ae.py:143-154 — as one node."""

    @staticmethod
    def forward(ctx, x, mask, H, scale, eps1, eps2, g1, b1, wq, bq, wo, bo, g2, b2, w1, c1, w2, c2):
        L = _lib.lib()
        B, S, D = x.shape
        M, Dh = B * S, D // H
        dev = x.device
        x2 = x.reshape(M, D)
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        f16 = BLOCK_BACKWARD == "fp16"
        _, h16, h_r, mu1, rs1 = _ln_fwd(x2, g1, b1, eps1, want_r=not f16)
        qkv = _linear16(h16, wq, bq, None, M, D, 3 * D)
        ao = torch.empty(M, D, dtype=torch.float32, device=dev)
        stats = torch.empty(B, H, S, 2, dtype=torch.float32, device=dev)
        qkv16 = torch.empty(M, 3 * D, dtype=torch.float16, device=dev)
        _lib.check(L.mla_attention_forward(qkv.data_ptr(), mask.data_ptr() if mask is not None else None, ao.data_ptr(),
                                           stats.data_ptr(), qkv16.data_ptr(), B, S, H, Dh, scale, _lib.stream_ptr()),
                   "mla_attention_forward")
        del qkv
        ao16, ao_r = _cast_round(ao, False, want_r=not f16)
        x1 = _linear16(ao16, wo, bo, x2, M, D, D)
        _, h2_16, h2_r, mu2, rs2 = _ln_fwd(x1, g2, b2, eps2, want_r=not f16)
        u = _linear16(h2_16, w1, c1, None, M, D, 4 * D)
        g16, g_r = _cast_round(u, True, want_r=not f16)
        y = _linear16(g16, w2, c2, x1, M, 4 * D, D)
        # the x operands of the four weight gradients: the fp16 copies the forward GEMMs consumed, or TF32-rounded fp32
        ops = (h16, ao16, h2_16, g16) if f16 else (h_r, ao_r, h2_r, g_r)
        ctx.save_for_backward(x2, mask, mu1, rs1, ops[0], qkv16, stats, ao, ops[1], x1, mu2, rs2, ops[2], u, ops[3],
                              g1, wq, wo, g2, w1, w2)
        ctx.f16 = f16
        ctx.cfg = (B, S, D, H, scale)
        return y.view(B, S, D)

    @staticmethod
    def backward(ctx, dy):
        (x2, mask, mu1, rs1, h_r, qkv16, stats, ao, ao_r, x1, mu2, rs2, h2_r, u, g_r, g1, wq, wo, g2, w1, w2) = ctx.saved_tensors
        B, S, D, H, scale = ctx.cfg
        L = _lib.lib()
        M, Dh = B * S, D // H
        dev = dy.device
        dy2 = dy.reshape(M, D)
        if dy2.dtype != torch.float32 or not dy2.is_contiguous():
            dy2 = dy2.float().contiguous()
        if ctx.f16:
            def operand(t, uu):                       # -> (fp16 operand, 1/F), bias gradient
                o16, col, inv = _grad_operand16(t, uu)
                return (o16, inv), col

            def grads(x_op, d_op, w, K, N):
                return _linear_grads16(x_op, d_op[0], d_op[1], w, M, K, N)
        else:
            def operand(t, uu):                       # -> (TF32-rounded operand, None), bias gradient
                r, col = _round_colsum(t, uu)
                return (r, None), col

            def grads(x_op, d_op, w, K, N):
                return _linear_grads(x_op, d_op[0], w, M, K, N)
        # MLP: y = x1 + fc2(gelu(fc1(LN2(x1))))
        d_op, dc2 = operand(dy2, None)
        dg, dw2 = grads(g_r, d_op, w2, 4 * D, D)
        d_op, dc1 = operand(dg, u)
        del dg
        dh2, dw1 = grads(h2_r, d_op, w1, D, 4 * D)
        dx1, dg2, db2 = _ln_bwd(dh2, x1, mu2, rs2, g2, dy2)
        # attention: x1 = x + fc(attn(qkv(LN1(x))))
        d_op, dbo = operand(dx1, None)
        dao, dwo = grads(ao_r, d_op, wo, D, D)
        dqkv = torch.empty(M, 3 * D, dtype=torch.float32, device=dev)
        ws = torch.empty(L.mla_attention_backward_workspace_bytes(B, S, H, Dh), dtype=torch.uint8, device=dev)
        _lib.check(L.mla_attention_backward(qkv16.data_ptr(), mask.data_ptr() if mask is not None else None, ao.data_ptr(),
                                            dao.data_ptr(), stats.data_ptr(), dqkv.data_ptr(), B, S, H, Dh, scale,
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mla_attention_backward")
        d_op, dbq = operand(dqkv, None)
        del dqkv
        dh1, dwq = grads(h_r, d_op, wq, D, 3 * D)
        dx, dg1, db1 = _ln_bwd(dh1, x2, mu1, rs1, g1, dx1)
        return (dx.view(B, S, D), None, None, None, None, None, dg1, db1, dwq, dbq, dwo, dbo, dg2, db2, dw1, dc1, dw2, dc2)


class _LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dimension on the native kernels (the encoder's final layer_norm)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        y, _, _, mean, rstd = _ln_fwd(x2, w, b, eps, want_y=True)
        ctx.save_for_backward(x2, mean, rstd, w)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, w = ctx.saved_tensors
        dy2 = dy.reshape(x2.shape)
        if dy2.dtype != torch.float32 or not dy2.is_contiguous():
            dy2 = dy2.float().contiguous()
        dx, dw, db = _ln_bwd(dy2, x2, mean, rstd, w, None)
        return dx.view(dy.shape), dw, db, None


class NativeLinear(nn.Linear):
    """nn.Linear (same parameters, same default init) whose CUDA forward / backward are the library's GEMM kernels."""

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("mla_b200 m3ae encoders run on CUDA only (no CPU fallback); got %s" % x.device)
        if self.in_features % 64 or self.out_features % 64:
            raise RuntimeError("NativeLinear needs in/out features that are multiples of 64")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features)
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        return _LinearFn.apply(x2, self.weight, self.bias).view(*lead, self.out_features)


class TransformerMLP(nn.Module):                       # m3ae.py:65-83
    def __init__(self, dim, out_dim):
        super().__init__()
        self.fc1 = NativeLinear(dim, 4 * dim)
        self.fc2 = NativeLinear(4 * dim, out_dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))           # exact (erf) GELU, dropout rates are 0


class Attention(nn.Module):                            # m3ae.py:86-125
    def __init__(self, dim, num_heads):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv_linear = NativeLinear(dim, dim * 3, bias=True)
        self.fc = NativeLinear(dim, dim)

    def forward(self, x, padding_mask=None):
        if padding_mask is not None:
            padding_mask = padding_mask.float().contiguous()
        out = _AttentionFn.apply(self.qkv_linear(x), padding_mask, self.num_heads, self.scale)
        return self.fc(out)


class Block(nn.Module):                                # m3ae.py:128-154 (pre-LN)
    def __init__(self, emb_dim, num_heads):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(emb_dim)
        self.attention = Attention(emb_dim, num_heads)
        self.layer_norm2 = nn.LayerNorm(emb_dim)
        self.transformer_mlp = TransformerMLP(emb_dim, emb_dim)

    def forward(self, x, padding_mask=None):
        if FUSED_BLOCK:
            if not x.is_cuda:
                raise RuntimeError("mla_b200 m3ae encoders run on CUDA only (no CPU fallback); got %s" % x.device)
            a, m = self.attention, self.transformer_mlp
            mask = padding_mask.float().contiguous() if padding_mask is not None else None
            return _BlockFn.apply(x, mask, a.num_heads, a.scale, self.layer_norm1.eps, self.layer_norm2.eps,
                                  self.layer_norm1.weight, self.layer_norm1.bias, a.qkv_linear.weight, a.qkv_linear.bias,
                                  a.fc.weight, a.fc.bias, self.layer_norm2.weight, self.layer_norm2.bias,
                                  m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias)
        x = x + self.attention(self.layer_norm1(x), padding_mask)
        return x + self.transformer_mlp(self.layer_norm2(x))


class Transformer(nn.Module):                          # m3ae.py:157-179
    def __init__(self, emb_dim, depth, num_heads):
        super().__init__()
        self.blocks = nn.ModuleList([Block(emb_dim, num_heads) for _ in range(depth)])
        self.layer_norm = nn.LayerNorm(emb_dim)

    def forward(self, x, padding_mask=None):
        for blk in self.blocks:
            x = blk(x, padding_mask)
        if FUSED_BLOCK and x.is_cuda:
            return _LayerNormFn.apply(x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
        return self.layer_norm(x)


def sincos_1d(embed_dim, positions):
    """[len(positions), embed_dim]: sin of pos * 10000^(-2i/D) for the first half, cos for the second (m3ae.py:181-194)."""
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float32) / (embed_dim / 2.0))
    ang = np.einsum("m,d->md", np.asarray(positions, dtype=np.float32).reshape(-1), omega)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_2d(embed_dim, length):
    """[length, embed_dim] for a square grid: half the channels encode one grid axis, half the other (m3ae.py:207-224;
    the reference's meshgrid puts the column index first)."""
    g = int(round(math.sqrt(length)))
    assert g * g == length
    col, row = np.meshgrid(np.arange(g, dtype=np.float32), np.arange(g, dtype=np.float32))
    return np.concatenate([sincos_1d(embed_dim // 2, col), sincos_1d(embed_dim // 2, row)], axis=1)


class MaskedMultimodalAutoencoder(nn.Module):
    """Encoder half of the reference's M3AE (m3ae.py:272-370): only what forward_representation uses."""

    def __init__(self, text_vocab_size, config_updates=None):
        super().__init__()
        cfg = dict(model_type="small", emb_dim=1024, depth=24, num_heads=16)      # get_default_config, m3ae.py:273-297
        cfg.update({k: v for k, v in dict(config_updates or {}).items() if k in cfg})
        if cfg["model_type"] is not None:                                         # a named size overrides the numbers
            cfg["emb_dim"], cfg["depth"], cfg["num_heads"] = _SIZES[cfg["model_type"]]
        assert text_vocab_size > 0
        self.text_vocab_size = text_vocab_size
        self.emb_dim, self.depth, self.num_heads = cfg["emb_dim"], cfg["depth"], cfg["num_heads"]
        D = self.emb_dim
        self.text_embedding = nn.Embedding(text_vocab_size, D)
        self.text_embedding.weight.data.normal_(0.0, 1.0)
        self.image_embedding = NativeLinear(768, D)
        nn.init.xavier_uniform_(self.image_embedding.weight)
        self.encoder_image_type_embedding = nn.Parameter(torch.empty(1, 1, D).normal_(0.02))    # mean 0.02, std 1
        self.encoder_text_type_embedding = nn.Parameter(torch.empty(1, 1, D).normal_(0.02))
        self.cls_token = nn.Parameter(torch.empty(1, 1, D).normal_(0.02))
        self.encoder = Transformer(D, self.depth, self.num_heads)
        self._pos = {}
        self._used = None            # which inputs the last forward_representation read: (image?, text?)

    def hot_parameters(self):
        """The parameters the last forward_representation read, i.e. the ones that receive gradients. The unused input
        branch (text embedding table of an image encoder, image projection of a text encoder, their type embeddings)
        keeps grad None and is skipped by SGD — no weight decay, no momentum — as in the reference, where autograd
        never touches it (main.py:435-440)."""
        if self._used is None:
            return list(self.parameters())
        image, text = self._used
        ps = [self.cls_token] + list(self.encoder.parameters())
        if image:
            ps += list(self.image_embedding.parameters()) + [self.encoder_image_type_embedding]
        if text:
            ps += list(self.text_embedding.parameters()) + [self.encoder_text_type_embedding]
        return ps

    def _pos_embed(self, kind, length, device):
        key = (kind, length, device)
        t = self._pos.get(key)
        if t is None:
            tab = sincos_2d(self.emb_dim, length) if kind == "2d" else sincos_1d(self.emb_dim, np.arange(length))
            t = torch.from_numpy(tab.astype(np.float32))[None].to(device)
            self._pos[key] = t
        return t

    def forward_representation(self, image, text, text_padding_mask, deterministic=False):
        B = image.shape[0] if image is not None else text.shape[0]
        dev = image.device if image is not None else text.device
        self._used = (image is not None, text is not None)
        parts = [self.cls_token.expand(B, 1, self.emb_dim)]
        masks = [torch.zeros(B, 1, dtype=torch.float32, device=dev)]
        if image is not None:
            parts.append(self.image_embedding(image) + self._pos_embed("2d", image.shape[1], dev)
                         + self.encoder_image_type_embedding)
            masks.append(torch.zeros(B, image.shape[1], dtype=torch.float32, device=dev))
        if text is not None:
            parts.append(self.text_embedding(text) + self._pos_embed("1d", text.shape[1], dev)
                         + self.encoder_text_type_embedding)
            masks.append(text_padding_mask.float())
        return self.encoder(torch.cat(parts, dim=1), torch.cat(masks, dim=1))


_N_CLASSES = {"MVSA": 3, "Food101": 101, "CREMAD": 6}


class M3AEClassifier(nn.Module):
    """basic_model.py:127-200. forward(token [B,1,L] int64, padding_mask [B,1,L], visual [B,3,256,256]) -> (a, v), each
    [B, 768]: `a` is the TEXT encoder's token mean, `v` the image encoder's (the reference's naming)."""

    def __init__(self, args, model_config=None, text_vocab_size=30522):
        super().__init__()
        if args.dataset not in _N_CLASSES:
            raise NotImplementedError("Incorrect dataset name {}".format(args.dataset))
        if args.fusion_method != "concat":
            raise NotImplementedError("mla_b200 implements the concat head only")
        if getattr(args, "modulation", "Normal") == "QMF":
            raise NotImplementedError("QMF is outside the MLA hot path (SURVEY.md section 2)")
        model_config = dict(model_config or {"model_type": "base"})               # basic_model.py:163
        # the reference hard-codes a 768-wide head next to its 'base' encoders under --gs_flag (basic_model.py:150) and a
        # 1536-wide concatenated one without it (:153); sized from the encoder here so other model sizes work too
        emb = _SIZES[model_config["model_type"]][0] if model_config.get("model_type") else model_config["emb_dim"]
        self.fusion_module = ConcatFusion(input_dim=emb if args.gs_flag else 2 * emb, output_dim=_N_CLASSES[args.dataset])
        self.mae_a = MaskedMultimodalAutoencoder(text_vocab_size, model_config)
        self.mae_v = MaskedMultimodalAutoencoder(text_vocab_size, model_config)
        # basic_model.py:167-174 loads pretrained encoders from hard-coded placeholder paths ("/path/to/..."); here they
        # are optional arguments, loaded non-strictly like the reference does
        for enc, key in ((self.mae_a, "m3ae_ckpt_audio"), (self.mae_v, "m3ae_ckpt_visual")):
            path = getattr(args, key, None)
            if path:
                enc.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
        self.args = args

    def forward(self, token, padding_mask, visual):
        B, Cc, H, W = visual.shape
        p = 16                                             # 'b c (h p1) (w p2) -> b (h w) (c p1 p2)'
        patches = visual.reshape(B, Cc, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), Cc * p * p)
        a = self.mae_a.forward_representation(None, token.squeeze(1), padding_mask.squeeze(1))
        v = self.mae_v.forward_representation(patches, None, None)
        return a.mean(dim=1), v.mean(dim=1)
