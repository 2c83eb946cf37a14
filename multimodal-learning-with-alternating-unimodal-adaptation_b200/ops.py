"""Tensor-level wrappers over the C ABI (one Python function per entry point).

These are thin: argument checking, workspace sizing, pointer extraction, error mapping.
"""
import ctypes

import torch

from . import _lib

_ws = {}


def _workspace(key, nbytes, device):
    ws = _ws.setdefault((key, device), _lib.Workspace())
    return ws.get(nbytes, device)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mla_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)


def _f32(t, name):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    return t


def gs_project(P, grad_w, alpha, feat=None, feat_sum=None, inv_batch=None, mode=0):
    """In-place P update + gradient projection (utils/utils.py:34-41). See mla_gs_project."""
    L = _lib.lib()
    _need_cuda(P, grad_w, feat, feat_sum)
    if (feat is None) == (feat_sum is None):
        raise RuntimeError("gs_project: pass exactly one of feat / feat_sum")
    D = P.shape[0]
    if P.dim() != 2 or P.shape[1] != D:
        raise RuntimeError("gs_project: P must be (D, D)")
    if feat is not None:
        B = feat.shape[0]
        if feat.shape[1] != D:
            raise RuntimeError("gs_project: feature width %d does not match P (%d)" % (feat.shape[1], D))
    else:
        if inv_batch is None:
            raise RuntimeError("gs_project: feat_sum needs inv_batch")
        B = 1
        if feat_sum.numel() != D:
            raise RuntimeError("gs_project: feat_sum must have D elements")
    if inv_batch is None:
        inv_batch = 1.0 / B
    C = 0 if grad_w is None else grad_w.shape[0]
    if grad_w is not None and grad_w.shape[1] != D:
        raise RuntimeError("gs_project: grad width %d does not match P (%d)" % (grad_w.shape[1], D))
    nbytes = L.mla_gs_project_workspace_bytes(B, D, C)
    if nbytes == 0:
        raise RuntimeError("gs_project: unsupported shape B=%d D=%d C=%d" % (B, D, C))
    ws = _workspace("gs", nbytes, P.device)
    rc = L.mla_gs_project(_lib.ptr(_f32(P, "P")), _lib.ptr(_f32(feat, "feat")), _lib.ptr(_f32(feat_sum, "feat_sum")),
                          float(inv_batch), float(alpha), _lib.ptr(_f32(grad_w, "grad_w")), B, D, C, int(mode),
                          ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "mla_gs_project")


def head_ce(feat, weight, bias, label, grad_scale=None, need_grad=True, need_logits=True, need_dfeat=True,
            out=None):
    """One head turn. Returns dict(logits, loss, dW, db, dfeat, feat_sum); loss is a 1-element tensor."""
    L = _lib.lib()
    _need_cuda(feat, weight, bias, label)
    B, D = feat.shape
    C = weight.shape[0]
    if weight.shape[1] != D:
        raise RuntimeError("head_ce: weight is %s, features are %d wide" % (tuple(weight.shape), D))
    if label.dtype != torch.int64:
        raise RuntimeError("head_ce: label must be int64")
    if label.numel() != B:
        raise RuntimeError("head_ce: %d labels for a batch of %d" % (label.numel(), B))
    if not label.is_contiguous():
        label = label.contiguous()
    if not (feat.is_contiguous() and weight.is_contiguous()):
        raise RuntimeError("head_ce: feat and weight must be contiguous")
    dev = feat.device
    o = out if out is not None else {}

    def buf(name, shape, want=True):
        if not want:
            o[name] = None
            return None
        t = o.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float32, device=dev)
            o[name] = t
        return t

    logits = buf("logits", (B, C), need_logits)
    loss = buf("loss", (1,))
    dW = buf("dW", (C, D), need_grad)
    db = buf("db", (C,), need_grad)
    dfeat = buf("dfeat", (B, D), need_grad and need_dfeat)
    feat_sum = buf("feat_sum", (D,))
    if grad_scale is None:
        grad_scale = 1.0 / B
    nbytes = L.mla_head_ce_workspace_bytes(B, D, C)
    ws = _workspace("head", nbytes, dev)
    rc = L.mla_head_ce(_lib.ptr(_f32(feat, "feat")), _lib.ptr(_f32(weight, "weight")), _lib.ptr(_f32(bias, "bias")),
                       _lib.ptr(label), B, D, C, _lib.ptr(logits), _lib.ptr(loss), _lib.ptr(dW), _lib.ptr(db),
                       _lib.ptr(dfeat), _lib.ptr(feat_sum), float(grad_scale), ws.data_ptr(), ws.numel(),
                       _lib.stream_ptr())
    _lib.check(rc, "mla_head_ce")
    return o


def fuse_eval(logits, label=None, dynamic=True, fixed_w=None, hits=None, num=None, want_fused=True,
              want_argmax=True, want_entropy=False):
    """Fusion + accuracy counters (main.py:65-106, 640-676). Returns (fused, w, argmax[, entropy])."""
    L = _lib.lib()
    _need_cuda(*logits)
    M = len(logits)
    B, C = logits[0].shape
    dev = logits[0].device
    for t in logits:
        if tuple(t.shape) != (B, C):
            raise RuntimeError("fuse_eval: all logit matrices must have the same shape")
        _f32(t, "logits")
    fused = torch.empty((B, C), dtype=torch.float32, device=dev) if want_fused else None
    w = torch.empty((M,), dtype=torch.float32, device=dev)
    ent = torch.empty((M,), dtype=torch.float32, device=dev) if (dynamic and want_entropy) else None
    argmax = torch.empty((M + 1, B), dtype=torch.int32, device=dev) if want_argmax else None
    if (hits is None) != (num is None):
        raise RuntimeError("fuse_eval: hits and num come together")
    if hits is not None:
        if label is None or label.dtype != torch.int64:
            raise RuntimeError("fuse_eval: counting needs int64 labels")
        if hits.dtype != torch.int64 or num.dtype != torch.int64 or tuple(hits.shape) != (M + 1, C) \
                or num.numel() != C:
            raise RuntimeError("fuse_eval: hits must be int64 (M+1, C) and num int64 (C)")
    ptrs = (ctypes.c_void_p * M)(*[_lib.ptr(t) for t in logits])
    fw = None
    if not dynamic:
        if fixed_w is None or len(fixed_w) != M:
            raise RuntimeError("fuse_eval: fixed fusion needs M weights")
        fw = (ctypes.c_float * M)(*[float(x) for x in fixed_w])
    nbytes = L.mla_fuse_eval_workspace_bytes(M, B, C)
    if nbytes == 0:
        raise RuntimeError("fuse_eval: unsupported shape M=%d B=%d C=%d" % (M, B, C))
    ws = _workspace("fuse", nbytes, dev)
    rc = L.mla_fuse_eval(ptrs, M, B, C, 1 if dynamic else 0, fw, _lib.ptr(label) if hits is not None else None,
                         _lib.ptr(fused), _lib.ptr(w), _lib.ptr(ent), _lib.ptr(argmax), _lib.ptr(hits), _lib.ptr(num),
                         ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "mla_fuse_eval")
    if want_entropy:
        return fused, w, argmax, ent
    return fused, w, argmax


# ----------------------------------------------------------------------------------------------
# OGM / OGM-GE (main.py:312-410)
# ----------------------------------------------------------------------------------------------
def ogm_scores(logits, label, out=None):
    """score[m] = sum_b softmax(logits[m])[b][label[b]] (main.py:315-317 / 373-374) as a device tensor of M floats."""
    L = _lib.lib()
    _need_cuda(*logits, label)
    M = len(logits)
    B, C = logits[0].shape
    for t in logits:
        if tuple(t.shape) != (B, C):
            raise RuntimeError("ogm_scores: all logit matrices must have the same shape")
        _f32(t, "logits")
    if label.dtype != torch.int64 or label.numel() != B:
        raise RuntimeError("ogm_scores: label must be int64 with one entry per sample")
    score = out if out is not None else torch.empty(M, dtype=torch.float32, device=label.device)
    ptrs = (ctypes.c_void_p * M)(*[_lib.ptr(t) for t in logits])
    ws = _workspace("ogm", L.mla_ogm_scores_workspace_bytes(M, B), label.device)
    rc = L.mla_ogm_scores(ptrs, M, _lib.ptr(label.contiguous()), B, C, _lib.ptr(score), ws.data_ptr(), ws.numel(),
                          _lib.stream_ptr())
    _lib.check(rc, "mla_ogm_scores")
    return score


def ogm_coeff(score, alpha, out=None):
    """Per-modality gradient coefficients from the (global-batch) scores, on the device (main.py:319-334 / 376-384)."""
    L = _lib.lib()
    _need_cuda(score)
    M = score.numel()
    coeff = out if out is not None else torch.empty(M, dtype=torch.float32, device=score.device)
    rc = L.mla_ogm_coeff(_lib.ptr(_f32(score, "score")), M, float(alpha), _lib.ptr(coeff), _lib.stream_ptr())
    _lib.check(rc, "mla_ogm_coeff")
    return coeff


def ogm_modulate(flat_grad, seg_off, seg_len, max_len, coeff, noise=None, seg_std=None):
    """flat_grad[segments] = flat_grad * coeff (+ noise * seg_std) in one launch (main.py:393-408)."""
    L = _lib.lib()
    _need_cuda(flat_grad, seg_off, seg_len, coeff, noise, seg_std)
    if seg_off.dtype != torch.int64 or seg_len.dtype != torch.int64 or seg_off.numel() != seg_len.numel():
        raise RuntimeError("ogm_modulate: seg_off / seg_len must be int64 device arrays of equal length")
    rc = L.mla_ogm_modulate(_lib.ptr(_f32(flat_grad, "grad")), _lib.ptr(seg_off), _lib.ptr(seg_len), seg_off.numel(),
                            int(max_len), _lib.ptr(_f32(coeff, "coeff")), _lib.ptr(_f32(noise, "noise")),
                            _lib.ptr(_f32(seg_std, "seg_std")), _lib.stream_ptr())
    _lib.check(rc, "mla_ogm_modulate")


# ----------------------------------------------------------------------------------------------
# Encoder convolutions (NHWC activations, [Cout][R][S][Cin] weights)
# ----------------------------------------------------------------------------------------------
def frames_to_batch(src, desc, B, T, size, max_crop_h, mean, std, out=None, status=None, bicubic=False):
    """uint8 HWC RGB frames -> the visual input tensor [B, 3, T, OH, OW] (dataset/dataset.py:123-161, 251-256, 414-421).
    `src` = packed frames (uint8, CUDA), `desc` = int32 CUDA tensor [nframes, 14] (see mla_frames_to_batch), `status` =
    optional int32 CUDA scalar."""
    L = _lib.lib()
    _need_cuda(src, desc, out, status)
    if src.dtype != torch.uint8 or desc.dtype != torch.int32 or desc.dim() != 2 or desc.shape[1] != 14:
        raise RuntimeError("frames_to_batch: src must be uint8 and desc int32 [nframes, 14]")
    OH, OW = (size, size) if isinstance(size, int) else size
    nframes = desc.shape[0]
    if out is None:
        out = torch.empty(B, 3, T, OH, OW, dtype=torch.float32, device=src.device)
    elif tuple(out.shape) != (B, 3, T, OH, OW) or out.dtype != torch.float32:
        raise RuntimeError("frames_to_batch: out must be float32 [B, 3, T, OH, OW]")
    if status is not None and (status.dtype != torch.int32 or status.numel() != 1):
        raise RuntimeError("frames_to_batch: status must be one int32")
    nbytes = L.mla_frames_to_batch_workspace_bytes(nframes, OH, OW, int(max_crop_h))
    if nbytes == 0:
        raise RuntimeError("frames_to_batch: unsupported shape")
    ws = _workspace("frames", nbytes, src.device)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    rc = L.mla_frames_to_batch(_lib.ptr(src), src.numel(), _lib.ptr(desc), nframes, B, T, OH, OW, int(max_crop_h),
                               1 if bicubic else 0, m3, s3,
                               _lib.ptr(out), _lib.ptr(status), ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "mla_frames_to_batch")
    return out


def spec_to_batch(fbank, params, amp, noise, mean, std, skip_norm=False, out=None):
    """Filterbank batch [B, T, F] -> masked / normalised / noised / rolled batch (dataset/dataset.py:281-294, 312-321).
    `params` int32 CUDA [B, 6] = (f0, f1, t0, t1, shift, add_noise), `amp` float32 CUDA [B], `noise` float32 [B, T, F] or None."""
    L = _lib.lib()
    _need_cuda(fbank, params, amp, noise, out)
    if fbank.dim() != 3 or params.dtype != torch.int32 or tuple(params.shape) != (fbank.shape[0], 6) or amp.numel() != fbank.shape[0]:
        raise RuntimeError("spec_to_batch: fbank [B, T, F], params int32 [B, 6], amp [B]")
    if noise is not None and noise.shape != fbank.shape:
        raise RuntimeError("spec_to_batch: noise must have the shape of fbank")
    B, T, F = fbank.shape
    if out is None:
        out = torch.empty_like(fbank)
    rc = L.mla_spec_to_batch(_lib.ptr(_f32(fbank, "fbank")), _lib.ptr(params), _lib.ptr(_f32(amp, "amp")),
                             _lib.ptr(_f32(noise, "noise")), float(mean), float(std), 1 if skip_norm else 0, B, T, F,
                             _lib.ptr(_f32(out, "out")), _lib.stream_ptr())
    _lib.check(rc, "mla_spec_to_batch")
    return out


def _conv_out(x, k, stride, pad):
    return (x + 2 * pad - k) // stride + 1


def conv2d_fprop(x, w, stride, pad, out=None):
    """x [N,H,W,Cin] contiguous; w [Cout,R,S,Cin] contiguous -> y [N,OH,OW,Cout]."""
    L = _lib.lib()
    _need_cuda(x, w)
    N, H, W, Cin = x.shape
    Cout, R, S, _ = w.shape
    OH, OW = _conv_out(H, R, stride, pad), _conv_out(W, S, stride, pad)
    y = out if out is not None else torch.empty((N, OH, OW, Cout), dtype=torch.float32, device=x.device)
    rc = L.mla_conv2d_fprop(_lib.ptr(_f32(x, "x")), _lib.ptr(_f32(w, "w")), _lib.ptr(y), N, H, W, Cin, Cout, R, S,
                            stride, pad, _lib.stream_ptr())
    _lib.check(rc, "mla_conv2d_fprop")
    return y


def conv2d_dgrad(dy, w, x_shape, stride, pad, out=None, accumulate=False):
    """dy [N,OH,OW,Cout], w [Cout,R,S,Cin] -> dx [N,H,W,Cin] (out += if accumulate)."""
    L = _lib.lib()
    _need_cuda(dy, w)
    N, H, W, Cin = x_shape
    Cout, R, S, _ = w.shape
    dx = out if out is not None else torch.empty((N, H, W, Cin), dtype=torch.float32, device=dy.device)
    rc = L.mla_conv2d_dgrad(_lib.ptr(_f32(dy, "dy")), _lib.ptr(_f32(w, "w")), _lib.ptr(dx), N, H, W, Cin, Cout, R, S,
                            stride, pad, 1 if accumulate else 0, _lib.stream_ptr())
    _lib.check(rc, "mla_conv2d_dgrad")
    return dx


def conv2d_wgrad(x, dy, w_shape, stride, pad, out=None):
    """x [N,H,W,Cin], dy [N,OH,OW,Cout] -> dw [Cout,R,S,Cin]."""
    L = _lib.lib()
    _need_cuda(x, dy)
    N, H, W, Cin = x.shape
    Cout, R, S, _ = w_shape
    dw = out if out is not None else torch.empty((Cout, R, S, Cin), dtype=torch.float32, device=x.device)
    nbytes = L.mla_conv2d_wgrad_workspace_bytes(N, H, W, Cin, Cout, R, S, stride, pad)
    if nbytes == 0:
        raise RuntimeError("conv2d_wgrad: unsupported shape")
    ws = _workspace("wgrad", nbytes, x.device)
    rc = L.mla_conv2d_wgrad(_lib.ptr(_f32(x, "x")), _lib.ptr(_f32(dy, "dy")), _lib.ptr(dw), N, H, W, Cin, Cout, R, S,
                            stride, pad, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "mla_conv2d_wgrad")
    return dw
