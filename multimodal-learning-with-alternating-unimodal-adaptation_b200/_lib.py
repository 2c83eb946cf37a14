"""ctypes binding of libmla_b200.so (the C ABI declared in include/mla_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised. torch is used only to own device memory and streams; every pointer handed to the
library is `tensor.data_ptr()`.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmla_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "mla_b200.h")

_c_int, _c_float, _c_size_t, _c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
_c_ll = ctypes.c_longlong

_SIGNATURES = {
    "mla_abi_version": (_c_int, []),
    "mla_error_string": (ctypes.c_char_p, [_c_int]),
    "mla_device_sm_count": (_c_int, []),
    "mla_launch_count": (ctypes.c_uint64, []),
    "mla_gs_project_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "mla_gs_project": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_float, _c_float, _c_void_p, _c_int, _c_int,
                                _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "mla_head_ce_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "mla_head_ce": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p,
                             _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_float, _c_void_p, _c_size_t,
                             _c_void_p]),
    "mla_fuse_eval_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "mla_fuse_eval": (_c_int, [ctypes.POINTER(_c_void_p), _c_int, _c_int, _c_int, _c_int,
                               ctypes.POINTER(_c_float), _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                               _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "mla_conv2d_fprop": (_c_int, [_c_void_p] * 3 + [_c_int] * 9 + [_c_void_p]),
    "mla_conv2d_fprop_stat_tiles": (_c_int, [_c_int] * 7),
    "mla_conv2d_fprop16_stat_tiles": (_c_int, [_c_int] * 9),
    "mla_conv2d_fprop_bnstats": (_c_int, [_c_void_p] * 3 + [_c_int] * 9 + [_c_void_p, _c_void_p]),
    "mla_bn_stats_from_partials": (_c_int, [_c_void_p, _c_int, _c_ll, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                            _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                            _c_size_t, _c_void_p]),
    "mla_conv2d_fprop16": (_c_int, [_c_void_p] * 3 + [_c_int] * 9 + [_c_void_p, _c_void_p]),
    "mla_conv2d_dgrad16": (_c_int, [_c_void_p] * 3 + [_c_int] * 10 + [_c_void_p]),
    "mla_conv2d_wgrad16_workspace_bytes": (_c_size_t, [_c_int] * 9),
    "mla_conv2d_wgrad16": (_c_int, [_c_void_p] * 3 + [_c_int] * 9 + [_c_void_p, _c_size_t, _c_void_p]),
    "mla_cast16": (_c_int, [_c_void_p, _c_void_p, _c_ll, _c_int, _c_void_p]),
    "mla_filter_transpose16": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "mla_conv2d_dgrad": (_c_int, [_c_void_p] * 3 + [_c_int] * 10 + [_c_void_p]),
    "mla_conv2d_wgrad_workspace_bytes": (_c_size_t, [_c_int] * 9),
    "mla_conv2d_wgrad": (_c_int, [_c_void_p] * 3 + [_c_int] * 9 + [_c_void_p, _c_size_t, _c_void_p]),
    "mla_attention_forward": (_c_int, [_c_void_p] * 5 + [_c_int] * 4 + [_c_float, _c_void_p]),
    "mla_attention_backward_workspace_bytes": (_c_size_t, [_c_int] * 4),
    "mla_attention_backward": (_c_int, [_c_void_p] * 6 + [_c_int] * 4 + [_c_float, _c_void_p, _c_size_t, _c_void_p]),
    "mla_linear_forward16": (_c_int, [_c_void_p] * 5 + [_c_int] * 3 + [_c_void_p]),
    "mla_linear_dgrad": (_c_int, [_c_void_p] * 3 + [_c_int] * 3 + [_c_void_p]),
    "mla_grad_operand16": (_c_int, [_c_void_p] * 5 + [_c_ll, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "mla_linear_dgrad16": (_c_int, [_c_void_p] * 4 + [_c_int] * 3 + [_c_void_p]),
    "mla_linear_wgrad16": (_c_int, [_c_void_p] * 4 + [_c_int] * 3 + [_c_void_p, _c_size_t, _c_void_p]),
    "mla_layernorm_forward": (_c_int, [_c_void_p] * 3 + [_c_float, _c_ll, _c_int] + [_c_void_p] * 6),
    "mla_layernorm_backward_workspace_bytes": (_c_size_t, [_c_ll, _c_int]),
    "mla_layernorm_backward": (_c_int, [_c_void_p] * 6 + [_c_ll, _c_int] + [_c_void_p] * 4 + [_c_size_t, _c_void_p]),
    "mla_cast_round": (_c_int, [_c_void_p] * 3 + [_c_ll, _c_int, _c_void_p]),
    "mla_round_colsum_workspace_bytes": (_c_size_t, [_c_ll, _c_int]),
    "mla_round_colsum": (_c_int, [_c_void_p] * 4 + [_c_ll, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "mla_stem_im2col": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_ll, _c_ll, _c_ll] + [_c_int] * 8 + [_c_void_p]),
    "mla_stem_im2col16": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_ll, _c_ll, _c_ll] + [_c_int] * 8 + [_c_void_p]),
    "mla_round_tf32": (_c_int, [_c_void_p, _c_void_p, _c_ll, _c_void_p]),
    "mla_pad_rows": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "mla_bn_workspace_bytes": (_c_size_t, [_c_ll, _c_int]),
    "mla_bn_train_stats": (_c_int, [_c_void_p, _c_ll, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_float,
                                    _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t,
                                    _c_void_p]),
    "mla_bn_eval_coeffs": (_c_int, [_c_void_p] * 4 + [_c_float, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "mla_bn_apply": (_c_int, [_c_void_p] * 6 + [_c_int, _c_void_p, _c_ll, _c_int, _c_void_p]),
    "mla_bn_apply_mask": (_c_int, [_c_void_p] * 6 + [_c_int, _c_void_p, _c_void_p, _c_ll, _c_int, _c_void_p]),
    "mla_bn_backward_mask": (_c_int, [_c_void_p] * 6 + [_c_ll, _c_int] + [_c_void_p] * 5 + [_c_size_t, _c_void_p]),
    "mla_bn_apply_ex": (_c_int, [_c_void_p] * 6 + [_c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int,
                                _c_void_p]),
    "mla_bn_backward_ex": (_c_int, [_c_void_p] * 7 + [_c_ll, _c_int] + [_c_void_p] * 6 + [_c_size_t, _c_void_p]),
    "mla_bn_relu_maxpool_ex": (_c_int, [_c_void_p] * 7 + [_c_int] * 4 + [_c_void_p]),
    "mla_bn_backward": (_c_int, [_c_void_p] * 6 + [_c_ll, _c_int] + [_c_void_p] * 5 + [_c_size_t, _c_void_p]),
    "mla_bn_relu_maxpool": (_c_int, [_c_void_p] * 5 + [_c_int] * 4 + [_c_void_p]),
    "mla_maxpool_relu_backward": (_c_int, [_c_void_p] * 4 + [_c_int] * 4 + [_c_void_p]),
    "mla_avgpool_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "mla_avgpool_backward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "mla_bn_backward_f16": (_c_int, [_c_void_p] * 6 + [_c_ll, _c_int] + [_c_void_p] * 6 + [_c_size_t, _c_void_p]),
    "mla_conv2d_dgrad16_f16": (_c_int, [_c_void_p] * 4 + [_c_int] * 10 + [_c_void_p]),
    "mla_conv2d_wgrad16_f16": (_c_int, [_c_void_p] * 4 + [_c_int] * 9 + [_c_void_p, _c_size_t, _c_void_p]),
    "mla_filter_transpose16_batch": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "mla_stem_s2d_input_elems": (_c_ll, [_c_int] * 3),
    "mla_stem_s2d_tiles": (_c_int, [_c_int] * 3),
    "mla_stem_s2d_pack": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_ll, _c_ll, _c_ll, _c_int, _c_int, _c_int, _c_void_p]),
    "mla_stem_s2d_weights": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p]),
    "mla_stem_s2d_fprop": (_c_int, [_c_void_p] * 3 + [_c_int] * 3 + [_c_void_p, _c_void_p]),
    "mla_stem_s2d_wgrad_workspace_bytes": (_c_size_t, []),
    "mla_stem_s2d_wgrad": (_c_int, [_c_void_p] * 4 + [_c_int] * 4 + [_c_void_p, _c_size_t, _c_void_p]),
    "mla_bn_relu_maxpool16": (_c_int, [_c_void_p] * 6 + [_c_int] * 4 + [_c_void_p]),
    "mla_pool_bn_backward_f16": (_c_int, [_c_void_p] * 6 + [_c_int] * 4 + [_c_void_p] * 5 + [_c_size_t, _c_void_p]),
    "mla_ogm_scores_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "mla_ogm_scores": (_c_int, [ctypes.POINTER(_c_void_p), _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p,
                                _c_size_t, _c_void_p]),
    "mla_ogm_coeff": (_c_int, [_c_void_p, _c_int, _c_float, _c_void_p, _c_void_p]),
    "mla_ogm_modulate": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_ll, _c_void_p, _c_void_p, _c_void_p,
                                  _c_void_p]),
    "mla_spec_to_batch": (_c_int, [_c_void_p] * 4 + [_c_float, _c_float] + [_c_int] * 4 + [_c_void_p, _c_void_p]),
    "mla_frames_to_batch_workspace_bytes": (_c_size_t, [_c_int] * 4),
    "mla_frames_to_batch": (_c_int, [_c_void_p, _c_ll, _c_void_p] + [_c_int] * 7 + [ctypes.POINTER(_c_float)] * 2 +
                            [_c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
}

_lib = None


def declared_symbols():
    """Every function name declared with MLA_API in include/mla_b200.h."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    return sorted(set(re.findall(r"MLA_API\s+[\w\s\*]+?\b(mla_\w+)\s*\(", text)))


def lib():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmla_b200.so is not built (%s). Run `python __graft_entry__.py build` — there is no "
            "CPU or PyTorch fallback for the MLA hot path." % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    if handle.mla_abi_version() != 1:
        raise RuntimeError("libmla_b200.so ABI version mismatch")
    _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().mla_error_string(code)
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", code))


def ptr(t):
    """Device pointer of a tensor (None -> NULL). The tensor must be CUDA and contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mla_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("mla_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class Workspace:
    """Grow-only device scratch owned by the caller side (the library never allocates)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self.buf


def launch_count():
    return int(lib().mla_launch_count())
