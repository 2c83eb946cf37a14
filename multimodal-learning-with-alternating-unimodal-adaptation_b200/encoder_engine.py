"""Encoder execution engine: runs a ResNet-18 parameter container (backbone.py) entirely on the
kernels of libmla_b200.so — forward AND backward (the reference gets its backward from autograd
over cuDNN, main.py:435).

Data layout: activations NHWC fp32; conv weights [Cout][R][S][Cin], which is the channels_last
memory of the reference's OIHW parameters (converted in place on first use, after the seeded
initialisation so that initial values stay bit-identical to the reference's).

Per encoder and input shape a `ResNetPlan` owns every intermediate buffer (nothing is allocated
in the steady state) and issues the launch sequence:
  stem   im2col (raw NCHW/NCTHW input, frame fold by index arithmetic) -> tcgen05 GEMM -> BN
         statistics -> fused BN+ReLU+MaxPool
  block  conv3x3 (tcgen05 implicit GEMM) -> BN stats -> BN+ReLU -> conv3x3 -> BN stats ->
         [1x1/2 downsample conv -> BN stats] -> fused BN + residual(+its BN) + ReLU
  head   global average pool over (T,h,w)
and the mirrored backward (BN backward = 2-stage reduction + apply, dgrad / wgrad on tcgen05,
residual gradients accumulated in the dgrad epilogue).
"""
import torch

from . import _lib

import os

BACKEND = "native-tcgen05"
USE_F16 = os.environ.get("MLA_F16", "1") != "0"     # 2-byte conv operands (kind::f16): fp16 forward, power-of-two-scaled fp16 backward
USE_GRAPHS = os.environ.get("MLA_GRAPHS", "1") != "0"               # replay the plans' launch sequences as CUDA graphs
_OVERLAP_WGRAD = os.environ.get("MLA_OVERLAP_WGRAD", "1") != "0"    # wgrad kernels on a side stream of the plan
_USE_RELU_MASK = os.environ.get("MLA_RELU_MASK", "1") != "0"      # A/B switch (bitmask vs reading the activation)
_STEM_KP = {1: 64, 3: 160}     # K = 49*Cin padded to a multiple of 32 (tcgen05 k-blocks of 32 tf32)
_STEM_KP16 = {1: 64, 3: 192}   # ... to a multiple of 64 (k-blocks of 64 2-byte elements)
# 2-byte stem: the im2col matrix leaves as ONE fp16 copy (operand of fprop16 and of wgrad16) instead of TF32 fp32, the stem
# GEMMs run at the kind::f16 rate, and BN backward writes dy as scaled fp16 only. Needs USE_F16.
STEM_F16 = True
# im2col-free stem (csrc/stem_s2d.cu): space-to-depth input, sliding-window TMA, fp16 conv output, BN backward with the
# MaxPool / ReLU backward folded in. Needs USE_F16 and STEM_F16; MLA_STEM_S2D=0 goes back to the im2col GEMM.
STEM_S2D = os.environ.get("MLA_STEM_S2D", "1") != "0"


def _p(t):
    return None if t is None else t.data_ptr()


# Optional per-launch timing of the convolution kernels (bench.py's roofline leg): when CONV_TIMING is a
# list, every conv call is bracketed by CUDA events on the launching stream and appended as
# (kind, algorithmic FLOPs, start event, end event). Off (None) in normal operation.
CONV_TIMING = None
GRAPH_LAUNCHES = 0        # kernel launches replayed through CUDA graphs (the library's own counter only sees eager ones)


def _conv_timer_begin():
    if CONV_TIMING is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _conv_timer_end(e0, kind, N, H, W, Cin, Cout, R, stride, pad):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    CONV_TIMING.append((kind, 2.0 * N * OH * OW * Cout * Cin * R * R, e0, e1))


def _chk(rc, what):
    if rc != 0:
        _lib.check(rc, what)


def _krsc(p):
    """Make a conv weight parameter channels_last in place (memory [Cout][R][S][Cin])."""
    if p.dim() == 4 and not p.data.permute(0, 2, 3, 1).is_contiguous():
        p.data = p.data.contiguous(memory_format=torch.channels_last)
    return p


def _flatten_params(net):
    """Re-home every parameter of an encoder into ONE contiguous fp32 buffer (values preserved; conv
    weights become channels_last = [Cout][R][S][Cin]). One launch then produces the TF32-rounded copy
    of all weights that the tensor-core convolutions read (`net._mla_wr`), at the same offsets."""
    flat = net.__dict__.get("_mla_flat")
    params = list(net.parameters())
    if flat is not None and all(p.data.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
                                for p in params):
        return
    dev = params[0].device
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]          # 16-byte aligned slots
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    off = 0
    offsets = {}
    for p, n in zip(params, sizes):
        if p.dim() == 4:
            o, i, r, s = p.shape
            view = torch.as_strided(flat, p.shape, (r * s * i, 1, s * i, i), storage_offset=off)
        else:
            view = torch.as_strided(flat, p.shape, p.data.contiguous().stride(), storage_offset=off)
        view.copy_(p.data)
        p.data = view
        offsets[id(p)] = off
        off += n
    net.__dict__["_mla_flat"] = flat
    net.__dict__["_mla_wr"] = torch.empty_like(flat)
    net.__dict__["_mla_off"] = offsets


def _grad_buffer(p):
    """The tensor backward writes d(loss)/dp into: the pre-attached p.grad (e.g. a view of the flat
    all-reduce bucket) when its memory matches p's, otherwise a fresh tensor with p's layout."""
    g = p.grad
    if g is None or g.stride() != p.stride() or g.dtype != torch.float32:
        g = torch.empty_like(p)
        p.grad = g
    return g


class _BN:
    def __init__(self, bn, dev):
        self.bn = bn
        C = bn.num_features
        self.C = C
        self.mean = torch.empty(C, device=dev)
        self.invstd = torch.empty(C, device=dev)
        self.scale = torch.empty(C, device=dev)
        self.shift = torch.empty(C, device=dev)
        # [F, 1 / F]: the power-of-two scale of this layer's fp16 dy (written by the BN backward, read by dgrad / wgrad)
        self.gscale = torch.ones(2, device=dev)


class ResNetPlan:
    def __init__(self, net, x):
        self.net = net
        dev = x.device
        self.dev = dev
        if net.modality == "visual":
            B, Cin, T, H, W = x.shape
        else:
            B, Cin, H, W = x.shape
            T = 1
        self.B, self.T, self.Cin, self.H, self.W = B, T, Cin, H, W
        self.N = B * T
        self.key = tuple(x.shape)
        self.L = _lib.lib()
        self._pool = {}
        N = self.N
        _flatten_params(net)
        self.flat, self.wr, self.woff = net._mla_flat, net._mla_wr, net._mla_off
        self._params = list(net.parameters())
        # ---- stem
        self.f16 = USE_F16
        self.stem16 = bool(USE_F16 and STEM_F16)
        self.Kp = (_STEM_KP16 if self.stem16 else _STEM_KP)[Cin]
        self.OH0, self.OW0 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        self.PH, self.PW = (self.OH0 + 2 - 3) // 2 + 1, (self.OW0 + 2 - 3) // 2 + 1
        self.M0 = N * self.OH0 * self.OW0
        e = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=dev)   # noqa: E731
        half = lambda *s: torch.empty(s, dtype=torch.float16, device=dev)          # noqa: E731
        self.s2d = bool(self.stem16 and STEM_S2D and Cin <= 4)
        if self.s2d:
            self.col = self.col16 = None
            self.xs16 = torch.empty(self.L.mla_stem_s2d_input_elems(N, H, W), dtype=torch.float16, device=dev)
            self.w2_16 = half(64, 256)
            self.y0_16 = half(N, self.OH0, self.OW0, 64)
            self.stem_tiles = self.L.mla_stem_s2d_tiles(N, H, W)
            self.stem_ws = torch.empty(self.L.mla_stem_s2d_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev)
        elif self.stem16:
            self.col = None
            self.col16 = half(self.M0, self.Kp)
            self.wpad16 = half(64, self.Kp)
        else:
            self.col = e(self.M0, self.Kp)
        if not self.s2d:
            self.wpad = e(64, self.Kp)
            self.dwpad = e(64, self.Kp)
            self.y0 = e(N, self.OH0, self.OW0, 64)
        self.p0 = e(N, self.PH, self.PW, 64)
        self.p0_16 = half(N, self.PH, self.PW, 64) if self.f16 else None
        self.w16 = torch.empty(self.flat.numel(), dtype=torch.float16, device=dev) if self.f16 else None     # fp16 weights
        self.wt16 = torch.empty(self.flat.numel(), dtype=torch.float16, device=dev) if self.f16 else None    # transposed, fp16
        self.idx0 = e(N, self.PH, self.PW, 64, dt=torch.uint8)
        self.bn0 = _BN(net.bn1, dev)
        # ---- residual blocks
        self.blocks = []
        h, w, cin = self.PH, self.PW, 64
        for li in range(1, 5):
            for blk in getattr(net, "layer%d" % li):
                s = blk.stride
                cout = blk.conv1.out_channels
                ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
                d = dict(blk=blk, stride=s, cin=cin, cout=cout, h=h, w=w, ho=ho, wo=wo,
                         y1=e(N, ho, wo, cout), a1=e(N, ho, wo, cout), y2=e(N, ho, wo, cout), out=e(N, ho, wo, cout),
                         # ReLU sign bitmasks of a1 / out (1 bit per element): what BN backward reads instead of them
                         m1=e(N * ho * wo * cout // 32, dt=torch.int32), m2=e(N * ho * wo * cout // 32, dt=torch.int32),
                         # fp16 copies of a1 / out: the operands of the kind::f16 forward convolutions AND weight gradients
                         a1_16=half(N, ho, wo, cout) if self.f16 else None, out_16=half(N, ho, wo, cout) if self.f16 else None,
                         bn1=_BN(blk.bn1, dev), bn2=_BN(blk.bn2, dev), yd=None, bnd=None)
                if blk.downsample is not None:
                    d["yd"] = e(N, ho, wo, cout)
                    d["bnd"] = _BN(blk.downsample[1], dev)
                self.blocks.append(d)
                h, w, cin = ho, wo, cout
        self.rows = T * h * w                   # NHWC rows averaged per sample
        self.C_out = cin
        self.bns = [self.bn0] + [b[k] for b in self.blocks for k in ("bn1", "bn2", "bnd") if b[k] is not None]
        self.nbt = [b.bn.num_batches_tracked for b in self.bns if b.bn.num_batches_tracked is not None]
        nb = max(self.L.mla_bn_workspace_bytes(self.M0, 64),
                 max(self.L.mla_bn_workspace_bytes(N * b["ho"] * b["wo"], b["cout"]) for b in self.blocks))
        tl = self.L.mla_conv2d_fprop_stat_tiles
        tl16 = self.L.mla_conv2d_fprop16_stat_tiles
        self.stat_part = torch.empty(max([tl(N, self.OH0, self.OW0, 1, 1, 1, 0) * 2 * 64,
                                          (self.stem_tiles if self.s2d else 0) * 2 * 64] +
                                         [max(tl(N, b["h"], b["w"], 3, 3, b["stride"], 1), tl(N, b["ho"], b["wo"], 3, 3, 1, 1),
                                              tl(N, b["h"], b["w"], 1, 1, b["stride"], 0),
                                              tl16(N, b["h"], b["w"], b["cin"], b["cout"], 3, 3, b["stride"], 1),
                                              tl16(N, b["ho"], b["wo"], b["cout"], b["cout"], 3, 3, 1, 1)) * 2 * b["cout"]
                                          for b in self.blocks]),
                                     dtype=torch.float32, device=dev)          # per-tile BN partial sums (fprop epilogue)
        self.bn_ws = torch.zeros(nb, dtype=torch.uint8, device=dev)      # ticket counters start at 0 (mla_b200.h)
        nw = 0 if self.s2d else (self.L.mla_conv2d_wgrad16_workspace_bytes if self.stem16
                                 else self.L.mla_conv2d_wgrad_workspace_bytes)(N, self.OH0, self.OW0, self.Kp, 64, 1, 1, 1, 0)
        for b in self.blocks:
            nw = max(nw, self.L.mla_conv2d_wgrad_workspace_bytes(N, b["h"], b["w"], b["cin"], b["cout"], 3, 3,
                                                                  b["stride"], 1),
                     self.L.mla_conv2d_wgrad_workspace_bytes(N, b["ho"], b["wo"], b["cout"], b["cout"], 3, 3, 1, 1))
        if self.f16:
            for b in self.blocks:
                nw = max(nw, self.L.mla_conv2d_wgrad16_workspace_bytes(N, b["h"], b["w"], b["cin"], b["cout"], 3, 3, b["stride"], 1),
                         self.L.mla_conv2d_wgrad16_workspace_bytes(N, b["ho"], b["wo"], b["cout"], b["cout"], 3, 3, 1, 1),
                         self.L.mla_conv2d_wgrad16_workspace_bytes(N, b["h"], b["w"], b["cin"], b["cout"], 1, 1, b["stride"], 0))
        self.wg_ws = torch.empty(max(nw, 256), dtype=torch.uint8, device=dev)
        if self.f16:      # one launch transposes every BasicBlock filter into its fp16 [Cin][R][S][Cout] copy (dgrad16 operand)
            import numpy as np
            seg, tile0 = [], 0
            for b in self.blocks:
                blk = b["blk"]
                ws_ = [(blk.conv1.weight, b["cout"], 9, b["cin"]), (blk.conv2.weight, b["cout"], 9, b["cout"])]
                if b["yd"] is not None:
                    ws_.append((blk.downsample[0].weight, b["cout"], 1, b["cin"]))
                for w, co, rs, ci in ws_:
                    seg.append((self.woff[id(w)], co, rs, ci, tile0))
                    tile0 += rs * ((co + 31) // 32) * ((ci + 31) // 32)
            tab = np.array(seg, dtype=np.dtype([("off", "<i8"), ("co", "<i4"), ("rs", "<i4"), ("ci", "<i4"), ("t0", "<i4")]))
            self.tr_table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
            self.tr_nseg, self.tr_tiles = len(seg), tile0
        self.trained_forward = False
        self.serial = 0                 # number of the last training forward (activations are single-buffered)
        self._param_ptrs = [p.data_ptr() for p in self._params]
        self.feat_static = torch.empty(self.B, self.C_out, dtype=torch.float32, device=dev)    # pooled feature (graph output)
        self.dfeat_static = torch.empty(self.B, self.C_out, dtype=torch.float32, device=dev)   # its gradient (graph input)
        self._graphs, self._warm = {}, {}
        self.wstream = None            # side stream of the weight-gradient kernels (created on first backward)
        self._slot_events = {}         # (buffer slot, shape) -> event of the last wgrad that read it

    # ------------------------------------------------------------------ small helpers
    def tmp(self, slot, shape):
        k = (slot, tuple(shape))
        t = self._pool.get(k)
        if t is None:
            t = torch.empty(shape, dtype=torch.float32, device=self.dev)
            self._pool[k] = t
        return t

    def _wptr(self, w):
        """Device pointer of the TF32-rounded copy of a parameter (plain buffers: their own pointer)."""
        off = self.woff.get(id(w))
        return w.data_ptr() if off is None else self.wr.data_ptr() + 4 * off

    def _w16ptr(self, w):
        return self.w16.data_ptr() + 2 * self.woff[id(w)]

    def _wt16ptr(self, w):
        return self.wt16.data_ptr() + 2 * self.woff[id(w)]

    def _conv_bn16(self, x16, w, y, N, H, W, Cin, Cout, R, stride, pad, b, training, st):
        """fp16 x fp16 forward convolution (+ BN partial sums in training), then the BN coefficients."""
        t = _conv_timer_begin()
        _chk(self.L.mla_conv2d_fprop16(_p(x16), self._w16ptr(w), _p(y), N, H, W, Cin, Cout, R, R, stride, pad,
                                       _p(self.stat_part) if training else None, st), "mla_conv2d_fprop16")
        _conv_timer_end(t, "fprop16", N, H, W, Cin, Cout, R, stride, pad)
        if not training:
            self._bn_coeffs(y, 0, b, False, st)
            return
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
        M = N * OH * OW
        bn = b.bn
        ntiles = self.L.mla_conv2d_fprop16_stat_tiles(N, H, W, Cin, Cout, R, R, stride, pad)
        _chk(self.L.mla_bn_stats_from_partials(_p(self.stat_part), ntiles, M, b.C, _p(bn.weight), _p(bn.bias),
                                               _p(bn.running_mean), _p(bn.running_var), float(bn.momentum), float(bn.eps),
                                               _p(b.mean), _p(b.invstd), _p(b.scale), _p(b.shift), _p(self.bn_ws),
                                               self.bn_ws.numel(), st), "mla_bn_stats_from_partials")

    def _dgrad16(self, dy16, gscale, w, dx, N, H, W, Cin, Cout, R, stride, pad, acc, st):
        """dx (+)= (1 / F) * dgrad(fp16 dy * F, fp16 transposed filter); gscale = the producing BN's [F, 1 / F]."""
        t = _conv_timer_begin()
        _chk(self.L.mla_conv2d_dgrad16_f16(_p(dy16), self._wt16ptr(w), gscale.data_ptr() + 4, _p(dx), N, H, W, Cin, Cout, R, R,
                                           stride, pad, 1 if acc else 0, st), "mla_conv2d_dgrad16_f16")
        _conv_timer_end(t, "dgrad16", N, H, W, Cin, Cout, R, stride, pad)

    def _conv(self, x, w, y, N, H, W, Cin, Cout, R, stride, pad, st, k_alg=None):
        t = _conv_timer_begin()
        _chk(self.L.mla_conv2d_fprop(_p(x), self._wptr(w), _p(y), N, H, W, Cin, Cout, R, R, stride, pad, st), "mla_conv2d_fprop")
        _conv_timer_end(t, "fprop", N, H, W, Cin if k_alg is None else k_alg, Cout, R, stride, pad)

    def _conv_bn(self, x, w, y, N, H, W, Cin, Cout, R, stride, pad, b, training, st, k_alg=None):
        """Convolution + BatchNorm coefficients of its output. Training: the conv epilogue emits per-tile
        (sum, sum^2) partials from its fp32 accumulators and one small launch finalises them (no statistics
        pass over y). Eval: plain conv + coefficients from the running statistics."""
        if not training:
            self._conv(x, w, y, N, H, W, Cin, Cout, R, stride, pad, st, k_alg=k_alg)
            self._bn_coeffs(y, 0, b, False, st)
            return
        t = _conv_timer_begin()
        _chk(self.L.mla_conv2d_fprop_bnstats(_p(x), self._wptr(w), _p(y), N, H, W, Cin, Cout, R, R, stride, pad,
                                              _p(self.stat_part), st), "mla_conv2d_fprop_bnstats")
        _conv_timer_end(t, "fprop", N, H, W, Cin if k_alg is None else k_alg, Cout, R, stride, pad)
        OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
        M = N * OH * OW
        bn = b.bn
        ntiles = self.L.mla_conv2d_fprop_stat_tiles(N, H, W, R, R, stride, pad)
        _chk(self.L.mla_bn_stats_from_partials(_p(self.stat_part), ntiles, M, b.C, _p(bn.weight), _p(bn.bias),
                                               _p(bn.running_mean), _p(bn.running_var), float(bn.momentum), float(bn.eps),
                                               _p(b.mean), _p(b.invstd), _p(b.scale), _p(b.shift), _p(self.bn_ws),
                                               self.bn_ws.numel(), st), "mla_bn_stats_from_partials")

    def _dgrad(self, dy, w, dx, N, H, W, Cin, Cout, R, stride, pad, acc, st):
        t = _conv_timer_begin()
        _chk(self.L.mla_conv2d_dgrad(_p(dy), self._wptr(w), _p(dx), N, H, W, Cin, Cout, R, R, stride, pad, 1 if acc else 0, st),
             "mla_conv2d_dgrad")
        _conv_timer_end(t, "dgrad", N, H, W, Cin, Cout, R, stride, pad)

    def _wgrad(self, x, dy, dw, N, H, W, Cin, Cout, R, stride, pad, st, k_alg=None):
        t = _conv_timer_begin()
        _chk(self.L.mla_conv2d_wgrad(_p(x), _p(dy), _p(dw), N, H, W, Cin, Cout, R, R, stride, pad, _p(self.wg_ws),
                                     self.wg_ws.numel(), st), "mla_conv2d_wgrad")
        _conv_timer_end(t, "wgrad", N, H, W, Cin if k_alg is None else k_alg, Cout, R, stride, pad)

    def _wgrad16(self, x16, dy16, gscale, dw, N, H, W, Cin, Cout, R, stride, pad, st, k_alg=None):
        t = _conv_timer_begin()
        _chk(self.L.mla_conv2d_wgrad16_f16(_p(x16), _p(dy16), gscale.data_ptr() + 4, _p(dw), N, H, W, Cin, Cout, R, R, stride,
                                           pad, _p(self.wg_ws), self.wg_ws.numel(), st), "mla_conv2d_wgrad16_f16")
        _conv_timer_end(t, "wgrad16", N, H, W, Cin if k_alg is None else k_alg, Cout, R, stride, pad)

    def _bn_coeffs(self, y, M, b, training, st):
        bn = b.bn
        if training:
            _chk(self.L.mla_bn_train_stats(_p(y), M, b.C, _p(bn.weight), _p(bn.bias), _p(bn.running_mean),
                                           _p(bn.running_var), float(bn.momentum), float(bn.eps), _p(b.mean),
                                           _p(b.invstd), _p(b.scale), _p(b.shift), _p(self.bn_ws), self.bn_ws.numel(), st),
                 "mla_bn_train_stats")
        else:
            _chk(self.L.mla_bn_eval_coeffs(_p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var),
                                           float(bn.eps), b.C, _p(b.scale), _p(b.shift), st), "mla_bn_eval_coeffs")

    def _bn_bwd(self, dz, z, y, b, M, dy, g_out, st, mask=None, dy16=None):
        bn = b.bn
        dg, db = _grad_buffer(bn.weight), _grad_buffer(bn.bias)
        if dy16 is not None:           # 2-byte path: dy leaves as fp16 * F (power of two from the data), F -> b.gscale
            _chk(self.L.mla_bn_backward_f16(_p(dz), _p(mask), _p(y), _p(b.mean), _p(b.invstd), _p(bn.weight), M, b.C, _p(dg),
                                            _p(db), _p(dy16), _p(g_out), _p(b.gscale), _p(self.bn_ws), self.bn_ws.numel(), st),
                 "mla_bn_backward_f16")
            return
        _chk(self.L.mla_bn_backward_ex(_p(dz), _p(z), _p(mask), _p(y), _p(b.mean), _p(b.invstd), _p(bn.weight), M, b.C,
                                       _p(dg), _p(db), _p(dy), _p(dy16), _p(g_out), _p(self.bn_ws), self.bn_ws.numel(), st),
             "mla_bn_backward")

    def tmp16(self, slot, shape):
        k = (slot, tuple(shape), "f16")
        t = self._pool.get(k)
        if t is None:
            t = torch.empty(shape, dtype=torch.float16, device=self.dev)
            self._pool[k] = t
        return t

    # ------------------------------------------------------------------------ CUDA graphs
    def _run(self, key, fn):
        """Run `fn` (a fixed launch sequence over buffers this plan owns) — eagerly the first time (function
        attributes, lazy allocations), then captured once into a CUDA graph and replayed: one launch instead of
        ~80-250, so the host never delays the streams (the two encoders and the weight gradients run on concurrent
        streams and every microsecond of enqueue latency on one of them is exposed)."""
        if not USE_GRAPHS or CONV_TIMING is not None:
            return fn()
        g = self._graphs.get(key)
        if g is None:
            if self._warm.get(key, 0) < 1:
                self._warm[key] = 1
                return fn()
            try:
                g = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    fn()
                g.mla_nodes = _lib.launch_count() - n0      # kernels of ours inside the graph
            except Exception as e:                      # capture unsupported for some reason: stay eager, loudly
                import warnings
                warnings.warn("mla_b200: CUDA graph capture of %s failed (%s); running eagerly" % (key, e))
                self._graphs[key] = False
                torch.cuda.synchronize()
                return fn()
            self._graphs[key] = g
        if g is False:
            return fn()
        g.replay()
        global GRAPH_LAUNCHES
        GRAPH_LAUNCHES += g.mla_nodes

    # ------------------------------------------------------------------------ forward
    def forward(self, x, training):
        L, N, st = self.L, self.N, _lib.stream_ptr()
        net = self.net
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        HW = self.H * self.W
        if net.modality == "visual":       # [B,3,T,H,W]: n = b*T + t
            sB, sT, sC = 3 * self.T * HW, HW, self.T * HW
        else:                              # [B,1,H,W]
            sB, sT, sC = self.Cin * HW, 0, HW
        # the only launch that reads the caller's buffer (its address changes from batch to batch): outside the graph
        if self.s2d:
            _chk(L.mla_stem_s2d_pack(_p(x), _p(self.xs16), N, self.T, sB, sT, sC, self.Cin, self.H, self.W, st),
                 "mla_stem_s2d_pack")
        elif self.stem16:
            _chk(L.mla_stem_im2col16(_p(x), _p(self.col16), None, N, self.T, sB, sT, sC,
                                     self.Cin, self.H, self.W, 7, 7, 2, 3, self.Kp, st), "mla_stem_im2col16")
        else:
            _chk(L.mla_stem_im2col(_p(x), _p(self.col), N, self.T, sB, sT, sC, self.Cin, self.H, self.W, 7, 7, 2, 3, self.Kp,
                                   st), "mla_stem_im2col")
        self._run(("fwd", bool(training)), lambda: self._forward_body(training))
        self.trained_forward = training
        if training:
            self.serial += 1
        return self.feat_static.clone()

    def _forward_body(self, training):
        L, N, st = self.L, self.N, _lib.stream_ptr()
        net = self.net
        K = 49 * self.Cin
        if not (self.f16 and self.s2d):       # the TF32-rounded weights are read by the kind::tf32 kernels only
            _chk(L.mla_round_tf32(_p(self.flat), _p(self.wr), self.flat.numel(), st), "mla_round_tf32")
        if self.s2d:
            # the stem as a 4x4 / stride 1 convolution over 2x2 space-to-depth cells, straight from the packed input
            _chk(L.mla_stem_s2d_weights(_p(net.conv1.weight), _p(self.w2_16), self.Cin, st), "mla_stem_s2d_weights")
            t = _conv_timer_begin()
            _chk(L.mla_stem_s2d_fprop(_p(self.xs16), _p(self.w2_16), _p(self.y0_16), N, self.H, self.W,
                                      _p(self.stat_part) if training else None, st), "mla_stem_s2d_fprop")
            _conv_timer_end(t, "fprop16", N, self.OH0, self.OW0, K, 64, 1, 1, 0)
            if training:
                bn = self.bn0.bn
                _chk(L.mla_bn_stats_from_partials(_p(self.stat_part), self.stem_tiles, self.M0, 64, _p(bn.weight),
                                                  _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), float(bn.momentum),
                                                  float(bn.eps), _p(self.bn0.mean), _p(self.bn0.invstd), _p(self.bn0.scale),
                                                  _p(self.bn0.shift), _p(self.bn_ws), self.bn_ws.numel(), st),
                     "mla_bn_stats_from_partials")
            else:
                self._bn_coeffs(None, 0, self.bn0, False, st)
        elif self.stem16:
            # the stem as an fp16 GEMM over the fp16 im2col matrix (same 10-bit operand mantissa as the TF32 path)
            _chk(L.mla_pad_rows(_p(net.conv1.weight), _p(self.wpad), 64, K, self.Kp, 0, st), "mla_pad_rows")
            _chk(L.mla_cast16(_p(self.wpad), _p(self.wpad16), self.wpad.numel(), 0, st), "mla_cast16")
            t = _conv_timer_begin()
            _chk(L.mla_conv2d_fprop16(_p(self.col16), _p(self.wpad16), _p(self.y0), N, self.OH0, self.OW0, self.Kp, 64, 1, 1, 1, 0,
                                      _p(self.stat_part) if training else None, st), "mla_conv2d_fprop16")
            _conv_timer_end(t, "fprop16", N, self.OH0, self.OW0, K, 64, 1, 1, 0)
            if training:
                bn = self.bn0.bn
                _chk(L.mla_bn_stats_from_partials(_p(self.stat_part), (self.M0 + 127) // 128, self.M0, 64, _p(bn.weight),
                                                  _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), float(bn.momentum),
                                                  float(bn.eps), _p(self.bn0.mean), _p(self.bn0.invstd), _p(self.bn0.scale),
                                                  _p(self.bn0.shift), _p(self.bn_ws), self.bn_ws.numel(), st),
                     "mla_bn_stats_from_partials")
            else:
                self._bn_coeffs(self.y0, 0, self.bn0, False, st)
        else:
            _chk(L.mla_pad_rows(self._wptr(net.conv1.weight), _p(self.wpad), 64, K, self.Kp, 0, st), "mla_pad_rows")
            self._conv_bn(self.col, self.wpad, self.y0, N, self.OH0, self.OW0, self.Kp, 64, 1, 1, 0, self.bn0, training, st,
                          k_alg=K)
        if self.s2d:
            _chk(L.mla_bn_relu_maxpool16(_p(self.y0_16), _p(self.bn0.scale), _p(self.bn0.shift), _p(self.p0), _p(self.p0_16),
                                         _p(self.idx0), N, self.OH0, self.OW0, 64, st), "mla_bn_relu_maxpool16")
        else:
            _chk(L.mla_bn_relu_maxpool_ex(_p(self.y0), _p(self.bn0.scale), _p(self.bn0.shift), _p(self.p0), _p(self.p0_16),
                                          None, _p(self.idx0), N, self.OH0, self.OW0, 64, st),
                 "mla_bn_relu_maxpool")
        f16 = self.f16
        if f16:      # fp16 copy of every parameter (same offsets as the flat buffer): the B operand of fprop16
            _chk(L.mla_cast16(_p(self.flat), _p(self.w16), self.flat.numel(), 0, st), "mla_cast16")
        xin, xin16 = self.p0, self.p0_16
        for b in self.blocks:
            blk, s, cin, cout = b["blk"], b["stride"], b["cin"], b["cout"]
            M = N * b["ho"] * b["wo"]
            if f16:
                self._conv_bn16(xin16, blk.conv1.weight, b["y1"], N, b["h"], b["w"], cin, cout, 3, s, 1, b["bn1"], training, st)
            else:
                self._conv_bn(xin, blk.conv1.weight, b["y1"], N, b["h"], b["w"], cin, cout, 3, s, 1, b["bn1"], training, st)
            mk1, mk2 = (_p(b["m1"]), _p(b["m2"])) if training else (None, None)
            # a1 in fp32 is only read by the TF32 wgrad / the unmasked BN backward: not written on the 2-byte path
            a1_32 = None if (f16 and _USE_RELU_MASK) else b["a1"]
            _chk(L.mla_bn_apply_ex(_p(b["y1"]), _p(b["bn1"].scale), _p(b["bn1"].shift), None, None, None, 1, _p(a1_32),
                                   mk1, _p(b["a1_16"]), None, M, cout, st), "mla_bn_apply")
            if f16:
                self._conv_bn16(b["a1_16"], blk.conv2.weight, b["y2"], N, b["ho"], b["wo"], cout, cout, 3, 1, 1, b["bn2"],
                                training, st)
            else:
                self._conv_bn(b["a1"], blk.conv2.weight, b["y2"], N, b["ho"], b["wo"], cout, cout, 3, 1, 1, b["bn2"],
                              training, st)
            if b["yd"] is not None:
                if f16:
                    self._conv_bn16(xin16, blk.downsample[0].weight, b["yd"], N, b["h"], b["w"], cin, cout, 1, s, 0, b["bnd"],
                                    training, st)
                else:
                    self._conv_bn(xin, blk.downsample[0].weight, b["yd"], N, b["h"], b["w"], cin, cout, 1, s, 0, b["bnd"],
                                  training, st)
                _chk(L.mla_bn_apply_ex(_p(b["y2"]), _p(b["bn2"].scale), _p(b["bn2"].shift), _p(b["yd"]),
                                       _p(b["bnd"].scale), _p(b["bnd"].shift), 1, _p(b["out"]), mk2, _p(b["out_16"]),
                                       None, M, cout, st), "mla_bn_apply")
            else:
                _chk(L.mla_bn_apply_ex(_p(b["y2"]), _p(b["bn2"].scale), _p(b["bn2"].shift), _p(xin), None, None, 1,
                                       _p(b["out"]), mk2, _p(b["out_16"]), None, M, cout, st),
                     "mla_bn_apply")
            xin, xin16 = b["out"], b["out_16"]
        _chk(L.mla_avgpool_forward(_p(xin), _p(self.feat_static), self.B, self.rows, self.C_out, st), "mla_avgpool_forward")
        if training and self.nbt:
            torch._foreach_add_(self.nbt, 1)

    # ----------------------------------------------------------------------- backward
    def tail_bucket_offset(self):
        """Element offset, inside the encoder's flat parameter / gradient buffer, of the first layer4 parameter: everything
        from there on (75 % of a ResNet-18's parameters) has its gradient complete after backward segment 0."""
        return self.bucket_offsets()[-1]

    def bucket_offsets(self):
        """(first layer2 parameter, first layer4 parameter) as element offsets into the flat buffer: the gradient ranges
        [off4, n), [off2, off4), [0, off2) are complete after backward segments 0, 1, 2."""
        firsts = [next(iter(self.net.layer2.parameters())), next(iter(self.net.layer4.parameters()))]
        offs, off = [None, None], 0
        for q in self._params:
            for i, f in enumerate(firsts):
                if q is f:
                    offs[i] = off
            off += q.numel()
        if None in offs:
            raise RuntimeError("layer2 / layer4 parameters not found in the encoder's parameter list")
        return tuple(offs)

    def backward(self, dfeat, serial=None, on_segment=None):
        """Native backward. `on_segment(k)` (optional) makes the pass run as three replayed launch sequences — k = 0: pooling
        + layer4 (called when every layer4 gradient is complete on the current stream), k = 1: layer3 + layer2, k = 2: layer1
        + stem — so that a data-parallel caller can start the all-reduce of each bucket while the rest of the backward still
        runs; only the small layer1 + stem bucket (0.16 M of 11.2 M parameters) is exposed at the end."""
        if not self.trained_forward:
            raise RuntimeError("encoder backward needs a training-mode forward on the same plan")
        if serial is not None and serial != self.serial:
            # a plan owns ONE set of activation buffers per input shape: a later forward of the same shape overwrote the
            # ones this backward pass would differentiate through
            raise RuntimeError("encoder backward for forward #%d, but forward #%d of the same input shape has overwritten "
                               "its activations: run backward before the next forward" % (serial, self.serial))
        self.dfeat_static.copy_(dfeat)
        params = self._params
        if params[0].grad is not None and params[-1].grad is not None:
            # gradient buffers pre-attached by the caller (train_epoch: views of the flat all-reduce bucket): their
            # addresses are stable, so the launch sequence can be replayed as a graph keyed on them
            for p in params:
                _grad_buffer(p)
            if len(self._graphs) > 12:
                self._graphs.clear(); self._warm.clear()
            key = (params[0].grad.data_ptr(), params[-1].grad.data_ptr())
            if on_segment is None:
                self._run(("bwd",) + key, self._backward_body)
            else:
                for k in range(3):
                    self._run(("bwd%d" % k,) + key, lambda k=k: self._backward_body(k))
                    on_segment(k)
        else:
            self._backward_body()                        # fresh gradient tensors every call (plain autograd use): eager
            if on_segment is not None:
                for k in range(3):
                    on_segment(k)
        self.trained_forward = False

    def _backward_body(self, seg=None):
        """seg None: the whole pass; 0: transposed filters + pooling + layer4 (its weight gradients joined at the end);
        1: layer3 + layer2, 2: layer1 + stem, each continuing from the gradient the previous segment left in
        `self._seg_dout`."""
        L, N, st = self.L, self.N, _lib.stream_ptr()
        net = self.net
        dfeat = self.dfeat_static
        first, last = seg in (None, 0), seg in (None, 2)
        nblk = len(self.blocks)
        split = nblk - len(net.layer4)                   # blocks [split, nblk) belong to layer4
        self._slot_events.clear()                        # only events of THIS pass order its buffer reuse
        # Weight gradients leave the critical path: every wgrad (and its split-K reduction) runs on the plan's own side
        # stream as soon as its dy exists, concurrently with the BN-backward / dgrad chain that continues on `cur`.
        # A dy buffer is only rewritten after the wgrad that read it has finished (events per buffer slot).
        cur = torch.cuda.current_stream()
        if self.wstream is None:
            self.wstream = torch.cuda.Stream()
        wsm = self.wstream if _OVERLAP_WGRAD else cur
        wst = wsm.cuda_stream
        wsm.wait_stream(cur)
        events = self._slot_events

        def buf(slot, shape):                       # a buffer about to be overwritten on `cur`
            ev = events.pop((slot, tuple(shape)), None)
            if ev is not None:
                cur.wait_event(ev)
            return self.tmp(slot, shape)

        def buf16(slot, shape):                     # same for the fp16 gradient buffers (dgrad16 / wgrad16 operands)
            ev = events.pop((slot, tuple(shape)), None)
            if ev is not None:
                cur.wait_event(ev)
            return self.tmp16(slot, shape)

        def wgrad_async(x, dy, slot, dw, *geom, k_alg=None, two_byte=False, gscale=None):
            if wsm is not cur:
                ready = torch.cuda.Event()
                ready.record(cur)
                wsm.wait_event(ready)
            with torch.cuda.stream(wsm):
                if two_byte:
                    self._wgrad16(x, dy, gscale, dw, *geom, wst, k_alg=k_alg)
                else:
                    self._wgrad(x, dy, dw, *geom, wst, k_alg=k_alg)
            if wsm is not cur:
                done = torch.cuda.Event()
                done.record(wsm)
                events[(slot, tuple(dy.shape))] = done

        f16 = self.f16
        if first:
            if f16:  # transposed fp16 filters: the K-major B operand of dgrad16 (weights are those of this step's forward)
                _chk(L.mla_filter_transpose16_batch(_p(self.flat), _p(self.wt16), _p(self.tr_table), self.tr_nseg,
                                                    self.tr_tiles, 0, st), "mla_filter_transpose16_batch")
            dout = self.tmp("dXa", self.blocks[-1]["out"].shape)
            _chk(L.mla_avgpool_backward(_p(dfeat), _p(dout), self.B, self.rows, self.C_out, st), "mla_avgpool_backward")
        else:
            dout = self._seg_dout
        split2 = len(net.layer1)                         # blocks [0, split2) belong to layer1
        hi, lo = {None: (nblk - 1, 0), 0: (nblk - 1, split), 1: (split - 1, split2), 2: (split2 - 1, 0)}[seg]
        for i in range(hi, lo - 1, -1):
            b = self.blocks[i]
            blk, s, cin, cout = b["blk"], b["stride"], b["cin"], b["cout"]
            xin = self.blocks[i - 1]["out"] if i > 0 else self.p0
            M = N * b["ho"] * b["wo"]
            shp = b["out"].shape
            par = i & 1
            g = self.tmp("g%d" % par, shp)
            xin_h = (self.blocks[i - 1]["out_16"] if i > 0 else self.p0_16) if f16 else None
            z2, mk2 = (None, b["m2"]) if _USE_RELU_MASK else (b["out"], None)
            z1, mk1 = (None, b["m1"]) if _USE_RELU_MASK else (b["a1"], None)
            da1 = self.tmp("da", shp)
            if f16:
                # 2-byte path: the gradients exist as scaled fp16 only; wgrad16 reads (x fp16, dy fp16), dgrad16 (dy fp16, w^T fp16)
                dy2 = buf16("dy2h_%d" % par, shp)
                self._bn_bwd(dout, z2, b["y2"], b["bn2"], M, None, g, st, mask=mk2, dy16=dy2)
                wgrad_async(b["a1_16"], dy2, "dy2h_%d" % par, _grad_buffer(blk.conv2.weight), N, b["ho"], b["wo"], cout, cout, 3,
                            1, 1, two_byte=True, gscale=b["bn2"].gscale)
                self._dgrad16(dy2, b["bn2"].gscale, blk.conv2.weight, da1, N, b["ho"], b["wo"], cout, cout, 3, 1, 1, False, st)
                dy1 = buf16("dy1h_%d" % par, shp)
                self._bn_bwd(da1, z1, b["y1"], b["bn1"], M, None, None, st, mask=mk1, dy16=dy1)
                wgrad_async(xin_h, dy1, "dy1h_%d" % par, _grad_buffer(blk.conv1.weight), N, b["h"], b["w"], cin, cout, 3, s, 1,
                            two_byte=True, gscale=b["bn1"].gscale)
            else:
                dy2 = buf("dy2_%d" % par, shp)
                self._bn_bwd(dout, z2, b["y2"], b["bn2"], M, dy2, g, st, mask=mk2)
                wgrad_async(b["a1"], dy2, "dy2_%d" % par, _grad_buffer(blk.conv2.weight), N, b["ho"], b["wo"], cout, cout, 3, 1, 1)
                self._dgrad(dy2, blk.conv2.weight, da1, N, b["ho"], b["wo"], cout, cout, 3, 1, 1, False, st)
                dy1 = buf("dy1_%d" % par, shp)
                self._bn_bwd(da1, z1, b["y1"], b["bn1"], M, dy1, None, st, mask=mk1)
                wgrad_async(xin, dy1, "dy1_%d" % par, _grad_buffer(blk.conv1.weight), N, b["h"], b["w"], cin, cout, 3, s, 1)
            if f16:
                def dg(dy, w, dx, *geom_acc, gs=None):
                    self._dgrad16(dy, gs, w, dx, *geom_acc)
            else:
                def dg(dy, w, dx, *geom_acc, gs=None):
                    self._dgrad(dy, w, dx, *geom_acc)
            if b["yd"] is not None:
                if f16:
                    dyd = buf16("dydh", shp)
                    self._bn_bwd(g, None, b["yd"], b["bnd"], M, None, None, st, dy16=dyd)
                    wgrad_async(xin_h, dyd, "dydh", _grad_buffer(blk.downsample[0].weight), N, b["h"], b["w"], cin, cout, 1, s, 0,
                                two_byte=True, gscale=b["bnd"].gscale)
                else:
                    dyd = buf("dyd", shp)
                    self._bn_bwd(g, None, b["yd"], b["bnd"], M, dyd, None, st)
                    wgrad_async(xin, dyd, "dyd", _grad_buffer(blk.downsample[0].weight), N, b["h"], b["w"], cin, cout, 1, s, 0)
                dx = self.tmp("dXa", xin.shape)         # xin.shape != out.shape here, so never aliases dout
                # the 3x3 dgrad touches every pixel of dx and goes first; the 1x1/2 shortcut then ADDS into the
                # one output parity class it reaches (its other classes are skipped, not zero-filled)
                dg(dy1, blk.conv1.weight, dx, N, b["h"], b["w"], cin, cout, 3, s, 1, False, st, gs=b["bn1"].gscale)
                dg(dyd, blk.downsample[0].weight, dx, N, b["h"], b["w"], cin, cout, 1, s, 0, True, st, gs=b["bnd"].gscale)
            else:
                dx = g                                  # identity shortcut: dX starts as the masked gradient
                dg(dy1, blk.conv1.weight, dx, N, b["h"], b["w"], cin, cout, 3, s, 1, True, st, gs=b["bn1"].gscale)
            dout = dx
        if not last:
            self._seg_dout = dout
            cur.wait_stream(wsm)                        # every layer4 gradient is complete on `cur`
            return
        if self.s2d:
            # stem: BN backward with the MaxPool / ReLU backward gathered inside its two passes (the dense gradient of the
            # pre-pool activation never exists), then the weight gradient straight from the packed input
            bn = self.bn0.bn
            dy0 = buf16("dy0", self.y0_16.shape)
            _chk(L.mla_pool_bn_backward_f16(_p(dout), _p(self.idx0), _p(self.y0_16), _p(self.bn0.mean), _p(self.bn0.invstd),
                                            _p(bn.weight), N, self.OH0, self.OW0, 64, _p(_grad_buffer(bn.weight)),
                                            _p(_grad_buffer(bn.bias)), _p(dy0), _p(self.bn0.gscale), _p(self.bn_ws),
                                            self.bn_ws.numel(), st), "mla_pool_bn_backward_f16")
            dw0 = _grad_buffer(net.conv1.weight)
            if wsm is not cur:
                ready = torch.cuda.Event()
                ready.record(cur)
                wsm.wait_event(ready)
            with torch.cuda.stream(wsm):
                t = _conv_timer_begin()
                _chk(L.mla_stem_s2d_wgrad(_p(self.xs16), _p(dy0), self.bn0.gscale.data_ptr() + 4, _p(dw0), N, self.H, self.W,
                                          self.Cin, _p(self.stem_ws), self.stem_ws.numel(), wst), "mla_stem_s2d_wgrad")
                _conv_timer_end(t, "wgrad16", N, self.OH0, self.OW0, 49 * self.Cin, 64, 1, 1, 0)
            cur.wait_stream(wsm)
            return
        # stem: maxpool+relu backward, BN backward, weight gradient (no dgrad: the input needs none)
        g0 = self.tmp("g0", self.y0.shape)
        _chk(L.mla_maxpool_relu_backward(_p(dout), _p(self.p0), _p(self.idx0), _p(g0), N, self.OH0, self.OW0, 64, st),
             "mla_maxpool_relu_backward")
        if self.stem16:
            dy0 = buf16("dy0", self.y0.shape)            # scaled fp16 only: nothing reads the stem's dy in fp32
            self._bn_bwd(g0, None, self.y0, self.bn0, self.M0, None, None, st, dy16=dy0)
            wgrad_async(self.col16, dy0, "dy0", self.dwpad, N, self.OH0, self.OW0, self.Kp, 64, 1, 1, 0, k_alg=49 * self.Cin,
                        two_byte=True, gscale=self.bn0.gscale)
        else:
            dy0 = buf("dy0", self.y0.shape)
            self._bn_bwd(g0, None, self.y0, self.bn0, self.M0, dy0, None, st)
            wgrad_async(self.col, dy0, "dy0", self.dwpad, N, self.OH0, self.OW0, self.Kp, 64, 1, 1, 0, k_alg=49 * self.Cin)
        with torch.cuda.stream(wsm):
            _chk(L.mla_pad_rows(_p(self.dwpad), _p(_grad_buffer(net.conv1.weight)), 64, 49 * self.Cin, self.Kp, 1, wst),
                 "mla_pad_rows")
        cur.wait_stream(wsm)                            # every parameter gradient is complete on `cur`


class _EncoderFn(torch.autograd.Function):
    """Bridges a ResNetPlan into autograd: forward returns the pooled feature; backward runs the
    plan's native backward, which writes parameter gradients in place (p.grad)."""

    @staticmethod
    def forward(ctx, plan, x, anchor):
        ctx.plan = plan
        out = plan.forward(x, True)
        ctx.serial = plan.serial
        return out

    @staticmethod
    def backward(ctx, dfeat):
        # the native backward OVERWRITES its gradient buffers; autograd semantics are accumulation, so gradients that were
        # already there (a second backward without zero_grad, another loss term) are added back afterwards
        plan = ctx.plan
        old = [(p, p.grad.clone()) for p in plan._params if p.grad is not None]
        plan.backward(dfeat, ctx.serial)
        for p, g in old:
            p.grad.add_(g)
        return None, None, None


def _plan(net, x):
    plans = net.__dict__.setdefault("_mla_plans", {})
    key = (tuple(x.shape), x.device)
    p = plans.get(key)
    if p is not None and p._param_ptrs != [q.data_ptr() for q in p._params]:
        # a parameter was re-homed after the plan captured its address (p.data = ..., load_state_dict(assign=True), a
        # dtype / device round trip): the plan's flat buffer, offsets and CUDA graphs point at stale memory
        torch.cuda.synchronize(x.device)
        plans.clear()
        p = None
    if p is None:
        if len(plans) >= 4:
            torch.cuda.synchronize(x.device)     # buffers of the dropped plans may still be in use on side streams
            plans.clear()
        p = ResNetPlan(net, x)
        plans[key] = p
    return p


def resnet_pooled(net, x):
    """[B,1,H,W] / [B,3,T,H,W] -> pooled feature [B,512] (backbone + adaptive_avg_pool + flatten)."""
    if not x.is_cuda:
        raise RuntimeError("mla_b200 encoders run on CUDA only (no CPU fallback); got %s" % x.device)
    plan = _plan(net, x)
    if net.training and torch.is_grad_enabled():
        out = _EncoderFn.apply(plan, x, net.conv1.weight)
        out._mla_plan = plan          # lets train_epoch run the native backward directly, on a stream of its choice
        out._mla_serial = plan.serial
        return out
    with torch.no_grad():
        return plan.forward(x, net.training)


def resnet_feature_map(net, x):
    """Layer4 feature map in the reference's NCHW layout [N,512,h,w] (API completeness; no grad)."""
    plan = _plan(net, x)
    with torch.no_grad():
        plan.forward(x, net.training)
    return plan.blocks[-1]["out"].permute(0, 3, 1, 2)
