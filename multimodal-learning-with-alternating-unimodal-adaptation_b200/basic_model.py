"""AVClassifier (reference models/basic_model.py:14-77).

forward(audio [B,1,H,W], visual [B,3,T,H,W]) -> (a, v), each [B,512], under --gs_flag (512-wide shared head), or
(a, v, out) with the 1024-wide concatenated head without it (joint training, basic_model.py:73-75); features require
grad in training. Sub-module names (`fusion_module.fc_out`, `audio_net`, `visual_net`) and their creation order follow
the reference so state dicts and seeded initialisation are identical. Only the concat head is in scope (SURVEY.md §2:
QMF / film / gated / sum are out of scope and raise).
"""
import os

import torch
import torch.nn as nn

from .backbone import resnet18
from .fusion_modules import ConcatFusion

_N_CLASSES = {"CREMAD": 6}
OVERLAP_ENCODERS = os.environ.get("MLA_OVERLAP", "1") != "0"     # audio / visual encoder work on two CUDA streams


class AVClassifier(nn.Module):
    def __init__(self, args):
        super().__init__()
        if args.dataset not in _N_CLASSES:
            raise NotImplementedError("Incorrect dataset name {}".format(args.dataset))
        n_classes = _N_CLASSES[args.dataset]
        if args.fusion_method != "concat":
            raise NotImplementedError("mla_b200 implements the concat head of the --gs_flag path only "
                                      "(got fusion_method={})".format(args.fusion_method))
        if getattr(args, "modulation", "Normal") == "QMF":
            raise NotImplementedError("QMF is outside the MLA --gs_flag path (SURVEY.md §2)")
        # basic_model.py:31-34: 512-wide shared head under gs_flag, 1024 (concat) otherwise
        self.fusion_module = ConcatFusion(input_dim=512 if args.gs_flag else 1024, output_dim=n_classes)
        self.audio_net = resnet18(modality="audio")
        self.visual_net = resnet18(modality="visual")
        self.args = args

    def _streams(self, device):
        st = self.__dict__.get("_mla_streams")
        if st is None:
            st = (torch.cuda.Stream(device), torch.cuda.Stream(device))
            self.__dict__["_mla_streams"] = st
        return st

    def forward_streams(self, audio, visual):
        """Both encoders launched on their own CUDA streams, NOT joined: returns [(a, stream_a), (v, stream_v)].
        The caller waits on a modality's stream when it needs that feature — train_epoch runs the whole audio turn
        (head, GS, audio backward) while the longer visual forward is still in flight. `forward` is this + a join."""
        cur = torch.cuda.current_stream(audio.device)
        sa, sv = self._streams(audio.device)
        sa.wait_stream(cur)                     # inputs and the latest parameter update are ready
        sv.wait_stream(cur)
        with torch.cuda.stream(sv):
            v = self.visual_net.pooled(visual)      # backbone + adaptive_avg_pool3d over (T,H,W) + flatten
        with torch.cuda.stream(sa):
            a = self.audio_net.pooled(audio)        # backbone + adaptive_avg_pool2d + flatten
        return [(a, sa), (v, sv)]

    def features(self, audio, visual):
        """(a, v) without the head (what train_epoch / valid consume in both modes)."""
        return self._features(audio, visual)

    def forward(self, audio, visual):
        a, v = self._features(audio, visual)
        if not self.args.gs_flag:
            return self.fusion_module(a, v)          # basic_model.py:73-75: (a, v, out)
        return a, v

    def _features(self, audio, visual):
        if audio.is_cuda and OVERLAP_ENCODERS:
            cur = torch.cuda.current_stream(audio.device)
            (a, sa), (v, sv) = self.forward_streams(audio, visual)
            for t, s in ((a, sa), (v, sv)):
                cur.wait_stream(s)
                t.record_stream(cur)
        else:
            a = self.audio_net.pooled(audio)
            v = self.visual_net.pooled(visual)
        return a, v
