"""Test-time fusion helpers with the reference's names (main.py:65-106).

`calculate_entropy` softmaxes over dim 0 — the BATCH axis — and sums over everything, so the
"per-sample" gating weights are one scalar per modality per batch. All three run
through the single fusion kernel (csrc/fuse_eval.cu); each returns 0-dim CUDA tensors that
broadcast against [B, C] exactly like the reference's.
"""
from . import ops


def calculate_entropy(output):
    """main.py:65-70."""
    _, _, _, ent = ops.fuse_eval([output.contiguous()], dynamic=True, want_fused=False, want_argmax=False,
                                 want_entropy=True)
    return ent[0]


def _gating(*outs):
    _, w, _ = ops.fuse_eval([o.contiguous() for o in outs], dynamic=True, want_fused=False, want_argmax=False)
    return tuple(w[i] for i in range(len(outs)))


def calculate_gating_weights(encoder_output_1, encoder_output_2):
    """main.py:72-87."""
    return _gating(encoder_output_1, encoder_output_2)


def calculate_gating_weights3(encoder_output_1, encoder_output_2, encoder_output_3):
    """main.py:89-106."""
    return _gating(encoder_output_1, encoder_output_2, encoder_output_3)
