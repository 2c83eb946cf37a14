"""GSPlugin — drop-in for the reference's gradient-projection hook (utils/utils.py:12-41).

Same public surface: `.Pl` (D x D fp32 CUDA tensor, initialised to I), `.exp_count` (int),
`before_update(model, before_batch_input, batch_index, len_dataloader, train_exp_counter)`.
The body (utils.py:34-41) runs as ONE fused sm_100a kernel (csrc/gs_project.cu).

Behaviours kept from the reference, on purpose:
  F1  the hook only acts on a parameter literally named "module.weight"; main.py passes the
      bare nn.Linear (`weight`/`bias`), so AS PUBLISHED it is a no-op. Default here is the
      same. `force_projection=True` also fires on a bare `weight` ("what the paper meant").
  F2  the denominator is ELEMENTWISE: alpha + k_i r_j (mode 0). mode=1 is canonical OWM.
  F3  P is divided by its Frobenius norm after every update; only weight.grad is touched;
      the call with train_exp_counter == 0 is skipped.
Differences, also on purpose:
  F4  Pl is sized from the head's in_features on first use (the reference hard-codes 512).
  -   Pl is updated in place and never joins the autograd graph (the reference leaks a graph).
"""
import torch

from . import ops


class GSPlugin:
    def __init__(self, gs_flag=True, dim=512, device=None, force_projection=False, mode=0):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("GSPlugin needs a CUDA device: the projection runs as an sm_100a kernel "
                                   "and there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.Pl = torch.eye(dim, dtype=torch.float32, device=self.device)      # utils.py:19-20
        self.exp_count = 0                                                     # utils.py:21
        self.force_projection = bool(force_projection)
        self.mode = int(mode)

    # -- helpers ---------------------------------------------------------------------------
    @staticmethod
    def alpha(batch_index, len_dataloader):
        """utils.py:26-27, Python double arithmetic; converted to fp32 at the C boundary."""
        lamda = batch_index / len_dataloader + 1
        return 1.0 * 0.1 ** lamda

    def _gated_weight(self, model):
        """The parameter the reference's name test (utils.py:30-32) selects, or None."""
        for n, w in model.named_parameters():
            if n == "module.weight":
                return w
        if self.force_projection:
            for n, w in model.named_parameters():
                if n == "weight":
                    return w
        return None

    def _ensure_dim(self, D):
        if self.Pl.shape[0] != D:
            if self.exp_count != 0 and not torch.equal(self.Pl, torch.eye(self.Pl.shape[0], device=self.device)):
                raise RuntimeError("GSPlugin: head width changed from %d to %d after updates" % (self.Pl.shape[0], D))
            self.Pl = torch.eye(D, dtype=torch.float32, device=self.device)

    # -- the reference's entry point -------------------------------------------------------
    @torch.no_grad()
    def before_update(self, model, before_batch_input, batch_index, len_dataloader, train_exp_counter,
                      feat_sum=None, inv_batch=None):
        """utils.py:24-41. `feat_sum`/`inv_batch` (keyword-only extension) replace the batch mean
        of `before_batch_input` by an already reduced global sum (data-parallel training)."""
        alpha = self.alpha(batch_index, len_dataloader)
        if train_exp_counter == 0:                                            # utils.py:29
            return
        w = self._gated_weight(model)
        if w is None:
            return
        if w.grad is None:
            raise RuntimeError("GSPlugin.before_update: %s has no gradient" % "weight")
        D = w.shape[1]
        self._ensure_dim(D)
        grad = w.grad.data
        if not grad.is_contiguous():
            grad = grad.contiguous()
            w.grad.data = grad
        if feat_sum is not None:
            ops.gs_project(self.Pl, grad, alpha, feat_sum=feat_sum, inv_batch=inv_batch, mode=self.mode)
        else:
            feat = before_batch_input.detach()
            if feat.dtype != torch.float32 or not feat.is_contiguous():
                feat = feat.float().contiguous()
            ops.gs_project(self.Pl, grad, alpha, feat=feat, mode=self.mode)

    # -- persistence (extension: the reference never saves Pl / exp_count) -------------------
    def state_dict(self):
        return {"Pl": self.Pl.clone(), "exp_count": self.exp_count, "mode": self.mode,
                "force_projection": self.force_projection}

    def load_state_dict(self, sd):
        self.Pl = sd["Pl"].to(self.device, torch.float32).contiguous()
        self.exp_count = int(sd["exp_count"])
        self.mode = int(sd.get("mode", self.mode))
        self.force_projection = bool(sd.get("force_projection", self.force_projection))
