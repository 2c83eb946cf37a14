"""train_epoch / valid — the reference's L4 drivers for the --gs_flag path (main.py:127-484,
486-679), same signatures and return values, B200-native underneath:

  * per modality turn the head forward+loss+backward is ONE fused call (head kernel), the GS
    hook is one fused kernel, the encoder backward gets dfeat directly;
  * no per-step `.item()` syncs: epoch losses accumulate on the device and are read once;
  * evaluation does fusion, argmax and per-class counters in one kernel per batch — the
    reference's per-sample numpy loop (~8 D2H syncs per sample, main.py:659-676) is gone;
  * multi-GPU is one process per GPU with two NCCL all-reduces per turn (dist.py).
Out of scope (raises): the joint-training branch without --gs_flag (main.py:165-418).
"""
import torch
import torch.nn as nn

from . import basic_model
from . import dist as mdist
from . import ops
from .fusion_modules import head_turn

N_CLASSES = {"MVSA": 3, "CREMAD": 6, "Food101": 101, "IEMOCAP": 4}


class ModuleHolder(nn.Module):
    """Stand-in for nn.DataParallel's `.module` indirection (main.py:732): keeps
    `model.module.fusion_module.fc_out` and the `module.`-prefixed state-dict keys of the
    reference's checkpoints, without DataParallel's per-step replicate/scatter/gather."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def _unpack(args, data_packet, device):
    """Batch tuple layouts of the reference's datasets (main.py:143-162)."""
    nb = dict(non_blocking=True)
    if args.lorb == "m3ae":
        if args.modal3:
            token, padding_mask, image, spec, label, idx = data_packet
            return (token.to(device, **nb), padding_mask.to(device, **nb), image.to(device, **nb),
                    spec.to(device, **nb)), label.to(device, **nb)
        token, padding_mask, image, label, idx = data_packet
        return (token.to(device, **nb), padding_mask.to(device, **nb), image.to(device, **nb)), label.to(device, **nb)
    spec, image, label = data_packet[0], data_packet[1], data_packet[2]
    spec, image = spec.to(device, **nb), image.to(device, **nb)
    if getattr(args, "clip", False):
        return (spec, image), label.to(device, **nb)
    return (spec.unsqueeze(1).float(), image.float()), label.to(device, **nb)


def _batches_on_device(args, dataloader, device):
    """Yields (inputs, label) on `device`. Host batches (pinned) are copied on a side stream ONE BATCH AHEAD into
    two alternating sets of preallocated device buffers, so the H2D transfer of batch i+1 overlaps the compute of
    batch i and nothing is allocated per step (the reference copies synchronously at the top of every iteration,
    main.py:160-162)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        for pkt in dataloader:
            yield _unpack(args, pkt, device)
        return
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    it = iter(dataloader)
    bufs = [None, None]
    count = [0]

    def load():
        try:
            pkt = next(it)
        except StopIteration:
            return None
        if all((not torch.is_tensor(t)) or t.is_cuda for t in pkt):
            return _unpack(args, pkt, device), None                      # already resident: nothing to overlap
        k = count[0] & 1
        count[0] += 1
        if bufs[k] is None or any(torch.is_tensor(t) and (b.shape != t.shape or b.dtype != t.dtype)
                                  for t, b in zip(pkt, bufs[k])):
            bufs[k] = [torch.empty(t.shape, dtype=t.dtype, device=dev) if torch.is_tensor(t) else t for t in pkt]
        # buffer set k was last read by the step before the one now being enqueued: everything on `main` so far
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for t, b in zip(pkt, bufs[k]):
                if torch.is_tensor(t):
                    b.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
        return _unpack(args, tuple(bufs[k]), device), ev

    nxt = load()
    while nxt is not None:
        (inputs, label), ev = nxt
        if ev is not None:
            main.wait_event(ev)
        nxt = load()
        yield inputs, label


class _TurnState:
    """Per-model buffers reused across steps: packed head buffer and flat encoder grads."""

    def __init__(self, net):
        fc = net.fusion_module.fc_out
        C, D = fc.weight.shape
        dev = fc.weight.device
        # [dW | db | sum_b feat], every section 16-byte aligned (the kernels use float4 accesses)
        o_db = (C * D + 3) // 4 * 4
        o_fs = (o_db + C + 3) // 4 * 4
        self.packed = torch.zeros(o_fs + D, dtype=torch.float32, device=dev)
        self.head_out = {"dW": self.packed[:C * D].view(C, D), "db": self.packed[o_db:o_db + C],
                         "feat_sum": self.packed[o_fs:o_fs + D]}
        self.encoders = encoder_param_groups(net)
        self.flat = [mdist.FlatGrads(g) for g in self.encoders]
        self._dfeat = {}
        self._side = {}

    def dfeat_buffer(self, m, feat):
        """d(loss)/d(feature) of modality m: its own buffer, because the encoder backward that reads it may still be
        running on a side stream when the next modality's head turn writes its dfeat."""
        t = self._dfeat.get(m)
        if t is None or t.shape != feat.shape or t.device != feat.device:
            t = torch.empty_like(feat)
            self._dfeat[m] = t
        return t

    def side_stream(self, m):
        s = self._side.get(m)
        if s is None:
            s = torch.cuda.Stream()
            self._side[m] = s
        return s


def encoder_param_groups(net):
    """Parameters of each modality's encoder, in the fixed turn order a -> v -> t (main.py:432-466)."""
    names = []
    for cand in (("audio_net", "visual_net"), ("mae_a", "mae_v", "mae_t")):
        if all(hasattr(net, n) for n in cand[:2]):
            names = [n for n in cand if hasattr(net, n)]
            break
    if not names:
        raise RuntimeError("model has no known encoders (audio_net/visual_net or mae_a/mae_v[/mae_t])")
    # an encoder may name the parameters its forward really reads (CAVMAEFT carries an unused visual branch): the others
    # keep grad None and the optimiser skips them, as in the reference
    return [[p for p in getattr(getattr(net, n), "hot_parameters", getattr(net, n).parameters)()] for n in names]


def train_epoch(args, epoch, model, device, dataloader, optimizer, scheduler,
                gs_plugin=None, writer=None, gs_flag=False, av_alpha=0.5,
                txt_history=None, img_history=None, audio_history=None):
    """main.py:127-484. Returns (loss, loss_a, loss_v[, loss_t]) as Python floats."""
    if not gs_flag:
        raise NotImplementedError("mla_b200 implements MLA's alternating step (--gs_flag) only; the joint-training "
                                  "branch (main.py:165-418) is out of scope")
    net = _unwrap(model)
    model.train()
    print("Start training ... ")
    fc = net.fusion_module.fc_out
    world = mdist.world_size()
    len_dataloader = len(dataloader)
    n_mod = len(encoder_param_groups(net))
    acc = torch.zeros(1 + n_mod, dtype=torch.float64, device=device)      # _loss, _loss_a, _loss_v[, _loss_t]
    # the reference reads the losses back every step with .item() (main.py:472-476), stalling the stream; here an
    # opt-in pinned log receives them by non-blocking copies (args.step_loss_log = True), read after the epoch
    step_log = None
    if getattr(args, "step_loss_log", False) and torch.device(device).type == "cuda":
        step_log = torch.empty(len_dataloader, 1 + n_mod, dtype=torch.float64).pin_memory()

    for batch_step, (inputs, label) in enumerate(_batches_on_device(args, dataloader, device)):
        optimizer.zero_grad()                                              # main.py:164
        if basic_model.OVERLAP_ENCODERS and hasattr(net, "forward_streams") and inputs[0].is_cuda:
            pairs = net.forward_streams(*inputs)                           # main.py:421-431, one stream per encoder
            feats, feat_streams = [f for f, _ in pairs], [s for _, s in pairs]
        else:
            feats, feat_streams = model(*inputs), None                     # main.py:421-431
        st = getattr(net, "_mla_turn_state", None)
        if st is None:       # after the first forward: the engine has fixed the parameter memory layouts by now
            st = _TurnState(net)
            net._mla_turn_state = st
        if len(feats) != n_mod:
            raise RuntimeError("model returned %d features for %d encoders" % (len(feats), n_mod))
        B = feats[0].shape[0]
        inv_global = 1.0 / (B * world)
        losses = []
        pending = []
        for m, feat in enumerate(feats):                                   # a -> v -> (t)
            if feat_streams is not None:           # this turn starts as soon as ITS encoder's forward has finished
                torch.cuda.current_stream().wait_stream(feat_streams[m])
                feat.record_stream(torch.cuda.current_stream())
            fdet = feat.detach()
            st.head_out["dfeat"] = st.dfeat_buffer(m, fdet)                # one dfeat buffer per modality (see below)
            o = head_turn(fc, fdet, label, grad_scale=inv_global, out=st.head_out)   # main.py:432-435 (head part)
            if world > 1:                                                  # the small head all-reduce
                mdist.allreduce_sum_(st.packed)
            plan = getattr(feat, "_mla_plan", None)
            if plan is not None:
                # main.py:435 (encoder part), native backward launched directly. Every encoder but the last runs its
                # backward (and gradient all-reduce) on its own side stream, concurrently with the next turns: this
                # turn's optimizer.step() then only sees the head; the encoder's own SGD update follows when its
                # gradients are complete (same arithmetic: SGD treats parameters independently, and the next turns
                # never read this encoder's parameters).
                deferred = basic_model.OVERLAP_ENCODERS and m < n_mod - 1
                stream = st.side_stream(m) if deferred else torch.cuda.current_stream()
                if deferred:
                    stream.wait_stream(torch.cuda.current_stream())        # dfeat is ready
                with torch.cuda.stream(stream):
                    st.flat[m].attach()
                    plan.backward(o["dfeat"])
                    if world > 1 and not deferred:                         # encoder-gradient all-reduce
                        mdist.allreduce_sum_(st.flat[m].flat)
                if deferred:
                    # its all-reduce is issued with the deferred update below: collectives of one communicator run in
                    # issue order, so issuing it here would park the NEXT turn's small head all-reduce behind this
                    # whole backward pass and serialise the two encoders again
                    st.flat[m].detach()
                    pending.append((m, stream))
            else:                                  # autograd encoders (m3ae): gradients ACCUMULATE into the views
                st.flat[m].attach(zero=True)
                feat.backward(o["dfeat"])
                if world > 1:
                    mdist.allreduce_sum_(st.flat[m].flat)
            gs_plugin.before_update(fc, fdet, batch_step, len_dataloader, gs_plugin.exp_count,
                                    feat_sum=o["feat_sum"], inv_batch=inv_global)     # main.py:437-438
            optimizer.step()                                               # main.py:439
            optimizer.zero_grad()                                          # main.py:440
            gs_plugin.exp_count += 1                                       # main.py:442
            losses.append(o["loss"].clone())
        if pending:                                                        # deferred encoder updates (main.py:439)
            for m, stream in pending:
                torch.cuda.current_stream().wait_stream(stream)
                if world > 1:
                    mdist.allreduce_sum_(st.flat[m].flat)
                st.flat[m].attach()
            optimizer.step()
            optimizer.zero_grad()
        # main.py:472 (fp32, like the reference; with three modalities the mixed loss still ignores the third)
        mix = losses[0] * av_alpha + losses[1] * (1 - av_alpha)
        step_vec = torch.cat([mix] + losses).double()                      # (_loss, _loss_a, _loss_v[, _loss_t]) of this step
        acc += step_vec
        if step_log is not None:                                           # per-step device->host read, asynchronous
            step_log[batch_step].copy_(step_vec, non_blocking=True)
    scheduler.step()                                                       # main.py:481
    if world > 1:
        mdist.allreduce_sum_(acc)
        acc /= world
    vals = (acc / len_dataloader).tolist()                                 # the only device->host sync of the epoch
    if step_log is not None:
        train_epoch.last_step_losses = step_log                            # complete: the .tolist() above synchronised
    return tuple(vals)


@torch.no_grad()
def valid(args, model, device, dataloader, gs_flag=False, av_alpha=0.5,
          a_alpha=0.35, v_alpha=0.25, t_alpha=0.4):
    """main.py:486-679 (gs branch). Returns (acc, acc_a, acc_v[, acc_t]) micro-accuracies."""
    if args.dataset not in N_CLASSES:
        raise NotImplementedError("Incorrect dataset name {}".format(args.dataset))
    if not gs_flag:
        raise NotImplementedError("mla_b200 implements the --gs_flag evaluation branch only (main.py:622-679)")
    n_classes = N_CLASSES[args.dataset]
    net = _unwrap(model)
    model.eval()
    mdist.broadcast_buffers_(net)                                          # DataParallel keeps replica 0's BN statistics
    fc = net.fusion_module.fc_out
    n_mod = 3 if getattr(args, "modal3", False) else 2
    num = torch.zeros(n_classes, dtype=torch.int64, device=device)
    hits = torch.zeros(n_mod + 1, n_classes, dtype=torch.int64, device=device)
    if args.dynamic:
        fixed = None
    elif n_mod == 3:
        fixed = (a_alpha, v_alpha, t_alpha)                                # main.py:649
    else:
        fixed = (av_alpha, 1 - av_alpha)                                   # main.py:651
    for inputs, label in _batches_on_device(args, dataloader, device):
        feats = model(*inputs)                                             # main.py:624-634
        logits = [fc(f.contiguous()) for f in feats]                       # main.py:636-639
        if mdist.is_dist():
            # the entropy weights are a GLOBAL-batch quantity: gather the tiny logits
            logits = [mdist.all_gather_rows(x) for x in logits]
            label = mdist.all_gather_rows(label)
        ops.fuse_eval(logits, label, dynamic=bool(args.dynamic), fixed_w=fixed, hits=hits, num=num,
                      want_fused=False, want_argmax=False)                 # main.py:640-676
    h = hits.sum(dim=1).tolist()
    n = float(num.sum().item())
    return tuple(x / n for x in h)
