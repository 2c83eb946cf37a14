"""train_epoch / valid — the reference's L4 drivers for the --gs_flag path (main.py:127-484,
486-679), same signatures and return values, B200-native underneath:

  * per modality turn the head forward+loss+backward is ONE fused call (head kernel), the GS
    hook is one fused kernel, the encoder backward gets dfeat directly;
  * no per-step `.item()` syncs: epoch losses accumulate on the device and are read once;
  * evaluation does fusion, argmax and per-class counters in one kernel per batch — the
    reference's per-sample numpy loop (~8 D2H syncs per sample, main.py:659-676) is gone;
  * multi-GPU is one process per GPU with two NCCL all-reduces per turn (dist.py).
Without --gs_flag the same kernels run the reference's JOINT-training step (main.py:165-168, 269-310, 412-418: one
concatenated head, one backward through every encoder, one optimiser step) with its OGM / OGM-GE gradient modulation
(main.py:312-410) — the "MLA vs joint" comparison line. Out of scope (raises): QMF, --lorb large, --clip, sum/film/gated.
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
        # data parallel: every encoder's gradients travel as two asynchronous buckets (layer4 first, while the rest of the
        # backward runs); encoders whose backward is deferred to a side stream use their own communicator so that their
        # buckets never queue in front of the next turn's small head all-reduce
        self.buckets = [mdist.BucketedAllReduce(fg.flat, mdist.aux_group("enc%d" % m) if m < len(self.flat) - 1 else None)
                        for m, fg in enumerate(self.flat)]

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
        return _train_epoch_joint(args, epoch, model, device, dataloader, optimizer, scheduler)
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
            deferred_m = None
            if feat_streams is not None:           # this turn starts as soon as ITS encoder's forward has finished
                torch.cuda.current_stream().wait_stream(feat_streams[m])
                feat.record_stream(torch.cuda.current_stream())
            fdet = feat.detach()
            st.head_out["dfeat"] = st.dfeat_buffer(m, fdet)                # one dfeat buffer per modality (see below)
            o = head_turn(fc, fdet, label, grad_scale=inv_global, out=st.head_out)   # main.py:432-435 (head part)
            # the small head all-reduce ([dW | db | sum_b feat], 14 KB). Only the GS projection and the head's SGD
            # read its result — the encoder backward differentiates the LOCAL loss through dfeat — so it is issued here and
            # awaited after the backward has been queued: its latency hides behind the first backward kernels
            head_ar = mdist.allreduce_sum_async(st.packed) if world > 1 else None
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
                    if world > 1:
                        # encoder-gradient all-reduce, bucketed and overlapped — the layer4 bucket (75 % of the
                        # bytes) leaves while layer3 .. stem are still being differentiated, layer3 + layer2 (24 %) during
                        # layer1 + stem; only the 0.6 MB layer1 + stem bucket of the last encoder is exposed
                        bk, (off2, off4) = st.buckets[m], plan.bucket_offsets()
                        spans = ((off4, st.flat[m].flat.numel()), (off2, off4), (0, off2))
                        plan.backward(o["dfeat"], getattr(feat, "_mla_serial", None),
                                      on_segment=lambda k, bk=bk, spans=spans: bk.span(*spans[k]))
                        if not deferred:
                            bk.wait()
                    else:
                        plan.backward(o["dfeat"], getattr(feat, "_mla_serial", None))
                if deferred:
                    # (a deferred encoder's buckets use their own communicator: on the main one they would park the NEXT
                    # turn's small head all-reduce behind this whole backward pass and serialise the two encoders again)
                    st.flat[m].detach()
                    pending.append((m, stream))
                    deferred_m = m
            else:                                  # autograd encoders (m3ae): gradients ACCUMULATE into the views
                st.flat[m].attach(zero=True)
                feat.backward(o["dfeat"])
                if world > 1:
                    mdist.allreduce_sum_(st.flat[m].flat)
            if head_ar is not None:
                head_ar.wait()
            gs_plugin.before_update(fc, fdet, batch_step, len_dataloader, gs_plugin.exp_count,
                                    feat_sum=o["feat_sum"], inv_batch=inv_global)     # main.py:437-438
            optimizer.step()                                               # main.py:439
            optimizer.zero_grad()                                          # main.py:440
            if deferred_m is not None:
                # the deferred encoder's own SGD update (its share of main.py:439), queued on ITS side stream behind its
                # backward pass and gradient buckets: it runs while the next turn is busy instead of at the end of the step.
                # Only this encoder's gradients are attached at this point (the head's were just cleared), so the caller's
                # optimizer touches nothing else; parameters and momentum buffers of different encoders are disjoint.
                with torch.cuda.stream(pending[-1][1]):
                    if world > 1:
                        st.buckets[deferred_m].wait()                      # its buckets were issued during the backward
                    st.flat[deferred_m].attach()
                    optimizer.step()
                    optimizer.zero_grad()
                deferred_m = None
            gs_plugin.exp_count += 1                                       # main.py:442
            losses.append(o["loss"].clone())
        for m, stream in pending:                                          # the next forward reads the updated encoders
            torch.cuda.current_stream().wait_stream(stream)
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


# ------------------------------------------------------------------------------------------------------------------
# Joint training (no --gs_flag): main.py:165-168, 269-310 (forward + per-modality logit slices), 312-410 (OGM / OGM-GE),
# 412-418 (step, loss accumulation). SURVEY section 8 f1 / f2.
# ------------------------------------------------------------------------------------------------------------------
def _check_joint_scope(args):
    if getattr(args, "modulation", "Normal") == "QMF":
        raise NotImplementedError("QMF (main.py:171-225) is outside the hot path (SURVEY.md section 2)")
    if getattr(args, "fusion_method", "concat") != "concat":
        raise NotImplementedError("only the concat head is implemented (got fusion_method={})".format(args.fusion_method))
    if getattr(args, "lorb", "base") == "large" or getattr(args, "clip", False):
        raise NotImplementedError("--lorb large / --clip are out of scope (SURVEY.md section 2)")
    if getattr(args, "modulation", "Normal") not in ("Normal", "OGM", "OGM_GE"):
        raise NotImplementedError("unknown modulation {}".format(args.modulation))


def _encoder_names(net):
    for cand in (("audio_net", "visual_net"), ("mae_a", "mae_v", "mae_t")):
        if all(hasattr(net, n) for n in cand[:2]):
            return [n for n in cand if hasattr(net, n)]
    raise RuntimeError("model has no known encoders (audio_net/visual_net or mae_a/mae_v[/mae_t])")


def ogm_coeff_index(name, modal3):
    """Which coefficient the reference's name test applies to the parameters of encoder `name`: main.py:350,357,362 for
    three modalities ('mae_a' / 'mae_v' / 'mae_t' in layer), main.py:396,403 otherwise ('audio' / 'visual' in layer — which
    the two-modality m3ae encoders `mae_a` / `mae_v` never match: as published, OGM is a no-op for them). None = untouched."""
    if modal3:
        return {"mae_a": 0, "mae_v": 1, "mae_t": 2}.get(name)
    if "audio" in name:
        return 0
    if "visual" in name:
        return 1
    return None


class _JointState:
    """Per-model buffers of the joint step: packed [dW | db] of the concatenated head, flat per-encoder gradient buckets,
    and for OGM the segment tables of every encoder's 4-D parameters inside its bucket."""

    def __init__(self, net, modal3):
        fc = net.fusion_module.fc_out
        C, Dcat = fc.weight.shape
        dev = fc.weight.device
        o_db = (C * Dcat + 3) // 4 * 4
        self.packed = torch.zeros(o_db + (C + 3) // 4 * 4, dtype=torch.float32, device=dev)
        self.head_out = {"dW": self.packed[:C * Dcat].view(C, Dcat), "db": self.packed[o_db:o_db + C]}
        self.names = _encoder_names(net)
        self.encoders = encoder_param_groups(net)
        self.flat = [mdist.FlatGrads(g) for g in self.encoders]
        self.score = torch.zeros(len(self.names), dtype=torch.float32, device=dev)
        self.coeff = torch.ones(len(self.names), dtype=torch.float32, device=dev)
        self.slice_out = [dict() for _ in self.names]
        self.seg = []
        for name, fg in zip(self.names, self.flat):
            ci = ogm_coeff_index(name, modal3)
            offs, lens, views, off = [], [], [], 0
            for p, v in zip(fg.params, fg.views):
                if p.dim() == 4 and ci is not None:                       # len(parms.grad.size()) == 4: conv weights only
                    offs.append(off); lens.append(p.numel()); views.append(v)
                off += p.numel()
            if offs:
                self.seg.append(dict(ci=ci, off=torch.tensor(offs, dtype=torch.int64, device=dev),
                                     len=torch.tensor(lens, dtype=torch.int64, device=dev), max_len=max(lens), views=views,
                                     params=[p for p in fg.params if p.dim() == 4], noise=None))
            else:
                self.seg.append(None)
        self._side = {}

    def side_stream(self, m):
        s = self._side.get(m)
        if s is None:
            s = torch.cuda.Stream()
            self._side[m] = s
        return s


def _modality_logits(fc, feats, label, outs):
    """out_m = feat_m W[:, m-th slice]^T + b / M (main.py:276-308) with its mean CE, through the head kernel (forward only)."""
    M = len(feats)
    D = fc.weight.shape[1] // M
    b = fc.bias.detach() / M
    res = []
    for m, f in enumerate(feats):
        w = fc.weight.detach()[:, m * D:(m + 1) * D].contiguous()
        res.append(ops.head_ce(f.detach().contiguous(), w, b, label, need_grad=False, out=outs[m]))
    return res


def _train_epoch_joint(args, epoch, model, device, dataloader, optimizer, scheduler):
    """main.py:127-168, 269-418 for --modulation Normal / OGM / OGM_GE. Returns (loss, loss_a, loss_v[, loss_t])."""
    _check_joint_scope(args)
    net = _unwrap(model)
    model.train()
    print("Start training ... ")
    fc = net.fusion_module.fc_out
    world = mdist.world_size()
    modal3 = bool(getattr(args, "modal3", False))
    names = _encoder_names(net)
    n_mod = len(names)
    if fc.weight.shape[1] % n_mod:
        raise RuntimeError("joint training needs the concatenated head (width = %d x feature width); build the model "
                           "without --gs_flag" % n_mod)
    modulate = args.modulation in ("OGM", "OGM_GE")
    acc = torch.zeros(1 + n_mod, dtype=torch.float64, device=device)
    len_dataloader = len(dataloader)
    for batch_step, (inputs, label) in enumerate(_batches_on_device(args, dataloader, device)):
        optimizer.zero_grad()                                              # main.py:164
        if basic_model.OVERLAP_ENCODERS and hasattr(net, "forward_streams") and inputs[0].is_cuda:
            pairs = net.forward_streams(*inputs)                           # main.py:166 / 227-232 / 265-268
            feats = []
            for f, s in pairs:
                torch.cuda.current_stream().wait_stream(s)
                f.record_stream(torch.cuda.current_stream())
                feats.append(f)
        else:
            feats = list(net.features(*inputs)) if hasattr(net, "features") else list(model(*inputs))[:n_mod]
        st = getattr(net, "_mla_joint_state", None)
        if st is None:
            st = _JointState(net, modal3)
            net._mla_joint_state = st
        B = feats[0].shape[0]
        D = feats[0].shape[1]
        inv_global = 1.0 / (B * world)
        cat = torch.cat([f.detach() for f in feats], dim=1)                # fusion_modules.py:22 / 32
        o = head_turn(fc, cat, label, grad_scale=inv_global, out=st.head_out)     # main.py:309 + head part of :313
        if world > 1:
            mdist.allreduce_sum_(st.packed)
        per_mod = _modality_logits(fc, feats, label, st.slice_out)         # main.py:276-308, 310-312 (logging, OGM scores)
        # main.py:313 (encoder part): every encoder's backward from its slice of d(loss)/d(cat)
        native = [getattr(f, "_mla_plan", None) for f in feats]
        pending = []
        for m, (f, plan) in enumerate(zip(feats, native)):
            dslice = o["dfeat"][:, m * D:(m + 1) * D]
            if plan is not None:
                side = basic_model.OVERLAP_ENCODERS and m < n_mod - 1
                stream = st.side_stream(m) if side else torch.cuda.current_stream()
                if side:
                    stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(stream):
                    st.flat[m].attach()
                    plan.backward(dslice, getattr(f, "_mla_serial", None))
                if side:
                    pending.append(stream)
            else:
                st.flat[m].attach(zero=True)
                f.backward(dslice.contiguous())
        for stream in pending:
            torch.cuda.current_stream().wait_stream(stream)
        if world > 1:
            for fg in st.flat:
                mdist.allreduce_sum_(fg.flat)
        if modulate:                                                       # main.py:315-410
            ops.ogm_scores([r["logits"] for r in per_mod], label, out=st.score)
            if world > 1:
                mdist.allreduce_sum_(st.score)                             # the scores are sums over the GLOBAL batch
            ops.ogm_coeff(st.score, args.alpha, out=st.coeff)
            if args.modulation_starts <= epoch <= args.modulation_ends:    # main.py:343 / 393
                for m, seg in enumerate(st.seg):
                    if seg is not None:
                        _ogm_apply(st, m, seg, args.modulation == "OGM_GE")
            train_epoch.last_ogm = (st.score, st.coeff)
        optimizer.step()                                                   # main.py:412
        step_vec = torch.cat([o["loss"]] + [r["loss"] for r in per_mod]).double()     # main.py:414-418
        acc += step_vec
    scheduler.step()                                                       # main.py:481
    if world > 1:
        mdist.allreduce_sum_(acc)
        acc /= world
    return tuple((acc / len_dataloader).tolist())


def _ogm_apply(st, m, seg, ge):
    """main.py:393-408 for one encoder: grad = grad * coeff [+ N(0, std(grad) + 1e-8)] over its 4-D parameters."""
    flat = st.flat[m].flat
    coeff = st.coeff[seg["ci"]:seg["ci"] + 1]
    if not ge:
        ops.ogm_modulate(flat, seg["off"], seg["len"], seg["max_len"], coeff)
        return
    if seg["noise"] is None:
        seg["noise"] = torch.empty_like(flat)
        seg["nviews"] = [torch.as_strided(seg["noise"], v.shape, v.stride(), storage_offset=int(o))
                         for v, o in zip(seg["views"], seg["off"].tolist())]
    stds = []
    for v, nv in zip(seg["views"], seg["nviews"]):
        # the reference draws zeros_like(grad).normal_(0, std) parameter by parameter in named_parameters() order from the
        # default CUDA generator: the same draws, in the same logical element order, scaled inside the kernel
        stds.append(v.std().double() + 1e-8)
        nv.copy_(torch.empty(v.shape, dtype=torch.float32, device=v.device).normal_())
    seg_std = torch.stack(stds).float()
    ops.ogm_modulate(flat, seg["off"], seg["len"], seg["max_len"], coeff, noise=seg["noise"], seg_std=seg_std)


@torch.no_grad()
def _valid_joint(args, model, net, fc, n_mod, n_classes, device, dataloader):
    """main.py:538-620, 653-679: the joint head's prediction plus the per-modality slices (bias / M)."""
    num = torch.zeros(n_classes, dtype=torch.int64, device=device)
    hits = torch.zeros(n_mod + 2, n_classes, dtype=torch.int64, device=device)
    outs = [dict() for _ in range(n_mod)]
    for inputs, label in _batches_on_device(args, dataloader, device):
        feats = list(net.features(*inputs)) if hasattr(net, "features") else list(model(*inputs))[:n_mod]
        feats = [f.contiguous() for f in feats]
        out = fc(torch.cat(feats, dim=1))                                  # main.py:543 / 575-579 / 598
        logits = [out] + [r["logits"] for r in _modality_logits(fc, feats, label, outs)]
        if mdist.is_dist():
            sizes = mdist.gather_sizes(label.shape[0], label.device)
            logits = [mdist.all_gather_rows(x, sizes) for x in logits]
            label = mdist.all_gather_rows(label, sizes)
        # fused = 1 * out + 0 * out_a + ... = out exactly: rows of `hits` are [out, out, out_a, out_v(, out_t)]
        ops.fuse_eval(logits, label, dynamic=False, fixed_w=(1.0,) + (0.0,) * n_mod, hits=hits, num=num,
                      want_fused=False, want_argmax=False)
    h = hits.sum(dim=1).tolist()
    n = float(num.sum().item())
    return tuple(h[i] / n for i in [0] + list(range(2, n_mod + 2)))


@torch.no_grad()
def valid(args, model, device, dataloader, gs_flag=False, av_alpha=0.5,
          a_alpha=0.35, v_alpha=0.25, t_alpha=0.4):
    """main.py:486-679 (gs branch). Returns (acc, acc_a, acc_v[, acc_t]) micro-accuracies."""
    if args.dataset not in N_CLASSES:
        raise NotImplementedError("Incorrect dataset name {}".format(args.dataset))
    if not gs_flag:
        _check_joint_scope(args)
    n_classes = N_CLASSES[args.dataset]
    net = _unwrap(model)
    model.eval()
    mdist.broadcast_buffers_(net)                                          # DataParallel keeps replica 0's BN statistics
    fc = net.fusion_module.fc_out
    n_mod = 3 if getattr(args, "modal3", False) else 2
    if not gs_flag:
        return _valid_joint(args, model, net, fc, n_mod, n_classes, device, dataloader)
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
            # (every rank then counts the same gathered batch: the counters are identical on all ranks by construction).
            # Shards may differ in size on the last batch: sizes are exchanged once per batch
            sizes = mdist.gather_sizes(label.shape[0], label.device)
            logits = [mdist.all_gather_rows(x, sizes) for x in logits]
            label = mdist.all_gather_rows(label, sizes)
        ops.fuse_eval(logits, label, dynamic=bool(args.dynamic), fixed_w=fixed, hits=hits, num=num,
                      want_fused=False, want_argmax=False)                 # main.py:640-676
    h = hits.sum(dim=1).tolist()
    n = float(num.sum().item())
    return tuple(x / n for x in h)
