"""Command-line driver with the reference's flag surface (main.py:18-63, 697-964) for the
MLA --gs_flag path, on synthetic data of the BASELINE.json shapes.

    python -m mla_b200.main --train --lorb base --gs_flag --dynamic --dataset CREMAD --ckpt_path ckpt \
        --synthetic --steps 8 --epochs 1
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m mla_b200.main ... (one process per GPU)

Every reference flag is accepted with the reference's default. Real datasets are out of scope
(they need private files under /data1/...), so data is always synthetic here.
"""
import argparse
import os

import torch
import torch.optim as optim

from . import dist as mdist
from .basic_model import AVClassifier
from .engine import ModuleHolder, train_epoch, valid
from .gs_plugin import GSPlugin
from .m3ae import M3AEClassifier
from .cav_mae import Modal3Classifier
from .utils import setup_seed, weight_init


def get_arguments(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", default="CREMA-D", type=str)
    p.add_argument("--modulation", default="Normal", type=str, choices=["Normal", "OGM", "OGM_GE", "QMF"])
    p.add_argument("--fusion_method", default="concat", type=str, choices=["sum", "concat", "gated", "film"])
    p.add_argument("--fps", default=1, type=int)
    p.add_argument("--use_video_frames", default=3, type=int)
    p.add_argument("--batch_size", default=64, type=int)
    p.add_argument("--epochs", default=100, type=int)
    p.add_argument("--optimizer", default="sgd", type=str, choices=["sgd", "adam"])
    p.add_argument("--learning_rate", default=0.001, type=float)
    p.add_argument("--lr_decay_step", default=70, type=int)
    p.add_argument("--lr_decay_ratio", default=0.1, type=float)
    p.add_argument("--modulation_starts", default=0, type=int)
    p.add_argument("--modulation_ends", default=50, type=int)
    p.add_argument("--alpha", default=0.3, type=float)
    p.add_argument("--ckpt_path", required=True, type=str)
    p.add_argument("--train", action="store_true")
    p.add_argument("--use_tensorboard", default=True, type=bool)
    p.add_argument("--tensorboard_path", default="ckpt/", type=str)
    p.add_argument("--random_seed", default=0, type=int)
    p.add_argument("--gpu_ids", default="0, 1, 2", type=str)
    p.add_argument("--lorb", default="m3ae", type=str)
    p.add_argument("--gs_flag", action="store_true")
    p.add_argument("--av_alpha", default=0.5, type=float)
    p.add_argument("--cav_opti", action="store_true")
    p.add_argument("--cav_lrs", action="store_true")
    p.add_argument("--cav_augnois", action="store_true")
    p.add_argument("--modal3", action="store_true")
    p.add_argument("--dynamic", action="store_true")
    p.add_argument("--a_alpha", default=0.35, type=float)
    p.add_argument("--v_alpha", default=0.25, type=float)
    p.add_argument("--t_alpha", default=0.4, type=float)
    p.add_argument("--clip", action="store_true")
    p.add_argument("--ckpt_load_path_train", default=None, type=str)
    # additions (not in the reference)
    p.add_argument("--synthetic", action="store_true", help="synthetic batches of the BASELINE.json shapes")
    p.add_argument("--steps", default=8, type=int, help="synthetic batches per epoch")
    p.add_argument("--force_projection", action="store_true",
                   help="fire the GS projection on the bare Linear (the published hook is a no-op, SURVEY F1)")
    return p.parse_args(argv)


class SyntheticAVLoader:
    """len()-able iterable of CREMA-D-shaped batches (spec [B,257,188], image [B,3,T,224,224],
    label, idx), pinned host memory, seeded per rank."""

    def __init__(self, batch_size, steps, seed, frames=2, spec_hw=(257, 188), image_hw=(224, 224), n_classes=6,
                 distinct=2):
        g = torch.Generator().manual_seed(seed)
        self.batches = []
        for _ in range(min(distinct, steps)):
            spec = torch.randn(batch_size, *spec_hw, generator=g)
            image = torch.randn(batch_size, 3, frames, *image_hw, generator=g)
            label = torch.randint(0, n_classes, (batch_size,), generator=g)
            idx = torch.zeros(batch_size, 1, dtype=torch.long)
            if torch.cuda.is_available():
                spec, image, label = spec.pin_memory(), image.pin_memory(), label.pin_memory()
            self.batches.append((spec, image, label, idx))
        self.steps = steps

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            yield self.batches[i % len(self.batches)]


class SyntheticTextImageLoader:
    """len()-able iterable of Food-101-shaped m3ae batches (token [B,1,256] int64, padding_mask [B,1,L], image
    [B,3,256,256], label, idx), pinned host memory, seeded per rank (dataset layout: dataset/*.py __getitem__).
    audio_len > 0: IEMOCAP-shaped three-modality batches (token, padding_mask, image, spec [B,audio_len,128], label, idx)."""

    def __init__(self, batch_size, steps, seed, text_len=256, image_hw=(256, 256), n_classes=101, vocab=30522, distinct=2,
                 audio_len=0):
        g = torch.Generator().manual_seed(seed)
        self.batches = []
        for _ in range(min(distinct, steps)):
            token = torch.randint(0, vocab, (batch_size, 1, text_len), generator=g)
            n_valid = torch.randint(max(1, text_len // 4), text_len + 1, (batch_size,), generator=g)
            pm = (torch.arange(text_len)[None, :] >= n_valid[:, None]).long()[:, None, :]
            image = torch.randn(batch_size, 3, *image_hw, generator=g)
            label = torch.randint(0, n_classes, (batch_size,), generator=g)
            idx = torch.zeros(batch_size, 1, dtype=torch.long)
            spec = torch.randn(batch_size, audio_len, 128, generator=g) if audio_len else None
            if torch.cuda.is_available():
                token, pm, image, label = token.pin_memory(), pm.pin_memory(), image.pin_memory(), label.pin_memory()
                spec = spec.pin_memory() if audio_len else None
            self.batches.append((token, pm, image, spec, label, idx) if audio_len else (token, pm, image, label, idx))
        self.steps = steps

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            yield self.batches[i % len(self.batches)]


def build_model(args, device):
    """main.py:707-734 for the paths in scope."""
    if args.lorb == "large" or args.clip:
        raise NotImplementedError("implemented: --lorb base (AVClassifier), --lorb m3ae (M3AEClassifier) and --lorb m3ae "
                                  "--modal3 (Modal3Classifier); CAVClassifier / CLIPClassifier are out of scope "
                                  "(SURVEY.md section 2)")
    if args.lorb == "m3ae":                                      # main.py:708-712 (no weight_init, like the reference)
        model = Modal3Classifier(args) if args.modal3 else M3AEClassifier(args)
    else:
        model = AVClassifier(args)
        model.apply(weight_init)
    if args.ckpt_load_path_train:
        loaded = torch.load(args.ckpt_load_path_train, map_location="cpu")["model"]
        state = {k[7:]: v for k, v in loaded.items()}            # strip 'module.' (main.py:723)
        state.pop("fusion_module.fc_out.weight", None)
        state.pop("fusion_module.fc_out.bias", None)
        model.load_state_dict(state, strict=False)
        print("Trained model loaded!")
    return ModuleHolder(model.to(device))


def checkpoint_name(args, epoch, acc):
    """main.py:900-909."""
    return ("best_model_of_dataset_{}_{}_alpha_{}_optimizer_{}_modulate_starts_{}_ends_{}_epoch_{}_acc_{}.pth".format(
        args.dataset, args.modulation, args.alpha, args.optimizer, args.modulation_starts, args.modulation_ends, epoch, acc))


def save_checkpoint(path, args, epoch, acc, model, optimizer, scheduler, gs_plugin=None):
    """The reference's checkpoint dictionary (main.py:911-921: saved_epoch, modulation, alpha, fusion, acc, model with the
    DataParallel 'module.' prefix, optimizer, scheduler) plus one extension the reference lacks: the GSPlugin state
    (`Pl`, `exp_count`), without which a resumed --gs_flag run restarts the projection from the identity."""
    torch.save({"saved_epoch": epoch, "modulation": args.modulation, "alpha": args.alpha, "fusion": args.fusion_method,
                "acc": acc, "model": model.state_dict(), "optimizer": optimizer.state_dict(),
                "scheduler": scheduler.state_dict(),
                "gs_plugin": gs_plugin.state_dict() if gs_plugin is not None else None}, path)


def load_checkpoint(path, model, optimizer=None, scheduler=None, gs_plugin=None, map_location="cpu"):
    """main.py:946-953 (evaluation: strict load of loaded['model'] into the wrapped model) and, when optimizer / scheduler /
    gs_plugin are given, a full training resume. Returns the loaded dictionary."""
    loaded = torch.load(path, map_location=map_location)
    model.load_state_dict(loaded["model"])                           # strict, 'module.'-prefixed keys (main.py:952)
    if optimizer is not None and loaded.get("optimizer") is not None:
        optimizer.load_state_dict(loaded["optimizer"])
    if scheduler is not None and loaded.get("scheduler") is not None:
        scheduler.load_state_dict(loaded["scheduler"])
    if gs_plugin is not None and loaded.get("gs_plugin") is not None:
        gs_plugin.load_state_dict(loaded["gs_plugin"])
    return loaded


def main(av_alpha=0.5):
    args = get_arguments()
    if args.dataset == "CREMA-D":
        args.dataset = "CREMAD"
    rank, world = mdist.init_from_env()
    setup_seed(args.random_seed)
    device = torch.device("cuda", torch.cuda.current_device())
    model = build_model(args, device)
    optimizer = optim.SGD(model.parameters(), lr=args.learning_rate, momentum=0.9, weight_decay=1e-4)   # main.py:749
    scheduler = optim.lr_scheduler.StepLR(optimizer, args.lr_decay_step, args.lr_decay_ratio)           # main.py:760
    if args.lorb == "m3ae" and args.modal3:
        def Loader(b, n, seed):
            return SyntheticTextImageLoader(b, n, seed, n_classes=4, audio_len=1024)
    else:
        Loader = SyntheticTextImageLoader if args.lorb == "m3ae" else SyntheticAVLoader
    train_loader = Loader(args.batch_size, args.steps, seed=1 + rank)
    test_loader = Loader(args.batch_size, max(1, args.steps // 2), seed=1001 + rank)
    gs = GSPlugin(force_projection=args.force_projection) if args.gs_flag else None                     # main.py:819
    if args.train:
        best_acc = 0.0
        for epoch in range(args.epochs):
            if rank == 0:
                print("Epoch: {}: ".format(epoch))
            losses = train_epoch(args, epoch, model, device, train_loader, optimizer, scheduler, gs_plugin=gs,
                                 gs_flag=args.gs_flag, av_alpha=av_alpha)
            accs = valid(args, model, device, test_loader, gs_flag=args.gs_flag, av_alpha=av_alpha,
                         a_alpha=args.a_alpha, v_alpha=args.v_alpha, t_alpha=args.t_alpha)
            if rank == 0:
                print("Loss: {:.4f}, Acc: {:.4f}, Acc_a: {:.4f}, Acc_v: {:.4f}".format(losses[0], accs[0], accs[1],
                                                                                       accs[2])
                      + (", Acc_t: {:.4f}".format(accs[3]) if len(accs) > 3 else ""))
                if accs[0] > best_acc:
                    best_acc = float(accs[0])
                    os.makedirs(args.ckpt_path, exist_ok=True)
                    save_checkpoint(os.path.join(args.ckpt_path, checkpoint_name(args, epoch, accs[0])), args, epoch,
                                    accs[0], model, optimizer, scheduler, gs)
    else:
        load_checkpoint(args.ckpt_path, model)                                                         # main.py:946-953
        print("Trained model loaded!")
        accs = valid(args, model, device, test_loader, gs_flag=args.gs_flag, av_alpha=args.av_alpha,
                     a_alpha=args.a_alpha, v_alpha=args.v_alpha, t_alpha=args.t_alpha)
        if rank == 0:
            print("Acc: {:.4f}, Acc_a: {:.4f}, Acc_v: {:.4f}".format(*accs[:3]))


if __name__ == "__main__":
    main(av_alpha=0.55)        # main.py:968
