"""Dataset tuple producer for the visual modality — the batch-level, on-GPU form of what the reference's datasets do per
sample on CPU workers (dataset/dataset.py:123-161 `AVDataset.__getitem__`; the same Compose at :448-480 and :753-803):

    train:  RandomResizedCrop(224) -> RandomHorizontalFlip() -> ToTensor() -> Normalize(IMAGENET_MEAN, IMAGENET_STD)
    test:   Resize(size=(224, 224))                          -> ToTensor() -> Normalize(...)
    frames of a sample stacked on a new dim 1 -> [3, T, 224, 224]
and the deterministic transform of the other datasets (CAVDataset dataset.py:251-256, M3AEDataset test mode :414-421):
    center: Resize(s, BICUBIC) [shorter side -> s] -> CenterCrop(s) -> ToTensor() -> Normalize(...)
(M3AEDataset's TRAINING transform is timm's create_transform — RandomResizedCropAndInterpolation + colour jitter from the
un-vendored timm==0.4.5 — and is not reproduced.)

`FrameBatchProducer` takes the DECODED frames (uint8 HWC RGB arrays: decoding stays on the host, PIL / the loader's
workers), draws the crop / flip parameters on the host exactly as torchvision does — same functions, same order of torch
RNG draws per frame, so a seeded run picks the same crops as the reference's Compose — and hands geometry + pixels to
`mla_frames_to_batch`, whose output is bit-identical to the torchvision-on-PIL path (csrc/frame_producer.cu). The audio
member of the tuple needs no producer: the reference loads pre-computed spectrogram arrays (`np.load`, dataset.py:116).
"""
import math

import numpy as np
import torch

from . import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)       # dataset/dataset.py:131, 137
IMAGENET_STD = (0.229, 0.224, 0.225)


def random_resized_crop_params(height, width, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """torchvision.transforms.RandomResizedCrop.get_params restated draw for draw (torch's global CPU generator): up to ten
    attempts of (area fraction ~ U(scale), log aspect ratio ~ U(log ratio), then i, j ~ randint), else the central fallback
    crop. Returns (top, left, h, w)."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w = width
        h = height
    return (height - h) // 2, (width - w) // 2, h, w


def resize_center_crop_params(height, width, size):
    """Resize(size) + CenterCrop(size) on a PIL image (torchvision.transforms.functional): the shorter side becomes `size`,
    the longer int(size * long / short); the crop origin is int(round((extent - size) / 2.0)) (Python's round).
    Returns (RH, RW, oy, ox)."""
    if width <= height:
        rw, rh = size, int(size * height / width)
    else:
        rh, rw = size, int(size * width / height)
    return rh, rw, int(round((rh - size) / 2.0)), int(round((rw - size) / 2.0))


class FrameBatchProducer:
    """producer(batch_of_frames) -> CUDA tensor [B, 3, T, size, size], the `image` member of the reference's batch tuple
    (main.py:159). `batch_of_frames`: B samples x T frames, each a uint8 numpy array [H, W, 3] (any sizes).

    mode 'train' draws RandomResizedCrop + RandomHorizontalFlip parameters per frame in the reference's order (crop, then
    flip: one Compose call per frame, dataset.py:147-150); mode 'test' resizes whole frames; mode 'center' = Resize(size) +
    CenterCrop(size). `interpolation` 'bilinear' (AVDataset) or 'bicubic' (CAVDataset / M3AEDataset). Explicit `params` (a
    list of (top, left, h, w, flip) per frame, sample-major) override the random draws."""

    def __init__(self, size=224, mode="train", mean=IMAGENET_MEAN, std=IMAGENET_STD, device="cuda", interpolation="bilinear"):
        if mode not in ("train", "test", "center"):
            raise ValueError("mode must be 'train', 'test' or 'center'")
        if interpolation not in ("bilinear", "bicubic"):
            raise ValueError("interpolation must be 'bilinear' or 'bicubic'")
        self.bicubic = interpolation == "bicubic"
        self.size, self.mode, self.mean, self.std = int(size), mode, tuple(mean), tuple(std)
        self.device = torch.device(device)
        self._pinned = None
        self._status = None

    def _params(self, frames):
        out = []
        for f in frames:
            H, W = f.shape[:2]
            if self.mode == "train":
                i, j, h, w = random_resized_crop_params(H, W)
                flip = bool(torch.rand(1) < 0.5)                      # RandomHorizontalFlip.forward
                out.append((i, j, h, w, flip))
            else:
                out.append((0, 0, H, W, False))
        return out

    def __call__(self, batch_of_frames, params=None, out=None, check=True):
        B = len(batch_of_frames)
        T = len(batch_of_frames[0])
        frames = [np.ascontiguousarray(f) for sample in batch_of_frames for f in sample]
        if any(len(s) != T for s in batch_of_frames):
            raise ValueError("every sample needs the same number of frames")
        for f in frames:
            if f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != 3:
                raise ValueError("frames must be uint8 [H, W, 3] arrays")
        if params is None:
            params = self._params(frames)
        if len(params) != len(frames):
            raise ValueError("one (top, left, h, w, flip) tuple per frame")
        total = sum(f.size for f in frames)
        if self._pinned is None or self._pinned.numel() < total:
            self._pinned = torch.empty(total, dtype=torch.uint8).pin_memory() if self.device.type == "cuda" else torch.empty(total, dtype=torch.uint8)
        host = self._pinned[:total].numpy()
        desc = np.zeros((len(frames), 14), np.int32)
        off = 0
        for n, (f, (i, j, h, w, flip)) in enumerate(zip(frames, params)):
            host[off:off + f.size] = f.reshape(-1)
            lo = off & 0xFFFFFFFF
            if self.mode == "center":                              # the (whole-frame) box is resized, then centre-cropped
                rh, rw, oy, ox = resize_center_crop_params(h, w, self.size)
            else:
                rh, rw, oy, ox = self.size, self.size, 0, 0
            desc[n] = (lo - (1 << 32) if lo >= (1 << 31) else lo, off >> 32, f.shape[0], f.shape[1], i, j, h, w,
                       1 if flip else 0, n, rh, rw, oy, ox)
            off += f.size
        src = self._pinned[:total].to(self.device, non_blocking=True)
        ddesc = torch.from_numpy(desc).to(self.device, non_blocking=True)
        if self._status is None:
            self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        res = ops.frames_to_batch(src, ddesc, B, T, self.size, max(p[2] for p in params), self.mean, self.std, out=out,
                                  status=self._status if check else None, bicubic=self.bicubic)
        if check:
            bad = int(self._status.item())
            if bad:
                raise RuntimeError("frame %d has an invalid descriptor (crop outside the frame, or a crop / output ratio above "
                                   "15.5 bilinear / 7.5 bicubic)" % (bad - 1))
        return res


def mask_interval(mask_param, size):
    """torchaudio.functional.mask_along_axis's draws (p = 1.0): value = rand * mask_param, min_value = rand * (size - value);
    the masked interval is [long(min_value), long(min_value) + long(value)). Two draws from torch's global CPU generator."""
    if mask_param < 1:
        return 0, 0
    value = torch.rand(1) * mask_param
    min_value = torch.rand(1) * (size - value)
    start = int(min_value.long())
    return start, start + int(value.long())


class SpecBatchProducer:
    """The audio member of the CAV-MAE-style tuples for a whole batch (dataset/dataset.py:296-321, `CAVDataset.__getitem__`):
    pre-computed filterbank arrays [T, F] -> SpecAugment (train + augnois: FrequencyMasking(48), TimeMasking(192), :281-294)
    -> (x - norm_mean) / norm_std -> (train + augnois + noise) x + rand(T, F) * np.random.rand() / 10, rolled by
    np.random.randint(-1024, 1024) frames. Parameters and the noise field are drawn on the host, per sample, in the
    reference's order and from the same generators (torch's CPU generator, numpy's global one); the arithmetic runs in one
    kernel pass (`mla_spec_to_batch`) with torch's roundings."""

    def __init__(self, mode="train", augnois=True, noise=True, skip_norm=False, norm_mean=-5.081, norm_std=4.4849,
                 freqm=48, timem=192, device="cuda"):
        self.train = mode == "train"
        self.augnois, self.noise, self.skip_norm = bool(augnois), bool(noise), bool(skip_norm)
        self.mean, self.std, self.freqm, self.timem = float(norm_mean), float(norm_std), int(freqm), int(timem)
        self.device = torch.device(device)

    def __call__(self, fbanks, out=None):
        x = torch.stack([torch.as_tensor(np.asarray(f), dtype=torch.float32) for f in fbanks])          # [B, T, F]
        B, T, F = x.shape
        params = np.zeros((B, 6), np.int32)
        amp = np.zeros(B, np.float32)
        noise = None
        for b in range(B):
            if self.train and self.augnois:
                f0, f1 = mask_interval(self.freqm, F)              # FrequencyMasking on the transposed [F, T] view
                t0, t1 = mask_interval(self.timem, T)              # TimeMasking
                params[b, :4] = (f0, f1, t0, t1)
            if self.noise and self.train and self.augnois:
                if noise is None:
                    noise = torch.zeros(B, T, F)
                noise[b] = torch.rand(T, F)
                amp[b] = np.float32(np.random.rand())
                params[b, 4] = np.random.randint(-1024, 1024)
                params[b, 5] = 1
        dev = self.device
        return ops.spec_to_batch(x.to(dev), torch.from_numpy(params).to(dev), torch.from_numpy(amp).to(dev),
                                 None if noise is None else noise.to(dev), self.mean, self.std, self.skip_norm, out=out)
