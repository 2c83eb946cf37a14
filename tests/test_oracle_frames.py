"""CPU: the visual dataset transform (dataset/dataset.py:123-161). The arithmetic lives in libraries the
reference calls (torchvision.transforms over Pillow); both are installed, so the oracle's numpy restatement and the host
mirror's parameter sampling are pinned by running the real thing, plus a committed fixture (tests/golden/frames.npz,
tests/golden/make_golden_frames.py) in case a future Pillow changes its resampler."""
import os

import numpy as np
import pytest
import torch

from oracle import mla_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def _frame(rng, H, W):
    """A smooth image with edges (noise alone would hide coefficient errors behind rounding)."""
    y, x = np.mgrid[0:H, 0:W]
    img = np.stack([127 + 120 * np.sin(x / 7.0 + c) * np.cos(y / 11.0 - c) for c in range(3)], -1)
    img += rng.normal(0, 12, img.shape)
    img[H // 3: H // 2, W // 4: W // 2] = 255 - img[H // 3: H // 2, W // 4: W // 2]
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("H,W,size", [(360, 480, 224), (224, 224, 224), (100, 130, 224), (513, 300, 224), (37, 41, 64),
                                      (225, 223, 224), (720, 1280, 224)])
def test_resize_restatement_is_pillow_bit_for_bit(H, W, size):
    from PIL import Image
    img = _frame(np.random.default_rng(H + W), H, W)
    ref = np.asarray(Image.fromarray(img).resize((size, size), Image.BILINEAR))
    assert np.array_equal(orc.pil_resize_u8(img, size, size), ref)


def test_transform_restatement_matches_the_reference_compose():
    """dataset.py:133-138 (test mode) and :126-132 with fixed crop parameters (train mode), through torchvision itself."""
    from PIL import Image
    from torchvision import transforms
    from torchvision.transforms import functional as F
    rng = np.random.default_rng(3)
    frames = [[_frame(rng, 120, 160), _frame(rng, 120, 160)], [_frame(rng, 97, 131), _frame(rng, 200, 150)]]
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    test_tf = transforms.Compose([transforms.Resize(size=(224, 224)), transforms.ToTensor(), transforms.Normalize(mean, std)])
    ref = torch.stack([torch.stack([test_tf(Image.fromarray(f)) for f in s], 1) for s in frames])
    params = [(0, 0, f.shape[0], f.shape[1], False) for s in frames for f in s]
    assert np.array_equal(orc.frames_to_tensor(frames, params, 224, mean, std), ref.numpy())
    params = [(10, 20, 80, 100, True), (0, 0, 120, 160, False), (5, 3, 90, 60, True), (50, 40, 120, 100, False)]
    tail = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean, std)])
    ref = []
    for s, f in enumerate([f for smp in frames for f in smp]):
        i, j, h, w, flip = params[s]
        img = F.resized_crop(Image.fromarray(f), i, j, h, w, [224, 224])        # RandomResizedCrop.forward
        if flip:
            img = F.hflip(img)                                                   # RandomHorizontalFlip.forward
        ref.append(tail(img))
    ref = torch.stack(ref).view(2, 2, 3, 224, 224).permute(0, 2, 1, 3, 4)
    assert np.array_equal(orc.frames_to_tensor(frames, params, 224, mean, std), ref.numpy())


def test_crop_parameter_sampling_follows_torchvision_draw_for_draw():
    from torchvision import transforms
    from mla_b200.dataset import random_resized_crop_params
    for seed, (H, W) in enumerate([(360, 480), (64, 512), (500, 40), (224, 224)]):
        torch.manual_seed(seed)
        ours = [random_resized_crop_params(H, W) + (bool(torch.rand(1) < 0.5),) for _ in range(20)]
        torch.manual_seed(seed)
        dummy = torch.zeros(3, H, W)
        ref = []
        for _ in range(20):
            p = transforms.RandomResizedCrop.get_params(dummy, [0.08, 1.0], [3.0 / 4.0, 4.0 / 3.0])
            ref.append(tuple(p) + (bool(torch.rand(1) < 0.5),))
        assert ours == ref


def test_committed_fixture():
    g = np.load(os.path.join(HERE, "golden", "frames.npz"))
    frames = [[g["f00"], g["f01"]], [g["f10"], g["f11"]]]
    params = [tuple(int(v) for v in p[:4]) + (bool(p[4]),) for p in g["params"]]
    out = orc.frames_to_tensor(frames, params, int(g["size"]), g["mean"], g["std"])
    assert np.array_equal(out, g["expected"])


@pytest.mark.parametrize("H,W,size", [(360, 480, 224), (480, 360, 224), (224, 224, 224), (301, 500, 224), (256, 341, 256),
                                      (1000, 700, 256), (230, 229, 224)])
def test_bicubic_resize_center_crop_restatement_matches_the_reference_compose(H, W, size):
    """dataset.py:251-256 (CAVDataset.preprocess) / :414-421 (M3AEDataset.preprocess_test) through torchvision itself."""
    import PIL
    from PIL import Image
    from torchvision import transforms
    from mla_b200.dataset import resize_center_crop_params
    img = _frame(np.random.default_rng(H * 3 + W), H, W)
    tf = transforms.Compose([transforms.Resize(size, interpolation=transforms.InterpolationMode.BICUBIC),
                             transforms.CenterCrop(size)])
    ref = np.asarray(tf(Image.fromarray(img)))
    assert np.array_equal(orc.resize_center_crop_u8(img, size), ref)
    rh, rw, oy, ox = resize_center_crop_params(H, W, size)
    full = np.asarray(Image.fromarray(img).resize((rw, rh), PIL.Image.BICUBIC))
    assert np.array_equal(full[oy:oy + size, ox:ox + size], ref)               # the host mirror's window arithmetic


def test_audio_pipeline_restatement_and_host_draws_match_torchaudio():
    """dataset.py:281-294, 301-321 executed with torchaudio vs the oracle's numpy restatement fed with the parameters the
    host mirror draws (mla_b200.dataset.mask_interval + numpy's generator) under the same seeds."""
    import torchaudio
    from mla_b200.dataset import mask_interval
    rng = np.random.default_rng(5)
    fbank = rng.normal(-5, 4, (1024, 128)).astype(np.float32)
    for seed in range(4):
        torch.manual_seed(seed); np.random.seed(seed)
        fb = torch.transpose(torch.tensor(fbank), 0, 1).unsqueeze(0)
        fb = torchaudio.transforms.TimeMasking(192)(torchaudio.transforms.FrequencyMasking(48)(fb))
        ref = torch.transpose(fb.squeeze(0), 0, 1)
        ref = (ref - (-5.081)) / (4.4849)
        ref = ref + torch.rand(ref.shape[0], ref.shape[1]) * np.random.rand() / 10
        ref = torch.roll(ref, np.random.randint(-1024, 1024), 0)
        torch.manual_seed(seed); np.random.seed(seed)
        f0, f1 = mask_interval(48, 128)
        t0, t1 = mask_interval(192, 1024)
        noise = torch.rand(1024, 128).numpy()
        amp = np.float32(np.random.rand())
        shift = np.random.randint(-1024, 1024)
        out = orc.spec_augment(fbank, (f0, f1, t0, t1, shift, 1), amp, noise, -5.081, 4.4849)
        assert np.array_equal(out, ref.numpy())
