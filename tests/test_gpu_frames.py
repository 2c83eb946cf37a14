"""GPU: the visual dataset tuple producer — csrc/frame_producer.cu through the C ABI
(`FrameBatchProducer` -> ops.frames_to_batch -> mla_frames_to_batch) against torchvision / Pillow run live on the host,
against the committed fixture of the reference's train-mode Compose (dataset/dataset.py:126-132), and against the numpy
oracle. Integer resampling + IEEE fp32 normalisation: the bar is BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

from oracle import mla_oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _frame(rng, H, W):
    y, x = np.mgrid[0:H, 0:W]
    img = np.stack([127 + 120 * np.sin(x / 7.0 + c) * np.cos(y / 11.0 - c) for c in range(3)], -1)
    img += rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.fixture(scope="module")
def producer_cls(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from mla_b200.dataset import FrameBatchProducer
    return FrameBatchProducer


@pytest.mark.parametrize("shapes", [[(360, 480)] * 4, [(224, 224), (100, 130), (513, 300), (37, 41)],
                                    [(720, 1280), (225, 223), (1080, 1920), (64, 3000)]])
def test_eval_transform_equals_torchvision_bit_for_bit(producer_cls, shapes):
    """dataset.py:133-138: Resize((224, 224)) -> ToTensor -> Normalize, two frames per sample."""
    from PIL import Image
    from torchvision import transforms
    rng = np.random.default_rng(len(shapes) + shapes[0][0])
    frames = [[_frame(rng, *shapes[0]), _frame(rng, *shapes[1])], [_frame(rng, *shapes[2]), _frame(rng, *shapes[3])]]
    tf = transforms.Compose([transforms.Resize(size=(224, 224)), transforms.ToTensor(), transforms.Normalize(MEAN, STD)])
    ref = torch.stack([torch.cat([tf(Image.fromarray(f)).unsqueeze(1).float() for f in s], 1) for s in frames])   # :147-156
    out = producer_cls(224, "test")(frames)
    assert out.shape == (2, 3, 2, 224, 224) and out.is_cuda
    assert torch.equal(out.cpu(), ref)


def test_train_transform_reproduces_the_seeded_reference_compose(producer_cls):
    """Fixture: RandomResizedCrop -> RandomHorizontalFlip -> ToTensor -> Normalize under torch.manual_seed(7), executed by
    torchvision (tests/golden/make_golden_frames.py). Same seed here: same crops, same flips, same bits."""
    g = np.load(os.path.join(HERE, "golden", "frames.npz"))
    frames = [[g["f00"], g["f01"]], [g["f10"], g["f11"]]]
    prod = producer_cls(int(g["size"]), "train", g["mean"], g["std"])
    torch.manual_seed(int(g["seed"]))
    drawn = prod._params([f for s in frames for f in s])
    assert [list(p[:4]) + [int(p[4])] for p in drawn] == g["params"].tolist()
    torch.manual_seed(int(g["seed"]))
    out = prod(frames)
    assert np.array_equal(out.cpu().numpy(), g["expected"])


def test_explicit_crops_flips_and_batch_layout_vs_oracle(producer_cls):
    rng = np.random.default_rng(11)
    B, T = 5, 3
    frames = [[_frame(rng, int(rng.integers(40, 400)), int(rng.integers(40, 400))) for _ in range(T)] for _ in range(B)]
    params = []
    for s in frames:
        for f in s:
            H, W = f.shape[:2]
            h, w = int(rng.integers(8, H + 1)), int(rng.integers(8, W + 1))
            params.append((int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1)), h, w, bool(rng.integers(0, 2))))
    for size in (224, 96):
        out = producer_cls(size, "train")(frames, params=params)
        assert np.array_equal(out.cpu().numpy(), orc.frames_to_tensor(frames, params, size, MEAN, STD))
    again = producer_cls(96, "train")(frames, params=params)
    assert torch.equal(out, again)                                              # deterministic


def test_identity_resize_and_errors(producer_cls):
    rng = np.random.default_rng(5)
    f = _frame(rng, 224, 224)
    out = producer_cls(224, "test")([[f]])
    x = torch.from_numpy(f).permute(2, 0, 1).float().div(255)
    ref = (x - torch.tensor(MEAN)[:, None, None]) / torch.tensor(STD)[:, None, None]
    assert torch.equal(out.cpu()[0, :, 0], ref)                                 # scale 1: weights (1, 0) -> the input pixels
    prod = producer_cls(32, "train")
    with pytest.raises(RuntimeError):
        prod([[f]], params=[(200, 0, 100, 100, False)])                         # crop leaves the frame
    with pytest.raises(RuntimeError):
        producer_cls(8, "test")([[f]])                                          # 224 / 8 = 28 > 15.5: outside the tap budget
    with pytest.raises(RuntimeError):
        producer_cls(24, "test", interpolation="bicubic")([[f]])                # 224 / 24 = 9.3 > 7.5 for the bicubic filter
    with pytest.raises(ValueError):
        prod([[f.astype(np.float32)]])


@pytest.mark.parametrize("shapes,size", [([(360, 480), (480, 360), (224, 224), (301, 500)], 224),
                                         ([(256, 341), (1000, 700), (230, 229), (260, 256)], 256)])
def test_bicubic_resize_center_crop_equals_torchvision_bit_for_bit(producer_cls, shapes, size):
    """dataset.py:251-256 (CAVDataset.preprocess, size 224) and :414-421 (M3AEDataset.preprocess_test, size 256):
    Resize(s, BICUBIC) -> CenterCrop(s) -> ToTensor -> Normalize. Only the crop window is computed on the GPU."""
    from PIL import Image
    from torchvision import transforms
    rng = np.random.default_rng(size + shapes[1][0])
    frames = [[_frame(rng, *hw)] for hw in shapes]
    mean, std = [0.4850, 0.4560, 0.4060], [0.2290, 0.2240, 0.2250]
    tf = transforms.Compose([transforms.Resize(size, interpolation=transforms.InterpolationMode.BICUBIC),
                             transforms.CenterCrop(size), transforms.ToTensor(), transforms.Normalize(mean=mean, std=std)])
    ref = torch.stack([tf(Image.fromarray(s[0])) for s in frames])
    out = producer_cls(size, "center", mean, std, interpolation="bicubic")(frames)
    assert out.shape == (4, 3, 1, size, size)
    assert torch.equal(out.cpu()[:, :, 0], ref)


def _reference_cav_audio(fbank, train, augnois, noise, skip_norm, norm_mean=-5.081, norm_std=4.4849):
    """dataset/dataset.py:281-294 + 301-321 restated line by line on the CPU with torchaudio itself."""
    import torchaudio
    fbank = torch.tensor(fbank)
    if train and augnois:
        freqm = torchaudio.transforms.FrequencyMasking(48)
        timem = torchaudio.transforms.TimeMasking(192)
        fb = torch.transpose(fbank, 0, 1).unsqueeze(0)
        fb = timem(freqm(fb))
        fbank = torch.transpose(fb.squeeze(0), 0, 1)
    if not skip_norm:
        fbank = (fbank - norm_mean) / (norm_std)
    if noise and train and augnois:
        fbank = fbank + torch.rand(fbank.shape[0], fbank.shape[1]) * np.random.rand() / 10
        fbank = torch.roll(fbank, np.random.randint(-1024, 1024), 0)
    return fbank


@pytest.mark.parametrize("train,augnois,noise,skip_norm", [(True, True, True, False), (True, True, False, False),
                                                            (False, True, True, False), (True, False, True, True)])
def test_spec_producer_equals_the_reference_audio_pipeline_bit_for_bit(built_lib, train, augnois, noise, skip_norm):
    from mla_b200.dataset import SpecBatchProducer
    rng = np.random.default_rng(3)
    fbanks = [rng.normal(-5, 4, (1024, 128)).astype(np.float32) for _ in range(3)]
    torch.manual_seed(11); np.random.seed(11)
    ref = torch.stack([_reference_cav_audio(f, train, augnois, noise, skip_norm) for f in fbanks])
    torch.manual_seed(11); np.random.seed(11)
    out = SpecBatchProducer("train" if train else "test", augnois, noise, skip_norm)(fbanks)
    assert out.shape == (3, 1024, 128)
    assert torch.equal(out.cpu(), ref)
    # the kernel against the numpy oracle on explicit parameters (masks at the borders, negative and zero shifts)
    from mla_b200 import ops
    params = np.array([[0, 48, 1000, 1024, -1023, 1], [127, 128, 0, 1, 0, 1], [10, 10, 5, 5, 512, 0]], np.int32)
    amp = np.array([0.73, 0.01, 0.5], np.float32)
    nz = rng.random((3, 1024, 128)).astype(np.float32)
    x = np.stack(fbanks)
    got = ops.spec_to_batch(torch.from_numpy(x).cuda(), torch.from_numpy(params).cuda(), torch.from_numpy(amp).cuda(),
                            torch.from_numpy(nz).cuda(), -5.081, 4.4849, skip_norm).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], orc.spec_augment(x[b], params[b], amp[b], nz[b], -5.081, 4.4849, skip_norm))
