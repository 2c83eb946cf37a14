"""Fixture of the visual dataset transform: the reference's train-mode Compose
(dataset/dataset.py:126-132: RandomResizedCrop(size) -> RandomHorizontalFlip -> ToTensor -> Normalize) executed by
torchvision / Pillow on four small synthetic frames under torch.manual_seed(7), frames stacked as dataset.py:147-156 does.
Run in the build container:  python tests/golden/make_golden_frames.py   -> tests/golden/frames.npz
(the crop / flip parameters torchvision drew are recorded as well: the host mirror must draw the same ones)."""
import os

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
SIZE = 56            # the reference uses 224; 56 keeps the fixture small (the algorithm is size-generic, 224 is tested live)
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def frame(rng, H, W):
    y, x = np.mgrid[0:H, 0:W]
    img = np.stack([127 + 120 * np.sin(x / 7.0 + c) * np.cos(y / 11.0 - c) for c in range(3)], -1)
    img += rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


class Recorder(transforms.RandomResizedCrop):
    """RandomResizedCrop that remembers the parameters it drew (same draws: it calls the parent's get_params)."""
    log = []

    def forward(self, img):
        i, j, h, w = self.get_params(img, self.scale, self.ratio)
        Recorder.log.append([i, j, h, w])
        return transforms.functional.resized_crop(img, i, j, h, w, self.size, self.interpolation, antialias=self.antialias)


class FlipRecorder(transforms.RandomHorizontalFlip):
    log = []

    def forward(self, img):
        flip = bool(torch.rand(1) < self.p)
        FlipRecorder.log.append(flip)
        return transforms.functional.hflip(img) if flip else img


def main():
    rng = np.random.default_rng(0)
    frames = [[frame(rng, 90, 120), frame(rng, 90, 120)], [frame(rng, 64, 80), frame(rng, 130, 97)]]
    tf = transforms.Compose([Recorder(SIZE), FlipRecorder(), transforms.ToTensor(), transforms.Normalize(MEAN, STD)])
    torch.manual_seed(7)
    out = torch.stack([torch.cat([tf(Image.fromarray(f)).unsqueeze(1).float() for f in s], 1) for s in frames])
    params = np.array([p + [int(f)] for p, f in zip(Recorder.log, FlipRecorder.log)], np.int32)
    np.savez_compressed(os.path.join(HERE, "frames.npz"), f00=frames[0][0], f01=frames[0][1], f10=frames[1][0],
                        f11=frames[1][1], params=params, size=np.int32(SIZE), mean=np.float32(MEAN), std=np.float32(STD),
                        seed=np.int32(7), expected=out.numpy())
    print("wrote frames.npz", out.shape, params.tolist())


if __name__ == "__main__":
    main()
