"""ResNet-18 audio / visual encoders (reference models/backbone.py:54-160, 211-213).

The module tree is a PARAMETER CONTAINER with the reference's names and creation order
(conv1, bn1, layerN.M.{conv1,bn1,conv2,bn2,downsample.0,downsample.1}) so that
  * released checkpoints load (state-dict keys, incl. BN buffers / num_batches_tracked), and
  * seeding reproduces the reference's initial weights bit for bit (same RNG consumption).
The arithmetic does not go through these nn.Modules' forward: `ResNet.forward` hands the
parameters to the encoder engine (encoder_engine.py), which runs the convolution / BN /
pooling kernels of libmla_b200.so and implements the backward pass.
"""
import torch
import torch.nn as nn

from . import encoder_engine

_STAGES = ((64, 1), (128, 2), (256, 2), (512, 2))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    def __init__(self, modality):
        super().__init__()
        if modality not in ("audio", "visual"):
            raise NotImplementedError("Incorrect modality, should be audio or visual but got {}".format(modality))
        self.modality = modality
        self.conv1 = nn.Conv2d(1 if modality == "audio" else 3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for i, (planes, stride) in enumerate(_STAGES, start=1):
            down = None
            if stride != 1 or inplanes != planes:
                down = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
            blocks = [BasicBlock(inplanes, planes, stride, down), BasicBlock(planes, planes)]
            setattr(self, "layer%d" % i, nn.Sequential(*blocks))
            inplanes = planes
        # backbone.py:101-106 — same order as nn.Module.modules()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.normal_(m.weight, mean=1, std=0.02)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        """audio: [B,1,H,W]; visual: [B,3,T,H,W] (frames folded into the batch, backbone.py:144-147).
        Returns the layer4 feature map in the reference's layout [N,512,h,w]."""
        return encoder_engine.resnet_feature_map(self, x)

    def pooled(self, x, frames_per_sample=1):
        """Feature map + global average pool fused (basic_model.py:56-65): returns [B,512]."""
        return encoder_engine.resnet_pooled(self, x)


def resnet18(modality, **kwargs):
    return ResNet(modality)
