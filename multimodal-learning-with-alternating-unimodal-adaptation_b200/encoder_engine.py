"""Encoder execution engine: runs a ResNet parameter container (backbone.py) on the device.

INTERIM (round 1, first slice): the convolution / batch-norm / pooling arithmetic is issued
as library calls (cuDNN through torch.nn.functional) while the hand-written sm_100a
implicit-GEMM kernels are brought up; `BACKEND` names what is running and bench.py reports it.
The head, GS projection and fusion kernels are already native (libmla_b200.so).
"""
import torch
import torch.nn.functional as F

BACKEND = "cudnn-interim"


def _bn(x, bn, training):
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)


def resnet_feature_map(net, x):
    if not x.is_cuda:
        raise RuntimeError("mla_b200 encoders run on CUDA only (no CPU fallback); got %s" % x.device)
    training = net.training
    if net.modality == "visual":
        B, C, T, H, W = x.shape
        x = x.permute(0, 2, 1, 3, 4).contiguous().view(B * T, C, H, W)
    x = F.conv2d(x, net.conv1.weight, None, 2, 3)
    x = F.relu(_bn(x, net.bn1, training))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        for blk in getattr(net, "layer%d" % li):
            identity = x
            out = F.conv2d(x, blk.conv1.weight, None, blk.stride, 1)
            out = F.relu(_bn(out, blk.bn1, training))
            out = F.conv2d(out, blk.conv2.weight, None, 1, 1)
            out = _bn(out, blk.bn2, training)
            if blk.downsample is not None:
                identity = F.conv2d(x, blk.downsample[0].weight, None, blk.stride, 0)
                identity = _bn(identity, blk.downsample[1], training)
            x = F.relu(out + identity)
    return x


def resnet_pooled(net, x):
    batch = x.shape[0]
    fm = resnet_feature_map(net, x)                       # [B*T, 512, h, w]
    n, c, h, w = fm.shape
    return fm.view(batch, n // batch, c, h * w).mean(dim=(1, 3))
