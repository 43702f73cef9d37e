"""Partial pin of the conv-trunk oracle (oracle/magma_oracle.py::resnet_forward, "parity unpinned": openai/CLIP is not
vendored). torchvision IS in the build image, and CLIP's Bottleneck with stride 1 is structurally torchvision's
`Bottleneck` (conv1x1-bn-relu, conv3x3-bn-relu, conv1x1-bn, + identity or conv1x1-bn downsample, relu): for a trunk
whose blocks all have stride 1 the oracle must agree with torchvision's independent implementation run on the same
weights. What stays unpinned is only what CLIP changes relative to torchvision: the 3-conv stem and the anti-aliasing
average pools of the stride-2 blocks (restated from the published architecture)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import magma_oracle as O

tv = pytest.importorskip("torchvision.models.resnet")


def test_stride1_bottleneck_stage_matches_torchvision():
    cfg = O.OracleConfig(rn_width=16, rn_layers=(3,), rn_image=32)
    pre = "image_prefix.enc"
    w = O.init_resnet_weights(cfg, seed=3, pre=pre)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(2, 3, 32, 32, generator=g)
    got = O.resnet_forward(images, w, cfg, pre=pre)  # [b, hw, 4*width]

    def bn(name, c):
        m = nn.BatchNorm2d(c)
        m.load_state_dict({k: w[f"{name}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}, strict=False)
        return m.eval()

    x = images
    for i, stride, co in ((1, 2, 8), (2, 1, 8), (3, 1, 16)):  # CLIP's stem, straight from torch.nn
        x = F.relu(bn(f"{pre}.bn{i}", co)(F.conv2d(x, w[f"{pre}.conv{i}.weight"], stride=stride, padding=1)))
    x = F.avg_pool2d(x, 2)
    inpl = 16
    for b in range(3):  # layer1: stride 1 everywhere -> torchvision's Bottleneck, unmodified
        p = f"{pre}.layer1.{b}"
        down = None
        if inpl != 64:
            conv = nn.Conv2d(inpl, 64, 1, bias=False)
            conv.weight.data.copy_(w[f"{p}.downsample.0.weight"])
            down = nn.Sequential(conv, bn(f"{p}.downsample.1", 64))
        blk = tv.Bottleneck(inpl, 16, stride=1, downsample=down).eval()
        for i, c in ((1, 16), (2, 16), (3, 64)):
            getattr(blk, f"conv{i}").weight.data.copy_(w[f"{p}.conv{i}.weight"])
            setattr(blk, f"bn{i}", bn(f"{p}.bn{i}", c))
        with torch.no_grad():
            x = blk(x)
        inpl = 64
    want = x.reshape(2, 64, -1).permute(0, 2, 1)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
