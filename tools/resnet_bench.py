"""Forward time of the CLIP RN50x16 conv trunk (clip_resnet_large, the encoder of MAGMA_v1.yml) at B=8, 384 px."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from magma_b200.image_encoders import RESNET_CONFIGS, B200ModifiedResNet

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "clip_resnet_large"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
layers, width, R = RESNET_CONFIGS[name]
net = B200ModifiedResNet(layers, width, R, device=dev).init_weights(0)
x = torch.randn(B, 3, R, R, device=dev).to(torch.bfloat16)
for _ in range(3):
    y = net(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    y = net(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"[RESNET] {name} B={B} {R}px -> {tuple(y.shape)}: {ms:.2f} ms/forward  ({B / ms * 1e3:.0f} images/s)  "
      f"finite={bool(torch.isfinite(y.float()).all())} mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
