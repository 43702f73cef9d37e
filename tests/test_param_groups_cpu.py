"""Optimizer parameter groups on the flat arena (magma/utils.py:120-215): the image encoder's own learning rate and the
weight-decay exemptions, as contiguous arena runs. Pure host logic."""
import torch
import torch.nn as nn

from magma_b200 import dp
from magma_b200.utils import no_weight_decay_names


def layout(numels):
    offs, total = dp.arena_layout(numels)
    return offs, total


def test_default_config_is_one_run_over_the_whole_arena():
    names = ["lm.transformer.h.1.mlp.1.adapter.0.weight", "lm.transformer.h.1.mlp.1.adapter.0.bias",
             "image_prefix.proj.weight", "image_prefix.ln.weight"]
    numels = [1000, 10, 500, 64]
    offs, total = layout(numels)
    segs = dp.optimizer_segments(names, numels, offs, [False, True, False, True], 1.0, None, 0.0)
    assert segs == [(0, total, 1.0, 0.0)]  # -> ParamArena.adamw_step takes the single fused launch


def test_image_encoder_gets_its_own_rate():
    """MAGMA_v1.yml: lr 8e-4, image_enc_lr 2e-6, freeze_img_encoder false."""
    names = ["lm.transformer.h.0.mlp.1.adapter.0.weight", "image_prefix.enc.conv1.weight",
             "image_prefix.enc.ln_pre.weight", "image_prefix.proj.weight"]
    numels = [128, 70, 64, 200]
    offs, total = layout(numels)
    scale = 2.0e-6 / 8.0e-4
    segs = dp.optimizer_segments(names, numels, offs, [False, False, True, False], 1.0, scale, 0.0)
    assert segs == [(0, 128, 1.0, 0.0), (128, 128 + 128 + 64, scale, 0.0), (320, total, 1.0, 0.0)]
    # runs tile the arena exactly
    assert segs[0][0] == 0 and segs[-1][1] == total and all(a[1] == b[0] for a, b in zip(segs, segs[1:]))


def test_weight_decay_skips_biases_and_layernorms():
    names = ["a.weight", "a.bias", "ln.weight", "ln.bias", "b.weight"]
    numels = [64, 64, 64, 64, 64]
    offs, total = layout(numels)
    segs = dp.optimizer_segments(names, numels, offs, [False, True, True, True, False], 1.0, None, 0.1)
    assert segs == [(0, 64, 1.0, 0.1), (64, 256, 1.0, 0.0), (256, total, 1.0, 0.1)]


def test_no_weight_decay_names_follow_the_reference_rule():
    from magma_b200.adapters import Adapter

    class Toy(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(4, 4)
            self.ln = nn.LayerNorm(4)
            self.emb = nn.Embedding(3, 4)
            self.free = nn.Parameter(torch.zeros(2))       # e.g. CLIP class_embedding / proj: decays
            self.ad = Adapter(dim=8, downsample_factor=2, add_layernorm=True)

    nd = no_weight_decay_names(Toy())
    assert {"lin.bias", "ln.weight", "ln.bias", "emb.weight"} <= nd
    assert "lin.weight" not in nd and "free" not in nd
    # adapter with a leading LayerNorm (adapters.py:16-17): LN params and biases exempt, projection weights decay
    ad = {n for n in nd if n.startswith("ad.")}
    assert ad == {"ad.adapter.0.weight", "ad.adapter.0.bias", "ad.adapter.1.bias", "ad.adapter.3.bias"}
