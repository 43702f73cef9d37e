"""CPU checks of the conv-trunk host logic (no GPU): BatchNorm folding and the packed (kh, kw, c) weight layout that
mb200_im2col3x3 + mb200_gemm consume reproduce conv2d + eval BatchNorm of the oracle restatement, block by block; the
ModifiedResNet module exposes openai/CLIP's state-dict names; the encoder factory keeps the reference's dispatch
(magma/image_encoders.py:48-91)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))

from magma_b200 import image_encoders as ie  # noqa: E402
from oracle import magma_oracle as O  # noqa: E402
from tools.model_check import im2col3x3_reference  # noqa: E402


def _net_and_weights(cfg, seed):
    w = O.init_resnet_weights(cfg, seed=seed)
    net = ie.B200ModifiedResNet(cfg.rn_layers, cfg.rn_width, cfg.rn_image, device="cpu")
    missing, unexpected = net.load_state_dict({k[len("image_prefix.enc."):]: v for k, v in w.items()}, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing)
    return net, w


def test_state_dict_names_follow_openai_clip():
    cfg = O.OracleConfig(rn_width=16, rn_layers=(1, 2, 1, 1), rn_image=64)
    net, w = _net_and_weights(cfg, 0)
    names = set(net.state_dict())
    for k in ("conv1.weight", "bn3.running_var", "layer1.0.downsample.0.weight", "layer1.0.downsample.1.running_mean",
              "layer2.1.conv2.weight", "layer4.0.bn3.bias"):
        assert k in names, k
    assert "layer2.1.downsample.0.weight" not in names  # only the first block of a stage has a projection shortcut
    assert net.output_dim == 16 * 32 and net.input_resolution == 64


def test_folded_packed_weights_reproduce_conv_bn():
    """im2col (torch statement of the kernel's layout) x packed weights + folded bias == conv2d -> eval BatchNorm."""
    cfg = O.OracleConfig(rn_width=16, rn_layers=(1, 1, 1, 1), rn_image=64)
    net, w = _net_and_weights(cfg, 1)
    g = torch.Generator().manual_seed(0)
    # stem conv1: 3 input channels padded to 8, stride 2
    x = torch.randn(2, 3, 16, 16, generator=g)
    want = O._bn_eval(F.conv2d(x, w["image_prefix.enc.conv1.weight"], stride=2, padding=1), w, "image_prefix.enc.bn1")
    x8 = torch.zeros(2, 16, 16, 8)
    x8[..., :3] = x.permute(0, 2, 3, 1)
    s = net.bn1.weight / torch.sqrt(net.bn1.running_var + net.bn1.eps)
    wf = torch.cat([net.conv1.weight * s[:, None, None, None], torch.zeros(8, 5, 3, 3)], 1).permute(0, 2, 3, 1).reshape(8, -1)
    got = (im2col3x3_reference(x8, 2) @ wf.t() + (net.bn1.bias - net.bn1.running_mean * s)).view(2, 8, 8, 8).permute(0, 3, 1, 2)
    assert torch.allclose(got, want, atol=1e-5)
    packed, bias = ie.fold_conv_bn(net.conv1.weight, net.bn1, pad_cin_to=8)
    assert packed.shape == (8, 72) and packed.dtype == torch.bfloat16
    assert torch.allclose(packed.float(), wf, atol=2e-2, rtol=1e-2) and torch.allclose(bias.float(), net.bn1.bias - net.bn1.running_mean * s, atol=1e-2)
    # a 3x3 stride-1 bottleneck conv and a 1x1 conv
    blk = net.layer2[0]
    x = torch.randn(2, 32, 8, 8, generator=g)
    p = "image_prefix.enc.layer2.0"
    want = O._bn_eval(F.conv2d(x, w[f"{p}.conv2.weight"], padding=1), w, f"{p}.bn2")
    packed, bias = ie.fold_conv_bn(blk.conv2.weight, blk.bn2)
    got = (im2col3x3_reference(x.permute(0, 2, 3, 1).contiguous(), 1) @ packed.float().t() + bias.float())
    assert ((got.view(2, 8, 8, 32).permute(0, 3, 1, 2) - want).norm() / want.norm()) < 1e-2
    want = O._bn_eval(F.conv2d(x, w[f"{p}.conv3.weight"]), w, f"{p}.bn3")
    packed, bias = ie.fold_conv_bn(blk.conv3.weight, blk.bn3)
    got = x.permute(0, 2, 3, 1).reshape(-1, 32) @ packed.float().t() + bias.float()
    assert ((got.view(2, 8, 8, 128).permute(0, 3, 1, 2) - want).norm() / want.norm()) < 1e-2


def test_oracle_trunk_geometry_matches_reference_tables():
    """RN50x16 at 384 px gives 12x12 = 144 tokens of 3072 (magma/image_prefix.py:13,20); RN50x4 gives 2560 (:19)."""
    assert ie.RESNET_CONFIGS["clip_resnet_large"] == ((6, 8, 18, 8), 96, 384)
    assert ie.RESNET_CONFIGS["clip_resnet"][1] * 32 == 2560
    cfg = O.OracleConfig(rn_width=16, rn_layers=(1, 1, 1, 1), rn_image=96)
    y = O.resnet_forward(torch.randn(1, 3, 96, 96), O.init_resnet_weights(cfg, 2), cfg)
    assert y.shape == (1, 9, 512)
    specs = O.resnet_block_specs(O.OracleConfig())
    assert len(specs) == 40 and specs[0][1:] == (96, 96, 1) and specs[6][1:] == (384, 192, 2) and specs[-1][1:] == (3072, 768, 1)


def test_factory_dispatch():
    with pytest.raises(NotImplementedError):
        ie.get_image_encoder("nfresnet50")
    with pytest.raises(ValueError):
        ie.get_image_encoder("clip_unknown")
    assert isinstance(ie.get_image_encoder("clip_resnet", device="cpu"), ie.B200ModifiedResNet)


def test_checkpoint_key_adapter_round_trip():
    """Fork-named (GPT-Neo lineage) LM keys map onto the HF GPT-J names and back; CLIP's own mlp.c_fc / c_proj names
    under image_prefix.enc are left alone; causal-mask buffers are dropped (magma_b200/checkpoint.py)."""
    from magma_b200.checkpoint import convert_reference_state_dict, to_reference_names

    cfg = O.OracleConfig(d=64, n_layer=2, n_head=2, rotary_dim=16, vocab=128, vit_width=32, vit_layers=1, vit_heads=2,
                         vit_patch=8, vit_image=16, vit_mlp=64, enc_out_dim=32,
                         attn_adapter={"adapter_type": "normal", "downsample_factor": 8})
    w = O.init_weights(cfg)
    ref = to_reference_names(w)
    assert "lm.transformer.h.0.attn.attn_block.attention.q_proj.weight" in ref
    assert "lm.transformer.h.1.mlp.0.c_fc.weight" in ref and "lm.transformer.h.1.mlp.0.fc_in.weight" not in ref
    assert "image_prefix.enc.transformer.resblocks.0.mlp.c_fc.weight" in ref
    ref["lm.transformer.h.0.attn.attn_block.attention.bias"] = torch.zeros(1)
    ref["lm.transformer.h.0.attn.attn_block.attention.masked_bias"] = torch.zeros(1)
    back, report = convert_reference_state_dict(ref)
    assert set(back) == set(w) and not report["collisions"] and len(report["dropped"]) == 2
    assert all(torch.equal(back[k], w[k]) for k in w)
