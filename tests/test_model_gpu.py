"""GPU parity tests of the model-level runtime (C ABI mb200_gptj_forward/backward, mb200_vit_forward, decode) and of
the drop-in Python API (Magma.forward / embed / generate, B200Engine) against the CPU oracle on identical seeded
inputs, and against the golden fixtures produced by the REFERENCE ITSELF (tests/golden/, oracle/make_golden.py).

Tolerances (bf16 kernels vs fp32 oracle): |dloss| < 2e-2; logits rel-Frobenius < 3e-2; gradients
rel-Frobenius < 3e-2 when every adapter ReLU is decided away from zero, < 1.5e-1 with free ReLU masks (a flipped mask
entry is an O(1) relative error in that entry; the flip rate is ~2^-8, so ~sqrt(2^-8) relative Frobenius error is the
expected bf16-vs-fp32 disagreement, not a kernel defect). Integer results (labels, greedy token ids) are exact."""
import os

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", ["lm", "lm_variants", "vit", "resnet", "magma", "generate", "sampling"])
def test_model_group(group):
    import torch

    from tools import model_check

    assert getattr(model_check, "group_" + group)(torch.device("cuda:0"))


@pytest.mark.parametrize("tag", ["v1_mlp_normal", "v2_mlp_attn_normal", "parallel", "no_adapters"])
def test_magma_matches_reference_golden(golden_dir, tag):
    import torch

    from _gpu_util import build_magma_from_weights, gpu_device, rel
    from conftest import oracle_cfg_from_record

    dev = gpu_device()
    rec = torch.load(os.path.join(golden_dir, f"magma_{tag}.pt"), weights_only=False)
    cfg = oracle_cfg_from_record(rec)
    # the reference ran in fp32; the CUDA path stores weights/activations in bf16
    model = build_magma_from_weights(rec["weights"], cfg, rec["adapter_config"], rec["S"], dev, vit_name="clip_vit_golden")
    model.eval()
    out = model(rec["images"].to(dev), rec["captions"].to(dev))
    assert abs(float(out.loss) - float(rec["loss"])) < 2e-2
    assert rel(out.logits, rec["logits"]) < 3e-2
    if rec["grads"]:
        out.loss.backward()
        sd = dict(model.named_parameters())
        # Down-projection (adapter.0) gradients pass through the ReLU mask of the bottleneck: where a pre-activation is
        # within the bf16 error of zero the mask differs from the fp32 reference's, and in the tiny golden configuration
        # (8-16 hidden units x 64 tokens) a handful of flipped (token, unit) entries is 10-20 % of the tensor. The same
        # kernels are held to 5e-3 with decided masks in tools/model_check.py::group_lm (test_model_group[lm]).
        bad = {}
        for k, gref in rec["grads"].items():
            assert sd[k].grad is not None, k
            err = rel(sd[k].grad, gref)
            if err >= (3e-1 if ".adapter.0." in k else 1.5e-1):
                bad[k] = round(err, 4)
        assert not bad, bad


def test_vit_embed_generate_match_reference_golden(golden_dir):
    import torch

    from _gpu_util import build_magma_from_weights, gpu_device, rel
    from conftest import oracle_cfg_from_record

    dev = gpu_device()
    rec = torch.load(os.path.join(golden_dir, "magma_v1_mlp_normal.pt"), weights_only=False)
    cfg = oracle_cfg_from_record(rec)
    model = build_magma_from_weights(rec["weights"], cfg, rec["adapter_config"], rec["S"], dev, vit_name="clip_vit_golden")
    model.eval()
    feats = model.image_prefix.enc(rec["images"].to(dev))
    assert rel(feats, rec["enc_feats"]) < 3e-2
    emb = model.embed([rec["images"].to(dev), rec["text"].to(dev)])
    assert emb.shape == rec["embeddings"].shape and rel(emb, rec["embeddings"]) < 3e-2
    toks = model.generate(rec["embeddings"].to(dev).to(torch.bfloat16), max_steps=10, temperature=0.0, decode=False).cpu()
    ref = rec["greedy_tokens"]
    n = min(toks.shape[1], ref.shape[1])
    agree = (toks[:, :n] == ref[:, :n]).float().mean().item()
    # token ids are exact up to the first bf16 near-tie between the top two logits
    assert torch.equal(toks[:, : rec["embeddings"].shape[1] + 1], ref[:, : rec["embeddings"].shape[1] + 1])
    assert agree > 0.85, agree


def test_magma_forward_asserts_like_the_reference():
    import torch

    from _gpu_util import build_magma_from_weights, gpu_device
    from oracle import magma_oracle as O
    from tools.model_check import small_cfg

    dev = gpu_device()
    cfg = small_cfg()
    w = {k: v.to(torch.bfloat16).float() for k, v in O.init_weights(cfg, seed=3).items()}
    model = build_magma_from_weights(w, cfg, {"mlp": {"adapter_type": "normal", "downsample_factor": 4}}, 32, dev)
    images, captions = O.synthetic_batch(cfg, 2, 32, seed=1)
    with pytest.raises(AssertionError, match="captions"):
        model(images.to(dev), None)
    with pytest.raises(AssertionError, match="padded to sequence length"):
        model(images.to(dev), captions[:, :20].to(dev))
    with pytest.raises(AssertionError, match="not both"):
        model(images.to(dev), captions.to(dev), input_embeddings=torch.zeros(2, 2, cfg.d, device=dev))
    with pytest.raises(ValueError, match="already added"):
        model.add_adapters(location="mlp")
    with pytest.raises(ValueError, match="Expected 2d or 4d"):
        model.embed([torch.zeros(3, device=dev)])


def test_full_size_step_properties():
    """BASELINE config-2 sizes (GPT-J-6B + ViT-L/14, B=8, S=128): size-independent properties instead of an oracle
    run — loss ~ ln(V) at random init, finite gradients on every trainable tensor, gradient accumulation is linear,
    the fused optimizer step lowers the loss on the same batch, and a frozen LM weight never changes."""
    import math

    import torch

    from magma_b200.config import MultimodalConfig
    from magma_b200.magma import Magma
    from magma_b200.train_loop import B200Engine

    dev = torch.device("cuda:0")
    mc = MultimodalConfig(batch_size=8, train_steps=1, encoder_name="clip_vit_large",
                          adapter_config={"mlp": {"adapter_type": "normal", "downsample_factor": 4}}, image_seq_len=2,
                          image_embed_dropout_prob=0.0, use_image_embed_layernorm=True, image_size=224, seq_len=128,
                          lr=1e-3, warmup_num_steps=2)
    model = Magma(mc, device=dev, init_seed=0)
    model.train()
    B, S = 8, 128
    g = torch.Generator().manual_seed(0)
    images = torch.randn(B, 3, 224, 224, generator=g).to(dev).to(torch.bfloat16)
    captions = torch.randint(0, 50256, (B, S), generator=g)
    captions[:, 90:] = 50256
    captions = captions.to(dev)
    out = model(images, captions)
    assert out.logits.shape == (B, S, 50258)
    assert abs(float(out.loss) - math.log(50258)) < 1.0
    out.loss.backward()
    n_train = 0
    for n, p in model.named_parameters():
        if p.requires_grad:
            n_train += p.numel()
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert n_train == 28 * 8_393_728 + 768 * 8192 + 8192 + 2 * 4096  # 241.3 M trainable
    g1 = model.arena.grad.clone()
    model(images, captions).loss.backward()  # grads live -> accumulate
    assert ((model.arena.grad - 2 * g1).norm() / (2 * g1).norm()).item() < 1e-2
    frozen_before = model.lm.transformer.h[5].attn.out_proj.weight.clone()
    for p in model.parameters():
        p.grad = None
    model.arena.grad.zero_()
    eng = B200Engine(model, mc)
    losses = []
    for _ in range(3):
        o = eng(images, captions)
        eng.backward(o.loss)
        eng.step()
        losses.append(float(o.loss))
    assert losses[-1] < losses[0], losses
    assert torch.equal(frozen_before, model.lm.transformer.h[5].attn.out_proj.weight)
