"""INTEGRATION.md section 1, executed: the REFERENCE'S OWN `magma.magma.Magma` class (imported from /root/reference under
oracle/ref_shims.py — import-time stubs only) constructed over THIS package's factories and modules —
`magma_b200.language_model.get_gptj`, `magma_b200.image_prefix.ImagePrefix` (-> `magma_b200.image_encoders`),
`magma_b200.adapters.{Adapter, ParallelAdapter, AdapterWrapper, ParallelAdapterWrapper}` — exactly the import redirection
a maintainer would make. The reference's `__init__` (magma/magma.py:29-100: resize_token_embeddings, transformer.h
get/setattr in add_adapters, the "adapter"-in-name freeze loop), its `forward` (magma.py:238-276, with its own Python
`build_labels`), its `embed` and the reference's own `sampling.generate` then run against our modules; results are held to
the oracle and to this package's own `Magma` class on the same weights.

The kernels are emulated (fixture `emul_ops`: there is no GPU in the build container, and /root/reference does not exist on
the GPU box), so what this pins is the duck-typed surface, not kernel arithmetic. Skipped where the
reference tree is absent."""
import os
import sys

import pytest
import torch

from oracle import magma_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "magma")), reason="reference tree not mounted")


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.fixture
def reference_over_b200(emul_ops, monkeypatch):
    from oracle import ref_shims

    ref = ref_shims.load_reference()
    import magma.config as ref_config
    import magma_b200.adapters as BA
    import magma_b200.image_prefix as BP
    import magma_b200.language_model as BL
    from magma_b200.image_encoders import register_vit
    from test_e2e_dryrun_cpu import oracle_weights, tiny_cfg

    cfg = tiny_cfg()
    register_vit("clip_vit_dropin", cfg.vit_width, cfg.vit_layers, cfg.vit_heads, cfg.vit_patch, cfg.vit_image,
                 cfg.vit_mlp, cfg.enc_out_dim)
    lm_cfg = BL.GPTJConfig(vocab_size=cfg.vocab + 4, hidden_size=cfg.d, num_layers=cfg.n_layer, num_heads=cfg.n_head,
                           rotary_dim=cfg.rotary_dim)
    cpu = torch.device("cpu")

    class Tok(ref_shims.StubTokenizer):  # ids of the tiny vocabulary; len() drives resize_token_embeddings (magma.py:50)
        cls_token_id, eos_token_id, pad_token_id = cfg.image_token, cfg.eos_token, cfg.eos_token

        def __len__(self):
            return cfg.vocab

    # the import redirection of INTEGRATION.md section 1, applied to the names magma/magma.py imported
    monkeypatch.setattr(ref.magma, "get_gptj", lambda **kw: BL.get_gptj(config=lm_cfg, device=cpu))
    monkeypatch.setattr(ref.magma, "get_tokenizer", lambda *a, **k: Tok())
    monkeypatch.setattr(ref.magma, "ImagePrefix", lambda config, out_dim: BP.ImagePrefix(config, out_dim, device=cpu))
    for name in ("Adapter", "ParallelAdapter", "AdapterWrapper", "ParallelAdapterWrapper"):
        monkeypatch.setattr(ref.magma, name, getattr(BA, name))
    return ref, ref_config, cfg, oracle_weights(cfg), cpu


def _load(model, w16):
    missing, unexpected = model.load_state_dict(w16, strict=False)
    missing = [k for k in missing if not k.startswith(("word_embedding.", "transformer."))]
    assert not missing and not unexpected, (missing, unexpected)
    model.lm.invalidate()  # the one call INTEGRATION.md asks for after add_adapters / weight loading
    model.image_prefix.enc.invalidate()


@pytest.mark.parametrize("adapters", [{"mlp": {"adapter_type": "normal", "downsample_factor": 4}},
                                      {"mlp": {"adapter_type": "normal", "downsample_factor": 8},
                                       "attention": {"adapter_type": "normal", "downsample_factor": 8}}])
def test_reference_magma_class_runs_over_b200_modules(reference_over_b200, adapters):
    ref, ref_config, cfg, w16, cpu = reference_over_b200
    import dataclasses

    cfg = dataclasses.replace(cfg, mlp_adapter=adapters.get("mlp"), attn_adapter=adapters.get("attention"))
    w16 = __import__("test_e2e_dryrun_cpu").oracle_weights(cfg)
    S, B = 16, 3
    rc = ref_config.MultimodalConfig(batch_size=B, train_steps=1, encoder_name="clip_vit_dropin", adapter_config=adapters,
                                     image_seq_len=cfg.image_seq_len, use_image_embed_layernorm=True,
                                     image_embed_dropout_prob=0.0, image_size=cfg.vit_image, freeze_img_encoder=True)
    model = ref.magma.Magma(rc, device=cpu)            # the REFERENCE's class
    assert type(model).__module__ == "magma.magma"
    assert model.lm.lm_head.weight.shape[0] == cfg.vocab and model.word_embedding.weight.shape[0] == cfg.vocab
    assert model.mlp_adapter_added and model.attn_adapter_added == ("attention" in adapters)
    model.seq_len = S                                   # plain attribute
    _load(model, w16)
    model.eval()
    trainable = [n for n, p in model.named_parameters()
                 if p.requires_grad and ("adapter" in n or n.startswith(("image_prefix.proj", "image_prefix.ln")))]
    assert any(".adapter." in n for n in trainable)
    images, captions = O.synthetic_batch(cfg, B, S, seed=11)
    images = images.to(torch.bfloat16)
    out = model(images, captions)                       # magma/magma.py:238-276 incl. the reference's build_labels
    params = {k: v.clone().requires_grad_(k in trainable) for k, v in w16.items()}
    loss_o, logits_o, _ = O.magma_forward(images.float(), captions, params, cfg)
    assert abs(float(out.loss.detach()) - float(loss_o.detach())) < 2e-2
    assert rel(out.logits, logits_o.detach()) < 3e-2
    loss_o.backward()
    out.loss.backward()
    sd = dict(model.named_parameters())
    bad = {k: round(rel(sd[k].grad, params[k].grad), 4) for k in trainable if rel(sd[k].grad, params[k].grad) > 5e-2}
    assert not bad, bad
    # the reference's own decode loop (magma/sampling.py:43-121) over our LM == this package's generate on the same model
    with torch.no_grad():
        emb = model.image_prefix(images)
        emb = torch.cat([emb, model.word_embedding(captions[:, :3])], dim=1)
        toks_ref_loop = ref.sampling.generate(model, emb, max_steps=6, temperature=0.0, decode=False)
    from magma_b200.sampling import generate as b200_generate

    toks_b200 = b200_generate(model, emb, max_steps=6, temperature=0.0, decode=False)
    n = min(toks_ref_loop.shape[1], toks_b200.shape[1])
    assert torch.equal(toks_ref_loop[:, :n].cpu(), toks_b200[:, :n].cpu())


def test_reference_class_keeps_its_assertions_and_double_add_guard(reference_over_b200):
    ref, ref_config, cfg, w16, cpu = reference_over_b200
    rc = ref_config.MultimodalConfig(batch_size=2, train_steps=1, encoder_name="clip_vit_dropin",
                                     adapter_config={"mlp": {"adapter_type": "normal", "downsample_factor": 4}},
                                     image_seq_len=cfg.image_seq_len, image_size=cfg.vit_image)
    model = ref.magma.Magma(rc, device=cpu)
    with pytest.raises(ValueError, match="already added"):
        model.add_adapters(location="mlp")
    model.seq_len = 16
    images, captions = O.synthetic_batch(cfg, 2, 16, seed=1)
    with pytest.raises(AssertionError):
        model(images.to(torch.bfloat16), captions[:, :10])  # magma.py:249-251: captions must be padded to seq_len
