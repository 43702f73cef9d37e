"""Helpers for the GPU parity tests: build the CUDA-backed Magma from oracle/reference-named weights."""
import os

import torch


def gpu_device():
    """cuda:0 — or the CPU when tests/test_default_path_replay_cpu.py replays a GPU test body on emulated kernels."""
    return torch.device(os.environ.get("MB200_TEST_DEVICE", "cuda:0"))


def build_magma_from_weights(w, cfg, adapter_config, S, dev, vit_name="clip_vit_test"):
    from magma_b200.config import MultimodalConfig
    from magma_b200.image_encoders import register_vit
    from magma_b200.language_model import GPTJConfig
    from magma_b200.magma import Magma

    register_vit(vit_name, cfg.vit_width, cfg.vit_layers, cfg.vit_heads, cfg.vit_patch, cfg.vit_image, cfg.vit_mlp,
                 cfg.enc_out_dim)
    mc = MultimodalConfig(batch_size=2, train_steps=1, encoder_name=vit_name, adapter_config=adapter_config,
                          image_seq_len=cfg.image_seq_len, image_embed_dropout_prob=0.0,
                          use_image_embed_layernorm=True, image_size=cfg.vit_image, seq_len=S)
    mc._lm_config = GPTJConfig(vocab_size=cfg.vocab, hidden_size=cfg.d, num_layers=cfg.n_layer, num_heads=cfg.n_head,
                               rotary_dim=cfg.rotary_dim)
    model = Magma(mc, device=dev, init_seed=None)
    model.eos_token, model.image_token = cfg.eos_token, cfg.image_token
    missing, unexpected = model.load_state_dict(w, strict=False)
    # Magma registers lm.transformer.wte / .h a second time as word_embedding / transformer (magma/magma.py:52-53):
    # those alias keys share storage with the lm.* keys that were loaded
    missing = [k for k in missing if not k.startswith(("word_embedding.", "transformer."))]
    assert not missing and not unexpected, (missing, unexpected)
    model.lm.invalidate()
    model.lm.attach_arena(model.arena)
    model.image_prefix.enc.invalidate()
    return model


def rel(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return ((got - want).norm() / (want.norm() + 1e-12)).item()
