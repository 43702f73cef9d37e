"""Dry run of the product's host-only schedule files on the CPU.

magma_b200/csrc/vit_train.cu (ViT training forward + backward) contains no kernels: it carves a workspace and issues
primitive C-ABI operators. Here the SAME source file is compiled as plain C++ against oracle/cabi_emul.cpp (a scalar
CPU emulation of those primitives with bf16 storage, test infrastructure only) and run on CPU tensors, so every
pointer offset, leading dimension, operand major, batch stride and accumulate flag of the schedule is held to torch
autograd of the oracle (oracle/magma_oracle.py::vit_forward). What this cannot check is the CUDA kernels themselves —
that is the job of the `-m gpu` tests."""
import ctypes

import pytest
import torch

from magma_b200._lib import VitGradsC, VitLayerC, VitLayerGradsC, VitModelC
from oracle import magma_oracle as O

LAYER_KEYS = [("ln1_g", "ln_1.weight"), ("ln1_b", "ln_1.bias"), ("w_qkv", "attn.in_proj_weight"),
              ("b_qkv", "attn.in_proj_bias"), ("w_out", "attn.out_proj.weight"), ("b_out", "attn.out_proj.bias"),
              ("ln2_g", "ln_2.weight"), ("ln2_b", "ln_2.bias"), ("w_fc", "mlp.c_fc.weight"), ("b_fc", "mlp.c_fc.bias"),
              ("w_proj", "mlp.c_proj.weight"), ("b_proj", "mlp.c_proj.bias")]
TOP_KEYS = [("cls", "class_embedding"), ("pos", "positional_embedding"), ("ln_pre_g", "ln_pre.weight"),
            ("ln_pre_b", "ln_pre.bias"), ("ln_post_g", "ln_post.weight"), ("ln_post_b", "ln_post.bias")]


@pytest.fixture(scope="module")
def emul():
    from oracle import build_emul

    L = ctypes.CDLL(build_emul.build())
    L.mb200_last_error.restype = ctypes.c_char_p
    L.mb200_vit_train_workspace_bytes.restype = ctypes.c_size_t
    return L


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def make_case(seed=0, B=3):
    cfg = O.OracleConfig(d=64, n_layer=1, n_head=2, rotary_dim=8, vocab=64, enc_out_dim=48, vit_width=64, vit_layers=2,
                         vit_heads=4, vit_patch=8, vit_image=32, vit_mlp=128)
    pre = "image_prefix.enc"
    w = {k: v for k, v in O.init_weights(cfg, seed=seed).items() if k.startswith(pre)}
    g = torch.Generator().manual_seed(seed + 1)
    for k in w:  # larger weights than the 0.02 init so every gradient is well above bf16 noise
        if k.endswith(("in_proj_weight", "out_proj.weight", "c_fc.weight", "c_proj.weight", "conv1.weight", ".proj")):
            w[k] = w[k] * 4
    w16 = {k: v.to(torch.bfloat16) for k, v in w.items()}
    images = torch.randn(B, 3, cfg.vit_image, cfg.vit_image, generator=g).to(torch.bfloat16)
    dfeats = torch.randn(B, cfg.enc_out_dim, generator=g).to(torch.bfloat16)
    return cfg, pre, w16, images, dfeats


def c_model(cfg, pre, w16, keep):
    K = 3 * cfg.vit_patch ** 2
    ldk = (K + 7) // 8 * 8
    conv = torch.zeros(cfg.vit_width, ldk, dtype=torch.bfloat16)
    conv[:, :K] = w16[f"{pre}.conv1.weight"].reshape(cfg.vit_width, K)
    proj_t = w16[f"{pre}.proj"].t().contiguous()
    layers = (VitLayerC * cfg.vit_layers)()
    for i in range(cfg.vit_layers):
        for f, k in LAYER_KEYS:
            setattr(layers[i], f, w16[f"{pre}.transformer.resblocks.{i}.{k}"].data_ptr())
    m = VitModelC()
    m.n_layer, m.width, m.n_head, m.patch = cfg.vit_layers, cfg.vit_width, cfg.vit_heads, cfg.vit_patch
    m.image, m.mlp, m.out_dim = cfg.vit_image, cfg.vit_mlp, cfg.enc_out_dim
    m.w_conv, m.ld_conv = conv.data_ptr(), ldk
    for f, k in TOP_KEYS:
        setattr(m, f, w16[f"{pre}.{k}"].data_ptr())
    m.proj_t = proj_t.data_ptr()
    m.layers = ctypes.cast(layers, ctypes.POINTER(VitLayerC))
    keep += [conv, proj_t, layers]
    return m


def c_grads(cfg, pre, w16, keep, fill=0.0):
    """fp32 gradient buffers with each parameter's own shape; returns (struct, {param name: tensor})."""
    out = {}

    def buf(name):
        out[name] = torch.full(w16[name].shape, fill, dtype=torch.float32)
        return out[name].data_ptr()

    lg = (VitLayerGradsC * cfg.vit_layers)()
    for i in range(cfg.vit_layers):
        for f, k in LAYER_KEYS:
            setattr(lg[i], f, buf(f"{pre}.transformer.resblocks.{i}.{k}"))
    G = VitGradsC()
    G.w_conv = buf(f"{pre}.conv1.weight")
    for f, k in TOP_KEYS:
        setattr(G, f, buf(f"{pre}.{k}"))
    G.proj = buf(f"{pre}.proj")
    G.layers = ctypes.cast(lg, ctypes.POINTER(VitLayerGradsC))
    keep.append(lg)
    return G, out


def oracle_grads(cfg, pre, w16, images, dfeats):
    wf = {k: v.float().requires_grad_(True) for k, v in w16.items()}
    feats = O.vit_forward(images.float(), wf, cfg, pre=pre)
    feats.backward(dfeats.float())
    return feats.detach(), {k: v.grad for k, v in wf.items()}


def run_emul(L, cfg, pre, w16, images, dfeats, accumulate=0, fill=0.0):
    keep = []
    m = c_model(cfg, pre, w16, keep)
    B = images.shape[0]
    n = L.mb200_vit_train_workspace_bytes(ctypes.byref(m), B)
    assert n > 0
    ws = torch.empty(n + 256, dtype=torch.uint8)
    off = (-ws.data_ptr()) % 256  # the product hands the schedule a 256-byte aligned torch allocation
    wsp = ctypes.c_void_p(ws.data_ptr() + off)
    feats = torch.empty(B, cfg.enc_out_dim, dtype=torch.bfloat16)
    rc = L.mb200_vit_forward_train(ctypes.byref(m), ptr(images), ptr(feats), B, wsp, ctypes.c_size_t(n), None)
    assert rc == 0, L.mb200_last_error().decode()
    G, grads = c_grads(cfg, pre, w16, keep, fill=fill)
    rc = L.mb200_vit_backward(ctypes.byref(m), ctypes.byref(G), ptr(dfeats), accumulate, B, wsp, ctypes.c_size_t(n), None)
    assert rc == 0, L.mb200_last_error().decode()
    return feats, grads


def test_vit_train_schedule_matches_oracle_autograd(emul):
    cfg, pre, w16, images, dfeats = make_case()
    feats_o, grads_o = oracle_grads(cfg, pre, w16, images, dfeats)
    feats, grads = run_emul(emul, cfg, pre, w16, images, dfeats)
    assert rel(feats, feats_o) < 2e-2
    assert set(grads) == set(grads_o)
    bad = {k: round(rel(g, grads_o[k]), 4) for k, g in grads.items() if rel(g, grads_o[k]) > 4e-2}
    assert not bad, bad  # bf16 activations vs the fp32 oracle


def test_vit_train_schedule_accumulates(emul):
    """accumulate != 0 adds into every gradient buffer (gradient accumulation, train_loop.py:10-19)."""
    cfg, pre, w16, images, dfeats = make_case(seed=3, B=2)
    _, g0 = run_emul(emul, cfg, pre, w16, images, dfeats, accumulate=0, fill=7.0)   # overwrite: the 7s must vanish
    _, g1 = run_emul(emul, cfg, pre, w16, images, dfeats, accumulate=1, fill=0.5)
    for k in g0:
        assert torch.allclose(g1[k], g0[k] + 0.5, rtol=1e-4, atol=1e-4), k


def test_vit_train_schedule_rejects_a_small_workspace(emul):
    cfg, pre, w16, images, dfeats = make_case(B=1)
    keep = []
    m = c_model(cfg, pre, w16, keep)
    ws = torch.empty(1024, dtype=torch.uint8)
    feats = torch.empty(1, cfg.enc_out_dim, dtype=torch.bfloat16)
    rc = emul.mb200_vit_forward_train(ctypes.byref(m), ptr(images), ptr(feats), 1, ptr(ws), ctypes.c_size_t(1024), None)
    assert rc != 0 and b"workspace too small" in emul.mb200_last_error()
