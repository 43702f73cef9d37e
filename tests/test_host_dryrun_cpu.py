"""Dry run of the HOST-SIDE Python schedules on the CPU: magma_b200/ops.py is pointed at the CPU emulation of the
primitive C-ABI operators (fixture `emul_ops`, oracle/cabi_emul.cpp), so the code paths a GPU run takes through
image_prefix.py, adapters.py, the conv trunk of image_encoders.py and arena.py execute here on CPU tensors and are
compared with the oracle / torch. This checks argument plumbing (shapes, strides, majors, accumulate flags, autograd
wiring) — never the CUDA kernels, which only the `-m gpu` tests can."""
import types

import pytest
import torch
import torch.nn as nn

from oracle import magma_oracle as O


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


class _StubEncoder(nn.Module):
    input_resolution = 32

    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def forward(self, x):
        return self.feats


@pytest.mark.parametrize("seq_encoder", [False, True])
def test_image_prefix_forward_and_backward(emul_ops, monkeypatch, seq_encoder):
    """ImagePrefix.forward (magma/image_prefix.py:78-109) + the hand-written backward of _PrefixFn, including the
    gradient w.r.t. the encoder features that a trainable encoder consumes (freeze_img_encoder: false)."""
    from magma_b200 import image_prefix as IP
    from magma_b200.config import MultimodalConfig

    d, enc_dim, B = 64, 48, 3
    name = "clip_resnet_dry" if seq_encoder else "clip_vit_dry"
    IP.ENCODER_OUT_DIMS[name] = enc_dim
    if seq_encoder:
        IP.ENCODER_SEQ_LENS[name] = 4
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(*((B, 4, enc_dim) if seq_encoder else (B, enc_dim)), generator=g).to(torch.bfloat16)
    feats_leaf = feats.float().requires_grad_(True)
    monkeypatch.setattr(IP, "get_image_encoder", lambda *a, **k: _StubEncoder(feats_leaf.to(torch.bfloat16)))
    cfg = MultimodalConfig(batch_size=B, train_steps=1, encoder_name=name, image_seq_len=2,
                           image_embed_dropout_prob=0.0, use_image_embed_layernorm=True)
    try:
        mod = IP.ImagePrefix(cfg, out_dim=d, device=torch.device("cpu"))
        with torch.no_grad():
            mod.proj.weight.copy_((torch.randn(mod.proj.weight.shape, generator=g) * 0.1).to(torch.bfloat16).float())
            mod.proj.bias.copy_((torch.randn(mod.proj.bias.shape, generator=g) * 0.1).to(torch.bfloat16).float())
            mod.ln.weight.copy_((1 + 0.1 * torch.randn(d, generator=g)).to(torch.bfloat16).float())
            mod.ln.bias.copy_((0.1 * torch.randn(d, generator=g)).to(torch.bfloat16).float())
        mod.eval()
        out = mod(torch.zeros(B, 3, 32, 32))
        gout = torch.randn(out.shape, generator=g).to(torch.bfloat16)
        out.backward(gout)
        # oracle
        ocfg = O.OracleConfig(d=d, image_seq_len=2, enc_out_dim=enc_dim)
        w = {"image_prefix.proj.weight": mod.proj.weight.detach().clone().requires_grad_(True),
             "image_prefix.proj.bias": mod.proj.bias.detach().clone().requires_grad_(True),
             "image_prefix.ln.weight": mod.ln.weight.detach().clone().requires_grad_(True),
             "image_prefix.ln.bias": mod.ln.bias.detach().clone().requires_grad_(True)}
        f_o = feats.float().requires_grad_(True)
        want = O.image_prefix_from_features(f_o, w, ocfg, fixed_seq=seq_encoder)
        want.backward(gout.float())
        assert out.shape == want.shape and rel(out, want.detach()) < 1e-2
        assert rel(mod.proj.weight.grad, w["image_prefix.proj.weight"].grad) < 2e-2
        assert rel(mod.proj.bias.grad, w["image_prefix.proj.bias"].grad) < 2e-2
        assert rel(mod.ln.weight.grad, w["image_prefix.ln.weight"].grad) < 2e-2
        assert rel(mod.ln.bias.grad, w["image_prefix.ln.bias"].grad) < 2e-2
        assert rel(feats_leaf.grad, f_o.grad) < 2e-2          # d feats: what _VitTrainFn.backward receives
    finally:
        IP.ENCODER_OUT_DIMS.pop(name, None)
        IP.ENCODER_SEQ_LENS.pop(name, None)


def test_image_prefix_dropout_uses_one_mask_for_forward_and_backward(emul_ops, monkeypatch):
    from magma_b200 import ops

    g = torch.Generator().manual_seed(1)
    x = torch.randn(16, 64, generator=g).to(torch.bfloat16)
    y, mask = ops.dropout_fwd(x, 0.25, seed=7)
    keep = mask.bool().view_as(y)
    assert 0.6 < keep.float().mean() < 0.9
    assert torch.equal(y[~keep], torch.zeros_like(y[~keep]))
    assert rel(y[keep], x[keep].float() / 0.75) < 5e-3
    gy = torch.randn(16, 64, generator=g).to(torch.bfloat16)
    gx = ops.dropout_apply(gy, mask, 0.25)
    assert torch.equal(gx[~keep], torch.zeros_like(gx[~keep])) and rel(gx[keep], gy[keep].float() / 0.75) < 5e-3
    y2, mask2 = ops.dropout_fwd(x, 0.25, seed=7)
    assert torch.equal(mask, mask2) and torch.equal(y, y2)      # counter-based: reproducible per seed


def test_standalone_adapter_forward_and_backward(emul_ops):
    """Adapter.forward = adapter(x) + x (magma/adapters.py:38-39) through _AdapterFn (two GEMMs forward, four backward)."""
    from magma_b200.adapters import Adapter

    g = torch.Generator().manual_seed(2)
    ad = Adapter(dim=64, downsample_factor=4)
    with torch.no_grad():
        for p in ad.parameters():
            p.copy_((torch.randn(p.shape, generator=g) * 0.2).to(torch.bfloat16).float())
        ad.down.bias.copy_(torch.where(torch.rand(16, generator=g) < 0.5, -3.0, 3.0))  # decided ReLU masks
    x = torch.randn(2, 5, 64, generator=g).to(torch.bfloat16).float().requires_grad_(True)
    y = ad(x)
    gy = torch.randn(y.shape, generator=g).to(torch.bfloat16).float()
    y.backward(gy)
    w = {f"a.adapter.{i}.{n}": getattr(ad.adapter[i], n).detach().clone().requires_grad_(True)
         for i in (0, 2) for n in ("weight", "bias")}
    xo = x.detach().clone().requires_grad_(True)
    want = O.adapter_forward(xo, w, "a")
    want.backward(gy)
    assert rel(y, want.detach()) < 1e-2 and rel(x.grad, xo.grad) < 2e-2
    for i in (0, 2):
        for n in ("weight", "bias"):
            assert rel(getattr(ad.adapter[i], n).grad, w[f"a.adapter.{i}.{n}"].grad) < 2e-2, (i, n)


def test_conv_trunk_forward_schedule(emul_ops):
    """B200ModifiedResNet._forward_eager (NHWC im2col + GEMM with folded BatchNorm, anti-aliasing pools, residual +
    ReLU_POST epilogue) against the oracle's NCHW restatement on a tiny trunk."""
    from magma_b200.image_encoders import B200ModifiedResNet

    cfg = O.OracleConfig(rn_width=16, rn_layers=(1, 2, 1, 1), rn_image=64)
    w = O.init_resnet_weights(cfg, seed=4, pre="enc")
    enc = B200ModifiedResNet(cfg.rn_layers, cfg.rn_width, cfg.rn_image, device=torch.device("cpu"))
    missing, unexpected = enc.load_state_dict({k[4:]: v for k, v in w.items()}, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(2, 3, 64, 64, generator=g).to(torch.bfloat16)
    got = enc._forward_eager(images)
    want = O.resnet_forward(images.float(), w, cfg, pre="enc")
    assert got.shape == want.shape == (2, 4, 16 * 32)
    assert rel(got, want) < 2e-2


def test_arena_optimizer_matches_torch_adamw_with_param_groups(emul_ops):
    """ParamArena.adamw_step over dp.optimizer_segments == torch.optim.AdamW with the reference's parameter groups
    (magma/utils.py:164-215: image encoder at its own rate, biases / LayerNorm exempt from weight decay) and global-norm
    clipping (config.py:126), three steps; the bf16 compute copy follows the fp32 master."""
    from magma_b200 import dp
    from magma_b200.arena import ParamArena

    g = torch.Generator().manual_seed(3)
    names = ["lm.transformer.h.0.mlp.1.adapter.0.weight", "lm.transformer.h.0.mlp.1.adapter.0.bias",
             "image_prefix.enc.proj", "image_prefix.enc.ln_post.weight", "image_prefix.proj.weight"]
    shapes = [(16, 8), (16,), (8, 12), (8,), (12, 8)]
    no_decay = [False, True, False, True, False]
    mine = [nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    ref = [nn.Parameter(p.detach().clone()) for p in mine]
    arena = ParamArena(list(zip(names, mine)), torch.device("cpu"))
    lr, enc_lr, wd, clip = 1e-2, 1e-4, 0.1, 0.5
    segs = dp.optimizer_segments(names, [p.numel() for p in mine], arena.offsets, no_decay, lr, enc_lr, wd)
    assert len(segs) == 5
    groups = [{"params": [r], "lr": enc_lr if n.startswith("image_prefix.enc.") else lr,
               "weight_decay": 0.0 if nd else wd} for r, n, nd in zip(ref, names, no_decay)]
    opt = torch.optim.AdamW(groups, betas=(0.9, 0.95), eps=1e-8)
    for _ in range(3):
        grads = [torch.randn(*s, generator=g) for s in shapes]
        for p, r, gr in zip(mine, ref, grads):
            arena.grad_of(p).copy_(gr)
            r.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_(ref, clip)
        opt.step()
        arena.adamw_step(lr=lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd, max_norm=clip, segments=segs)
    for p, r, n in zip(mine, ref, names):
        assert torch.allclose(p.detach(), r.detach(), rtol=2e-5, atol=2e-6), n
        assert torch.equal(arena.shadow_of(p), p.detach().to(torch.bfloat16)), n
    assert float(arena.grad.abs().sum()) == 0.0    # zero_grad folded into the kernel


def test_conv_trunk_training_forward_and_backward(emul_ops, monkeypatch):
    """freeze_img_encoder: false with a CLIP conv trunk (MAGMA_v1.yml / MAGMA_v2.yml): BatchNorm in training mode and the
    hand-scheduled backward of B200ModifiedResNet against torch autograd of the oracle — features, every conv / BN
    parameter gradient, and the running statistics. The oracle is run like-with-like: straight-through bf16 rounding
    wherever the CUDA path stores a tensor (`store=`), and the SAME ReLU pattern as the recorded activations (`relu=`).
    In pure fp32 the same graph differs from the bf16 run by ~5 % in the features (18 BatchNorm units with batch
    statistics) and ~40 % in the gradients (0.3 % of the pre-activations sit inside bf16 rounding of zero and flip their
    mask entry) — neither says anything about the schedule."""
    from magma_b200.image_encoders import B200ModifiedResNet

    cfg = O.OracleConfig(rn_width=16, rn_layers=(1, 2, 1, 1), rn_image=64)
    w = O.init_resnet_weights(cfg, seed=6, pre="enc")
    # the kernels read the bf16 compute copy of the convolution weights: give both sides bf16-representable values
    w = {k: (v.to(torch.bfloat16).float() if v.ndim == 4 else v) for k, v in w.items()}
    enc = B200ModifiedResNet(cfg.rn_layers, cfg.rn_width, cfg.rn_image, device=torch.device("cpu"))
    enc.load_state_dict({k[4:]: v for k, v in w.items()}, strict=False)
    for p in enc.parameters():
        p.requires_grad = True
    enc.train()
    g = torch.Generator().manual_seed(1)
    B = 4
    images = torch.randn(B, 3, 64, 64, generator=g).to(torch.bfloat16)
    feats, tape = enc._train_forward(images)
    dfeats = torch.randn(feats.shape, generator=g).to(torch.bfloat16)
    enc._train_backward(tape, dfeats)
    # the ReLU pattern of the bf16 run, in network order, as NCHW masks
    units = list(tape["stem"]) + [u for blk in tape["blocks"] for u in blk["units"]]
    masks = [(u["y"] > 0).view(B, -1, u["y"].shape[1]) for u in units if u["y"] is not None]
    it = iter(masks)

    def masked_relu(x):
        m = next(it)                                   # [B, H*W, C] -> [B, C, H, W]
        return x * m.permute(0, 2, 1).reshape(x.shape).to(x.dtype)

    wo = {k: (v.clone().requires_grad_(True) if not k.endswith(("running_mean", "running_var")) else v.clone())
          for k, v in w.items()}
    def store(v):                                      # bf16 storage, identity gradient
        return v + (v.detach().to(torch.bfloat16).float() - v.detach())

    want = O.resnet_forward(images.float(), wo, cfg, pre="enc", train_bn=True, relu=masked_relu, store=store)
    assert next(it, None) is None                      # every recorded mask was consumed: same number of ReLUs
    want.backward(dfeats.float())
    assert feats.shape == want.shape and rel(feats, want.detach()) < 1.5e-2
    sd = dict(enc.named_parameters())
    errs = {k[4:]: rel(sd[k[4:]].grad, v.grad) for k, v in wo.items() if v.requires_grad}
    assert len(errs) == len(sd)
    bad = {k: round(e, 3) for k, e in errs.items() if e > 3e-2}
    assert not bad, bad
    assert sorted(errs.values())[len(errs) // 2] < 1.5e-2   # median
    bufs = dict(enc.named_buffers())
    for k, v in wo.items():
        if k.endswith(("running_mean", "running_var")):
            assert torch.allclose(bufs[k[4:]], v, rtol=2e-2, atol=2e-3), k
    assert int(bufs["bn1.num_batches_tracked"]) == 1
    # through the autograd function (what ImagePrefix calls): same features, gradients land on the parameters
    for p in enc.parameters():
        p.grad = None
    f2 = enc(images)
    assert f2.requires_grad and rel(f2, want.detach()) < 3e-2
    f2.backward(dfeats)
    assert all(p.grad is not None for p in enc.parameters())
    # eval mode afterwards: running statistics, folded weights rebuilt from the current parameters
    monkeypatch.setenv("MB200_RESNET_GRAPH", "0")      # CUDA graphs are a GPU matter; same launches either way
    enc.eval()
    with torch.no_grad():
        e1 = enc(images)
    w_eval = {k: v.detach() for k, v in wo.items()}
    for k, v in enc.named_buffers():                  # the module has now seen two training batches
        if not k.endswith("num_batches_tracked"):
            w_eval["enc." + k] = v.clone()
    assert rel(e1, O.resnet_forward(images.float(), w_eval, cfg, pre="enc")) < 2e-2


def test_conv_training_primitives_are_what_the_schedule_assumes(emul_ops):
    """The emulated col2im3x3 / avgpool_nhwc_bwd are the adjoints of the forward layout operators (checked against
    torch autograd of unfold / avg_pool2d), and col_moments / channel_affine compute the documented expressions —
    the same reference formulas the GPU test of the real kernels uses (tests/test_training_paths_gpu.py)."""
    import torch.nn.functional as F

    from magma_b200 import ops

    g = torch.Generator().manual_seed(0)
    for B, H, W, C, s in ((2, 6, 6, 8, 1), (2, 6, 6, 8, 2), (1, 7, 5, 16, 2)):
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        xb = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16)
        cols, ho, wo = ops.im2col3x3(xb, s)
        x = xb.float().requires_grad_(True)
        xu = F.unfold(x.permute(0, 3, 1, 2), 3, padding=1, stride=s)                       # [B, C*9, Ho*Wo], (c, kh, kw)
        xu = xu.view(B, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, 9 * C)      # (kh*3+kw, c)
        assert (ho, wo) == (Ho, Wo) and torch.equal(cols.float(), xu.detach())
        dcols = torch.randn(cols.shape, generator=g).to(torch.bfloat16)
        xu.backward(dcols.float())
        assert rel(ops.col2im3x3(dcols, B, H, W, C, s), x.grad) < 5e-3
    dy = torch.randn(2, 3, 4, 16, generator=g).to(torch.bfloat16)
    x = torch.zeros(2, 6, 8, 16).requires_grad_(True)
    F.avg_pool2d(x.permute(0, 3, 1, 2), 2).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(ops.avgpool_nhwc_bwd(dy, 6, 8, 2), x.grad) < 5e-3
    u, v, m = (torch.randn(50, 16, generator=g).to(torch.bfloat16) for _ in range(3))
    o1, o2 = ops.col_moments(u, v, m)
    um = u.float() * (m.float() > 0)
    assert rel(o1, um.sum(0)) < 1e-5 and rel(o2, (um * v.float()).sum(0)) < 1e-5
    a1, a2, c0 = (torch.randn(16, generator=g) for _ in range(3))
    y = ops.channel_affine(u, a1, x2=v, a2=a2, c0=c0, mask=m, res=v, relu=True)
    assert rel(y, F.relu(um * a1 + v.float() * a2 + c0 + v.float())) < 5e-3


def test_frozen_conv_trunk_can_follow_the_reference_train_mode_batchnorm(emul_ops, monkeypatch):
    """Reference-literal option: a FROZEN trunk under train() still normalises with batch statistics and updates its
    running statistics (magma/magma.py:98-100 only clears requires_grad). Off by default (GPU-verified folded path)."""
    from magma_b200.image_encoders import B200ModifiedResNet

    monkeypatch.setenv("MB200_RESNET_GRAPH", "0")
    cfg = O.OracleConfig(rn_width=16, rn_layers=(1, 1, 1, 1), rn_image=64)
    w = O.init_resnet_weights(cfg, seed=8, pre="enc")
    w = {k: (v.to(torch.bfloat16).float() if v.ndim == 4 else v) for k, v in w.items()}
    enc = B200ModifiedResNet(cfg.rn_layers, cfg.rn_width, cfg.rn_image, device=torch.device("cpu"))
    enc.load_state_dict({k[4:]: v for k, v in w.items()}, strict=False)
    enc.train()
    g = torch.Generator().manual_seed(0)
    images = torch.randn(4, 3, 64, 64, generator=g).to(torch.bfloat16)
    folded = enc(images)                                           # default: eval statistics even under train()
    assert rel(folded, O.resnet_forward(images.float(), {k: v.clone() for k, v in w.items()}, cfg, pre="enc")) < 2e-2
    rm0 = enc.bn1.running_mean.clone()
    enc.bn_batch_stats_when_frozen = True
    got = enc(images)
    assert not got.requires_grad and not torch.equal(rm0, enc.bn1.running_mean)

    def store(v):
        return v.to(torch.bfloat16).float()

    want = O.resnet_forward(images.float(), {k: v.clone() for k, v in w.items()}, cfg, pre="enc", train_bn=True, store=store)
    assert rel(got, want) < 1.5e-2 and rel(got, folded) > 5e-2     # batch statistics, not the running ones
