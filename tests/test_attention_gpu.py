"""GPU parity tests of the fused attention kernels — single-tile forward / backward (C ABI mb200_attn_fwd_tile /
mb200_attn_bwd_tile) and the multi-tile forward for any sequence length (mb200_attn_fwd_flash) — against a torch-fp32
restatement of GPTJAttention._attn (hf:gptj/modeling_gptj.py:136-149; non-causal: hf:clip/modeling_clip.py:282-330) and
its autograd. Tolerance: relative Frobenius < 2e-2 (bf16 outputs; probabilities are rounded to bf16 before P*V on both
sides)."""
import math

import pytest

pytestmark = pytest.mark.gpu


def _ref(qkv, B, S, H, hd):
    import torch

    q, k, v = qkv.float().view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    mask = torch.ones(S, S, device=qkv.device).tril().bool()
    att = att.masked_fill(~mask, torch.finfo(torch.float32).min)
    p = torch.softmax(att, -1)
    o = (p.to(torch.bfloat16).float() @ v).permute(0, 2, 1, 3).reshape(B * S, H * hd)
    return q, k, v, p, o


@pytest.mark.parametrize("S,H,hd", [(128, 4, 256), (70, 3, 256), (128, 2, 128), (33, 2, 64), (1, 1, 64)])
def test_attn_tile_forward_backward(S, H, hd):
    import torch

    from magma_b200 import ops

    dev = torch.device("cuda:0")
    B = 2
    g = torch.Generator(device=dev).manual_seed(S * 1000 + hd)
    qkv = (0.5 * torch.randn(B * S, 3 * H * hd, device=dev, generator=g)).to(torch.bfloat16)
    O, P = ops.attn_fwd_tile(qkv, B, S, H, hd)
    q, k, v, p_ref, o_ref = _ref(qkv, B, S, H, hd)
    rel = lambda a, b: ((a.float() - b).norm() / (b.norm() + 1e-12)).item()
    assert rel(O, o_ref) < 2e-2
    assert rel(P[..., :S], p_ref) < 2e-2
    assert (P[..., :S].float().triu(1) == 0).all()  # causal mask: exact zeros above the diagonal
    dO = torch.randn(B * S, H * hd, device=dev, generator=g).to(torch.bfloat16)
    o_ref.backward(dO.float())
    dqkv = ops.attn_bwd_tile(qkv, dO, P, B, S, H, hd).float().view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    for got, want, name in zip(dqkv, (q.grad, k.grad, v.grad), "qkv"):
        assert rel(got, want) < 3e-2, (name, rel(got, want))


def test_attn_tile_inverse_rope_matches_separate_kernel():
    import torch

    from magma_b200 import ops

    dev = torch.device("cuda:0")
    B, S, H, hd, rot = 2, 96, 2, 128, 64
    qkv = (0.5 * torch.randn(B * S, 3 * H * hd, device=dev)).to(torch.bfloat16)
    dO = torch.randn(B * S, H * hd, device=dev).to(torch.bfloat16)
    _, P = ops.attn_fwd_tile(qkv, B, S, H, hd)
    plain = ops.attn_bwd_tile(qkv, dO, P, B, S, H, hd)
    tab = ops.rope_table(S, rot, 0, device=dev)
    fused = ops.attn_bwd_tile(qkv, dO, P, B, S, H, hd, rope_tab=tab, rot=rot)
    want = ops.rope_(plain.clone(), S, H, hd, rot, pos0=0, inverse=True)
    assert ((fused.float() - want.float()).norm() / want.float().norm()).item() < 1e-2


def test_attn_tile_rejects_unsupported_shapes():
    import torch

    from magma_b200 import ops
    from magma_b200._lib import MB200Error

    dev = torch.device("cuda:0")
    with pytest.raises(MB200Error, match="unsupported"):
        ops.attn_fwd_tile(torch.zeros(2 * 129, 3 * 64, device=dev, dtype=torch.bfloat16), 2, 129, 1, 64)
    with pytest.raises(MB200Error, match="unsupported"):
        ops.attn_fwd_tile(torch.zeros(2 * 16, 3 * 48, device=dev, dtype=torch.bfloat16), 2, 16, 1, 48)


def _dev():
    """cuda:0 — or the CPU when tests/test_attention_twin_cpu.py replays a test body on emulated kernels."""
    import os

    import torch

    return torch.device(os.environ.get("MB200_TEST_DEVICE", "cuda:0"))


def _ref_general(q, k, v, causal):
    """q [B,H,Sq,hd], k / v [B,H,Sk,hd] (fp32): softmax(q k^T / sqrt(hd) [+ causal mask with offset Sk - Sq]) v with
    the probabilities rounded to bf16 before P V."""
    import torch

    Sq, Sk, hd = q.shape[2], k.shape[2], q.shape[3]
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:
        i = torch.arange(Sq, device=q.device)[:, None]
        j = torch.arange(Sk, device=q.device)[None, :]
        att = att.masked_fill(j > i + (Sk - Sq), torch.finfo(torch.float32).min)
    p = torch.softmax(att, -1)
    return p, p.to(torch.bfloat16).float() @ v


@pytest.mark.parametrize("S,H,hd,causal", [(129, 2, 256, True), (257, 4, 64, False), (300, 2, 256, True),
                                            (128, 2, 128, True), (1, 1, 64, False), (640, 2, 128, True),
                                            (2048, 2, 256, True), (70, 3, 192, False), (384, 2, 64, True),
                                            (256, 2, 128, False)])
def test_attn_flash_forward_matches_materialised_softmax(S, H, hd, causal):
    import torch

    from magma_b200 import ops

    dev = _dev()
    B = 2 if S < 2048 else 1
    g = torch.Generator().manual_seed(S * 1000 + hd)
    qkv = (0.5 * torch.randn(B * S, 3 * H * hd, generator=g)).to(torch.bfloat16).to(dev)
    O, P, stats = ops.attn_fwd_flash(qkv, B, S, H, hd, causal=causal, want_p=True, want_stats=True)
    q, k, v = qkv.float().view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    p_ref, o_ref = _ref_general(q, k, v, causal)
    o_ref = o_ref.permute(0, 2, 1, 3).reshape(B * S, H * hd)
    rel = lambda a, b: ((a.float() - b).norm() / (b.norm() + 1e-12)).item()
    assert rel(O, o_ref) < 2e-2
    assert rel(P[..., :S], p_ref) < 2e-2
    assert (P[..., S:] == 0).all()
    if causal:
        assert (P[..., :S].float().triu(1) == 0).all()  # exact zeros above the diagonal
    # row statistics: maximum of the scaled scores and 1 / sum exp(s - max)
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:
        att = att.masked_fill(~torch.ones(S, S, device=dev).tril().bool(), float("-inf"))
    m_ref = att.max(-1).values
    assert (stats[..., 0] - m_ref).abs().max().item() < 2e-3 * (1 + m_ref.abs().max().item())
    assert rel(stats[..., 1], 1.0 / torch.exp(att - m_ref[..., None]).sum(-1)) < 1e-3
    O2 = ops.attn_fwd_flash(qkv, B, S, H, hd, causal=causal)  # without the optional outputs
    assert torch.equal(O2, O)


@pytest.mark.parametrize("hd,pos0,S,Smax", [(256, 100, 60, 512), (64, 129, 130, 300), (128, 0, 200, 256)])
def test_attn_flash_over_a_kv_cache_with_offset_causal_mask(hd, pos0, S, Smax):
    """Prefill continuation: S new queries attend to pos0 cached + S new keys; cache rows beyond pos0 + S hold NaN and
    must never be read (the tensor maps end at Sk, TMA zero-fills beyond)."""
    import torch

    from magma_b200 import ops

    dev = _dev()
    B, H = 2, 2
    g = torch.Generator().manual_seed(pos0 + S)
    Sk = pos0 + S
    qkv = (0.5 * torch.randn(B * S, 3 * H * hd, generator=g)).to(torch.bfloat16).to(dev)
    kc = torch.full((B, H, Smax, hd), float("nan"), dtype=torch.bfloat16, device=dev)
    vc = torch.full((B, H, Smax, hd), float("nan"), dtype=torch.bfloat16, device=dev)
    kc[:, :, :Sk] = (0.5 * torch.randn(B, H, Sk, hd, generator=g)).to(torch.bfloat16).to(dev)
    vc[:, :, :Sk] = (0.5 * torch.randn(B, H, Sk, hd, generator=g)).to(torch.bfloat16).to(dev)
    O = ops.attn_fwd_flash(qkv, B, S, H, hd, causal=True, kcache=kc, vcache=vc, pos0=pos0)
    q = qkv.float().view(B, S, 3, H, hd)[:, :, 0].permute(0, 2, 1, 3)
    _, o_ref = _ref_general(q, kc[:, :, :Sk].float(), vc[:, :, :Sk].float(), True)
    o_ref = o_ref.permute(0, 2, 1, 3).reshape(B * S, H * hd)
    assert torch.isfinite(O.float()).all()
    assert ((O.float() - o_ref).norm() / o_ref.norm()).item() < 2e-2
