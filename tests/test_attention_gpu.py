"""GPU parity tests of the fused single-tile attention kernels (C ABI mb200_attn_fwd_tile / mb200_attn_bwd_tile)
against a torch-fp32 restatement of GPTJAttention._attn (hf:gptj/modeling_gptj.py:136-149) and its autograd.
Tolerance: relative Frobenius < 2e-2 (bf16 outputs; probabilities are rounded to bf16 before P*V on both sides)."""
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
