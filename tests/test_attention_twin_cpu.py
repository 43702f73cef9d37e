"""CPU replay of the multi-tile attention test bodies of tests/test_attention_gpu.py on the emulated operator
(oracle/cabi_emul.cpp::mb200_attn_fwd_flash) at small sizes: keeps the test bodies and the emulation that the schedule
dry runs rely on honest; the kernel itself is verified by the `-m gpu` run only."""
import pytest

import test_attention_gpu as A


@pytest.mark.parametrize("S,H,hd,causal", [(129, 1, 64, True), (70, 2, 64, False), (33, 1, 128, True)])
def test_flash_forward_body_on_emulation(emul_ops, monkeypatch, S, H, hd, causal):
    monkeypatch.setenv("MB200_TEST_DEVICE", "cpu")
    A.test_attn_flash_forward_matches_materialised_softmax(S, H, hd, causal)


def test_flash_cache_body_on_emulation(emul_ops, monkeypatch):
    monkeypatch.setenv("MB200_TEST_DEVICE", "cpu")
    A.test_attn_flash_over_a_kv_cache_with_offset_causal_mask(64, 20, 30, 64)
