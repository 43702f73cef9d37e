"""Replay of the DEFAULT Python paths on the CPU. The GPU parity harness itself (tools/model_check.py groups lm,
lm_variants, vit, magma — what tests/test_model_gpu.py runs on a B200) and __graft_entry__.smoke() are executed here on
CPU tensors: primitive operators emulated (oracle/cabi_emul.cpp), and the model-level entry points of engine.cu
(mb200_gptj_forward / backward, mb200_vit_forward) provided by oracle/cabi_emul_models.cpp, which delegates to the
product's host-only schedules. This shows that the Python of the default path — Magma.forward, _EmbedLMFn, the pointer
tables of language_model.py / image_encoders.py, chunked backward, ParamArena, B200Engine — still works after a change,
without a GPU. It says nothing about engine.cu's own schedule or about any kernel: those are GPU-tested only."""
import pytest
import torch


@pytest.fixture
def replay(emul_ops, monkeypatch):
    from magma_b200.magma import Magma

    monkeypatch.setattr(Magma, "_require_cuda", lambda self: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    return torch.device("cpu")


@pytest.mark.parametrize("group", ["lm", "lm_variants", "vit", "magma"])
def test_gpu_parity_harness_replays_on_emulated_kernels(replay, group, capsys):
    from tools import model_check

    assert getattr(model_check, "group_" + group)(replay)
    assert "[FAIL]" not in capsys.readouterr().out


def test_smoke_replays_on_emulated_kernels(replay, capsys):
    import __graft_entry__ as entry

    entry.smoke(_device="cpu", _decode=False)     # KV-cache decoding lives in engine.cu (GPU only)
    assert "smoke (no decode) OK" in capsys.readouterr().out


@pytest.mark.parametrize("tag", ["v1_mlp_normal", "v2_mlp_attn_normal", "parallel", "no_adapters"])
def test_reference_goldens_replay_on_emulated_kernels(replay, monkeypatch, golden_dir, tag):
    """The fixtures the REFERENCE ITSELF produced (tests/golden/magma_*.pt: loss, logits, every trainable gradient of
    Magma.forward under four adapter wirings) against this package's Python + host-only schedules on emulated kernels —
    the body of tests/test_model_gpu.py::test_magma_matches_reference_golden, unchanged."""
    import test_model_gpu as G

    monkeypatch.setenv("MB200_TEST_DEVICE", "cpu")
    G.test_magma_matches_reference_golden(golden_dir, tag)


def test_reference_assertion_behaviour_replays(replay, monkeypatch):
    import test_model_gpu as G

    monkeypatch.setenv("MB200_TEST_DEVICE", "cpu")
    G.test_magma_forward_asserts_like_the_reference()
