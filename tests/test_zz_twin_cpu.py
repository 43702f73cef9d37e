"""CPU replay of the test BODIES in tests/test_zz_unverified_gpu.py that do not need engine.cu: the same functions, run
on CPU tensors with the kernels emulated (fixture `emul_ops`) and the LM on the general schedule. This does not verify
any kernel; it makes sure that when those GPU tests run for the first time, a failure is about the kernels and not about
a typo in the test."""
import pytest

import test_zz_unverified_gpu as Z

REPLAYABLE = [
    "test_quick_gelu_bwd_matches_autograd",
    "test_layernorm_param_grad_rows_matches_the_small_kernel_and_fp32",
    "test_conv_trunk_training_kernels_match_torch",
    "test_conv_trunk_training_matches_oracle_like_with_like",
    "test_trainable_vit_gradients_match_oracle_autograd",
    "test_encoder_learning_rate_group_and_weight_decay_exemptions",
]


@pytest.mark.parametrize("name", REPLAYABLE)
def test_replay_on_emulated_kernels(emul_ops, monkeypatch, name):
    from magma_b200.magma import Magma

    monkeypatch.setenv("MB200_TEST_DEVICE", "cpu")
    monkeypatch.setenv("MB200_FORCE_GENERAL", "1")      # the LM through csrc/gptj_sched.cu (engine.cu is GPU only)
    monkeypatch.setattr(Magma, "_require_cuda", lambda self: None)
    getattr(Z, name)()
