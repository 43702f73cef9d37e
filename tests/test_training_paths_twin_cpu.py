"""CPU replay of test BODIES in tests/test_training_paths_gpu.py: the same functions, run on CPU tensors with the kernels
emulated (fixture `emul_ops`). This does not verify
any kernel; it keeps the test bodies themselves exercised in the CPU suite, so a failure on the GPU is about the kernels and not
about the test."""
import pytest

import test_training_paths_gpu as Z

REPLAYABLE = [
    "test_quick_gelu_bwd_matches_autograd",
    "test_layernorm_param_grad_rows_matches_the_small_kernel_and_fp32",
    "test_conv_trunk_training_kernels_match_torch",
    "test_conv_trunk_training_units_match_fp32_locally",
    "test_conv_trunk_training_matches_oracle_like_with_like",
    "test_trainable_vit_gradients_match_oracle_autograd",
    "test_encoder_learning_rate_group_and_weight_decay_exemptions",
    "test_frozen_vit_is_unchanged_by_the_training_path",
    "test_engine_checkpoint_resume_continues_the_same_trajectory",
    "test_optimizer_on_its_own_stream_gives_the_same_trajectory",
    "test_fused_attention_paths_agree_with_the_batched_gemm_path",
    "test_preprocess_inputs_image_and_text_to_embeddings",
]

REPLAYABLE_WITH_ARGS = {
    "test_adapter_forms_with_layernorm_and_scale_match_oracle": [("normal", "normal", True, True),
                                                                 ("scaled_parallel", "scaled_parallel", False, False),
                                                                 ("scaled_parallel", "normal", True, True)],
}


@pytest.mark.parametrize("name", REPLAYABLE)
def test_replay_on_emulated_kernels(emul_ops, monkeypatch, name):
    from magma_b200.magma import Magma

    monkeypatch.setenv("MB200_TEST_DEVICE", "cpu")
    monkeypatch.setattr(Magma, "_require_cuda", lambda self: None)
    fn = getattr(Z, name)
    import inspect

    if "monkeypatch" in inspect.signature(fn).parameters:
        fn(monkeypatch)
    elif "tmp_path" in inspect.signature(fn).parameters:
        import pathlib
        import tempfile

        with tempfile.TemporaryDirectory() as d:
            fn(pathlib.Path(d))
    else:
        fn()


@pytest.mark.parametrize("args", REPLAYABLE_WITH_ARGS["test_adapter_forms_with_layernorm_and_scale_match_oracle"])
def test_replay_adapter_forms(emul_ops, monkeypatch, args):
    from magma_b200.magma import Magma

    monkeypatch.setenv("MB200_TEST_DEVICE", "cpu")
    monkeypatch.setattr(Magma, "_require_cuda", lambda self: None)
    Z.test_adapter_forms_with_layernorm_and_scale_match_oracle(*args)
