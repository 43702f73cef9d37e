"""Training-state checkpoints (magma/utils.py:89-117 via the engine's save_checkpoint / load_checkpoint) and the small
loader helpers either side of train_step — host logic, exercised on CPU tensors with a stand-in arena."""
import types

import pytest
import torch

from magma_b200 import checkpoint as ck
from magma_b200 import dp
from magma_b200.utils import collate_fn, count_parameters, cycle


def fake_arena(seed=0, with_moments=True):
    names = ["lm.transformer.h.1.mlp.1.adapter.0.weight", "lm.transformer.h.1.mlp.1.adapter.0.bias",
             "image_prefix.proj.weight"]
    numels = [96, 8, 40]
    offs, total = dp.arena_layout(numels)
    g = torch.Generator().manual_seed(seed)
    a = types.SimpleNamespace(names=names, offsets=offs, numel=total, master=torch.randn(total, generator=g),
                              exp_avg=torch.randn(total, generator=g) if with_moments else None,
                              exp_avg_sq=torch.rand(total, generator=g) if with_moments else None, step_count=17,
                              synced=0)
    a.sync_shadow = lambda force=False: setattr(a, "synced", a.synced + 1)
    return a


def test_optimizer_state_round_trip_and_layout_check(tmp_path):
    a = fake_arena(0)
    module_state = {"lm.transformer.h.1.mlp.1.adapter.0.weight": torch.randn(8, 12)}
    d = ck.save_training_checkpoint(tmp_path, "global_step17", module_state, ck.arena_optimizer_state(a),
                                    {"global_step": 17, "config": {"lr": 1e-3}})
    assert (tmp_path / "latest").read_text() == "global_step17"
    path, payload, optim = ck.read_training_checkpoint(tmp_path)
    assert path == d and payload["global_step"] == 17 and payload["config"] == {"lr": 1e-3}
    assert torch.equal(payload["module"]["lm.transformer.h.1.mlp.1.adapter.0.weight"],
                       module_state["lm.transformer.h.1.mlp.1.adapter.0.weight"])
    b = fake_arena(1, with_moments=False)  # a freshly built model: no moments yet
    ck.load_arena_optimizer_state(b, optim)
    assert torch.equal(b.master, a.master) and torch.equal(b.exp_avg, a.exp_avg) and torch.equal(b.exp_avg_sq, a.exp_avg_sq)
    assert b.step_count == 17 and b.synced == 1  # the bf16 compute copy is refreshed from the restored master
    c = fake_arena(2)
    ck.load_arena_optimizer_state(c, optim, load_optimizer_states=False)   # weights only
    assert torch.equal(c.master, a.master) and not torch.equal(c.exp_avg, a.exp_avg)
    c.names = c.names[:-1] + ["image_prefix.ln.weight"]
    with pytest.raises(RuntimeError, match="does not match"):
        ck.load_arena_optimizer_state(c, optim)


def test_nothing_to_load_returns_none(tmp_path):
    assert ck.read_training_checkpoint(tmp_path) == (None, None, None)
    (tmp_path / "latest").write_text("global_step5")
    assert ck.read_training_checkpoint(tmp_path) == (None, None, None)


def test_reference_named_checkpoint_is_readable_by_from_checkpoint_adapter(tmp_path):
    sd = {"lm.transformer.h.0.attn.q_proj.weight": torch.zeros(2, 2), "lm.transformer.h.0.mlp.0.fc_in.weight": torch.ones(2, 2),
          "image_prefix.enc.transformer.resblocks.0.mlp.c_fc.weight": torch.ones(1)}
    ck.save_training_checkpoint(tmp_path, "t", sd, {"names": []}, {}, reference_names=True)
    _, payload, _ = ck.read_training_checkpoint(tmp_path, "t")
    assert "lm.transformer.h.0.attn.attention.q_proj.weight" in payload["module"]
    back, _ = ck.convert_reference_state_dict(payload["module"])
    assert set(back) == set(sd)


def test_loader_helpers():
    batch = [(torch.zeros(1, 3, 4, 4), torch.arange(10)[None]), (torch.ones(1, 3, 4, 4), torch.arange(10)[None] + 10)]
    images, caps = collate_fn(batch, seq_len=6)
    assert images.shape == (2, 3, 4, 4) and caps.shape == (2, 6) and caps[1, 0] == 10
    it = cycle([1, 2])
    assert [next(it) for _ in range(5)] == [1, 2, 1, 2, 1]
    m = torch.nn.Linear(3, 2)
    m.bias.requires_grad = False
    assert count_parameters(m) == 6
