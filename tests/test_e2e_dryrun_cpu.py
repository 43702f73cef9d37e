"""Whole-model dry run on the CPU: Magma.forward -> loss.backward() -> B200Engine.step() with the primitive operators
emulated (fixture `emul_ops`) and the LM running through its host-only schedule (csrc/gptj_sched.cu, compiled into the
emulation library). What runs here is the product's own Python — Magma / ImagePrefix / the trainable encoders /
ParamArena / B200Engine / checkpointing — and the two host-only C++ schedules; what is emulated are the kernels. The
result is held to torch autograd of the oracle (oracle/magma_oracle.py::magma_forward)."""
import pytest
import torch

from oracle import magma_oracle as O


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def tiny_cfg(**kw):
    base = dict(d=64, n_layer=2, n_head=4, rotary_dim=8, vocab=96, image_seq_len=2, enc_out_dim=48, vit_width=64,
                vit_layers=2, vit_heads=4, vit_patch=8, vit_image=32, vit_mlp=128, eos_token=90, image_token=91)
    base.update(kw)
    return O.OracleConfig(**base)


def build(monkeypatch, cfg, w16, S, encoder="clip_vit_dry", freeze_enc=False, adapter_config=None, **mc_kw):
    from magma_b200.config import MultimodalConfig
    from magma_b200.image_encoders import register_vit
    from magma_b200.language_model import GPTJConfig
    from magma_b200.magma import Magma

    monkeypatch.setattr(Magma, "_require_cuda", lambda self: None)   # test-only: kernels are emulated (emul_ops)
    register_vit("clip_vit_dry", cfg.vit_width, cfg.vit_layers, cfg.vit_heads, cfg.vit_patch, cfg.vit_image, cfg.vit_mlp,
                 cfg.enc_out_dim)
    mc = MultimodalConfig(batch_size=2, train_steps=1, encoder_name=encoder,
                          adapter_config=adapter_config or {"mlp": dict(cfg.mlp_adapter)},
                          image_seq_len=cfg.image_seq_len,
                          image_embed_dropout_prob=0.0, use_image_embed_layernorm=True, image_size=cfg.vit_image,
                          seq_len=S, freeze_img_encoder=freeze_enc, **mc_kw)
    mc._lm_config = GPTJConfig(vocab_size=cfg.vocab, hidden_size=cfg.d, num_layers=cfg.n_layer, num_heads=cfg.n_head,
                               rotary_dim=cfg.rotary_dim)
    model = Magma(mc, device=torch.device("cpu"), init_seed=None)
    model.eos_token, model.image_token = cfg.eos_token, cfg.image_token
    if w16 is not None:
        missing, unexpected = model.load_state_dict(w16, strict=False)
        missing = [k for k in missing if not k.startswith(("word_embedding.", "transformer."))]
        assert not missing and not unexpected, (missing, unexpected)
    model.lm.invalidate()
    model.lm.attach_arena(model.arena)
    model.image_prefix.enc.invalidate()
    return model, mc


def oracle_weights(cfg, seed=5):
    w = O.init_weights(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k in w:
        if ".adapter." in k:  # O(0.05) adapters with decided ReLU masks (bias +-3), like tools/model_check.py
            w[k] = torch.randn(w[k].shape, generator=g) * (0.05 if k.endswith("weight") else 0.02)
            if k.endswith("adapter.0.bias"):
                w[k] = torch.where(torch.rand(w[k].shape, generator=g) < 0.5, -3.0, 3.0)
        elif k.startswith("image_prefix.enc.") and k.endswith(("in_proj_weight", "out_proj.weight", "c_fc.weight",
                                                                "c_proj.weight", "conv1.weight", ".proj")):
            w[k] = w[k] * 4
        elif k.startswith("lm.") and k.endswith(("proj.weight", "fc_in.weight", "fc_out.weight")):
            w[k] = w[k] * 3
    return {k: v.to(torch.bfloat16).float() for k, v in w.items()}


def test_magma_with_a_trainable_vit_end_to_end(emul_ops, monkeypatch, tmp_path):
    """MAGMA_v1's optimizer settings on the ViT encoder: every trainable gradient (adapters, image prefix, all ViT
    parameters) against the oracle's autograd; then engine steps with image_enc_lr and a checkpoint round trip."""
    from magma_b200.train_loop import B200Engine
    from magma_b200.utils import load_model, save_model

    cfg = tiny_cfg()
    S, B = 16, 3
    w16 = oracle_weights(cfg)
    model, mc = build(monkeypatch, cfg, w16, S, lr=1e-2, image_enc_lr=1e-4, warmup_num_steps=2)
    model.eval()
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    assert any(n.startswith("image_prefix.enc.") for n in names) and model.arena.numel > 0
    images, captions = O.synthetic_batch(cfg, B, S, seed=11)
    images = images.to(torch.bfloat16).float()
    params = {k: v.clone().requires_grad_(k in names) for k, v in w16.items()}
    loss_o, logits_o, labels_o = O.magma_forward(images, captions, params, cfg)
    loss_o.backward()
    out = model(images, captions)
    assert abs(float(out.loss.detach()) - float(loss_o.detach())) < 2e-2
    assert rel(out.logits, logits_o.detach()) < 3e-2
    out.loss.backward()
    sd = dict(model.named_parameters())
    bad = {k: round(rel(sd[k].grad, params[k].grad), 4) for k in names if rel(sd[k].grad, params[k].grad) > 5e-2}
    assert not bad, bad
    # a second identical pass accumulates (gradient accumulation: grads are live)
    g1 = {k: sd[k].grad.clone() for k in names}
    model(images, captions).loss.backward()
    assert max(rel(sd[k].grad, 2 * g1[k]) for k in names) < 1e-2
    # engine: WarmupLR gives lr = 0 at step 0; the encoder moves at image_enc_lr / lr of the others' rate
    for p in model.parameters():
        p.grad = None
    model.arena.grad.zero_()
    model.train()
    eng = B200Engine(model, mc, n_buckets=2)
    before = {k: sd[k].detach().clone() for k in names}
    losses = []
    for _ in range(4):
        o = eng(images, captions)
        eng.backward(o.loss)
        eng.step()
        losses.append(float(o.loss.detach()))
    assert losses[-1] < losses[0] and abs(losses[1] - losses[0]) < 1e-6
    enc_move = max(float((sd[k].detach() - before[k]).abs().max()) for k in names if k.startswith("image_prefix.enc."))
    oth_move = max(float((sd[k].detach() - before[k]).abs().max()) for k in names if not k.startswith("image_prefix.enc."))
    assert 0 < enc_move < 5e-2 * oth_move
    assert len(eng._segments) == 3 and eng._segments[1][2] == pytest.approx(1e-2)
    # checkpoint round trip through save_model / load_model (magma/utils.py:89-117)
    save_model(eng, str(tmp_path), eng.global_step, config=mc)
    want = []
    for _ in range(2):
        o = eng(images, captions)
        eng.backward(o.loss)
        eng.step()
        want.append(float(o.loss.detach()))
    model2, mc2 = build(monkeypatch, cfg, w16, S, lr=1e-2, image_enc_lr=1e-4, warmup_num_steps=2)
    model2.train()
    eng2 = B200Engine(model2, mc2, n_buckets=2)
    assert load_model(eng2, str(tmp_path)) == 4 and eng2.global_step == 4
    got = []
    for _ in range(2):
        o = eng2(images, captions)
        eng2.backward(o.loss)
        eng2.step()
        got.append(float(o.loss.detach()))
    assert got == pytest.approx(want, abs=1e-5)


def test_magma_with_a_trainable_conv_trunk_trains(emul_ops, monkeypatch, tmp_path):
    """MAGMA_v1.yml's encoder type (a CLIP conv trunk, freeze_img_encoder: false): the whole step runs and learns, and a
    (trainable-only) checkpoint carries the BatchNorm running statistics the folded eval path reads."""
    from magma_b200.image_encoders import register_resnet
    from magma_b200.train_loop import B200Engine
    from magma_b200.utils import load_model, save_model

    cfg = tiny_cfg(vit_image=64)
    register_resnet("clip_resnet_dry", (1, 1, 1, 1), 16, 64)          # 4 prefix tokens of width 512
    S, B = 16, 4
    model, mc = build(monkeypatch, cfg, None, S, encoder="clip_resnet_dry", lr=5e-3, image_enc_lr=5e-4,
                      warmup_num_steps=2)
    model.lm.init_weights(seed=0)
    model.image_prefix.enc.init_weights(seed=1)
    model.finalize()
    model.train()
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    assert any("enc.layer4" in n for n in names) and any(n.endswith("bn1.weight") for n in names)
    images, captions = O.synthetic_batch(cfg, B, S, seed=3, prefix_len=4)
    eng = B200Engine(model, mc, n_buckets=2)
    rm0 = model.image_prefix.enc.bn1.running_mean.clone()
    losses = []
    for _ in range(6):
        o = eng(images.to(torch.bfloat16), captions)
        eng.backward(o.loss)
        eng.step()
        losses.append(float(o.loss.detach()))
    assert all(l == l for l in losses) and losses[-1] < losses[0] - 0.05, losses
    assert not torch.equal(rm0, model.image_prefix.enc.bn1.running_mean)   # BatchNorm ran in training mode
    # resume: a fresh model + load_model must give the same EVAL features (running_mean / running_var travel with the
    # trainable-only file) and a checkpoint stripped of them must be refused, not loaded with statistics of (0, 1)
    save_model(eng, str(tmp_path), eng.global_step, config=mc)
    payload = torch.load(tmp_path / f"global_step{eng.global_step}" / "mp_rank_00_model_states.pt", weights_only=False)
    assert payload["trainable_only"] and "image_prefix.enc.bn1.running_var" in payload["module"]
    assert not any(k.startswith("lm.transformer.h.0.attn") for k in payload["module"])       # frozen LM stays out
    model.eval()
    with torch.no_grad():
        want = model.image_prefix(images.to(torch.bfloat16)).float()
    model2, mc2 = build(monkeypatch, cfg, None, S, encoder="clip_resnet_dry", lr=5e-3, image_enc_lr=5e-4,
                        warmup_num_steps=2)
    model2.lm.init_weights(seed=0)
    model2.image_prefix.enc.init_weights(seed=1)
    model2.finalize()
    eng2 = B200Engine(model2, mc2, n_buckets=2)
    assert load_model(eng2, str(tmp_path)) == eng.global_step
    assert torch.equal(model2.image_prefix.enc.bn1.running_mean, model.image_prefix.enc.bn1.running_mean)
    model2.eval()
    with torch.no_grad():
        got = model2.image_prefix(images.to(torch.bfloat16)).float()
    assert rel(got, want) < 1e-6
    for k in [k for k in payload["module"] if k.endswith(("running_mean", "running_var"))]:
        del payload["module"][k]
    torch.save(payload, tmp_path / f"global_step{eng.global_step}" / "mp_rank_00_model_states.pt")
    with pytest.raises(RuntimeError, match="lacks"):
        eng2.load_checkpoint(str(tmp_path))


def _dp_worker(rank, world, port, q):
    """One data-parallel rank of the whole-model dry run (gloo): shard of the global batch -> engine step."""
    import ctypes
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from magma_b200 import _lib, dp, ops
        from magma_b200.magma import Magma
        from magma_b200.train_loop import B200Engine
        from magma_b200.utils import reduce_losses
        from oracle import build_emul

        L = ctypes.CDLL(build_emul.build())                 # what the emul_ops fixture does, in this child process
        L.mb200_last_error.restype = ctypes.c_char_p
        _lib._lib, ops._stream = L, (lambda: None)

        class MP:                                           # minimal monkeypatch stand-in for build()
            @staticmethod
            def setattr(obj, name, val):
                setattr(obj, name, val)

        cfg = tiny_cfg()
        S, GB = 16, 4
        w16 = oracle_weights(cfg)
        images, captions = O.synthetic_batch(cfg, GB, S, seed=11)
        images = images.to(torch.bfloat16)

        def run(model_world, shard):
            model, mc = build(MP, cfg, w16, S, lr=1e-2, image_enc_lr=1e-3, warmup_num_steps=2, gradient_clipping=0.0)
            model.train()
            eng = B200Engine(model, mc, n_buckets=2)
            eng.world = model_world
            lo, hi = shard
            out = []
            for _ in range(3):  # the body of train_step (train_loop.py:7-21) without its .cuda() copies
                o = eng(images[lo:hi], captions[lo:hi])
                eng.backward(o.loss)
                eng.step()
                out.append(float(reduce_losses(o.loss.detach())))
            return model, out

        lo, hi = dp.shard_batch(GB, rank, world)
        model, losses = run(world, (lo, hi))
        flat = model.arena.master.clone()
        # every rank must hold identical parameters after the exchanged steps
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        moved = float((flat - B200_initial(cfg, w16, S, MP)).abs().max()) > 0
        q.put((rank, same, moved, losses))
    except Exception as exc:  # surface the failure instead of letting the parent wait for its queue timeout
        q.put((rank, False, False, repr(exc)))
        raise
    finally:
        dist.destroy_process_group()


def B200_initial(cfg, w16, S, MP):
    model, _ = build(MP, cfg, w16, S)
    return model.arena.master.clone()


def test_two_rank_data_parallel_engine_step(emul_ops):
    """B200Engine over gloo, world size 2, whole model on emulated kernels: after three train_steps on different shards
    the ranks hold bit-identical parameters (the slice-wise gradient all-reduce + 1/world in the fused optimizer), the
    parameters moved, and train_step returned the cross-rank mean loss on every rank."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "ranks diverged"
    assert all(r[2] for r in res), "parameters did not move"
    assert res[0][3] == pytest.approx(res[1][3], abs=1e-6)      # reduce_losses: the same mean on both ranks


def test_magma_with_layernorm_and_scaled_adapters_end_to_end(emul_ops, monkeypatch):
    """adapter_config with add_layernorm (adapters.py:16-17) and scaled_parallel (adapters.py:57-61) through
    Magma.add_adapters -> language_model._cmodel_ex -> csrc/gptj_sched.cu, against the oracle's autograd."""
    mlp = {"adapter_type": "scaled_parallel", "downsample_factor": 4}
    attn = {"adapter_type": "normal", "downsample_factor": 8}
    cfg = tiny_cfg(mlp_adapter=mlp, attn_adapter=attn)
    S, B = 16, 3
    w = oracle_weights(cfg)
    g = torch.Generator().manual_seed(9)
    for l in range(cfg.n_layer):
        for pre in (f"lm.transformer.h.{l}.mlp", f"lm.transformer.h.{l}.attn"):   # add_layernorm shifts the indices
            for i, j in ((2, 3), (0, 1)):
                for sfx in ("weight", "bias"):
                    w[f"{pre}.adapter.{j}.{sfx}"] = w.pop(f"{pre}.adapter.{i}.{sfx}")
            w[f"{pre}.adapter.0.weight"] = (1.0 + 0.1 * torch.randn(cfg.d, generator=g)).to(torch.bfloat16).float()
            w[f"{pre}.adapter.0.bias"] = (0.1 * torch.randn(cfg.d, generator=g)).to(torch.bfloat16).float()
        w[f"lm.transformer.h.{l}.mlp.adapter_scale"] = torch.tensor([0.75 + 0.25 * l])
    ac = {"mlp": dict(mlp, add_layernorm=True), "attention": dict(attn, add_layernorm=True)}
    model, mc = build(monkeypatch, cfg, w, S, freeze_enc=True, adapter_config=ac)
    model.eval()
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    assert any(n.endswith("adapter_scale") for n in names) and any(n.endswith("adapter.0.weight") for n in names)
    images, captions = O.synthetic_batch(cfg, B, S, seed=11)
    images = images.to(torch.bfloat16).float()
    # the LM is what this test is about: feed the oracle's prefix embeddings instead of images
    with torch.no_grad():
        prefix = O.image_prefix(images, w, cfg).to(torch.bfloat16)
    params = {k: v.clone().requires_grad_(k in names) for k, v in w.items()}
    loss_o, logits_o, _ = O.magma_forward(None, captions, params, cfg, input_embeddings=prefix.float())
    loss_o.backward()
    out = model(None, captions, input_embeddings=prefix)
    assert abs(float(out.loss.detach()) - float(loss_o.detach())) < 2e-2 and rel(out.logits, logits_o.detach()) < 3e-2
    out.loss.backward()
    sd = dict(model.named_parameters())
    lm_names = [n for n in names if n.startswith("lm.")]
    bad = {k: round(rel(sd[k].grad, params[k].grad), 4) for k in lm_names if rel(sd[k].grad, params[k].grad) > 5e-2}
    assert not bad, bad


def test_gradient_accumulation_matches_the_mean_of_micro_batch_gradients(emul_ops, monkeypatch):
    """gradient_accumulation_steps = 2 (MAGMA_v1.yml uses 8): engine.backward scales each micro-batch loss by 1/2 and
    accumulates; engine.step only acts at the boundary (train_loop.py:9-19 under DeepSpeed)."""
    from magma_b200.train_loop import B200Engine

    cfg = tiny_cfg()
    S = 16
    w16 = oracle_weights(cfg)
    images, captions = O.synthetic_batch(cfg, 4, S, seed=21)
    images = images.to(torch.bfloat16)
    model, mc = build(monkeypatch, cfg, w16, S, freeze_enc=False, gradient_accumulation_steps=2, lr=1e-2,
                      warmup_num_steps=2)
    model.eval()
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    sd = dict(model.named_parameters())

    def grads_of(lo, hi):
        for p in model.parameters():
            p.grad = None
        model.arena.grad.zero_()
        model(images[lo:hi], captions[lo:hi]).loss.backward()
        return {n: sd[n].grad.clone() for n in names}

    ga, gb = grads_of(0, 2), grads_of(2, 4)
    for p in model.parameters():
        p.grad = None
    model.arena.grad.zero_()
    eng = B200Engine(model, mc, n_buckets=2)
    before = model.arena.master.clone()
    o = eng(images[0:2], captions[0:2])
    eng.backward(o.loss)
    eng.step()                                        # not a boundary: nothing moves, gradients stay
    assert torch.equal(model.arena.master, before) and eng.global_step == 0
    o = eng(images[2:4], captions[2:4])
    eng.backward(o.loss)
    got = {n: sd[n].grad.clone() for n in names}
    bad = {n: round(rel(got[n], 0.5 * (ga[n] + gb[n])), 4) for n in names if rel(got[n], 0.5 * (ga[n] + gb[n])) > 1e-2}
    assert not bad, bad
    eng.step()                                        # boundary: WarmupLR gives lr(0) = 0, but the step is counted
    assert eng.global_step == 1 and float(model.arena.grad.abs().sum()) == 0.0


@pytest.mark.parametrize("reference_names", [False, True])
def test_from_checkpoint_round_trip(emul_ops, monkeypatch, tmp_path, reference_names):
    """Magma.from_checkpoint (magma/magma.py:278-301) on a full state dict written in DeepSpeed's file layout — with this
    package's parameter names, and with the reference fork's names (attn.attention.*, mlp.c_fc / c_proj), which
    checkpoint.convert_reference_state_dict maps back: the loaded model reproduces the logits of the one that was saved."""
    from magma_b200 import checkpoint as ck
    from magma_b200.magma import Magma

    cfg = tiny_cfg()
    S = 16
    w16 = oracle_weights(cfg)
    model, mc = build(monkeypatch, cfg, w16, S, freeze_enc=True)
    model.eval()
    images, captions = O.synthetic_batch(cfg, 2, S, seed=5)
    images = images.to(torch.bfloat16)
    with torch.no_grad():
        want = model(images, captions).logits.float().clone()
    sd = {k: v for k, v in model.state_dict().items() if not k.startswith(("word_embedding.", "transformer."))}
    d = ck.save_training_checkpoint(tmp_path, "global_step0", sd, ck.arena_optimizer_state(model.arena), {"global_step": 0},
                                    reference_names=reference_names)
    path = str(tmp_path / "global_step0" / "mp_rank_00_model_states.pt")
    if reference_names:
        keys = torch.load(path, weights_only=False)["module"].keys()
        assert any(".attention.q_proj." in k for k in keys) and any(".c_fc." in k and k.startswith("lm.") for k in keys)
    loaded = Magma.from_checkpoint(mc, path, device="cpu")
    loaded.eos_token, loaded.image_token = cfg.eos_token, cfg.image_token
    with torch.no_grad():
        got = loaded(images, captions).logits.float()
    assert torch.equal(got, want)
    with pytest.raises(FileNotFoundError):
        Magma.from_checkpoint(mc, str(tmp_path / "nope.pt"), device="cpu")


def test_save_model_then_from_checkpoint_as_in_the_reference(emul_ops, monkeypatch, tmp_path, capsys):
    """The reference's workflow (train.py:120-140 -> README.md:74): save_model(...) then
    Magma.from_checkpoint(<tag>/mp_rank_00_model_states.pt). `full=True` writes a self-contained file; the default
    trainable-only file is loaded on top of the initialised frozen weights with a message, not an opaque key error."""
    from magma_b200.magma import Magma
    from magma_b200.train_loop import B200Engine
    from magma_b200.utils import save_model

    cfg = tiny_cfg()
    S = 16
    w16 = oracle_weights(cfg)
    model, mc = build(monkeypatch, cfg, w16, S, freeze_enc=True, lr=1e-2, warmup_num_steps=2)
    model.train()
    images, captions = O.synthetic_batch(cfg, 2, S, seed=5)
    images = images.to(torch.bfloat16)
    eng = B200Engine(model, mc, n_buckets=2)
    for _ in range(3):
        o = eng(images, captions)
        eng.backward(o.loss)
        eng.step()
    model.eval()
    with torch.no_grad():
        want = model(images, captions).logits.float().clone()
    save_model(eng, str(tmp_path / "full"), eng.global_step, config=mc, full=True)
    save_model(eng, str(tmp_path / "part"), eng.global_step, config=mc)
    f_full = tmp_path / "full" / "global_step3" / "mp_rank_00_model_states.pt"
    f_part = tmp_path / "part" / "global_step3" / "mp_rank_00_model_states.pt"
    assert f_part.stat().st_size < 0.5 * f_full.stat().st_size
    loaded = Magma.from_checkpoint(mc, str(f_full), device="cpu")
    loaded.eos_token, loaded.image_token = cfg.eos_token, cfg.image_token
    with torch.no_grad():
        assert torch.equal(loaded(images, captions).logits.float(), want)
    # trainable-only: frozen weights come from the constructor (random here), so only the trained tensors are checked
    capsys.readouterr()
    part = Magma.from_checkpoint(mc, str(f_part), device="cpu")
    assert "trainable-only checkpoint" in capsys.readouterr().out
    sd_m, sd_p = model.state_dict(), part.state_dict()
    trained = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trained and all(torch.equal(sd_m[n].float(), sd_p[n].float()) for n in trained)
