"""Host-side helpers mirroring magma/utils.py for the hot path: build_labels (device kernel), tokenizer seam,
distributed helpers (torch.distributed over NCCL, env contract of utils.py:255-269)."""
import os

import torch
import torch.distributed as dist

from . import ops


def is_main():
    return (not dist.is_initialized()) or dist.get_rank() == 0


def print_main(*msg):
    if is_main():
        print(*msg)


def reduce_losses(losses):
    """magma/utils.py:26-34 — SUM all-reduce / world size."""
    if dist.is_initialized():
        losses = losses.detach().clone()
        dist.all_reduce(losses, dist.ReduceOp.SUM)
        return losses / dist.get_world_size()
    return losses


def cycle(loader):
    """magma/utils.py:37-40 — endless iterator over a DataLoader (what train_step's `next(train_loader)` pulls from)."""
    while True:
        for data in loader:
            yield data


def collate_fn(batch_data, seq_len=2048):
    """magma/datasets/dataset.py:155-160: [(image [1,3,R,R], caption ids [1,n]), ...] -> (images [b,3,R,R],
    captions [b, <=seq_len])."""
    images, captions = zip(*batch_data)
    return torch.cat(images), torch.cat([c[:, :seq_len] for c in captions])


def count_parameters(model):
    """magma/utils.py:241-245 — number of trainable parameters."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def get_world_info():
    """magma/utils.py:255-259."""
    return int(os.environ["LOCAL_RANK"]), int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])


def save_model(model_engine, save_dir, global_step, config=None, full=None):
    """magma/utils.py:89-96 — config.yml next to the engine checkpoint, client state {global_step, config}.
    `full=True` (or MB200_SAVE_FULL=1; not a config field, so config.yml stays readable by the reference's schema) also
    writes the frozen LM / encoder weights, which makes <tag>/mp_rank_00_model_states.pt loadable by
    `Magma.from_checkpoint` on its own as in the reference; the default keeps the 12 GB of weights that never change out
    of every periodic save (from_checkpoint then loads it on top of the initialised frozen weights and says so)."""
    import yaml

    os.makedirs(save_dir, exist_ok=True)
    cfg = config.to_dict() if config is not None else None
    if cfg is not None and is_main():
        with open(os.path.join(str(save_dir), "config.yml"), "w") as f:
            yaml.dump(cfg, f, default_flow_style=False)
    if full is None:
        full = os.environ.get("MB200_SAVE_FULL", "0") == "1"
    model_engine.save_checkpoint(save_dir, client_state={"global_step": global_step, "config": cfg},
                                 trainable_only=not full)


def load_model(model_engine, load_dir, load_optimizer_states=True, load_lr_scheduler_states=True):
    """magma/utils.py:99-117 — returns the global step to resume from, 0 when nothing could be loaded."""
    try:
        load_path, sd = model_engine.load_checkpoint(load_dir, load_optimizer_states=load_optimizer_states,
                                                     load_lr_scheduler_states=load_lr_scheduler_states)
    except AssertionError as e:
        load_path, sd = None, None
        print(e)
    if load_path is None:
        print("Model loading failed - starting from global step 0")
        return 0
    return sd["global_step"]


def build_labels(input_embeddings, captions, eos_token, device=None):
    """Drop-in for magma/utils.py:334-364 (same signature). The Python double loop of the reference (one device
    sync per token) is one integer kernel here; the result is bit-identical."""
    shape = input_embeddings.shape[:2]
    assert captions.shape[1] >= shape[1]
    if shape[1] == 0:  # reference quirk (utils.py:352-355): captions[:, :-0] is empty -> labels is [b, 0]
        return torch.empty(captions.shape[0], 0, dtype=torch.int64, device=captions.device)
    return ops.build_labels(captions.contiguous(), int(shape[1]), int(eos_token))


def no_weight_decay_names(model):
    """Names of the parameters the reference exempts from weight decay (magma/utils.py:120-146): every parameter of a
    LayerNorm or Embedding module, and every parameter called `bias`."""
    import torch.nn as nn

    from .image_encoders import _LN
    from .language_model import LayerNorm, WordEmbedding

    out = set()
    for mname, mod in model.named_modules():
        exempt = isinstance(mod, (nn.LayerNorm, nn.Embedding, _LN, LayerNorm, WordEmbedding))
        for pname, p in mod._parameters.items():
            if p is not None and (exempt or pname == "bias"):
                out.add(f"{mname}.{pname}" if mname else pname)
    return out


def configure_param_groups(model, config):
    """magma/utils.py:164-215 for the flat arena: [(lo, hi, lr_scale, weight_decay)] runs of `model.arena`, where
    lr_scale multiplies the scheduled learning rate (the image encoder's group trains at image_enc_lr / lr of it —
    DeepSpeed's WarmupDecayLR scales every group's own max lr by the same schedule factor, config.py:103-123)."""
    from . import dp

    arena = model.arena
    nd = no_weight_decay_names(model)
    enc_scale = None
    if getattr(config, "image_enc_lr", None) is not None and config.lr:
        enc_scale = float(config.image_enc_lr) / float(config.lr)
    return dp.optimizer_segments(arena.names, [p.numel() for p in arena.params], arena.offsets,
                                 [n in nd for n in arena.names], 1.0, enc_scale,
                                 float(getattr(config, "weight_decay", 0.0) or 0.0))


class IdTokenizer:
    """Offline stand-in for GPT2TokenizerFast + '<|image|>' (magma/utils.py:43-58): same ids and length
    (eos=pad=50256, cls=50257, len=50258). Text <-> id conversion needs the GPT-2 vocabulary files, which are not
    available offline; encode/decode here are byte-level placeholders."""

    cls_token_id = 50257
    eos_token_id = 50256
    pad_token_id = 50256
    padding_side = "right"

    def __init__(self, sequence_length=2048):
        self.model_max_length = sequence_length

    def __len__(self):
        return 50258

    def encode(self, text, return_tensors=None, **_):
        ids = list(text.encode("utf-8"))
        return torch.tensor([ids], dtype=torch.long) if return_tensors == "pt" else ids

    def decode(self, ids, **_):
        return bytes(int(i) % 256 for i in ids).decode("utf-8", errors="replace")


def get_tokenizer(name="gpt2", sequence_length=2048):
    """magma/utils.py:43-58. Uses the real GPT-2 tokenizer when its files are cached locally, else IdTokenizer."""
    if name != "gpt2":
        raise ValueError(f"Tokenizer {name} not recognized")
    try:
        from transformers import GPT2TokenizerFast

        tok = GPT2TokenizerFast.from_pretrained("gpt2", local_files_only=True)
        tok.pad_token_id = tok.eos_token_id
        tok.padding_side = "right"
        tok.model_max_length = sequence_length
        tok.add_special_tokens({"cls_token": "<|image|>"})
        if len(tok) != 50258 or tok.eos_token_id != 50256 or tok.cls_token_id != 50257:
            raise RuntimeError("incomplete local GPT-2 tokenizer files")
        return tok
    except Exception:
        return IdTokenizer(sequence_length)


def init_distributed(backend="nccl"):
    """One process per GPU; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (utils.py:255-269)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            import datetime

            # a short collective timeout: a rank that stops participating must fail fast, not hang the box
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                    timeout=datetime.timedelta(seconds=int(os.environ.get("MB200_NCCL_TIMEOUT_S", "180"))))
        else:
            dist.init_process_group(backend)
        return dist.get_rank(), world, local_rank
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return 0, 1, local_rank
