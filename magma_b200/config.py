"""MultimodalConfig — same field names / YAML format as the reference (magma/config.py:20-94) so the reference's
`configs/*.yml` load unchanged. The DeepSpeed-specific derived dicts of the reference (`config.py:96-137`) are
replaced by plain optimizer/scheduler fields consumed by magma_b200.train_loop.B200Engine."""
import uuid
from dataclasses import asdict, dataclass
from pathlib import Path

import yaml


def load_config(path, config_dir=Path("configs")):
    path = str(path)
    if not path.endswith(".yml"):
        path += ".yml"
    p = Path(path)
    if not p.exists():
        p = Path(config_dir) / path
    with open(p, "r") as f:
        return yaml.safe_load(f)


@dataclass
class MultimodalConfig:
    # training
    batch_size: int
    train_steps: int
    optimizer_name: str = "AdamW"
    lr: float = 8.0e-4
    image_enc_lr: float = None
    min_lr: float = 0.0
    lr_decay_iters: int = None
    gradient_accumulation_steps: int = 1
    image_size: int = 256
    eval_every: int = 250
    eval_steps: int = 25
    zero_stage: int = 2
    gradient_clipping: float = 1.0
    warmup_num_steps: int = 100
    weight_decay: float = 0.00
    run_blind: bool = False
    fine_tune: bool = False
    load_optimizer: bool = True
    # checkpointing
    save_every: int = 2500
    save: str = None
    load: str = None
    # data
    train_dataset_name: str = "conceptual_captions"
    eval_dataset_name: str = "/data/conceptual_captions"
    train_dataset_dir: str = "/data/coco_data"
    eval_dataset_dir: str = "/data/coco_data"
    eval_dataset_pct: float = 0.1
    # architecture
    encoder_name: str = "clip"
    tokenizer_name: str = "gpt2"
    lm_name: str = "EleutherAI/gpt-j-6B"
    image_seq_len: int = 2
    pretrained_img_encoder: bool = False
    seq_len: int = None  # present but never read by the reference; magma_b200 uses it to set Magma.seq_len
    # freezing
    freeze_lm: bool = True
    freeze_img_encoder: bool = True
    image_embed_dropout_prob: float = 0.0
    use_image_embed_layernorm: bool = False
    # adapters
    adapter_config: dict = None
    class_dict: dict = None
    # logging
    name: str = None
    log_every: int = 1
    wandb_project: str = "magma"

    def __post_init__(self):
        self.is_classifier = self.class_dict is not None
        if self.adapter_config is None:
            self.adapter_config = {}
        self.lr_scheduler = "WarmupLR" if self.lr_decay_iters is None else "WarmupDecayLR"
        if self.name is None:
            self.name = str(uuid.uuid4())[:8]

    @classmethod
    def from_yml(cls, path):
        return cls(**load_config(path))

    def to_dict(self):
        return asdict(self)

    def lr_at(self, step: int) -> float:
        """WarmupLR / WarmupDecayLR as configured for DeepSpeed in the reference (config.py:103-123):
        log-warmup to `lr` over warmup_num_steps, then constant or linear decay to min_lr at lr_decay_iters."""
        import math

        w = max(self.warmup_num_steps, 2)
        if step < w:
            gamma = math.log(step + 1) / math.log(w)
            return self.min_lr + (self.lr - self.min_lr) * gamma
        if self.lr_decay_iters is None:
            return self.lr
        frac = max(0.0, (self.lr_decay_iters - step) / max(1.0, self.lr_decay_iters - w))
        return self.min_lr + (self.lr - self.min_lr) * frac
