"""Data-parallel host logic (pure Python, device-agnostic so it is testable with gloo on CPU).

The reference's only parallelism is DP under DeepSpeed ZeRO-2 (train.py:103-111). Here every rank
holds the full (frozen) model and a flat fp32 gradient arena for the ~0.24 B trainable parameters; the arena is
laid out in the order gradients become ready in backward (last GPT-J layer first, image prefix last) and is
all-reduced slice by slice as backward proceeds."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

ALIGN = 64  # elements; keeps every parameter view 256-byte aligned


def arena_layout(numels: Sequence[int]) -> Tuple[List[int], int]:
    """Offsets (in elements) of each parameter in the flat arena and the total padded length."""
    offs, off = [], 0
    for n in numels:
        offs.append(off)
        off += (n + ALIGN - 1) // ALIGN * ALIGN
    return offs, off


def backward_order_key(name: str):
    """Sort key: GPT-J layers descending (their gradients are produced first), then everything else."""
    if ".transformer.h." in name:
        return (0, -int(name.split(".transformer.h.")[1].split(".")[0]))
    return (1, 0)


def layer_chunks(n_layer: int, n_buckets: int) -> List[Tuple[int, int]]:
    """[(layer_hi, layer_lo), ...] covering [0, n_layer) from the top down in n_buckets nearly equal chunks."""
    nb = max(1, min(n_buckets, n_layer))
    bounds = [round(i * n_layer / nb) for i in range(nb + 1)]
    return [(bounds[i + 1], bounds[i]) for i in reversed(range(nb))]


def slice_for(names: Sequence[str], numels: Sequence[int], offsets: Sequence[int], prefixes: Sequence[str]):
    """(lo, hi) element range of the arena covering all parameters whose name starts with one of `prefixes`
    (None, None when there is none)."""
    lo = hi = None
    for n, k, o in zip(names, numels, offsets):
        if any(n.startswith(p) for p in prefixes):
            e = o + (k + ALIGN - 1) // ALIGN * ALIGN
            lo = o if lo is None else min(lo, o)
            hi = e if hi is None else max(hi, e)
    return lo, hi


def allreduce_slice(flat_grad: torch.Tensor, lo, hi, group=None, comm_dtype=None):
    """SUM all-reduce of one contiguous arena slice (averaging by 1/world is folded into the optimizer kernel).
    comm_dtype=torch.bfloat16 exchanges a bf16 copy of the slice (half the NVLink bytes; the reference exchanged fp16
    gradients under DeepSpeed fp16, train.py:103-111) and writes the summed values back into the fp32 arena."""
    if lo is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    view = flat_grad[lo:hi]
    if comm_dtype is None or comm_dtype == view.dtype:
        return dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
    buf = view.to(comm_dtype)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    view.copy_(buf)
    return None


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank `rank` takes samples [lo, hi) of the global batch (rank i takes [8i, 8i+8))."""
    per = global_batch // world
    assert per * world == global_batch, "global batch must divide evenly across ranks"
    return rank * per, (rank + 1) * per


def optimizer_segments(names: Sequence[str], numels: Sequence[int], offsets: Sequence[int], no_decay: Sequence[bool],
                       lr: float, image_enc_lr, weight_decay: float, enc_prefix: str = "image_prefix.enc."):
    """Parameter groups of the reference optimizer (magma/utils.py:120-215) mapped onto the flat arena: contiguous runs
    [(lo, hi, lr, weight_decay), ...] of parameters sharing a learning rate and decay. The image encoder gets
    `image_enc_lr` when it is set (utils.py:173-177); LayerNorm / embedding parameters and biases are exempt from
    weight decay (utils.py:128-146). With weight_decay == 0 and no separate encoder rate this is ONE run — the whole
    arena in one fused kernel launch."""
    segs = []
    for n, k, o, nd in zip(names, numels, offsets, no_decay):
        plr = image_enc_lr if (image_enc_lr is not None and n.startswith(enc_prefix)) else lr
        pwd = 0.0 if (nd or weight_decay == 0.0) else weight_decay
        hi = o + (k + ALIGN - 1) // ALIGN * ALIGN
        if segs and segs[-1][1] == o and segs[-1][2] == plr and segs[-1][3] == pwd:
            segs[-1] = (segs[-1][0], hi, plr, pwd)
        else:
            segs.append((o, hi, plr, pwd))
    return segs
