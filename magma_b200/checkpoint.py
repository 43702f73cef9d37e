"""Checkpoint key adapter for `Magma.from_checkpoint` (magma/magma.py:278-301).

The published MAGMA checkpoint (`mp_rank_00_model_states.pt["module"]`, README.md:74) was written by the reference
with the finetuneanon/transformers GPT-Neo(jax, rotary) fork as the LM, whose parameter names differ from the HF GPT-J
names this package uses (language_model.py). The fork is not vendored, so the fork-side names below are RECALLED from
its GPT-Neo lineage, not read from source (parity unpinned): `attn.attention.{q,k,v,out}_proj`, `mlp.c_fc`,
`mlp.c_proj`, causal-mask buffers `attn.attention.{bias,masked_bias}`. Anything that does not match a rule is passed
through unchanged and reported, so a differing real checkpoint fails loudly in `load_state_dict` rather than silently.
"""
import re
from typing import Dict, List, Tuple

import torch

# (pattern, replacement) applied in order to every key
_RENAMES: List[Tuple[re.Pattern, str]] = [
    (re.compile(r"\.attention\.(q_proj|k_proj|v_proj|out_proj)\."), r".\1."),  # attn.attention.q_proj -> attn.q_proj
    (re.compile(r"\.c_fc\."), ".fc_in."),
    (re.compile(r"\.c_proj\."), ".fc_out."),
]
# non-parameter buffers of the fork's attention (causal mask, mask fill value) and learned positions (absent with rotary)
_DROP = re.compile(r"(\.attention\.(bias|masked_bias)$)|(\.attn\.(bias|masked_bias)$)|(\.transformer\.wpe\.)")


def convert_reference_state_dict(sd: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, list]]:
    """Fork-named reference state dict -> this package's names. Returns (converted, report) with
    report = {"renamed": [(old, new)], "dropped": [keys], "collisions": [keys]}."""
    out, report = {}, {"renamed": [], "dropped": [], "collisions": []}
    for k, v in sd.items():
        if not k.startswith("image_prefix.") and _DROP.search(k):
            report["dropped"].append(k)
            continue
        nk = k
        if not k.startswith("image_prefix."):  # CLIP's own ViT blocks are legitimately named mlp.c_fc / mlp.c_proj
            for pat, rep in _RENAMES:
                nk = pat.sub(rep, nk)
        if nk != k:
            report["renamed"].append((k, nk))
        if nk in out:
            report["collisions"].append(nk)
        out[nk] = v
    return out, report


def to_reference_names(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse mapping (this package -> fork names), for writing checkpoints the reference can read."""
    out = {}
    for k, v in sd.items():
        nk = k
        if not k.startswith("image_prefix."):
            nk = re.sub(r"\.(attn|attn_block|module)\.(q_proj|k_proj|v_proj|out_proj)\.",
                        lambda m: f".{m.group(1)}.attention.{m.group(2)}.", k)
            nk = nk.replace(".fc_in.", ".c_fc.").replace(".fc_out.", ".c_proj.")
        out[nk] = v
    return out
