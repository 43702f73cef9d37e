"""Image encoders — drop-in for magma/image_encoders.py (`get_image_encoder`) re-backed by the CUDA runtime.

The reference obtains its encoders from openai/CLIP (`clip.load(name)[0].visual`, image_encoders.py:65) and timm;
neither package is vendored. This module implements CLIP's VisionTransformer family natively (patchify-conv as
im2col + tcgen05 GEMM, 24x [LN -> QKV GEMM+bias -> attention -> out GEMM+bias+residual -> LN -> fc GEMM+bias+
QuickGELU -> proj GEMM+bias+residual], ln_post(CLS) @ proj) with openai/CLIP's parameter names, and extends the
reference's name table with `clip_vit_large` (ViT-L/14, BASELINE.json config 2) following the reference's own ViT
convention (pooled [b, D] features; `"clip"` = ViT-B/32 -> 512, image_prefix.py:18). The conv-trunk encoders
(`nfresnet50`, `clip_resnet`, `clip_resnet_large`) are a later row of the scope table and raise.
"""
import ctypes

import torch
import torch.nn as nn

from . import ops
from ._lib import MB200Error, VitLayerC, VitModelC, check, lib

# name -> (width, layers, heads, patch, input_resolution, mlp, out_dim)
VIT_CONFIGS = {
    "clip": (768, 12, 12, 32, 224, 3072, 512),             # ViT-B/32 (image_encoders.py:56-57)
    "ViT-B/32": (768, 12, 12, 32, 224, 3072, 512),
    "clip_vit_large": (1024, 24, 16, 14, 224, 4096, 768),  # ViT-L/14 (extension)
    "ViT-L/14": (1024, 24, 16, 14, 224, 4096, 768),
}


def register_vit(name, width, layers, heads, patch, input_resolution, mlp, out_dim):
    """Add a CLIP-ViT geometry under `name` (must contain "clip", like the reference's dispatch at
    image_encoders.py:87) — used by the tests for small configurations."""
    from . import image_prefix

    VIT_CONFIGS[name] = (width, layers, heads, patch, input_resolution, mlp, out_dim)
    image_prefix.ENCODER_OUT_DIMS[name] = out_dim


def _p(*shape, device):
    return nn.Parameter(torch.empty(*shape, dtype=torch.bfloat16, device=device), requires_grad=False)


class _Lin(nn.Module):
    def __init__(self, i, o, device):
        super().__init__()
        self.weight, self.bias = _p(o, i, device=device), _p(o, device=device)


class _LN(nn.Module):
    def __init__(self, d, device):
        super().__init__()
        self.weight, self.bias = _p(d, device=device), _p(d, device=device)


class _Attn(nn.Module):
    def __init__(self, w, device):
        super().__init__()
        self.in_proj_weight, self.in_proj_bias = _p(3 * w, w, device=device), _p(3 * w, device=device)
        self.out_proj = _Lin(w, w, device)


class _MLP(nn.Module):
    def __init__(self, w, m, device):
        super().__init__()
        self.c_fc, self.c_proj = _Lin(w, m, device), _Lin(m, w, device)


class _ResBlock(nn.Module):
    def __init__(self, w, m, device):
        super().__init__()
        self.ln_1, self.attn, self.ln_2, self.mlp = _LN(w, device), _Attn(w, device), _LN(w, device), _MLP(w, m, device)


class _Transformer(nn.Module):
    def __init__(self, w, n, m, device):
        super().__init__()
        self.resblocks = nn.ModuleList([_ResBlock(w, m, device) for _ in range(n)])


class _Conv(nn.Module):
    def __init__(self, w, patch, device):
        super().__init__()
        self.weight = _p(w, 3, patch, patch, device=device)


class B200VisionTransformer(nn.Module):
    """CLIP VisionTransformer with openai/CLIP state-dict names (conv1, class_embedding, positional_embedding,
    ln_pre, transformer.resblocks.{i}.{ln_1,attn.{in_proj_weight,in_proj_bias,out_proj},ln_2,mlp.{c_fc,c_proj}},
    ln_post, proj). Attribute `input_resolution` is read by Magma.__init__ (magma/magma.py:69)."""

    def __init__(self, width, layers, heads, patch, input_resolution, mlp, out_dim, device=None):
        super().__init__()
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._device = dev
        self.input_resolution = input_resolution
        self.width, self.layers, self.heads, self.patch, self.mlp_dim, self.output_dim = width, layers, heads, patch, mlp, out_dim
        T = (input_resolution // patch) ** 2 + 1
        self.conv1 = _Conv(width, patch, dev)
        self.class_embedding = _p(width, device=dev)
        self.positional_embedding = _p(T, width, device=dev)
        self.ln_pre = _LN(width, dev)
        self.transformer = _Transformer(width, layers, mlp, dev)
        self.ln_post = _LN(width, dev)
        self.proj = _p(width, out_dim, device=dev)
        self._cache = None
        self._ws = {}

    @torch.no_grad()
    def init_weights(self, seed=0, std=0.02):
        g = torch.Generator(device=self._device).manual_seed(seed)
        for name, p in self.named_parameters():
            r = std * torch.randn(p.shape, generator=g, device=self._device)
            if name.endswith(("ln_1.weight", "ln_2.weight", "ln_pre.weight", "ln_post.weight")):
                r = 1.0 + r
            p.data.copy_(r)
        self._cache = None
        return self

    def invalidate(self):
        self._cache = None

    def _cmodel(self):
        if self._cache is not None:
            return self._cache
        for n, p in self.named_parameters():
            if p.dtype != torch.bfloat16 or not p.is_cuda:
                raise MB200Error(f"ViT parameter {n} must be bf16 on CUDA")
        K = 3 * self.patch * self.patch
        ldk = (K + 7) // 8 * 8
        conv = torch.zeros(self.width, ldk, dtype=torch.bfloat16, device=self._device)
        conv[:, :K] = self.conv1.weight.data.reshape(self.width, K)
        proj_t = self.proj.data.t().contiguous()
        layers = (VitLayerC * self.layers)()
        for i, b in enumerate(self.transformer.resblocks):
            L = layers[i]
            L.ln1_g, L.ln1_b = b.ln_1.weight.data_ptr(), b.ln_1.bias.data_ptr()
            L.w_qkv, L.b_qkv = b.attn.in_proj_weight.data_ptr(), b.attn.in_proj_bias.data_ptr()
            L.w_out, L.b_out = b.attn.out_proj.weight.data_ptr(), b.attn.out_proj.bias.data_ptr()
            L.ln2_g, L.ln2_b = b.ln_2.weight.data_ptr(), b.ln_2.bias.data_ptr()
            L.w_fc, L.b_fc = b.mlp.c_fc.weight.data_ptr(), b.mlp.c_fc.bias.data_ptr()
            L.w_proj, L.b_proj = b.mlp.c_proj.weight.data_ptr(), b.mlp.c_proj.bias.data_ptr()
        m = VitModelC()
        m.n_layer, m.width, m.n_head, m.patch = self.layers, self.width, self.heads, self.patch
        m.image, m.mlp, m.out_dim = self.input_resolution, self.mlp_dim, self.output_dim
        m.w_conv, m.ld_conv = conv.data_ptr(), ldk
        m.cls, m.pos = self.class_embedding.data_ptr(), self.positional_embedding.data_ptr()
        m.ln_pre_g, m.ln_pre_b = self.ln_pre.weight.data_ptr(), self.ln_pre.bias.data_ptr()
        m.ln_post_g, m.ln_post_b = self.ln_post.weight.data_ptr(), self.ln_post.bias.data_ptr()
        m.proj_t = proj_t.data_ptr()
        m.layers = ctypes.cast(layers, ctypes.POINTER(VitLayerC))
        self._cache = (m, layers, conv, proj_t)
        return self._cache

    def forward(self, x):
        """[b, 3, R, R] -> [b, out_dim] (pooled CLS features, like clip's `.visual`)."""
        if any(p.requires_grad for p in self.parameters()) and torch.is_grad_enabled():
            raise MB200Error("training the image encoder (freeze_img_encoder: false) is not supported yet: "
                             "the ViT backward pass is a later row of the scope table")
        B, C, R, R2 = x.shape
        if C != 3 or R != self.input_resolution or R2 != R:
            raise ValueError(f"expected [b,3,{self.input_resolution},{self.input_resolution}], got {tuple(x.shape)}")
        x = x.to(device=self._device, dtype=torch.bfloat16).contiguous()
        m = self._cmodel()[0]
        if B not in self._ws:
            n = lib().mb200_vit_workspace_bytes(ctypes.byref(m), B)
            self._ws[B] = torch.empty(n, dtype=torch.uint8, device=self._device)
        ws = self._ws[B]
        feats = torch.empty(B, self.output_dim, dtype=torch.bfloat16, device=self._device)
        check(lib().mb200_vit_forward(ctypes.byref(m), ops._ptr(x), ops._ptr(feats), B, ops._ptr(ws),
                                      ctypes.c_size_t(ws.numel()), ops._stream()))
        return feats


def clip_encoder(device=None, name: str = "clip") -> nn.Module:
    """magma/image_encoders.py:48-76."""
    if name in ("clip_resnet", "RN50x4", "clip_resnet_large", "RN50x16"):
        raise NotImplementedError(f"CLIP ModifiedResNet encoder '{name}' (conv trunk) is not re-backed yet "
                                  "(SURVEY.md §8f rank 1); use 'clip' or 'clip_vit_large'")
    if name not in VIT_CONFIGS:
        raise ValueError(f"encoder {name} not recognized")
    return B200VisionTransformer(*VIT_CONFIGS[name], device=device)


def get_image_encoder(name: str, device=None, pretrained: bool = False) -> nn.Module:
    """magma/image_encoders.py:79-91. Weights are uninitialised/random: no checkpoint source exists offline."""
    if name == "nfresnet50":
        raise NotImplementedError("nfresnet50 (timm conv trunk) is not re-backed yet (SURVEY.md §8f rank 1)")
    if "clip" in name or name in VIT_CONFIGS:
        return clip_encoder(device=device, name=name)
    raise ValueError(f"image encoder {name} not recognized")
