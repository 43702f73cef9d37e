"""Thin Python wrappers over the C ABI (include/magma_b200.h). Tensors are torch CUDA tensors used purely as
device-memory handles; every op is enqueued on torch's current CUDA stream."""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, check, lib

ACT_NONE, ACT_GELU_NEW, ACT_QUICK_GELU, ACT_RELU = 0, 1, 2, 3
DACT_NONE, DACT_GELU_NEW, DACT_RELU = 0, 1, 3


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _mat_meta(t, name):
    """(rows, cols, ld, nb0, nb1, bs0, bs1) of a [..., rows, cols] tensor with unit inner stride."""
    if t.dim() < 2 or t.dim() > 4:
        raise ValueError(f"{name}: expected 2-4 dims, got {t.dim()}")
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        raise ValueError(f"{name}: innermost stride must be 1, got {t.stride()}")
    rows, cols = t.shape[-2], t.shape[-1]
    ld = t.stride(-2) if rows > 1 else max(cols, t.stride(-2))
    nb0 = t.shape[-3] if t.dim() >= 3 else 1
    bs0 = t.stride(-3) if t.dim() >= 3 else 0
    nb1 = t.shape[-4] if t.dim() >= 4 else 1
    bs1 = t.stride(-4) if t.dim() >= 4 else 0
    return rows, cols, ld, nb0, nb1, bs0, bs1


def gemm(
    A,
    B,
    out=None,
    *,
    a_mn=False,
    b_mn=False,
    out_dtype=torch.bfloat16,
    alpha=1.0,
    bias=None,
    act=ACT_NONE,
    aux_out=None,
    aux_in=None,
    dact=DACT_NONE,
    res1=None,
    res2=None,
    accumulate=False,
    force_bn=0,
):
    """C[..., M, N] = epilogue(alpha * A @ B^T) on the tcgen05 GEMM core.

    A is [..., M, K] (a_mn=False) or [..., K, M] (a_mn=True); B is [..., N, K] (b_mn=False) or
    [..., K, N] (b_mn=True). Up to two leading batch dims with arbitrary (8-element aligned) strides.
    """
    if A.dtype != torch.bfloat16 or B.dtype != torch.bfloat16:
        raise TypeError("gemm operands must be bf16")
    ra, ca, lda, anb0, anb1, abs0, abs1 = _mat_meta(A, "A")
    rb, cb, ldb, bnb0, bnb1, bbs0, bbs1 = _mat_meta(B, "B")
    M, K = (ca, ra) if a_mn else (ra, ca)
    N, Kb = (cb, rb) if b_mn else (rb, cb)
    if K != Kb:
        raise ValueError(f"gemm K mismatch: {K} vs {Kb}")
    nb0, nb1 = max(anb0, bnb0), max(anb1, bnb1)
    # broadcast of a non-batched operand over the batch: stride 0 is not TMA-legal, so only allow equal batch
    if (anb0, anb1) != (nb0, nb1) or (bnb0, bnb1) != (nb0, nb1):
        raise ValueError("gemm: A and B must have identical batch dims")
    if out is None:
        shape = ([nb1] if A.dim() >= 4 else []) + ([nb0] if A.dim() >= 3 else []) + [M, N]
        out = torch.empty(shape, dtype=out_dtype, device=A.device)
    rc_, cc, ldc, cnb0, cnb1, cbs0, cbs1 = _mat_meta(out, "out")
    if (rc_, cc) != (M, N) or (cnb0, cnb1) != (nb0, nb1):
        raise ValueError(f"gemm: out shape {tuple(out.shape)} does not match M={M} N={N} batch=({nb1},{nb0})")
    g = GemmArgs()
    g.M, g.N, g.K, g.nb0, g.nb1 = M, N, K, nb0, nb1
    g.c_dtype = 1 if out.dtype == torch.float32 else 0
    if out.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("gemm out must be bf16 or f32")
    g.A.ptr, g.A.ld, g.A.bs0, g.A.bs1, g.A.mn_major = A.data_ptr(), lda, abs0, abs1, int(a_mn)
    g.B.ptr, g.B.ld, g.B.bs0, g.B.bs1, g.B.mn_major = B.data_ptr(), ldb, bbs0, bbs1, int(b_mn)
    g.C, g.ldc, g.c_bs0, g.c_bs1 = out.data_ptr(), ldc, cbs0, cbs1
    g.alpha, g.act, g.dact, g.accumulate = float(alpha), int(act), int(dact), int(bool(accumulate))
    for name, t in (("bias", bias), ("aux_out", aux_out), ("aux_in", aux_in), ("res1", res1), ("res2", res2)):
        if t is not None:
            if t.dtype != torch.bfloat16:
                raise TypeError(f"gemm {name} must be bf16")
            setattr(g, name, t.data_ptr())
    for name, t in (("aux_out", aux_out), ("aux_in", aux_in)):
        if t is not None and _mat_meta(t, name)[2:] != (ldc, cnb0, cnb1, cbs0, cbs1):
            raise ValueError(f"gemm {name} must share out's strides")
    ld_res = 0
    for name, t in (("res1", res1), ("res2", res2)):
        if t is not None:
            m = _mat_meta(t, name)
            if ld_res and m[2] != ld_res:
                raise ValueError("res1/res2 must share a row stride")
            ld_res = m[2]
    g.ld_res = ld_res
    g.force_bn = force_bn
    check(lib().mb200_gemm(ctypes.byref(g), _stream()))
    return out
