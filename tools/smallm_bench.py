"""Device-side timing of the small-M (decode) GEMMs: each shape is launched over a rotation of weight buffers larger than
L2 (so every launch streams its weights from HBM), captured in a CUDA graph (no host launch cost in the number), for the
tile widths and K splits selected through MB200_SMALLM_BN / MB200_SMALLM_SPLIT; "plan" is what plan_small_m picks.

  python tools/smallm_bench.py [--m 32] [--grid]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

SHAPES = [  # (name, N, K) of the GPT-J-6B decode step
    ("qkv", 12288, 4096),
    ("out", 4096, 4096),
    ("fc_in", 16384, 4096),
    ("fc_out", 4096, 16384),
    ("adapter_down", 1024, 4096),
    ("adapter_up", 4096, 1024),
    ("lm_head", 50258, 4096),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--grid", action="store_true", help="sweep tile width x split instead of the plan only")
    a = ap.parse_args()
    import torch

    from magma_b200 import ops

    dev = torch.device("cuda:0")
    M = a.m
    ws = torch.empty(16 * M * 4096, device=dev, dtype=torch.float32)
    for name, N, K in SHAPES:
        nbuf = max(2, int(300e6 // (N * K * 2)) + 1)
        Bs = [(torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16) for _ in range(nbuf)]
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        ldc = (N + 7) // 8 * 8
        C = torch.empty(M, ldc, device=dev, dtype=torch.bfloat16)[:, :N]
        want = (A.float() @ Bs[0].float().t())
        combos = [(None, None)]
        if a.grid:
            combos += [(bn, sp) for bn in (64, 128, 256) for sp in (1, 2, 3, 4, 6, 8)]
        for bn, sp in combos:
            for k, v in (("MB200_SMALLM_BN", bn), ("MB200_SMALLM_SPLIT", sp)):
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = str(v)
            ops.gemm(A, Bs[0], out=C, splitk_ws=ws)
            err = ((C.float() - want).norm() / want.norm()).item()
            reps = 2 * nbuf
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for i in range(3):
                    ops.gemm(A, Bs[i % nbuf], out=C, splitk_ws=ws)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for i in range(reps):
                        ops.gemm(A, Bs[i % nbuf], out=C, splitk_ws=ws)
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(3):
                    g.replay()
                e1.record(s)
                torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / (3 * reps) * 1e3
            tag = "plan" if bn is None else f"bn={bn} split={sp}"
            print(f"[SMALLM] {name:13s} M={M} N={N} K={K} {tag:18s} {us:7.1f} us  {N * K * 2 / us / 1e3:6.0f} GB/s  "
                  f"rel_err={err:.1e}", flush=True)
        del Bs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
