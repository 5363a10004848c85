"""Bring-up probe (GPU box): CTA-pair (cta_group::2, M=256) UMMA conventions.

Prints, per (N, K, mode), the error of each 128-row x N/2-column quadrant of D against the expected product, so a
wrong operand split shows up as a pattern instead of a bare failure.  Run under `timeout`.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
for mode in (0, 1):
    for (N, K) in [(64, 64), (256, 64), (128, 128), (32, 16)]:
        g = torch.Generator().manual_seed(N + K)
        A, B = torch.randn(256, K, generator=g), torch.randn(N, K, generator=g)
        want = (A.double() @ B.double().T)
        D = torch.full((256, N), float("nan"), device=dev)
        a, b = A.to(dev), B.to(dev)
        err = lib.dcl_debug_umma_pair_gemm(N, K, L.ptr(a), L.ptr(b), L.ptr(D), mode, L.stream_ptr())
        torch.cuda.synchronize()
        got = D.double().cpu()
        scale = want.abs().max()
        quads = []
        for r in range(2):
            for c in range(2):
                blk = (slice(128 * r, 128 * r + 128), slice(c * N // 2, (c + 1) * N // 2))
                quads.append(((got[blk] - want[blk]).abs().max() / scale).item())
        print(f"mode={mode} N={N} K={K}: launch={err} quadrant rel_err (r0c0 r0c1 r1c0 r1c1) = "
              + " ".join(f"{q:.2e}" for q in quads), flush=True)
