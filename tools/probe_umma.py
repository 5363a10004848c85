"""Bring-up probe (GPU box): which UMMA descriptor convention does the hardware follow?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
for (N, K) in [(64, 64), (256, 64), (64, 128), (32, 16)]:
    g = torch.Generator().manual_seed(N + K)
    A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g)
    want = (A.double() @ B.double().T)
    for swap in (0, 1, 2, 3):
        D = torch.full((128, N), float("nan"), device=dev)
        a, b = A.to(dev), B.to(dev)
        err = lib.dcl_debug_umma_gemm(N, K, L.ptr(a), L.ptr(b), L.ptr(D), swap, L.stream_ptr())
        torch.cuda.synchronize()
        e = ((D.double().cpu() - want).abs().max() / want.abs().max()).item()
        print(f"N={N} K={K} variant={swap} (bit0 swap LBO/SBO, bit1 B MN-major): launch={err} rel_err={e:.3e}", flush=True)
