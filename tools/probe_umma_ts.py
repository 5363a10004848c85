"""Bring-up probe (GPU box): A-operand-from-TMEM form of tcgen05.mma — which TMEM layout does the hardware expect
for a bf16 A tile?  Prints the error of each variant against bf16(A) bf16(B)^T.  Run under `timeout`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_net_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
for (N, K) in [(64, 64), (256, 64), (128, 128), (32, 16)]:
    g = torch.Generator().manual_seed(N + K)
    A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g)
    want = A.bfloat16().double() @ B.bfloat16().double().T
    for variant in (0, 1):
        D = torch.full((128, N), float("nan"), device=dev)
        a, b = A.to(dev), B.to(dev)
        err = lib.dcl_debug_umma_ts_gemm(N, K, L.ptr(a), L.ptr(b), L.ptr(D), variant, L.stream_ptr())
        torch.cuda.synchronize()
        e = ((D.double().cpu() - want).abs().max() / want.abs().max()).item()
        print(f"N={N} K={K} variant={variant} ({'2 bf16 per column' if variant == 0 else '1 bf16 per column'}): "
              f"launch={err} rel_err={e:.3e}", flush=True)
