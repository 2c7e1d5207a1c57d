"""Tiny driver for ncu captures: runs each hot operator a few times on ViT-L @896 bs=8 shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multihmr_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "attention"
dev = torch.device("cuda:0")
B, T, D = 8, 4097, 1024
M = B * T
if what == "attention":
    qkv = torch.randn(M, 3 * D, device=dev).half()
    out = torch.empty(M, D, device=dev, dtype=torch.float16)
    for _ in range(3):
        ops.attention(qkv, B, T, D, out=out)
else:
    shapes = {"qkv": (3 * D, D, ops.EPI_BIAS_F16), "proj": (D, D, ops.EPI_LS_RESID_F32),
              "fc1": (4 * D, D, ops.EPI_BIAS_GELU_F16), "fc2": (D, 4 * D, ops.EPI_LS_RESID_F32)}
    N, K, epi = shapes[what]
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) * 0.03).half()
    bias, gamma = torch.randn(N, device=dev), torch.rand(N, device=dev)
    f16 = epi in (ops.EPI_BIAS_F16, ops.EPI_BIAS_GELU_F16)
    out = torch.zeros(M, N, device=dev, dtype=torch.float16 if f16 else torch.float32)
    for _ in range(3):
        ops.gemm_f16(a, w, epi, out, bias=bias, gamma=gamma)
torch.cuda.synchronize()
