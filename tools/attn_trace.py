"""Runs mhmr_op_attention once in trace mode (MHMR_ATTN_ABLATE=7|8, MHMR_ATTN_TRACE=<file>)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multihmr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T, D = 8, 4097, 1024
qkv = torch.randn(B * T, 3 * D, device=dev).half()
out = torch.empty(B * T, D, device=dev, dtype=torch.float16)
ops.attention(qkv, B, T, D, out=out)
torch.cuda.synchronize()
