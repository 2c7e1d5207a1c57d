import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multihmr_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 32776, 3072, 1024
bn = int(sys.argv[1]) if len(sys.argv) > 1 else 512
a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) * 0.03).half()
bias = torch.randn(N, device=dev); out = torch.zeros(M, N, device=dev, dtype=torch.float16)
for _ in range(3):
    ops.gemm_f16(a, w, ops.EPI_BIAS_F16, out, bias=bias, block_n=bn)
torch.cuda.synchronize()
