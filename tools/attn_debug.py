"""Debug driver: one attention call per process, prints max error vs torch or the exception."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from multihmr_b200 import ops
B, T, D = [int(v) for v in sys.argv[1:4]]
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(B * 1000 + T + D)
qkv = torch.randn(B * T, 3 * D, generator=g).to(dev).half()
out = ops.attention(qkv, B, T, D)
torch.cuda.synchronize()
H = D // 64
q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
d = (out.float() - ref).abs()
rows = d.max(dim=1).values
bad = (rows > 4e-3).nonzero().flatten().tolist()
print("B,T,D", B, T, D, "max err", d.max().item(), "nan", torch.isnan(out).sum().item(), "bad rows", bad[:12], len(bad))
''' % ROOT

if __name__ == "__main__":
    cases = [a.split(",") for a in sys.argv[1:]] or [["1", "130", "128"], ["1", "257", "64"], ["1", "600", "64"]]
    for c in cases:
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, *c], capture_output=True, text=True, timeout=25)
            print(c, "rc", r.returncode, r.stdout.strip()[-400:], "|", r.stderr.strip()[-600:], flush=True)
        except subprocess.TimeoutExpired:
            print(c, "TIMEOUT (hang)", flush=True)
