"""Times mhmr_op_attention (ViT-L @896, batch 8): the product kernel, without its exponentials
(MHMR_ATTN_ABLATE=1) and with the barrier / TMEM protocol only (=4); one process each."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
from multihmr_b200 import ops
sys.path.insert(0, os.path.join(%r, "tools"))
from bench_ops import timeit
dev = torch.device("cuda:0")
B, T, D = 8, 4097, 1024
qkv = torch.randn(B * T, 3 * D, device=dev).half()
out = torch.empty(B * T, D, device=dev, dtype=torch.float16)
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
ms = timeit(lambda: ops.attention(qkv, B, T, D, out=out), iters=15, flush=flush)
print(json.dumps(dict(ms=round(ms, 4), tflops=round(4.0 * B * T * T * D / ms / 1e9, 1))))
''' % (ROOT, ROOT)

res = []
envs = [dict(), dict(MHMR_ATTN_ABLATE="1"), dict(MHMR_ATTN_ABLATE="4")]
for env in envs:
    e = dict(os.environ, MHMR_ATTN_VERBOSE="1")
    e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
    line = (r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "") + " | " + r.stderr.strip()[-200:]
    res.append(dict(env=env, result=line))
    print(res[-1], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "attn_ablate.json"), "w"), indent=1)
