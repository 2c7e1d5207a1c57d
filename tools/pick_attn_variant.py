"""Reads the attention timings of the staged GPU run (gpurun_out/r02b_attn_*.json) and prints shell exports that
select the fastest validated variant (B=8, T=4097, D=1024 case)."""
import glob
import json
import os
import re

best = None
for path in sorted(glob.glob("gpurun_out/r02b_attn_*.json")):
    tag = re.sub(r".*r02b_attn_|\.json", "", path)
    try:
        rows = json.load(open(path))["attention"]
    except Exception:
        continue
    ms = [r["ms"] for r in rows if r["T"] == 4097][0]
    print(f"# {tag}: {ms} ms")
    env = {"new": "MHMR_ATTN_V1=0", "v1": "MHMR_ATTN_V1=1", "poly1": "MHMR_ATTN_V1=0 MHMR_ATTN_POLY=1",
           "poly2": "MHMR_ATTN_V1=0 MHMR_ATTN_POLY=2", "poly3": "MHMR_ATTN_V1=0 MHMR_ATTN_POLY=3",
           "tail_poly0": "MHMR_ATTN_V1=0 MHMR_ATTN_TAIL=1", "tail_poly2": "MHMR_ATTN_V1=0 MHMR_ATTN_TAIL=1 MHMR_ATTN_POLY=2",
           "helper_tail0": "MHMR_ATTN_V1=0 MHMR_ATTN_HELPER=1", "helper_tail1": "MHMR_ATTN_V1=0 MHMR_ATTN_HELPER=1 MHMR_ATTN_TAIL=1",
           "token1_tail1": "MHMR_ATTN_V1=0 MHMR_ATTN_TOKEN=1 MHMR_ATTN_TAIL=1",
           "token2_tail1": "MHMR_ATTN_V1=0 MHMR_ATTN_TOKEN=2 MHMR_ATTN_TAIL=1",
           "token1_helper_tail1": "MHMR_ATTN_V1=0 MHMR_ATTN_TOKEN=1 MHMR_ATTN_HELPER=1 MHMR_ATTN_TAIL=1"}.get(tag)
    if env is None:
        continue
    # a variant only qualifies if its unit tests passed
    need = {"token1_tail1": ["token_test", "tail_test"], "token2_tail1": ["token_test", "tail_test"],
            "token1_helper_tail1": ["token_test", "tail_test", "helper_test"]}.get(tag)
    if need is not None:
        ok_all = True
        for lg in need:
            pth = f"gpurun_out/r02b_attn_{lg}.log"
            txt = open(pth).read() if os.path.exists(pth) else "failed"
            ok_all = ok_all and ("failed" not in txt) and ("passed" in txt)
        if ok_all and (best is None or ms < best[0]):
            best = (ms, env, tag)
        continue
    log = {"poly1": "poly2_test", "poly2": "poly2_test", "poly3": "poly2_test", "tail_poly0": "tail_test",
           "tail_poly2": "tail_test", "helper_tail0": "helper_test", "helper_tail1": "helper_test"}.get(tag)
    ok = True
    if tag == "new":
        ok = "NEW ATTENTION FAILED" not in open("gpurun_out/r02b_attn.log").read()
    if log is not None:
        txt = open(f"gpurun_out/r02b_attn_{log}.log").read() if os.path.exists(f"gpurun_out/r02b_attn_{log}.log") else "failed"
        ok = ("failed" not in txt) and ("error" not in txt.lower()) and ("passed" in txt)
        if tag == "tail_poly2":
            t2 = open("gpurun_out/r02b_attn_poly2_test.log").read() if os.path.exists("gpurun_out/r02b_attn_poly2_test.log") else "failed"
            ok = ok and ("failed" not in t2) and ("passed" in t2)
        if tag == "helper_tail1":
            t2 = open("gpurun_out/r02b_attn_tail_test.log").read() if os.path.exists("gpurun_out/r02b_attn_tail_test.log") else "failed"
            ok = ok and ("failed" not in t2) and ("passed" in t2)
    if ok and (best is None or ms < best[0]):
        best = (ms, env, tag)
if best:
    print(f"# best: {best[2]} {best[0]} ms")
    print("export " + best[1])
