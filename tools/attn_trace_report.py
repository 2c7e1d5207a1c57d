"""Turns an attention timeline trace (MHMR_ATTN_ABLATE=7 MHMR_ATTN_TRACE=file, tools/attn_trace.py) into a
markdown table: SM clocks of the per-key-tile events of the sub-partition-0 softmax warps of both query tiles."""
import sys

import numpy as np

path, out = sys.argv[1], sys.argv[2]
a = np.fromfile(path, dtype=np.uint32).reshape(8, 40, 16).astype(np.int64)
names = ["S in regs", "args ready", "token taken", "exp run done", "sum+pack done", "S(j+1) ready", "P stored"]
lines = ["# Attention kernel: SM-clock timeline of one CTA (ViT-L @896, batch 8, steady state)", "",
         "`MHMR_ATTN_ABLATE=7 MHMR_ATTN_TRACE=trace.bin python tools/attn_trace.py`; clocks relative to the moment the",
         "tile-0 warp of SM sub-partition 0 has S(j) in registers.  T0 / T1 = softmax warp of query tile 0 / 1 that",
         "share the sub-partition's MUFU unit.", ""]
for c in range(2):
    t = a[c]
    lines += [f"## traced CTA {c}", "", "| key tile | " + " | ".join("T0 " + n for n in names) + " | " +
              " | ".join("T1 " + n for n in names) + " | period |", "|---:|" + "---:|" * 15]
    for j in range(4, 12):
        base = int(t[j, 0])
        e0 = [(int(t[j, k]) - base) & 0xffffffff for k in range(7)]
        e1 = [(int(t[j, 8 + k]) - base) & 0xffffffff for k in range(7)]
        per = (base - int(t[j - 1, 0])) & 0xffffffff
        lines.append(f"| {j} | " + " | ".join(str(v) for v in e0 + e1) + f" | {per} |")
    lines.append("")
per = [(int(a[c][j, 0]) - int(a[c][j - 1, 0])) & 0xffffffff for c in range(2) for j in range(4, 30)]
lines += [f"median period per key tile (two query tiles): {int(np.median(per))} clk = "
          f"{2 * 1024 / np.median(per):.2f} of the MUFU.EX2 bound (2 x 1024 clk per sub-partition)", ""]
open(out, "w").write("\n".join(lines))
print("\n".join(lines[-3:]))
