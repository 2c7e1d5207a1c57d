"""Turns the ncu exports brought back in gpurun_out/ into the committed summaries under profiles/.
  python tools/summarize_profiles.py <tag> <launch-list csv> <ncu-rep with --set full>[,<ncu-rep>...]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"mhmr::\(anonymous namespace\)::|mhmr::<unnamed>::|unnamed>::|<unnamed>::|void ", "", name)[:64]


rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 10]
ix = {h: i for i, h in enumerate(rows[0])}
agg, tot = collections.OrderedDict(), 0.0
for r in rows[1:]:
    v = float(r[ix["Metric Value"]])
    v = v / 1000 if r[ix["Metric Unit"]] == "ns" else (v * 1000 if r[ix["Metric Unit"]] == "ms" else v)
    a = agg.setdefault(short(r[ix["Kernel Name"]]), [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as fh:
    fh.write(f"# {tag}: every kernel launch of ONE forward (multiHMR_896_L, batch 8, 1xB200)\n\n")
    fh.write("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/prof_forward.py`\n")
    fh.write("(cold-cache, serialised launches: compare SHARES, not absolutes)\n\n")
    fh.write(f"launches: {len(rows) - 1}, summed device time: {tot / 1000:.2f} ms\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |\n")

rows = []
for one in rep.split(","):   # several captures of the same command: same columns, concatenate the launches
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if not rows:
        rows = part
    else:
        assert part[0] == rows[0], "captures with different metric sets"
        rows += part[2:]
ix = {h: i for i, h in enumerate(rows[0])}
cols = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem KB")]
seen = set()
with open(os.path.join(out_dir, f"{tag}_kernels_full.md"), "w") as fh:
    fh.write(f"# {tag}: `ncu --set full --clock-control none` of the distinct kernels of one forward\n\n")
    fh.write("(first captured launch of each kernel; units as reported by ncu; report: gpurun_out/ scratch)\n\n")
    fh.write("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
    for r in rows[2:]:
        k = short(r[ix["Kernel Name"]])
        if k in seen or k.startswith("at::"):
            continue
        seen.add(k)
        vals = []
        for m, _ in cols:
            v = r[ix[m]] if m in ix else ""
            try:
                vals.append(f"{float(v):.2f}")
            except ValueError:
                vals.append(v)
        fh.write(f"| `{k}` | " + " | ".join(vals) + " |\n")
# DRAM traffic per launch of the dominant kernels -> profiles/ncu_traffic.json (read by bench.py for roofline.traffic)
import json

fam = {"attn_fwd_kernel": "attention", "gemm_tc2_kernel<0>": "gemm_qkv", "gemm_tc2_kernel<1>": "gemm_fc1",
       "gemm_tc2_kernel<3>": "gemm_proj",
       # LayerNorm folded into the GEMMs (default): qkv / fc1 consumers, proj / fc2 on the split residual stream
       "gemm_tc2_kernel<7>": "gemm_qkv", "gemm_tc2_kernel<8>": "gemm_fc1", "gemm_tc2_kernel<6>": "gemm_proj"}
traffic = {"_source": f"profiles/{tag}_kernels_full.md (ncu --set full --clock-control none, first captured launch of each "
                      f"kernel, multiHMR_896_L batch 8)"}
seen = set()
for r in rows[2:]:
    k = short(r[ix["Kernel Name"]])
    for prefix, name in fam.items():
        if k.startswith(prefix) and name not in seen:
            seen.add(name)
            def val(metric):
                v, u = float(r[ix[metric]]), rows[1][ix[metric]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic[name] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                             "kernel": k}
json.dump(traffic, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)
print("wrote", os.listdir(out_dir))
