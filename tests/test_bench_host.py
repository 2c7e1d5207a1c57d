"""Host-side logic of bench.py (no GPU): the algorithmic FLOP figure behind `roofline.achieved`, the workload table,
the usable-CPU detection of the CPU baseline and the one-JSON-line guarantee of `emit`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_vit_flops_match_survey_8d():
    # 672_S 0.1969, 672_B 0.5895, 672_L 1.9172, 896_L 4.1295, 1288_L 12.1670 TFLOP per image
    want = {("dinov2_vits14", 672): 0.1969, ("dinov2_vitb14", 672): 0.5895, ("dinov2_vitl14", 672): 1.9172,
            ("dinov2_vitl14", 896): 4.1295, ("dinov2_vitl14", 1288): 12.1670}
    for (bb, s), tf in want.items():
        got = bench.vit_flops_per_image(bb, s) / 1e12
        assert abs(got - tf) < 5e-4 * tf + 5e-5, (bb, s, got, tf)


def test_workloads_are_the_baseline_configs():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    text = " ".join(base["configs"])
    assert "multiHMR_896_L" in base["metric"] and "bs=8" in base["metric"]
    c = bench.CONFIGS
    assert (c["c3"]["name"], c["c3"]["img_size"], c["c3"]["batch_per_gpu"]) == ("multiHMR_896_L", 896, 8)
    assert (c["c2"]["img_size"], c["c2"]["batch_per_gpu"]) == (672, 4) and "multiHMR_672_L" in text
    assert (c["c5"]["img_size"], c["c5"]["batch_per_gpu"], c["c5"]["target_persons_per_image"]) == (1288, 2, 20)
    assert "1288" in text
    try:
        bench.set_workload("c5")
        assert bench.METRIC == "images/sec multiHMR_1288_L_bedlam bs=2" and bench.WORKLOAD["img_size"] == 1288
    finally:
        bench.set_workload("c3")
    assert bench.METRIC == "images/sec multiHMR_896_L bs=8"


def test_usable_cpus_respects_affinity_and_quota():
    info = bench.usable_cpus()
    assert 1 <= info["usable"] <= info["affinity"] <= info["os_cpu_count"]
    if info["cgroup_quota"] is not None:
        assert info["usable"] <= int(info["cgroup_quota"] + 0.999)


def test_emit_puts_exactly_one_json_line_on_stdout():
    """Libraries write to fd 1 behind Python's back (NCCL banner, child processes): after `_reserve_stdout` fd 1 is
    stderr and only `emit` reaches the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._reserve_stdout(); "
            "os.write(1, b'noise from a library\\n'); print('python print noise'); "
            "bench.emit({'metric': 'm', 'value': 1.5})") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.5}
    assert "noise from a library" in r.stderr and "python print noise" in r.stderr
