#!/bin/bash
# refinement chain v2 (weights in shared memory, lane-per-column fc1, packed FMAs): GPU suite + bench c3 / c2 / c5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_fullsize_gpu.py tests/test_stages_gpu.py -q -x > gpurun_out/r02i_pytest.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/r02i_pytest.log | tail -5
grep -n "^FAILED\|^ERROR" gpurun_out/r02i_pytest.log | head -20
timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err
cut -c1-160 gpurun_out/r02i_bench_n1.json; tail -3 gpurun_out/r02i_bench_n1.err
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_$c.json 2> gpurun_out/r02i_bench_$c.err
  cut -c1-160 gpurun_out/r02i_bench_$c.json
done
