#!/bin/bash
# refine chain with 16 persons per pass: full GPU suite + bench c3 / c2 / c5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/r02i_pytest.log | tail -5
grep -n "^FAILED\|^ERROR" gpurun_out/r02i_pytest.log | head -20
timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err
cut -c1-200 gpurun_out/r02i_bench_n1.json
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_$c.json 2> gpurun_out/r02i_bench_$c.err
  cut -c1-200 gpurun_out/r02i_bench_$c.json
done
