#!/bin/bash
# residual prefetch (cp.async) in the split-stream epilogue of proj / fc2: GPU suite + bench c3 / c2 / c5
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x -k "folded or layerscale" > gpurun_out/r02k_unit.log 2>&1
tail -2 gpurun_out/r02k_unit.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/r02k_pytest.log | tail -3
grep -n "^FAILED\|^ERROR" gpurun_out/r02k_pytest.log | head
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err
cut -c1-160 gpurun_out/r02k_bench_n1.json; tail -2 gpurun_out/r02k_bench_n1.err
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_$c.json 2> gpurun_out/r02k_bench_$c.err
  cut -c1-160 gpurun_out/r02k_bench_$c.json
done
