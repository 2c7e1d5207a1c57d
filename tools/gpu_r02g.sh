#!/bin/bash
# folded LayerNorm (norm1 -> qkv, norm2 -> fc1 in the GEMM epilogues): unit test, full GPU suite, A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x -k "folded or public" > gpurun_out/r02g_fold_unit.log 2>&1
tail -3 gpurun_out/r02g_fold_unit.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/r02g_pytest.log | tail -5
grep -n "^FAILED\|^ERROR" gpurun_out/r02g_pytest.log | head -20
for f in 1 0 1 0; do
  MHMR_LN_FOLD=$f timeout 600 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_fold$f.json 2> gpurun_out/r02g_bench_fold$f.err
  cut -c1-200 gpurun_out/r02g_bench_fold$f.json
done
export MHMR_PROF_BATCH=8
MHMR_LN_FOLD=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/r02g_launches_fold1.csv python tools/prof_forward.py > gpurun_out/r02g_prof_fold1.log 2>&1
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_$c.json 2> gpurun_out/r02g_bench_$c.err
  cut -c1-200 gpurun_out/r02g_bench_$c.json
done
