#!/bin/bash
# validation of the rewritten skinny linear (+ PDL), the automatic SIMT-tail choice, then bench + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02f_pytest.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/r02f_pytest.log | tail -5
if grep -q "failed" gpurun_out/r02f_pytest.log; then
  MHMR_PDL=0 timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_pytest_nopdl.log 2>&1
  tail -3 gpurun_out/r02f_pytest_nopdl.log
fi
timeout 300 python tools/bench_ops.py --what attention --out gpurun_out/r02f_attn.json > gpurun_out/r02f_ops.log 2>&1
cat gpurun_out/r02f_ops.log | grep attention
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
cat gpurun_out/r02f_bench_n1.json
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_$c.json 2> gpurun_out/r02f_bench_$c.err
  cat gpurun_out/r02f_bench_$c.json | cut -c1-400
done
export MHMR_PROF_BATCH=8
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02f_launches.csv python tools/prof_forward.py > gpurun_out/r02f_prof_launch.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"skinny|attn_fwd|gemm_tc2_kernel<3>" -c 12 -o gpurun_out/r02f_full python tools/prof_forward.py > gpurun_out/r02f_prof_full.log 2>&1
ls -la gpurun_out/r02f_full.ncu-rep
