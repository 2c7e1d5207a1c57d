#!/bin/bash
# Round-2 validation + measurement run (1 GPU).  Ordered by importance, every stage under its own timeout.
mkdir -p gpurun_out
# 1. the whole GPU suite on the default configuration (persistent attention, refinement on)
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02e_pytest.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/r02e_pytest.log | tail -5
# 2. headline bench (with the CPU leg and the 672x672 secondary leg)
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
cat gpurun_out/r02e_bench_n1.json
# 3. attention variants, isolated kernel timing (short timeouts: a hang costs 2 minutes, not 15)
echo "variants" > gpurun_out/r02e_ops.log
for v in "MHMR_ATTN_V1=1" "MHMR_ATTN_V1=0" "MHMR_ATTN_TAIL=1" "MHMR_ATTN_TOKEN=1" "MHMR_ATTN_TOKEN=2" "MHMR_ATTN_TAIL=1 MHMR_ATTN_TOKEN=1"; do
  tag=$(echo $v | tr ' =' '__')
  echo "== $v" >> gpurun_out/r02e_ops.log
  env $v timeout 150 python -m pytest tests/test_attention_gpu.py -q > gpurun_out/r02e_attn_test_$tag.log 2>&1
  tail -1 gpurun_out/r02e_attn_test_$tag.log >> gpurun_out/r02e_ops.log
  env $v timeout 150 python tools/bench_ops.py --what attention --out gpurun_out/r02e_attn_$tag.json >> gpurun_out/r02e_ops.log 2>&1
done
grep "==\|attention\|passed\|failed" gpurun_out/r02e_ops.log
# 4. the other BASELINE configs + the reference arm
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_$c.json 2> gpurun_out/r02e_bench_$c.err
  cat gpurun_out/r02e_bench_$c.json
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref.json 2> gpurun_out/r02e_bench_ref.err
cat gpurun_out/r02e_bench_ref.json
# 5. profiles: launch list of one forward, --set full of the main kernels, attention timeline
export MHMR_PROF_BATCH=8
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches.csv python tools/prof_forward.py > gpurun_out/r02_prof_launch.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"attn_fwd|gemm_tc2|layernorm|skinny|smplx_vertex|im2col|hph_cross|refine" -c 40 \
   -o gpurun_out/r02_full python tools/prof_forward.py > gpurun_out/r02_prof_full.log 2>&1
ls -la gpurun_out/r02_full.ncu-rep
MHMR_ATTN_ABLATE=7 MHMR_ATTN_TRACE=gpurun_out/r02_trace.bin timeout 200 python tools/attn_trace.py > gpurun_out/r02_trace.log 2>&1
python tools/attn_trace_report.py gpurun_out/r02_trace.bin gpurun_out/r02_attention_timeline.md >> gpurun_out/r02_trace.log 2>&1
tail -2 gpurun_out/r02_trace.log
