#!/bin/bash
# profiling run (1 GPU): ncu launch list of one forward, --set full captures of the attention / GEMM / LN / refine
# kernels, then the bench with the CPU leg.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export MHMR_PROF_BATCH=8
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches.csv python tools/prof_forward.py > gpurun_out/r02_prof_launch.log 2>&1
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"attn_fwd|gemm_tc2|layernorm|skinny|smplx_vertex|im2col|hph_cross|pack_records" -c 60 \
   -o gpurun_out/r02_full python tools/prof_forward.py > gpurun_out/r02_prof_full.log 2>&1
ls -la gpurun_out/r02_full.ncu-rep
MHMR_ATTN_ABLATE=7 MHMR_ATTN_TRACE=gpurun_out/r02_trace.bin timeout 300 python tools/attn_trace.py > gpurun_out/r02_trace.log 2>&1
python tools/attn_trace_report.py gpurun_out/r02_trace.bin gpurun_out/r02_attention_timeline.md >> gpurun_out/r02_trace.log 2>&1
tail -2 gpurun_out/r02_trace.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cat gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
cat gpurun_out/r02_bench_ref.json
for c in c2 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  cat gpurun_out/r02_bench_$c.json
done
