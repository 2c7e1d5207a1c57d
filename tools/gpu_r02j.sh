#!/bin/bash
# final evidence of round 2: GPU suite, bench lines (c3 with the CPU leg, c2, c5, reference arm), ncu launch list and
# full-section captures of the main kernels of the final library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1
grep -n "passed\|failed\|error" gpurun_out/r02j_pytest.log | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err
cut -c1-200 gpurun_out/r02j_bench_n1.json
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_$c.json 2> gpurun_out/r02j_bench_$c.err
  cut -c1-200 gpurun_out/r02j_bench_$c.json
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j_bench_ref.json 2> gpurun_out/r02j_bench_ref.err
cut -c1-300 gpurun_out/r02j_bench_ref.json
export MHMR_PROF_BATCH=8
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02j_launches.csv python tools/prof_forward.py > gpurun_out/r02j_prof_launch.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"attn_fwd|gemm_tc2|split_rowstats|im2col" -c 10 \
   -o gpurun_out/r02j_full_vit python tools/prof_forward.py > gpurun_out/r02j_prof_full_vit.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"layernorm|refine|smplx_vertex|hph_cross|person_gather|rowdot" -c 8 \
   -o gpurun_out/r02j_full_head python tools/prof_forward.py > gpurun_out/r02j_prof_full_head.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"skinny" -c 2 \
   -o gpurun_out/r02j_full_skinny python tools/prof_forward.py > gpurun_out/r02j_prof_full_skinny.log 2>&1
ls -la gpurun_out/r02j_full_*.ncu-rep
