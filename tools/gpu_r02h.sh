#!/bin/bash
# why do fc1 / proj slow down with the folded LayerNorm?  isolated per-launch times + full sections, both modes
mkdir -p gpurun_out
export MHMR_PROF_BATCH=8
for f in 1 0; do
  MHMR_LN_FOLD=$f timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/r02h_launches_fold$f.csv python tools/prof_forward.py > gpurun_out/r02h_prof_fold$f.log 2>&1
done
MHMR_LN_FOLD=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"gemm_tc2_kernel<[678]>" -s 8 -c 8 -o gpurun_out/r02h_full_fold1 python tools/prof_forward.py > gpurun_out/r02h_full1.log 2>&1
MHMR_LN_FOLD=0 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"gemm_tc2_kernel<[013]>" -s 8 -c 8 -o gpurun_out/r02h_full_fold0 python tools/prof_forward.py > gpurun_out/r02h_full0.log 2>&1
ls -la gpurun_out/r02h_*
