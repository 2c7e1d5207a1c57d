#!/bin/bash
# staged GPU run: new attention kernel first (falls back to the round-1 kernel for the rest if it fails)
mkdir -p gpurun_out
export MHMR_ATTN_V1=0
timeout 900 python -m pytest tests/test_attention_gpu.py -q -x > gpurun_out/r02b_attn.log 2>&1
if [ $? -ne 0 ]; then echo "NEW ATTENTION FAILED -> MHMR_ATTN_V1" >> gpurun_out/r02b_attn.log; export MHMR_ATTN_V1=1; fi
tail -3 gpurun_out/r02b_attn.log
timeout 600 python tools/bench_ops.py --what attention --out gpurun_out/r02b_attn_new.json > gpurun_out/r02b_ops.log 2>&1
MHMR_ATTN_V1=1 timeout 600 python tools/bench_ops.py --what attention --out gpurun_out/r02b_attn_v1.json >> gpurun_out/r02b_ops.log 2>&1
for p in 1 2 3; do
  echo "POLY=$p" >> gpurun_out/r02b_ops.log
  MHMR_ATTN_POLY=$p timeout 600 python tools/bench_ops.py --what attention --out gpurun_out/r02b_attn_poly$p.json >> gpurun_out/r02b_ops.log 2>&1
done
MHMR_ATTN_POLY=2 timeout 900 python -m pytest tests/test_attention_gpu.py -q > gpurun_out/r02b_attn_poly2_test.log 2>&1
tail -2 gpurun_out/r02b_attn_poly2_test.log
MHMR_ATTN_TAIL=1 timeout 900 python -m pytest tests/test_attention_gpu.py -q > gpurun_out/r02b_attn_tail_test.log 2>&1
tail -2 gpurun_out/r02b_attn_tail_test.log
MHMR_ATTN_HELPER=1 timeout 900 python -m pytest tests/test_attention_gpu.py -q > gpurun_out/r02b_attn_helper_test.log 2>&1
tail -2 gpurun_out/r02b_attn_helper_test.log
for tl in 0 1; do
  echo "HELPER=1 TAIL=$tl" >> gpurun_out/r02b_ops.log
  MHMR_ATTN_HELPER=1 MHMR_ATTN_TAIL=$tl timeout 600 python tools/bench_ops.py --what attention --out gpurun_out/r02b_attn_helper_tail$tl.json >> gpurun_out/r02b_ops.log 2>&1
done
MHMR_ATTN_TOKEN=1 timeout 900 python -m pytest tests/test_attention_gpu.py -q > gpurun_out/r02b_attn_token_test.log 2>&1
tail -2 gpurun_out/r02b_attn_token_test.log
for tk in 1 2; do
  echo "TOKEN=$tk TAIL=1" >> gpurun_out/r02b_ops.log
  MHMR_ATTN_TOKEN=$tk MHMR_ATTN_TAIL=1 timeout 600 python tools/bench_ops.py --what attention --out gpurun_out/r02b_attn_token${tk}_tail1.json >> gpurun_out/r02b_ops.log 2>&1
done
echo "TOKEN=1 HELPER=1 TAIL=1" >> gpurun_out/r02b_ops.log
MHMR_ATTN_TOKEN=1 MHMR_ATTN_HELPER=1 MHMR_ATTN_TAIL=1 timeout 600 python tools/bench_ops.py --what attention --out gpurun_out/r02b_attn_token1_helper_tail1.json >> gpurun_out/r02b_ops.log 2>&1
for p in 0 2; do
  echo "TAIL=1 POLY=$p" >> gpurun_out/r02b_ops.log
  MHMR_ATTN_TAIL=1 MHMR_ATTN_POLY=$p timeout 600 python tools/bench_ops.py --what attention --out gpurun_out/r02b_attn_tail_poly$p.json >> gpurun_out/r02b_ops.log 2>&1
done
grep 'attention\|POLY\|TAIL\|HELPER\|TOKEN' gpurun_out/r02b_ops.log
timeout 1800 python -m pytest tests -m gpu -q -s --deselect tests/test_attention_gpu.py > gpurun_out/r02b_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r02b_pytest.log | tail -3
if grep -q "failed" gpurun_out/r02b_pytest.log; then
  MHMR_PDL=0 timeout 1800 python -m pytest tests -m gpu -q -s --deselect tests/test_attention_gpu.py > gpurun_out/r02b_pytest_nopdl.log 2>&1
  grep -n "passed\|failed" gpurun_out/r02b_pytest_nopdl.log | tail -3
fi
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
cat gpurun_out/r02b_bench.json

# ---- second half: fastest validated attention variant -> full test pass, profiles, bench legs
python tools/pick_attn_variant.py > gpurun_out/r02b_pick.txt 2>&1
cat gpurun_out/r02b_pick.txt
eval "$(grep '^export' gpurun_out/r02b_pick.txt)"
env | grep MHMR_ > gpurun_out/r02b_env.txt
timeout 1200 python -m pytest tests/test_e2e_gpu.py tests/test_fullsize_gpu.py -q > gpurun_out/r02b_pytest_best.log 2>&1
grep -n "passed\|failed" gpurun_out/r02b_pytest_best.log | tail -2
bash tools/gpu_r02c.sh
