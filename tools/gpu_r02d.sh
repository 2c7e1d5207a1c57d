#!/bin/bash
# multi-GPU run (gpurun --gpus N): 2-rank NCCL parity test of the sharded path, then the scaling bench at N = 2..NGPU
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_parallel_gpu.py -q -s > gpurun_out/r02j_parallel_test.log 2>&1
tail -3 gpurun_out/r02j_parallel_test.log
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02j_scale_n1.json 2> gpurun_out/r02j_scale_n1.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_scale_n$n.json 2> gpurun_out/r02j_scale_n$n.err
    fi
    cat gpurun_out/r02j_scale_n$n.json
  fi
done
