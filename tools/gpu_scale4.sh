#!/bin/bash
# scaling point at N = 4 (gpurun --gpus 4): torchrun, one rank per GPU, image shards + one all-gather of record blocks
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_scale_n4.json 2> gpurun_out/r02k_scale_n4.err
cat gpurun_out/r02k_scale_n4.json | cut -c1-400
tail -3 gpurun_out/r02k_scale_n4.err
