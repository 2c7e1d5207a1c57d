"""One profiled forward of the bench workload (multiHMR_896_L, batch 8) for ncu:
   ncu --profile-from-start off ... python tools/prof_forward.py
Only the region between cudaProfilerStart/Stop (one full forward after a warm-up forward) is captured."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multihmr_b200 import synth  # noqa: E402
from multihmr_b200.model import Model  # noqa: E402

w = bench.WORKLOAD
B, S = int(os.environ.get("MHMR_PROF_BATCH", w["batch_per_gpu"])), w["img_size"]
dev = torch.device("cuda:0")
sd, bm = bench.build_workload(det_bias=0.0)
x = synth.make_images_u8(B, S, seed=w["seed"]).to(dev)   # uint8 HWC: the fused loader, as in bench.py
K = synth.make_cameras(B, S, seed=w["seed"]).to(dev)
idx = synth.make_forced_idx(B, S // 14, w["target_persons_per_image"], seed=w["seed"])
m = Model(backbone=w["backbone"], img_size=S, max_batch=B, max_persons=128, body_model=bm, device=dev)
m.load_state_dict(sd)
m.finalize()
m(x, idx=idx, K=K, is_training=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m(x, idx=idx, K=K, is_training=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per forward:", m.last_launch_count())
