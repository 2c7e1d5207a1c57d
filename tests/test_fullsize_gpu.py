"""Full BASELINE.json sizes on the GPU: the engine vs the oracle restatement evaluated in fp32 ON THE SAME
GPU (plain torch ops, TF32 off; the oracle is the checker, not the product).  c3-like: multiHMR_896_L,
c5: multiHMR_1288_L with 20 forced persons per image (stress of the cross-attention decoder + batched LBS)."""
import pytest
import torch

import parity_util as pu

pytestmark = pytest.mark.gpu


def _oracle_on_gpu(sd, bm, cfg_kw, x, K, idx, dev):
    from oracle import multihmr_ref, smplx_ref

    sd_d = {k: v.to(dev) for k, v in sd.items()}
    body = smplx_ref.SMPLXShim(bm, 10).to(dev)
    cfg = multihmr_ref.RefConfig(**cfg_kw)
    with torch.no_grad():
        out = multihmr_ref.model_forward(sd_d, body, cfg, x.to(dev), K.to(dev), idx=tuple(i.to(dev) for i in idx),
                                         is_training=True)
    return {k: v.float().cpu() for k, v in out.items()}


# BASELINE.json configs at their stated sizes: c2 = multiHMR_672_L batch 4, c3 = multiHMR_896_L batch 8,
# c5 = multiHMR_1288_L batch 2 with 20 persons per image.  One tolerance for all: 1e-3 max-abs on scores,
# SMPL-X parameters and 3-D vertices (north_star), PVE < 1 mm.
@pytest.mark.parametrize("name,backbone,S,B,persons", [
    ("c2_672_L_b4", "dinov2_vitl14", 672, 4, [2, 1, 3, 2]),
    ("c3_896_L_b8", "dinov2_vitl14", 896, 8, [2, 3, 1, 2, 0, 1, 2, 3]),
    ("c5_1288_L_b2_20ppl", "dinov2_vitl14", 1288, 2, [20, 20]),
])
def test_engine_vs_oracle_full_size(cuda_device, name, backbone, S, B, persons):
    from multihmr_b200 import synth

    seed = 21
    sd = synth.make_state_dict(backbone, S, seed=seed)
    bm = synth.make_body_model(seed)
    x, K = synth.make_images(B, S, seed), synth.make_cameras(B, S, jitter=True, seed=seed)
    idx = synth.make_forced_idx(B, S // 14, persons, seed)
    ref = _oracle_on_gpu(sd, bm, dict(backbone=backbone, img_size=S), x, K, idx, cuda_device)
    torch.cuda.empty_cache()
    m = pu.build_engine(dict(backbone=backbone, img_size=S, batch=B), sd, bm, max_persons=64)
    out = m(x, idx=idx, K=K, is_training=True)
    keys = ["scores", "offset", "dist", "expression", "rotmat", "shape", "rotvec", "loc", "v3d", "j3d", "j2d", "v2d",
            "transl", "transl_pelvis", "dist_postprocessed"]
    bad = pu.compare(out, ref, keys, focal=float(K[:, 0, 0].max()), verbose=True)
    # PVE (train.py:387): mean per-vertex error in mm
    pve = (out["v3d"].cpu() - ref["v3d"]).norm(dim=-1).mean().item() * 1000
    print(f"{name}: PVE vs oracle = {pve:.4f} mm over {sum(persons)} persons")
    assert pve < 1.0
    assert not bad, bad
