"""GPU parity proper: the engine (sm_100a kernels behind the C-ABI) vs the golden fixtures = outputs of
the unmodified reference `Model.forward` on identical seeded inputs.  Tolerance: 1e-3 abs on detection
scores, SMPL-X parameters and 3-D vertices (BASELINE.json north_star); see parity_util.TOL."""
import pytest
import torch

import parity_util as pu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["s_224_S_forced", "s_448_B_forced", "c1_672_S_forced"])
def test_forced_idx_matches_reference(cuda_device, name):
    case, sd, bm, x, K, idx = pu.build_inputs(name)
    gold = pu.load_golden(name)
    m = pu.build_engine(case, sd, bm)
    out = m(x, idx=idx, K=K, is_training=True)
    bad = pu.compare(out, gold, [k for k in gold if k != "idx"], focal=float(K[:, 0, 0].max()), verbose=True)
    assert not bad, bad


def test_natural_detection_matches_reference(cuda_device):
    name = "s_224_S_detect"
    case, sd, bm, x, K, _ = pu.build_inputs(name)
    gold = pu.load_golden(name)
    m = pu.build_engine(case, sd, bm)
    persons = m(x, K=K, det_thresh=0.3, nms_kernel_size=3)
    assert len(persons) == gold["scores"].shape[0]
    assert set(persons[0]) == {"scores", "loc", "transl", "transl_pelvis", "rotvec", "expression", "shape", "v3d",
                               "j3d", "j2d"}                                         # model.py:329-347
    got = {k: torch.stack([p[k] for p in persons]) for k in gold}
    bad = pu.compare(got, gold, list(gold), focal=float(K[:, 0, 0].max()), verbose=True)
    assert not bad, bad


def test_no_detection_returns_empty_list(cuda_device):
    case, sd, bm, x, K, _ = pu.build_inputs("s_224_S_detect")
    m = pu.build_engine(case, sd, bm)
    assert m(x, K=K, det_thresh=0.999, nms_kernel_size=3) == []                      # model.py:241-243


def test_detection_set_and_order_vs_oracle(cuda_device):
    """Many natural detections: same cells in torch.where order, except cells within 1e-3 of the threshold
    or of a 3x3 tie (score discontinuities)."""
    from multihmr_b200 import synth
    from oracle import dinov2_ref, multihmr_ref

    backbone, S, B = "dinov2_vits14", 224, 3
    sd = synth.make_state_dict(backbone, S, seed=11, det_bias=-1.0)
    bm = synth.make_body_model(11)
    x, K = synth.make_images(B, S, 11), synth.make_cameras(B, S, seed=11)
    m = pu.build_engine(dict(backbone=backbone, img_size=S, batch=B), sd, bm, max_persons=256)
    t, P = m.forward_raw(x, K, det_thresh=0.3, nms_kernel_size=3)
    with torch.no_grad():
        z = dinov2_ref.get_intermediate_layers(x, sd, backbone, "backbone.encoder.")
        s_raw, _, _ = multihmr_ref.detection(z, sd, 1, 0.3, None, False)            # no NMS: raw scores
        s_nms, _, idx = multihmr_ref.detection(z, sd, 3, 0.3, None, False)
    ref_cells = set(zip(idx[0].tolist(), idx[1].tolist(), idx[2].tolist()))
    got = t["det_idx"][:, :P].cpu()
    got_list = list(zip(got[0].tolist(), got[1].tolist(), got[2].tolist()))
    assert got_list == sorted(got_list), "persons must come in (b, y, x) order"
    s_raw = s_raw[..., 0]
    ambiguous = set()
    pooled = torch.nn.functional.max_pool2d(s_raw[:, None], 3, 1, 1)[:, 0]
    second = (pooled - s_raw).abs()
    for b, y, xx in ref_cells.symmetric_difference(got_list):
        near_thr = abs(s_raw[b, y, xx].item() - 0.3) < 1e-3
        near_tie = second[b, y, xx].item() < 1e-3
        assert near_thr or near_tie, ("detection differs away from a discontinuity", (b, y, xx))
        ambiguous.add((b, y, xx))
    assert P >= 5 and len(ambiguous) <= max(2, P // 10)
    err = (t["scores_map"].cpu() - s_nms[..., 0]).abs()
    assert (err > 1e-3).sum().item() <= 2 * len(ambiguous) + 2  # NMS flips only at ambiguous cells


def test_capacity_overflow_is_an_error_not_a_truncation(cuda_device):
    from multihmr_b200 import _lib, synth

    backbone, S = "dinov2_vits14", 224
    sd = synth.make_state_dict(backbone, S, seed=11, det_bias=2.0)  # nearly every NMS maximum fires
    m = pu.build_engine(dict(backbone=backbone, img_size=S, batch=1), sd, synth.make_body_model(11), max_persons=4)
    with pytest.raises(_lib.MhmrError, match="max_persons"):
        m(synth.make_images(1, S, 11), K=synth.make_cameras(1, S, seed=11), det_thresh=0.3, nms_kernel_size=3)


def test_batch_invariance_and_smaller_batches(cuda_device):
    """Images are independent units: running images one by one gives the same persons."""
    case, sd, bm, x, K, idx = pu.build_inputs("s_224_S_forced")
    m = pu.build_engine(case, sd, bm)
    full = m(x, idx=idx, K=K, is_training=True)
    b_idx = idx[0]
    for b in range(case["batch"]):
        sel = b_idx == b
        if sel.sum() == 0:
            continue
        sub = tuple(t[sel] if i else torch.zeros_like(t[sel]) for i, t in enumerate(idx))
        one = m(x[b:b + 1], idx=sub, K=K[b:b + 1], is_training=True)
        for k in ("v3d", "rotmat", "shape", "dist", "loc"):
            assert (one[k] - full[k][sel.to(full[k].device)]).abs().max().item() <= 1e-4, k
