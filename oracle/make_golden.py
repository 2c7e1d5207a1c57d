"""Pins the oracle against the reference's OWN Python and writes the golden fixtures.

Runs ONLY in the build container (needs /root/reference).  It imports /root/reference/model.py UNMODIFIED;
the three third-party packages that are neither vendored nor installed (roma, smplx, the torch.hub
DINOv2 model) plus the render-only imports (pyrender, trimesh) are shimmed through sys.modules with the
restatements of this package.  For each case it
  1. builds `Model(**ckpt_args)`, loads the synthetic state_dict (strict=False, like demo.py:103),
  2. runs `model(x, K=K, ...)` on seeded synthetic inputs,
  3. runs oracle.multihmr_ref.model_forward on the same inputs and asserts they agree (<= 2e-5 abs),
  4. stores the reference outputs under tests/golden/<case>.npz (small tensors only; inputs and weights
     are regenerated from the seeds by multihmr_b200.synth).

Usage:  python -m oracle.make_golden            (from the repo root)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REFERENCE = "/root/reference"

from multihmr_b200 import synth  # noqa: E402
from oracle import dinov2_ref, multihmr_ref, roma_ref, smplx_ref  # noqa: E402

CASES = {
    # name: (backbone, img_size, batch, persons per image (forced idx) or None for natural detection)
    "c1_672_S_forced": dict(backbone="dinov2_vits14", img_size=672, batch=1, persons=[1], seed=0),
    "s_224_S_forced": dict(backbone="dinov2_vits14", img_size=224, batch=3, persons=[2, 0, 3], seed=1),
    "s_224_S_detect": dict(backbone="dinov2_vits14", img_size=224, batch=2, persons=None, seed=2,
                           det_bias=-2.1),
    "s_448_B_forced": dict(backbone="dinov2_vitb14", img_size=448, batch=2, persons=[3, 1], seed=3,
                           jitter=True),
    "s_224_S_asymK": dict(backbone="dinov2_vits14", img_size=224, batch=2, persons=[2, 2], seed=4, jitter=True,
                          asymmetric=True),
    # the ViT-L architecture (depth 24, 16 heads, D = 1024) of the headline configs, at a size the CPU reference
    # runs in seconds; rectangular per-image intrinsics
    "s_280_L_forced": dict(backbone="dinov2_vitl14", img_size=280, batch=2, persons=[2, 1], seed=5, jitter=True),
    # DINOv2-like massive-activation channels (synth.add_outlier_channels): residual values of O(100)
    "s_224_S_outliers": dict(backbone="dinov2_vits14", img_size=224, batch=2, persons=[2, 1], seed=6, jitter=True,
                             outliers=True),
}

_CURRENT_BM = {}


def install_shims():
    roma = types.ModuleType("roma")
    roma.special_gramschmidt = roma_ref.special_gramschmidt
    roma.rotvec_to_rotmat = roma_ref.rotvec_to_rotmat
    roma.rotmat_to_rotvec = roma_ref.rotmat_to_rotvec
    sys.modules["roma"] = roma

    smplx = types.ModuleType("smplx")

    def create(model_path, model_type="smplx", gender="neutral", use_pca=False, flat_hand_mean=True,
               num_betas=10, **kw):
        assert model_type == "smplx" and not use_pca and flat_hand_mean
        return smplx_ref.SMPLXShim(_CURRENT_BM["bm"], num_betas=num_betas)

    smplx.create = create
    jn = types.ModuleType("smplx.joint_names")
    jn.JOINT_NAMES = smplx_ref.JOINT_NAMES
    smplx.joint_names = jn
    sys.modules["smplx"] = smplx
    sys.modules["smplx.joint_names"] = jn
    for name in ("pyrender", "trimesh"):
        sys.modules[name] = types.ModuleType(name)
    torch.hub.load = lambda repo, name, pretrained=False, **kw: dinov2_ref.HubModelShim(name)


def build_reference_model(case, sd, bm, mean):
    _CURRENT_BM["bm"] = bm
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="mhmr_ref_")
    os.makedirs(os.path.join(tmp, "models"))
    np.savez(os.path.join(tmp, "models", "smpl_mean_params.npz"), pose=mean["pose"].numpy(),
             shape=mean["shape"].numpy(), cam=mean["cam"].numpy())
    os.chdir(tmp)  # MEAN_PARAMS is a relative path (utils/constants.py:8)
    try:
        if REFERENCE not in sys.path:
            sys.path.insert(0, REFERENCE)
        from model import Model  # the reference, unmodified

        model = Model(backbone=case["backbone"], img_size=case["img_size"], num_betas=10)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(k.startswith("smpl_layer.") for k in missing), [k for k in missing if not k.startswith("smpl_layer.")][:5]
        return model.eval()
    finally:
        os.chdir(cwd)


def flatten_persons(persons):
    keys = ("scores", "loc", "transl", "transl_pelvis", "rotvec", "expression", "shape", "v3d", "j3d", "j2d")
    return {k: torch.stack([p[k] for p in persons]) for k in keys}


def run_case(name, case, out_dir):
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    seed = case["seed"]
    sd = synth.make_state_dict(case["backbone"], case["img_size"], seed=seed, det_bias=case.get("det_bias", -4.0))
    if case.get("outliers"):
        synth.add_outlier_channels(sd, case["backbone"], seed)
    bm = synth.make_body_model(seed)
    mean = synth.make_mean_params(seed)
    x = synth.make_images(case["batch"], case["img_size"], seed)
    K = synth.make_cameras(case["batch"], case["img_size"], jitter=case.get("jitter", False), seed=seed,
                           asymmetric=case.get("asymmetric", False))
    res = case["img_size"] // 14
    cfg = multihmr_ref.RefConfig(backbone=case["backbone"], img_size=case["img_size"])
    body = smplx_ref.SMPLXShim(bm, 10)
    model = build_reference_model(case, sd, bm, mean)
    gold = {}
    with torch.no_grad():
        if case["persons"] is not None:
            idx = synth.make_forced_idx(case["batch"], res, case["persons"], seed)
            ref = model(x, idx=idx, K=K, is_training=True)
            mine = multihmr_ref.model_forward(sd, body, cfg, x, K, idx=idx, is_training=True)
            keys = ["scores", "offset", "dist", "expression", "rotmat", "shape", "rotvec", "loc", "v3d", "j3d",
                    "j2d", "v2d", "transl", "transl_pelvis", "dist_postprocessed"]
            for k in keys:
                gold[k] = ref[k]
            gold["idx"] = torch.stack(idx)
        else:
            persons = model(x, K=K, det_thresh=0.3, nms_kernel_size=3, is_training=False)
            assert len(persons) > 0, "no natural detections: adjust det_bias"
            ref = flatten_persons(persons)
            mine = flatten_persons(multihmr_ref.model_forward(sd, body, cfg, x, K, det_thresh=0.3,
                                                              nms_kernel_size=3))
            keys = list(ref.keys())
            for k in keys:
                gold[k] = ref[k]
    worst = 0.0
    for k in keys:
        assert ref[k].shape == mine[k].shape, (k, ref[k].shape, mine[k].shape)
        err = (ref[k].float() - mine[k].float()).abs().max().item()
        worst = max(worst, err)
        if os.environ.get("MHMR_VERBOSE"):
            print(f"   {k:20s} max|ref|={ref[k].abs().max().item():10.4f} err={err:.2e}")
        assert err <= 2e-5 * max(1.0, ref[k].abs().max().item()), (name, k, err)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **{k: v.numpy() for k, v in gold.items()})
    n_p = gold[keys[-1]].shape[0]
    print(f"{name}: reference == oracle (max abs err {worst:.2e}), persons={n_p}, keys={len(keys)}")


def run_normalize_rgb(out_dir):
    """Golden of the preprocessing step: the reference's own `normalize_rgb`
    (utils/image.py:12-24) applied to every uint8 value in every channel -> a [3, 256] fp32 table, plus its output
    on a seeded random image.  The host restatement (multihmr_b200.api.normalize_rgb) and the device kernel must
    reproduce it bit for bit."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_ref_utils_image", os.path.join(REFERENCE, "utils", "image.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)       # [1, 256, 3]
    table = mod.normalize_rgb(ramp)[:, 0, :]                                          # [3, 256]
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    out = mod.normalize_rgb(img)
    from multihmr_b200 import api

    assert np.array_equal(api.normalize_rgb(img), out), "host restatement of normalize_rgb differs from the reference"
    np.savez_compressed(os.path.join(out_dir, "normalize_rgb.npz"), table=np.ascontiguousarray(table), image=img,
                        normalized=np.ascontiguousarray(out))
    print(f"normalize_rgb: table {table.shape} {table.dtype}, restatement bit-exact")


def make_matching_cases():
    """Seeded 2-D keypoint sets (pixels) for the matching golden: persons are blobs of 127 joints; predictions are
    noisy copies of some ground truths plus spurious far-away ones (false positives) and missing ones (misses)."""
    rng = np.random.default_rng(123)
    cases = []
    for (G, keep, extra) in [(1, 1, 0), (3, 3, 0), (4, 2, 1), (2, 2, 3), (5, 4, 2), (3, 0, 2), (2, 0, 0), (6, 6, 1)]:
        centers = rng.uniform(100, 800, size=(G, 1, 2))
        gt = centers + rng.normal(0, 40, size=(G, 127, 2))
        order = rng.permutation(G)[:keep]
        preds = [gt[g] + rng.normal(0, 6, size=(127, 2)) for g in order]
        for _ in range(extra):
            c = rng.uniform(2000, 3000, size=(1, 2)) if rng.random() < 0.5 else centers[rng.integers(G)] + rng.normal(0, 60, size=(1, 2))
            preds.append(c + rng.normal(0, 40, size=(127, 2)))
        perm = rng.permutation(len(preds)) if preds else []
        pred = np.stack([preds[i] for i in perm]) if len(preds) else np.zeros((0, 127, 2))
        cases.append((pred.astype(np.float32), gt.astype(np.float32)))
    return cases


def run_eval_matching(out_dir):
    """Golden of the evaluation helpers: the reference's OWN match_2d_greedy / get_bbx_overlap /
    compute_prf1 (utils/training.py, loaded as a file: it only needs numpy + torch) on seeded keypoint sets; the
    restatement oracle/eval_ref.py must agree exactly."""
    import importlib.util

    from oracle import eval_ref

    spec = importlib.util.spec_from_file_location("_ref_utils_training", os.path.join(REFERENCE, "utils", "training.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = {}
    for i, (pred, gt) in enumerate(make_matching_cases()):
        vm = np.ones_like(gt[..., 0]).astype(np.bool_)
        if len(pred):
            best, fps, misses = mod.match_2d_greedy(pred, gt, vm)
        else:  # train.py:363 builds an empty array when nothing is detected: every ground truth is a miss
            best, fps, misses = np.zeros((0, 2), dtype=np.int64), [], list(range(len(gt)))
        best = np.asarray(best, dtype=np.int64).reshape(-1, 2)
        mine = eval_ref.match_2d_greedy(pred, gt, vm)
        assert np.array_equal(mine[0], best) and list(mine[1]) == [int(v) for v in fps] and list(mine[2]) == [int(v) for v in misses], i
        ious = np.array([[mod.get_bbx_overlap(p, g) for g in gt] for p in pred]).reshape(len(pred), len(gt))
        mine_iou = np.array([[eval_ref.get_bbx_overlap(p, g) for g in gt] for p in pred]).reshape(len(pred), len(gt))
        assert np.allclose(ious, mine_iou, rtol=0, atol=0)
        gold[f"pred{i}"], gold[f"gt{i}"], gold[f"best{i}"] = pred, gt, best
        gold[f"fp{i}"], gold[f"miss{i}"] = np.asarray(fps, dtype=np.int64), np.asarray(misses, dtype=np.int64)
        gold[f"iou{i}"] = ious.astype(np.float64)
    prf = []
    for (c, m, f) in [(0, 0, 0), (10, 2, 1), (7, 7, 3), (25, 0, 0), (13, 5, 9)]:
        r = mod.compute_prf1(c, m, f)
        assert tuple(r) == tuple(eval_ref.compute_prf1(c, m, f))
        prf.append([c, m, f, *r])
    gold["prf1"] = np.asarray(prf, dtype=np.float64)
    gold["n_cases"] = np.asarray(len(make_matching_cases()))
    np.savez_compressed(os.path.join(out_dir, "eval_matching.npz"), **gold)
    print(f"eval_matching: {int(gold['n_cases'])} cases, restatement == reference (matches, false positives, misses, IoU, PRF1)")


def main():
    assert os.path.isdir(REFERENCE), "make_golden needs the reference checkout (build container only)"
    install_shims()
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    if not only or "normalize_rgb" in only:
        run_normalize_rgb(out_dir)
    if not only or "eval_matching" in only:
        run_eval_matching(out_dir)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        run_case(name, case, out_dir)


if __name__ == "__main__":
    main()
