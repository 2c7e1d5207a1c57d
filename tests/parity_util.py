"""Shared helpers of the parity tests: rebuild the seeded synthetic case, load the golden fixture
(outputs of the UNMODIFIED reference, written by oracle/make_golden.py), run the engine."""
import math
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Must stay in sync with oracle/make_golden.py::CASES (the fixture generator).
CASES = {
    "c1_672_S_forced": dict(backbone="dinov2_vits14", img_size=672, batch=1, persons=[1], seed=0),
    "s_224_S_forced": dict(backbone="dinov2_vits14", img_size=224, batch=3, persons=[2, 0, 3], seed=1),
    "s_224_S_detect": dict(backbone="dinov2_vits14", img_size=224, batch=2, persons=None, seed=2, det_bias=-2.1),
    "s_448_B_forced": dict(backbone="dinov2_vitb14", img_size=448, batch=2, persons=[3, 1], seed=3, jitter=True),
    "s_224_S_asymK": dict(backbone="dinov2_vits14", img_size=224, batch=2, persons=[2, 2], seed=4, jitter=True,
                          asymmetric=True),
    # ViT-L depth (24 layers) with the synthetic body model's large extent (|v3d| up to 4.3 m): the bulk fp16
    # pass alone gives max |dv3d| = 1.07e-3 here; with the fp32 refinement of the detected tokens' residual
    # streams it is ~2e-4, inside the 1e-3 contract like every other case
    "s_280_L_forced": dict(backbone="dinov2_vitl14", img_size=280, batch=2, persons=[2, 1], seed=5, jitter=True),
    # DINOv2-like massive-activation channels (synth.add_outlier_channels): residual values of O(100)
    "s_224_S_outliers": dict(backbone="dinov2_vits14", img_size=224, batch=2, persons=[2, 1], seed=6, jitter=True,
                             outliers=True),
}

# Absolute tolerances vs the fp32 reference (BASELINE.json north_star: 1e-3 abs on scores / SMPL-X
# parameters / 3-D vertices).  Pixel-space quantities scale with the image size: 1e-3 of a 14-px patch
# step per metre-level 1e-3 is not meaningful, so they get 2e-2 px (about 1e-5 of the image width x 2).
TOL = {
    "scores": 1e-3, "offset": 1e-3, "dist": 1e-3, "dist_postprocessed": 1e-3, "expression": 1e-3,
    "rotmat": 1e-3, "shape": 1e-3, "rotvec": 1e-3, "v3d": 1e-3, "j3d": 1e-3, "transl": 1e-3,
    "transl_pelvis": 1e-3, "loc": 2e-2, "j2d": None, "v2d": None,
}


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def build_inputs(name):
    from multihmr_b200 import synth

    case = CASES[name]
    seed = case["seed"]
    sd = synth.make_state_dict(case["backbone"], case["img_size"], seed=seed, det_bias=case.get("det_bias", -4.0))
    if case.get("outliers"):
        synth.add_outlier_channels(sd, case["backbone"], seed)
    bm = synth.make_body_model(seed)
    x = synth.make_images(case["batch"], case["img_size"], seed)
    K = synth.make_cameras(case["batch"], case["img_size"], jitter=case.get("jitter", False), seed=seed,
                           asymmetric=case.get("asymmetric", False))
    idx = None
    if case["persons"] is not None:
        idx = synth.make_forced_idx(case["batch"], case["img_size"] // 14, case["persons"], seed)
    return case, sd, bm, x, K, idx


def build_engine(case, sd, bm, max_batch=None, max_persons=64, **kw):
    from multihmr_b200.model import Model

    m = Model(backbone=case["backbone"], img_size=case["img_size"], num_betas=10,
              max_batch=max_batch or case["batch"], max_persons=max_persons, body_model=bm, **kw)
    m.load_state_dict(sd, strict=False)
    return m


def projection_tolerance(ref_3d, K_focal, tol3d=1e-3):
    """Pixel tolerance for K.(p/p_z): |d(u)| <= f/z * (|dx| + |u-c|/f*|dz|) ~ 2 f tol / z_min."""
    zmin = ref_3d[..., 2].abs().min().clamp_min(1e-3).item()
    return 3.0 * K_focal * tol3d / zmin + 1e-2


def _rotvec_to_rotmat(rv):
    from oracle import roma_ref

    return roma_ref.rotvec_to_rotmat(rv)


def compare(got: dict, ref: dict, keys, focal=None, verbose=False):
    """Returns list of (key, err, tol) that fail; prints a table when verbose.  Tolerances are the fixed
    contract of TOL (1e-3 abs on scores / SMPL-X parameters / 3-D outputs): there is no per-case scaling."""
    bad = []
    for k in keys:
        g, r = got[k].detach().float().cpu(), ref[k].float()
        assert g.shape == r.shape, (k, tuple(g.shape), tuple(r.shape))
        if k == "rotvec":
            # axis-angle is discontinuous at angle = pi (r and -r are the same rotation): compare the
            # vectors away from pi and the rotations they encode everywhere.
            near_pi = r.norm(dim=-1) > math.pi - 0.05
            err_vec = ((g - r).abs().amax(dim=-1) * (~near_pi)).max().item() if r.numel() else 0.0
            err_rot = (_rotvec_to_rotmat(g) - _rotvec_to_rotmat(r)).abs().max().item() if r.numel() else 0.0
            err = max(err_vec, err_rot)
        else:
            err = (g - r).abs().max().item() if r.numel() else 0.0
        tol = TOL.get(k, 1e-3)
        if tol is None:
            src = ref["j3d"] if k == "j2d" else ref["v3d"]
            tol = projection_tolerance(src, focal, tol3d=1e-3)
        if verbose:
            print(f"  {k:20s} max|ref|={r.abs().max().item() if r.numel() else 0:10.4f} err={err:.3e} tol={tol:.1e}"
                  f" {'OK' if err <= tol else 'FAIL'}")
        if not err <= tol:
            bad.append((k, err, tol))
    return bad
