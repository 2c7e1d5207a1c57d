"""The inference glue the reference keeps in demo.py:27-126, with identical signatures, so that callers
(`demo.py:334`, `app.py:132`, `train.py:356`) can switch to this package for the `Model.forward` path:

    open_image, get_camera_parameters, load_model, forward_model

Only the forward itself is B200-native; image decoding stays on the host (PIL), as in the reference.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .model import Model

from .preprocess import IMG_NORM_MEAN, IMG_NORM_STD, device_table, normalize_rgb, normalize_rgb_table  # noqa: F401

SMPLX_DIR = "models"                         # utils/constants.py:7
CACHE_DIR_MULTIHMR = "models/multiHMR"       # utils/constants.py:9


def normalize_rgb_device(img_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [B,H,W,3] on a CUDA device -> normalised fp32 [B,3,H,W], bit-identical to `normalize_rgb` on the
    host (hand-written kernel, `mhmr_op_normalize_u8`); 4x less to upload than the fp32 image."""
    from . import ops

    return ops.normalize_u8(img_u8.contiguous(), device_table(img_u8.device))


def open_image(img_path, img_size, device=torch.device("cuda"), fused=False):
    """Open, resize keeping the aspect ratio, zero-pad to a square, normalise (demo.py:27-51).  On a CUDA
    device the padded uint8 image is uploaded and normalised there (same values, a quarter of the bytes).
    `fused=True` returns the uint8 [1,S,S,3] device tensor itself: `Model.forward` / `forward_model` then run the
    fused loader (uint8 -> normalised fp16 patch rows, `mhmr_forward_u8`) and the fp32 image never exists."""
    from PIL import Image, ImageOps

    img_pil = Image.open(img_path).convert("RGB")
    img_pil_full = img_pil.copy()
    img_pil = ImageOps.contain(img_pil, (img_size, img_size))
    img_pil = ImageOps.pad(img_pil, size=(img_size, img_size))
    device = torch.device(device)
    if device.type == "cuda" and img_size % 4 == 0:
        u8 = torch.from_numpy(np.ascontiguousarray(np.asarray(img_pil))).unsqueeze(0).to(device)
        if fused:
            return u8, img_pil_full
        return normalize_rgb_device(u8), img_pil_full
    x = torch.from_numpy(normalize_rgb(np.asarray(img_pil))).unsqueeze(0).to(device)
    return x, img_pil_full


def get_focalLength_from_fieldOfView(fov=60, img_size=512):
    """utils/camera.py:50-60."""
    return img_size / (2 * np.tan(np.radians(fov) / 2))


def get_camera_parameters(img_size, fov=60, p_x=None, p_y=None, device=torch.device("cuda")):
    """K [1,3,3] from image size, field of view and principal point (demo.py:53-68)."""
    K = torch.eye(3)
    focal = get_focalLength_from_fieldOfView(fov=fov, img_size=img_size)
    K[0, 0], K[1, 1] = focal, focal
    if p_x is not None and p_y is not None:
        K[0, -1], K[1, -1] = p_x * img_size, p_y * img_size
    else:
        K[0, -1], K[1, -1] = img_size // 2, img_size // 2
    return K.unsqueeze(0).to(device)


def body_model_from_smplx_npz(path: str, num_betas: int = 10) -> dict:
    """Reads `SMPLX_NEUTRAL.npz` the way `smplx.create(..., use_pca=False, flat_hand_mean=True)` does
    (blocks/smpl_layer.py:38): shapedirs[..., :num_betas] + 10 expression directions (stored after the 300
    shape components), posedirs reshaped to [486, 3V], kinematic parents from kintree_table[0]."""
    d = np.load(path, allow_pickle=True)
    t = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a)).to(dt)
    sdirs = t(d["shapedirs"])
    V = sdirs.shape[0]
    n_shape = 300 if sdirs.shape[-1] >= 310 else sdirs.shape[-1] - 10
    parents = t(d["kintree_table"][0].astype(np.int64), torch.int64)
    parents[0] = -1
    required = ("lmk_faces_idx", "lmk_bary_coords")
    for k in required:
        if k not in d:
            raise KeyError(f"{path} lacks {k} (static face landmarks)")
    if "extra_joints_idxs" in d:
        extra = t(d["extra_joints_idxs"], torch.int64)
    else:  # smplx.vertex_ids.vertex_ids['smplx']: nose, eyes, ears, feet, finger tips
        extra = torch.tensor([9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
                              5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022], dtype=torch.int64)
    return {
        "v_template": t(d["v_template"]), "shapedirs": sdirs[..., :num_betas],
        "shapedirs_extra": sdirs[..., num_betas:num_betas + 1], "expr_dirs": sdirs[..., n_shape:n_shape + 10],
        "posedirs": t(d["posedirs"]).reshape(-1, 486).T.contiguous(), "J_regressor": t(d["J_regressor"]),
        "parents": parents, "lbs_weights": t(d["weights"]), "faces": t(d["f"].astype(np.int64), torch.int64),
        "lmk_faces_idx": t(d["lmk_faces_idx"].astype(np.int64), torch.int64),
        "lmk_bary_coords": t(d["lmk_bary_coords"]), "extra_joints_idxs": extra, "num_verts": V,
    }


def load_model(model_name, device=torch.device("cuda"), max_batch=8, max_persons=None):
    """Open a checkpoint, build the engine from its saved arguments, load the weights (demo.py:70-106).
    No download is attempted (this build has no network): a missing file is an error."""
    ckpt_path = os.path.join(CACHE_DIR_MULTIHMR, model_name + ".pt")
    if not os.path.isfile(ckpt_path):
        raise FileNotFoundError(f"{ckpt_path} not found (place the reference checkpoint there)")
    if "anny" in ckpt_path:
        raise NotImplementedError("the Anny variant (multi_hmr_anny/) is outside this build (SURVEY.md §8f)")
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    kwargs = dict(vars(ckpt["args"]))
    kwargs["type"] = ckpt["args"].train_return_type
    kwargs["img_size"] = ckpt["args"].img_size[0]
    smplx_npz = os.path.join(SMPLX_DIR, "smplx", "SMPLX_NEUTRAL.npz")
    bm = body_model_from_smplx_npz(smplx_npz, kwargs.get("num_betas", 10))
    model = Model(max_batch=max_batch, max_persons=max_persons, body_model=bm, device=device, **kwargs)
    model.load_state_dict(ckpt["model_state_dict"], strict=False)
    return model.finalize()


class HostBatchLoader:
    """Double-buffered upload of (image batch, intrinsics) from pinned host memory on a copy stream: a serving loop
    submits batch i+1 while batch i is inside `Model.forward`, so the host->device copy (19 MB of uint8 for 8 images
    at 896x896) overlaps the previous forward instead of preceding its own.  `get()` makes the current stream wait
    for the copy and hands the device tensors to `forward_model`."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pending = None

    def submit(self, images: torch.Tensor, K: torch.Tensor):
        with torch.cuda.stream(self.stream):
            x = images.to(self.device, non_blocking=True)
            k = K.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (x, k, ev)

    @property
    def pending(self) -> bool:
        return self._pending is not None

    def get(self):
        x, k, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        x.record_stream(cur)
        k.record_stream(cur)
        return x, k


def forward_model(model, input_image, camera_parameters, det_thresh=0.3, nms_kernel_size=1):
    """One forward on an image batch and its intrinsics (demo.py:108-126).  The reference wraps the call in
    no_grad + fp16 autocast; here precision is fixed by the kernels (fp16 tensor-core operands, fp32
    accumulation / residual stream / softmax / head)."""
    return model(input_image, is_training=False, nms_kernel_size=int(nms_kernel_size), det_thresh=det_thresh,
                 K=camera_parameters)
