"""Host statement of the reference's image normalisation (utils/image.py:8-24) and the [3][256] table that the
device kernels (`mhmr_op_normalize_u8`, the fused loader of `mhmr_forward_u8`) use to reproduce it bit for bit."""
from __future__ import annotations

import numpy as np

IMG_NORM_MEAN = [0.485, 0.456, 0.406]       # utils/image.py:8
IMG_NORM_STD = [0.229, 0.224, 0.225]        # utils/image.py:9


def normalize_rgb(img: np.ndarray, imagenet_normalization: bool = True) -> np.ndarray:
    """uint8 HWC -> float32 CHW (utils/image.py:12-24)."""
    img = img.astype(np.float32) / 255.0
    img = np.transpose(img, (2, 0, 1))
    if imagenet_normalization:
        img = (img - np.asarray(IMG_NORM_MEAN).reshape(3, 1, 1)) / np.asarray(IMG_NORM_STD).reshape(3, 1, 1)
    return img.astype(np.float32)


def normalize_rgb_table() -> np.ndarray:
    """[3, 256] fp32: `normalize_rgb` of every uint8 value in every channel (the table of the device kernels)."""
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)
    return np.ascontiguousarray(normalize_rgb(ramp)[:, 0, :])


_LUT_CACHE = {}


def device_table(device):
    """The table as a cached device tensor."""
    import torch

    key = (device.type, device.index)
    if key not in _LUT_CACHE:
        _LUT_CACHE[key] = torch.from_numpy(normalize_rgb_table()).to(device)
    return _LUT_CACHE[key]
