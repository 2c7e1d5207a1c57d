"""Preprocessing: `normalize_rgb` / `open_image` (reference utils/image.py:12-24,
demo.py:27-51).  The golden table and image come from the reference's own `normalize_rgb`
(oracle/make_golden.py::run_normalize_rgb); host restatement and device kernel must match it bit for bit."""
import os

import numpy as np
import pytest
import torch

import parity_util as pu


def _golden():
    with np.load(os.path.join(pu.GOLDEN_DIR, "normalize_rgb.npz")) as f:
        return {k: f[k] for k in f.files}


def test_host_normalize_rgb_matches_reference_bit_for_bit():
    from multihmr_b200 import api

    g = _golden()
    assert np.array_equal(api.normalize_rgb(g["image"]), g["normalized"])
    assert np.array_equal(api.normalize_rgb_table(), g["table"])
    assert api.normalize_rgb(g["image"]).dtype == np.float32
    raw = api.normalize_rgb(g["image"], imagenet_normalization=False)
    assert np.array_equal(raw, np.transpose(g["image"].astype(np.float32) / 255.0, (2, 0, 1)))


def test_open_image_host_path(tmp_path):
    from PIL import Image

    from multihmr_b200 import api

    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(60, 100, 3), dtype=np.uint8)   # landscape: padded top and bottom
    path = os.path.join(tmp_path, "img.png")
    Image.fromarray(img).save(path)
    x, full = api.open_image(path, 56, device=torch.device("cpu"))
    assert tuple(x.shape) == (1, 3, 56, 56) and x.dtype == torch.float32 and full.size == (100, 60)
    black = torch.tensor(api.normalize_rgb(np.zeros((1, 1, 3), np.uint8))).reshape(3)
    assert torch.equal(x[0, :, 0, 0], black) and torch.equal(x[0, :, -1, -1], black)  # zero padding, then normalised
    assert not torch.equal(x[0, :, 28, 28], black)


@pytest.mark.gpu
def test_device_normalize_matches_reference_bit_for_bit(cuda_device):
    from multihmr_b200 import api, ops

    g = _golden()
    lut = torch.from_numpy(np.ascontiguousarray(g["table"])).to(cuda_device)  # (npz keeps the Fortran order)
    gen = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (3, 56, 64, 3), generator=gen, dtype=torch.uint8)
    img[0, 0, :, 0] = torch.arange(64, dtype=torch.uint8) * 4          # every region of the table
    img[1, :, :, 1] = 255
    out = ops.normalize_u8(img.to(cuda_device), lut).cpu()
    ref = torch.stack([torch.from_numpy(api.normalize_rgb(im.numpy())) for im in img])
    assert torch.equal(out, ref)
    assert torch.equal(api.normalize_rgb_device(img.to(cuda_device)).cpu(), ref)
    with pytest.raises(AssertionError):  # W % 4 != 0 is rejected, never silently mishandled
        ops.normalize_u8(torch.zeros(1, 8, 6, 3, dtype=torch.uint8, device=cuda_device), lut)


@pytest.mark.gpu
def test_open_image_device_path_equals_host_path(cuda_device, tmp_path):
    from PIL import Image

    from multihmr_b200 import api

    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(200, 120, 3), dtype=np.uint8)   # portrait: padded left and right
    path = os.path.join(tmp_path, "img.png")
    Image.fromarray(img).save(path)
    x_dev, _ = api.open_image(path, 224, device=cuda_device)
    x_host, _ = api.open_image(path, 224, device=torch.device("cpu"))
    assert x_dev.is_cuda and torch.equal(x_dev.cpu(), x_host)


@pytest.mark.gpu
def test_fused_uint8_loader_equals_normalize_then_forward(cuda_device):
    """Model.forward on the uint8 HWC image (fused loader: uint8 -> normalised fp16 patch rows in
    one kernel, mhmr_forward_u8) gives bit-identical outputs to Model.forward on normalize_rgb(image) — with and
    without the fp32 refinement of the detected tokens (which re-reads the pixels of its patches)."""
    import parity_util as pu
    from multihmr_b200 import api, synth

    case, sd, bm, _, K, idx = pu.build_inputs("s_224_S_forced")
    u8 = synth.make_images_u8(case["batch"], case["img_size"], seed=3)
    x32 = torch.from_numpy(np.stack([api.normalize_rgb(im.numpy()) for im in u8]))
    for refine in (True, False):
        m = pu.build_engine(case, sd, bm, refine_central=refine)
        a = {k: v.clone() for k, v in m(x32, idx=idx, K=K, is_training=True).items()}
        b = m(u8.to(cuda_device), idx=idx, K=K, is_training=True)
        for k in ("scores", "v3d", "rotmat", "shape", "dist", "loc", "j2d", "offset"):
            assert torch.equal(a[k], b[k]), (refine, k)


@pytest.mark.gpu
def test_host_batch_loader_hands_over_the_submitted_batch(cuda_device):
    """api.HostBatchLoader: the batch submitted on the copy stream arrives intact on the compute stream, two batches
    in flight never alias."""
    from multihmr_b200 import api, synth

    loader = api.HostBatchLoader(cuda_device)
    a = synth.make_images_u8(2, 224, seed=1).pin_memory()
    b = synth.make_images_u8(2, 224, seed=2).pin_memory()
    K = synth.make_cameras(2, 224, seed=1).pin_memory()
    loader.submit(a, K)
    xa, Ka = loader.get()
    loader.submit(b, K)
    assert not loader._pending is None and loader.pending
    xb, _ = loader.get()
    torch.cuda.synchronize()
    assert torch.equal(xa.cpu(), a) and torch.equal(xb.cpu(), b) and torch.equal(Ka.cpu(), K)
    assert xa.data_ptr() != xb.data_ptr()
