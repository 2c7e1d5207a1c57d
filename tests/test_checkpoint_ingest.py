"""Checkpoint ingest: `api.load_model` / `api.body_model_from_smplx_npz` on files written in
the reference's own formats -- the dict `train.py:195-207` saves ({'epoch', 'iter', 'model_state_dict', 'args':
argparse.Namespace}, body-model buffers excluded) and the `SMPLX_NEUTRAL.npz` layout `smplx.create` reads
(blocks/smpl_layer.py:38).  The published checkpoints and the licensed SMPL-X file are not available offline, so
the files carry the seeded synthetic weights of the golden cases; the loaded engine must reproduce the golden
outputs of the unmodified reference."""
import argparse
import os

import numpy as np
import pytest
import torch

import parity_util as pu


def write_smplx_npz(path, bm):
    """The keys and shapes of SMPLX_NEUTRAL.npz that smplx.create(use_pca=False, flat_hand_mean=True) consumes."""
    V = bm["v_template"].shape[0]
    sdirs = np.zeros((V, 3, 400), np.float32)          # 300 shape + 100 expression components
    sdirs[..., :10] = bm["shapedirs"].numpy()
    sdirs[..., 10:11] = bm["shapedirs_extra"].numpy()
    sdirs[..., 300:310] = bm["expr_dirs"].numpy()
    kintree = np.zeros((2, 55), np.uint32)
    kintree[0] = bm["parents"].numpy().astype(np.int64).astype(np.uint32)  # root parent = 2**32 - 1, as in the file
    kintree[1] = np.arange(55)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, v_template=bm["v_template"].numpy(), shapedirs=sdirs,
             posedirs=bm["posedirs"].numpy().T.reshape(V, 3, 486).copy(), J_regressor=bm["J_regressor"].numpy(),
             kintree_table=kintree, weights=bm["lbs_weights"].numpy(), f=bm["faces"].numpy().astype(np.uint32),
             lmk_faces_idx=bm["lmk_faces_idx"].numpy().astype(np.int64), lmk_bary_coords=bm["lmk_bary_coords"].numpy(),
             extra_joints_idxs=bm["extra_joints_idxs"].numpy())


def reference_args(case):
    """The argparse.Namespace `train.py` stores in a checkpoint: model arguments + (ignored) training arguments."""
    return argparse.Namespace(
        backbone=case["backbone"], pretrained_backbone=0, img_size=[case["img_size"], case["img_size"]],
        camera_embedding="geometric", camera_embedding_num_bands=16, camera_embedding_max_resolution=64,
        nearness=1, xat_depth=2, xat_num_heads=8, dict_smpl_layer=None, person_center="head", clip_dist=1,
        num_betas=10, train_return_type="params", lr=5e-5, batch_size=8, max_iter=500000, weight_decay=1e-2,
        ckpt_dir="logs/checkpoints", det_thresh=0.2, training_data="bedlam", amp=1)


def test_body_model_from_smplx_npz_cpu(tmp_path):
    from multihmr_b200 import api, synth

    bm = synth.make_body_model(5)
    path = os.path.join(tmp_path, "models", "smplx", "SMPLX_NEUTRAL.npz")
    write_smplx_npz(path, bm)
    got = api.body_model_from_smplx_npz(path, num_betas=10)
    for k in ("v_template", "shapedirs", "shapedirs_extra", "expr_dirs", "posedirs", "J_regressor", "lbs_weights",
              "lmk_bary_coords"):
        assert torch.equal(got[k], bm[k]), k
    for k in ("faces", "lmk_faces_idx", "extra_joints_idxs"):
        assert torch.equal(got[k], bm[k].to(torch.int64)), k
    assert got["parents"][0].item() == -1 and torch.equal(got["parents"][1:], bm["parents"][1:])
    assert got["num_verts"] == bm["v_template"].shape[0]
    got11 = api.body_model_from_smplx_npz(path, num_betas=11)  # the neutral_11 layer: 11th (kid) shape component
    assert torch.equal(got11["shapedirs"][..., :10], bm["shapedirs"])
    assert torch.equal(got11["shapedirs"][..., 10:11], bm["shapedirs_extra"])


def test_load_model_missing_checkpoint_is_an_error(tmp_path, monkeypatch):
    from multihmr_b200 import api

    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError):
        api.load_model("multiHMR_896_L", device=torch.device("cpu"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["s_224_S_forced", "s_448_B_forced"])
def test_load_model_reproduces_reference_outputs(cuda_device, tmp_path, monkeypatch, name):
    from multihmr_b200 import api

    case, sd, bm, x, K, idx = pu.build_inputs(name)
    gold = pu.load_golden(name)
    monkeypatch.chdir(tmp_path)   # the reference resolves 'models/...' relative to the working directory
    write_smplx_npz(os.path.join("models", "smplx", "SMPLX_NEUTRAL.npz"), bm)
    os.makedirs(api.CACHE_DIR_MULTIHMR, exist_ok=True)
    ckpt = {"epoch": 3, "iter": 1234, "args": reference_args(case),
            "model_state_dict": {k: v for k, v in sd.items() if "smpl_layer" not in k}}
    torch.save(ckpt, os.path.join(api.CACHE_DIR_MULTIHMR, "multiHMR_synthetic.pt"))
    model = api.load_model("multiHMR_synthetic", device=cuda_device, max_batch=case["batch"], max_persons=16)
    assert model.img_size == case["img_size"] and model.patch_size == 14
    out = model(x.to(cuda_device), idx=idx, K=K.to(cuda_device), is_training=True)
    torch.cuda.synchronize()
    bad = pu.compare(out, gold, [k for k in gold if k != "idx"], focal=float(K[:, 0, 0].max()))
    assert not bad, bad
    persons = api.forward_model(model, x.to(cuda_device), K.to(cuda_device), det_thresh=0.3, nms_kernel_size=3)
    assert isinstance(persons, list)
