"""Seeded synthetic assets with the exact names/shapes the reference loads (there is no network for the
published checkpoints, `SMPLX_NEUTRAL.npz` or `smpl_mean_params.npz`):

  make_state_dict  : `model_state_dict` keys of a Multi-HMR checkpoint (SURVEY.md Appendix B;
                     reference demo.py:87-103, train.py:195-207)
  make_mean_params : contents of models/smpl_mean_params.npz used at model.py:440-477
  make_body_model  : SMPL-X-shaped body model (V=10475, 55 joints, 486 pose-corrective features,
                     51 static landmarks, 21 vertex-picked joints) as consumed by smplx.create at
                     blocks/smpl_layer.py:38

Everything is generated on the CPU from torch.Generator seeds so the GPU box, the build container and
the golden-fixture script see bit-identical inputs.
"""
from __future__ import annotations

import math

import torch

BACKBONES = {
    "dinov2_vits14": dict(embed_dim=384, depth=12, num_heads=6),
    "dinov2_vitb14": dict(embed_dim=768, depth=12, num_heads=12),
    "dinov2_vitl14": dict(embed_dim=1024, depth=24, num_heads=16),
}
PATCH = 14
NUM_VERTS = 10475
NUM_FACES = 20908
NUM_JOINTS = 55
SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]
HPH_DIM = 1024
CAMERA_EMBED_DIM = 99


def _gen(seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(seed)


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g) * std


def make_state_dict(backbone: str = "dinov2_vitl14", img_size: int = 896, num_betas: int = 10,
                    xat_depth: int = 2, xat_num_heads: int = 8, seed: int = 0,
                    det_bias: float = -4.0) -> dict:
    """Random-init weights with trained-like scales.  `det_bias` shifts the detection logit so that only a
    few cells per image pass the 0.3 threshold."""
    cfg = BACKBONES[backbone]
    D, depth = cfg["embed_dim"], cfg["depth"]
    C = D + CAMERA_EMBED_DIM
    res = img_size // PATCH
    g = _gen(seed)
    sd = {}

    def lin(name, out_f, in_f, std=0.02, bias=True, bias_std=0.02):
        sd[name + ".weight"] = _randn(g, out_f, in_f, std=std)
        if bias:
            sd[name + ".bias"] = _randn(g, out_f, std=bias_std)

    def ln(name, dim):
        sd[name + ".weight"] = 1.0 + _randn(g, dim, std=0.1)
        sd[name + ".bias"] = _randn(g, dim, std=0.05)

    e = "backbone.encoder."
    sd[e + "cls_token"] = _randn(g, 1, 1, D, std=0.02)
    sd[e + "pos_embed"] = _randn(g, 1, 1 + 37 * 37, D, std=0.02)
    sd[e + "mask_token"] = torch.zeros(1, D)
    sd[e + "patch_embed.proj.weight"] = _randn(g, D, 3, PATCH, PATCH, std=0.02)
    sd[e + "patch_embed.proj.bias"] = _randn(g, D, std=0.02)
    for i in range(depth):
        b = f"{e}blocks.{i}."
        ln(b + "norm1", D)
        lin(b + "attn.qkv", 3 * D, D)
        lin(b + "attn.proj", D, D)
        sd[b + "ls1.gamma"] = 0.05 + 0.95 * torch.rand(D, generator=g)
        ln(b + "norm2", D)
        lin(b + "mlp.fc1", 4 * D, D)
        lin(b + "mlp.fc2", D, 4 * D)
        sd[b + "ls2.gamma"] = 0.05 + 0.95 * torch.rand(D, generator=g)
    ln(e + "norm", D)

    lin("mlp_classif.0", D, D)
    lin("mlp_classif.2", 1, D, std=0.07)
    sd["mlp_classif.2.bias"] = torch.full((1,), float(det_bias))
    lin("mlp_offset.0", D, D)
    lin("mlp_offset.2", 2, D, std=0.02)

    h = "x_attention_head."
    for nm in ("cross_queries_x", "cross_queries_y", "cross_values_x", "cross_values_y"):
        sd[h + nm] = _randn(g, res, C, std=0.2)
    mean = make_mean_params(seed)
    init_pose = torch.eye(3).reshape(1, 3, 3).repeat(53, 1, 1)[:, :, :2].flatten(1).reshape(1, -1)
    init_pose[:, : 24 * 6] = mean["pose"].float()
    sd[h + "init_body_pose"] = init_pose
    init_betas = mean["shape"].float().unsqueeze(0)
    sd[h + "init_betas_kid"] = torch.cat([init_betas, torch.zeros(1, 1)], 1)
    if num_betas == 11:
        init_betas = torch.cat([init_betas, torch.zeros(1, 1)], 1)
    sd[h + "init_betas"] = init_betas
    sd[h + "init_cam"] = mean["cam"].float().unsqueeze(0)
    sd[h + "init_expression"] = torch.zeros(1, 10)

    t = h + "transformer."
    token_dim = 318 + num_betas + 3 + C
    sd[t + "pos_embedding"] = _randn(g, 1, 1, HPH_DIM)
    lin(t + "to_token_embedding", HPH_DIM, token_dim)
    inner = xat_num_heads * 32
    for l in range(xat_depth):
        p = f"{t}transformer.layers.{l}."
        ln(p + "0.norm", HPH_DIM)
        lin(p + "0.fn.to_qkv", 3 * inner, HPH_DIM, bias=False)
        lin(p + "0.fn.to_out.0", HPH_DIM, inner)
        ln(p + "1.norm", HPH_DIM)
        lin(p + "1.fn.to_kv", 2 * inner, C, bias=False, std=0.03)
        lin(p + "1.fn.to_q", inner, HPH_DIM, bias=False, std=0.03)
        lin(p + "1.fn.to_out.0", HPH_DIM, inner)
        ln(p + "2.norm", HPH_DIM)
        lin(p + "2.fn.net.0", HPH_DIM, HPH_DIM)
        lin(p + "2.fn.net.3", HPH_DIM, HPH_DIM)
    # Trained-like conditioning of the 6D pose head: joints 24..52 start from the degenerate init
    # [1,0,0,1,0,0] (model.py:444-450: two identical columns after utils/humans.py:20), so a trained decoder
    # must add ~[0,0,0,-1,1,0] to produce a valid rotation.  Random weights without that offset make the
    # Gram-Schmidt step arbitrarily ill-conditioned (SURVEY.md §7 "hard parts"), which would measure the
    # conditioning of the fixture rather than the kernels.
    lin(h + "decpose", 318, HPH_DIM, std=0.005)
    pose_bias = sd[h + "decpose.bias"].view(53, 6)
    pose_bias[24:] += torch.tensor([0.0, 0.0, 0.0, -1.0, 1.0, 0.0])
    lin(h + "decshape", num_betas, HPH_DIM, std=0.01)
    lin(h + "deccam", 3, HPH_DIM, std=0.005)
    lin(h + "decexpression", 10, HPH_DIM, std=0.01)
    return sd


def add_outlier_channels(sd: dict, backbone: str, seed: int = 0, magnitude: float = 60.0) -> dict:
    """DINOv2-like 'massive activations' on top of make_state_dict (in place; the random stream of make_state_dict
    is untouched, so the other fixtures do not change): a few residual channels carry values of O(100) — on a few
    tokens from the position embedding on, and on every token after two MLP blocks whose fc2 rows / LayerScale for
    those channels are large.  Trained ViTs have such channels; N(0, 0.02) weights alone do not, and fp16
    tensor-core operands (Xn16 / QKV16 / H16) are stressed differently by them."""
    cfg = BACKBONES[backbone]
    D, depth = cfg["embed_dim"], cfg["depth"]
    g = _gen(seed + 909)
    ch = torch.randperm(D, generator=g)[:4]
    e = "backbone.encoder."
    pos = sd[e + "pos_embed"].clone()
    cells = 1 + torch.randperm(37 * 37, generator=g)[:40]          # ~3 % of the pretraining grid
    for c in ch[:2]:
        pos[0, cells, c] += magnitude * (0.5 + torch.rand(cells.numel(), generator=g))
    sd[e + "pos_embed"] = pos
    for l in (depth // 3, depth // 2):
        b = f"{e}blocks.{l}."
        w = sd[b + "mlp.fc2.weight"].clone()
        w[ch[2:]] *= 25.0
        sd[b + "mlp.fc2.weight"] = w
        gma = sd[b + "ls2.gamma"].clone()
        gma[ch[2:]] = 1.0
        sd[b + "ls2.gamma"] = gma
    return sd


def make_mean_params(seed: int = 0) -> dict:
    """pose[144] (24 joints x 6D, in the (a1, a2) order rot6d_to_rotmat reads, utils/humans.py:20),
    shape[10], cam[3]."""
    g = _gen(seed + 101)
    rv = _randn(g, 24, 3, std=0.3)
    ang = rv.norm(dim=1, keepdim=True).clamp_min(1e-8)
    ax = rv / ang
    Kx = torch.zeros(24, 3, 3)
    Kx[:, 0, 1], Kx[:, 0, 2] = -ax[:, 2], ax[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = ax[:, 2], -ax[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -ax[:, 1], ax[:, 0]
    R = torch.eye(3)[None] + torch.sin(ang)[:, :, None] * Kx + (1 - torch.cos(ang))[:, :, None] * (Kx @ Kx)
    pose6 = torch.cat([R[:, :, 0], R[:, :, 1]], dim=1).reshape(-1)  # [a1(3), a2(3)] per joint
    return {"pose": pose6, "shape": _randn(g, 10, std=0.5), "cam": torch.tensor([0.9, 0.0, 0.0])}


def make_body_model(seed: int = 0, num_verts: int = NUM_VERTS, num_faces: int = NUM_FACES) -> dict:
    g = _gen(seed + 202)
    V = num_verts
    bm = {}
    bm["v_template"] = _randn(g, V, 3) * torch.tensor([0.25, 0.45, 0.12])
    bm["shapedirs"] = _randn(g, V, 3, 10, std=0.01)
    bm["shapedirs_extra"] = _randn(g, V, 3, 1, std=0.01)  # 11th (kid) component of the neutral_11 layer
    bm["expr_dirs"] = _randn(g, V, 3, 10, std=0.005)
    bm["posedirs"] = _randn(g, (NUM_JOINTS - 1) * 9, V * 3, std=1e-3)
    Jr = torch.zeros(NUM_JOINTS, V)
    for j in range(NUM_JOINTS):
        ids = torch.randint(0, V, (32,), generator=g)
        w = torch.rand(32, generator=g)
        Jr[j].index_add_(0, ids, w / w.sum())
    bm["J_regressor"] = Jr
    W = torch.zeros(V, NUM_JOINTS)
    ids = torch.randint(0, NUM_JOINTS, (V, 4), generator=g)
    w = torch.rand(V, 4, generator=g) + 0.05
    W.scatter_add_(1, ids, w / w.sum(dim=1, keepdim=True))
    bm["lbs_weights"] = W
    bm["parents"] = torch.tensor(SMPLX_PARENTS, dtype=torch.int64)
    bm["faces"] = torch.randint(0, V, (num_faces, 3), generator=g)
    bm["lmk_faces_idx"] = torch.randint(0, num_faces, (51,), generator=g)
    bary = torch.rand(51, 3, generator=g) + 0.1
    bm["lmk_bary_coords"] = bary / bary.sum(dim=1, keepdim=True)
    bm["extra_joints_idxs"] = torch.randint(0, V, (21,), generator=g)
    return bm


def make_images(batch: int, img_size: int, seed: int = 0) -> torch.Tensor:
    """fp32 NCHW in the range normalize_rgb produces (utils/image.py:8-24)."""
    g = _gen(seed + 303)
    return torch.randn(batch, 3, img_size, img_size, generator=g).clamp_(-2.1, 2.6)


def make_images_u8(batch: int, img_size: int, seed: int = 0) -> torch.Tensor:
    """uint8 [B,S,S,3] RGB (HWC), what PIL + ImageOps.pad yield in demo.py:33-47 before normalize_rgb."""
    g = _gen(seed + 313)
    return torch.randint(0, 256, (batch, img_size, img_size, 3), generator=g, dtype=torch.uint8)


def make_cameras(batch: int, img_size: int, fov_deg=60.0, jitter: bool = False, seed: int = 0,
                 asymmetric: bool = False) -> torch.Tensor:
    """K [B,3,3] as demo.py:get_camera_parameters (demo.py:53-68); optional per-image fov jitter;
    `asymmetric` adds fx != fy and an off-centre principal point with cx != cy (exercises the (row, col)
    ordering of model.py:164-178)."""
    g = _gen(seed + 404)
    K = torch.eye(3).repeat(batch, 1, 1)
    for b in range(batch):
        fov = fov_deg + (float(torch.rand(1, generator=g)) * 20 - 10 if jitter else 0.0)
        f = img_size / (2 * math.tan(math.radians(fov) / 2))
        K[b, 0, 0] = K[b, 1, 1] = f
        K[b, 0, 2] = K[b, 1, 2] = img_size // 2
        if asymmetric:
            K[b, 1, 1] = f * (0.9 + 0.05 * b)
            K[b, 0, 2] = img_size * (0.40 + 0.03 * b)
            K[b, 1, 2] = img_size * (0.57 - 0.02 * b)
    return K


def make_forced_idx(batch: int, res: int, persons_per_image, seed: int = 0):
    """Distinct (b, y, x) cells in torch.where order, as the `idx=` argument of Model.forward
    (model.py:150-151, train.py:171-176): tuple (b, y, x, c=0) of int64 tensors."""
    g = _gen(seed + 505)
    if isinstance(persons_per_image, int):
        persons_per_image = [persons_per_image] * batch
    bs, ys, xs = [], [], []
    for b, n in enumerate(persons_per_image):
        cells = torch.randperm(res * res, generator=g)[:n].sort().values
        bs += [b] * n
        ys += (cells // res).tolist()
        xs += (cells % res).tolist()
    t = lambda v: torch.tensor(v, dtype=torch.int64)
    return (t(bs), t(ys), t(xs), torch.zeros(len(bs), dtype=torch.int64))
