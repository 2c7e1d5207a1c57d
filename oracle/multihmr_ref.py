"""CPU restatement (plain torch, fp32) of the reference's own hot-path code: `Model.forward` and everything
it calls inside /root/reference (model.py, blocks/{camera_embed,cross_attn_transformer,smpl_layer}.py,
utils/{camera,image,tensor_manip,humans}.py).  Functional style over a flat `state_dict`; every function
cites the reference file:line it follows.  The third-party arithmetic comes from dinov2_ref / smplx_ref /
roma_ref.  Pinned against the unmodified reference by oracle/make_golden.py + tests/test_oracle_golden.py.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import dinov2_ref, roma_ref, smplx_ref

PATCH = 14


@dataclass
class RefConfig:
    backbone: str = "dinov2_vitl14"
    img_size: int = 896
    num_betas: int = 10
    xat_depth: int = 2
    xat_num_heads: int = 8
    person_center_idx: int = 15   # 'head' in JOINT_NAMES (model.py:45, blocks/smpl_layer.py:41-45)
    fovn: float = 60.0            # model.py:68
    num_bands: int = 16           # model.py:38
    max_resolution: int = 64      # model.py:39
    dim_head: int = 32            # model.py:127


# ------------------------------------------------------------------ utils/camera.py
def inverse_perspective_projection(points, K, distance):
    """utils/camera.py:30-48."""
    pts = torch.cat([points, torch.ones_like(points[..., :1])], -1)
    pts = torch.einsum("bij,bkj->bki", torch.inverse(K), pts)
    return pts if distance is None else pts * distance


def perspective_projection(x, K):
    """utils/camera.py:14-27."""
    y = x / x[:, :, -1].unsqueeze(-1)
    return torch.einsum("bij,bkj->bki", K, y)[:, :, :2]


def focal_from_fov(fov, img_size):
    """utils/camera.py:50-60."""
    return img_size / (2 * math.tan(math.radians(fov) / 2))


# ------------------------------------------------------------------ blocks/camera_embed.py
def fourier_features(pos, num_bands, max_resolution):
    """blocks/camera_embed.py:39-58: [pos | sin(pi f p) | cos(pi f p)], f = linspace(1, res/2, bands) per dim,
    feature order dim-major then band."""
    b, n, d = pos.shape
    freqs = torch.stack([torch.linspace(1.0, max_resolution / 2, num_bands, device=pos.device)
                         for _ in range(d)], 0)                  # [d, bands]
    feats = (pos[:, :, :, None] * freqs[None, None]).reshape(b, n, -1)
    feats = torch.cat([torch.sin(math.pi * feats), torch.cos(math.pi * feats)], -1)
    return torch.cat([pos, feats], -1)


def embed_camera(K, h, w, cfg):
    """model.py:160-187.  NB the (row, col) grid is fed as (x, y) to the un-projection."""
    bs = K.shape[0]
    pts = torch.stack([torch.arange(h).reshape(-1, 1).repeat(1, w),
                       torch.arange(w).reshape(1, -1).repeat(h, 1)], -1).to(K.device).float()
    pts = pts * PATCH + PATCH // 2
    pts = pts.reshape(1, -1, 2).repeat(bs, 1, 1)
    rays = inverse_perspective_projection(pts, K, torch.ones(bs, pts.shape[1], 1, device=K.device))
    return fourier_features(rays, cfg.num_bands, cfg.max_resolution).reshape(bs, h, w, -1)


# ------------------------------------------------------------------ model.py heads
def regression_mlp(x, sd, name, lin=F.linear):
    """model.py:596-609 with two layers: Linear -> ReLU -> Linear."""
    y = torch.relu(lin(x, sd[name + ".0.weight"], sd[name + ".0.bias"]))
    return lin(y, sd[name + ".2.weight"], sd[name + ".2.bias"])


def nms(heat, kernel):
    """model.py:620-638."""
    pad = (kernel - 1) // 2 if kernel not in (2, 4) else (1 if kernel == 2 else 2)
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=pad)
    if hmax.shape[2] > heat.shape[2]:
        hmax = hmax[:, :, : heat.shape[2], : heat.shape[3]]
    return heat * (hmax == heat).float()


def detection(z, sd, nms_kernel_size, det_thresh, idx, is_training, lin=F.linear):
    """model.py:133-158 (+ _sigmoid :641-643, unpatch utils/image.py:39-51, apply_threshold :612-617)."""
    B, N, _ = z.shape
    h = w = int(math.sqrt(N))
    s = torch.clamp(torch.sigmoid(regression_mlp(z, sd, "mlp_classif", lin)), min=1e-4, max=1 - 1e-4)
    scores = s.reshape(B, h, w, 1).permute(0, 3, 1, 2)           # unpatch(patch_size=1): [B,1,h,w]
    if not is_training:
        if nms_kernel_size > 1:
            scores = nms(scores, nms_kernel_size)
        idx = torch.where(scores.permute(0, 2, 3, 1) >= (det_thresh[0] if isinstance(det_thresh, list)
                                                         else det_thresh))
    else:
        assert idx is not None
    scores_det = scores[idx[0], idx[3], idx[1], idx[2]]
    return scores.permute(0, 2, 3, 1), scores_det, idx


# ------------------------------------------------------------------ blocks/cross_attn_transformer.py
def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _heads(t, h):
    b, n, _ = t.shape
    return t.reshape(b, n, h, -1).permute(0, 2, 1, 3)


def self_attention(x, mask, sd, p, heads, dim_head, lin=F.linear):
    """Attention.forward, cross_attn_transformer.py:129-159 (PreNorm applied by the caller)."""
    q, k, v = lin(x, sd[p + "to_qkv.weight"], None).chunk(3, dim=-1)
    q, k, v = (_heads(t, heads) for t in (q, k, v))
    if mask is not None:
        q, k, v = (t * mask[:, None, :, None] for t in (q, k, v))
    dots = torch.matmul(q, k.transpose(-1, -2)) * dim_head**-0.5
    if mask is not None:
        dots = dots - (1 - mask)[:, None, None, :] * 10e10
    attn = dots.softmax(dim=-1)
    if mask is not None:
        attn = attn * mask[:, None, None, :]
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(x.shape[0], x.shape[1], -1)
    return lin(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def cross_attention(x, context, mask, sd, p, heads, dim_head, lin=F.linear, lin_kv=None):
    """CrossAttention.forward, cross_attn_transformer.py:185-205 (context is NOT normalised)."""
    k, v = (lin_kv or lin)(context, sd[p + "to_kv.weight"], None).chunk(2, dim=-1)
    q = lin(x, sd[p + "to_q.weight"], None)
    q, k, v = (_heads(t, heads) for t in (q, k, v))
    if mask is not None:
        q = q * mask[:, None, :, None]
    dots = torch.matmul(q, k.transpose(-1, -2)) * dim_head**-0.5
    if mask is not None:
        dots = dots - (1 - mask).float()[:, None, :, None] * 1e6
    out = torch.matmul(dots.softmax(dim=-1), v)
    if mask is not None:
        out = out * mask[:, None, :, None]
    out = out.permute(0, 2, 1, 3).reshape(x.shape[0], x.shape[1], -1)
    return lin(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def transformer_decoder(token, context, mask, sd, cfg, lin=F.linear, lin_kv=None):
    """TransformerDecoder.forward :351-359 + TransformerCrossAttn.forward :239-261 (dropouts are identity)."""
    t = "x_attention_head.transformer."
    x = lin(token, sd[t + "to_token_embedding.weight"], sd[t + "to_token_embedding.bias"])
    x = x + sd[t + "pos_embedding"][:, 0][:, None, :]
    for l in range(cfg.xat_depth):
        p = f"{t}transformer.layers.{l}."
        if mask is not None:
            x = x * mask[:, :, None]
        x = self_attention(_ln(x, sd, p + "0.norm"), mask, sd, p + "0.fn.", cfg.xat_num_heads, cfg.dim_head, lin) + x
        x = cross_attention(_ln(x, sd, p + "1.norm"), context, mask, sd, p + "1.fn.", cfg.xat_num_heads,
                            cfg.dim_head, lin, lin_kv) + x
        y = _ln(x, sd, p + "2.norm")
        y = lin(F.gelu(lin(y, sd[p + "2.fn.net.0.weight"], sd[p + "2.fn.net.0.bias"])),
                sd[p + "2.fn.net.3.weight"], sd[p + "2.fn.net.3.bias"])
        x = y + x
    if mask is not None:
        x = x * mask[:, :, None]
    return x


# ------------------------------------------------------------------ HPH (model.py:352-593)
def rot6d_to_rotmat(x):
    """utils/humans.py:12-22."""
    return roma_ref.special_gramschmidt(x.reshape(-1, 2, 3).permute(0, 2, 1).contiguous())


def hph_forward(z_central, z_all_tokens, idx, sd, cfg, lin=F.linear, lin_kv=None, taps=None):
    """HPH.forward model.py:527-593 with cross_attn_inputs :479-525, rebatch / pad_to_max
    (utils/tensor_manip.py:7-45).  `z_all_tokens` is [B, N, C] (the reference carries one [C,h,w] copy per
    person, model.py:278-280, and keeps one per non-empty image at :511 — same values)."""
    h_ = "x_attention_head."
    b_idx, y_idx, x_idx = idx[0], idx[1], idx[2]
    P = b_idx.shape[0]
    res = int(math.sqrt(z_all_tokens.shape[1]))
    imgs, dense, counts = torch.unique(b_idx, sorted=True, return_inverse=True, return_counts=True)  # rebatch
    q = z_central + sd[h_ + "cross_queries_x"][y_idx] + sd[h_ + "cross_queries_y"][x_idx]      # :500-504
    maxp = int(counts.max())
    Bp = imgs.shape[0]
    C = q.shape[1]
    padded = q.new_zeros(Bp, maxp, C)
    mask = q.new_zeros(Bp, maxp)
    slot = torch.arange(P, device=q.device) - (torch.cumsum(counts, 0) - counts)[dense]          # pad_to_max
    padded[dense, slot] = q
    mask[dense, slot] = 1.0
    context = z_all_tokens[imgs].clone()                                                        # :511
    vals = sd[h_ + "cross_values_x"][y_idx] + sd[h_ + "cross_values_y"][x_idx]                  # :514-517
    context[dense, y_idx * res + x_idx] = context[dense, y_idx * res + x_idx] + vals
    init_pose, init_betas = sd[h_ + "init_body_pose"], sd[h_ + "init_betas"]
    init_cam, init_expr = sd[h_ + "init_cam"], sd[h_ + "init_expression"]
    ex = lambda t: t.expand(Bp, maxp, -1)
    token = torch.cat([padded, ex(init_pose), ex(init_betas), ex(init_cam)], dim=-1)            # :550
    out = transformer_decoder(token, context, mask, sd, cfg, lin, lin_kv)                       # :555
    tok = out[dense, slot]                                                                      # :558-561
    if taps is not None:
        taps["token_out"] = tok
    dec = lambda n, init: lin(tok, sd[h_ + n + ".weight"], sd[h_ + n + ".bias"]) + init          # :571-575
    pose6 = dec("decpose", init_pose)
    betas = dec("decshape", init_betas)
    cam = dec("deccam", init_cam)
    expr = dec("decexpression", init_expr)
    if taps is not None:
        taps["pose6d"] = pose6
    rotmat = rot6d_to_rotmat(pose6).view(P, 53, 3, 3)                                           # :578-583
    return rotmat, betas, expr, cam


# ------------------------------------------------------------------ blocks/smpl_layer.py
def smpl_layer_forward(body, pose, shape, loc, dist, K, expression, person_center_idx):
    """SMPL_Layer.forward blocks/smpl_layer.py:47-155; `body` is an smplx_ref.SMPLXShim."""
    bs = pose.shape[0]
    z3 = pose.new_zeros(bs, 3)
    out = body(betas=shape, global_orient=z3, body_pose=pose[:, 1:22].flatten(1),
               left_hand_pose=pose[:, 22:37].flatten(1), right_hand_pose=pose[:, 37:52].flatten(1),
               jaw_pose=pose[:, 52:53].flatten(1), expression=expression.flatten(1), leye_pose=z3,
               reye_pose=z3)
    verts, j3d = out.vertices, out.joints
    R = roma_ref.rotvec_to_rotmat(pose[:, 0])
    pelvis = j3d[:, [0]]
    j3d = (R.unsqueeze(1) @ (j3d - pelvis).unsqueeze(-1)).squeeze(-1)
    verts = (R.unsqueeze(1) @ (verts - pelvis).unsqueeze(-1)).squeeze(-1)
    transl = inverse_perspective_projection(loc.unsqueeze(1), K, dist.unsqueeze(1))[:, 0]
    center = j3d[:, [person_center_idx]]
    verts = verts - center
    j3d = j3d - center
    j3d_cam = j3d + transl.unsqueeze(1)
    v_cam = verts + transl.unsqueeze(1)
    return {"v3d": v_cam, "j3d": j3d_cam, "j2d": perspective_projection(j3d_cam, K),
            "v2d": perspective_projection(v_cam, K), "transl": transl, "transl_pelvis": j3d_cam[:, [0]]}


# ------------------------------------------------------------------ Model.forward (model.py:205-349)
def model_forward(sd, body, cfg, x, K, idx=None, det_thresh=0.3, nms_kernel_size=3, is_training=False,
                  emulate=None, taps=None):
    """Returns the person list (inference) or the flat dict (is_training=True), like the reference."""
    lin = F.linear if emulate is None else emulate
    z = dinov2_ref.get_intermediate_layers(x, sd, cfg.backbone, "backbone.encoder.", emulate, taps)  # :229
    B, N, D = z.shape
    h = w = int(math.sqrt(N))
    if taps is not None:
        taps["z"] = z
    scores, scores_det, idx = detection(z, sd, nms_kernel_size, det_thresh, idx, is_training, lin)  # :233-240
    if taps is not None:
        taps["scores"] = scores
    if len(idx[0]) == 0 and not is_training:
        return []
    b_idx, y_idx, x_idx = idx[0], idx[1], idx[2]
    z_central = z[b_idx, y_idx * w + x_idx]                                                  # :246-255
    offset = regression_mlp(z_central, sd, "mlp_offset")                                     # :258
    K_det = K[b_idx]
    z_K = embed_camera(K, h, w, cfg)                                                         # :262
    if taps is not None:
        taps["z_K"] = z_K
    z_central = torch.cat([z_central, z_K[b_idx, y_idx, x_idx]], 1)                          # :263-265
    z_all = torch.cat([z, z_K.reshape(B, N, -1)], 2)                                         # :266-268
    loc = (torch.stack([x_idx, y_idx]).permute(1, 0) + 0.5 + offset) * PATCH                 # :272-275
    lin_kv = None if emulate is None else emulate
    rotmat, shape, expression, cam = hph_forward(z_central, z_all, idx, sd, cfg,
                                                 F.linear, lin_kv, taps)                     # :281-283
    rotvec = roma_ref.rotmat_to_rotvec(rotmat)                                               # :291
    dist_pp = cam[:, 0][:, None]
    focal = K_det[:, [0], [0]]
    dist = dist_pp * (focal / focal_from_fov(cfg.fovn, x.shape[-1]))                         # :189-193
    dist = torch.clamp(torch.exp(dist) - 1e-10, 0, 50)                                       # :196-201
    out = {"dist_postprocessed": dist_pp, "scores": scores, "offset": offset, "dist": dist,
           "expression": expression, "rotmat": rotmat, "shape": shape, "rotvec": rotvec, "loc": loc}
    out.update(smpl_layer_forward(body, rotvec, shape, loc, dist, K_det, expression,
                                  cfg.person_center_idx))                                   # :319-321
    if is_training:
        return out
    keys = ("loc", "transl", "transl_pelvis", "rotvec", "expression", "shape", "v3d", "j3d", "j2d")
    persons = []
    for i in range(b_idx.shape[0]):                                                          # :329-347
        person = {"scores": scores_det[i]}
        person.update({k: out[k][i] for k in keys})
        persons.append(person)
    return persons
