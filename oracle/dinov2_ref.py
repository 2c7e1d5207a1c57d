"""Restatement of the DINOv2 ViT forward that `blocks/dinov2.py:12,25` obtains through
`torch.hub.load('facebookresearch/dinov2', 'dinov2_vit{s,b,l}14')` (hub `main`, unpinned; not vendored in
/root/reference, not installed here).  Follows the published `DinoVisionTransformer` of
facebookresearch/dinov2 `dinov2/models/vision_transformer.py`:
  patch_embed (Conv2d 14x14/14) -> [cls | patches] + interpolate_pos_encoding(bicubic, scale_factor with
  the 0.1 offset) -> depth x { x += ls1(attn(norm1 x)); x += ls2(mlp(norm2 x)) } -> norm -> drop cls,
which is what `get_intermediate_layers(x)[0]` (n=1, norm=True, return_class_token=False) returns.
A second local source for the block structure is transformers' modeling_dinov2.py (LN eps 1e-6,
LayerScale, exact GELU, qkv bias).  TEST INFRASTRUCTURE ONLY — fp32, plain torch."""
import math

import torch
import torch.nn.functional as F

ARCHS = {
    "dinov2_vits14": dict(embed_dim=384, depth=12, num_heads=6),
    "dinov2_vitb14": dict(embed_dim=768, depth=12, num_heads=12),
    "dinov2_vitl14": dict(embed_dim=1024, depth=24, num_heads=16),
}
PATCH = 14
PRETRAIN_GRID = 37  # 518 / 14
LN_EPS = 1e-6
INTERP_OFFSET = 0.1


def interpolate_pos_encoding(pos_embed: torch.Tensor, w: int, h: int) -> torch.Tensor:
    """pos_embed [1, 1+37^2, D] -> [1, 1+(w/14)*(h/14), D] (fp32 bicubic, antialias off)."""
    N = pos_embed.shape[1] - 1
    w0, h0 = w // PATCH, h // PATCH
    if w0 * h0 == N and w == h:
        return pos_embed
    pos = pos_embed.float()
    cls_pos, patch_pos = pos[:, 0], pos[:, 1:]
    D = pos.shape[-1]
    M = int(math.sqrt(N))
    assert M * M == N
    sx = float(w0 + INTERP_OFFSET) / M
    sy = float(h0 + INTERP_OFFSET) / M
    patch_pos = F.interpolate(patch_pos.reshape(1, M, M, D).permute(0, 3, 1, 2), mode="bicubic",
                              antialias=False, scale_factor=(sx, sy))
    assert (w0, h0) == tuple(patch_pos.shape[-2:])
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, D)
    return torch.cat((cls_pos.unsqueeze(0), patch_pos), dim=1)


def vit_block(x, sd, pre, num_heads, emulate=None):
    """One pre-LN block with LayerScale; `emulate` optionally rounds GEMM operands (precision studies)."""
    lin = _linear if emulate is None else emulate
    B, T, D = x.shape
    hd = D // num_heads
    y = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], LN_EPS)
    qkv = lin(y, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(B, T, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if emulate is None:
        o = F.scaled_dot_product_attention(q, k, v)  # softmax(q k^T / sqrt(hd)) v
    else:
        o = emulate.attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, T, D)
    o = lin(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    x = x + sd[pre + "ls1.gamma"] * o
    y = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], LN_EPS)
    y = lin(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
    y = F.gelu(y)
    y = lin(y, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + sd[pre + "ls2.gamma"] * y


def _linear(x, w, b):
    return F.linear(x, w, b)


def prepare_tokens(x_img, sd, prefix="", emulate=None):
    B, _, H, W = x_img.shape
    w = sd[prefix + "patch_embed.proj.weight"]
    b = sd[prefix + "patch_embed.proj.bias"]
    if emulate is None:
        x = F.conv2d(x_img, w, b, stride=PATCH)
    else:
        x = emulate.conv(x_img, w, b, PATCH)
    x = x.flatten(2).transpose(1, 2)  # [B, N, D], row-major (y, x)
    cls = sd[prefix + "cls_token"].expand(B, -1, -1)
    x = torch.cat((cls, x), dim=1)
    return x + interpolate_pos_encoding(sd[prefix + "pos_embed"], W, H)


def get_intermediate_layers(x_img, sd, name, prefix="", emulate=None, taps=None):
    """== encoder.get_intermediate_layers(x)[0]: final-norm'ed patch tokens [B, N, D]."""
    cfg = ARCHS[name]
    x = prepare_tokens(x_img, sd, prefix, emulate)
    if taps is not None:
        taps["tokens0"] = x
    for i in range(cfg["depth"]):
        x = vit_block(x, sd, f"{prefix}blocks.{i}.", cfg["num_heads"], emulate)
        if taps is not None and i == 0:
            taps["block0"] = x
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], LN_EPS)
    return x[:, 1:]


class HubModelShim(torch.nn.Module):
    """What `torch.hub.load('facebookresearch/dinov2', name)` must look like for blocks/dinov2.py:
    attributes patch_size / embed_dim, method get_intermediate_layers(x) -> tuple, and parameters whose
    state_dict keys equal the hub model's (`cls_token`, `pos_embed`, `patch_embed.proj.*`, `blocks.i.*`,
    `norm.*`).  Used by oracle/make_golden.py to run the UNMODIFIED reference model.py in this container."""

    def __init__(self, name):
        super().__init__()
        cfg = ARCHS[name]
        D, depth = cfg["embed_dim"], cfg["depth"]
        self.name_, self.patch_size, self.embed_dim = name, PATCH, D
        P = torch.nn.Parameter
        z = torch.zeros
        self.cls_token = P(z(1, 1, D))
        self.pos_embed = P(z(1, 1 + PRETRAIN_GRID**2, D))
        self.mask_token = P(z(1, D))
        self.patch_embed = torch.nn.Module()
        self.patch_embed.proj = torch.nn.Conv2d(3, D, PATCH, PATCH)
        blocks = []
        for _ in range(depth):
            blk = torch.nn.Module()
            blk.norm1 = torch.nn.LayerNorm(D, eps=LN_EPS)
            blk.norm2 = torch.nn.LayerNorm(D, eps=LN_EPS)
            blk.attn = torch.nn.Module()
            blk.attn.qkv = torch.nn.Linear(D, 3 * D)
            blk.attn.proj = torch.nn.Linear(D, D)
            blk.ls1 = torch.nn.Module()
            blk.ls1.gamma = P(torch.ones(D))
            blk.ls2 = torch.nn.Module()
            blk.ls2.gamma = P(torch.ones(D))
            blk.mlp = torch.nn.Module()
            blk.mlp.fc1 = torch.nn.Linear(D, 4 * D)
            blk.mlp.fc2 = torch.nn.Linear(4 * D, D)
            blocks.append(blk)
        self.blocks = torch.nn.ModuleList(blocks)
        self.norm = torch.nn.LayerNorm(D, eps=LN_EPS)

    def get_intermediate_layers(self, x):
        sd = {k: v for k, v in self.state_dict().items()}
        return (get_intermediate_layers(x, sd, self.name_),)
