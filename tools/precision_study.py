"""Per-rounding-point error budget of the engine's precision contract, on the CPU (oracle `emulate` hook).

Emulates where the engine rounds to fp16 (tensor-core operands / fp16 workspaces) inside the fp32 oracle and
reports the output error vs the golden fixture (= the unmodified reference) with each rounding class switched
off / split in two fp16 terms.  TEST/DIAGNOSTIC TOOL, never on the product path.

    python tools/precision_study.py [case] [--classes ...]
"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_util as pu  # noqa: E402
from oracle import multihmr_ref, smplx_ref  # noqa: E402


def r16(t):
    return t.to(torch.float16).to(torch.float32)


def split16(t):
    """two-term fp16 representation (hi + lo): ~22 bits"""
    hi = r16(t)
    return hi + r16(t - hi)


class Emu:
    """rounding points:  <gemm>.in / <gemm>.w / <gemm>.out for gemm in {patch,qkv,proj,fc1,fc2,cls0,kv};
    attn.p (P to fp16), attn.o == proj.in (O16).  mode per point: 'r' round, 'n' none, 's' split (2-term)."""

    def __init__(self, sd, modes=None, default="r", layer_modes=None):
        self.names = {}
        for k, v in sd.items():
            if not k.endswith(".weight"):
                continue
            for tag, cls in (("attn.qkv", "qkv"), ("attn.proj", "proj"), ("mlp.fc1", "fc1"), ("mlp.fc2", "fc2"),
                             ("mlp_classif.0", "cls0"), ("mlp_classif.2", "cls2"), ("to_kv", "kv")):
                if tag in k:
                    layer = -1
                    if "blocks." in k:
                        layer = int(k.split("blocks.")[1].split(".")[0])
                    self.names[id(v)] = (cls, layer)
        self.modes = modes or {}
        self.default = default
        self.layer_modes = layer_modes or {}   # {(point, layer): mode}

    def mode(self, point, layer=-1):
        if (point, layer) in self.layer_modes:
            return self.layer_modes[(point, layer)]
        return self.modes.get(point, self.default)

    def apply(self, t, point, layer=-1):
        m = self.mode(point, layer)
        if m == "r":
            return r16(t)
        if m == "s":
            return split16(t)
        return t

    def __call__(self, x, w, b):
        cls, layer = self.names.get(id(w), (None, -1))
        if cls is None:
            return F.linear(x, w, b)
        if cls == "cls2":
            return F.linear(self.apply(x, "cls2.in"), w, b)
        y = F.linear(self.apply(x, cls + ".in", layer), self.apply(w, cls + ".w", layer), b)
        if cls == "fc1":
            return y  # GELU follows; the rounding of H16 is fc2.in
        if cls in ("qkv",):
            y = self.apply(y, "qkv.out", layer)
        return y

    def attention(self, q, k, v):
        # S fp32, online softmax == plain softmax up to fp32 rounding; P rounded to fp16 un-normalised
        s = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
        m = s.amax(dim=-1, keepdim=True)
        p = torch.exp(s - m)
        l = p.sum(dim=-1, keepdim=True)
        o = torch.matmul(self.apply(p, "attn.p"), v) / l
        return o  # rounding to O16 is proj.in

    def conv(self, x_img, w, b, patch):
        return F.conv2d(self.apply(x_img, "patch.in"), self.apply(w, "patch.w"), b, stride=patch)


POINTS = ["patch.in", "patch.w", "qkv.in", "qkv.w", "qkv.out", "attn.p", "proj.in", "proj.w", "fc1.in", "fc1.w",
          "fc2.in", "fc2.w", "cls0.in", "cls0.w", "kv.in", "kv.w"]


def run(name, emu_factory):
    case, sd, bm, x, K, idx = pu.build_inputs(name)
    gold = pu.load_golden(name)
    cfg = multihmr_ref.RefConfig(backbone=case["backbone"], img_size=case["img_size"])
    body = smplx_ref.SMPLXShim(bm, 10)
    res = {}
    with torch.no_grad():
        ref = multihmr_ref.model_forward(sd, body, cfg, x, K, idx=idx, is_training=True, taps=(tr := {}))
        for label, emu in emu_factory(sd):
            taps = {}
            out = multihmr_ref.model_forward(sd, body, cfg, x, K, idx=idx, is_training=True, emulate=emu, taps=taps)
            e = {}
            for k in ("v3d", "rotmat", "shape", "dist", "scores", "expression", "transl"):
                d = (out[k] - gold[k]).abs()
                e[k] = (d.max().item(), d.pow(2).mean().sqrt().item())
            dz = (taps["z"] - tr["z"])
            e["z"] = (dz.abs().max().item(), dz.pow(2).mean().sqrt().item())
            res[label] = e
            print(f"{label:28s} v3d max {e['v3d'][0]:.3e} rms {e['v3d'][1]:.3e} | rotmat max {e['rotmat'][0]:.3e} rms "
                  f"{e['rotmat'][1]:.3e} | z rms {e['z'][1]:.3e} | shape {e['shape'][0]:.2e} dist {e['dist'][0]:.2e} "
                  f"scores {e['scores'][0]:.2e}", flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", nargs="?", default="s_280_L_forced")
    ap.add_argument("--study", default="ablate")
    a = ap.parse_args()
    torch.set_num_threads(min(32, os.cpu_count()))

    def factory(sd):
        yield "all-rounded (engine r01)", Emu(sd)
        yield "none (fp32)", Emu(sd, default="n")
        if a.study == "ablate":
            for p in POINTS:
                yield f"without {p}", Emu(sd, modes={p: "n"})
        elif a.study == "groups":
            g = {
                "weights": [p for p in POINTS if p.endswith(".w")],
                "acts": [p for p in POINTS if not p.endswith(".w")],
                "attn-branch": ["qkv.in", "qkv.w", "qkv.out", "attn.p", "proj.in", "proj.w"],
                "mlp-branch": ["fc1.in", "fc1.w", "fc2.in", "fc2.w"],
                "head": ["cls0.in", "cls0.w", "kv.in", "kv.w"],
                "resid-writers(proj,fc2)": ["proj.in", "proj.w", "fc2.in", "fc2.w"],
                "ln-consumers(qkv,fc1)": ["qkv.in", "qkv.w", "fc1.in", "fc1.w"],
            }
            for label, pts in g.items():
                yield f"without {label}", Emu(sd, modes={p: "n" for p in pts})
        elif a.study == "layers":
            depth = 24 if "L" in a.case.split("_")[2] else 12
            for lo in range(0, depth, 4):
                lm = {(p, l): "n" for p in POINTS for l in range(lo, lo + 4)}
                yield f"without layers {lo}-{lo + 3}", Emu(sd, layer_modes=lm)

    run(a.case, factory)




# ---------------------------------------------------------------------------------------------------
# "refine" study: bulk pass with fp16 operands for every token + an fp32 second pass of the residual
# streams of the detected (central) tokens only, attending over the K/V of the bulk pass.
class RefineEmu(Emu):
    """pass 'record': behaves like Emu and records (k, v) per attention call; pass 'replay': fp32 linears,
    attention of the fp32 queries over the recorded K/V."""

    def __init__(self, sd, reuse_o=False, **kw):
        super().__init__(sd, **kw)
        self.kv = []
        self.o = []
        self.reuse_o = reuse_o   # replay the bulk pass's fp16 attention OUTPUT instead of re-attending
        self.replay = False
        self.i = 0

    def __call__(self, x, w, b):
        if self.replay:
            return F.linear(x, w, b)
        return super().__call__(x, w, b)

    def attention(self, q, k, v):
        if not self.replay:
            self.kv.append((k, v))
            o = super().attention(q, k, v)
            self.o.append(r16(o))
            return o
        k, v = self.kv[self.i]
        self.i += 1
        if self.reuse_o:
            return self.o[self.i - 1]
        s = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
        return torch.matmul(s.softmax(dim=-1), v)

    def conv(self, x_img, w, b, patch):
        if self.replay:
            return F.conv2d(x_img, w, b, stride=patch)
        return super().conv(x_img, w, b, patch)


def refine_study(name, reuse_o=False):
    from oracle import dinov2_ref
    case, sd, bm, x, K, idx = pu.build_inputs(name)
    gold = pu.load_golden(name)
    cfg = multihmr_ref.RefConfig(backbone=case["backbone"], img_size=case["img_size"])
    body = smplx_ref.SMPLXShim(bm, 10)
    with torch.no_grad():
        emu = RefineEmu(sd, reuse_o=reuse_o)
        zA = dinov2_ref.get_intermediate_layers(x, sd, cfg.backbone, "backbone.encoder.", emu)
        emu.replay = True
        zB = dinov2_ref.get_intermediate_layers(x, sd, cfg.backbone, "backbone.encoder.", emu)
        zR = dinov2_ref.get_intermediate_layers(x, sd, cfg.backbone, "backbone.encoder.")
        print("z rms err: bulk %.3e  refined rows %.3e" % ((zA - zR).pow(2).mean().sqrt(), (zB - zR).pow(2).mean().sqrt()))
        for label, zc_src in (("bulk only", zA), ("central rows refined", zB)):
            # tail of multihmr_ref.model_forward with z_central taken from zc_src
            emu2 = Emu(sd)
            B, N, D = zA.shape
            h = w = int(math.sqrt(N))
            scores, scores_det, idx_ = multihmr_ref.detection(zA, sd, 3, 0.3, idx, True, emu2)
            b_idx, y_idx, x_idx = idx[0], idx[1], idx[2]
            z_central = zc_src[b_idx, y_idx * w + x_idx]
            offset = multihmr_ref.regression_mlp(z_central, sd, "mlp_offset")
            K_det = K[b_idx]
            z_K = multihmr_ref.embed_camera(K, h, w, cfg)
            z_central = torch.cat([z_central, z_K[b_idx, y_idx, x_idx]], 1)
            z_all = torch.cat([zA, z_K.reshape(B, N, -1)], 2)
            loc = (torch.stack([x_idx, y_idx]).permute(1, 0) + 0.5 + offset) * 14
            rotmat, shape, expression, cam = multihmr_ref.hph_forward(z_central, z_all, idx, sd, cfg, F.linear, emu2)
            rotvec = multihmr_ref.roma_ref.rotmat_to_rotvec(rotmat)
            dist_pp = cam[:, 0][:, None]
            focal = K_det[:, [0], [0]]
            dist = dist_pp * (focal / multihmr_ref.focal_from_fov(cfg.fovn, x.shape[-1]))
            dist = torch.clamp(torch.exp(dist) - 1e-10, 0, 50)
            out = {"rotmat": rotmat, "shape": shape, "dist": dist, "scores": scores, "expression": expression, "loc": loc}
            out.update(multihmr_ref.smpl_layer_forward(body, rotvec, shape, loc, dist, K_det, expression, 15))
            print(label, {k: "%.3e" % (out[k] - gold[k]).abs().max().item() for k in
                          ("v3d", "rotmat", "shape", "dist", "scores", "expression", "transl", "loc", "j3d")})


if __name__ == "__main__":
    if "--refine" in sys.argv:
        torch.set_num_threads(min(32, os.cpu_count()))
        names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["s_280_L_forced"]
        refine_study(names[0], reuse_o="--reuse-o" in sys.argv)
    else:
        main()
