"""Image-level data parallelism over the GPUs of one box.

The reference is single-GPU with no collectives (README.md:107); images are independent units of
`Model.forward`, so the batch is split into contiguous shards, every rank runs the full single-GPU engine
on its shard with replicated weights, and ONE all-gather of fixed-size per-person records returns the
result to every rank (NCCL over NVLink on GPUs; gloo in the CPU tests).  Ordering after the gather is
rank-major = global (b, y, x) order because the shards are contiguous.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

# per-person record layout (fp32): [global image index, score, loc(2), transl(3), transl_pelvis(3),
# rotvec(159), expression(10), shape(nb), v3d(3V), j3d(381), j2d(254)]
FIELDS = (("img", 1), ("scores", 1), ("loc", 2), ("transl", 3), ("transl_pelvis", 3), ("rotvec", 159),
          ("expression", 10))


def shard_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of `batch` images for `rank`; the first batch % world ranks get one more."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def record_layout(num_betas: int, num_verts: int):
    fields = FIELDS + (("shape", num_betas), ("v3d", 3 * num_verts), ("j3d", 381), ("j2d", 254))
    offs, o = {}, 0
    for name, n in fields:
        offs[name] = (o, n)
        o += n
    return offs, o


def pack_records(t: dict, P: int, img_offset: int, max_persons: int, num_betas: int, num_verts: int):
    """Device-side packing of the engine outputs (max_persons-sized tensors) into [max_persons, R]."""
    offs, R = record_layout(num_betas, num_verts)
    dev = t["v3d"].device
    rec = torch.zeros(max_persons, R, device=dev, dtype=torch.float32)
    if P > 0:
        src = {"img": (t["det_idx"][0, :P].float() + img_offset)[:, None], "scores": t["det_score"][:P, None],
               "loc": t["loc"][:P], "transl": t["transl"][:P], "transl_pelvis": t["transl_pelvis"][:P],
               "rotvec": t["rotvec"][:P].reshape(P, -1), "expression": t["expression"][:P], "shape": t["shape"][:P],
               "v3d": t["v3d"][:P].reshape(P, -1), "j3d": t["j3d"][:P].reshape(P, -1),
               "j2d": t["j2d"][:P].reshape(P, -1)}
        for name, (o, n) in offs.items():
            rec[:P, o:o + n] = src[name]
    return rec


def unpack_records(rec: torch.Tensor, num_betas: int, num_verts: int) -> list[dict]:
    """[P, R] -> list of person dicts with the reference's keys (model.py:329-347) + 'img' (global index)."""
    offs, _ = record_layout(num_betas, num_verts)
    shapes = {"loc": (2,), "transl": (3,), "transl_pelvis": (1, 3), "rotvec": (53, 3), "expression": (10,),
              "shape": (num_betas,), "v3d": (num_verts, 3), "j3d": (127, 3), "j2d": (127, 2)}
    persons = []
    for i in range(rec.shape[0]):
        p = {"img": int(rec[i, 0].item()), "scores": rec[i, 1]}
        for name, shp in shapes.items():
            o, n = offs[name]
            p[name] = rec[i, o:o + n].reshape(shp)
        persons.append(p)
    return persons


def all_gather_persons(rec: torch.Tensor, count: int, group=None) -> tuple[torch.Tensor, list[int]]:
    """One all-gather of the padded record blocks (+ one of the counts).  Returns the valid records of all
    ranks concatenated in rank order and the per-rank counts.  Every rank must pass the same max_persons."""
    world = dist.get_world_size(group)
    cnt = torch.tensor([count], device=rec.device, dtype=torch.int32)
    counts = torch.empty(world, device=rec.device, dtype=torch.int32)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    out = torch.empty(world * rec.shape[0], rec.shape[1], device=rec.device, dtype=rec.dtype)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    counts = counts.tolist()
    blocks = out.view(world, rec.shape[0], rec.shape[1])
    valid = torch.cat([blocks[r, :counts[r]] for r in range(world)], dim=0)
    return valid, counts


class ShardedModel:
    """Runs `model` (a multihmr_b200.Model on this rank's GPU) on this rank's contiguous image shard of a
    global batch and all-gathers the persons."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def forward(self, x_global, K_global, det_thresh=0.3, nms_kernel_size=3, gather=True):
        lo, hi = shard_range(x_global.shape[0], self.world, self.rank)
        m = self.model
        if hi > lo:
            t, P = m.forward_raw(x_global[lo:hi], K_global[lo:hi], det_thresh=det_thresh,
                                 nms_kernel_size=nms_kernel_size)
            rec = pack_records(t, P, lo, m.max_persons, m.num_betas, m.num_verts)
        else:
            _, R = record_layout(m.num_betas, m.num_verts)
            P, rec = 0, torch.zeros(m.max_persons, R, device=m.device)
        if not gather:
            return rec[:P], [P]
        return all_gather_persons(rec, P, self.group)

    __call__ = forward
