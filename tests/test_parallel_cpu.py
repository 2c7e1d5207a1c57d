"""CPU, world_size 2 over gloo: image sharding and the single all-gather of the per-rank record blocks
(host-side logic of the multi-GPU path; the engine itself is replaced by recorded per-rank outputs, the pack
kernel by its plain-torch statement — the kernel is checked against that statement in the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_outputs(rank, P, max_persons, V, nb):
    g = torch.Generator().manual_seed(100 + rank)
    r = lambda *s: torch.randn(*s, generator=g)
    det = torch.zeros(3, max_persons, dtype=torch.int32)
    det[0, :P] = torch.sort(torch.randint(0, 4, (P,), generator=g)).values.int()
    return {"det_idx": det, "det_score": r(max_persons), "loc": r(max_persons, 2), "transl": r(max_persons, 3),
            "transl_pelvis": r(max_persons, 3), "rotvec": r(max_persons, 53, 3), "expression": r(max_persons, 10),
            "shape": r(max_persons, nb), "v3d": r(max_persons, V, 3), "j3d": r(max_persons, 127, 3),
            "j2d": r(max_persons, 127, 2)}


class _FakeModel:
    """What RecordGather needs from a Model on the CPU path."""
    num_betas, num_verts, max_persons, max_batch = 10, 50, 6, 1
    device = torch.device("cpu")


def _expected_records(counts, max_persons, V, nb, B_per):
    from multihmr_b200 import parallel

    _, R = parallel.record_layout(nb, V)
    rows = []
    for r, P in enumerate(counts):
        t = _fake_outputs(r, P, max_persons, V, nb)
        block = parallel.pack_block_torch(t, P, r * B_per, max_persons, nb, V)
        rows.append(block[parallel.HEADER_WORDS:].view(max_persons, R)[:P])
    return torch.cat(rows)


def _worker(rank, world, port, counts, capacity, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multihmr_b200 import parallel

    m = _FakeModel()
    B_per = 4
    t = _fake_outputs(rank, counts[rank], m.max_persons, m.num_verts, m.num_betas)
    g = parallel.RecordGather(m, rank, world, capacity=capacity)
    g.gather_async(t, rank * B_per, counts[rank])
    valid, got_counts = g.result()
    persons = parallel.unpack_records(valid, m.num_betas, m.num_verts)
    q.put((rank, got_counts, [p["img"] for p in persons], valid.numpy().copy()))  # numpy: no fd passing
    dist.barrier()
    dist.destroy_process_group()


# capacity 4 < 6 detections on one rank: exercises the full-size second round (nothing is ever truncated)
@pytest.mark.parametrize("counts,capacity", [([3, 2], 4), ([0, 4], 4), ([6, 0], 4), ([5, 6], 6)])
def test_record_gather_world2(counts, capacity):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, counts, capacity, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = _expected_records(counts, 6, 50, 10, 4)
    for rank, got_counts, imgs, valid in results:
        assert got_counts == counts
        assert torch.equal(torch.from_numpy(valid), expect)  # every rank sees all persons, rank-major order
        assert imgs == sorted(imgs)                       # = global (b, y, x) order for contiguous shards
        assert all((i >= 4) == (k >= counts[0]) for k, i in enumerate(imgs))


def test_shard_range_is_a_contiguous_partition():
    from multihmr_b200.parallel import shard_range

    for B in (1, 7, 8, 64, 65):
        for W in (1, 2, 3, 8):
            spans = [shard_range(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_record_roundtrip():
    from multihmr_b200 import parallel

    t = _fake_outputs(0, 3, 5, 20, 10)
    _, R = parallel.record_layout(10, 20)
    block = parallel.pack_block_torch(t, 3, 8, 4, 10, 20)
    hdr = block[:parallel.HEADER_WORDS].view(torch.int32).tolist()
    assert hdr[:5] == [3, 3, 4, R, 8]
    rec = block[parallel.HEADER_WORDS:].view(4, R)
    assert torch.count_nonzero(rec[3]) == 0               # unused slot is zero-filled
    persons = parallel.unpack_records(rec[:3], 10, 20)
    assert len(persons) == 3 and persons[0]["v3d"].shape == (20, 3) and persons[0]["transl_pelvis"].shape == (1, 3)
    assert torch.equal(persons[1]["rotvec"], t["rotvec"][1])
    assert persons[2]["img"] == int(t["det_idx"][0, 2]) + 8
