"""Evaluation metrics.  CPU: the restatement oracle/eval_ref.py against the golden outputs of
the reference's OWN utils/training.py functions (tests/golden/eval_matching.npz, oracle/make_golden.py).  GPU: the
device kernels behind the C-ABI against the golden and against the restatement."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_matching.npz")


def _cases():
    with np.load(GOLD) as f:
        n = int(f["n_cases"])
        return [{k: f[f"{k}{i}"] for k in ("pred", "gt", "best", "fp", "miss", "iou")} for i in range(n)], f["prf1"]


def test_matching_restatement_vs_reference_golden():
    from oracle import eval_ref

    cases, prf = _cases()
    for c in cases:
        vm = np.ones_like(c["gt"][..., 0]).astype(np.bool_)
        best, fps, misses = eval_ref.match_2d_greedy(c["pred"], c["gt"], vm)
        assert np.array_equal(best, c["best"]) and fps == c["fp"].tolist() and misses == c["miss"].tolist()
        for p in range(len(c["pred"])):
            for g in range(len(c["gt"])):
                assert eval_ref.get_bbx_overlap(c["pred"][p], c["gt"][g]) == c["iou"][p, g]
    for row in prf:
        assert tuple(eval_ref.compute_prf1(int(row[0]), int(row[1]), int(row[2]))) == tuple(row[3:])


def test_procrustes_restatement_recovers_a_known_similarity():
    """Second source for roma.rigid_points_registration: a known similarity transform is recovered, and the
    alignment error equals the one of an independent SVD (numpy) solution, reflections included."""
    from oracle import roma_ref

    rng = np.random.default_rng(5)
    for trial in range(6):
        x = rng.normal(size=(300, 3))
        Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        s, t = rng.uniform(0.5, 2.0), rng.normal(size=3)
        y = s * x @ Q.T + t + rng.normal(size=x.shape) * (0.0 if trial < 3 else 0.05)
        if trial == 5:
            y[:, 2] *= -1  # mirrored target: the best ROTATION is not the best orthogonal matrix
        R, tt, ss = roma_ref.rigid_points_registration(torch.tensor(x), torch.tensor(y), compute_scaling=True)
        R, tt, ss = R.numpy(), tt.numpy(), ss.item()
        assert abs(np.linalg.det(R) - 1) < 1e-9
        if trial < 3:
            assert np.abs(R - Q).max() < 1e-9 and abs(ss - s) < 1e-9 and np.abs(tt - t).max() < 1e-9
        # independent Umeyama
        xm, ym = x.mean(0), y.mean(0)
        H = (y - ym).T @ (x - xm)
        U, D, Vt = np.linalg.svd(H)
        S = np.diag([1, 1, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
        R2 = U @ S @ Vt
        s2 = np.trace(np.diag(D) @ S) / ((x - xm) ** 2).sum()
        assert np.abs(R - R2).max() < 1e-9 and abs(ss - s2) < 1e-9


@pytest.mark.gpu
def test_device_matching_vs_reference_golden(cuda_device):
    from multihmr_b200 import metrics

    cases, _ = _cases()
    for c in cases:
        pred, gt = torch.from_numpy(c["pred"]).to(cuda_device), torch.from_numpy(c["gt"]).to(cuda_device)
        pairs, n_pairs, p2g, g2p = metrics.match_2d_greedy(pred, gt)
        n = int(n_pairs.item())
        assert np.array_equal(pairs[:n].cpu().numpy().astype(np.int64), c["best"])
        assert [i for i, g in enumerate(p2g.tolist()) if g < 0] == c["fp"].tolist()
        assert [i for i, p in enumerate(g2p.tolist()) if p < 0] == c["miss"].tolist()


@pytest.mark.gpu
def test_device_pve_and_pa_pve_vs_restatement(cuda_device):
    from multihmr_b200 import metrics
    from oracle import eval_ref

    g = torch.Generator().manual_seed(9)
    V, P, G = 10475, 3, 4
    gt_v = torch.randn(G, V, 3, generator=g) * 0.4 + torch.tensor([0.0, 0.0, 5.0])
    gt_p = gt_v[:, 100].clone()
    order = [2, 0, 3]
    pred_v = torch.stack([1.1 * gt_v[i] @ torch.linalg.qr(torch.randn(3, 3, generator=g))[0].T
                          + torch.randn(V, 3, generator=g) * 0.01 for i in order])
    pred_p = pred_v[:, 100].clone()
    pairs = torch.tensor([[0, 2], [1, 0], [2, 3]], dtype=torch.int32, device=cuda_device)
    n_pairs = torch.tensor([3], dtype=torch.int32, device=cuda_device)
    err, pa = metrics.points_error(pred_v.to(cuda_device), gt_v.to(cuda_device), pairs, n_pairs,
                                   pred_p.to(cuda_device), gt_p.to(cuda_device))
    for m, (pid, gid) in enumerate(pairs.tolist()):
        e, pe = eval_ref.points_errors(pred_v[pid] - pred_p[pid], gt_v[gid] - gt_p[gid])
        assert abs(err[m].item() - e.item()) <= 1e-3 * max(1.0, e.item()), (m, err[m].item(), e.item())
        assert abs(pa[m].item() - pe.item()) <= 2e-3 * max(1.0, pe.item()), (m, pa[m].item(), pe.item())


@pytest.mark.gpu
def test_evaluator_on_engine_outputs(cuda_device):
    """Trainer.evaluate's loop on real engine outputs: ground truth = the golden outputs of the unmodified reference
    on the natural-detection case -> every person matched, PVE below 1 mm (the 'PVE vs ref' of BASELINE.json)."""
    import parity_util as pu
    from multihmr_b200 import metrics

    name = "s_224_S_detect"
    case, sd, bm, x, K, _ = pu.build_inputs(name)
    gold = pu.load_golden(name)
    m = pu.build_engine(case, sd, bm)
    persons = m(x, K=K, det_thresh=0.3, nms_kernel_size=3)
    ev = metrics.Evaluator()
    # the reference evaluates image by image (batch size 1 in train.py:346); here all persons of the batch at once
    gt = {"j2d": gold["j2d"].to(cuda_device), "v3d": gold["v3d"].to(cuda_device),
          "transl_pelvis": gold["transl_pelvis"].to(cuda_device)}
    pairs = ev.update(persons, gt)
    s = ev.summary()
    print(s)
    assert pairs.shape[0] == gold["j2d"].shape[0] and s["recall"] == 100.0 and s["precision"] == 100.0
    assert s["pve"] < 1.0 and s["pa_pve"] < 1.0
