"""Restatement of the three `roma` functions the hot path calls (roma is a pip dependency, unpinned in
reference requirements.txt:5, absent from /root/reference and from this image; published algorithm of
naver/roma `roma/mappings.py`).  Call sites: utils/humans.py:21 (special_gramschmidt), model.py:291
(rotmat_to_rotvec), blocks/smpl_layer.py:107 (rotvec_to_rotmat).  TEST INFRASTRUCTURE ONLY."""
import torch


def special_gramschmidt(M: torch.Tensor, epsilon: float = 0.0) -> torch.Tensor:
    """[...,3,2] -> [...,3,3]: orthonormalise the two columns, third = cross product (columns stacked)."""
    shape = M.shape[:-2]
    M = M.reshape(-1, 3, 2)
    a, b = M[:, :, 0], M[:, :, 1]
    e1 = a / torch.clamp_min(torch.norm(a, dim=-1, keepdim=True), epsilon)
    b = b - torch.sum(e1 * b, dim=-1, keepdim=True) * e1
    e2 = b / torch.clamp_min(torch.norm(b, dim=-1, keepdim=True), epsilon)
    e3 = torch.cross(e1, e2, dim=-1)
    return torch.stack((e1, e2, e3), dim=-1).reshape(*shape, 3, 3)


def rotvec_to_rotmat(rotvec: torch.Tensor, epsilon: float = 1e-6) -> torch.Tensor:
    """Rodrigues formula; first-order expansion below `epsilon` rad."""
    shape = rotvec.shape[:-1]
    rv = rotvec.reshape(-1, 3)
    theta = torch.norm(rv, dim=-1)
    small = theta < epsilon
    axis = rv / torch.clamp_min(theta[:, None], epsilon)
    kx, ky, kz = axis[:, 0], axis[:, 1], axis[:, 2]
    s, c = torch.sin(theta), torch.cos(theta)
    omc = 1 - c
    xs, ys, zs = kx * s, ky * s, kz * s
    xyc, xzc, yzc = kx * ky * omc, kx * kz * omc, ky * kz * omc
    xxc, yyc, zzc = kx**2 * omc, ky**2 * omc, kz**2 * omc
    R = torch.stack([1 - yyc - zzc, xyc - zs, xzc + ys,
                     xyc + zs, 1 - xxc - zzc, -xs + yzc,
                     xzc - ys, xs + yzc, 1 - xxc - yyc], dim=-1).reshape(-1, 3, 3)
    x, y, z = rv[:, 0], rv[:, 1], rv[:, 2]
    one = torch.ones_like(x)
    R1 = torch.stack([one, -z, y, z, one, -x, -y, x, one], dim=-1).reshape(-1, 3, 3)
    R = torch.where(small[:, None, None], R1, R)
    return R.reshape(*shape, 3, 3)


def rotmat_to_unitquat(R: torch.Tensor) -> torch.Tensor:
    """Largest-of(diagonal, trace) branch selection (scipy-style), XYZW, normalised."""
    shape = R.shape[:-2]
    m = R.reshape(-1, 3, 3)
    n = m.shape[0]
    dec = torch.empty((n, 4), dtype=m.dtype, device=m.device)
    dec[:, :3] = m.diagonal(dim1=1, dim2=2)
    dec[:, 3] = dec[:, :3].sum(dim=1)
    choice = dec.argmax(dim=1)
    q = torch.empty((n, 4), dtype=m.dtype, device=m.device)

    ind = torch.nonzero(choice != 3, as_tuple=True)[0]
    i = choice[ind]
    j = (i + 1) % 3
    k = (j + 1) % 3
    q[ind, i] = 1 - dec[ind, 3] + 2 * m[ind, i, i]
    q[ind, j] = m[ind, j, i] + m[ind, i, j]
    q[ind, k] = m[ind, k, i] + m[ind, i, k]
    q[ind, 3] = m[ind, k, j] - m[ind, j, k]

    ind = torch.nonzero(choice == 3, as_tuple=True)[0]
    q[ind, 0] = m[ind, 2, 1] - m[ind, 1, 2]
    q[ind, 1] = m[ind, 0, 2] - m[ind, 2, 0]
    q[ind, 2] = m[ind, 1, 0] - m[ind, 0, 1]
    q[ind, 3] = 1 + dec[ind, 3]

    q = q / torch.norm(q, dim=1)[:, None]
    return q.reshape(*shape, 4)


def unitquat_to_rotvec(quat: torch.Tensor) -> torch.Tensor:
    shape = quat.shape[:-1]
    q = quat.reshape(-1, 4).clone()
    q[q[:, 3] < 0] *= -1  # shortest arc: w >= 0
    half = torch.atan2(torch.norm(q[:, :3], dim=1), q[:, 3])
    angle = 2 * half
    small = angle.abs() <= 1e-3
    scale = torch.where(small, 2 + angle**2 / 12 + 7 * angle**4 / 2880,
                        angle / torch.sin(torch.where(small, torch.ones_like(angle), angle) / 2))
    return (scale[:, None] * q[:, :3]).reshape(*shape, 3)


def rotmat_to_rotvec(R: torch.Tensor) -> torch.Tensor:
    return unitquat_to_rotvec(rotmat_to_unitquat(R))


def special_procrustes(M: torch.Tensor):
    """roma.special_procrustes(M, return_singular_values=True): the rotation closest to M (Frobenius), through the
    SVD with the reflection fix on the last singular direction.  Returns (R, signed singular values)."""
    U, D, Vh = torch.linalg.svd(M)
    det = torch.det(U) * torch.det(Vh)
    sign = torch.ones_like(D)
    sign[..., -1] = det
    R = (U * sign[..., None, :]) @ Vh
    return R, D * sign


def rigid_points_registration(x: torch.Tensor, y: torch.Tensor, compute_scaling: bool = False):
    """roma.rigid_points_registration (utils/rigid_points_registration of naver/roma; call sites train.py:391,424):
    the similarity (R, t, s) minimising sum ||s R x_i + t - y_i||^2 (Kabsch / Umeyama).  x, y: [..., n, 3]."""
    xmean, ymean = x.mean(dim=-2, keepdim=True), y.mean(dim=-2, keepdim=True)
    xhat, yhat = x - xmean, y - ymean
    M = torch.einsum("...ki,...kj->...ij", yhat, xhat)
    R, DS = special_procrustes(M)
    if compute_scaling:
        scale = DS.sum(dim=-1) / (xhat**2).sum(dim=(-1, -2))
        t = ymean.squeeze(-2) - scale[..., None] * (R @ xmean.squeeze(-2)[..., None]).squeeze(-1)
        return R, t, scale
    t = ymean.squeeze(-2) - (R @ xmean.squeeze(-2)[..., None]).squeeze(-1)
    return R, t
