"""Restatement of the `smplx` body model as configured by the reference at blocks/smpl_layer.py:38
(`smplx.create(SMPLX_DIR, 'smplx', gender='neutral', use_pca=False, flat_hand_mean=True, num_betas=...)`)
and called at blocks/smpl_layer.py:104.  `smplx` is a pip dependency (requirements.txt:6, unpinned),
absent from /root/reference and from this image; this follows the published vchoutas/smplx
`body_models.SMPLX.forward` + `lbs.lbs` algorithm:
  v_shaped = v_template + shapedirs . [betas|expression];  J = J_regressor . v_shaped
  R_j = batch_rodrigues(full_pose_j) (angle = ||r + 1e-8||);  pose_feature = (R_1.. - I)
  v_posed = v_shaped + pose_feature . posedirs;  kinematic chain -> A_j;  verts = (sum_j W_vj A_j) v_posed
  joints = [55 LBS joints | 21 vertex-picked joints | 51 static face landmarks] (+ zero transl).
TEST INFRASTRUCTURE ONLY — fp32, plain torch."""
import torch

# First 55 = kinematic joints of SMPL-X; the rest are vertex-picked extras and face landmarks.  Only the
# positions of 'pelvis' (0) and 'head' (15) and len >= 127 matter to the reference
# (utils/humans.py:25-26, blocks/smpl_layer.py:41-45).
_BODY = ["pelvis", "left_hip", "right_hip", "spine1", "left_knee", "right_knee", "spine2", "left_ankle",
         "right_ankle", "spine3", "left_foot", "right_foot", "neck", "left_collar", "right_collar", "head",
         "left_shoulder", "right_shoulder", "left_elbow", "right_elbow", "left_wrist", "right_wrist", "jaw",
         "left_eye_smplhf", "right_eye_smplhf"]
_FINGERS = ["index", "middle", "pinky", "ring", "thumb"]
_HANDS = [f"{s}_{f}{i}" for s in ("left", "right") for f in _FINGERS for i in (1, 2, 3)]
_EXTRA = ["nose", "right_eye", "left_eye", "right_ear", "left_ear", "left_big_toe", "left_small_toe",
          "left_heel", "right_big_toe", "right_small_toe", "right_heel"] + \
         [f"{s}_{f}" for s in ("left", "right") for f in ("thumb", "index", "middle", "ring", "pinky")]
JOINT_NAMES = _BODY + _HANDS + _EXTRA + [f"face_landmark_{i}" for i in range(51)] + \
              [f"face_contour_{i}" for i in range(17)]
assert len(JOINT_NAMES) == 144 and JOINT_NAMES.index("head") == 15

NUM_JOINTS = 55
NUM_BODY_JOINTS = 21
# Standard SMPL-X kinematic tree (parent of each of the 55 joints).
PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
           20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
           21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]


def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    cos = torch.cos(angle)[:, :, None]
    sin = torch.sin(angle)[:, :, None]
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    zeros = torch.zeros_like(rx)
    Kx = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return ident + sin * Kx + (1 - cos) * torch.bmm(Kx, Kx)


def batch_rigid_transform(rot_mats, joints, parents):
    """rot_mats [B,J,3,3], joints [B,J,3] -> posed joints [B,J,3], relative transforms A [B,J,4,4]."""
    B, J = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    T = torch.zeros(B, J, 4, 4, dtype=joints.dtype, device=joints.device)
    T[:, :, :3, :3] = rot_mats
    T[:, :, :3, 3:] = rel
    T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], T[:, i]))
    G = torch.stack(chain, dim=1)
    posed = G[:, :, :3, 3]
    jh = torch.nn.functional.pad(joints, [0, 0, 0, 1])
    A = G - torch.nn.functional.pad(torch.matmul(G, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, A


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    B = max(betas.shape[0], pose.shape[0])
    v_shaped = v_template + torch.einsum("bl,mkl->bmk", betas, shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)
    rot = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
    ident = torch.eye(3, dtype=betas.dtype, device=betas.device)
    pose_feature = (rot[:, 1:] - ident).reshape(B, -1)
    v_posed = torch.matmul(pose_feature, posedirs).view(B, -1, 3) + v_shaped
    J_posed, A = batch_rigid_transform(rot, J, parents)
    nj = J_regressor.shape[0]
    T = torch.matmul(lbs_weights.unsqueeze(0).expand(B, -1, -1), A.view(B, nj, 16)).view(B, -1, 4, 4)
    homo = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype, device=betas.device)], 2)
    verts = torch.matmul(T, homo.unsqueeze(-1))[:, :, :3, 0]
    return verts, J_posed


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary):
    B = vertices.shape[0]
    tri = faces[lmk_faces_idx]  # [L,3] vertex ids
    lmk_v = vertices[:, tri]    # [B,L,3,3]
    return torch.einsum("blfi,lf->bli", lmk_v, lmk_bary)


class Output:
    def __init__(self, vertices, joints):
        self.vertices, self.joints = vertices, joints


class SMPLXShim(torch.nn.Module):
    """`smplx.create(...)` stand-in built from a body-model dict (see multihmr_b200.synth.make_body_model):
    v_template [V,3], shapedirs [V,3,nb], expr_dirs [V,3,10], posedirs [486,3V], J_regressor [55,V],
    parents [55], lbs_weights [V,55], faces [F,3], lmk_faces_idx [51], lmk_bary_coords [51,3],
    extra_joints_idxs [21]."""

    def __init__(self, bm: dict, num_betas: int = 10):
        super().__init__()
        reg = self.register_buffer
        sd = bm["shapedirs"]
        if sd.shape[-1] < num_betas:  # num_betas=11 layer: one extra (kid) component
            sd = torch.cat([sd, bm["shapedirs_extra"][..., : num_betas - sd.shape[-1]]], dim=-1)
        reg("v_template", bm["v_template"].float())
        reg("shapedirs", sd[..., :num_betas].float())
        reg("expr_dirs", bm["expr_dirs"].float())
        reg("posedirs", bm["posedirs"].float())
        reg("J_regressor", bm["J_regressor"].float())
        reg("parents", bm["parents"].long())
        reg("lbs_weights", bm["lbs_weights"].float())
        reg("faces_tensor", bm["faces"].long())
        reg("lmk_faces_idx", bm["lmk_faces_idx"].long())
        reg("lmk_bary_coords", bm["lmk_bary_coords"].float())
        reg("extra_joints_idxs", bm["extra_joints_idxs"].long())
        P = torch.nn.Parameter
        self.num_betas = num_betas
        self.faces = bm["faces"].cpu().numpy()
        self.global_orient = P(torch.zeros(1, 3))
        self.expression = P(torch.zeros(1, 10))
        self.leye_pose = P(torch.zeros(1, 3))
        self.reye_pose = P(torch.zeros(1, 3))
        self.transl = P(torch.zeros(1, 3))

    def forward(self, betas, global_orient, body_pose, left_hand_pose, right_hand_pose, jaw_pose,
                expression, leye_pose, reye_pose):
        B = betas.shape[0]
        full_pose = torch.cat([global_orient.reshape(-1, 1, 3), body_pose.reshape(-1, NUM_BODY_JOINTS, 3),
                               jaw_pose.reshape(-1, 1, 3), leye_pose.reshape(-1, 1, 3),
                               reye_pose.reshape(-1, 1, 3), left_hand_pose.reshape(-1, 15, 3),
                               right_hand_pose.reshape(-1, 15, 3)], dim=1).reshape(-1, 165)
        # flat_hand_mean=True -> pose_mean == 0
        comps = torch.cat([betas, expression], dim=-1)
        dirs = torch.cat([self.shapedirs, self.expr_dirs], dim=-1)
        verts, joints = lbs(comps, full_pose, self.v_template, dirs, self.posedirs, self.J_regressor,
                            self.parents, self.lbs_weights)
        lmk = vertices2landmarks(verts, self.faces_tensor, self.lmk_faces_idx, self.lmk_bary_coords)
        joints = torch.cat([joints, verts[:, self.extra_joints_idxs], lmk], dim=1)
        joints = joints + self.transl.unsqueeze(1)
        verts = verts + self.transl.unsqueeze(1)
        return Output(verts, joints)
