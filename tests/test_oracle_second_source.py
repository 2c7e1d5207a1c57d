"""Second-source pinning of the third-party restatements in oracle/ (DESIGN.md §2): the packages the reference
imports (facebookresearch/dinov2 via torch.hub, roma, smplx) are neither vendored nor installed, so their
restatements are checked against INDEPENDENT implementations of the same published algorithms that do exist in
this image:
  * oracle.dinov2_ref  vs  HuggingFace transformers' Dinov2Model (same architecture, different code base and
    state-dict layout: separate q/k/v Linears, `layer_scale1.lambda1`, ...), on random weights mapped key by key;
  * oracle.roma_ref    vs  scipy.spatial.transform.Rotation (rotation vector <-> matrix).
CPU only."""
import numpy as np
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _hf_to_hub_state_dict(hf_sd, depth):
    """HuggingFace Dinov2Model keys -> facebookresearch/dinov2 hub keys (the ones Multi-HMR checkpoints carry
    under `backbone.encoder.`)."""
    sd = {
        "cls_token": hf_sd["embeddings.cls_token"], "pos_embed": hf_sd["embeddings.position_embeddings"],
        "patch_embed.proj.weight": hf_sd["embeddings.patch_embeddings.projection.weight"],
        "patch_embed.proj.bias": hf_sd["embeddings.patch_embeddings.projection.bias"],
        "norm.weight": hf_sd["layernorm.weight"], "norm.bias": hf_sd["layernorm.bias"],
    }
    for i in range(depth):
        h, b = f"encoder.layer.{i}.", f"blocks.{i}."
        a = h + "attention.attention."
        sd[b + "attn.qkv.weight"] = torch.cat([hf_sd[a + "query.weight"], hf_sd[a + "key.weight"], hf_sd[a + "value.weight"]])
        sd[b + "attn.qkv.bias"] = torch.cat([hf_sd[a + "query.bias"], hf_sd[a + "key.bias"], hf_sd[a + "value.bias"]])
        sd[b + "attn.proj.weight"] = hf_sd[h + "attention.output.dense.weight"]
        sd[b + "attn.proj.bias"] = hf_sd[h + "attention.output.dense.bias"]
        sd[b + "ls1.gamma"] = hf_sd[h + "layer_scale1.lambda1"]
        sd[b + "ls2.gamma"] = hf_sd[h + "layer_scale2.lambda1"]
        for n in ("norm1", "norm2"):
            sd[b + n + ".weight"], sd[b + n + ".bias"] = hf_sd[h + n + ".weight"], hf_sd[h + n + ".bias"]
        for n in ("fc1", "fc2"):
            sd[b + f"mlp.{n}.weight"], sd[b + f"mlp.{n}.bias"] = hf_sd[h + f"mlp.{n}.weight"], hf_sd[h + f"mlp.{n}.bias"]
    return sd


@pytest.mark.parametrize("name,dim,heads", [("dinov2_vits14", 384, 6), ("dinov2_vitb14", 768, 12)])
def test_dinov2_restatement_matches_transformers(name, dim, heads, monkeypatch):
    from transformers import Dinov2Config, Dinov2Model

    from oracle import dinov2_ref

    depth = 3  # the block is the same at every depth; three keep the CPU test short
    monkeypatch.setitem(dinov2_ref.ARCHS, name, dict(embed_dim=dim, depth=depth, num_heads=heads))
    torch.manual_seed(0)
    cfg = Dinov2Config(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads, mlp_ratio=4,
                       image_size=518, patch_size=14, qkv_bias=True, layerscale_value=1.0, use_swiglu_ffn=False,
                       layer_norm_eps=1e-6, hidden_act="gelu", attn_implementation="eager")
    hf = Dinov2Model(cfg).eval()
    with torch.no_grad():  # non-trivial LayerScale / LayerNorm / embeddings (the defaults are ones and zeros)
        for k, p in hf.named_parameters():
            if "lambda1" in k:
                p.uniform_(0.05, 1.0)
            elif "norm" in k and k.endswith("weight"):
                p.uniform_(0.5, 1.5)
            elif k.endswith("bias") or "cls_token" in k or "position_embeddings" in k:
                p.normal_(0, 0.05)
    sd = _hf_to_hub_state_dict(hf.state_dict(), depth)
    # 518 = the pre-training resolution: 37 x 37 patches, so neither side interpolates the position embedding
    x = torch.randn(1, 3, 518, 518)
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state[:, 1:]          # final LayerNorm applied, cls dropped
        got = dinov2_ref.get_intermediate_layers(x, sd, name)
    assert got.shape == ref.shape == (1, 37 * 37, dim)
    err = (got - ref).abs().max().item()
    assert err <= 2e-4 * max(1.0, ref.abs().max().item()), err


def test_dinov2_pos_embed_interpolation_identity_at_native_grid():
    from oracle import dinov2_ref

    pos = torch.randn(1, 1 + 37 * 37, 16)
    assert torch.equal(dinov2_ref.interpolate_pos_encoding(pos, 518, 518), pos)
    out = dinov2_ref.interpolate_pos_encoding(pos, 224, 224)
    assert out.shape == (1, 1 + 16 * 16, 16) and torch.equal(out[:, 0], pos[:, 0])


def test_roma_restatement_matches_scipy():
    from scipy.spatial.transform import Rotation

    from oracle import roma_ref

    rng = np.random.default_rng(0)
    rv = rng.normal(size=(512, 3))
    rv = rv / np.linalg.norm(rv, axis=1, keepdims=True) * rng.uniform(1e-4, np.pi - 1e-3, size=(512, 1))
    R_ref = Rotation.from_rotvec(rv).as_matrix()
    R = roma_ref.rotvec_to_rotmat(torch.from_numpy(rv)).numpy()
    assert np.abs(R - R_ref).max() < 1e-9
    back = roma_ref.rotmat_to_rotvec(torch.from_numpy(R_ref)).numpy()
    assert np.abs(back - Rotation.from_matrix(R_ref).as_rotvec()).max() < 1e-7
    # Gram-Schmidt of (a, b) equals the QR-based orthonormal frame with positive diagonal and det +1
    M = rng.normal(size=(128, 3, 2))
    Q = roma_ref.special_gramschmidt(torch.from_numpy(M)).numpy()
    for m, q in zip(M, Q):
        qq, rr = np.linalg.qr(m)
        qq = qq * np.sign(np.diag(rr))
        assert np.abs(q[:, :2] - qq).max() < 1e-9
        assert np.abs(q[:, 2] - np.cross(qq[:, 0], qq[:, 1])).max() < 1e-9


def _naive_lbs(betas, pose, bm):
    """Linear blend skinning written from the SMPL definition (Loper et al. 2015, eq. 2-4) in float64 numpy with
    explicit per-joint world transforms and a per-vertex weighted sum -- a structurally different evaluation of
    the same formula as oracle.smplx_ref.lbs (which follows smplx.lbs: relative transforms, batched matmuls)."""
    from scipy.spatial.transform import Rotation

    vt = bm["v_template"].double().numpy()
    sdirs = np.concatenate([bm["shapedirs"].double().numpy(), bm["expr_dirs"].double().numpy()], axis=2)
    v_shaped = vt + sdirs @ betas
    J = bm["J_regressor"].double().numpy() @ v_shaped
    R = Rotation.from_rotvec(pose.reshape(-1, 3)).as_matrix()
    feat = (R[1:] - np.eye(3)).reshape(-1)
    v_posed = v_shaped + (feat @ bm["posedirs"].double().numpy()).reshape(-1, 3)
    parents = bm["parents"].numpy()
    world_R, world_t = [R[0]], [J[0]]
    for j in range(1, len(parents)):
        p = parents[j]
        world_R.append(world_R[p] @ R[j])
        world_t.append(world_R[p] @ (J[j] - J[p]) + world_t[p])
    W = bm["lbs_weights"].double().numpy()
    verts = np.zeros_like(v_posed)
    for j in range(len(parents)):  # vertex in the rest frame of joint j, moved by the joint's world transform
        verts += W[:, j:j + 1] * ((v_posed - J[j]) @ world_R[j].T + world_t[j])
    return verts, np.stack(world_t)


def test_smplx_lbs_restatement_matches_naive_float64_evaluation():
    from multihmr_b200 import synth
    from oracle import smplx_ref

    bm = synth.make_body_model(3)
    rng = np.random.default_rng(1)
    betas = rng.normal(size=20)
    pose = rng.normal(size=(55, 3)) * 0.4
    shapedirs = torch.cat([bm["shapedirs"], bm["expr_dirs"]], dim=2).double()
    v, j = smplx_ref.lbs(torch.from_numpy(betas)[None], torch.from_numpy(pose).reshape(1, -1), bm["v_template"].double(),
                         shapedirs, bm["posedirs"].double(), bm["J_regressor"].double(), bm["parents"],
                         bm["lbs_weights"].double())
    v_ref, j_ref = _naive_lbs(betas, pose, bm)
    assert np.abs(v[0].numpy() - v_ref).max() < 1e-7   # (batch_rodrigues adds 1e-8 to the vector before the norm)
    assert np.abs(j[0].numpy() - j_ref).max() < 1e-7


def test_smplx_lbs_root_rotation_is_rigid():
    from scipy.spatial.transform import Rotation

    from multihmr_b200 import synth
    from oracle import smplx_ref

    bm = synth.make_body_model(4)
    betas = torch.zeros(1, 20, dtype=torch.float64)
    pose = torch.zeros(1, 55 * 3, dtype=torch.float64)
    args = (bm["v_template"].double(), torch.cat([bm["shapedirs"], bm["expr_dirs"]], dim=2).double(),
            bm["posedirs"].double(), bm["J_regressor"].double(), bm["parents"], bm["lbs_weights"].double())
    v0, j0 = smplx_ref.lbs(betas, pose, *args)
    rv = np.array([0.3, -1.1, 0.7])
    pose[0, :3] = torch.from_numpy(rv)
    v1, j1 = smplx_ref.lbs(betas, pose, *args)
    R = Rotation.from_rotvec(rv).as_matrix()
    root = j0[0, 0].numpy()
    assert np.abs(v1[0].numpy() - ((v0[0].numpy() - root) @ R.T + root)).max() < 1e-6
    assert np.abs(j1[0].numpy() - ((j0[0].numpy() - root) @ R.T + root)).max() < 1e-6
