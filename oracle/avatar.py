"""Oracle: DreamWaltzG.animate and what it calls (rows R4, R5, R7, R8, R9).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain torch on the CPU, differentiable.
Follows:
  MLP.forward                         reference core/nerf/nerf_model.py:12-33
  DeformNetwork.forward               reference core/deformation/deform_model.py:102-143
  GaussianModel activations           reference core/gaussian/gaussian_model.py:25-33
  static/dynamic_mlp_forward          reference core/system/avatar.py:1283-1294
  non_rigid_transform                 reference core/system/avatar.py:1464-1498
  MeshBindingGaussianModel            reference core/system/avatar.py:1009-1079
  compute_normal                      reference utils/mesh.py:34-97
  animate                             reference core/system/avatar.py:1500-1588
Pinned against the reference's own DeformNetwork / MLP / mesh-bound code by
tests/golden/make_golden.py.
"""
import torch
import torch.nn.functional as F

from . import lbs as olbs
from . import threep as tp


def mlp_forward(x, weights, biases):
    """nerf_model.py:28-33: Linear+ReLU chain, no activation after the last layer."""
    n = len(weights)
    for l in range(n):
        x = F.linear(x, weights[l], biases[l])
        if l != n - 1:
            x = F.relu(x)
    return x


def deform_forward(enc, body_pose, p):
    """deform_model.py:102-143 with D=4, W=64, no skip, is_6dof=False.
    p: dict layers.{i}.weight/bias, gaussian_warp/rotation/scaling.weight/bias."""
    h = torch.cat([enc, body_pose.expand(enc.shape[0], -1)], dim=-1)
    i = 0
    while f'layers.{i}.weight' in p:
        h = F.leaky_relu(F.linear(h, p[f'layers.{i}.weight'], p[f'layers.{i}.bias']))
        i += 1
    d_xyz = F.linear(h, p['gaussian_warp.weight'], p['gaussian_warp.bias'])
    d_scale = F.linear(h, p['gaussian_scaling.weight'], p['gaussian_scaling.bias'])
    d_rot = F.linear(h, p['gaussian_rotation.weight'], p['gaussian_rotation.bias'])
    return d_xyz, d_scale, d_rot


def static_heads(enc, sigma_w, sigma_b, fix_opacities=False):
    """avatar.py:1283-1290: out[:,0] -> sigmoid opacity (or 1), out[:,1:4] -> sigmoid colour."""
    o = mlp_forward(enc, sigma_w, sigma_b)
    colors = torch.sigmoid(o[:, 1:])
    opac = torch.ones_like(o[:, :1]) if fix_opacities else torch.sigmoid(o[:, :1])
    return colors, opac


def non_rigid(positions, offsets, d_scales, quats_param, init_offset=0.01, init_scale=1e-3, max_scale=0.01):
    """avatar.py:1464-1498 with the shipped flags: use_non_rigid_offsets, use_non_rigid_scales,
    learn_scale=False, use_non_rigid_rotations=False, learn_quaternions=True."""
    pos = positions + offsets * init_offset
    scales = (torch.exp(d_scales) * init_scale).clamp_max(max_scale)
    quats = F.normalize(quats_param)
    return pos, scales, quats


# ------------------------------------------------------------------------- mesh-bound part
def compute_vertex_normals(vertices, faces):
    """utils/mesh.py:34-97 (single mesh)."""
    i0, i1, i2 = faces[:, 0], faces[:, 1], faces[:, 2]
    v0, v1, v2 = vertices[i0], vertices[i1], vertices[i2]
    fn = torch.cross(v1 - v0, v2 - v0, dim=-1)
    fn = fn / torch.sqrt(torch.clamp((fn * fn).sum(-1, keepdim=True), min=1e-20))
    vn = torch.zeros_like(vertices)
    vn = vn.index_add(0, i0, fn).index_add(0, i1, fn).index_add(0, i2, fn)
    dotp = (vn * vn).sum(-1, keepdim=True)
    vn = torch.where(dotp > 1e-20, vn, torch.tensor([0.0, 0.0, 1.0], dtype=vn.dtype))
    vn = vn / torch.sqrt(torch.clamp((vn * vn).sum(-1, keepdim=True), min=1e-20))
    return vn


def mesh_positions(vertex_coords, triangles, bary_raw):
    """avatar.py:1009-1025: bary/sum(bary) then einsum('fnv,fvc->fnc')."""
    bary = bary_raw / bary_raw.sum(dim=-1, keepdim=True)
    tri = vertex_coords[triangles]                   # [F,3,3]
    return torch.einsum('fnv,fvc->fnc', bary, tri).reshape(-1, 3)


def mesh_scales_quats(vertex_coords, positions, triangles, bary_raw, scales_param, n_per_tri=6, eps=1e-9):
    """avatar.py:1027-1079.  NB the normal interpolation uses the RAW _bary_coords (:1056)."""
    Fn = triangles.shape[0]
    p2t = torch.arange(Fn).unsqueeze(-1).expand(-1, n_per_tri).reshape(-1)
    p2v = triangles[p2t]                              # [N,3]
    p0 = positions
    pv = vertex_coords[p2v]
    p1, p2, p3 = pv[:, 0], pv[:, 1], pv[:, 2]
    vn = compute_vertex_normals(vertex_coords, triangles)
    pn = (vn[p2v] * bary_raw.reshape(-1, 3)[:, :, None]).sum(dim=1)
    nrm = lambda v: torch.linalg.vector_norm(v, dim=-1, keepdim=True)
    dot = lambda a, b: (a * b).sum(-1, keepdim=True)
    v0 = pn / (nrm(pn) + eps)
    ref = torch.tensor((1.0, 0.0, 0.0), dtype=positions.dtype).expand_as(p0)
    v1 = torch.cross(v0, ref, dim=1)
    v1 = v1 / (nrm(v1) + eps)
    v2 = torch.cross(v0, v1, dim=1)
    v2 = v2 / (nrm(v2) + eps)
    R = torch.stack((v0, v1, v2), dim=2)
    R = R * torch.tensor([1.0, -1.0, -1.0], dtype=R.dtype).view(1, 3, 1)       # rows 1,2 negated (:1068)
    s0 = torch.zeros_like(v0[:, :1])
    s1 = (dot(p1 - p0, v1).abs() + dot(p2 - p0, v1).abs() + dot(p3 - p0, v1).abs()) / n_per_tri
    s2 = (dot(p1 - p0, v2).abs() + dot(p2 - p0, v2).abs() + dot(p3 - p0, v2).abs()) / n_per_tri
    s1 = s1 * torch.clamp(scales_param[:, 1:2], min=0.5, max=2.0)
    s2 = s2 * torch.clamp(scales_param[:, 2:3], min=0.5, max=2.0)
    scales = torch.cat((s0, s1, s2), dim=1)
    quats = tp.standardize_quaternion(tp.matrix_to_quaternion(R))
    return scales, quats


# ---------------------------------------------------------------------------------- animate
def animate(model, avatar, nets, grid_encode, smpl_canonical, smpl_observed, bound=2.0):
    """avatar.py:1500-1588.  ``grid_encode(x)`` is the differentiable grid encoder
    (positions [B,3] -> [B,32]); nets = {'sigma_w','sigma_b','deform'}.
    Returns dict(positions, opacities, colors, quaternions, scales) for all N Gaussians."""
    _, cnl_V, cnl_tr = olbs.glbs_forward(model, **smpl_canonical)
    _, obs_V, obs_tr = olbs.glbs_forward(model, **smpl_observed)
    positions = avatar['_positions']
    W = olbs.normalise_lbs_weights(avatar['_lbs_weights'])
    cnl_pos = olbs.lbs_transform(positions, cnl_tr, W)
    enc = grid_encode(cnl_pos)
    colors, opac = static_heads(enc, nets['sigma_w'], nets['sigma_b'])
    body_pose = smpl_observed.get('body_pose', torch.zeros(1, 63))
    d_xyz, d_scale, _ = deform_forward(enc, body_pose, nets['deform'])
    pos, scales, quats = non_rigid(positions, d_xyz, d_scale, avatar['_quaternions'])
    pos, quats = olbs.lbs_transform(pos, obs_tr, W, quaternions=quats)
    out = {'positions': pos, 'opacities': opac, 'colors': colors, 'quaternions': quats, 'scales': scales}
    meshes = avatar.get('meshes') or ({'hands': avatar['mesh']} if avatar.get('mesh') is not None else {})
    betas = avatar.get('_betas')                 # avatar.py:1551-1553: learn_hand_betas / learn_face_betas
    learn = avatar.get('learn_betas_parts', ())
    if betas is not None and learn:
        _, cnl_Vb, _ = olbs.glbs_forward(model, **smpl_canonical, extra_betas=betas)
        _, obs_Vb, _ = olbs.glbs_forward(model, **smpl_observed, extra_betas=betas)
    for part, mesh in meshes.items():
        vidx = mesh['predefined_vertex_indices']
        use_b = betas is not None and part in learn
        cnl_T, obs_T = (cnl_Vb if use_b else cnl_V).squeeze(0), (obs_Vb if use_b else obs_V).squeeze(0)
        cnl_vc = cnl_T.transform_points(mesh['_vertex_coords'], indices=vidx)
        cnl_p = mesh_positions(cnl_vc, mesh['triangles'], mesh['_bary_coords'])
        m_enc = grid_encode(cnl_p)
        m_col, m_op = static_heads(m_enc, nets['sigma_w'], nets['sigma_b'], fix_opacities=True)
        obs_vc = obs_T.transform_points(mesh['_vertex_coords'], indices=vidx)
        m_pos = mesh_positions(obs_vc, mesh['triangles'], mesh['_bary_coords'])
        m_sc, m_q = mesh_scales_quats(obs_vc, m_pos, mesh['triangles'], mesh['_bary_coords'], mesh['_scales'])
        out = {'positions': torch.cat([out['positions'], m_pos]), 'opacities': torch.cat([out['opacities'], m_op]),
               'colors': torch.cat([out['colors'], m_col]), 'quaternions': torch.cat([out['quaternions'], m_q]),
               'scales': torch.cat([out['scales'], m_sc])}
    return out
