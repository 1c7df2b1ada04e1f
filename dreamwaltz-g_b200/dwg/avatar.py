"""R4-R9: the per-step avatar path (DreamWaltzG.animate and what it calls) on the device.

Mirrors reference core/system/avatar.py:1097-1635 for the shipped configuration of
scripts/train_wo_expr.sh (use_non_rigid_offsets/scales, learn_scale=False,
use_non_rigid_rotations=False, learn_quaternions=True, frozen LBS weights, mesh-bound hands).
State-dict key names follow the reference so that checkpoints map 1:1:
  _positions, _scales, _quaternions, _lbs_weights, nerf_encoder.embeddings,
  nerf_opacity_and_color_net.net.{i}.{weight,bias},
  nerf_scale_and_quaternion_net.{layers.{i},gaussian_warp,gaussian_rotation,gaussian_scaling}.{weight,bias},
  mesh_binding_gaussians.hands.{_bary_coords,_vertex_coords,_scales}
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lbs as dlbs
from . import ops


@dataclass
class GaussianOutput:
    """Field names of reference core/gaussian/gaussian_utils.py:20-33."""
    positions: Optional[torch.Tensor] = None
    sh_features: Optional[torch.Tensor] = None
    opacities: Optional[torch.Tensor] = None
    quaternions: Optional[torch.Tensor] = None
    scales: Optional[torch.Tensor] = None
    colors: Optional[torch.Tensor] = None
    cov3D: Optional[torch.Tensor] = None
    offsets: Optional[torch.Tensor] = None
    lbs_weights: Optional[torch.Tensor] = None


class MLP(nn.Module):
    """Parameters of core/nerf/nerf_model.py:12-33 (state-dict names net.{i}.{weight,bias}).  The
    arithmetic runs in the fused kernel dwg_avatar_mlp_fwd/bwd (ops.avatar_mlp); there is no torch path."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        self.num_layers = num_layers
        self.net = nn.ModuleList([nn.Linear(dim_in if l == 0 else dim_hidden,
                                            dim_out if l == num_layers - 1 else dim_hidden, bias=bias)
                                  for l in range(num_layers)])


class DeformNetwork(nn.Module):
    """Parameters of core/deformation/deform_model.py:61-143 (D=4, W=64, no skip, is_6dof=False); see MLP."""

    def __init__(self, xyz_input_ch=32, pose_input_ch=63, D=4, W=64):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(xyz_input_ch + pose_input_ch, W)] + [nn.Linear(W, W) for _ in range(D - 1)])
        self.gaussian_warp = nn.Linear(W, 3)
        self.gaussian_rotation = nn.Linear(W, 4)
        self.gaussian_scaling = nn.Linear(W, 3)


class GridEncoder(nn.Module):
    """core/nerf/gridencoder/grid.py:99-166 (tiled grid, smoothstep) on the dwg kernel."""

    def __init__(self, device='cuda', bound=2.0, **kw):
        super().__init__()
        self.spec = ops.GridSpec(device, bound=bound, **kw)
        self.embeddings = nn.Parameter(torch.empty(self.spec.n_rows, 2, device=device).uniform_(-1e-4, 1e-4))
        self.output_dim = self.spec.num_levels * 2

    def forward(self, inputs, bound=None):
        return ops.grid_encode(inputs.reshape(-1, 3), self.embeddings, self.spec).view(*inputs.shape[:-1], self.output_dim)


class MeshBindingGaussianModel(nn.Module):
    """core/system/avatar.py:921-1094."""

    def __init__(self, mesh: dict, n_per_triangle=6, device='cuda', lbs_model=None):
        super().__init__()
        self.n_per_triangle = n_per_triangle
        self.register_buffer('predefined_vertex_indices', mesh['predefined_vertex_indices'].to(device))
        self.register_buffer('triangles', mesh['triangles'].to(device))
        Fn = self.triangles.shape[0]
        p2t = torch.arange(Fn, device=device)[..., None].expand(-1, n_per_triangle).reshape(-1)
        self.register_buffer('points_to_vertices', self.triangles[p2t])
        self._bary_coords = nn.Parameter(mesh['_bary_coords'].clone().to(device))
        self._vertex_coords = nn.Parameter(mesh['_vertex_coords'].clone().to(device), requires_grad=False)
        self._scales = nn.Parameter(mesh['_scales'].clone().to(device))
        # constants as device buffers (no per-step host->device copies; CUDA-graph capturable)
        self.register_buffer('_zaxis', torch.tensor([0.0, 0.0, 1.0], device=device))
        self.register_buffer('_xaxis', torch.tensor([1.0, 0.0, 0.0], device=device))
        self.register_buffer('_flip', torch.tensor([1.0, -1.0, -1.0], device=device).view(1, 3, 1))
        self._build_tables(lbs_model)

    def _build_tables(self, lbs_model=None):
        """Static tables of the kernel path (ops.glbs_vertices / ops.mesh_gaussians); rebuilt after a checkpoint changed the mesh."""
        device = self.triangles.device
        tri = self.triangles.cpu()
        Vp = int(self.predefined_vertex_indices.shape[0])
        order = torch.argsort(tri.reshape(-1), stable=True)                  # incidences grouped by vertex, triangle order kept
        counts = torch.bincount(tri.reshape(-1), minlength=Vp)
        nb = lambda n, t: self.register_buffer(n, t.to(device), persistent=False)
        nb('_tri_i32', tri.to(torch.int32).contiguous())
        nb('_adj_ptr', torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts, 0)]).to(torch.int32))
        nb('_adj_tri', (order // 3).to(torch.int32))
        self._has_dirs = lbs_model is not None
        if lbs_model is not None:
            vi = self.predefined_vertex_indices
            ns = lbs_model._shapedirs_full.shape[1]
            V = lbs_model.v_template.shape[0]
            nb('_sdirs_sel', lbs_model._shapedirs_full.view(V, 3, ns)[vi].contiguous())                     # [Vp,3,400]
            nb('_pdirs_sel', lbs_model.posedirs.data.view(-1, V, 3)[:, vi].permute(1, 2, 0).contiguous())   # [Vp,3,486]
            nb('_w_sel', lbs_model.lbs_weights.data[vi].contiguous())                                       # [Vp,55]

    def posed_vertex_coords(self, jt):
        """The part's predefined vertices under the GLBS vertex transform of joint state ``jt`` (ops.glbs_joints)."""
        return ops.glbs_vertices(jt, self._sdirs_sel, self._pdirs_sel, self._w_sel, self._vertex_coords)

    def gaussians(self, vertex_coords):
        """(positions, scales, quaternions) = get_positions + get_scales_and_quaternions as one kernel each way."""
        return ops.mesh_gaussians(self._bary_coords, self._scales, vertex_coords, self._tri_i32, self._adj_ptr, self._adj_tri, self.n_per_triangle)

    def get_positions(self, vertex_coords):
        bary = self._bary_coords / self._bary_coords.sum(dim=-1, keepdim=True)
        return torch.einsum('fnv,fvc->fnc', bary, vertex_coords[self.triangles]).reshape(-1, 3)

    def get_scales_and_quaternions(self, vertex_coords, positions, eps=1e-9):
        dot = lambda a, b: (a * b).sum(-1, keepdim=True)
        nrm = lambda v: torch.linalg.vector_norm(v, dim=-1, keepdim=True)
        pv = vertex_coords[self.points_to_vertices]
        p0, p1, p2, p3 = positions, pv[:, 0], pv[:, 1], pv[:, 2]
        # vertex normals (utils/mesh.py:34-97)
        i0, i1, i2 = self.triangles[:, 0], self.triangles[:, 1], self.triangles[:, 2]
        fn = torch.cross(vertex_coords[i1] - vertex_coords[i0], vertex_coords[i2] - vertex_coords[i0], dim=-1)
        fn = fn / torch.sqrt(torch.clamp((fn * fn).sum(-1, keepdim=True), min=1e-20))
        vn = torch.zeros_like(vertex_coords).index_add(0, i0, fn).index_add(0, i1, fn).index_add(0, i2, fn)
        vn = torch.where(dot(vn, vn) > 1e-20, vn, self._zaxis)
        vn = vn / torch.sqrt(torch.clamp(dot(vn, vn), min=1e-20))
        pn = (vn[self.points_to_vertices] * self._bary_coords.reshape(-1, 3)[:, :, None]).sum(dim=1)
        v0 = pn / (nrm(pn) + eps)
        ref = self._xaxis.expand_as(p0)
        v1 = torch.cross(v0, ref, dim=1)
        v1 = v1 / (nrm(v1) + eps)
        v2 = torch.cross(v0, v1, dim=1)
        v2 = v2 / (nrm(v2) + eps)
        R = torch.stack((v0, v1, v2), dim=2) * self._flip
        s1 = (dot(p1 - p0, v1).abs() + dot(p2 - p0, v1).abs() + dot(p3 - p0, v1).abs()) / self.n_per_triangle
        s2 = (dot(p1 - p0, v2).abs() + dot(p2 - p0, v2).abs() + dot(p3 - p0, v2).abs()) / self.n_per_triangle
        s1 = s1 * torch.clamp(self._scales[:, 1:2], min=0.5, max=2.0)
        s2 = s2 * torch.clamp(self._scales[:, 2:3], min=0.5, max=2.0)
        scales = torch.cat((torch.zeros_like(s1), s1, s2), dim=1)
        return scales, dlbs.standardize_quaternion(dlbs.matrix_to_quaternion(R))


class DreamWaltzGAvatar(nn.Module):
    """The animate() path of reference DreamWaltzG (avatar.py:1500-1588) on dwg kernels."""

    def __init__(self, body_model: dict, avatar: dict, device='cuda', nerf_bound=2.0,
                 init_offset=0.01, init_scale=1e-3, max_scale=0.01, learn_hand_betas=False, learn_face_betas=False):
        super().__init__()
        self.device = device
        self.lbs_model = dlbs.GeneralLinearBlendSkinning(body_model, device=device)
        self._positions = nn.Parameter(avatar['_positions'].clone().to(device))
        self._scales = nn.Parameter(avatar['_scales'].clone().to(device))
        self._quaternions = nn.Parameter(avatar['_quaternions'].clone().to(device))
        self._lbs_weights = nn.Parameter(avatar['_lbs_weights'].clone().to(device), requires_grad=False)
        self.nerf_bound = nerf_bound
        self.nerf_encoder = GridEncoder(device=device, bound=nerf_bound)
        self.nerf_opacity_and_color_net = MLP(32, 4, 64, 3).to(device)
        self.nerf_scale_and_quaternion_net = DeformNetwork(32, 63, 4, 64).to(device)
        self.mesh_binding_gaussians = nn.ModuleDict()
        meshes = avatar.get('meshes') or ({'hands': avatar['mesh']} if avatar.get('mesh') is not None else {})
        for part, mesh in meshes.items():                       # predefined_body_parts: 'hands' [, 'face'] (scripts/train_w_expr.sh:9)
            self.mesh_binding_gaussians[part] = MeshBindingGaussianModel(mesh, device=device, lbs_model=self.lbs_model)
        # avatar.py:1222-1225: optional learnable shape offset used by the mesh-bound parts
        self.learn_hand_betas, self.learn_face_betas = learn_hand_betas, learn_face_betas
        self.learn_betas = learn_hand_betas or learn_face_betas
        self._betas = nn.Parameter(self.lbs_model.betas.data.clone(), requires_grad=self.learn_betas)
        self.init_offset, self.init_scale, self.max_scale = init_offset, init_scale, max_scale
        self.smpl_canonical_inputs = {}
        self._canonical_cache = None
        self.use_kernels = True          # False: GLBS module + torch mesh ops (kept for parity tests and the learn_betas case)

    def get_lbs_weights(self):
        return self._lbs_weights / self._lbs_weights.sum(dim=-1, keepdim=True)          # avatar.py:914-917

    def _joint_pose_transform(self, transforms):
        return dlbs.RigidTransform.compose(transforms['J_pose_rigid'], transforms['G_transl_offset']).squeeze(0)

    def _mlp_params(self):
        named = dict(self.named_parameters())
        return [named[n] for n in ops.AVATAR_MLP_PARAMS]

    def animate(self, smpl_observed_inputs: Optional[dict] = None) -> GaussianOutput:
        if smpl_observed_inputs is None:
            smpl_observed_inputs = self.smpl_canonical_inputs
        if self.learn_betas or not self.use_kernels:
            return self._animate_torch(smpl_observed_inputs)     # gradient w.r.t. the shape offset: autograd through the GLBS module
        # ---- R1: joint state of the canonical (cached: constant while the shape is frozen; the reference recomputes it
        # every step, avatar.py:1508) and the observed pose -- ONE kernel each (dwg_glbs_joints)
        if self._canonical_cache is None:
            cj = self.lbs_model.joint_transforms(**self.smpl_canonical_inputs)
            self._canonical_cache = (cj, [gm.posed_vertex_coords(cj) for gm in self.mesh_binding_gaussians.values()])
        cnl_j, cnl_vcs = self._canonical_cache
        obs_j = self.lbs_model.joint_transforms(**smpl_observed_inputs)
        positions = self._positions
        W = self.get_lbs_weights()
        n_unc = positions.shape[0]
        # canonical positions of every Gaussian (unconstrained first, then the mesh-bound parts): ONE grid encode and ONE
        # fused MLP launch serve all of them (avatar.py:1519-1535,1567-1572)
        canon = [ops.lbs_skin(W, cnl_j['A_t'], positions)]
        for gm, cvc in zip(self.mesh_binding_gaussians.values(), cnl_vcs):
            canon.append(gm.gaussians(cvc)[0])
        enc = self.nerf_encoder(torch.cat(canon, dim=0) if len(canon) > 1 else canon[0], bound=self.nerf_bound)
        body_pose = smpl_observed_inputs.get('body_pose')
        if body_pose is None:
            body_pose = self.lbs_model.body_pose
        colors, opacities, pos, scales = ops.avatar_mlp(enc, positions, body_pose, self._mlp_params(), n_unc,
                                                        self.init_offset, self.init_scale, self.max_scale)
        quats = F.normalize(self._quaternions)
        pos, quats = ops.lbs_skin(W, obs_j['A_t'], pos, quats)
        if len(self.mesh_binding_gaussians) == 0:
            return GaussianOutput(positions=pos, opacities=opacities, colors=colors, quaternions=quats, scales=scales)
        all_pos, all_q, all_sc = [pos], [quats], [scales]
        for gm in self.mesh_binding_gaussians.values():         # R5: two kernels per part (vertices, Gaussians)
            m_pos, m_sc, m_q = gm.gaussians(gm.posed_vertex_coords(obs_j))
            all_pos.append(m_pos); all_q.append(m_q); all_sc.append(m_sc)
        return GaussianOutput(positions=torch.cat(all_pos, dim=0), opacities=opacities, colors=colors,
                              quaternions=torch.cat(all_q, dim=0), scales=torch.cat(all_sc, dim=0))

    def _animate_torch(self, smpl_observed_inputs):
        """The same path on the torch GLBS module / torch mesh ops (autograd reaches the shape offset `_betas`)."""
        with torch.no_grad():
            _, cnl_V, cnl_tr = self.lbs_model.forward(**self.smpl_canonical_inputs)
            _, obs_V, obs_tr = self.lbs_model.forward(**smpl_observed_inputs)
        positions = self._positions
        W = self.get_lbs_weights()
        cnl_jt = self._joint_pose_transform(cnl_tr)
        obs_jt = self._joint_pose_transform(obs_tr)
        n_unc = positions.shape[0]
        sq = lambda t: t.squeeze(0) if t.SE3.dim() == 4 else t
        cnl_T, obs_T = sq(cnl_V), sq(obs_V)
        cnl_Tb, obs_Tb = cnl_T, obs_T
        if self.learn_betas:                                    # avatar.py:1551-1553
            _, cVb, _ = self.lbs_model.forward(**self.smpl_canonical_inputs, extra_betas=self._betas)
            _, oVb, _ = self.lbs_model.forward(**smpl_observed_inputs, extra_betas=self._betas)
            cnl_Tb, obs_Tb = sq(cVb), sq(oVb)
        with_betas = lambda part: (part == 'hands' and self.learn_hand_betas) or (part == 'face' and self.learn_face_betas)
        canon = [cnl_jt.transform_points(positions, weights=W)]
        for part, gm in self.mesh_binding_gaussians.items():
            cnl_vc = (cnl_Tb if with_betas(part) else cnl_T).transform_points(gm._vertex_coords, indices=gm.predefined_vertex_indices)
            canon.append(gm.get_positions(cnl_vc))
        enc = self.nerf_encoder(torch.cat(canon, dim=0) if len(canon) > 1 else canon[0], bound=self.nerf_bound)
        body_pose = smpl_observed_inputs.get('body_pose')
        if body_pose is None:
            body_pose = torch.zeros(1, 63, device=self.device)
        colors, opacities, pos, scales = ops.avatar_mlp(enc, positions, body_pose, self._mlp_params(), n_unc,
                                                        self.init_offset, self.init_scale, self.max_scale)
        quats = F.normalize(self._quaternions)
        pos, quats = obs_jt.transform_points_and_quaternions(pos, quats, W)
        if len(self.mesh_binding_gaussians) == 0:
            return GaussianOutput(positions=pos, opacities=opacities, colors=colors, quaternions=quats, scales=scales)
        all_pos, all_q, all_sc = [pos], [quats], [scales]
        for part, gm in self.mesh_binding_gaussians.items():
            obs_vc = (obs_Tb if with_betas(part) else obs_T).transform_points(gm._vertex_coords, indices=gm.predefined_vertex_indices)
            m_pos = gm.get_positions(obs_vc)
            m_sc, m_q = gm.get_scales_and_quaternions(obs_vc, m_pos)
            all_pos.append(m_pos); all_q.append(m_q); all_sc.append(m_sc)
        return GaussianOutput(positions=torch.cat(all_pos, dim=0), opacities=opacities, colors=colors,
                              quaternions=torch.cat(all_q, dim=0), scales=torch.cat(all_sc, dim=0))


class GaussianRenderer:
    """core/gaussian/gaussian_renderer.py:9-224 (the wrapper around the rasteriser)."""

    def __init__(self, sh_levels=4, bg_color=(0.0, 0.0, 0.0)):
        self.sh_levels = sh_levels
        self.bg_color = torch.tensor(bg_color)

    def render(self, data: dict, gaussians: GaussianOutput, return_2d_radii: bool = False, cam_dev=None, bg_image=None) -> dict:
        """bg_image [1,H,W,3] / [3,H,W] (optional): composited in the blend epilogue, outputs then carry image_fg too."""
        from .camera import raster_matrices
        view, proj, campos, tanfovx, tanfovy = raster_matrices(data)         # host tensors: no device sync
        means3D = gaussians.positions
        screenspace_points = torch.zeros(means3D.shape[0], 3, dtype=means3D.dtype, requires_grad=True, device=means3D.device)
        colors = gaussians.colors
        if colors is None:
            colors = ops.sh_colors(gaussians.sh_features, means3D, campos.to(means3D.device), self.sh_levels)
        H, W = data['image_height'], data['image_width']
        if bg_image is not None and bg_image.shape[-1] == 3:                 # reference layout [1,H,W,3] -> planar
            bg_image = bg_image.reshape(H, W, 3).permute(2, 0, 1).contiguous()
        res = ops.rasterize(
            means3D, screenspace_points, colors, gaussians.opacities, gaussians.scales, gaussians.quaternions,
            image_height=H, image_width=W, tanfovx=tanfovx, tanfovy=tanfovy,
            viewmatrix=view, projmatrix=proj, bg=self.bg_color, cam_dev=cam_dev, bg_image=bg_image)
        img, radii, depth, alpha = res[:4]
        out = {'image': img.permute(1, 2, 0).unsqueeze(0), 'depth': depth.permute(1, 2, 0).unsqueeze(0),
               'alpha': alpha.permute(1, 2, 0).unsqueeze(0), 'image_chw': img}
        if bg_image is not None:
            out['image_fg'] = res[4].permute(1, 2, 0).unsqueeze(0)
        if return_2d_radii:
            out['radii'] = radii
            out['viewspace_points'] = screenspace_points
        return out
