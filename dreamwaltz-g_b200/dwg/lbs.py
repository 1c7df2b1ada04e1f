"""Surface 4: host-side mirror of the reference LBS module on the device.

  RigidTransform                 <- core/human/inverse_lbs.py:15-260 (same fields and methods)
  GeneralLinearBlendSkinning     <- core/human/inverse_lbs.py:518-784 (forward signature and the
                                    returned (transform_J, transform_V, dict) triple)

The per-Gaussian skinning (weights=... paths) runs in the fused CUDA kernel (dwg_lbs_skin_*); the
55-joint kinematics and the per-vertex blend are tiny and stay as torch device ops (plumbing).
Third-party maths (smplx.lbs.*, pytorch3d.transforms.*) is restated here because those packages
are not dependencies of this library.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
                 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]


# ---------------------------------------------------------------- small quaternion helpers
def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def matrix_to_quaternion(m):
    """pytorch3d 0.7.5 semantics (argmax candidate, no sign standardisation)."""
    b = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(b + (9,)), dim=-1)
    e = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1)
    q_abs = torch.where(e > 0, torch.sqrt(torch.clamp(e, min=1e-38)), torch.zeros_like(e))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    sel = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, sel[..., None, None].expand(b + (1, 4))).squeeze(-2)


def standardize_quaternion(q):
    return torch.where(q[..., 0:1] < 0, -q, q)


def quaternion_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return standardize_quaternion(torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                                               aw * bx + ax * bw + ay * bz - az * by,
                                               aw * by - ax * bz + ay * bw + az * bx,
                                               aw * bz + ax * by - ay * bx + az * bw), -1))


# ------------------------------------------------------------------------------ RigidTransform
class RigidTransform:
    """SE3 container with the reference's method surface (inverse_lbs.py:15-260)."""

    def __init__(self, SE3=None, R=None, T=None):
        if SE3 is None:
            if R is not None and T is not None:
                SE3 = torch.zeros(*R.shape[:-2], 4, 4, dtype=R.dtype, device=R.device)
                SE3[..., :3, :3] = R
                SE3[..., :3, 3] = T
                SE3[..., 3, 3] = 1.0
            elif R is not None:
                SE3 = torch.eye(4, dtype=R.dtype, device=R.device).expand(*R.shape[:-2], 4, 4).contiguous()
                SE3[..., :3, :3] = R
            elif T is not None:
                SE3 = torch.eye(4, dtype=T.dtype, device=T.device).expand(*T.shape[:-1], 4, 4).contiguous()
                SE3[..., :3, 3] = T
            else:
                raise NotImplementedError
        self.SE3 = SE3
        self.R = SE3[..., :3, :3]
        self.T = SE3[..., :3, 3]

    @property
    def shape(self):
        return self.SE3.shape[:-2]

    def inverse(self):
        SE3 = self.SE3
        SE3[..., 3, :] = torch.tensor([0, 0, 0, 1], dtype=SE3.dtype, device=SE3.device)    # in place, as the reference (:122)
        Rt = SE3[..., :3, :3].transpose(-1, -2)
        out = torch.zeros_like(SE3)
        out[..., :3, :3] = Rt
        out[..., :3, 3] = -torch.matmul(Rt, SE3[..., :3, 3].unsqueeze(-1)).squeeze(-1)
        out[..., 3, 3] = 1.0
        return RigidTransform(SE3=out)

    def compose(self, *others):
        SE3 = self.SE3.clone()
        for other in others:
            if not isinstance(other, RigidTransform):
                raise ValueError('Only possible to compose RigidTransform objects; got %s' % type(other))
            SE3 = other.SE3 @ SE3
        return RigidTransform(SE3=SE3)

    def index(self, indices):
        return RigidTransform(SE3=self.SE3[indices])

    def weight(self, weights, qr_correct=False):
        assert not qr_correct, 'qr_correct is unused by the reference ("Too slow!")'
        return RigidTransform(SE3=torch.einsum('nj,jkl->nkl', weights, self.SE3))

    def squeeze(self, dim=0):
        self.SE3 = self.SE3.squeeze(dim)
        self.R = self.R.squeeze(dim)
        self.T = self.T.squeeze(dim)
        return self

    def _fusable(self, pts, weights):
        return weights is not None and pts.is_cuda and self.SE3.dim() == 3 and pts.dim() == 2

    def transform_points(self, points, indices=None, weights=None):
        assert indices is None or weights is None
        if self._fusable(points, weights):
            return ops.lbs_skin(weights, self.SE3, points)             # fused CUDA path
        R, T = self.R, self.T
        if indices is not None:
            R, T = R[indices], T[indices]
        if weights is not None:
            R = torch.einsum('nj,jkl->nkl', weights, R)
            T = torch.einsum('nj,jk->nk', weights, T)
        return torch.matmul(R, points.unsqueeze(-1))[..., :, 0] + T

    def transform_points_and_quaternions(self, points, quaternions, weights):
        """Fused form of transform_points(weights) + transform_quaternions(weights,
        flip_rotation_axis=True): ONE pass over the weights (avatar.py:1450-1459)."""
        return ops.lbs_skin(weights, self.SE3, points, quaternions)

    def transform_quaternions(self, quaternions, indices=None, weights=None, rotation_mode='quaternion',
                              flip_rotation_axis=False):
        assert indices is None or weights is None
        R = self.R
        if indices is not None:
            R = self.R[indices]
        if weights is not None:
            if flip_rotation_axis and self._fusable(quaternions, weights):
                dummy = torch.zeros(quaternions.shape[0], 3, device=quaternions.device, dtype=quaternions.dtype)
                return ops.lbs_skin(weights, self.SE3, dummy, quaternions)[1]
            R = torch.einsum('nj,jkl->nkl', weights, self.R)
        if flip_rotation_axis:
            sign = torch.tensor([1.0, -1.0, -1.0], dtype=R.dtype, device=R.device).view(1, 3, 1)
            rot = quaternion_to_matrix(quaternions) * sign
            rot = (R @ rot) * sign
            return matrix_to_quaternion(rot)
        if rotation_mode == 'matrix':
            return matrix_to_quaternion(R @ quaternion_to_matrix(quaternions))
        if rotation_mode == 'quaternion':
            return quaternion_multiply(matrix_to_quaternion(R), quaternions)
        raise AssertionError(rotation_mode)

    def __repr__(self):
        return f'SE3: {self.SE3},\r\nR: {self.R},\r\nT: {self.T}'


# ------------------------------------------------------------- smplx.lbs maths on the device
def batch_rodrigues(rot_vecs):
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(d, 1, dim=1)
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def _kinematic_levels(parents):
    depth = [0] * len(parents)
    for i in range(1, len(parents)):
        depth[i] = depth[parents[i]] + 1
    levels = []
    for d in range(1, max(depth) + 1):
        idx = [i for i in range(len(parents)) if depth[i] == d]
        levels.append((idx, [parents[i] for i in idx]))
    return levels


def device_levels(parents, device):
    """Kinematic levels as device index tensors (built once: no host->device copies per step, which
    also keeps the chain capturable in a CUDA graph)."""
    return {'par': torch.as_tensor(parents[1:], device=device),
            'levels': [(torch.as_tensor(i, device=device), torch.as_tensor(p, device=device)) for i, p in _kinematic_levels(parents)]}


def batch_rigid_transform(rot_mats, joints, parents, levels=None):
    """smplx.lbs.batch_rigid_transform; the chain is evaluated level by level (tree depth 10)
    instead of 54 sequential products."""
    B, J = joints.shape[:2]
    rel = joints.clone()
    if levels is None:
        levels = device_levels(parents, joints.device)
    par = levels['par']
    rel[:, 1:] = joints[:, 1:] - joints[:, par]
    M = torch.zeros(B, J, 4, 4, dtype=joints.dtype, device=joints.device)
    M[..., :3, :3] = rot_mats
    M[..., :3, 3] = rel
    M[..., 3, 3] = 1.0
    chain = M.clone()
    for idx, pidx in levels['levels']:
        chain[:, idx] = chain[:, pidx] @ M[:, idx]
    posed = chain[..., :3, 3]
    A = chain.clone()
    A[..., :3, 3] = chain[..., :3, 3] - torch.matmul(chain[..., :3, :3], joints.unsqueeze(-1)).squeeze(-1)
    return posed, A


class GeneralLinearBlendSkinning(nn.Module):
    """Device mirror of the reference module.  Construct from a dict of SMPL-X-shaped tensors
    (same names as the reference's parameters / state-dict keys)."""

    NUM_BODY_JOINTS = 21

    def __init__(self, model: dict, device='cuda'):
        super().__init__()
        p = lambda k: nn.Parameter(model[k].detach().clone().to(device), requires_grad=False)
        self.parents = list(model['parents'])
        self._levels = device_levels(self.parents, device)
        for k in ('betas', 'v_template', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights', 'pose_mean',
                  'expr_dirs', 'expression'):
            setattr(self, k, p(k))
        z = lambda *s: nn.Parameter(torch.zeros(*s, device=device), requires_grad=False)
        self.body_pose, self.global_orient = z(1, 63), z(1, 3)
        self.left_hand_pose, self.right_hand_pose = z(1, 45), z(1, 45)
        self.jaw_pose, self.leye_pose, self.reye_pose = z(1, 3), z(1, 3), z(1, 3)
        self.use_smplx, self.use_pca = True, False
        self.register_buffer('J_template', torch.einsum('ik,ji->jk', self.v_template.data, self.J_regressor.data))
        self.register_buffer('_shapedirs_full', torch.cat([self.shapedirs.data, self.expr_dirs.data], dim=-1)
                             .reshape(-1, self.shapedirs.shape[-1] + self.expr_dirs.shape[-1]).contiguous())
        # fused joint path (ops.glbs_joints): the joint regressor is linear, so J = J_template + JS . shape with the
        # constant JS = J_regressor . shapedirs [55*3, 400] -- no full-mesh blend shapes for the joints
        V, ns = self.v_template.shape[0], self._shapedirs_full.shape[1]
        JS = torch.einsum('ji,ick->jck', self.J_regressor.data, self._shapedirs_full.view(V, 3, ns)).reshape(-1, ns).contiguous()
        self.register_buffer('_JS', JS, persistent=False)
        self.register_buffer('_parents_i32', torch.tensor([p if p >= 0 else 0 for p in self.parents], dtype=torch.int32, device=device),
                             persistent=False)

    @torch.no_grad()
    def joint_transforms(self, body_pose=None, global_orient=None, left_hand_pose=None, right_hand_pose=None, expression=None,
                         transl=None, betas=None, extra_betas=None, **_):
        """The part of forward() that DreamWaltzG.animate consumes, as ONE kernel (dwg_glbs_joints): the joint transforms
        A (= tr['J_pose_rigid']), transl o A (= _joint_pose_transform), and the pose feature / shape vector that the
        per-part vertex kernel (dwg_glbs_vertices) needs.  Same argument semantics as forward() (jaw / eye poses always
        come from the module, inverse_lbs.py:617-619)."""
        betas = self.betas if betas is None else betas
        if extra_betas is not None:
            betas = betas + extra_betas
        d = lambda v, dflt: dflt if v is None else v
        parts = (d(global_orient, self.global_orient), d(body_pose, self.body_pose), self.jaw_pose, self.leye_pose, self.reye_pose,
                 d(left_hand_pose, self.left_hand_pose), d(right_hand_pose, self.right_hand_pose))
        return ops.glbs_joints(parts, self.pose_mean, betas, d(expression, self.expression), self.J_template, self._JS, self._parents_i32,
                               transl=transl)

    def get_full_shape(self, betas=None, expression=None, extra_betas=None):
        betas = self.betas if betas is None else betas
        if extra_betas is not None:
            betas = betas + extra_betas
        expression = expression if expression is not None else self.expression
        return torch.cat([betas, expression], dim=-1)

    def get_full_pose(self, body_pose=None, global_orient=None, left_hand_pose=None, right_hand_pose=None, **_):
        global_orient = global_orient if global_orient is not None else self.global_orient
        body_pose = body_pose if body_pose is not None else self.body_pose
        left_hand_pose = left_hand_pose if left_hand_pose is not None else self.left_hand_pose
        right_hand_pose = right_hand_pose if right_hand_pose is not None else self.right_hand_pose
        full = torch.cat([global_orient.reshape(-1, 1, 3), body_pose.reshape(-1, 21, 3),
                          self.jaw_pose.reshape(-1, 1, 3), self.leye_pose.reshape(-1, 1, 3), self.reye_pose.reshape(-1, 1, 3),
                          left_hand_pose.reshape(-1, 15, 3), right_hand_pose.reshape(-1, 15, 3)], dim=1).reshape(-1, 165)
        return full + self.pose_mean

    def get_full_transform(self, betas, pose):
        B = max(betas.shape[0], pose.shape[0])
        V = self.v_template.shape[0]
        shape_offsets = (self._shapedirs_full @ betas.t()).t().reshape(betas.shape[0], V, 3)    # blend_shapes
        v_shaped = self.v_template + shape_offsets
        J = torch.einsum('bik,ji->bjk', v_shaped, self.J_regressor)                              # vertices2joints
        rot_mats = batch_rodrigues(pose.view(-1, 3)).view(B, -1, 3, 3)
        ident = torch.eye(3, dtype=pose.dtype, device=pose.device)
        pose_feature = (rot_mats[:, 1:] - ident).view(B, -1)
        pose_offsets = torch.matmul(pose_feature, self.posedirs).view(B, -1, 3)
        _, A = batch_rigid_transform(rot_mats, J, self.parents, self._levels)
        T = torch.matmul(self.lbs_weights.unsqueeze(0).expand(B, -1, -1), A.view(B, -1, 16)).view(B, -1, 4, 4)
        return {
            'V_shape_offset': RigidTransform(T=shape_offsets),
            'V_pose_offset': RigidTransform(T=pose_offsets),
            'V_pose_rigid': RigidTransform(SE3=T),
            'J_shape_offset': RigidTransform(T=J - self.J_template),
            'J_pose_rigid': RigidTransform(SE3=A),
        }

    def forward(self, betas=None, body_pose=None, global_orient=None, left_hand_pose=None, right_hand_pose=None,
                jaw_pose=None, leye_pose=None, reye_pose=None, expression=None, transl=None,
                flame_betas=None, flame_expression=None, extra_betas=None):
        full_shape = self.get_full_shape(betas=betas, expression=expression, extra_betas=extra_betas)
        full_pose = self.get_full_pose(body_pose=body_pose, global_orient=global_orient,
                                       left_hand_pose=left_hand_pose, right_hand_pose=right_hand_pose)
        tr = self.get_full_transform(full_shape, full_pose)
        t_V = tr['V_shape_offset'].compose(tr['V_pose_offset'], tr['V_pose_rigid'])
        t_J = tr['J_shape_offset'].compose(tr['J_pose_rigid'])
        if transl is not None:
            t_tr = RigidTransform(T=transl)
            t_V, t_J = t_V.compose(t_tr), t_J.compose(t_tr)
            tr['G_transl_offset'] = t_tr
        else:
            N = full_shape.shape[0]
            tr['G_transl_offset'] = RigidTransform(SE3=torch.eye(4, dtype=full_shape.dtype, device=full_shape.device).expand(N, 4, 4))
        return t_J, t_V, tr
