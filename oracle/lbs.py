"""Oracle: SMPL-X LBS as transforms + linear-blend skinning of Gaussians (rows R1-R3).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, in plain torch on the CPU:
  * RigidTransform                      reference core/human/inverse_lbs.py:15-260
  * GeneralLinearBlendSkinning.forward  reference core/human/inverse_lbs.py:570-784
  * DreamWaltzG.lbs_transform           reference core/system/avatar.py:1426-1462
Third-party maths comes from oracle/threep.py.  Pinned against the reference's own code by
tests/golden/make_golden.py (the reference module is executed with threep as its smplx /
pytorch3d shims) -> tests/golden/lbs_*.npz.
"""
import torch

from . import threep as tp


# --------------------------------------------------------------------------- RigidTransform
class RigidTransform:
    """SE3 container; field and method semantics of inverse_lbs.py:15-260."""

    def __init__(self, SE3=None, R=None, T=None):
        if SE3 is None:
            if R is not None and T is not None:        # :73-88
                SE3 = torch.zeros(*R.shape[:-2], 4, 4, dtype=R.dtype)
                SE3[..., :3, :3] = R
                SE3[..., :3, 3] = T
                SE3[..., 3, 3] = 1.0
            elif R is not None:                         # :41-54
                SE3 = torch.eye(4, dtype=R.dtype).expand(*R.shape[:-2], 4, 4).contiguous()
                SE3[..., :3, :3] = R
            elif T is not None:                         # :57-70
                SE3 = torch.eye(4, dtype=T.dtype).expand(*T.shape[:-1], 4, 4).contiguous()
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
        """:107-143.  NB the reference overwrites the last row of self.SE3 in place (:122)."""
        SE3 = self.SE3
        SE3[..., 3, :] = torch.tensor([0, 0, 0, 1], dtype=SE3.dtype)
        Rt = SE3[..., :3, :3].transpose(-1, -2)
        t_inv = -torch.matmul(Rt, SE3[..., :3, 3].unsqueeze(-1)).squeeze(-1)
        out = torch.zeros_like(SE3)
        out[..., :3, :3] = Rt
        out[..., :3, 3] = t_inv
        out[..., 3, 3] = 1.0
        return RigidTransform(SE3=out)

    def compose(self, *others):
        """:145-159: later transforms multiply on the left."""
        SE3 = self.SE3.clone()
        for other in others:
            SE3 = other.SE3 @ SE3
        return RigidTransform(SE3=SE3)

    def index(self, indices):
        return RigidTransform(SE3=self.SE3[indices])

    def weight(self, weights):
        """:174-180 (qr_correct=False): linear blend of whole 4x4 matrices."""
        return RigidTransform(SE3=torch.einsum('nj,jkl->nkl', weights, self.SE3))

    def squeeze(self, dim=0):
        self.SE3 = self.SE3.squeeze(dim)
        self.R = self.R.squeeze(dim)
        self.T = self.T.squeeze(dim)
        return self

    def transform_points(self, points, indices=None, weights=None):
        """:190-210."""
        assert indices is None or weights is None
        R, T = self.R, self.T
        if indices is not None:
            R, T = R[indices], T[indices]
        if weights is not None:
            R = torch.einsum('nj,jkl->nkl', weights, R)
            T = torch.einsum('nj,jk->nk', weights, T)
        return torch.matmul(R, points.unsqueeze(-1))[..., :, 0] + T

    def transform_quaternions(self, quaternions, indices=None, weights=None,
                              rotation_mode='quaternion', flip_rotation_axis=False):
        """:212-251."""
        assert indices is None or weights is None
        R = self.R
        if indices is not None:
            R = self.R[indices]
        if weights is not None:
            R = torch.einsum('nj,jkl->nkl', weights, self.R)
        if flip_rotation_axis:                       # :237-242
            rot = tp.quaternion_to_matrix(quaternions)
            sign = torch.tensor([1.0, -1.0, -1.0], dtype=rot.dtype).view(1, 3, 1)
            rot = rot * sign
            rot = R @ rot
            rot = rot * sign
            return tp.matrix_to_quaternion(rot)
        if rotation_mode == 'matrix':
            return tp.matrix_to_quaternion(R @ tp.quaternion_to_matrix(quaternions))
        if rotation_mode == 'quaternion':
            return tp.quaternion_multiply(tp.matrix_to_quaternion(R), quaternions)
        raise AssertionError(rotation_mode)


# ------------------------------------------------------- GeneralLinearBlendSkinning.forward
def glbs_forward(model, body_pose=None, global_orient=None, left_hand_pose=None, right_hand_pose=None,
                 expression=None, transl=None, extra_betas=None, betas=None):
    """SMPL-X forward kinematics as transforms (inverse_lbs.py:719-784).

    ``model`` is a dict of SMPL-X-shaped tensors: v_template [V,3], shapedirs [V,3,300],
    expr_dirs [V,3,100], posedirs [486,3V], J_regressor [55,V], lbs_weights [V,55],
    parents [55], betas [1,300], expression [1,100], pose_mean [165], J_template [55,3],
    and the module's own zero jaw/leye/reye poses.

    Returns (transform_J, transform_V, dict) exactly like the reference.  Quirks kept:
    jaw/leye/reye always come from the module (:617-619); transl composed last (:773-777).
    """
    dt = model['v_template'].dtype
    zeros3 = torch.zeros(1, 3, dtype=dt)
    betas = model['betas'] if betas is None else betas
    if extra_betas is not None:
        betas = betas + extra_betas                                            # :578-579
    expression = expression if expression is not None else model['expression']
    full_shape = torch.cat([betas, expression], dim=-1)                       # :582

    global_orient = global_orient if global_orient is not None else zeros3
    body_pose = body_pose if body_pose is not None else torch.zeros(1, 63, dtype=dt)
    left_hand_pose = left_hand_pose if left_hand_pose is not None else torch.zeros(1, 45, dtype=dt)
    right_hand_pose = right_hand_pose if right_hand_pose is not None else torch.zeros(1, 45, dtype=dt)
    full_pose = torch.cat([global_orient.reshape(-1, 1, 3), body_pose.reshape(-1, 21, 3),
                           zeros3.reshape(-1, 1, 3), zeros3.reshape(-1, 1, 3), zeros3.reshape(-1, 1, 3),
                           left_hand_pose.reshape(-1, 15, 3), right_hand_pose.reshape(-1, 15, 3)],
                          dim=1).reshape(-1, 165)                              # :615-622
    full_pose = full_pose + model['pose_mean']                                 # :626

    # get_full_transform :652-717
    B = max(full_shape.shape[0], full_pose.shape[0])
    shapedirs = torch.cat([model['shapedirs'], model['expr_dirs']], dim=-1)
    shape_offsets = tp.blend_shapes(full_shape, shapedirs)
    v_shaped = model['v_template'] + shape_offsets
    J = tp.vertices2joints(model['J_regressor'], v_shaped)
    t_J_shape = RigidTransform(T=J - model['J_template'])
    ident = torch.eye(3, dtype=dt)
    rot_mats = tp.batch_rodrigues(full_pose.view(-1, 3)).view(B, -1, 3, 3)
    pose_feature = rot_mats[:, 1:, :, :] - ident
    pose_offsets = torch.matmul(pose_feature.view(B, -1), model['posedirs']).view(B, -1, 3)
    _, A = tp.batch_rigid_transform(rot_mats, J, model['parents'], dtype=dt)
    W = model['lbs_weights'].unsqueeze(0).expand(B, -1, -1)
    nj = model['J_regressor'].shape[0]
    T = torch.matmul(W, A.view(B, nj, 16)).view(B, -1, 4, 4)
    tr = {
        'V_shape_offset': RigidTransform(T=shape_offsets),
        'V_pose_offset': RigidTransform(T=pose_offsets),
        'V_pose_rigid': RigidTransform(SE3=T),
        'J_shape_offset': t_J_shape,
        'J_pose_rigid': RigidTransform(SE3=A),
    }
    t_V = tr['V_shape_offset'].compose(tr['V_pose_offset'], tr['V_pose_rigid'])
    t_J = tr['J_shape_offset'].compose(tr['J_pose_rigid'])
    if transl is not None:
        t_tr = RigidTransform(T=transl)
        t_V = t_V.compose(t_tr)
        t_J = t_J.compose(t_tr)
        tr['G_transl_offset'] = t_tr
    else:
        tr['G_transl_offset'] = RigidTransform(SE3=torch.eye(4, dtype=dt).expand(full_shape.shape[0], 4, 4))
    return t_J, t_V, tr


# ------------------------------------------------------------------ avatar.lbs_transform
def joint_pose_transform(transforms):
    """avatar.py:1446-1449: J_pose_rigid then G_transl_offset, batch dim squeezed -> SE3 [55,4,4]."""
    return transforms['J_pose_rigid'].compose(transforms['G_transl_offset']).squeeze(0)


def lbs_transform(positions, transforms, lbs_weights, quaternions=None):
    """avatar.py:1426-1462 with the default flags (no vertex/joint shape or pose offsets,
    configs/__init__.py:117-119)."""
    jt = joint_pose_transform(transforms)
    out = jt.transform_points(positions, weights=lbs_weights)
    if quaternions is not None:
        q = jt.transform_quaternions(quaternions, weights=lbs_weights, flip_rotation_axis=True)
        return out, q
    return out


def normalise_lbs_weights(w):
    """avatar.py:914-917 get_lbs_weights: row-normalise the stored weights."""
    return w / w.sum(dim=-1, keepdim=True)


# ---------------------------------------------- the fused op the CUDA kernel implements
def skin(W, A, x, q=None):
    """Spec of dwg_lbs_skin_fwd.  W [N,J], A [J,4,4] (joint SE3), x [N,3], q [N,4] or None.

      M_n = sum_j W[n,j] A[j,:3,:]          (3x4, linear blend, not re-orthogonalised)
      x'_n = M_n[:, :3] x_n + M_n[:, 3]
      q'_n = mat2quat( F (M_n[:, :3] (F quat2mat(q_n))) )   with F = diag(1,-1,-1) on rows

    which is transform_points(weights=W) + transform_quaternions(weights=W,
    flip_rotation_axis=True) of inverse_lbs.py:190-242.
    """
    jt = RigidTransform(SE3=A)
    xo = jt.transform_points(x, weights=W)
    if q is None:
        return xo
    qo = jt.transform_quaternions(q, weights=W, flip_rotation_axis=True)
    return xo, qo
