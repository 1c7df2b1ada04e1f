"""Oracle restatement of the un-vendored third-party maths the LBS path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: neither ``smplx`` nor
``pytorch3d`` is installed in the build image, so these follow the published algorithms
 and are anchored on the reference's call sites:

  smplx.lbs.{blend_shapes, vertices2joints, batch_rodrigues, batch_rigid_transform}
      pinned version: smplx @ git HEAD (reference scripts/install.sh:24)
      call sites: core/human/inverse_lbs.py:640,645,676,681,688,696
  pytorch3d.transforms.{quaternion_to_matrix, matrix_to_quaternion, quaternion_multiply,
                        standardize_quaternion}
      pinned version: pytorch3d 0.7.5 (reference scripts/install.sh:8)
      call sites: core/human/inverse_lbs.py:238,242,245-249; core/system/avatar.py:1077,1489

All functions are plain torch (CPU, any float dtype) and differentiable, so the oracle's
backward pass is torch autograd over these definitions.
"""
import torch
import torch.nn.functional as F

# SMPL-X kinematic tree (55 joints), SURVEY.md appendix A.
SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
                 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]


# ----------------------------------------------------------------------------- smplx.lbs
def blend_shapes(betas, shape_disps):
    """betas [B,L], shape_disps [V,3,L] -> [B,V,3]  (einsum 'bl,mkl->bmk')."""
    return torch.einsum('bl,mkl->bmk', betas, shape_disps)


def vertices2joints(J_regressor, vertices):
    """J_regressor [J,V], vertices [B,V,3] -> [B,J,3]  (einsum 'bik,ji->bjk')."""
    return torch.einsum('bik,ji->bjk', vertices, J_regressor)


def batch_rodrigues(rot_vecs):
    """Axis-angle [N,3] -> rotation matrices [N,3,3].

    angle = ||r + 1e-8||, d = r / angle, R = I + sin(angle) K + (1 - cos(angle)) K K.
    """
    batch_size = rot_vecs.shape[0]
    dtype = rot_vecs.dtype
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.unsqueeze(torch.cos(angle), dim=1)
    sin = torch.unsqueeze(torch.sin(angle), dim=1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((batch_size, 1), dtype=dtype)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view((batch_size, 3, 3))
    ident = torch.eye(3, dtype=dtype).unsqueeze(dim=0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def _transform_mat(R, t):
    """R [B,3,3], t [B,3,1] -> [B,4,4]."""
    return torch.cat([F.pad(R, [0, 0, 0, 1]), F.pad(t, [0, 0, 0, 1], value=1)], dim=2)


def batch_rigid_transform(rot_mats, joints, parents, dtype=torch.float32):
    """rot_mats [B,J,3,3], joints [B,J,3], parents [J] -> (posed_joints [B,J,3], A [B,J,4,4])."""
    joints = torch.unsqueeze(joints, dim=-1)
    rel_joints = joints.clone()
    parents = torch.as_tensor(parents, dtype=torch.long)
    rel_joints[:, 1:] = rel_joints[:, 1:] - joints[:, parents[1:]]
    transforms_mat = _transform_mat(rot_mats.reshape(-1, 3, 3),
                                    rel_joints.reshape(-1, 3, 1)).reshape(-1, joints.shape[1], 4, 4)
    transform_chain = [transforms_mat[:, 0]]
    for i in range(1, parents.shape[0]):
        transform_chain.append(torch.matmul(transform_chain[int(parents[i])], transforms_mat[:, i]))
    transforms = torch.stack(transform_chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_homogen = F.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - F.pad(torch.matmul(transforms, joints_homogen), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


# ------------------------------------------------------------------ pytorch3d.transforms
def quaternion_to_matrix(quaternions):
    """Real-first quaternions [...,4] -> [...,3,3]; two_s = 2 / sum(q*q) (no normalisation needed)."""
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    """sqrt(max(0,x)) with zero subgradient where x <= 0."""
    ret = torch.zeros_like(x)
    positive_mask = x > 0
    ret[positive_mask] = torch.sqrt(x[positive_mask])
    return ret


def matrix_to_quaternion(matrix):
    """[...,3,3] -> real-first quaternion [...,4]; pytorch3d 0.7.5: best-conditioned of four
    candidates (argmax of q_abs, first on ties), each divided by 2*max(q_abs, 0.1); no sign
    standardisation."""
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch_dim + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(
        torch.stack(
            [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22],
            dim=-1,
        )
    )
    quat_by_rijk = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
        ],
        dim=-2,
    )
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    sel = F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return quat_candidates[sel, :].reshape(batch_dim + (4,))


def standardize_quaternion(quaternions):
    return torch.where(quaternions[..., 0:1] < 0, -quaternions, quaternions)


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack((ow, ox, oy, oz), -1)


def quaternion_multiply(a, b):
    return standardize_quaternion(quaternion_raw_multiply(a, b))
