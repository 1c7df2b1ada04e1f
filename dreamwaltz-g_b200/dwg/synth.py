"""Seeded synthetic inputs with the shapes of the real assets.

No SMPL-X model file, avatar checkpoint or diffusion weight exists in this environment, so
the bench and the parity tests run on synthetic tensors of the right shapes:
  * an SMPL-X-shaped body model: V=10475 vertices on capsule limbs around a 55-joint
    skeleton (SMPL-X kinematic tree), 300+100 shape/expression directions, 486 pose-feature
    rows, a sparse joint regressor and <=4-sparse skinning weights;
  * an avatar: N_u unconstrained Gaussians sampled on that surface (+5 mm noise) with
    barycentrically mixed skinning weights, plus mesh-bound Gaussians (6 per triangle) on
    the hand triangles;
  * poses: real SMPL-X pose rows (8 rows of the reference's assets/motions/aist.npy are
    committed as tests/golden/poses.npz) or small random axis-angles.
Everything is numpy/torch on the CPU; callers move it to the device.
"""
import numpy as np
import torch

SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
                 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]
NUM_JOINTS = 55
NUM_VERTS = 10475
NUM_BETAS = 300
NUM_EXPR = 100


def _rest_joints():
    """Approximate SMPL-X rest skeleton (metres, y up, facing +z)."""
    J = np.zeros((55, 3), np.float64)
    J[0] = (0, -0.25, 0)
    J[1], J[2], J[3] = (0.06, -0.34, 0), (-0.06, -0.34, 0), (0, -0.13, 0)
    J[4], J[5], J[6] = (0.10, -0.72, 0), (-0.10, -0.72, 0), (0, 0.0, 0)
    J[7], J[8], J[9] = (0.09, -1.12, -0.03), (-0.09, -1.12, -0.03), (0, 0.06, 0)
    J[10], J[11], J[12] = (0.11, -1.18, 0.09), (-0.11, -1.18, 0.09), (0, 0.27, -0.02)
    J[13], J[14], J[15] = (0.05, 0.18, 0), (-0.05, 0.18, 0), (0, 0.36, 0)
    J[16], J[17] = (0.17, 0.22, -0.02), (-0.17, 0.22, -0.02)
    J[18], J[19] = (0.43, 0.21, -0.03), (-0.43, 0.21, -0.03)
    J[20], J[21] = (0.68, 0.21, -0.03), (-0.68, 0.21, -0.03)
    J[22], J[23], J[24] = (0, 0.33, 0.03), (0.03, 0.40, 0.07), (-0.03, 0.40, 0.07)
    zoff = [0.03, 0.01, -0.03, -0.01, 0.04]          # index, middle, pinky, ring, thumb
    for f in range(5):
        for k in range(3):
            x = 0.09 + 0.03 * k if f < 4 else 0.04 + 0.025 * k
            J[25 + 3 * f + k] = (0.68 + x, 0.21, -0.03 + zoff[f])
            J[40 + 3 * f + k] = (-0.68 - x, 0.21, -0.03 + zoff[f])
    return J


def _bone_radius(j):
    if j in (0, 3, 6, 9):
        return 0.11
    if j in (1, 2):
        return 0.07
    if j in (4, 5):
        return 0.05
    if j in (7, 8, 10, 11):
        return 0.035
    if j in (12,):
        return 0.05
    if j == 15:
        return 0.09
    if j in (13, 14, 16, 17):
        return 0.045
    if j in (18, 19):
        return 0.035
    if j in (20, 21):
        return 0.03
    if j in (22, 23, 24):
        return 0.02
    return 0.008            # fingers


def make_body_model(seed=0, dtype=torch.float32):
    """SMPL-X-shaped model dict (see oracle/lbs.py:glbs_forward for the field list)."""
    rng = np.random.default_rng(seed)
    J = _rest_joints()
    parents = np.array(SMPLX_PARENTS)
    children = {j: [c for c in range(55) if parents[c] == j] for j in range(55)}
    # one tube per joint: from the joint towards its first child (or a short stub)
    verts, faces, vjoint = [], [], []
    finger = lambda j: j >= 25
    for j in range(55):
        if children[j]:
            end = J[children[j][0]]
        else:
            d = J[j] - J[parents[j]]
            end = J[j] + d / (np.linalg.norm(d) + 1e-9) * (0.12 if j == 15 else 0.03 if not finger(j) else 0.02)
        rings, segs = (8, 12) if finger(j) else (12, 20) if j not in (15,) else (30, 40)
        if j in (0, 3, 6, 9):
            rings, segs = 14, 32
        axis = end - J[j]
        L = np.linalg.norm(axis) + 1e-9
        a = axis / L
        ref = np.array([0, 0, 1.0]) if abs(a[2]) < 0.9 else np.array([1.0, 0, 0])
        u = np.cross(a, ref); u /= np.linalg.norm(u)
        v = np.cross(a, u)
        r = _bone_radius(j)
        base = len(verts)
        for i in range(rings):
            t = i / (rings - 1)
            rr = r * (0.55 + 0.45 * np.sin(np.pi * min(max(t, 0.08), 0.92)))
            for s in range(segs):
                ang = 2 * np.pi * s / segs
                verts.append(J[j] + a * L * t + rr * (np.cos(ang) * u + np.sin(ang) * v))
                vjoint.append(j)
        for i in range(rings - 1):
            for s in range(segs):
                p0 = base + i * segs + s
                p1 = base + i * segs + (s + 1) % segs
                p2 = p0 + segs
                p3 = p1 + segs
                faces.append((p0, p1, p2)); faces.append((p1, p3, p2))
    verts = np.array(verts); vjoint = np.array(vjoint); faces = np.array(faces, np.int64)
    # pad / trim to exactly NUM_VERTS with extra head-sphere points (isolated vertices are legal)
    if len(verts) < NUM_VERTS:
        k = NUM_VERTS - len(verts)
        d = rng.normal(size=(k, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        verts = np.concatenate([verts, J[15] + np.array([0, 0.07, 0]) + 0.1 * d])
        vjoint = np.concatenate([vjoint, np.full(k, 15)])
    else:
        keep = NUM_VERTS
        faces = faces[(faces < keep).all(1)]
        verts, vjoint = verts[:keep], vjoint[:keep]
    V = NUM_VERTS
    # skinning weights: own joint, parent, first child, grandparent; distance-softmax
    W = np.zeros((V, 55), np.float64)
    for j in range(55):
        idx = np.nonzero(vjoint == j)[0]
        cand = [j]
        if parents[j] >= 0:
            cand.append(parents[j])
            if parents[parents[j]] >= 0:
                cand.append(parents[parents[j]])
        if children[j]:
            cand.append(children[j][0])
        cand = cand[:4]
        d = np.stack([np.linalg.norm(verts[idx] - J[c], axis=1) for c in cand], 1)
        w = np.exp(-d / 0.05)
        w /= w.sum(1, keepdims=True)
        for k, c in enumerate(cand):
            W[idx, c] += w[:, k]
    # joint regressor: mean of the 64 nearest vertices, then a rest-pose correction is folded in
    Jreg = np.zeros((55, V), np.float64)
    for j in range(55):
        d = np.linalg.norm(verts - J[j], axis=1)
        nn = np.argsort(d)[:64]
        Jreg[j, nn] = 1.0 / 64
    model = {
        'v_template': torch.tensor(verts, dtype=dtype),
        'faces': torch.tensor(faces, dtype=torch.long),
        'vertex_joint': torch.tensor(vjoint, dtype=torch.long),
        'shapedirs': torch.tensor(rng.normal(0, 2e-3, size=(V, 3, NUM_BETAS)), dtype=dtype),
        'expr_dirs': torch.tensor(rng.normal(0, 5e-4, size=(V, 3, NUM_EXPR)), dtype=dtype),
        'posedirs': torch.tensor(rng.normal(0, 1e-3, size=(486, 3 * V)), dtype=dtype),
        'J_regressor': torch.tensor(Jreg, dtype=dtype),
        'lbs_weights': torch.tensor(W, dtype=dtype),
        'parents': list(SMPLX_PARENTS),
        'betas': torch.zeros(1, NUM_BETAS, dtype=dtype),
        'expression': torch.zeros(1, NUM_EXPR, dtype=dtype),
        'pose_mean': torch.zeros(165, dtype=dtype),
    }
    model['J_template'] = torch.einsum('ik,ji->jk', model['v_template'], model['J_regressor'])
    return model


def hand_triangles(model, max_triangles=2500):
    """Triangle subset on both hands (wrist+finger tubes) for the mesh-bound Gaussians."""
    vj = model['vertex_joint']
    f = model['faces']
    on_hand = (vj[f] >= 25).all(1) | ((vj[f] == 20) | (vj[f] == 21) | (vj[f] >= 25)).all(1)
    tri = torch.nonzero(on_hand)[:, 0]
    if tri.numel() > max_triangles:
        sel = torch.linspace(0, tri.numel() - 1, max_triangles).long()
        tri = tri[sel]
    return tri


def face_triangles(model, max_triangles=2500):
    """Triangle subset on the head (head / jaw / eye joints 15, 22-24) for the mesh-bound 'face' part
    (predefined_body_parts=hands,face of scripts/train_w_expr.sh)."""
    vj = model['vertex_joint']
    f = model['faces']
    on_face = ((vj[f] == 15) | ((vj[f] >= 22) & (vj[f] <= 24))).all(1)
    tri = torch.nonzero(on_face)[:, 0]
    if tri.numel() > max_triangles:
        tri = tri[torch.linspace(0, tri.numel() - 1, max_triangles).long()]
    return tri


def _mesh_part(model, tri_sel, dtype):
    f, vt = model['faces'], model['v_template']
    tris = f[tri_sel]
    vidx, inv = torch.unique(tris.reshape(-1), return_inverse=True)
    bary6 = torch.tensor([[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3],
                          [1 / 6, 5 / 12, 5 / 12], [5 / 12, 1 / 6, 5 / 12], [5 / 12, 5 / 12, 1 / 6]], dtype=dtype)
    return {'predefined_vertex_indices': vidx, 'triangles': inv.reshape(-1, 3), '_vertex_coords': vt[vidx].clone(),
            '_bary_coords': bary6.expand(tris.shape[0], -1, -1).clone(), '_scales': torch.ones(tris.shape[0] * 6, 3, dtype=dtype)}


def make_avatar(model, n_unconstrained=135000, n_mesh_triangles=2500, seed=0, dtype=torch.float32, n_face_triangles=0):
    """Synthetic avatar state: unconstrained Gaussians + mesh-bound hand Gaussians (+ a mesh-bound 'face' part when
    n_face_triangles > 0: avatar['meshes'] = {'hands': ..., 'face': ...})."""
    g = torch.Generator().manual_seed(seed)
    f = model['faces']
    vt = model['v_template']
    # area-weighted triangle sampling
    a, b, c = vt[f[:, 0]], vt[f[:, 1]], vt[f[:, 2]]
    area = torch.linalg.norm(torch.cross(b - a, c - a, dim=-1), dim=-1) + 1e-12
    tri = torch.multinomial(area / area.sum(), n_unconstrained, replacement=True, generator=g)
    bary = torch.rand(n_unconstrained, 3, generator=g, dtype=dtype)
    bary = bary / bary.sum(-1, keepdim=True)
    pos = (bary[:, :, None] * vt[f[tri]]).sum(1) + 0.005 * torch.randn(n_unconstrained, 3, generator=g, dtype=dtype)
    W = (bary[:, :, None] * model['lbs_weights'][f[tri]]).sum(1)
    quats = torch.randn(n_unconstrained, 4, generator=g, dtype=dtype)
    scales = torch.log(torch.empty(n_unconstrained, 3, dtype=dtype).uniform_(0.0005, 0.01, generator=g))
    avatar = {
        '_positions': pos.contiguous(),
        '_lbs_weights': W.contiguous(),
        '_quaternions': quats,
        '_scales': scales,
    }
    tri_sel = hand_triangles(model, n_mesh_triangles)
    tris = f[tri_sel]
    vidx, inv = torch.unique(tris.reshape(-1), return_inverse=True)
    bary6 = torch.tensor([[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3],
                          [1 / 6, 5 / 12, 5 / 12], [5 / 12, 1 / 6, 5 / 12], [5 / 12, 5 / 12, 1 / 6]], dtype=dtype)
    avatar['mesh'] = {
        'predefined_vertex_indices': vidx,                       # into the V model vertices
        'triangles': inv.reshape(-1, 3),                         # remapped to [0, V_p)
        '_vertex_coords': vt[vidx].clone(),
        '_bary_coords': bary6.expand(tris.shape[0], -1, -1).clone(),
        '_scales': torch.ones(tris.shape[0] * 6, 3, dtype=dtype),
    }
    if n_face_triangles > 0:
        ft = face_triangles(model, n_face_triangles)
        if ft.numel() > 0:
            avatar['meshes'] = {'hands': avatar['mesh'], 'face': _mesh_part(model, ft, dtype)}
    return avatar


def random_pose(rng: np.random.Generator, scale=0.25, dtype=torch.float32):
    """Small random axis-angle SMPL-X inputs (dict of [1,*] tensors)."""
    t = lambda a: torch.tensor(a, dtype=dtype)
    return {
        'global_orient': t(rng.normal(0, 0.1, size=(1, 3))),
        'body_pose': t(rng.normal(0, scale, size=(1, 63))),
        'left_hand_pose': t(rng.normal(0, scale, size=(1, 45))),
        'right_hand_pose': t(rng.normal(0, scale, size=(1, 45))),
        'expression': t(np.zeros((1, 100))),
        'transl': t(rng.normal(0, 0.02, size=(1, 3))),
    }


def pose_from_row(row, dtype=torch.float32):
    """Reference data/human/demo.py:17-24 column layout of a [265] SMPL-X motion row."""
    r = torch.as_tensor(np.asarray(row), dtype=dtype).reshape(1, -1)
    return {
        'global_orient': r[:, 9:12].contiguous(), 'body_pose': r[:, 12:75].contiguous(),
        'left_hand_pose': r[:, 75:120].contiguous(), 'right_hand_pose': r[:, 120:165].contiguous(),
        'expression': r[:, 165:265].contiguous(),
    }


def random_gaussians(n, seed=0, dtype=torch.float32, extent=0.5, scale_range=(0.002, 0.02)):
    """Free-standing Gaussians for raster tests: positions in a box around the origin."""
    g = torch.Generator().manual_seed(seed)
    pos = (torch.rand(n, 3, generator=g, dtype=dtype) - 0.5) * 2 * extent
    scales = torch.empty(n, 3, dtype=dtype).uniform_(scale_range[0], scale_range[1], generator=g)
    quats = torch.randn(n, 4, generator=g, dtype=dtype)
    quats = quats / quats.norm(dim=-1, keepdim=True)
    opac = torch.sigmoid(2 + torch.randn(n, 1, generator=g, dtype=dtype))
    cols = torch.rand(n, 3, generator=g, dtype=dtype)
    return {'positions': pos, 'scales': scales, 'quaternions': quats, 'opacities': opac, 'colors': cols}
