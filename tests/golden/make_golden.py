"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN PYTHON for the in-tree parts of
the hot path.  Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

What is executed from /root/reference (unmodified source text):
  * core/human/inverse_lbs.py            imported as a module (RigidTransform,
                                         GeneralLinearBlendSkinning) with ``smplx.lbs`` and
                                         ``pytorch3d.transforms`` shimmed by oracle/threep.py
                                         (those two packages are not installed; PARITY UNPINNED
                                         for them, see oracle/threep.py)
  * core/gaussian/spherical_harmonics.py, core/gaussian/gaussian_utils.py   imported
  * core/deformation/deform_model.py     imported (DeformNetwork)
  * core/nerf/nerf_model.py::MLP, utils/mesh.py::compute_normal,
    core/system/avatar.py::{MeshBindingGaussianModel.get_positions, .bary_coord_activation,
    .get_scales_and_quaternions, DreamWaltzG.lbs_transform, .non_rigid_transform}
                                         extracted with ``ast`` and exec'd (their modules
                                         import CUDA JIT builds / missing packages)
The outputs are written next to this file as small .npz fixtures; tests compare the oracle
(and, on the GPU, the CUDA path) against them.  Nothing here is imported by the product.
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))

from oracle import threep  # noqa: E402
from dwg import synth      # noqa: E402


def install_shims():
    smplx = types.ModuleType('smplx')

    class SMPL:                      # noqa: N801
        pass

    class SMPLX(SMPL):               # noqa: N801
        pass
    smplx.SMPL, smplx.SMPLX = SMPL, SMPLX
    lbs = types.ModuleType('smplx.lbs')
    for n in ('blend_shapes', 'batch_rodrigues', 'vertices2joints', 'batch_rigid_transform'):
        setattr(lbs, n, getattr(threep, n))
    smplx.lbs = lbs
    p3d = types.ModuleType('pytorch3d')
    tr = types.ModuleType('pytorch3d.transforms')
    for n in ('quaternion_to_matrix', 'matrix_to_quaternion', 'quaternion_multiply', 'standardize_quaternion'):
        setattr(tr, n, getattr(threep, n))
    p3d.transforms = tr
    sys.modules.update({'smplx': smplx, 'smplx.lbs': lbs, 'pytorch3d': p3d, 'pytorch3d.transforms': tr})
    # configs imports pyrallis-free dataclasses only; loguru is installed
    sys.path.insert(0, REF)
    return SMPLX


def extract(path, names):
    """Return {name: source} for top-level defs/classes or Class.method in a reference file."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            if node.name in names:
                out[node.name] = ast.get_source_segment(src, node)
            if isinstance(node, ast.ClassDef):
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef) and f'{node.name}.{sub.name}' in names:
                        seg = ast.get_source_segment(src, sub)
                        # include decorators (e.g. @staticmethod)
                        first = min([d.lineno for d in sub.decorator_list] + [sub.lineno])
                        lines = src.splitlines()[first - 1:sub.end_lineno]
                        import textwrap
                        out[f'{node.name}.{sub.name}'] = textwrap.dedent('\n'.join(lines))
    missing = set(names) - set(out)
    assert not missing, missing
    return out


def small_model(SMPLX, V=64, seed=3):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, generator=g) * sc
    m = SMPLX()
    m.NUM_JOINTS, m.NUM_BODY_JOINTS = 54, 21
    m.faces = np.zeros((1, 3), np.int64)
    m.parents = torch.tensor(synth.SMPLX_PARENTS)
    m.betas = r(1, 300, sc=0.5)
    m.body_pose = torch.zeros(1, 63)
    m.global_orient = torch.zeros(1, 3)
    m.v_template = r(V, 3, sc=0.5)
    m.shapedirs = r(V, 3, 300, sc=2e-3)
    m.posedirs = r(486, 3 * V, sc=1e-3)
    Jr = torch.rand(55, V, generator=g)
    m.J_regressor = Jr / Jr.sum(1, keepdim=True)
    W = torch.rand(V, 55, generator=g) ** 8
    m.lbs_weights = W / W.sum(1, keepdim=True)
    m.use_pca = False
    m.left_hand_pose = torch.zeros(1, 45)
    m.right_hand_pose = torch.zeros(1, 45)
    m.pose_mean = r(165, sc=0.05)
    m.left_hand_components = torch.eye(45)
    m.right_hand_components = torch.eye(45)
    m.jaw_pose = torch.zeros(1, 3)
    m.leye_pose = torch.zeros(1, 3)
    m.reye_pose = torch.zeros(1, 3)
    m.expr_dirs = r(V, 3, 100, sc=5e-4)
    m.expression = torch.zeros(1, 100)
    return m


def main():
    torch.manual_seed(0)
    SMPLX = install_shims()
    import core.human.inverse_lbs as ref_lbs
    from core.gaussian.gaussian_utils import get_colors
    from core.deformation.deform_model import DeformNetwork

    # ------------------------------------------------------------------ poses (real SMPL-X rows)
    aist = np.load(os.path.join(REF, 'assets/motions/aist.npy'))
    talk = np.load(os.path.join(REF, 'assets/motions/talkshow.npy'))
    poses = np.concatenate([aist[[0, 30, 60, 90, 120, 150, 180, 210]], talk[[0, 60, 120, 180]]]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'poses.npz'), rows=poses)

    # ------------------------------------------------------------------ LBS (R1-R3)
    sm = small_model(SMPLX)
    glbs = ref_lbs.GeneralLinearBlendSkinning(sm)
    inp = synth.pose_from_row(poses[3])
    inp['transl'] = torch.tensor([[0.03, -0.02, 0.05]])
    inp['expression'] = inp['expression'] * 0 + torch.randn(1, 100) * 0.3
    extra_betas = torch.randn(1, 300) * 0.1
    with torch.no_grad():
        tJ, tV, tr = glbs.forward(**inp)
        tJ2, tV2, _ = glbs.forward(**inp, extra_betas=extra_betas)
    N = 256
    x = torch.randn(N, 3) * 0.4
    q = torch.randn(N, 4)
    Wn = torch.rand(N, 55) ** 6
    Wn = Wn / Wn.sum(1, keepdim=True)
    av = extract('core/system/avatar.py', ['DreamWaltzG.lbs_transform', 'DreamWaltzG.non_rigid_transform',
                                            'MeshBindingGaussianModel.get_positions',
                                            'MeshBindingGaussianModel.bary_coord_activation',
                                            'MeshBindingGaussianModel.get_scales_and_quaternions'])
    import typing
    ns = {'Dict': typing.Dict, 'Optional': typing.Optional, 'Tuple': typing.Tuple, 'List': typing.List,
          'torch': torch, 'nn': nn, 'F': F, 'RigidTransform': ref_lbs.RigidTransform, 'GaussianOutput': object,
          'quaternion_multiply': threep.quaternion_multiply, 'matrix_to_quaternion': threep.matrix_to_quaternion,
          'standardize_quaternion': threep.standardize_quaternion}
    exec(extract('utils/mesh.py', ['safe_normalize', 'dot', 'compute_normal'])['safe_normalize'], ns)
    for k, v in extract('utils/mesh.py', ['dot', 'compute_normal']).items():
        exec(v, ns)
    for k, v in av.items():
        exec(v.replace('-> (tuple[torch.Tensor, torch.Tensor] | torch.Tensor)', ''), ns)
    self_lbs = types.SimpleNamespace(use_vertex_shape_offsets=False, use_joint_shape_offsets=False,
                                     use_vertex_pose_offsets=False)
    with torch.no_grad():
        xo, qo = ns['lbs_transform'](self_lbs, x, tr, Wn, None, quaternions=q)
        xo_only = ns['lbs_transform'](self_lbs, x, tr, Wn, None)
        A = ref_lbs.RigidTransform.compose(tr['J_pose_rigid'], tr['G_transl_offset']).squeeze(0).SE3
    model_np = {k: getattr(sm, k).numpy() for k in ('betas', 'v_template', 'shapedirs', 'posedirs', 'J_regressor',
                                                    'lbs_weights', 'pose_mean', 'expr_dirs', 'expression')}
    np.savez_compressed(
        os.path.join(HERE, 'lbs_small.npz'),
        **{f'model_{k}': v for k, v in model_np.items()},
        **{f'inp_{k}': v.numpy() for k, v in inp.items()},
        extra_betas=extra_betas.numpy(),
        J_SE3=tJ.SE3.numpy(), V_SE3=tV.SE3.numpy(), J_SE3_extra=tJ2.SE3.numpy(), V_SE3_extra=tV2.SE3.numpy(),
        V_shape_offset=tr['V_shape_offset'].T.numpy(), V_pose_offset=tr['V_pose_offset'].T.numpy(),
        J_pose_rigid=tr['J_pose_rigid'].SE3.numpy(), J_shape_offset=tr['J_shape_offset'].T.numpy(),
        A=A.numpy(), x=x.numpy(), q=q.numpy(), W=Wn.numpy(), x_out=xo.numpy(), q_out=qo.numpy(),
        x_out_only=xo_only.numpy())

    # RigidTransform unit behaviours (inverse mutates, compose order, index/weight)
    se3 = torch.randn(5, 4, 4)
    rt = ref_lbs.RigidTransform(SE3=se3.clone())
    inv = rt.inverse().SE3
    comp = ref_lbs.RigidTransform(SE3=se3[:1].clone()).compose(ref_lbs.RigidTransform(SE3=se3[1:2].clone()),
                                                             ref_lbs.RigidTransform(T=torch.tensor([[1., 2., 3.]]))).SE3
    wts = torch.rand(7, 5)
    np.savez_compressed(os.path.join(HERE, 'rigid.npz'), se3=se3.numpy(), inv=inv.numpy(), se3_after_inv=rt.SE3.numpy(),
                        comp=comp.numpy(), wts=wts.numpy(),
                        weighted=ref_lbs.RigidTransform(SE3=se3.clone()).weight(wts).SE3.numpy(),
                        q_matrix_mode=ref_lbs.RigidTransform(SE3=se3.clone()).transform_quaternions(
                            q[:7], weights=wts, rotation_mode='matrix').numpy(),
                        q_quat_mode=ref_lbs.RigidTransform(SE3=se3.clone()).transform_quaternions(
                            q[:7], weights=wts, rotation_mode='quaternion').numpy(),
                        q_in=q[:7].numpy())

    # ------------------------------------------------------------------ SH (R10)
    sh = torch.randn(128, 25, 3) * 0.4
    pos = torch.randn(128, 3)
    campos = torch.tensor([0.3, -0.2, 2.0])
    d = F.normalize(pos - campos, dim=-1)
    shd = {'sh': sh.numpy(), 'pos': pos.numpy(), 'campos': campos.numpy()}
    for lv in (1, 2, 3, 4, 5):
        shd[f'colors_l{lv}'] = get_colors(sh, d, lv).numpy()
    np.savez_compressed(os.path.join(HERE, 'sh.npz'), **shd)

    # ------------------------------------------------------------------ MLPs (R7, R8, R9)
    exec(extract('core/nerf/nerf_model.py', ['MLP'])['MLP'], ns)
    mlp = ns['MLP'](32, 4, 64, 3, bias=True)
    dn = DeformNetwork(xyz_input_ch=32, pose_input_ch=63, D=4, W=64)
    enc = torch.randn(200, 32) * 0.1
    body_pose = torch.tensor(poses[2:3, 12:75])
    with torch.no_grad():
        mo = mlp(enc.clone())
        dx, dsc, dro = dn(enc.clone(), body_pose)
    md = {'enc': enc.numpy(), 'body_pose': body_pose.numpy(), 'mlp_out': mo.numpy(), 'd_xyz': dx.numpy(),
          'd_scale': dsc.numpy(), 'd_rot': dro.numpy()}
    md.update({f'mlp.{k}': v.detach().numpy() for k, v in mlp.state_dict().items()})
    md.update({f'deform.{k}': v.detach().numpy() for k, v in dn.state_dict().items()})
    # non_rigid_transform with shipped flags
    self_nr = types.SimpleNamespace(use_non_rigid_offsets=True, init_offset=0.01, use_non_rigid_scales=True,
                                    learn_scale=False, scale_activation=torch.exp, init_scale=1e-3, max_scale=0.01,
                                    use_non_rigid_rotations=False, learn_quaternions=True,
                                    non_rigid_rotation_mode='add',
                                    get_quaternions=lambda: F.normalize(q[:200]))
    gs = types.SimpleNamespace(positions=x[:200].clone(), offsets=dx.clone(), scales=dsc.clone(), quaternions=dro.clone())
    with torch.no_grad():
        gs = ns['non_rigid_transform'](self_nr, gs)
    md.update({'nr_pos_in': x[:200].numpy(), 'nr_q_param': q[:200].numpy(), 'nr_pos': gs.positions.numpy(),
               'nr_scales': gs.scales.numpy(), 'nr_quats': gs.quaternions.numpy()})
    np.savez_compressed(os.path.join(HERE, 'mlp.npz'), **md)

    # ------------------------------------------------------------------ mesh-bound Gaussians (R5)
    model = synth.make_body_model(0)
    avatar = synth.make_avatar(model, 16, 40, seed=1)
    mesh = avatar['mesh']
    bary = mesh['_bary_coords'] + torch.rand_like(mesh['_bary_coords']) * 0.05
    sc = torch.rand_like(mesh['_scales']) * 2.5 + 0.2
    self_mb = types.SimpleNamespace(_bary_coords=bary, triangles=mesh['triangles'], _scales=sc, _n_points_per_triangle=6,
                                    bary_coord_activation=ns['bary_coord_activation'],
                                    get_vertex_coords=lambda: mesh['_vertex_coords'])
    Fn = mesh['triangles'].shape[0]
    p2t = torch.arange(Fn)[..., None].expand(-1, 6).reshape(-1)
    self_mb.points_to_vertices = mesh['triangles'][p2t]
    vc = mesh['_vertex_coords'] + torch.randn_like(mesh['_vertex_coords']) * 0.002
    with torch.no_grad():
        mp = ns['get_positions'](self_mb, vertex_coords=vc)
        msc, mq = ns['get_scales_and_quaternions'](self_mb, vertex_coords=vc, positions=mp)
    np.savez_compressed(os.path.join(HERE, 'mesh.npz'), vertex_coords=vc.numpy(), triangles=mesh['triangles'].numpy(),
                        bary=bary.numpy(), scales_param=sc.numpy(), positions=mp.numpy(), scales=msc.numpy(),
                        quats=mq.numpy())
    print('golden vectors written to', HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f'  {f}: {os.path.getsize(os.path.join(HERE, f)) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
