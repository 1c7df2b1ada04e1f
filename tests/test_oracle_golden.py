"""CPU: the oracle restatements reproduce the vectors produced by the reference's own code
(tests/golden/make_golden.py).  This is what pins the oracle for the in-tree parts."""
import numpy as np
import torch

from oracle import avatar as oav
from oracle import lbs as olbs
from oracle import sh as osh
from dwg import synth

T = torch.tensor


def _model_from_golden(g):
    m = {k[len('model_'):]: T(v) for k, v in g.items() if k.startswith('model_')}
    m['parents'] = synth.SMPLX_PARENTS
    m['J_template'] = torch.einsum('ik,ji->jk', m['v_template'], m['J_regressor'])
    return m


def test_glbs_forward_matches_reference(golden):
    g = golden('lbs_small')
    m = _model_from_golden(g)
    inp = {k[len('inp_'):]: T(v) for k, v in g.items() if k.startswith('inp_')}
    tJ, tV, tr = olbs.glbs_forward(m, **inp)
    np.testing.assert_allclose(tJ.SE3.numpy(), g['J_SE3'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tV.SE3.numpy(), g['V_SE3'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tr['V_shape_offset'].T.numpy(), g['V_shape_offset'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(tr['V_pose_offset'].T.numpy(), g['V_pose_offset'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(tr['J_pose_rigid'].SE3.numpy(), g['J_pose_rigid'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tr['J_shape_offset'].T.numpy(), g['J_shape_offset'], rtol=1e-5, atol=1e-7)
    tJ2, tV2, _ = olbs.glbs_forward(m, **inp, extra_betas=T(g['extra_betas']))
    np.testing.assert_allclose(tV2.SE3.numpy(), g['V_SE3_extra'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tJ2.SE3.numpy(), g['J_SE3_extra'], rtol=1e-5, atol=1e-6)


def test_lbs_transform_and_skin_match_reference(golden):
    g = golden('lbs_small')
    m = _model_from_golden(g)
    inp = {k[len('inp_'):]: T(v) for k, v in g.items() if k.startswith('inp_')}
    _, _, tr = olbs.glbs_forward(m, **inp)
    xo, qo = olbs.lbs_transform(T(g['x']), tr, T(g['W']), quaternions=T(g['q']))
    np.testing.assert_allclose(xo.numpy(), g['x_out'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(qo.numpy(), g['q_out'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(olbs.joint_pose_transform(tr).SE3.numpy(), g['A'], rtol=1e-5, atol=1e-6)
    # the fused-op spec agrees with the reference too
    xs, qs = olbs.skin(T(g['W']), T(g['A']), T(g['x']), T(g['q']))
    np.testing.assert_allclose(xs.numpy(), g['x_out'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(qs.numpy(), g['q_out'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(olbs.skin(T(g['W']), T(g['A']), T(g['x'])).numpy(), g['x_out_only'], rtol=1e-5, atol=1e-6)


def test_rigid_transform_quirks(golden):
    g = golden('rigid')
    rt = olbs.RigidTransform(SE3=T(g['se3']).clone())
    inv = rt.inverse()
    np.testing.assert_allclose(inv.SE3.numpy(), g['inv'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rt.SE3.numpy(), g['se3_after_inv'], rtol=0, atol=0)   # in-place last-row overwrite
    se3 = T(g['se3'])
    comp = olbs.RigidTransform(SE3=se3[:1].clone()).compose(olbs.RigidTransform(SE3=se3[1:2].clone()),
                                                          olbs.RigidTransform(T=T([[1., 2., 3.]])))
    np.testing.assert_allclose(comp.SE3.numpy(), g['comp'], rtol=1e-6, atol=1e-6)
    w = T(g['wts'])
    np.testing.assert_allclose(olbs.RigidTransform(SE3=se3.clone()).weight(w).SE3.numpy(), g['weighted'], rtol=1e-6, atol=1e-6)
    for mode in ('matrix', 'quat'):
        out = olbs.RigidTransform(SE3=se3.clone()).transform_quaternions(
            T(g['q_in']), weights=w, rotation_mode='matrix' if mode == 'matrix' else 'quaternion')
        np.testing.assert_allclose(out.numpy(), g[f'q_{mode}_mode'], rtol=1e-5, atol=1e-6)


def test_sh_matches_reference(golden):
    g = golden('sh')
    for lv in (1, 2, 3, 4, 5):
        c = osh.sh_colors(T(g['sh']), T(g['pos']), T(g['campos']), lv)
        np.testing.assert_allclose(c.numpy(), g[f'colors_l{lv}'], rtol=1e-5, atol=1e-6)


def test_mlps_and_nonrigid_match_reference(golden):
    g = golden('mlp')
    w = [T(g[f'mlp.net.{i}.weight']) for i in range(3)]
    b = [T(g[f'mlp.net.{i}.bias']) for i in range(3)]
    np.testing.assert_allclose(oav.mlp_forward(T(g['enc']), w, b).numpy(), g['mlp_out'], rtol=1e-5, atol=1e-6)
    p = {k[len('deform.'):]: T(v) for k, v in g.items() if k.startswith('deform.')}
    dx, ds, dr = oav.deform_forward(T(g['enc']), T(g['body_pose']), p)
    np.testing.assert_allclose(dx.numpy(), g['d_xyz'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ds.numpy(), g['d_scale'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dr.numpy(), g['d_rot'], rtol=1e-5, atol=1e-6)
    pos, sc, qu = oav.non_rigid(T(g['nr_pos_in']), dx, ds, T(g['nr_q_param']))
    np.testing.assert_allclose(pos.numpy(), g['nr_pos'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sc.numpy(), g['nr_scales'], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(qu.numpy(), g['nr_quats'], rtol=1e-6, atol=1e-7)


def test_mesh_bound_matches_reference(golden):
    g = golden('mesh')
    tri = T(g['triangles'])
    p = oav.mesh_positions(T(g['vertex_coords']), tri, T(g['bary']))
    np.testing.assert_allclose(p.numpy(), g['positions'], rtol=1e-6, atol=1e-7)
    s, q = oav.mesh_scales_quats(T(g['vertex_coords']), p, tri, T(g['bary']), T(g['scales_param']))
    np.testing.assert_allclose(s.numpy(), g['scales'], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(q.numpy(), g['quats'], rtol=1e-5, atol=1e-6)
