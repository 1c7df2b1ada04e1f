"""CPU: self-consistency and known answers for the rasteriser oracle (oracle/oracle_c.c).

The third-party rasteriser is not installable here (parity unpinned), so the oracle is
anchored on (a) hand-computed known answers, (b) an independent dense torch/float64
formulation of the same image formation model whose autograd gradients must agree with the
oracle's analytic backward, (c) structural invariants (sortedness, range partition)."""
import math

import numpy as np
import torch

from dwg import camera, synth
from oracle import raster as orast


def _cam(H=64, W=64, radius=2.0, az=20.0, el=80.0, fov=50.0, bg=(0.1, 0.2, 0.3)):
    d = camera.make_camera(radius, az, el, fov, H, W)
    view, proj, campos, tfx, tfy = camera.raster_matrices(d)
    return orast.make_camera(H, W, tfx, tfy, view.numpy(), proj.numpy(), bg), (view, proj, tfx, tfy, bg)


def test_spec_expf_accuracy():
    x = -np.abs(np.random.default_rng(0).normal(0, 3, size=4000)).astype(np.float32)
    got = orast.spec_expf(x)
    ref = np.exp(x.astype(np.float64))
    rel = np.abs(got - ref) / ref
    assert rel.max() < 2.5e-7          # ~2 ulp
    assert orast.spec_expf(np.float32(0.0)) == 1.0
    assert orast.spec_expf(np.float32(-100.0)) == 0.0


def test_single_gaussian_known_answer():
    """One isotropic Gaussian on the optical axis: radius, conic, centre alpha by hand."""
    H = W = 64
    cam, (view, proj, tfx, tfy, bg) = _cam(H, W, radius=2.0, az=0.0, el=90.0, fov=60.0, bg=(0, 0, 0))
    s = 0.05
    o = orast.forward(cam, [[0.0, 0.0, 0.0]], [[s, s, s]], [[1.0, 0, 0, 0]], [0.8], [[1.0, 0.5, 0.25]])
    f = W / (2 * tfx)
    var = (f * s / 2.0) ** 2 + 0.3                      # J Sigma J^T + 0.3 at depth 2
    assert o['radii'][0] == math.ceil(3 * math.sqrt(var))
    np.testing.assert_allclose(o['depth'][0], 2.0, rtol=1e-6)
    np.testing.assert_allclose(o['xy'][0], [(W - 1) / 2, (H - 1) / 2], atol=1e-4)
    np.testing.assert_allclose(o['conic_opacity'][0], [1 / var, 0.0, 1 / var, 0.8], rtol=1e-5, atol=1e-7)
    # pixel (32,32) is 0.5 px from the mean in x and y
    a = 0.8 * math.exp(-0.5 * (0.25 + 0.25) / var)
    np.testing.assert_allclose(o['out_alpha'][32, 32], a, rtol=1e-5)
    np.testing.assert_allclose(o['color'][:, 32, 32], np.array([1.0, 0.5, 0.25]) * a, rtol=1e-5)
    np.testing.assert_allclose(o['out_depth'][32, 32], 2.0 * a, rtol=1e-5)
    np.testing.assert_allclose(o['final_T'][32, 32], 1 - a, rtol=1e-5)
    assert o['n_contrib'][32, 32] == 1
    # culled behind the near threshold (view z <= 0.2)
    o2 = orast.forward(cam, [[0.0, 0.0, 1.9]], [[s, s, s]], [[1.0, 0, 0, 0]], [0.8], [[1.0, 0.5, 0.25]])
    assert o2['radii'][0] == 0 and o2['P'] == 0


def test_alpha_saturation_and_early_stop():
    """Opaque stack: alpha clamps at 0.99 and blending stops once T(1-a) < 1e-4; the stopping
    Gaussian itself is not blended and n_contrib points at the last blended one."""
    H = W = 32
    cam, _ = _cam(H, W, radius=2.0, az=0.0, el=90.0, fov=60.0, bg=(1, 1, 1))
    n = 6
    pos = [[0.0, 0.0, -0.1 * i] for i in range(n)]       # camera at +z looking down -z: increasing depth
    o = orast.forward(cam, pos, [[2.0] * 3] * n, [[1.0, 0, 0, 0]] * n, [1.0] * n, [[0.5, 0.5, 0.5]] * n)
    # T after the 1st blend = 1 - 0.99f = 0.00999999; test_T of the 2nd = 9.99998e-5 < 1e-4 -> stop:
    # exactly one Gaussian is blended even though six overlap the pixel
    c = 16
    assert o['n_contrib'][c, c] == 1
    np.testing.assert_allclose(o['final_T'][c, c], 0.01, rtol=1e-5)
    np.testing.assert_allclose(o['out_alpha'][c, c], 0.99, rtol=1e-6)
    np.testing.assert_allclose(o['color'][:, c, c], 0.5 * 0.99 + 0.01 * 1.0, rtol=1e-5)


def _scene(n=60, seed=0, H=48, W=64):
    g = synth.random_gaussians(n, seed=seed, extent=0.45, scale_range=(0.01, 0.06))
    g['opacities'] = g['opacities'] * 0.85
    cam, meta = _cam(H, W, radius=2.2, az=35.0, el=75.0, fov=45.0)
    return g, cam, meta, H, W


def test_binning_invariants():
    g, cam, meta, H, W = _scene(400, 1)
    o = orast.forward(cam, g['positions'], g['scales'], g['quaternions'], g['opacities'], g['colors'])
    assert o['P'] == int(o['tiles_touched'].sum()) == len(o['keys'])
    assert np.all(np.diff(o['keys'].astype(np.uint64)) >= 0)                       # sorted
    tiles = (o['keys'] >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        s, e = o['ranges'][t]
        assert np.all(tiles[s:e] == t) and (e - s) == np.sum(tiles == t)
    # keys carry the depth bits of their Gaussian
    db = o['depth'].view(np.uint32)[o['vals']]
    assert np.all((o['keys'] & np.uint64(0xffffffff)).astype(np.uint32) == db)
    assert np.all(o['n_contrib'] <= (o['ranges'][:, 1] - o['ranges'][:, 0]).max())


def _dense_torch_render(g, meta, rect, order, H, W, dtype=torch.float64):
    """Independent formulation: explicit EWA projection + per-pixel compositing in torch."""
    view, proj, tfx, tfy, bg = meta
    view, proj = view.to(dtype), proj.to(dtype)
    P, S, Q, O, Cc = (g[k] for k in ('positions', 'scales', 'quaternions', 'opacities', 'colors'))
    fx, fy = W / (2 * tfx), H / (2 * tfy)
    ones = torch.ones(P.shape[0], 1, dtype=dtype)
    ph = torch.cat([P, ones], 1)
    pv = ph @ view
    hom = ph @ proj
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    r, x, y, z = Q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * S[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    tz = pv[:, 2]
    tx = torch.clamp(pv[:, 0] / tz, -1.3 * tfx, 1.3 * tfx) * tz
    ty = torch.clamp(pv[:, 1] / tz, -1.3 * tfy, 1.3 * tfy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / tz ** 2, zero, fy / tz, -fy * ty / tz ** 2], -1).reshape(-1, 2, 3)
    Wm = view[:3, :3].t()                                   # world -> view rotation (column-vector form)
    Tm = J @ Wm
    cov = Tm @ Sigma @ Tm.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    cx, cy, cz = c / det, -b / det, a / det
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing='ij')
    tile_x, tile_y = (xs // 16).long(), (ys // 16).long()
    T = torch.ones(H, W, dtype=dtype)
    done = torch.zeros(H, W, dtype=torch.bool)
    col = torch.zeros(3, H, W, dtype=dtype); dep = torch.zeros(H, W, dtype=dtype); alp = torch.zeros(H, W, dtype=dtype)
    bgv = torch.tensor(bg, dtype=dtype)
    for gi in order:
        x0, y0, x1, y1 = [int(v) for v in rect[gi]]
        in_rect = (tile_x >= x0) & (tile_x < x1) & (tile_y >= y0) & (tile_y < y1)
        dx, dy = px[gi] - xs, py[gi] - ys
        power = -0.5 * (cx[gi] * dx * dx + cz[gi] * dy * dy) - cy[gi] * dx * dy
        al = torch.clamp(O[gi] * torch.exp(power), max=0.99)
        valid = in_rect & (power <= 0) & (al >= 1.0 / 255.0) & ~done
        test_T = T * (1 - al)
        stop = valid & (test_T < 1e-4)
        done = done | stop
        use = valid & ~stop
        w = torch.where(use, al * T, torch.zeros_like(T))
        col = col + Cc[gi][:, None, None] * w
        dep = dep + tz[gi] * w
        alp = alp + w
        T = torch.where(use, test_T, T)
    return col + T * bgv[:, None, None], dep, alp


def test_backward_matches_autograd_of_dense_formulation():
    g, cam, meta, H, W = _scene(60, 3)
    o = orast.forward(cam, g['positions'], g['scales'], g['quaternions'], g['opacities'], g['colors'])
    vis = np.nonzero(o['radii'] > 0)[0]
    order = vis[np.argsort(o['depth'][vis], kind='stable')]
    gd = {k: v.double().clone().requires_grad_(True) for k, v in g.items()}
    gd['opacities'] = g['opacities'].double().reshape(-1).clone().requires_grad_(True)
    col, dep, alp = _dense_torch_render(gd, meta, o['rect'], order, H, W)
    np.testing.assert_allclose(col.detach().numpy(), o['color'], atol=2e-5)
    np.testing.assert_allclose(dep.detach().numpy(), o['out_depth'], atol=5e-5)
    np.testing.assert_allclose(alp.detach().numpy(), o['out_alpha'], atol=2e-5)
    rng = np.random.default_rng(0)
    dc, dd, da = rng.normal(size=(3, H, W)), rng.normal(size=(H, W)), rng.normal(size=(H, W))
    loss = (col * torch.tensor(dc)).sum() + (dep * torch.tensor(dd)).sum() + (alp * torch.tensor(da)).sum()
    loss.backward()
    b = orast.backward(cam, o, dc, dd, da)

    def close(got, ref, name, rtol=2e-3):
        ref = ref.numpy()
        scale = np.abs(ref).max() + 1e-12
        err = np.abs(got - ref).max() / scale
        assert err < rtol, (name, err)
    close(b['colors'], gd['colors'].grad, 'colors')
    close(b['opacities'].reshape(-1), gd['opacities'].grad, 'opacities')
    close(b['means3D'], gd['positions'].grad, 'means3D')
    close(b['scales'], gd['scales'].grad, 'scales')
    close(b['rots'], gd['quaternions'].grad, 'rots')
