"""Oracle: real spherical-harmonic colour evaluation (row R10).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows reference
core/gaussian/spherical_harmonics.py:117-172 (eval_sh, degree <= 4), get_colors
core/gaussian/gaussian_utils.py:12-17 and GaussianRenderer.compute_colors
core/gaussian/gaussian_renderer.py:72-105.  Written as basis-vector x coefficient so that it
is an independent formulation; pinned against the reference's own eval_sh (which imports
here) by tests/golden/make_golden.py -> tests/golden/sh.npz.
"""
import torch

K0 = 0.28209479177387814
K1 = 0.4886025119029199
K2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
K3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435)
K4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
      -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761)


def sh_basis(deg, d):
    """d [N,3] unit directions -> basis [N,(deg+1)^2] with the reference's signs/constants."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    one = torch.ones_like(x)
    b = [K0 * one]
    if deg > 0:
        b += [-K1 * y, K1 * z, -K1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [K2[0] * xy, K2[1] * yz, K2[2] * (2.0 * zz - xx - yy), K2[3] * xz, K2[4] * (xx - yy)]
    if deg > 2:
        b += [K3[0] * y * (3 * xx - yy), K3[1] * xy * z, K3[2] * y * (4 * zz - xx - yy),
              K3[3] * z * (2 * zz - 3 * xx - 3 * yy), K3[4] * x * (4 * zz - xx - yy),
              K3[5] * z * (xx - yy), K3[6] * x * (xx - 3 * yy)]
    if deg > 3:
        b += [K4[0] * xy * (xx - yy), K4[1] * yz * (3 * xx - yy), K4[2] * xy * (7 * zz - 1),
              K4[3] * yz * (7 * zz - 3), K4[4] * (zz * (35 * zz - 30) + 3), K4[5] * xz * (7 * zz - 3),
              K4[6] * (xx - yy) * (7 * zz - 1), K4[7] * xz * (xx - 3 * yy),
              K4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(b, dim=-1)


def sh_colors(sh_features, positions, campos, sh_levels):
    """sh_features [N,K>=sh_levels^2,3], positions [N,3], campos [3] -> rgb [N,3].

    dirs = normalize(positions - campos) (F.normalize, eps 1e-12); colour =
    clamp_min(sum_k basis_k * sh_k + 0.5, 0)."""
    d = torch.nn.functional.normalize(positions - campos.view(1, 3), dim=-1)
    B = sh_basis(sh_levels - 1, d)                             # [N,K]
    rgb = torch.einsum('nk,nkc->nc', B, sh_features[:, :sh_levels ** 2])
    return torch.clamp_min(rgb + 0.5, 0.0)
