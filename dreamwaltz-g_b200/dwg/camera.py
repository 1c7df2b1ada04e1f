"""R0: the camera ``data`` dict that feeds the hot path (format only, host side).

Mirrors reference data/camera/utils.py: angle2sphere :62-76, to_extrinsic :79-113,
to_projection :149-201, RandomCamera.__call__ :301-357 (defaults of RandomCamera4Avatar,
configs/__init__.py:310-315: radius U[1,2], fovy U[40,70], elevation U[60,120],
azimuth U[0,360], z_near 0.01, z_far 1000).  Tensors are float32, batch dim 1.
"""
import math

import numpy as np
import torch


def _normalize(v, eps=1e-20):
    return v / torch.sqrt(torch.clamp((v * v).sum(-1, keepdim=True), min=eps))


def to_extrinsic(radius, azimuth, elevation, at_vector=None):
    """radius/azimuth/elevation [B] (degrees) -> (extrinsic w2c [B,4,4], c2w [B,4,4]).
    Camera looks down +z; c2w columns = right, up, look-at (utils.py:101-113)."""
    B = radius.shape[0]
    az = azimuth * math.pi / 180.0
    el = elevation * math.pi / 180.0
    sph = torch.stack([radius * torch.sin(el) * torch.sin(az),
                       radius * torch.cos(el),
                       radius * torch.sin(el) * torch.cos(az)], dim=-1)
    if at_vector is None:
        at_vector = torch.zeros(B, 3, dtype=torch.float32)
    up = torch.tensor([[0.0, 1.0, 0.0]]).repeat(B, 1)
    pos = at_vector + sph
    look = _normalize(-sph)
    right = _normalize(torch.cross(look, up, dim=-1))
    up2 = _normalize(torch.cross(right, look, dim=-1))
    c2w = torch.eye(4, dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
    c2w[:, :3, :3] = torch.stack((right, up2, look), dim=-1)
    c2w[:, :3, 3] = pos
    return torch.inverse(c2w), c2w


def to_projection(tanfov, z_near=0.01, z_far=1000.0, tanfov_x=None):
    """utils.py:149-201: y flipped, z_sign=+1, z range (-1,1)."""
    B = tanfov.shape[0]
    max_y = tanfov * z_near
    max_x = max_y if tanfov_x is None else tanfov_x * z_near
    K = torch.zeros(B, 4, 4, dtype=torch.float32)
    K[:, 0, 0] = 2.0 * z_near / (2 * max_x)
    K[:, 1, 1] = -2.0 * z_near / (2 * max_y)
    K[:, 2, 2] = (z_far + z_near) / (z_far - z_near)
    K[:, 2, 3] = -(2 * z_far * z_near) / (z_far - z_near)
    K[:, 3, 2] = 1.0
    return K


def make_camera(radius, azimuth, elevation, fov, image_height, image_width, at=(0.0, 0.0, 0.0),
                z_near=0.01, z_far=1000.0):
    """Build the ``data`` dict for one view from scalar parameters (degrees)."""
    f32 = lambda v: torch.tensor([float(v)], dtype=torch.float32)
    radius, azimuth, elevation, fov = f32(radius), f32(azimuth), f32(elevation), f32(fov)
    tanfov = torch.tan(fov * math.pi / 180.0 * 0.5)
    extrinsic, c2w = to_extrinsic(radius, azimuth, elevation,
                                  at_vector=torch.tensor([list(at)], dtype=torch.float32))
    projection = to_projection(tanfov, z_near, z_far)
    return {
        'extrinsic': extrinsic, 'c2w': c2w, 'projection': projection,
        'mvp': torch.bmm(projection, extrinsic),
        'azimuth': azimuth, 'elevation': elevation, 'radius': radius, 'fov': fov, 'tanfov': tanfov,
        'z_far': z_far, 'z_near': z_near, 'image_height': int(image_height), 'image_width': int(image_width),
    }


def random_camera(rng: np.random.Generator, image_height, image_width,
                  radius_range=(1.0, 2.0), fovy_range=(40.0, 70.0),
                  elevation_range=(60.0, 120.0), azimuth_range=(0.0, 360.0)):
    """One draw of RandomCamera.__call__(size=1) with the RandomCamera4Avatar default ranges."""
    u = rng.uniform
    return make_camera(u(*radius_range), u(*azimuth_range), u(*elevation_range), u(*fovy_range),
                       image_height, image_width)


def raster_matrices(data):
    """gaussian_renderer.py:23-40: row-vector matrices the rasteriser consumes.
    viewmatrix = extrinsic^T, projmatrix = viewmatrix @ projection^T, campos = c2w[:3,3]."""
    view = data['extrinsic'][0].transpose(0, 1).contiguous()
    proj = (view @ data['projection'][0].transpose(0, 1)).contiguous()
    campos = data['c2w'][0, :3, 3].contiguous()
    tanfovy = float(data['tanfov'][0])
    tanfovx = float(data['tanfov_x'][0]) if 'tanfov_x' in data else tanfovy
    return view, proj, campos, tanfovx, tanfovy
