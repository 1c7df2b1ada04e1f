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
    return v / np.sqrt(np.maximum((v * v).sum(-1, keepdims=True), np.float32(eps)))


def to_extrinsic(radius, azimuth, elevation, at_vector=None):
    """radius/azimuth/elevation [B] (degrees) -> (extrinsic w2c [B,4,4], c2w [B,4,4]) float32 tensors.
    Camera looks down +z; c2w columns = right, up, look-at (utils.py:101-113).  Host arithmetic in numpy float32 (a few
    microseconds per view instead of ~25 tiny torch ops); c2w is a rigid transform, so its inverse is written in closed
    form [R^T | -R^T t] (the reference calls torch.inverse: same matrix to fp32 rounding, tests/test_camera_golden.py)."""
    f32 = np.float32
    radius, azimuth, elevation = (np.asarray(torch.as_tensor(x).numpy() if torch.is_tensor(x) else x, dtype=f32).reshape(-1)
                                  for x in (radius, azimuth, elevation))
    B = radius.shape[0]
    az = azimuth * f32(math.pi) / f32(180.0)
    el = elevation * f32(math.pi) / f32(180.0)
    sph = np.stack([radius * np.sin(el) * np.sin(az), radius * np.cos(el), radius * np.sin(el) * np.cos(az)], axis=-1).astype(f32)
    at = np.zeros((B, 3), f32) if at_vector is None else np.asarray(torch.as_tensor(at_vector).numpy() if torch.is_tensor(at_vector) else at_vector, f32).reshape(B, 3)
    up = np.tile(np.array([[0.0, 1.0, 0.0]], f32), (B, 1))
    pos = at + sph
    look = _normalize(-sph)
    right = _normalize(np.cross(look, up))
    up2 = _normalize(np.cross(right, look))
    R = np.stack((right, up2, look), axis=-1).astype(f32)                      # columns
    c2w = np.tile(np.eye(4, dtype=f32)[None], (B, 1, 1))
    c2w[:, :3, :3] = R
    c2w[:, :3, 3] = pos
    ext = np.tile(np.eye(4, dtype=f32)[None], (B, 1, 1))
    Rt = R.transpose(0, 2, 1)
    ext[:, :3, :3] = Rt
    ext[:, :3, 3] = -(Rt @ pos[:, :, None])[:, :, 0]
    return torch.from_numpy(ext), torch.from_numpy(c2w)


def to_projection(tanfov, z_near=0.01, z_far=1000.0, tanfov_x=None):
    """utils.py:149-201: y flipped, z_sign=+1, z range (-1,1)."""
    f32 = np.float32
    tanfov = np.asarray(torch.as_tensor(tanfov).numpy() if torch.is_tensor(tanfov) else tanfov, f32).reshape(-1)
    B = tanfov.shape[0]
    max_y = tanfov * f32(z_near)
    max_x = max_y if tanfov_x is None else np.asarray(torch.as_tensor(tanfov_x).numpy(), f32).reshape(-1) * f32(z_near)
    K = np.zeros((B, 4, 4), f32)
    K[:, 0, 0] = f32(2.0) * f32(z_near) / (max_x - (-max_x))
    K[:, 1, 1] = -f32(2.0) * f32(z_near) / (max_y - (-max_y))
    K[:, 2, 2] = f32((z_far + z_near) / (z_far - z_near))
    K[:, 2, 3] = f32(-(2 * z_far * z_near) / (z_far - z_near))
    K[:, 3, 2] = 1.0
    return torch.from_numpy(K)


def make_camera(radius, azimuth, elevation, fov, image_height, image_width, at=(0.0, 0.0, 0.0),
                z_near=0.01, z_far=1000.0):
    """Build the ``data`` dict for one view from scalar parameters (degrees)."""
    f32 = lambda v: torch.tensor([float(v)], dtype=torch.float32)
    radius, azimuth, elevation, fov = f32(radius), f32(azimuth), f32(elevation), f32(fov)
    tanfov = torch.from_numpy(np.tan(fov.numpy() * np.float32(math.pi) / np.float32(180.0) / np.float32(2.0)).astype(np.float32))
    extrinsic, c2w = to_extrinsic(radius, azimuth, elevation, at_vector=np.asarray([list(at)], np.float32))
    projection = to_projection(tanfov, z_near, z_far)
    return {
        'extrinsic': extrinsic, 'c2w': c2w, 'projection': projection,
        'mvp': torch.from_numpy(projection.numpy() @ extrinsic.numpy()),
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
