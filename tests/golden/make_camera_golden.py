"""Golden vectors for row R0 (the camera ``data`` dict) by EXECUTING THE REFERENCE'S OWN functions
(data/camera/utils.py: safe_normalize, get_tan_half_fov, angle2sphere, to_extrinsic, to_intrinsics, to_projection),
extracted with ``ast`` because the package __init__ imports smplx.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_camera_golden.py      -> tests/golden/camera.npz
"""
import ast
import os

import numpy as np
import torch
from torch import Tensor
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/data/camera/utils.py'
NAMES = ('safe_normalize', 'get_tan_half_fov', 'angle2sphere', 'to_extrinsic', 'to_intrinsics', 'to_projection')


def main():
    src = open(SRC).read()
    ns = {'torch': torch, 'Tensor': Tensor, 'Optional': Optional, 'np': np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in NAMES:
            exec(compile(ast.Module(body=[node], type_ignores=[]), SRC, 'exec'), ns)
    rng = np.random.default_rng(42)
    B = 24
    radius = torch.tensor(rng.uniform(1.0, 2.0, B), dtype=torch.float32)
    azimuth = torch.tensor(rng.uniform(0.0, 360.0, B), dtype=torch.float32)
    elevation = torch.tensor(rng.uniform(60.0, 120.0, B), dtype=torch.float32)
    fov = torch.tensor(rng.uniform(40.0, 70.0, B), dtype=torch.float32)
    at = torch.tensor(rng.uniform(-0.3, 0.3, (B, 3)), dtype=torch.float32)
    azimuth[:3] = torch.tensor([0.0, 90.0, 180.0]); elevation[:3] = torch.tensor([90.0, 60.0, 120.0])
    tanfov = ns['get_tan_half_fov'](fov)
    extrinsic, c2w = ns['to_extrinsic'](radius=radius, azimuth=azimuth, elevation=elevation, at_vector=at)
    projection = ns['to_projection'](tanfov=tanfov, z_far=1000.0, z_near=0.01)
    intrinsics = ns['to_intrinsics'](tanfov=tanfov, image_height=512, image_width=512)
    mvp = torch.bmm(projection, extrinsic)
    np.savez_compressed(os.path.join(HERE, 'camera.npz'), radius=radius.numpy(), azimuth=azimuth.numpy(), elevation=elevation.numpy(),
                        fov=fov.numpy(), at=at.numpy(), tanfov=tanfov.numpy(), extrinsic=extrinsic.numpy(), c2w=c2w.numpy(),
                        projection=projection.numpy(), intrinsics=intrinsics.numpy(), mvp=mvp.numpy())
    print('wrote camera.npz', B)


if __name__ == '__main__':
    main()
