"""Row R0 pinned: dwg/camera.py against the reference's OWN camera functions (data/camera/utils.py:62-201), executed by
tests/golden/make_camera_golden.py -> tests/golden/camera.npz.  fp32 tolerance: the reference inverts c2w numerically
(torch.inverse), dwg writes the rigid inverse in closed form."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dreamwaltz-g_b200'))
from dwg import camera  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'camera.npz'))


def test_extrinsic_projection_match_reference_functions():
    ext, c2w = camera.to_extrinsic(torch.from_numpy(G['radius']), torch.from_numpy(G['azimuth']), torch.from_numpy(G['elevation']),
                                   at_vector=torch.from_numpy(G['at']))
    np.testing.assert_allclose(c2w.numpy(), G['c2w'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(ext.numpy(), G['extrinsic'], rtol=0, atol=5e-6)
    proj = camera.to_projection(torch.from_numpy(G['tanfov']), 0.01, 1000.0)
    np.testing.assert_allclose(proj.numpy(), G['projection'], rtol=2e-7, atol=0)
    np.testing.assert_allclose(proj.numpy() @ ext.numpy(), G['mvp'], rtol=0, atol=2e-5)


def test_data_dict_of_one_view():
    i = 5
    d = camera.make_camera(G['radius'][i], G['azimuth'][i], G['elevation'][i], G['fov'][i], 512, 512, at=tuple(G['at'][i]))
    np.testing.assert_allclose(d['tanfov'].numpy(), G['tanfov'][i:i + 1], rtol=3e-7)
    np.testing.assert_allclose(d['extrinsic'].numpy()[0], G['extrinsic'][i], atol=5e-6)
    np.testing.assert_allclose(d['mvp'].numpy()[0], G['mvp'][i], atol=2e-5)
    view, proj, campos, tfx, tfy = camera.raster_matrices(d)                    # gaussian_renderer.py:23-40
    np.testing.assert_allclose(view.numpy(), G['extrinsic'][i].T, atol=5e-6)
    np.testing.assert_allclose(proj.numpy(), (G['projection'][i] @ G['extrinsic'][i]).T, atol=2e-5)
    np.testing.assert_allclose(campos.numpy(), G['c2w'][i][:3, 3], atol=2e-6)
    assert d['extrinsic'].dtype == torch.float32 and d['image_height'] == 512
