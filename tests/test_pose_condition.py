"""(f1) The ControlNet condition producer: oracle (oracle/pose.py) against images drawn by the REFERENCE'S OWN code
(open_pose.py adaptive_draw_poses + to_controlnet_pose, tests/golden/make_pose_golden.py -> pose.npz), and the CUDA kernels
(dwg_pose_keypoints_2d / dwg_pose_image) against both.  Stated tolerance vs cv2: filled circles are exact; the ellipse
polygons / thick lines are matched analytically, so at most 6 % of the DRAWN pixels (boundary pixels; < 0.3 % of the image)
may differ, and the masks overlap with IoU >= 0.95."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'dreamwaltz-g_b200'))
from oracle import pose as opose  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'pose.npz'))
CASES = sorted({k.split('.')[0] for k in G.files})


def _compare(img, gold):
    drawn = (gold.sum(-1) > 0) | (img.sum(-1) > 0)
    both = (gold.sum(-1) > 0) & (img.sum(-1) > 0)
    iou = both.sum() / max(drawn.sum(), 1)
    differ = (np.abs(img.astype(np.int32) - gold.astype(np.int32)).max(-1) > 2)
    return float(iou), float(differ.sum() / max((gold.sum(-1) > 0).sum(), 1))


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_drawing(name):
    gold, kp = G[f'{name}.image'], G[f'{name}.kp']
    H, W = gold.shape[:2]
    img = opose.draw(kp, H, W, draw_body=True, draw_hand=True, draw_face=True, flip_LR=name.startswith('flip'))
    iou, frac = _compare(img, gold)
    assert iou >= 0.95 and frac <= 0.06, (iou, frac)


def test_hand_edge_colours_and_projection():
    c = opose.hand_edge_colors()
    assert c.shape == (20, 3) and tuple(c[0]) == (255, 0, 0) and c.dtype == np.uint8
    ext = np.eye(4); ext[:3, 3] = (0.1, -0.2, 2.0)
    K = np.array([[500.0, 0, 256], [0, -500.0, 256], [0, 0, 1]])
    kp = np.array([[0.0, 0.0, 0.0], [0.1, 0.2, 0.5], [0.0, 0.0, -3.0]])
    p = opose.project(kp, ext, K)
    np.testing.assert_allclose(p[0], [500 * 0.1 / 2.0 + 256, -500 * -0.2 / 2.0 + 256])
    assert np.isnan(p[2]).all()                                  # behind the camera


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_pose_image_matches_oracle_and_reference(name):
    import torch
    from dwg import condition
    gold, kp = G[f'{name}.image'], G[f'{name}.kp']
    H, W = gold.shape[:2]
    prod = condition.PoseConditionProducer(H, W, device='cuda', draw_face=True, flip_LR=name.startswith('flip'))
    out = prod.draw(torch.from_numpy(kp.astype(np.float32)).cuda())
    assert out.shape == (1, 3, H, W) and out.dtype == torch.float32
    img = np.rint(out[0].permute(1, 2, 0).cpu().numpy() * 255.0).astype(np.uint8)
    ref = opose.draw(kp, H, W, draw_body=True, draw_hand=True, draw_face=True, flip_LR=name.startswith('flip'))
    assert (np.abs(img.astype(np.int32) - ref.astype(np.int32)).max(-1) > 0).mean() < 2e-4          # same predicates: float32 vs float64 boundary ties only
    iou, frac = _compare(img, gold)
    assert iou >= 0.95 and frac <= 0.06, (iou, frac)


@pytest.mark.gpu
def test_cuda_projection_and_depth_occlusion():
    import torch
    from dwg import camera, condition
    H = W = 512
    d = camera.make_camera(2.0, 30.0, 85.0, 50.0, H, W)
    rng = np.random.default_rng(3)
    kpw = rng.normal(0, 0.3, size=(128, 3)).astype(np.float32)
    prod = condition.PoseConditionProducer(H, W, device='cuda')
    fx, fy, cx, cy = prod.intrinsics(d)
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    ref = opose.project(kpw.astype(np.float64), d['extrinsic'][0].numpy().astype(np.float64), K)
    got = prod.project(torch.from_numpy(kpw).cuda(), d).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-3)
    # occlusion: an opaque render 0.5 in front of every keypoint hides the body / hand points (thresholds 0.2) and the face points (0.02)
    ext = d['extrinsic'][0].numpy()
    z = (kpw @ ext[:3, :3].T + ext[:3, 3])[:, 2]
    depth = torch.full((H, W), float(z.min() - 0.5), device='cuda')
    alpha = torch.ones(H, W, device='cuda')
    occ = prod.project(torch.from_numpy(kpw).cuda(), d, depth=depth, alpha=alpha).cpu().numpy()
    inside = (ref[:, 0] >= 0) & (ref[:, 0] < W) & (ref[:, 1] >= 0) & (ref[:, 1] < H)
    assert np.isnan(occ[inside]).all()
    transparent = prod.project(torch.from_numpy(kpw).cuda(), d, depth=depth, alpha=torch.zeros(H, W, device='cuda')).cpu().numpy()
    np.testing.assert_allclose(transparent, got)
    cond = prod(torch.from_numpy(kpw).cuda(), d)
    assert cond.shape == (1, 3, H, W) and float(cond.max()) <= 1.0 and float(cond.sum()) > 0
