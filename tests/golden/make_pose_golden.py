"""Golden images for (f1) the condition-image producer: the reference's OWN drawing code
(core/human/open_pose.py draw_poses, loaded from its file: pure cv2 / numpy) fed by its own keypoint packing
(core/human/smpl_condition.py to_controlnet_pose, extracted with ``ast``) on seeded 2-D keypoints.
matplotlib (absent here) is only used for colors.hsv_to_rgb; a colorsys shim supplies the same function.
Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_pose_golden.py       -> tests/golden/pose.npz
"""
import ast
import colorsys
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def case_keypoints(seed, H=512, W=512, missing=()):
    """[128,2] pixel coordinates of a plausible front-view skeleton + jitter; NaN = not visible (smpl_condition.py:209,221)."""
    rng = np.random.default_rng(seed)
    s = H / 512.0
    body = np.array([[256, 100], [256, 140], [215, 142], [195, 205], [185, 265], [297, 142], [317, 205], [327, 265],
                     [232, 270], [228, 360], [226, 450], [280, 270], [284, 360], [286, 450], [247, 92], [265, 92], [236, 98], [276, 98]], np.float64)
    def hand(wrist, sign):
        pts = [wrist]
        for f in range(5):
            base = wrist + np.array([sign * (6 + 3 * f), 10 + 2 * f])
            for k in range(4):
                pts.append(base + np.array([sign * (f - 2) * 2.5 * (k + 1), 7.0 * (k + 1)]))
        return np.array(pts)
    lh, rh = hand(body[7], +1), hand(body[4], -1)
    face = np.stack([256 + 22 * np.cos(np.linspace(0, 2 * np.pi, 68, endpoint=False)) * np.linspace(0.3, 1.0, 68),
                     104 + 26 * np.sin(np.linspace(0, 2 * np.pi, 68, endpoint=False)) * np.linspace(0.3, 1.0, 68)], 1)
    kp = np.concatenate([body, lh, rh, face]) * s + rng.normal(0, 2.0 * s, size=(128, 2))
    kp[list(missing)] = np.nan
    return kp


CASES = {'front512': dict(seed=1, H=512, W=512, missing=()),
         'occluded512': dict(seed=2, H=512, W=512, missing=(4, 7, 10, 16, 25, 26, 60, 100)),
         'front768': dict(seed=3, H=768, W=768, missing=(17,)),
         'flip512': dict(seed=4, H=512, W=512, missing=(3,), flip=True)}


def main():
    mpl = types.ModuleType('matplotlib')
    mpl.colors = types.ModuleType('matplotlib.colors')
    mpl.colors.hsv_to_rgb = lambda hsv: np.array(colorsys.hsv_to_rgb(*hsv))
    sys.modules['matplotlib'], sys.modules['matplotlib.colors'] = mpl, mpl.colors
    spec = importlib.util.spec_from_file_location('ref_open_pose', os.path.join(REF, 'core/human/open_pose.py'))
    op = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(op)
    src = open(os.path.join(REF, 'core/human/smpl_condition.py')).read()
    ns = {'np': np, 'Keypoint': op.Keypoint, 'BodyResult': op.BodyResult, 'PoseResult': op.PoseResult}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == 'to_controlnet_pose':
            exec(compile(ast.Module(body=[node], type_ignores=[]), 'smpl_condition.py', 'exec'), ns)
    out = {}
    for name, c in CASES.items():
        kp = case_keypoints(c['seed'], c['H'], c['W'], c['missing'])
        K = np.array([[1.0, 0, c['W'] / 2], [0, 1.0, c['H'] / 2], [0, 0, 1]])
        poses = ns['to_controlnet_pose'](kp[None], intrinsics=K)
        # smpl_condition.py:8,226-234: `adaptive_draw_poses as draw_poses` (radii scale with the size), flip_LR from the config
        img = op.adaptive_draw_poses(poses, H=c['H'], W=c['W'], draw_body=True, draw_hand=True, draw_face=True, flip_LR=c.get('flip', False))
        out[f'{name}.kp'] = kp
        out[f'{name}.image'] = img
        print(name, img.shape, int((img.sum(-1) > 0).sum()), 'drawn pixels')
    np.savez_compressed(os.path.join(HERE, 'pose.npz'), **out)


if __name__ == '__main__':
    main()
