"""Oracle: the OpenPose-style condition image (f1).  TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy restatement of the reference's per-view condition stage: projection (core/human/smpl_condition.py:199-212,
utils/point3d.py:32-42), keypoint packing (to_controlnet_pose :20-79) and drawing (core/human/open_pose.py:48-333,
adaptive_draw_poses) with cv2's primitives replaced by analytic coverage tests (cv2.circle == dx^2 + dy^2 <= r^2 exactly;
ellipse2Poly + fillConvexPoly ~ ellipse with semi-axes + 0.5; thick line ~ capsule).  PINNED with a stated tolerance
against images drawn by the reference's own code (tests/golden/make_pose_golden.py -> pose.npz): boundary pixels of the
ellipses / lines may differ.  Embree occlusion culling is NOT restated (open3d absent); the CUDA path substitutes a
depth test against the rendered Gaussians, tested separately.
"""
import colorsys

import numpy as np

BODY_COLORS = np.array([[255, 0, 0], [255, 85, 0], [255, 170, 0], [255, 255, 0], [170, 255, 0], [85, 255, 0], [0, 255, 0], [0, 255, 85],
                        [0, 255, 170], [0, 255, 255], [0, 170, 255], [0, 85, 255], [0, 0, 255], [85, 0, 255], [170, 0, 255], [255, 0, 255],
                        [255, 0, 170], [255, 0, 85]], np.float32)
LIMBS = [[2, 3], [2, 6], [3, 4], [4, 5], [6, 7], [7, 8], [2, 9], [9, 10], [10, 11], [2, 12], [12, 13], [13, 14], [2, 1], [1, 15], [15, 17],
         [1, 16], [16, 18]]
FLIP = [0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16]
EDGES = [[0, 1], [1, 2], [2, 3], [3, 4], [0, 5], [5, 6], [6, 7], [7, 8], [0, 9], [9, 10], [10, 11], [11, 12], [0, 13], [13, 14], [14, 15],
         [15, 16], [0, 17], [17, 18], [18, 19], [19, 20]]


def hand_edge_colors():
    """rint(hsv_to_rgb(e / 20, 1, 1) * 255) as cv2 receives the float colour (open_pose.py:211; saturate_cast rounds half to even)."""
    return np.array([np.rint(np.array(colorsys.hsv_to_rgb(e / 20.0, 1.0, 1.0)) * 255.0) for e in range(20)], np.float32).astype(np.uint8)


def project(kp_world, extrinsic, K):
    """smpl_condition.py:205-212: world -> camera -> pixels; z < 0 => NaN."""
    cam = kp_world @ extrinsic[:3, :3].T + extrinsic[:3, 3]
    out = np.full((kp_world.shape[0], 2), np.nan, np.float64)
    ok = ~(cam[:, 2] < 0)
    h = cam[ok] @ K.T
    out[ok] = h[:, :2] / h[:, 2:3]
    return out


def draw(kp2d, H, W, draw_body=True, draw_hand=True, draw_face=False, flip_LR=False):
    """kp2d [128,2] pixels (NaN = absent) -> uint8 [H,W,3]."""
    kp = np.asarray(kp2d, np.float32)
    body_r, stick, hand_r, hand_th, face_r = 4, 4, 4, 2, 3
    if H != 512 or W != 512:
        r = (H + W) / 2.0 / 512.0
        body_r, stick, hand_r, hand_th, face_r = (max(int(v * r), 1) for v in (body_r, stick, hand_r, hand_th, face_r))
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W, 3), np.float32)
    ok = ~np.isnan(kp).any(1)
    ix = np.where(ok, np.nan_to_num(kp[:, 0]), -1).astype(np.int64)
    iy = np.where(ok, np.nan_to_num(kp[:, 1]), -1).astype(np.int64)
    drawn = ok & (ix > 0) & (iy > 0)
    if draw_body:
        for i in range(18):
            k = FLIP[i] if flip_LR else i
            if drawn[k]:
                img[(xx - ix[k]) ** 2 + (yy - iy[k]) ** 2 <= body_r * body_r] = BODY_COLORS[i]
        for l, (a1, a2) in enumerate(LIMBS):
            k1, k2 = a1 - 1, a2 - 1
            if flip_LR:
                k1, k2 = FLIP[k1], FLIP[k2]
            if not (ok[k1] and ok[k2]):
                continue
            Y0, Y1, X0, X1 = kp[k1, 0], kp[k2, 0], kp[k1, 1], kp[k2, 1]
            mX, mY = np.float32(0.5) * (X0 + X1), np.float32(0.5) * (Y0 + Y1)
            length = np.sqrt((X0 - X1) ** 2 + (Y0 - Y1) ** 2)
            ang = int(np.degrees(np.arctan2(X0 - X1, Y0 - Y1)))
            cx, cy, a = int(mY), int(mX), int(length / 2)
            t = np.radians(ang)
            u = (xx - cx) * np.cos(t) + (yy - cy) * np.sin(t)
            w = -(xx - cx) * np.sin(t) + (yy - cy) * np.cos(t)
            m = (u / (a + 0.5)) ** 2 + (w / (stick + 0.5)) ** 2 <= 1.0
            img[m] = np.rint(0.4 * img[m] + 0.6 * BODY_COLORS[l])
    if draw_hand:
        cols = hand_edge_colors().astype(np.float32)
        hw0 = 0.5 if hand_th <= 1 else float((hand_th + 1) // 2)
        for hnd in range(2):
            base = 18 + 21 * hnd
            for k in range(base, base + 21):
                if drawn[k]:
                    img[(xx - ix[k]) ** 2 + (yy - iy[k]) ** 2 <= hand_r * hand_r] = (0, 0, 255)
            for e, (e1, e2) in enumerate(EDGES):
                k1, k2 = base + e1, base + e2
                if not (ok[k1] and ok[k2]):
                    continue
                x1, y1, x2, y2 = int(kp[k1, 0]), int(kp[k1, 1]), int(kp[k2, 0]), int(kp[k2, 1])
                if not (x1 > 0 and y1 > 0 and x2 > 0 and y2 > 0):
                    continue
                ex, ey = float(x2 - x1), float(y2 - y1)
                L2 = ex * ex + ey * ey
                tt = np.clip(((xx - x1) * ex + (yy - y1) * ey) / L2, 0, 1) if L2 > 0 else np.zeros_like(xx, np.float64)
                qx, qy = (xx - x1) - tt * ex, (yy - y1) - tt * ey
                mn, mx = min(abs(ex), abs(ey)), max(abs(ex), abs(ey))
                hw = hw0 + (0.5 * mn / mx if mx > 0 else 0.0)
                img[qx * qx + qy * qy <= hw * hw] = cols[e]
    if draw_face:
        for k in range(60, 128):
            if drawn[k]:
                img[(xx - ix[k]) ** 2 + (yy - iy[k]) ** 2 <= face_r * face_r] = 255
    return img.astype(np.uint8)
