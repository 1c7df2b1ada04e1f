"""(f1) The per-view ControlNet condition image, produced on the GPU.

Mirrors the reference's producer surface -- SMPLXPrompt.get_cond_images -> SMPL2Condition.__call__(condition_type='pose')
(core/human/smpl_prompt.py:229-263, core/human/smpl_condition.py:271-320,191-235) followed by prepare_image
(core/guidance/controlnet.py:33-55) -- but the result is the [1,3,H,W] float tensor the guidance consumes, already on
the device, rendered directly at the condition size (no PIL image, no LANCZOS resize, no host round trip):
    cond = producer(keypoints_world [128,3], data)                # data = the camera dict of the view
Occlusion culling uses the view's own rendered depth / alpha (pass render_outputs) instead of Embree ray casting
against the template mesh (utils/open3d.py:8-45): see csrc/pose_image.cu.
"""
import colorsys

import numpy as np
import torch

from ._lib import check, lib, ptr, stream


def _hand_edge_colors():
    return np.array([np.rint(np.array(colorsys.hsv_to_rgb(e / 20.0, 1.0, 1.0)) * 255.0) for e in range(20)], np.float32).astype(np.uint8)


class PoseConditionProducer:
    def __init__(self, height=512, width=512, device='cuda', draw_body=True, draw_hand=True, draw_face=False, flip_LR=False,
                 thres_body=0.2, thres_face=0.02, thres_hand=0.2):
        self.H, self.W, self.dev = int(height), int(width), device
        self.flags = (1 if draw_body else 0) | (2 if draw_hand else 0) | (4 if draw_face else 0) | (8 if flip_LR else 0)     # configs/__init__.py:442-446
        self.thres = (float(thres_body), float(thres_face), float(thres_hand))                                            # smpl_condition.py:101-104
        self.hand_colors = torch.from_numpy(_hand_edge_colors()).to(device)
        self._half = torch.tensor([self.W / 2.0, self.H / 2.0], device=device, dtype=torch.float32)      # (cx, cy); built once (graph-safe)

    def intrinsics(self, data):
        """to_intrinsics (data/camera/utils.py:116-147) at the render size, then adjust_intrinsics_size (:233-242) to the
        condition size: returns (fx, fy, cx, cy); fy < 0 (image y points down)."""
        tanfov = float(data['tanfov'][0])
        h, w = data['image_height'], data['image_width']
        f = h / (2.0 * tanfov)
        width_raw, height_raw = float(h // 2) * 2.0, float(w // 2) * 2.0        # principal point = (h // 2, w // 2), as the reference
        return f * self.W / width_raw, -f * self.H / height_raw, self.W / 2.0, self.H / 2.0

    def project(self, keypoints_world, data, depth=None, alpha=None, cam_dev=None):
        """[K,3] world keypoints -> [K,2] pixel coordinates at the condition size (NaN = behind the camera / occluded).
        cam_dev: the device-resident camera struct of ops.pack_camera (overrides data's extrinsic / fov: graph replay)."""
        kp = keypoints_world.to(self.dev, torch.float32).contiguous()
        K = kp.shape[0]
        intr = None
        if cam_dev is not None:
            ext = cam_dev[4:20].view(4, 4).t().contiguous()                       # viewmatrix = extrinsic^T
            h, w = data['image_height'], data['image_width']
            f = (h * 0.5) / cam_dev[3:4]                                          # h / (2 tanfov_y)
            intr = torch.cat([f * (self.W / (float(h // 2) * 2.0)), f * (-self.H / (float(w // 2) * 2.0)), self._half]).contiguous()
            fx = fy = cx = cy = 0.0
        else:
            ext = data['extrinsic'][0].to(self.dev, torch.float32).contiguous()
            fx, fy, cx, cy = self.intrinsics(data)
        out = torch.empty(K, 2, device=self.dev, dtype=torch.float32)
        Hd = Wd = 0
        if depth is not None:
            depth, alpha = depth.to(torch.float32).squeeze().contiguous(), alpha.to(torch.float32).squeeze().contiguous()     # [1,H,W,1] -> [H,W]
            assert depth.dim() == 2 and depth.shape == alpha.shape
            Hd, Wd = depth.shape
        check(lib().dwg_pose_keypoints_2d(ptr(kp), K, ptr(ext), ptr(intr), fx, fy, cx, cy, ptr(depth), ptr(alpha), Hd, Wd, float(self.W), float(self.H),
                                          *self.thres, ptr(out), stream()), 'dwg_pose_keypoints_2d')
        return out

    def draw(self, kp2d):
        """[128,2] pixel keypoints -> [1,3,H,W] float in [0,1] (the tensor prepare_image returns)."""
        kp2d = kp2d.to(self.dev, torch.float32).contiguous()
        assert kp2d.shape == (128, 2), 'expected the 128-keypoint OpenPose layout (smpl_condition.py:22)'
        out = torch.empty(1, 3, self.H, self.W, device=self.dev, dtype=torch.float32)
        check(lib().dwg_pose_image(ptr(kp2d), self.H, self.W, self.flags, ptr(self.hand_colors), ptr(out), stream()), 'dwg_pose_image')
        return out

    def __call__(self, keypoints_world, data, render_outputs=None, cam_dev=None):
        depth = alpha = None
        if render_outputs is not None:
            depth, alpha = render_outputs['depth'].detach(), render_outputs['alpha'].detach()
        return self.draw(self.project(keypoints_world, data, depth=depth, alpha=alpha, cam_dev=cam_dev))


class KeypointSource:
    """The 128 OpenPose keypoints of a posed SMPL-X body on the device: model joints (posed by dwg_glbs_joints) plus joints
    defined on mesh vertices (nose, eyes, ears, finger tips, face landmarks: smplx VertexJointSelector / landmarks), posed by
    dwg_glbs_vertices over the same per-vertex tables the mesh-bound Gaussians use.  ``joint_ids`` / ``vertex_ids``: for each
    of the 128 slots either a model-joint index or a template-vertex index (the other = -1), i.e. the permutation that
    smpl_to_openpose (core/human/smpl_utils.py:77-200) applies to the smplx joint list."""

    def __init__(self, lbs_model, joint_ids, vertex_ids):
        dev = lbs_model.v_template.device
        joint_ids, vertex_ids = np.asarray(joint_ids, np.int64), np.asarray(vertex_ids, np.int64)
        assert joint_ids.shape == vertex_ids.shape == (128,) and np.all((joint_ids >= 0) != (vertex_ids >= 0))
        self.lbs = lbs_model
        self.is_joint = torch.from_numpy(joint_ids >= 0).to(dev)
        self.jidx = torch.from_numpy(np.maximum(joint_ids, 0)).to(dev)
        vsel = np.unique(vertex_ids[vertex_ids >= 0])
        self.vslot = torch.from_numpy(np.searchsorted(vsel, np.maximum(vertex_ids, vsel.min() if vsel.size else 0))).to(dev)
        vi = torch.from_numpy(vsel).to(dev)
        ns, V = lbs_model._shapedirs_full.shape[1], lbs_model.v_template.shape[0]
        self.points = lbs_model.v_template.data[vi].contiguous()
        self.sdirs = lbs_model._shapedirs_full.view(V, 3, ns)[vi].contiguous()
        self.pdirs = lbs_model.posedirs.data.view(-1, V, 3)[:, vi].permute(1, 2, 0).contiguous()
        self.w = lbs_model.lbs_weights.data[vi].contiguous()

    @torch.no_grad()
    def __call__(self, jt):
        """jt = lbs_model.joint_transforms(**smpl_inputs) -> [128,3] world keypoints."""
        from . import ops
        verts = ops.glbs_vertices(jt, self.sdirs, self.pdirs, self.w, self.points)
        return torch.where(self.is_joint[:, None], jt['posed_joints'][self.jidx], verts[self.vslot])


def synthetic_keypoint_source(lbs_model, body_model):
    """Keypoint definition for the synthetic SMPL-X-shaped body of dwg.synth (the real asset's vertex ids / landmark files are
    not available): coco18 body from model joints + head-surface vertices, the 21-point hands from the finger joints + tip
    vertices, 68 face points on head vertices."""
    vj = body_model['vertex_joint'].numpy()
    vt = body_model['v_template'].numpy()
    J = body_model['J_template'].numpy()

    def nearest_vertex(target, joint):
        cand = np.nonzero(vj == joint)[0]
        if cand.size == 0:
            cand = np.arange(vt.shape[0])
        return int(cand[np.argmin(((vt[cand] - target) ** 2).sum(1))])
    head = J[15]
    jid, vid = -np.ones(128, np.int64), -np.ones(128, np.int64)
    body = [None, 12, 17, 19, 21, 16, 18, 20, 2, 5, 8, 1, 4, 7, None, None, None, None]      # smpl_utils.py:183-190 pattern (nose / eyes / ears on vertices)
    extras = {0: head + (0, 0.0, 0.11), 14: head + (-0.03, 0.04, 0.09), 15: head + (0.03, 0.04, 0.09), 16: head + (-0.08, 0.02, 0.0), 17: head + (0.08, 0.02, 0.0)}
    for i, j in enumerate(body):
        if j is None:
            vid[i] = nearest_vertex(extras[i], 15)
        else:
            jid[i] = j
    for hnd, (wrist, first) in enumerate(((20, 25), (21, 40))):
        base = 18 + 21 * hnd
        jid[base] = wrist
        order = [4, 0, 1, 2, 3]                                    # thumb, index, middle, ring, pinky in the smplx finger-joint blocks
        for f, blk in enumerate(order):
            j0 = first + 3 * blk
            for k in range(3):
                jid[base + 1 + 4 * f + k] = j0 + k
            tip = J[j0 + 2] + (J[j0 + 2] - J[j0 + 1])
            vid[base + 1 + 4 * f + 3] = nearest_vertex(tip, j0 + 2)
    ang = np.linspace(0, 2 * np.pi, 68, endpoint=False)
    for i in range(68):
        vid[60 + i] = nearest_vertex(head + (0.06 * np.cos(ang[i]), 0.02 + 0.07 * np.sin(ang[i]), 0.10), 15)
    return KeypointSource(lbs_model, jid, vid)
