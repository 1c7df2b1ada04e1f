"""(f4 / cfg5) Re-enactment inference: animate -> rasterise (+ background composite) -> uint8 frames, per pose of a motion
sequence.  Mirrors the loop of Trainer.evaluate (core/trainer.py:1019-1112) for scripts/inference_reenact.sh: render
with bg_mode / video background, post-process 'image', 'image_fg' (+alpha), 'depth' (/3), 'alpha', hand uint8 frames to
the writer.  Here the whole frame (animate + raster + frame_pack) is ONE CUDA graph with static pose / camera buffers,
and finished frames leave through a ring of pinned host buffers (asynchronous D2H; the host callback sees frame i while
the GPU renders frame i+1..).  No gradient state is kept (torch.no_grad: the fused MLP kernel stores no activations).
"""
import torch

from . import camera as dcam
from . import ops
from ._lib import check, lib, ptr, stream


def frame_pack(image_chw, image_fg_chw=None, depth=None, alpha=None, depth_div=3.0, out=None):
    """Planar fp32 render outputs -> dict of interleaved uint8 frames (dwg_frame_pack)."""
    H, W = image_chw.shape[-2:]
    dev = image_chw.device
    out = out if out is not None else {}
    if 'image' not in out:
        out['image'] = torch.empty(H, W, 3, device=dev, dtype=torch.uint8)
    if image_fg_chw is not None and alpha is not None and 'image_fg' not in out:
        out['image_fg'] = torch.empty(H, W, 4, device=dev, dtype=torch.uint8)
    if depth is not None and 'depth' not in out:
        out['depth'] = torch.empty(H, W, device=dev, dtype=torch.uint8)
    if alpha is not None and 'alpha' not in out:
        out['alpha'] = torch.empty(H, W, device=dev, dtype=torch.uint8)
    c = lambda t: None if t is None else t.contiguous()
    check(lib().dwg_frame_pack(ptr(c(image_chw)), ptr(c(image_fg_chw)), ptr(c(depth)), ptr(c(alpha)), ptr(out['image']),
                               ptr(out.get('image_fg')) if image_fg_chw is not None else None, ptr(out.get('depth')) if depth is not None else None,
                               ptr(out.get('alpha')) if alpha is not None else None, H, W, float(depth_div), stream()), 'dwg_frame_pack')
    return out


class Reenactor:
    def __init__(self, scene, bg_mode='white', outputs=('image', 'image_fg', 'depth', 'alpha'), ring=4):
        self.scene, self.bg_mode, self.outputs = scene, bg_mode, tuple(outputs)
        self.dev = next(scene.parameters()).device
        self._graph = None
        self._ring_n = ring

    @torch.no_grad()
    def render(self, data, cam_dev=None, out=None):
        """One frame, eager: -> dict of uint8 device tensors (keys = self.outputs)."""
        ro = self.scene(data, smpl_observed_inputs=data.get('smpl_inputs'), use_densifier=False, bg_mode=self.bg_mode, cam_dev=cam_dev)
        chw = lambda t: t[0].permute(2, 0, 1)
        fg = chw(ro['image_fg']) if 'image_fg' in self.outputs else None
        return frame_pack(ro['image_chw'], fg, ro['depth'][0, :, :, 0] if 'depth' in self.outputs else None,
                          ro['alpha'][0, :, :, 0] if ('alpha' in self.outputs or fg is not None) else None, out=out)

    def capture(self, data, warmup=2):
        dev = self.dev
        pose0 = data['smpl_inputs']
        self._layout, off = [], ops.CAMERA_WORDS
        for k, v in pose0.items():
            self._layout.append((k, off, v.numel(), tuple(v.shape)))
            off += (v.numel() + 3) // 4 * 4
        self._packed = torch.zeros(off, device=dev)
        self._pose = {k: self._packed[o:o + n].view(shape) for k, o, n, shape in self._layout}
        self._cam = self._packed[:ops.CAMERA_WORDS]
        self._template = {k: v for k, v in data.items() if k != 'smpl_inputs'}
        self._host_in = [torch.zeros(off).pin_memory() for _ in range(16)]
        self._in_ev = [None] * 16
        self._in_i = 0
        self._send(data)
        sdata = dict(self._template, smpl_inputs=self._pose)
        self._out = {}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.render(sdata, cam_dev=self._cam, out=self._out)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = lib().launches
        with torch.cuda.graph(g):
            self.render(sdata, cam_dev=self._cam, out=self._out)
        self.graph_launches = lib().launches - n0
        self._graph = g
        # pinned output ring
        self._host_out = [{k: torch.empty_like(v, device='cpu').pin_memory() for k, v in self._out.items()} for _ in range(self._ring_n)]
        self._out_ev = [torch.cuda.Event() for _ in range(self._ring_n)]
        return self

    def _send(self, data):
        self._in_i = (self._in_i + 1) % len(self._host_in)
        if self._in_ev[self._in_i] is not None:
            self._in_ev[self._in_i].synchronize()
        h = self._host_in[self._in_i]
        view, proj, campos, tfx, tfy = dcam.raster_matrices(data)
        ops.pack_camera(data['image_height'], data['image_width'], tfx, tfy, view, proj, self.scene.renderer.bg_color, 1.0, out=h[:ops.CAMERA_WORDS])
        for k, o, n, shape in self._layout:
            h[o:o + n].copy_(data['smpl_inputs'][k].reshape(-1))
        self._packed.copy_(h, non_blocking=True)
        if self._in_ev[self._in_i] is None:
            self._in_ev[self._in_i] = torch.cuda.Event()
        self._in_ev[self._in_i].record()

    def run(self, frames, on_frame=None):
        """frames: iterable of data dicts (host pose tensors).  on_frame(i, {name: pinned uint8 host tensor}) is called
        once frame i has landed in host memory (one ring slot behind the GPU).  Returns the number of frames."""
        assert self._graph is not None, 'capture() first'
        L = lib()
        pending = []
        n = 0
        for i, data in enumerate(frames):
            slot = i % self._ring_n
            if len(pending) == self._ring_n:                     # slot about to be reused: deliver its frame first
                j, s = pending.pop(0)
                self._out_ev[s].synchronize()
                if on_frame is not None:
                    on_frame(j, self._host_out[s])
            self._send(data)
            self._graph.replay()
            object.__setattr__(L, 'launches', L.launches + self.graph_launches)
            for k, v in self._out.items():
                self._host_out[slot][k].copy_(v, non_blocking=True)
            self._out_ev[slot].record()
            pending.append((i, slot))
            n += 1
        for j, s in pending:
            self._out_ev[s].synchronize()
            if on_frame is not None:
                on_frame(j, self._host_out[s])
        return n
