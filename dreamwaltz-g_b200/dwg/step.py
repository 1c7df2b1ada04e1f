"""R13 / R17: the per-step glue of the 3DGS SDS stage as a package API.

Mirrors the reference's call surfaces so that its trainer can hold these objects instead of its own:

  Scene.forward(data, smpl_observed_inputs, use_densifier, bg_mode)      core/system/scene.py:95-168
      avatar.animate -> renderer.render -> background composite (image_fg / image_bg / image),
      the composite fused into the rasteriser's blend epilogue (ops.rasterize(bg_image=...)).
  SDSTrainStep.render(data, bg_mode)                                      core/trainer.py:680-705 (stage 'gs')
  SDSTrainStep.train_forward(data) -> (total_loss, render_outputs, sd_outputs, text)     core/trainer.py:933-1017
  SDSTrainStep.step(data) -> same tuple, after zero-grad + backward (+ gradient all-reduce, + optimiser)
                                                                          core/trainer.py:856-891
The whole step (animate -> raster -> VAE -> ControlNet + UNet -> SDS gradient -> backward to every avatar parameter)
can be captured ONCE as a CUDA graph (`capture`); `step` then only refreshes the static input buffers (pose,
device-resident camera struct, prompt embeddings, condition image) from pinned host memory and replays.
No host synchronisation anywhere in the step.
"""
import time

import torch

from . import camera as dcam
from . import ops
from .avatar import GaussianOutput, GaussianRenderer
from .parallel import GradBucket

PURE_COLORS = {'black': 0.0, 'white': 1.0, 'gray': 0.5}      # core/system/background.py:14-27


class _PinnedRing:
    """Pre-packed pinned staging for the per-step inputs (camera struct + pose vectors in ONE buffer, one H2D copy per
    step).  The CPU runs several steps ahead of the GPU, so a single pinned buffer would be overwritten before its
    asynchronous copy has executed: a ring of slots, each guarded by an event, keeps every in-flight copy intact."""

    def __init__(self, words, device, slots=16):
        self.host = [torch.zeros(words).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.i = 0
        self.device = device

    def next(self):
        self.i = (self.i + 1) % len(self.host)
        ev = self.events[self.i]
        if ev is not None:
            ev.synchronize()                                    # normally long complete: the ring is deeper than the launch queue
        return self.host[self.i]

    def sent(self):
        ev = self.events[self.i]
        if ev is None:
            ev = self.events[self.i] = torch.cuda.Event()
        ev.record()


class Scene(torch.nn.Module):
    """core/system/scene.py: one avatar + renderer + background.  ``background``: None, or a callable
    ``(data, shape) -> image_bg [1,H,W,3]`` (the reference's MLPBackground / VideoBackground surface)."""

    def __init__(self, avatar, renderer=None, background=None, avatar_transl=None, avatar_scale=None):
        super().__init__()
        self.avatar = avatar
        self.renderer = renderer or GaussianRenderer()
        self.background = background
        self.avatar_transl, self.avatar_scale = avatar_transl, avatar_scale
        self._pure = {}

    def avatar_forward(self, smpl_observed_inputs=None) -> GaussianOutput:
        gs = self.avatar.animate(smpl_observed_inputs)          # avatar.forward() == animate(canonical inputs) in the shipped config
        if self.avatar_scale is not None:
            gs.positions = gs.positions * self.avatar_scale.unsqueeze(0)
            gs.scales = gs.scales * self.avatar_scale.unsqueeze(0)
        if self.avatar_transl is not None:
            gs.positions = gs.positions + self.avatar_transl.unsqueeze(0)
        return gs

    def _pure_bg(self, bg_mode, H, W, device):
        key = (bg_mode, H, W, str(device))
        if key not in self._pure:
            self._pure[key] = torch.full((3, H, W), PURE_COLORS[bg_mode], device=device, dtype=torch.float32)
        return self._pure[key]

    def forward(self, data, smpl_observed_inputs=None, use_densifier=True, bg_mode=None, cam_dev=None, **_):
        gs = self.avatar_forward(smpl_observed_inputs)
        H, W = data['image_height'], data['image_width']
        bg_image = None
        if bg_mode in PURE_COLORS:
            bg_image = self._pure_bg(bg_mode, H, W, gs.positions.device)
        elif self.background is not None:
            bg_image = self.background(data, (1, H, W, 3))
        out = self.renderer.render(data, gs, return_2d_radii=use_densifier, cam_dev=cam_dev, bg_image=bg_image)
        if bg_image is not None:
            out['image_bg'] = bg_image.permute(1, 2, 0).unsqueeze(0) if bg_image.dim() == 3 else bg_image
        else:
            out['image_fg'] = out['image']                      # scene.py:166
        return out


class SDSTrainStep:
    """trainer.train_forward + the gradient half of the training loop, on dwg objects."""

    def __init__(self, scene, guidance, text_embeds_dict, lambda_guidance=1.0, optimizer=None, allreduce=True, max_step=1,
                 cond_producer=None, keypoint_source=None):
        self.scene, self.guidance = scene, guidance
        self.text_embeds_dict = text_embeds_dict
        self.lambda_guidance = lambda_guidance
        self.optimizer = optimizer
        self.allreduce = allreduce
        # (f1) condition image produced on the device from the posed body's keypoints and the view's own depth / alpha
        self.cond_producer, self.keypoint_source = cond_producer, keypoint_source
        self.train_step, self.max_step = 0, max_step
        self.params = [p for p in scene.parameters() if p.requires_grad]
        self.bucket = GradBucket(self.params)
        self.dev = self.params[0].device
        # Eager steps and the graph capture run on ONE private stream: autograd binds each parameter's AccumulateGrad node to the
        # stream it was created on, and a node bound to the legacy default stream would break a later capture on another stream.
        self._stream = torch.cuda.Stream(device=self.dev)
        self._graph = None
        self._result_dev, self._res_host, self._res_ev = None, None, None
        self.graph_launches = 0
        self.host_ms = {}
        self.fixed_draws = None         # tests: {'timestep', 'noise', 'vae_eps'} tensors instead of the guidance's own random draws

    # ---- core/trainer.py:680-705
    def render(self, data, bg_mode=None, cam_dev=None):
        return self.scene(data, smpl_observed_inputs=data.get('smpl_inputs'), use_densifier=False, bg_mode=bg_mode, cam_dev=cam_dev)

    # ---- core/trainer.py:933-1017 (stage 'gs'; no text augmentation / sparsity loss in the shipped script)
    def train_forward(self, data, cam_dev=None):
        produce = self.cond_producer is not None
        # head start on the second stream (no-op under sub-graphs); a produced condition depends on the render and is embedded later
        self.guidance.prepare(self.text_embeds_dict, None if produce else data.get('cond_images'))
        render_outputs = self.render(data, cam_dev=cam_dev)
        if produce:
            jt = self.scene.avatar.lbs_model.joint_transforms(**data['smpl_inputs'])
            cond = self.cond_producer(self.keypoint_source(jt), data, render_outputs, cam_dev=cam_dev)
            data = dict(data, cond_images=cond)
            render_outputs['cond_images'] = cond
        sd_inputs = render_outputs['image_chw'].unsqueeze(0)                       # == image.permute(0,3,1,2).contiguous() without the copy
        sd_outputs = self.guidance(inputs=sd_inputs, text_embeds_dict=self.text_embeds_dict, train_step=self.train_step,
                                   max_iteration=self.max_step, cond_inputs=data.get('cond_images'), **(self.fixed_draws or {}))
        total_loss = sd_outputs['diffusion_loss'] * self.lambda_guidance
        render_outputs['regularizations'] = {}
        return total_loss, render_outputs, sd_outputs, None

    def _body(self, data, cam_dev=None):
        self.bucket.zero()                                      # optimizer.zero_grad(): grads are views of one flat buffer
        loss, ro, so, text = self.train_forward(data, cam_dev)
        loss.backward()
        return loss, ro, so, text

    # ---- the step's scalar results on the host, without stalling the pipeline
    @staticmethod
    def _result_vector(loss, so):
        with torch.no_grad():
            return torch.stack([loss.detach().float().reshape(()), so['gradients'].abs().mean().float(), so['timestep'].reshape(-1)[0].float()])

    def _queue_result(self):
        """Enqueue the asynchronous D2H copy of (loss, mean |SDS gradient|, timestep) of the step just enqueued into a pinned ring."""
        if self._result_dev is None:
            return
        if self._res_host is None:
            self._res_host = [torch.zeros(3).pin_memory() for _ in range(8)]
            self._res_ev = [torch.cuda.Event() for _ in range(8)]
        i = self.train_step % 8
        self._res_host[i].copy_(self._result_dev, non_blocking=True)
        self._res_ev[i].record()

    def fetch_result(self, lag=0):
        """(loss, mean |SDS gradient|, timestep) of the step issued ``lag`` steps ago (0 = the latest: waits for it).  Reading one
        step late (lag=1) never stalls: host input preparation of the next step overlaps the device (the reference's loop,
        core/trainer.py:856-891, does not read the loss back inside a step either)."""
        assert 0 <= lag < 8 and self._res_host is not None
        i = (self.train_step - lag) % 8
        self._res_ev[i].synchronize()
        v = self._res_host[i]
        return float(v[0]), float(v[1]), int(v[2])

    def _post(self):
        if self.allreduce:
            self.bucket.all_reduce()                            # ONE collective over the flat gradient buffer
        if self.optimizer is not None:
            self.optimizer.step()

    def step(self, data):
        """One SDS step on ``data`` (camera dict + 'smpl_inputs' + 'cond_images').  Replays the captured graph when
        capture() was called (data then only refreshes the static buffers)."""
        self.train_step += 1
        if self._graph is not None:
            return self._replay(data)
        cur = torch.cuda.current_stream()
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            out = self._body(data)
            self._result_dev = self._result_vector(out[0], out[2])
        cur.wait_stream(self._stream)
        self._queue_result()
        self._post()
        return out

    # ---- whole-step CUDA graph
    def capture(self, data, warmup=3):
        """Capture the whole step (on the private stream the eager steps use as well)."""
        g = self.guidance
        if getattr(g, 'time_sampling', 'uniform') != 'uniform' and not self.fixed_draws:
            raise NotImplementedError("whole-step capture draws the timestep inside the graph: only time_sampling='uniform' (the shipped "
                                      "setting) can be captured; the iteration-dependent modes run through the eager step")
        g._g = None
        g.use_default_generator = True                          # graph-safe philox state
        dev = self.dev
        # ONE static device buffer holds the camera struct and every pose vector (views into it feed the graph)
        pose0 = data['smpl_inputs']
        self._pose_layout, off = [], ops.CAMERA_WORDS
        for k, v in pose0.items():
            self._pose_layout.append((k, off, v.numel(), tuple(v.shape)))
            off += (v.numel() + 3) // 4 * 4
        packed = torch.zeros(off, device=dev)
        st = {'packed': packed, 'cam': packed[:ops.CAMERA_WORDS],
              'pose': {k: packed[o:o + n].view(shape) for k, o, n, shape in self._pose_layout},
              'embeds': {k: v.to(dev).clone() for k, v in self.text_embeds_dict.items()},
              'cond': data['cond_images'].to(dev).clone()}
        self._static = st
        self._template = {k: v for k, v in data.items() if k not in ('smpl_inputs', 'cond_images')}
        self._ring = _PinnedRing(off, dev)
        self.text_embeds_dict = st['embeds']
        sdata = dict(self._template, smpl_inputs=st['pose'], cond_images=st['cond'])
        self._send_inputs(data)

        def body():
            loss, ro, so, _ = self._body(sdata, cam_dev=st['cam'])
            self._result_dev = self._result_vector(loss, so)     # captured: the step's scalars in one 12-byte device buffer
            return loss, ro, so
        side = self._stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from ._lib import lib
        graph = torch.cuda.CUDAGraph()
        n0 = lib().launches
        with torch.cuda.graph(graph, stream=side):
            self._out = body()
        self.graph_launches = lib().launches - n0
        torch.cuda.synchronize()
        self._graph = graph
        return self

    def _send_inputs(self, data):
        """Pack this step's camera struct + host-side pose vectors into the next pinned slot and enqueue ONE async H2D
        copy of the packed block (device-resident or absent pose entries are handled per entry)."""
        h = self._ring.next()
        view, proj, campos, tfx, tfy = dcam.raster_matrices(data)
        ops.pack_camera(data['image_height'], data['image_width'], tfx, tfy, view, proj, self.scene.renderer.bg_color, 1.0, out=h[:ops.CAMERA_WORDS])
        pose = data.get('smpl_inputs') or {}
        all_host = True
        for k, o, n, shape in self._pose_layout:
            v = pose.get(k)
            if v is not None and not v.is_cuda:
                h[o:o + n].copy_(v.reshape(-1))
            else:
                all_host = False
        if all_host:
            self._static['packed'].copy_(h, non_blocking=True)
        else:
            self._static['cam'].copy_(h[:ops.CAMERA_WORDS], non_blocking=True)
            for k, o, n, shape in self._pose_layout:
                v = pose.get(k)
                if v is None:
                    continue                                    # absent key: the static buffer keeps its value
                self._static['pose'][k].copy_(v if v.is_cuda else h[o:o + n].view(shape), non_blocking=True)
        self._ring.sent()

    def _replay(self, data):
        from ._lib import lib
        t0 = time.perf_counter()
        st = self._static
        self._send_inputs(data)
        cond = data.get('cond_images')
        if cond is not None and cond.data_ptr() != st['cond'].data_ptr():
            st['cond'].copy_(cond, non_blocking=True)
        t1 = time.perf_counter()
        self._graph.replay()
        t2 = time.perf_counter()
        L = lib()
        object.__setattr__(L, 'launches', L.launches + self.graph_launches)
        self._queue_result()
        self._post()
        self.host_ms = {'inputs': (t1 - t0) * 1e3, 'graph_launch': (t2 - t1) * 1e3, 'post': (time.perf_counter() - t2) * 1e3}
        loss, ro, so = self._out
        return loss, ro, so, None

    def set_text_embeds(self, embeds):
        """Refresh the prompt embeddings (static buffers under the graph)."""
        if self._graph is None:
            self.text_embeds_dict = embeds
        else:
            for k, v in embeds.items():
                self._static['embeds'][k].copy_(v, non_blocking=True)
