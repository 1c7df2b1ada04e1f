"""Surface 3: the guidance object the trainer calls as ``self.diffusion(**sd_kwargs)``.

Mirrors reference core/guidance/basic.py:778-917 (__call__), :354-383/:420-438 (preprocess /
prepare_latents), :546-663 (calc_gradients), :213-226 (SpecifyGradient) and
core/guidance/controlnet.py:83-114 (_predict) for the shipped configuration: loss 'sds', weight
'sjc' (= 1), classifier-free guidance with the negative prompt, scale 50, uniform timesteps in
[0.02, 0.98] * 1000, ControlNet conditioning scale 1.  No host synchronisation: the timestep
stays on the device.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from . import model as M
from . import weights as Wt


class SpecifyGradient(torch.autograd.Function):
    """basic.py:213-226: forward returns ones[1]; backward injects the SDS gradient."""

    @staticmethod
    def forward(ctx, input_tensor, gt_grad):
        ctx.save_for_backward(gt_grad)
        return torch.ones([1], device=input_tensor.device, dtype=input_tensor.dtype)

    @staticmethod
    def backward(ctx, grad_scale):
        gt_grad, = ctx.saved_tensors
        return gt_grad * grad_scale, None


def alphas_cumprod(device, n=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32, device=device) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class ControlNetScoreDistillation:
    """SD1.5 + ControlNet(openpose) score distillation on dwg kernels."""

    def __init__(self, unet_sd, controlnet_sd, vae_sd, cfg=Wt.SD15, vae_cfg=Wt.VAE15, device='cuda', guidance_scale=50.0,
                 conditioning_scale=1.0, min_timestep=0.02, max_timestep=0.98, seed=0, default_image_size=512, input_interpolate=True,
                 guidance_adjust='constant', time_sampling='uniform', time_annealing='linear'):
        self.device = device
        self.unet = M.UNet(unet_sd, cfg, device)
        self.controlnet = M.ControlNet(controlnet_sd, cfg, device)
        self.vae = M.VAEEncoder(vae_sd, vae_cfg, device)
        self.guidance_scale, self.conditioning_scale = guidance_scale, conditioning_scale
        self.initial_guidance_scale, self.guidance_adjust = guidance_scale, guidance_adjust
        self.acp = alphas_cumprod(device)
        self.t_lo, self.t_hi = int(min_timestep * 1000), int(max_timestep * 1000)
        # guide.time_sampling / time_annealing (configs/__init__.py:261-262, core/guidance/time_prior.py:322-351); the shipped scripts use 'uniform'
        self.time_sampling, self.time_annealing = time_sampling, time_annealing
        self.default_image_size = default_image_size     # 512 (SD1.5) / 768 (SD2.1), basic.py:296-300
        self.input_interpolate = input_interpolate       # basic.py:360-362
        self.vae_scale_factor = 8
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self.use_default_generator = False      # True inside whole-step CUDA-graph capture (graph-safe philox state)
        self.timestep = None
        self.two_streams = True                 # ControlNet beside the UNet encoder (_controlnet_unet)
        self._side = None
        self._helpers = None
        self._prepared = None                   # results of prepare() waiting for the next __call__

    # ---- CUDA graphs: the diffusion blocks have static shapes; one capture each for
    # ControlNet+UNet, VAE forward and VAE backward removes ~1500 launches of host overhead per step
    def enable_graphs(self, image_hw=(512, 512), batch=2, cond_batch=1):
        from .._lib import lib
        dev, L = self.device, self.vae.cfg['latent']
        H, Wd = image_hw
        h, w = H // 8, Wd // 8
        ctx_dim = self.unet.cfg['ctx_dim']
        self._g = {}
        st = {
            'x2': torch.zeros(batch, L, h, w, device=dev), 't': torch.zeros(1, dtype=torch.long, device=dev),
            'ctx': torch.zeros(batch, 77, ctx_dim, device=dev), 'cond': torch.zeros(cond_batch, 3, H, Wd, device=dev),
            'img': torch.zeros(1, 3, H, Wd, device=dev), 'veps': torch.zeros(1, L, h, w, device=dev),
            'glat': torch.zeros(1, L, h, w, device=dev),
        }
        self._static = st

        def predict():
            self.timestep = st['t']
            return self._controlnet_unet(st['x2'], st['ctx'], st['cond'])

        tape = []

        def vae_f():
            tape.clear()
            return self.vae.forward(st['img'], st['veps'], tape)

        def vae_b():
            return self.vae.backward(tape, st['glat'])

        ops.STATS_ARENA.enabled = False              # sub-graphs are replayed independently: each GroupNorm zeroes its own statistics
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):                       # warm-up (allocator, cudaFuncSetAttribute, ...)
                predict(); vae_f(); vae_b()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        counts = {}
        for name, fn in (('predict', predict), ('vae_f', vae_f), ('vae_b', vae_b)):
            g = torch.cuda.CUDAGraph()
            n0 = lib().launches
            with torch.cuda.graph(g), torch.no_grad():
                out = fn()
            counts[name] = lib().launches - n0
            self._g[name] = (g, out)
        self._graph_launches = counts
        ops.STATS_ARENA.enabled = True
        return counts

    def _replay(self, name):
        from .._lib import lib
        g, out = self._g[name]
        g.replay()
        L = lib()
        object.__setattr__(L, 'launches', L.launches + self._graph_launches[name])
        return out

    # ---- inner seams (controlnet.py:83, vae.py:34)
    def encode_images(self, images01, eps=None):
        if eps is None:
            eps = torch.randn(images01.shape[0], self.vae.cfg['latent'], images01.shape[2] // 8, images01.shape[3] // 8,
                              device=images01.device, generator=None if self.use_default_generator else self.gen)
        if getattr(self, '_g', None):
            return _GraphedVaeEncode.apply(images01, eps, self)
        return M.vae_encode(self.vae, images01, eps)

    @torch.no_grad()
    def _predict(self, latents_model_input, text_embeddings, cond_inputs):
        cond = cond_inputs
        if getattr(self, '_g', None):
            st = self._static
            st['x2'].copy_(latents_model_input); st['t'].copy_(self.timestep.reshape(-1)[:1]); st['ctx'].copy_(text_embeddings)
            assert cond.shape[0] == st['cond'].shape[0], 'enable_graphs(cond_batch=...) must match the condition batch'
            st['cond'].copy_(cond)
            return self._replay('predict')
        return self._controlnet_unet(latents_model_input, text_embeddings, cond)

    @torch.no_grad()
    def prepare(self, text_embeds_dict, cond_inputs, batch_size=1, timestep=None, use_negative_text=True):
        """Optional head start (not part of the reference surface): everything of the step that does not depend on
        the rendered image -- the timestep draw, both networks' time-embedding projections and cross-attention
        K / V, the ControlNet condition embedding -- is enqueued on the second stream, to run while the caller
        animates and rasterises the avatar.  The next __call__ consumes it (same timestep semantics)."""
        # Under enable_graphs() the prediction is a replay of a captured graph that does its own (captured) preparation:
        # eager work enqueued here would never be consumed and its lane-1 split-K scratch could race the replay.
        if not self.two_streams or os.environ.get('DWG_NO_PREPARE') == '1' or getattr(self, '_g', None):
            self._prepared = None
            return
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        t = timestep if timestep is not None else self.get_timestep(batch_size)
        neg = text_embeds_dict['neg' if use_negative_text else 'null']
        ctx = torch.cat([neg, text_embeds_dict['text']], dim=0)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            ops.gemm_lane(1)
            try:
                pre_c = self.controlnet.prepare(t, ctx, ctx.shape[0], cond_inputs)
                pre_u = self.unet.prepare(t, ctx, ctx.shape[0])
            finally:
                ops.gemm_lane(0)
        self._prepared = {'t': t, 'ctx': ctx, 'controlnet': pre_c, 'unet': pre_u, 'embeds': (neg, text_embeds_dict['text']), 'cond': cond_inputs}

    def _controlnet_unet(self, x2, ctx, cond):
        """ControlNet and the UNet encoder + mid block are independent until the residuals are added
        (controlnet.py:98-114 / diffusers): the ControlNet is enqueued on a second stream (its own
        split-K scratch lane) and joins before the UNet decoder.  At batch 2 most layers leave SMs idle,
        so the two chains fill each other's gaps; fork/join are graph edges under CUDA-graph capture."""
        B = x2.shape[0]
        if not self.two_streams:
            self.unet.helper = self.controlnet.helper = None
            pre_c = self.controlnet.prepare(self.timestep, ctx, B, cond)
            skips_c, h_c = self.controlnet.features(x2, pre_c, cond)
            state = self.unet.encode(x2, self.timestep, ctx)
            down, mid = self.controlnet.residuals(skips_c, h_c, self.conditioning_scale, add_to=(state[1], state[0]))
            return self.unet.decode(state, down, mid, summed=True)
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        side = self._side
        if self._helpers is None and os.environ.get('DWG_NO_HELPERS') != '1':
            # one helper stream per network (own split-K lanes 2 / 3): ResNet shortcuts and V projections run beside their main chains
            self._helpers = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
        if self._helpers is not None:
            self.unet.helper, self.controlnet.helper = (self._helpers[0], 2), (self._helpers[1], 3)
        prep, self._prepared = self._prepared, None
        pre_c = prep['controlnet'] if prep else None
        pre_u = prep['unet'] if prep else None
        if prep:
            main.wait_stream(side)              # the UNet encoder (this stream) reads the prepared tensors
            for d in (pre_c, pre_u):
                for v in d.values():
                    for tns in (v.values() if isinstance(v, dict) else [v]):
                        for x in (tns if isinstance(tns, tuple) else [tns]):
                            x.record_stream(main)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.gemm_lane(1)
            try:
                if pre_c is None:
                    pre_c = self.controlnet.prepare(self.timestep, ctx, B, cond)
                skips_c, h_c = self.controlnet.features(x2, pre_c, cond)
            finally:
                ops.gemm_lane(0)
        state = self.unet.encode(x2, self.timestep, ctx, pre=pre_u)
        main.wait_stream(side)
        for r in skips_c + [h_c]:
            r.record_stream(main)
        # the zero convolutions run after the join with the UNet skips as their residual operand: skip + residual in one epilogue
        down, mid = self.controlnet.residuals(skips_c, h_c, self.conditioning_scale, add_to=(state[1], state[0]))
        return self.unet.decode(state, down, mid, summed=True)

    def scheduled_timestep(self, train_step, max_iteration):
        """The deterministic timestep modes of TimePrioritizedScheduler.get_timestep (time_prior.py:329-347) as a host integer,
        None for the random ones ('uniform', 'stage').  'annealed' covers the prior-free annealing functions ('linear': p = 1,
        'hifa': p = 0.5, optional ',t_begin,t_end[,p]' arguments; impulse window) -- time_prior.py:199-224."""
        lo, hi = self.t_lo, self.t_hi
        if self.time_sampling == 'constant':
            return (lo + hi) // 2
        if self.time_sampling == 'linear':
            return int(hi - (train_step - 1) * ((hi - lo) / (max_iteration - 1)))
        if self.time_sampling == 'annealed':
            kind, *args = self.time_annealing.split(',')
            if kind not in ('linear', 'hifa'):
                raise NotImplementedError(f'time_annealing {kind!r}: prior-weighted annealing (dreamtime / ddpm / p2) is not part of the hot path')
            p = 1.0 if kind == 'linear' else 0.5
            t_begin, t_end = hi, lo
            if len(args) >= 2:
                t_begin, t_end = int(args[0]), int(args[1])
            if len(args) == 3:
                p = float(args[2])
            assert t_begin >= t_end and hi >= t_begin and lo <= t_end
            return int(t_begin - (t_begin - t_end) * (train_step / max_iteration) ** p)
        return None

    def get_timestep(self, batch_size, train_step=0, max_iteration=1):
        """time_prior.py:322-351.  'uniform' (the shipped setting) and 'stage-K' draw on the device; the other modes are a host
        integer broadcast to the batch."""
        t = self.scheduled_timestep(train_step, max_iteration)
        if t is not None:
            return torch.full((batch_size,), int(t), dtype=torch.long, device=self.device)
        lo, hi = self.t_lo, self.t_hi
        if self.time_sampling.startswith('stage'):
            _, *a = self.time_sampling.split('-')
            n_stage = int(a[0]) if a else 2
            per = (hi - lo) // n_stage
            i_stage = min(train_step // max(max_iteration // n_stage, 1), n_stage - 1)
            hi = lo + per * (n_stage - i_stage)                  # stage_intervals[i_stage][1]; the lower end stays t_lo ("Important!")
        elif self.time_sampling != 'uniform':
            raise NotImplementedError(self.time_sampling)
        return torch.randint(lo, hi + 1, (batch_size,), device=self.device,
                             generator=None if self.use_default_generator else self.gen)

    def add_noise(self, latents, noise, t):
        a = self.acp[t].reshape(-1, 1, 1, 1)
        return a.sqrt() * latents + (1 - a).sqrt() * noise

    def get_guidance_scale(self, train_step, max_iteration):
        """basic.py:404-418."""
        s0 = self.initial_guidance_scale
        if self.guidance_adjust == 'constant':
            return s0
        if self.guidance_adjust == 'uniform':
            return float(np.random.uniform(7.5, s0))
        delta = (s0 - 7.5) / max(max_iteration - 1, 1)
        if self.guidance_adjust == 'linear':
            return s0 - (train_step - 1) * delta
        if self.guidance_adjust == 'linear_reverse':
            return 7.5 + (train_step - 1) * delta
        raise NotImplementedError(self.guidance_adjust)

    def prepare_image(self, image, width=None, height=None):
        """core/guidance/controlnet.py:33-55 (prepare_image): the reference hands the ControlNet condition over as a PIL image (or a
        list of them) and resizes it on the host with LANCZOS every step; tensors / lists of tensors pass through.  Returns a
        [B,3,height,width] fp32 tensor in [0,1] on this object's device.  The device-side producer (dwg.condition) never comes
        through here -- it writes the tensor directly."""
        width = width or self.default_image_size
        height = height or self.default_image_size
        if isinstance(image, torch.Tensor):
            return image.to(device=self.device, dtype=torch.float32)
        if not isinstance(image, (list, tuple)):
            image = [image]
        if isinstance(image[0], torch.Tensor):
            return torch.cat([i if i.dim() == 4 else i.unsqueeze(0) for i in image], dim=0).to(device=self.device, dtype=torch.float32)
        from PIL import Image
        arrs = [np.array(i.convert('RGB').resize((width, height), resample=Image.Resampling.LANCZOS))[None] for i in image]
        x = np.concatenate(arrs, axis=0).astype(np.float32) / 255.0
        return torch.from_numpy(x.transpose(0, 3, 1, 2).copy()).to(self.device)

    def prepare_latents(self, inputs):
        """basic.py:354-383: a 3-channel render that is not default_image_size^2 is resized bilinearly
        (align_corners=False, differentiable) before the VAE; cfg4 renders 1024^2 and feeds SD2.1 at 768^2."""
        size = (self.default_image_size, self.default_image_size)
        if self.input_interpolate and tuple(inputs.shape[-2:]) != size:
            inputs = F.interpolate(inputs, size, mode='bilinear', align_corners=False)
        assert inputs.shape[-2] % 8 == 0 and inputs.shape[-1] % 8 == 0, 'image size must be a multiple of the VAE factor 8'
        return inputs

    def __call__(self, inputs, text_embeds_dict, train_step=0, max_iteration=1, cond_inputs=None, timestep=None, noise=None,
                 vae_eps=None, use_negative_text=True, guidance_scale=None, **_):
        """inputs [1,3,H,W] in [0,1] (autograd-connected); cond_inputs [1,3,S,S] in [0,1] (device tensor).
        Returns the reference's dict: latents, timestep, sources, targets, gradients, diffusion_loss."""
        assert inputs.dim() == 4 and inputs.shape[1] == 3, 'inputs must be [B,3,H,W]'
        if cond_inputs is not None and not isinstance(cond_inputs, torch.Tensor):
            cond_inputs = self.prepare_image(cond_inputs)       # PIL image(s) / list of tensors, as the reference passes them
        inputs = self.prepare_latents(inputs)
        if not getattr(self, '_g', None):
            ops.STATS_ARENA.reset(inputs.device)               # ONE memset serves every GroupNorm statistic of this step
        self.guidance_scale = guidance_scale if guidance_scale is not None else self.get_guidance_scale(train_step, max_iteration)
        latents = self.encode_images(inputs, vae_eps)
        neg = text_embeds_dict['neg' if use_negative_text else 'null']
        prep = self._prepared
        if prep is not None and timestep is None and not getattr(self, '_g', None):
            # prepare() results are only valid for the inputs they were made from
            assert prep['embeds'][0] is neg and prep['embeds'][1] is text_embeds_dict['text'] and (prep['cond'] is None or prep['cond'] is cond_inputs), \
                'prepare() was called with other prompt embeddings / condition than this __call__'
            self.timestep = prep['t']                           # drawn by prepare()
        else:
            if prep is not None and self._side is not None:
                torch.cuda.current_stream().wait_stream(self._side)      # discard: join the side stream so nothing stays in flight
            self._prepared = None
            self.timestep = timestep if timestep is not None else self.get_timestep(inputs.shape[0], train_step, max_iteration)
        with torch.no_grad():
            if noise is None:
                noise = torch.randn(latents.shape, device=latents.device, generator=None if self.use_default_generator else self.gen)
            latents_noisy = self.add_noise(latents.detach(), noise, self.timestep)
            ctx = torch.cat([neg, text_embeds_dict['text']], dim=0)
            x2 = torch.cat([latents_noisy] * 2, dim=0)
            eps = self._predict(x2, ctx, cond_inputs)
            self._prepared = None                              # consumed (or never used): a later call must not see it
            e_u, e_c = eps.chunk(2)
            gradients, noise_pred = ops.sds_grad(e_u.contiguous(), e_c.contiguous(), noise, self.guidance_scale, 1.0)
        loss = SpecifyGradient.apply(latents, gradients)
        return {'latents': latents, 'timestep': self.timestep, 'sources': latents, 'targets': (latents - gradients).detach(),
                'gradients': gradients, 'noise_pred': noise_pred, 'diffusion_loss': loss}


class _GraphedVaeEncode(torch.autograd.Function):
    """vae_encode through the captured forward / backward graphs (static buffers)."""

    @staticmethod
    def forward(ctx, images01, eps, g):
        st = g._static
        st['img'].copy_(images01)
        st['veps'].copy_(eps)
        ctx.g = g
        return g._replay('vae_f').clone()

    @staticmethod
    def backward(ctx, grad):
        g = ctx.g
        g._static['glat'].copy_(grad)
        return g._replay('vae_b').clone(), None, None
