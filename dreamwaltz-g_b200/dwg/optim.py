"""(f2) The optimiser step of the 3DGS SDS stage as ONE fused kernel (dwg_adam_step).

Mirrors what the reference builds in DreamWaltzG.get_optimizer (core/system/avatar.py:1590-1635) and steps in
core/trainer.py:863-890:
  'avatar'  GaussianOptimizer (core/gaussian/gaussian_optimizer.py:49-141): Adam(lr=0, eps=1e-15) with per-name rates --
            positions on get_expon_lr_func (core/optim/optim_utils.py:5-40; lr_delay_mult is inert because
            lr_delay_steps stays 0) times the spatial scale, scales = scaling_lr * spatial scale, quaternions = rotation_lr;
  'nerf'    Adam(betas=(0.9, 0.99), eps=1e-15): grid encoder at 10 x lr, the two MLPs at lr;
  'mesh_*'  Adam(lr=0, eps=1e-15): bary coords at position_lr_init, mesh scales at scaling_lr (avatar.py:1081-1094).
All of them become hyper-parameter GROUPS of one flat update over the GradBucket's buffer: parameters are re-homed as
views of one flat fp32 buffer (same layout as the gradients), the moments are two more flat buffers.
"""
import ctypes

import numpy as np
import torch

from ._lib import check, lib, ptr, stream

# configs/__init__.py:75,152-157 (RenderConfig / NeRFConfig defaults)
DEFAULTS = dict(position_lr_init=0.00016, position_lr_final=0.0000016, feature_lr=0.0125, opacity_lr=0.01, scaling_lr=0.0025,
                rotation_lr=0.001, nerf_lr=1e-3)


def expon_lr(step, lr_init, lr_final, max_steps):
    """get_expon_lr_func(...)(step) with lr_delay_steps = 0 (core/optim/optim_utils.py:23-37)."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    t = np.clip(step / max_steps, 0, 1)
    return float(np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t))


class FusedAdam:
    """groups: list of dicts {'name', 'params': [...], 'lr', 'betas': (b1, b2), 'eps'}; every trainable parameter of the
    bucket must be in exactly one group."""

    def __init__(self, bucket, groups):
        self.bucket, self.groups = bucket, groups
        dev = bucket.flat.device
        gid = {}
        for k, g in enumerate(groups):
            for p in g['params']:
                assert id(p) not in gid, 'a parameter appears in two groups'
                gid[id(p)] = k
        assert all(id(p) in gid for p in bucket.params), 'every bucket parameter needs an optimiser group'
        n = bucket.flat.numel()
        ends = [bucket.offsets[i + 1] if i + 1 < len(bucket.params) else n for i in range(len(bucket.params))]
        self._seg_end = (ctypes.c_int64 * len(ends))(*ends)
        self._seg_group = (ctypes.c_int32 * len(ends))(*[gid[id(p)] for p in bucket.params])
        G = len(groups)
        fa = lambda key, i=None: (ctypes.c_float * G)(*[float(g[key] if i is None else g[key][i]) for g in groups])
        for g in groups:
            g.setdefault('betas', (0.9, 0.999)); g.setdefault('eps', 1e-8)
        self._b1, self._b2, self._eps = fa('betas', 0), fa('betas', 1), fa('eps')
        # parameters re-homed into one flat buffer (views), moments alongside
        self.flat_params = torch.zeros(n, device=dev, dtype=torch.float32)
        for p, o in zip(bucket.params, bucket.offsets):
            self.flat_params[o:o + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.flat_params[o:o + p.numel()].view_as(p)
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self.lr_dev = torch.tensor([g['lr'] for g in groups], device=dev, dtype=torch.float32)
        self._lr_host = [torch.zeros(G).pin_memory() for _ in range(8)]
        self._lr_i = 0
        self.current_iteration = 0

    @property
    def param_groups(self):
        return self.groups

    def set_lrs(self):
        """Push the groups' current 'lr' values to the device table (async, pinned ring)."""
        self._lr_i = (self._lr_i + 1) % len(self._lr_host)
        h = self._lr_host[self._lr_i]
        for k, g in enumerate(self.groups):
            h[k] = float(g['lr'])
        self.lr_dev.copy_(h, non_blocking=True)

    def zero_grad(self, set_to_none=False):
        self.bucket.zero()

    def step(self):
        b = self.bucket
        check(lib().dwg_adam_step(ptr(self.flat_params), ptr(b.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq), b.flat.numel(),
                                  len(b.params), self._seg_end, self._seg_group, len(self.groups), self._b1, self._b2, self._eps,
                                  ptr(self.lr_dev), ptr(self.step_dev), stream()), 'dwg_adam_step')
        self.current_iteration += 1

    def state_dict(self):
        return {'step': self.step_dev.clone(), 'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone(),
                'lr': [g['lr'] for g in self.groups], 'current_iteration': self.current_iteration}

    def load_state_dict(self, sd):
        self.step_dev.copy_(sd['step']); self.exp_avg.copy_(sd['exp_avg']); self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        for g, lr in zip(self.groups, sd['lr']):
            g['lr'] = lr
        self.current_iteration = sd.get('current_iteration', 0)
        self.set_lrs()


class AvatarOptimizer(FusedAdam):
    """All optimisers of DreamWaltzG.get_optimizer as groups of one fused update, with GaussianOptimizer's
    update_learning_rate (gaussian_optimizer.py:128-139)."""

    def __init__(self, avatar, bucket, iterations, **cfg):
        c = dict(DEFAULTS, **cfg)
        self.cfg, self.iterations = c, iterations
        named = dict(avatar.named_parameters())
        groups = []

        def add(name, names, lr, betas=(0.9, 0.999), eps=1e-15):
            ps = [named[n] for n in names if n in named and named[n].requires_grad]
            if ps:
                groups.append({'name': name, 'params': ps, 'lr': lr, 'betas': betas, 'eps': eps})
        add('positions', ['_positions'], c['position_lr_init'])
        add('scales', ['_scales'], c['scaling_lr'])
        add('quaternions', ['_quaternions'], c['rotation_lr'])
        add('nerf_encoder', [n for n in named if n.startswith('nerf_encoder.')], c['nerf_lr'] * 10, betas=(0.9, 0.99))
        add('nerf_mlps', [n for n in named if n.startswith('nerf_opacity_and_color_net.') or n.startswith('nerf_scale_and_quaternion_net.')],
            c['nerf_lr'], betas=(0.9, 0.99))
        add('mesh_bary_coords', [n for n in named if n.endswith('._bary_coords')], c['position_lr_init'])
        add('mesh_scales', [n for n in named if n.startswith('mesh_binding_gaussians.') and n.endswith('._scales')], c['scaling_lr'])
        super().__init__(bucket, groups)

    def update_learning_rate(self, spatial_scale=1.0, iteration=None):
        it = self.current_iteration if iteration is None else iteration
        lr = 0.0
        for g in self.groups:
            if g['name'] == 'positions':
                lr = expon_lr(it, self.cfg['position_lr_init'], self.cfg['position_lr_final'], self.iterations * 2)
                g['lr'] = lr * spatial_scale
            elif g['name'] == 'scales':
                lr = self.cfg['scaling_lr']
                g['lr'] = lr * spatial_scale
        self.set_lrs()
        return lr
