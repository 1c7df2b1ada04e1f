"""Reference-equivalent GPU arm (BASELINE.md section 3 "second baseline") -- BENCH / TEST INFRASTRUCTURE ONLY.

What the reference itself would execute per SDS step on ONE GPU, assembled from what can run here:
  * the avatar path (GLBS, lbs_transform, MLPs, mesh-bound hands) = the oracle's eager fp32 torch restatement of the
    reference's Python, moved to the device (the reference's own code is eager torch on cuda as well);
  * the grid encoder = the reference's OWN gridencoder.cu, compiled unmodified for sm_100a (oracle/_ref), driven by the
    autograd wrapper of core/nerf/gridencoder/grid.py:28-94 restated below;
  * the rasteriser = a plain-SIMT restatement of the published 3DGS algorithm (oracle/ref_gpu_raster.cu), a LABELLED
    STAND-IN for the third-party diff_gaussian_rasterization that cannot be installed (no network);
  * VAE encode (with input gradient), ControlNet and UNet = oracle/diffusion.py (diffusers' architectures as eager torch
    modules) in fp32 on cuDNN / cuBLAS / SDPA with torch's default flags (TF32 cuDNN convolutions, fp32 matmuls) -- the
    reference's default precision (configs/__init__.py:236,241) -- no CUDA graphs, no fusion, as the reference runs.
Same workload, seeds and synthetic weights as the dwg arm of bench.py.
"""
import ctypes
import os

import numpy as np
import torch

from . import build_ref

_raster = None


class RefCamera(ctypes.Structure):
    _fields_ = [('H', ctypes.c_int), ('W', ctypes.c_int), ('tanfovx', ctypes.c_float), ('tanfovy', ctypes.c_float),
                ('view', ctypes.c_float * 16), ('proj', ctypes.c_float * 16), ('bg', ctypes.c_float * 3), ('scale_modifier', ctypes.c_float)]


def raster_lib():
    global _raster
    if _raster is None:
        so = build_ref.RASTER_SO
        if not os.path.exists(so):
            so = build_ref.build_raster()
        _raster = ctypes.CDLL(so)
    return _raster


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class SimtRasterize(torch.autograd.Function):
    """rasterize_gaussians / rasterize_gaussians_backward through the SIMT stand-in; scan and sort by torch (CUB)."""

    @staticmethod
    def forward(ctx, means3D, colors, opacities, scales, rotations, cam):
        L = raster_lib()
        dev = means3D.device
        N, H, W = means3D.shape[0], cam.H, cam.W
        gx, gy = (W + 15) // 16, (H + 15) // 16
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        means3D, colors, scales, rotations = means3D.contiguous(), colors.contiguous(), scales.contiguous(), rotations.contiguous()
        opac = opacities.reshape(-1).contiguous()
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        radii = torch.zeros(N, device=dev, dtype=torch.int32)
        xy, depth, cov3D, conic_op = f(N, 2), f(N), f(N, 6), f(N, 4)
        rect = torch.zeros(N, 4, device=dev, dtype=torch.int32)
        tiles = torch.zeros(N, device=dev, dtype=torch.int32)
        L.refr_preprocess(N, _p(means3D), _p(scales), _p(rotations), _p(opac), ctypes.byref(cam), _p(radii), _p(xy), _p(depth), _p(cov3D),
                          _p(conic_op), _p(rect), _p(tiles), st)
        offs = torch.cumsum(tiles.long(), 0)
        P = int(offs[-1].item()) if N > 0 else 0                 # upstream reads num_rendered back to the host as well
        keys = torch.empty(max(P, 1), device=dev, dtype=torch.int64)
        vals = torch.empty(max(P, 1), device=dev, dtype=torch.int32)
        L.refr_duplicate(N, _p(radii), _p(depth), _p(rect), _p(offs), gx, _p(keys), _p(vals), st)
        keys, order = torch.sort(keys[:P])
        vals = vals[:P][order].contiguous()
        ranges = torch.zeros(gx * gy, 2, device=dev, dtype=torch.int32)
        L.refr_ranges(ctypes.c_int64(P), _p(keys), _p(ranges), st)
        color, odepth, oalpha = f(3, H, W), f(1, H, W), f(1, H, W)
        final_T = f(H, W)
        n_contrib = torch.zeros(H, W, device=dev, dtype=torch.int32)
        L.refr_render(ctypes.byref(cam), _p(ranges), _p(vals), _p(xy), _p(conic_op), _p(colors), _p(depth), _p(color), _p(odepth), _p(oalpha),
                      _p(final_T), _p(n_contrib), st)
        ctx.save_for_backward(means3D, colors, scales, rotations, radii, xy, depth, cov3D, conic_op, ranges, vals, final_T, n_contrib)
        ctx.cam, ctx.opac_shape = cam, opacities.shape
        ctx.mark_non_differentiable(radii)
        return color, radii, odepth, oalpha

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_depth, g_alpha):
        means3D, colors, scales, rotations, radii, xy, depth, cov3D, conic_op, ranges, vals, final_T, n_contrib = ctx.saved_tensors
        L = raster_lib()
        cam = ctx.cam
        dev, N = means3D.device, means3D.shape[0]
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        g_m2, g_conic, g_op, g_col, g_dep = f(N, 2), f(N, 3), f(N), f(N, 3), f(N)
        g_color = g_color.contiguous()
        g_depth = None if g_depth is None else g_depth.contiguous()
        g_alpha = None if g_alpha is None else g_alpha.contiguous()
        L.refr_render_bwd(ctypes.byref(cam), _p(ranges), _p(vals), _p(xy), _p(conic_op), _p(colors), _p(depth), _p(final_T), _p(n_contrib),
                          _p(g_color), _p(g_depth), _p(g_alpha), _p(g_m2), _p(g_conic), _p(g_op), _p(g_col), _p(g_dep), st)
        g_m3, g_s, g_r = f(N, 3), f(N, 3), f(N, 4)
        L.refr_preprocess_bwd(N, _p(means3D), _p(scales), _p(rotations), ctypes.byref(cam), _p(radii), _p(cov3D), _p(g_m2), _p(g_conic), _p(g_dep),
                              _p(g_m3), _p(g_s), _p(g_r), st)
        return g_m3, g_col, g_op.reshape(ctx.opac_shape), g_s, g_r, None


def make_camera(H, W, tanfovx, tanfovy, view, proj, bg=(0.0, 0.0, 0.0)):
    cam = RefCamera()
    cam.H, cam.W, cam.tanfovx, cam.tanfovy, cam.scale_modifier = int(H), int(W), float(tanfovx), float(tanfovy), 1.0
    v, p = np.asarray(view, np.float32).reshape(16), np.asarray(proj, np.float32).reshape(16)
    for i in range(16):
        cam.view[i], cam.proj[i] = float(v[i]), float(p[i])
    for i in range(3):
        cam.bg[i] = float(bg[i])
    return cam


class RefGridEncode(torch.autograd.Function):
    """core/nerf/gridencoder/grid.py:28-94 (_grid_encode) over the reference's own compiled kernel."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, S, H, gridtype, align_corners, interp, ge):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype)
        ge.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, gridtype, align_corners, interp, ge)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, align_corners, interp, ge = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        g_emb = torch.zeros_like(embeddings)
        g_in = torch.zeros_like(inputs)
        ge.grid_encode_backward(grad, inputs, embeddings, offsets, g_emb, B, D, C, L, S, H, dy_dx, g_in, gridtype, align_corners, interp)
        return g_in, g_emb, None, None, None, None, None, None, None


class RefGpuScene:
    """One SDS step of the cfg2 workload the way the reference would run it on a GPU (see module docstring)."""

    def __init__(self, device, tiny=False, n_unc=135000, n_tri=2500, img=512, seed_rank=0, poses=None, diffusion=True):
        import sys
        pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dreamwaltz-g_b200')
        if pkg not in sys.path:
            sys.path.insert(0, pkg)
        from dwg import synth                                    # synthetic INPUTS only (shapes of the real assets)
        from dwg.diffusion import weights as W
        from . import grid as ogrid
        self.dev, self.img = device, img
        self.ge = build_ref.load_module()
        to = lambda x: x.to(device) if torch.is_tensor(x) else x
        self.model = {k: to(v) for k, v in synth.make_body_model(0).items()}
        av = synth.make_avatar(synth.make_body_model(0), n_unc, n_tri, seed=0)
        self.av = {k: ({kk: to(vv) for kk, vv in v.items()} if isinstance(v, dict) else to(v)) for k, v in av.items()}
        self.cfg, self.vcfg = (W.TINY, W.TINY_VAE) if tiny else (W.SD15, W.VAE15)
        cu = lambda sd: {k: v.to(device) for k, v in sd.items()}
        if diffusion:
            self.unet, self.cn, self.vae = cu(W.make_unet(self.cfg)), cu(W.make_controlnet(self.cfg)), cu(W.make_vae_encoder(self.vcfg))
        offsets, pls, S, _, _ = ogrid.level_table()
        self.offsets = torch.from_numpy(offsets).to(device)
        self.S = float(np.log2(pls))
        g = torch.Generator().manual_seed(1)
        self.table = (torch.rand(int(offsets[-1]), 2, generator=g) - 0.5).to(device).requires_grad_(True)
        gw = torch.Generator().manual_seed(5)
        rnd = lambda *s: (torch.randn(*s, generator=gw) * 0.3).to(device).requires_grad_(True)
        self.nets = {'sigma_w': [rnd(64, 32), rnd(64, 64), rnd(4, 64)], 'sigma_b': [rnd(64), rnd(64), rnd(4)],
                     'deform': {**{f'layers.{i}.weight': rnd(64, 95 if i == 0 else 64) for i in range(4)}, **{f'layers.{i}.bias': rnd(64) for i in range(4)},
                                'gaussian_warp.weight': rnd(3, 64), 'gaussian_warp.bias': rnd(3), 'gaussian_rotation.weight': rnd(4, 64),
                                'gaussian_rotation.bias': rnd(4), 'gaussian_scaling.weight': rnd(3, 64), 'gaussian_scaling.bias': rnd(3)}}
        self.av['_positions'] = self.av['_positions'].clone().requires_grad_(True)
        self.av['_quaternions'] = self.av['_quaternions'].clone().requires_grad_(True)
        g2 = torch.Generator().manual_seed(7)
        ctx_dim = self.cfg['ctx_dim']
        self.emb = {'neg': torch.randn(1, 77, ctx_dim, generator=g2).to(device), 'text': torch.randn(1, 77, ctx_dim, generator=g2).to(device)}
        self.cond = (torch.rand(1, 3, img, img, generator=g2) > 0.97).float().to(device)
        self.rng = np.random.default_rng(1000 + seed_rank)
        self.rows = poses
        self.i = 0
        self.params = [self.table, self.av['_positions'], self.av['_quaternions']] + self.nets['sigma_w'] + self.nets['sigma_b'] + list(self.nets['deform'].values())

    def step(self):
        from dwg import camera, synth
        from . import avatar as oav, diffusion as od
        dev = self.dev
        for p in self.params:
            p.grad = None
        row = self.rows[self.i % len(self.rows)]
        self.i += 1
        obs = {k: v.to(dev) for k, v in synth.pose_from_row(row).items()}
        data = camera.random_camera(self.rng, self.img, self.img)
        view, proj, campos, tfx, tfy = camera.raster_matrices(data)
        enc = lambda x: RefGridEncode.apply(((x + 2.0) / 4.0), self.table, self.offsets, self.S, 16, 1, False, 1, self.ge)     # grid.py:153
        with torch.device(dev):                                  # the oracle's factory calls (torch.zeros / eye / tensor) land on the GPU
            gs = oav.animate(self.model, self.av, self.nets, enc, {}, obs)
        cam = make_camera(self.img, self.img, tfx, tfy, view.numpy(), proj.numpy())
        color, radii, depth, alpha = SimtRasterize.apply(gs['positions'], gs['colors'], gs['opacities'], gs['scales'], gs['quaternions'], cam)
        img = color.unsqueeze(0)
        h = self.img // 8
        veps = torch.randn(1, 4, h, h, device=dev)
        lat = od.vae_encode_latents(self.vae, self.vcfg, img, veps)
        t = torch.randint(20, 981, (1,), device=dev)
        noise = torch.randn_like(lat)
        with torch.no_grad():
            ln = od.add_noise(lat.detach(), noise, t)
            grad, _ = od.sds_gradient(self.unet, self.cn, self.cfg, ln, noise, t, self.emb['neg'], self.emb['text'], self.cond, 50.0)
        (lat * grad).sum().backward()                            # SpecifyGradient (basic.py:213-226) == this inner product
        return grad


def _render_frame(self):
    """One inference frame the way Trainer.evaluate produces it (trainer.py:1019-1112): no_grad render with a white
    background composite (scene.py:153-156), then per output tensor2image (utils/image.py:52-61): .cpu().numpy(), * 255, clip, uint8."""
    from dwg import camera, synth
    from . import avatar as oav
    dev = self.dev
    row = self.rows[self.i % len(self.rows)]
    self.i += 1
    obs = {k: v.to(dev) for k, v in synth.pose_from_row(row).items()}
    if not hasattr(self, '_cam'):
        data = camera.random_camera(self.rng, self.img, self.img)
        view, proj, campos, tfx, tfy = camera.raster_matrices(data)
        self._cam = make_camera(self.img, self.img, tfx, tfy, view.numpy(), proj.numpy())
    with torch.no_grad():
        enc = lambda x: RefGridEncode.apply(((x + 2.0) / 4.0), self.table, self.offsets, self.S, 16, 1, False, 1, self.ge)
        with torch.device(dev):
            gs = oav.animate(self.model, self.av, self.nets, enc, {}, obs)
        color, radii, depth, alpha = SimtRasterize.apply(gs['positions'], gs['colors'], gs['opacities'], gs['scales'], gs['quaternions'], self._cam)
        image_fg = color.permute(1, 2, 0).unsqueeze(0)
        a = alpha.permute(1, 2, 0).unsqueeze(0)
        image = image_fg + torch.ones_like(image_fg) * (1 - a)
        outs = {'image': image, 'image_fg': torch.cat([image_fg, a], dim=3), 'depth': depth.permute(1, 2, 0).unsqueeze(0) / 3.0, 'alpha': a}
        return {k: (v[0].detach().cpu().numpy() * 255.0).clip(0.0, 255.0).astype(np.uint8) for k, v in outs.items()}


RefGpuScene.render_frame = _render_frame


def time_steps(scene, steps, warmup):
    """CUDA-event timing of `steps` eager steps after `warmup`; returns ms per step."""
    for _ in range(warmup):
        scene.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        scene.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
