#!/usr/bin/env python
"""bench.py -- SDS steps/s of the DreamWaltz-G per-step hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one full SDS step on one random view of the cfg2 workload (150k-Gaussian SMPL-X
avatar, 512x512 render, SD1.5 + ControlNet-openpose shapes, synthetic weights / poses / cameras /
prompt embeddings): animate (LBS + grid + MLPs + mesh-bound hands) -> tile rasteriser forward ->
VAE encode -> ControlNet + UNet (CFG batch 2) -> SDS gradient -> backward through the VAE encoder,
the rasteriser and the avatar to every trainable parameter (optimiser excluded).
With N > 1 every rank takes its own view per step and the parameter gradients are summed with ONE
NCCL all-reduce (weak scaling: 1 view per GPU per step); value = views (SDS steps) per second of
the whole job.  `--impl reference` times the CPU oracle (the reference has no CPU path and
cannot be installed here) on the host cores; `--impl reference-gpu` (also embedded
in the default line as `ref_gpu`) times the reference-equivalent GPU step of BASELINE.md section 3.
`--config cfg4|cfg5` run BASELINE.json's other single-GPU configurations (300k / 1024^2 / SD2.1
shapes; 500-frame re-enactment, frames/s).  The e2e arm feeds every step from pinned HOST memory
(camera + pose block, prompt embeddings) and reads every step's 12-byte result back through the
step API's pinned ring, one step late (the last one inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'dreamwaltz-g_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# rank 0 prints exactly ONE JSON line on stdout.  Libraries (NCCL's version banner, ...) write to file descriptor 1
# directly, so main() moves fd 1 onto stderr and keeps the original stdout for the JSON line alone.
_JSON_OUT = sys.stdout


def _claim_stdout():
    global _JSON_OUT
    try:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
    except OSError:
        _JSON_OUT = sys.stdout

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_UNCONSTRAINED, N_MESH_TRI, IMG = 135000, 2500, 512
WORKLOAD = 'cfg2: 150k-Gaussian SMPL-X avatar SDS step @512^2, SD1.5 + ControlNet-openpose shapes (synthetic weights/data)'
# BASELINE.json configs[1] (cfg2, the benchmark), configs[3] (cfg4) and configs[4] (cfg5); cfg3 = cfg2 under torchrun
CONFIGS = {
    'cfg2': dict(n_unc=135000, n_tri=2500, n_face=0, img=512, sd='SD15', sd_size=512, workload=WORKLOAD),
    'cfg4': dict(n_unc=270000, n_tri=2500, n_face=2500, img=1024, sd='SD21', sd_size=768,
                 workload='cfg4: 300k-Gaussian avatar with mesh-bound hands + face and expression, rendered @1024^2, resized to 768^2 '
                          '(basic.py:360-366), SD2.1 + ControlNet shapes (synthetic weights/data)'),
    'cfg5': dict(n_unc=135000, n_tri=2500, n_face=0, img=1024, sd=None, sd_size=0,
                 workload='cfg5: inference_reenact -- 500-frame motion sequence, 150k Gaussians @1024^2, animate + rasterise + uint8 frames '
                          '(image, image_fg+alpha, depth, alpha) delivered to host memory'),
}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


def poses():
    rows = np.load(os.path.join(ROOT, 'tests', 'golden', 'poses.npz'))['rows']
    return rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------- dwg arm
class Workload:
    """Everything one rank needs, built from the PACKAGE's public objects (dwg.step.Scene / SDSTrainStep are the R17 API;
    nothing of the step lives in this file any more): avatar, renderer, guidance, per-step inputs."""

    def __init__(self, device, rank, n_unc=N_UNCONSTRAINED, n_tri=N_MESH_TRI, img=IMG, tiny=False, allreduce=False, n_face=0, sd='SD15',
                 sd_size=None, produce_cond=True):
        from dwg import avatar as dav, step as dstep, synth
        from dwg.diffusion import guidance as G, weights as W
        self.dev, self.rank, self.img = device, rank, img
        sd_size = sd_size or img
        model = synth.make_body_model(0)
        av = synth.make_avatar(model, n_unc, n_tri, seed=0, n_face_triangles=n_face)
        self.avatar = dav.DreamWaltzGAvatar(model, av, device=device)
        self.expr_rng = np.random.default_rng(77 + rank) if n_face else None
        with torch.no_grad():
            g = torch.Generator(device='cpu').manual_seed(1)
            self.avatar.nerf_encoder.embeddings.copy_((torch.rand(self.avatar.nerf_encoder.embeddings.shape, generator=g) - 0.5).to(device))
        self.renderer = dav.GaussianRenderer()
        self.scene = dstep.Scene(self.avatar, self.renderer)
        self.step_i = 0
        self.rng = np.random.default_rng(1000 + rank)          # dwg.parallel.rank_seed(1000, rank)
        self.pose_rows = poses()
        if sd is None:                                          # render-only workload (cfg5)
            return
        cfg, vcfg = (W.TINY, W.TINY_VAE) if tiny else (getattr(W, sd), W.VAE15)
        self.guidance = G.ControlNetScoreDistillation(W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg), cfg, vcfg, device,
                                                      seed=1000 + rank, default_image_size=sd_size)
        self.ctx_dim = cfg['ctx_dim']
        g = torch.Generator().manual_seed(7)
        # host-side (pinned) per-step inputs of the e2e arm
        self.h_embeds = {'neg': torch.randn(1, 77, self.ctx_dim, generator=g).pin_memory(), 'text': torch.randn(1, 77, self.ctx_dim, generator=g).pin_memory()}
        cond = (torch.rand(1, 3, sd_size, sd_size, generator=g) > 0.97).float()  # sparse skeleton-like image (8 x the latent size)
        self.h_cond = cond.pin_memory()
        self.d_embeds = {k: v.to(device) for k, v in self.h_embeds.items()}
        self.d_cond = self.h_cond.to(device)
        # (f1) the ControlNet condition image is produced ON THE DEVICE every step from the posed body's keypoints and the view's own
        # depth / alpha (the reference does this on the CPU per view: Embree + cv2 + PIL); --cond-from-host feeds a fixed image instead
        prod = ks = None
        if produce_cond:
            from dwg import condition
            prod = condition.PoseConditionProducer(sd_size, sd_size, device=device)
            ks = condition.synthetic_keypoint_source(self.avatar.lbs_model, model)
        self.produce_cond = produce_cond
        self.trainer = dstep.SDSTrainStep(self.scene, self.guidance, self.d_embeds, allreduce=allreduce, cond_producer=prod, keypoint_source=ks)
        self.params = self.trainer.params

    def next_view(self, device_inputs=True):
        """The reference's per-step ``data`` dict (camera fields of data/camera/utils.py:301-357 + smpl_inputs + cond_images)."""
        from dwg import camera, synth
        row = self.pose_rows[self.step_i % len(self.pose_rows)]
        self.step_i += 1
        data = camera.random_camera(self.rng, self.img, self.img)
        data['smpl_inputs'] = synth.pose_from_row(row)                               # host tensors; staged through pinned memory
        if self.expr_rng is not None:                                                # random_pose_sampler=...,expr (train_w_expr.sh:10)
            data['smpl_inputs']['expression'] = torch.from_numpy(self.expr_rng.normal(0, 0.5, size=(1, 100)).astype(np.float32))
        if hasattr(self, 'd_cond'):
            data['cond_images'] = self.d_cond if device_inputs else self.h_cond
        return data


def run_reenact(args, dev, rank, world, C):
    """cfg5: frames/s of the re-enactment inference loop (Trainer.evaluate for scripts/inference_reenact.sh) on one GPU."""
    from dwg import _lib, inference
    if rank != 0:
        return
    img = args.image or C['img']
    sc = Workload(dev, rank, n_unc=args.n_unconstrained or C['n_unc'], img=img, sd=None)
    re = inference.Reenactor(sc.scene, bg_mode='white')
    frames = [sc.next_view() for _ in range(args.frames)]
    cam0 = frames[0]
    for f in frames:                                            # a re-enactment sequence keeps ONE camera; only the pose changes
        for k in ('extrinsic', 'c2w', 'projection', 'mvp', 'tanfov', 'fov', 'azimuth', 'elevation', 'radius'):
            f[k] = cam0[k]
    re.capture(frames[0])
    L = _lib.lib()
    sampler = ClockSampler(int(os.environ.get('LOCAL_RANK', 0)))
    got = [0]

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    re.run(frames[:min(20, len(frames))])                        # warm-up
    sampler.start()
    n0 = L.launches

    def device_only():
        for f in frames:
            re._send(f)
            re._graph.replay()
    ms_dev = timed(device_only)
    t0 = time.perf_counter()
    ms_e2e = timed(lambda: re.run(frames, on_frame=lambda i, fr: got.__setitem__(0, got[0] + int(fr['image'][0, 0, 0] >= 0))))
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    per_frame_bytes = sum(v.numel() for v in re._out.values())
    out = {'metric': 'inference_reenact frames/sec (150k Gaussians @1024^2, 500-frame sequence)', 'value': round(len(frames) * 1000.0 / ms_dev, 2),
           'unit': 'frames/s', 'n_gpus': 1, 'steps': len(frames), 'warmup': min(20, len(frames)), 'ms_per_step': round(ms_dev / len(frames), 4),
           'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': C['workload'], 'name': 'cfg5', 'gaussians': int(sc.avatar._positions.shape[0] + sum(m._scales.shape[0] for m in sc.avatar.mesh_binding_gaussians.values())),
                      'image': img, 'cuda_graphs': 'whole frame', 'cache': 'L2 flushed by the workload itself: every frame streams the 48 MB grid table, '
                      '106 MB of skinning weights and 9.4 MB of output frames'},
           'e2e': {'value': round(len(frames) * 1000.0 / ms_e2e, 2), 'unit': 'frames/s', 'ms_per_step': round(ms_e2e / len(frames), 4),
                   'h2d_bytes_per_step': int(re._packed.numel() * 4), 'd2h_bytes_per_step': int(per_frame_bytes), 'frames_delivered': got[0],
                   'wall_s': round(wall, 3)},
           'gpu_launches': int(re.graph_launches), 'clocks': clocks}
    if not args.skip_ref_gpu:
        try:
            out['ref_gpu'] = ref_gpu_reenact(args, dev, img, min(len(frames), 60))
            out['ref_gpu']['speedup_dwg_over_ref_gpu'] = {'e2e': round(out['e2e']['value'] / out['ref_gpu']['value'], 2)}
        except Exception as e:
            out['ref_gpu'] = {'unavailable': f'{type(e).__name__}: {e}'}
    print(json.dumps(out), file=_JSON_OUT, flush=True)


def ref_gpu_reenact(args, dev, img, n_frames):
    """The reference's own inference loop on the GPU: eager animate + SIMT raster stand-in + the torch / numpy post-processing of
    Trainer.evaluate (trainer.py:1068-1084, utils/image.py:52-61: .cpu().numpy() per output, synchronous)."""
    from oracle import ref_gpu
    sc = ref_gpu.RefGpuScene(dev, n_unc=args.n_unconstrained or CONFIGS['cfg5']['n_unc'], img=img, poses=poses(), diffusion=False)
    for _ in range(3):
        sc.render_frame()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_frames):
        sc.render_frame()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {'value': round(n_frames / dt, 2), 'unit': 'frames/s', 'ms_per_step': round(1000.0 * dt / n_frames, 3), 'steps': n_frames,
            'what': 'oracle eager torch avatar path on the GPU + reference gridencoder.cu + plain-SIMT raster stand-in + per-output '
                    '(x * 255).clip().astype(uint8) on the host after .cpu() -- as Trainer.evaluate runs'}


def run_dwg(args):
    import torch.distributed as dist
    from dwg import _lib, ops
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    pk, pk_src = peaks()
    C = CONFIGS[args.config]
    if args.config == 'cfg5':
        return run_reenact(args, dev, rank, world, C)
    sc = Workload(dev, rank, tiny=args.tiny, n_unc=args.n_unconstrained or C['n_unc'], img=args.image or C['img'], allreduce=world > 1,
                  n_face=C['n_face'], sd=C['sd'], sd_size=(args.image or C['sd_size']) if args.config == 'cfg2' else C['sd_size'],
                  produce_cond=not args.cond_from_host)
    args.image = args.image or C['img']
    tr = sc.trainer
    graphed = False
    if not args.no_graphs and not args.no_step_graph:
        try:
            tr.capture(sc.next_view())
            graphed = True
        except Exception as e:                       # fall back to sub-graphs (reported in config)
            print(f'[bench] whole-step graph capture failed ({type(e).__name__}: {e}); using sub-graphs', file=sys.stderr)
            tr._graph = None
            torch.cuda.synchronize()
    if not graphed and not args.no_graphs:
        sc.guidance.use_default_generator = False
        sc.guidance.enable_graphs((sc.guidance.default_image_size,) * 2)
    L = _lib.lib()
    host_parts = {}

    def one_step(e2e):
        # e2e: this step's inputs come from (pinned) HOST memory -- pose, camera struct, prompt embeddings, condition image
        data = sc.next_view(device_inputs=not e2e)
        if e2e:
            if graphed:
                tr.set_text_embeds(sc.h_embeds)
            else:
                tr.text_embeds_dict = {k: v.to(dev, non_blocking=True) for k, v in sc.h_embeds.items()}
                data['cond_images'] = sc.h_cond.to(dev, non_blocking=True)
        if not graphed:
            data['smpl_inputs'] = {k: v.pin_memory().to(dev, non_blocking=True) for k, v in data['smpl_inputs'].items()}
        loss, ro, so, _ = tr.step(data)                 # ONE graph replay (+ ONE NCCL all-reduce of the flat gradient buffer when N > 1)
        for k, v in tr.host_ms.items():
            host_parts[k] = host_parts.get(k, 0.0) + v
        if e2e and tr.train_step > 1:
            # device -> host read of a step's result (12 bytes: loss, mean |SDS gradient|, timestep) through the step API's pinned
            # ring, ONE STEP LATE so that the host preparation of the next step's inputs overlaps the device; the last step's
            # result is read inside the timed region as well (timed())
            return tr.fetch_result(lag=1)
        return None

    cpu_ms = [0.0]

    def timed(steps, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.launches
        e0.record()
        t_cpu = time.perf_counter()
        for _ in range(steps):
            one_step(e2e)
        cpu_ms[0] = (time.perf_counter() - t_cpu) * 1000.0 / steps        # host enqueue time (no sync inside)
        if e2e:
            tr.fetch_result(lag=0)                                        # the last step's result, still inside the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, (L.launches - n0) / steps

    for _ in range(args.warmup):
        one_step(False)
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as pr:
            for _ in range(3):
                one_step(False)
            torch.cuda.synchronize()
        print(pr.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90), file=sys.stderr)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(args.steps, False)
    cpu_enqueue_ms = cpu_ms[0]
    for _ in range(min(2, args.warmup)):
        one_step(True)
    ms_e2e, _ = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- roofline of the dominant kernel (tcgen05 GEMM / implicit-GEMM conv): one instrumented,
    # un-graphed, single-stream guidance pass with CUDA events around every tensor-core launch; the GPU is kept
    # busy ahead of the CPU (torch.cuda._sleep) so the events bracket pure execution time.
    roof = None
    if rank == 0:
        g = sc.guidance
        saved = getattr(g, '_g', None)
        g._g = None
        saved_ts, g.two_streams = g.two_streams, False      # one stream: each launch is timed alone (no ControlNet || UNet overlap)
        g._prepared = None
        ops.PROFILE = []
        ops.PROFILE_BYTES = 0.0
        img = torch.rand(1, 3, sc.guidance.default_image_size, sc.guidance.default_image_size, device=dev, requires_grad=True)
        for _ in range(15):
            torch.cuda._sleep(int(4e8))          # ~3 s head start: the CPU enqueues the whole un-graphed pass while the GPU is parked
        res = g(img, sc.d_embeds, cond_inputs=sc.d_cond)
        res['diffusion_loss'].backward()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        g._g = saved
        g.two_streams = saved_ts
        tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
        tot_fl = sum(f for _, _, f, _ in prof)
        ach = tot_fl / (tot_ms * 1e-3) / 1e12
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        roof = {'bound': 'tensor', 'kernel': 'dwg::gemm::gemm_kernel (tcgen05 GEMM + implicit-GEMM conv)', 'achieved': round(ach, 1),
                'peak': peak, 'unit': 'TFLOP/s', 'frac': round(ach / peak, 4), 'traffic': None, 'launches': len(prof),
                'algorithmic_tflop_per_step': round(tot_fl / 1e12, 3), 'kernel_ms_per_step': round(tot_ms, 3),
                'peak_source': pk_src + ' bf16_tflops_sustained (kernel timed inside a long step)'}

    # ---- rasteriser roofline (SURVEY 8d formula): fwd + bwd of the last view, un-graphed, CUDA events on the stream
    roof_r = None
    if rank == 0:
        try:
            roof_r = raster_roofline(sc, args.image, pk, pk_src)
        except Exception as e:                       # reported, never fatal for the headline number
            roof_r = {'error': f'{type(e).__name__}: {e}'}
        if roof is not None:
            tj = os.path.join(ROOT, 'profiles', 'r2c_gemm_traffic.json')
            if os.path.exists(tj):
                t = json.load(open(tj))
                roof['traffic'] = round(t['gemm']['dram_bytes_per_launch'] / 1e6, 3)
                roof['traffic_unit'] = 'MB of DRAM read+write per launch, mean over the %d gemm_kernel launches of one step (ncu, %s)' % (
                    t['gemm']['launches'], 'profiles/r2c_gemm_traffic.json')
                n_gemm = sum(1 for _, _, _, kind in prof if not kind.startswith('attention'))
                roof['algorithmic_MB_per_launch'] = round(ops.PROFILE_BYTES / 1e6 / max(n_gemm, 1), 3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * 1000.0 / ms_step
    e2e_v = world * 1000.0 / ms_e2e
    h2d = (0 if sc.produce_cond else sc.h_cond.numel() * 4) + sum(v.numel() * 4 for v in sc.h_embeds.values()) + (265 + 40) * 4
    out = {
        'metric': 'SDS steps/sec (150k Gaussians, 512^2, SD1.5)' + ('' if world == 1 else ' -- single-view SDS steps (views) per second of the whole job'),
        'value': round(value, 3), 'unit': 'steps/s' if world == 1 else 'views/s', 'n_gpus': world,
        'optimizer_steps_per_s': round(1000.0 / ms_step, 3),
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms_step, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f16 tensor-core GEMMs (fp32 accumulate in TMEM), fp16 activations; fp32 geometry/raster', 'data': 'synthetic',
        'config': {'workload': C['workload'] if not args.tiny else 'TINY smoke configuration (not the benchmark workload)', 'name': args.config,
                   'gaussians': int(sc.avatar._positions.shape[0] + sum(m._scales.shape[0] for m in sc.avatar.mesh_binding_gaussians.values())),
                   'image': args.image, 'views_per_step': world, 'parallelism': f'view-dp{world} + 1 NCCL all-reduce' if world > 1 else 'single GPU',
                   'cache': 'inputs larger than L2 (2.6 GB of fp16 weights streamed every step; 126 MB L2)',
                   'cuda_graphs': ('whole step' if graphed else ('sub-graphs' if not args.no_graphs else False)),
                   'condition_image': 'produced on the device every step (keypoints -> projection -> depth-tested -> OpenPose image)' if sc.produce_cond
                   else 'fixed image copied from pinned host memory',
                   'e2e_result_read': 'every step, asynchronously through a pinned ring, one step late (the last one inside the timed region)'},
        'e2e': {'value': round(e2e_v, 3), 'unit': 'steps/s' if world == 1 else 'views/s', 'ms_per_step': round(ms_e2e, 3), 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': 12},
        'gpu_launches': int(round(launches)), 'host_enqueue_ms_per_step': round(cpu_enqueue_ms, 3),
        'host_ms_parts_per_step': {k: round(v / max(1, 2 * args.steps + args.warmup + min(2, args.warmup)), 3) for k, v in host_parts.items()}, 'clocks': clocks, 'roofline': roof,
        'roofline_raster': roof_r,
    }
    if world == 1 and not args.skip_ref_gpu and args.config == 'cfg2':
        # the north-star target is stated against the reference's single-GPU step: time the reference-equivalent GPU arm
        # on the same box right after the dwg arm (rank 0, N = 1 only; bounded to a few steps)
        try:
            out['ref_gpu'] = ref_gpu_block(args)
            out['ref_gpu']['speedup_dwg_over_ref_gpu'] = {'device_timed': round(value / out['ref_gpu']['value'], 2),
                                                          'e2e': round(e2e_v / out['ref_gpu']['value'], 2), 'target': 10.0}
        except Exception as e:
            out['ref_gpu'] = {'unavailable': f'{type(e).__name__}: {e}'}
    if not args.skip_cpu_baseline and world == 1:      # rank 0, N = 1 only (bounded sample)
        out['cpu_baseline'] = cpu_baseline(sample_only=True)
    print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def raster_roofline(sc, img, pk, pk_src):
    """Achieved GB/s of dwg_raster_forward + dwg_raster_backward against the HBM copy peak.
    Algorithmic bytes: fwd 56 N + 24 P + 44 P + 28 HW, bwd 44 P + 32 HW + 124 N, P measured on this view."""
    from dwg import camera, ops
    dev = sc.dev
    data = sc.next_view()
    pose = data['smpl_inputs']
    with torch.no_grad():
        gs = sc.avatar.animate({k: v.to(dev) for k, v in pose.items()})
    view, proj, campos, tfx, tfy = camera.raster_matrices(data)
    kw = dict(image_height=img, image_width=img, tanfovx=tfx, tanfovy=tfy, viewmatrix=view, projmatrix=proj, bg=torch.zeros(3))
    t = [v.detach().clone().requires_grad_(True) for v in (gs.positions, gs.colors, gs.opacities, gs.scales, gs.quaternions)]
    N = t[0].shape[0]
    m2 = torch.zeros(N, 3, device=dev, requires_grad=True)
    states = []
    color, radii, depth, alpha = ops.rasterize(t[0], m2, t[1], t[2], t[3], t[4], state_out=states, **kw)
    P = int(states[0].status.cpu().numpy()[1])
    gc = torch.randn_like(color)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, n=5):
        best = []
        for _ in range(n):
            flush.zero_()                                  # L2 flush between timed iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            best.append(a.elapsed_time(b))
        return float(np.median(best))

    def fwd():
        with torch.no_grad():
            ops.rasterize(t[0], m2, t[1], t[2], t[3], t[4], **kw)

    def fwdbwd():
        c, _, _, _ = ops.rasterize(t[0], m2, t[1], t[2], t[3], t[4], **kw)
        torch.autograd.backward([c], [gc])
    for _ in range(3):
        fwdbwd()
    ms_f, ms_fb = timed(fwd), timed(fwdbwd)
    HW = img * img
    bytes_f = 56 * N + 68 * P + 28 * HW
    bytes_b = 44 * P + 32 * HW + 124 * N
    ach = (bytes_f + bytes_b) / (ms_fb * 1e-3) / 1e9
    return {'bound': 'hbm', 'kernel': 'dwg_raster_forward + dwg_raster_backward (preprocess, bin, sort, pack, render fwd/bwd, preprocess bwd)',
            'achieved': round(ach, 1), 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': round(ach / pk['hbm_gbs'], 4), 'traffic': None,
            'algorithmic_MB': round((bytes_f + bytes_b) / 1e6, 2), 'ms_fwd': round(ms_f, 4), 'ms_fwd_bwd': round(ms_fb, 4),
            'instances_P': P, 'gaussians': int(N), 'peak_source': pk_src + ' hbm_gbs',
            'note': 'latency / load-balance bound: ~60 MB of compulsory traffic is ~10 us of HBM time; the stage is 14 launches and the '
                    'heaviest 16x16 tiles hold thousands of depth-ordered instances (DESIGN.md section 4)'}


# ----------------------------------------------------------------------------------------- CPU arm
class OracleScene:
    """The same workload on the CPU oracle (torch CPU fp32 + the C raster / grid oracle)."""

    def __init__(self, tiny=False, n_unc=N_UNCONSTRAINED, img=IMG):
        from dwg import synth
        from dwg.diffusion import weights as W
        from oracle import grid as ogrid
        torch.set_num_threads(min(os.cpu_count(), 32))       # more threads than this slows the CPU oracle down
        self.img = img
        self.model = synth.make_body_model(0)
        self.av = synth.make_avatar(self.model, n_unc, N_MESH_TRI, seed=0)
        self.cfg, self.vcfg = (W.TINY, W.TINY_VAE) if tiny else (W.SD15, W.VAE15)
        self.unet, self.cn, self.vae = W.make_unet(self.cfg), W.make_controlnet(self.cfg), W.make_vae_encoder(self.vcfg)
        self.offsets, _, _, self.scale, self.res = ogrid.level_table()
        g = torch.Generator().manual_seed(1)
        self.table = (torch.rand(int(self.offsets[-1]), 2, generator=g) - 0.5)
        gw = torch.Generator().manual_seed(5)
        rnd = lambda *s: torch.randn(*s, generator=gw) * 0.3
        self.nets = {'sigma_w': [rnd(64, 32), rnd(64, 64), rnd(4, 64)], 'sigma_b': [rnd(64), rnd(64), rnd(4)],
                     'deform': {**{f'layers.{i}.weight': rnd(64, 95 if i == 0 else 64) for i in range(4)}, **{f'layers.{i}.bias': rnd(64) for i in range(4)},
                                'gaussian_warp.weight': rnd(3, 64), 'gaussian_warp.bias': rnd(3), 'gaussian_rotation.weight': rnd(4, 64),
                                'gaussian_rotation.bias': rnd(4), 'gaussian_scaling.weight': rnd(3, 64), 'gaussian_scaling.bias': rnd(3)}}
        g2 = torch.Generator().manual_seed(7)
        self.emb = {'neg': torch.randn(1, 77, self.cfg['ctx_dim'], generator=g2), 'text': torch.randn(1, 77, self.cfg['ctx_dim'], generator=g2)}
        self.cond = (torch.rand(1, 3, img, img, generator=g2) > 0.97).float()
        self.rng = np.random.default_rng(1000)
        self.rows = poses()
        self.i = 0

    def step(self):
        from dwg import camera, synth
        from oracle import avatar as oav, diffusion as od, grid as ogrid, raster as orast
        row = self.rows[self.i % len(self.rows)]
        self.i += 1
        obs = synth.pose_from_row(row)
        data = camera.random_camera(self.rng, self.img, self.img)
        view, proj, campos, tfx, tfy = camera.raster_matrices(data)
        pos = self.av['_positions'].clone().requires_grad_(True)
        av = dict(self.av)
        av['_positions'] = pos
        av['_quaternions'] = self.av['_quaternions'].clone().requires_grad_(True)
        table = self.table

        class Enc(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x):
                out, dy, _ = ogrid.forward(x.detach().numpy(), table.numpy(), self.offsets, self.scale, self.res, bound=2.0)
                ctx.save_for_backward(x)
                ctx.dy = dy
                return torch.from_numpy(out)

            @staticmethod
            def backward(ctx, g):
                x, = ctx.saved_tensors
                _, gx = ogrid.backward(g.numpy(), x.detach().numpy(), tuple(table.shape), self.offsets, self.scale, self.res, dy_dx=ctx.dy, bound=2.0)
                return torch.from_numpy(gx)
        gs = oav.animate(self.model, av, self.nets, Enc.apply, {}, obs)
        cam = orast.make_camera(self.img, self.img, tfx, tfy, view.numpy(), proj.numpy(), (0, 0, 0))
        o = orast.forward(cam, gs['positions'].detach().numpy(), gs['scales'].detach().numpy(), gs['quaternions'].detach().numpy(),
                          gs['opacities'].detach().numpy(), gs['colors'].detach().numpy())
        img = torch.from_numpy(o['color']).unsqueeze(0).requires_grad_(True)
        veps = torch.randn(1, 4, self.img // 8, self.img // 8)
        lat = od.vae_encode_latents(self.vae, self.vcfg, img, veps)
        t = torch.tensor([int(self.rng.integers(20, 981))])
        noise = torch.randn_like(lat)
        with torch.no_grad():
            ln = od.add_noise(lat.detach(), noise, t)
            grad, _ = od.sds_gradient(self.unet, self.cn, self.cfg, ln, noise, t, self.emb['neg'], self.emb['text'], self.cond, 50.0)
        (lat * grad).sum().backward()
        b = orast.backward(cam, o, img.grad[0].numpy())
        torch.autograd.backward([gs['positions'], gs['colors'], gs['opacities'], gs['scales'], gs['quaternions']],
                                [torch.from_numpy(b['means3D']), torch.from_numpy(b['colors']), torch.from_numpy(b['opacities']),
                                 torch.from_numpy(b['scales']), torch.from_numpy(b['rots'])])
        return float(grad.abs().mean())


def cpu_baseline(sample_only=True, tiny=False):
    """The oracle timed on the host cores on a bounded sample: ONE full SDS step of the same workload."""
    t0 = time.time()
    sc = OracleScene(tiny=tiny)
    t1 = time.time()
    sc.step()
    dt = time.time() - t1
    return {'value': round(1.0 / dt, 5), 'unit': 'steps/s', 'cores': min(os.cpu_count(), 32), 'kind': 'port',
            'sample': f'1 full SDS step of the same workload on the CPU oracle ({dt:.1f} s; setup {t1 - t0:.0f} s not counted)'}


def ref_gpu_block(args, steps=4, warmup=2):
    """Reference-equivalent GPU arm (oracle/ref_gpu.py): eager fp32 torch modules on cuDNN / cuBLAS / SDPA with torch's default
    TF32 flags + the reference's own gridencoder.cu + a plain-SIMT raster stand-in, same workload, CUDA-event timed."""
    from oracle import ref_gpu
    dev = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    sc = ref_gpu.RefGpuScene(dev, tiny=args.tiny, n_unc=args.n_unconstrained or CONFIGS['cfg2']['n_unc'], img=args.image or CONFIGS['cfg2']['img'], poses=poses())
    ms = ref_gpu.time_steps(sc, steps, warmup)
    del sc
    torch.cuda.empty_cache()
    return {'value': round(1000.0 / ms, 4), 'unit': 'steps/s', 'ms_per_step': round(ms, 2), 'steps': steps, 'warmup': warmup,
            'dtype': 'f32 (torch defaults: TF32 cuDNN convolutions, fp32 matmuls), eager, no CUDA graphs',
            'what': 'oracle eager torch modules on the GPU (avatar path, VAE, ControlNet, UNet: cuDNN / cuBLAS / SDPA) + the reference\'s own '
                    'gridencoder.cu (unmodified, sm_100a) + plain-SIMT 3DGS raster stand-in (oracle/ref_gpu_raster.cu) -- BASELINE.md section 3'}


def run_reference_gpu(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    sampler = ClockSampler(int(os.environ.get('LOCAL_RANK', 0)))
    sampler.start()
    blk = ref_gpu_block(args, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 3)))
    clocks = sampler.stop()
    line = {'impl': 'reference-gpu', 'metric': 'SDS steps/sec (150k Gaussians, 512^2, SD1.5)', 'value': blk['value'], 'unit': 'steps/s',
            'n_gpus': 1, 'steps': blk['steps'], 'warmup': blk['warmup'], 'ms_per_step': blk['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': blk['dtype'], 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'note': blk['what']},
            'clocks': clocks, 'e2e': {'value': blk['value'], 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 8}, 'gpu_launches': 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    budget_s = 240.0
    sc = OracleScene(tiny=args.tiny)
    times = []
    t_begin = time.time()
    n_warm = 0
    for _ in range(args.warmup):
        sc.step()
        n_warm += 1
        if time.time() - t_begin > 60.0:
            break
    t_begin = time.time()
    for _ in range(args.steps):
        t0 = time.time()
        sc.step()
        times.append(time.time() - t0)
        if time.time() - t_begin + times[-1] > budget_s:
            break
    ms = 1000.0 * float(np.mean(times))
    v = 1000.0 / ms
    world = int(os.environ.get('WORLD_SIZE', 1))
    unit = 'steps/s' if world == 1 else 'views/s'      # one CPU step = one view; rank 0 alone runs this arm
    line = json.dumps({
        'impl': 'reference', 'metric': 'SDS steps/sec (150k Gaussians, 512^2, SD1.5)', 'value': round(v, 5), 'unit': unit,
        'n_gpus': world, 'steps': len(times), 'warmup': n_warm, 'ms_per_step': round(ms, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': 'reference has no CPU path and is not installable here (diffusers, smplx, pytorch3d, '
                   'diff_gaussian_rasterization absent, no network): this arm times the CPU oracle restatement of the same step'},
        'cpu_baseline': {'value': round(v, 5), 'unit': unit, 'cores': min(os.cpu_count(), 32), 'kind': 'port',
                         'sample': f'{len(times)} full SDS steps of the same workload'},
        'e2e': {'value': round(v, 5), 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0,
    })
    print(line, file=_JSON_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='dwg', choices=['dwg', 'reference', 'reference-gpu'])
    ap.add_argument('--tiny', action='store_true', help='reduced-width smoke configuration (NOT the benchmark workload)')
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--no-step-graph', action='store_true', help='capture only the diffusion sub-graphs')
    ap.add_argument('--profile', action='store_true', help='print a CUPTI kernel table of 3 steps to stderr')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--cond-from-host', action='store_true', help='feed a fixed condition image from host memory instead of producing it on the device')
    ap.add_argument('--skip-ref-gpu', action='store_true', help='do not time the reference-equivalent GPU arm after the dwg arm')
    ap.add_argument('--config', default='cfg2', choices=sorted(CONFIGS), help='BASELINE.json configuration (cfg2 = the benchmark)')
    ap.add_argument('--n-unconstrained', type=int, default=0, help='override the number of unconstrained Gaussians')
    ap.add_argument('--image', type=int, default=0, help='override the render size')
    ap.add_argument('--frames', type=int, default=500, help='cfg5: frames of the motion sequence')
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == 'reference':
        run_reference(args)
    elif args.impl == 'reference-gpu':
        run_reference_gpu(args)
    else:
        run_dwg(args)


if __name__ == '__main__':
    main()
