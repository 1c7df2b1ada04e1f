"""Micro-benchmarks of the individual hot-path kernels at the cfg2 sizes (150k Gaussians, 512^2).
CUDA-event timing on the launching stream, warm-up, L2 flushed between timed iterations.
Prints one JSON line per kernel with achieved algorithmic GB/s against MEASURED_PEAKS.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import avatar as dav, camera, ops, synth  # noqa: E402

DEV = 'cuda'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def report(name, ms, best, algo_bytes, pk, src, **extra):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({'kernel': name, 'ms': round(ms, 4), 'ms_best': round(best, 4), 'algo_MB': round(algo_bytes / 1e6, 2),
                      'GBps': round(gbs, 1), 'frac_of_hbm': round(gbs / pk['hbm_gbs'], 3), 'peak': src, **extra}), flush=True)


def gemm_bench(pk, src):
    torch.manual_seed(0)
    for (M, N, K) in ((8192, 320, 320), (8192, 1280, 320), (8192, 4096, 4096), (2048, 1280, 1280)):
        a = torch.randn(M, K, device=DEV).half(); b = torch.randn(N, K, device=DEV).half()
        ms, best = timeit(lambda: ops.gemm(a, b), iters=10)
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        ms2, _ = timeit(lambda: a @ b.t(), iters=10)
        print(json.dumps({'kernel': f'tcgen05_gemm_{M}x{N}x{K}', 'ms': round(ms, 4), 'TFLOPs': round(tf, 1), 'frac_of_bf16_peak': round(tf / pk['bf16_tflops'], 3), 'cublas_ms': round(ms2, 4), 'peak': src}), flush=True)
    for (Nn, H, C, Co) in ((2, 64, 320, 320), (2, 32, 640, 640), (2, 16, 1280, 1280), (1, 256, 128, 128), (1, 512, 128, 128)):
        x = torch.randn(Nn, H, H, C, device=DEV).half(); w = (torch.randn(Co, 3, 3, C, device=DEV) * 0.02).half()
        ms, best = timeit(lambda: ops.conv2d_nhwc(x, w), iters=10)
        tf = 2.0 * Nn * H * H * Co * C * 9 / (ms * 1e-3) / 1e12
        xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last); wc = w.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        ms2, _ = timeit(lambda: torch.nn.functional.conv2d(xc, wc, padding=1), iters=10)
        print(json.dumps({'kernel': f'tcgen05_conv3x3_{Nn}x{H}x{H}x{C}->{Co}', 'ms': round(ms, 4), 'TFLOPs': round(tf, 1), 'frac_of_bf16_peak': round(tf / pk['bf16_tflops'], 3), 'cudnn_bf16_ms': round(ms2, 4), 'peak': src}), flush=True)


def main():
    pk, src = peaks()
    if 'gemm' in sys.argv:
        gemm_bench(pk, src)
        return
    torch.manual_seed(0)
    Nu, J = 135000, 55
    W = torch.rand(Nu, J, device=DEV) ** 6
    W = W / W.sum(1, keepdim=True)
    A = torch.eye(4, device=DEV).repeat(J, 1, 1) + torch.randn(J, 4, 4, device=DEV) * 0.05
    x, q = torch.randn(Nu, 3, device=DEV), torch.randn(Nu, 4, device=DEV)
    L = ops.lib()
    P = ops.ptr
    xo, qo = torch.empty_like(x), torch.empty_like(q)
    st = ops.stream()
    ms, b = timeit(lambda: L.dwg_lbs_skin_fwd(P(W), P(A), P(x), P(q), P(xo), P(qo), Nu, J, st))
    report('lbs_skin_fwd(pos+quat)', ms, b, 276 * Nu, pk, src)
    ms, b = timeit(lambda: L.dwg_lbs_skin_fwd(P(W), P(A), P(x), None, P(xo), None, Nu, J, st))
    report('lbs_skin_fwd(pos)', ms, b, 244 * Nu, pk, src)
    gx, gq = torch.empty_like(x), torch.empty_like(q)
    ms, b = timeit(lambda: L.dwg_lbs_skin_bwd(P(W), P(A), P(x), P(q), P(xo), P(qo), P(gx), P(gq), None, None, Nu, J, st))
    report('lbs_skin_bwd(pos+quat)', ms, b, 304 * Nu, pk, src)
    # torch-eager equivalent of the reference (inverse_lbs.py:208-242) for orientation
    from dwg import lbs as dlbs
    def eager():
        R = torch.einsum('nj,jkl->nkl', W, A[:, :3, :3]); T = torch.einsum('nj,jk->nk', W, A[:, :3, 3])
        xx = torch.matmul(R, x.unsqueeze(-1))[..., 0] + T
        R2 = torch.einsum('nj,jkl->nkl', W, A[:, :3, :3])
        s = torch.tensor([1.0, -1.0, -1.0], device=DEV).view(1, 3, 1)
        return xx, dlbs.matrix_to_quaternion((R2 @ (dlbs.quaternion_to_matrix(q) * s)) * s)
    ms, b = timeit(eager)
    report('torch_eager_lbs_fwd(reference-equivalent)', ms, b, 276 * Nu, pk, src)

    # grid encoder
    spec = ops.GridSpec(DEV, bound=2.0)
    table = torch.empty(spec.n_rows, 2, device=DEV).uniform_(-0.1, 0.1)
    xg = (torch.rand(Nu, 3, device=DEV) - 0.5) * 1.6
    xg[:, 1] *= 1.1
    out = torch.empty(Nu, 32, device=DEV)
    ms, b = timeit(lambda: L.dwg_grid_encode_fwd(P(xg), 2.0, P(table), P(spec.offsets), P(spec.level_scale), P(spec.level_res),
                                                P(out), 32, 2, None, Nu, 16, 1, 0, 1, st))
    report('grid_encode_fwd', ms, b, 1164 * Nu, pk, src)
    gt, gxx = torch.zeros_like(table), torch.empty_like(xg)
    ms, b = timeit(lambda: L.dwg_grid_encode_bwd(P(out), 32, 2, P(xg), 2.0, P(table), P(spec.offsets), P(spec.level_scale),
                                                P(spec.level_res), P(gt), P(gxx), Nu, 16, 1, 0, 1, st))
    report('grid_encode_bwd', ms, b, 1164 * Nu, pk, src)

    # rasteriser on the synthetic avatar (full animate path gives realistic splat statistics)
    model = synth.make_body_model(0)
    av = synth.make_avatar(model, Nu, 2500, seed=0)
    m = dav.DreamWaltzGAvatar(model, av, device=DEV)
    rng = np.random.default_rng(0)
    obs = {k: v.to(DEV) for k, v in synth.random_pose(rng).items()}
    with torch.no_grad():
        gs = m.animate(obs)
    N = gs.positions.shape[0]
    for (H, Wd, rad, fov) in ((512, 512, 2.0, 50.0), (512, 512, 1.2, 45.0), (1024, 1024, 2.0, 50.0)):
        data = camera.make_camera(rad, 30.0, 85.0, fov, H, Wd)
        view, proj, campos, tfx, tfy = camera.raster_matrices(data)
        kw = dict(image_height=H, image_width=Wd, tanfovx=tfx, tanfovy=tfy, viewmatrix=view, projmatrix=proj, bg=torch.zeros(3))
        t = [v.detach().clone().requires_grad_(True) for v in (gs.positions, gs.colors, gs.opacities, gs.scales, gs.quaternions)]
        m2 = torch.zeros(N, 3, device=DEV, requires_grad=True)
        states = []
        color, radii, depth, alpha = ops.rasterize(t[0], m2, t[1], t[2], t[3], t[4], state_out=states, **kw)
        stt = states[0]
        status = stt.status.cpu().numpy()
        Pn = int(status[1])
        vis = int((radii > 0).sum())
        def fwd():
            ops.rasterize(t[0], m2, t[1], t[2], t[3], t[4], **kw)
        with torch.no_grad():
            ms, b = timeit(fwd)
        fb = 56 * N + 24 * Pn + 44 * Pn + 28 * H * Wd
        report(f'raster_fwd_{H}', ms, b, fb, pk, src, P=Pn, visible=vis, max_tile=int(status[2]), N=N)
        gc = torch.randn_like(color)
        def fwdbwd():
            c, _, d, a = ops.rasterize(t[0], m2, t[1], t[2], t[3], t[4], **kw)
            torch.autograd.backward([c], [gc])
        ms2, b2 = timeit(fwdbwd)
        bb = 44 * Pn + 32 * H * Wd + 124 * N
        report(f'raster_fwd+bwd_{H}', ms2, b2, fb + bb, pk, src, bwd_ms_est=round(ms2 - ms, 4))
    # whole animate (fwd) for orientation
    with torch.no_grad():
        ms, b = timeit(lambda: m.animate(obs))
    print(json.dumps({'kernel': 'animate_fwd_total(torch glue + dwg kernels)', 'ms': round(ms, 3)}))


if __name__ == '__main__':
    main()
