"""A short, deterministic sequence of the hot kernels for Nsight Compute captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_targets.py
    ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 4 -o gpurun_out/gemm python tools/ncu_targets.py gemm
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import avatar as dav, camera, ops, synth  # noqa: E402

DEV = 'cuda'
what = sys.argv[1] if len(sys.argv) > 1 else 'all'
torch.manual_seed(0)


def gemm_part():
    for (M, N, K) in ((8192, 320, 320), (8192, 4096, 4096), (512, 1280, 1280)):
        a = torch.randn(M, K, device=DEV).half(); b = torch.randn(N, K, device=DEV).half()
        res = torch.randn(M, N, device=DEV).half(); bias = torch.randn(N, device=DEV)
        for _ in range(2):
            ops.gemm(a, b, bias=bias, residual=res)
    x = torch.randn(2, 64, 64, 320, device=DEV).half(); w = (torch.randn(320, 3, 3, 320, device=DEV) * 0.02).half()
    for _ in range(2):
        ops.conv2d_nhwc(x, w)
    x = torch.randn(1, 512, 512, 128, device=DEV).half(); w = (torch.randn(128, 3, 3, 128, device=DEV) * 0.02).half()
    ops.conv2d_nhwc(x, w)
    q = torch.randn(2, 4096, 320, device=DEV).half(); k = torch.randn(2, 4096, 320, device=DEV).half()
    vt = torch.randn(2, 320, 4096, device=DEV).half()
    for _ in range(2):
        ops.attention(q, k, vt, 8, 4096)
    gx = torch.randn(2, 64, 64, 320, device=DEV).half()
    ops.group_norm(gx, torch.ones(320, device=DEV), torch.zeros(320, device=DEV), 32, 1e-5, True)


def geom_part():
    model = synth.make_body_model(0)
    av = synth.make_avatar(model, 135000, 2500, seed=0)
    m = dav.DreamWaltzGAvatar(model, av, device=DEV)
    with torch.no_grad():
        m.nerf_encoder.embeddings.uniform_(-0.5, 0.5)
    rng = np.random.default_rng(0)
    obs = {k: v.to(DEV) for k, v in synth.random_pose(rng).items()}
    data = camera.make_camera(2.0, 30.0, 85.0, 50.0, 512, 512)
    r = dav.GaussianRenderer()
    for _ in range(2):
        gs = m.animate(obs)
        out = r.render(data, gs)
        out['image'].square().sum().backward()
    torch.cuda.synchronize()


def nn_part():
    x = torch.randn(2, 64, 64, 320, device=DEV).half()
    g, b = torch.ones(320, device=DEV), torch.zeros(320, device=DEV)
    for _ in range(2):
        ops.group_norm(x, g, b, 32, 1e-5, True)
    t = torch.randn(8192, 320, device=DEV).half()
    for _ in range(2):
        ops.layer_norm(t, g, b)
    gg = torch.randn(8192, 2560, device=DEV).half()
    for _ in range(2):
        ops.geglu(gg)
    for _ in range(2):
        ops.add(x, x)
    xv = torch.randn(1, 512, 512, 128, device=DEV).half()
    gv, bv = torch.ones(128, device=DEV), torch.zeros(128, device=DEV)
    y, st = ops.group_norm(xv, gv, bv, 32, 1e-6, True, return_stats=True)
    ops.group_norm_bwd(xv, y, st, gv, bv, 32, 1e-6, True)
    # GroupNorm from the producing conv's epilogue statistics (gn_apply_cs: one node instead of memset + stats + apply)
    wc = (torch.randn(320, 3, 3, 320, device=DEV) * 0.02).half()
    for _ in range(2):
        yc = ops.conv2d_nhwc(x, wc, stats=True)
        ops.group_norm(yc, g, b, 32, 1e-5, True, colstats=yc._cs)
    a = torch.randn(128, 1280, device=DEV).half(); w = torch.randn(1280, 1280, device=DEV).half()
    for _ in range(2):
        ops.gemm(a, w)                      # split-K + finalize
    a2 = torch.randn(2, 320, device=DEV).half(); w2 = torch.randn(1280, 320, device=DEV).half()
    for _ in range(2):
        ops.gemm(a2, w2, act='silu')


if what in ('nn',):
    nn_part()
if what in ('all', 'gemm'):
    gemm_part()
if what in ('all', 'geom'):
    geom_part()
torch.cuda.synchronize()
print('done')
