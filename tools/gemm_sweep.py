"""Sweep the tile width (BN) and split-K factor of the tcgen05 GEMM on the step's shapes and compare
the planner's automatic choice with the measured optimum.  python tools/gemm_sweep.py [--quick]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402
from dwg._lib import lib  # noqa: E402

GEMMS = [  # M, N, K, act, residual
    (8192, 320, 320, None, False), (8192, 320, 320, None, True), (8192, 2560, 320, 'geglu', False), (8192, 320, 1280, None, True),
    (2048, 640, 640, None, True), (2048, 5120, 640, 'geglu', False), (2048, 640, 2560, None, True),
    (512, 1280, 1280, None, True), (512, 10240, 1280, 'geglu', False), (512, 1280, 5120, None, True),
    (128, 1280, 1280, None, False), (128, 1280, 5120, None, True), (4096, 512, 512, None, False), (640, 1024, 640, None, False),
    (1280, 256, 1280, None, False), (320, 4096, 320, None, False), (4096, 512, 4096, None, False), (4096, 4096, 512, None, False),
]
CONVS = [  # Nimg, H, W, Cin, Cout, k, residual
    (1, 512, 512, 128, 128, 3, False), (1, 256, 256, 256, 256, 3, False), (1, 128, 128, 512, 512, 3, False), (1, 64, 64, 512, 512, 3, True),
    (2, 64, 64, 320, 320, 3, True), (2, 32, 32, 640, 640, 3, True), (2, 16, 16, 1280, 1280, 3, True), (2, 8, 8, 1280, 1280, 3, True),
    (2, 16, 16, 2560, 1280, 3, False), (2, 8, 8, 2560, 1280, 3, False), (2, 32, 32, 1280, 640, 3, False), (2, 64, 64, 640, 320, 3, False),
    (2, 16, 16, 1280, 1280, 1, True), (2, 8, 8, 1280, 1280, 1, True), (2, 64, 64, 320, 320, 1, True), (2, 32, 32, 640, 640, 1, True),
]


def timeit(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def sweep(name, fn, geglu, quick):
    L = lib()
    plan = (ctypes.c_int * 3)()
    L.dwg_gemm_tune(0, 0)
    t_auto = timeit(fn)
    L.dwg_gemm_last_plan(plan)
    auto = tuple(plan)
    res = []
    bns = (64, 128, 192, 256) if geglu else ((32, 64, 128, 160, 256) if quick else (32, 64, 96, 128, 160, 192, 224, 256))
    kss = (1,) if geglu else ((1, 2, 4, 8, 16) if quick else (1, 2, 3, 4, 6, 8, 12, 16, 24, 32))
    for bn in bns:
        for ks in kss:
            L.dwg_gemm_tune(bn, ks)
            fn()
            L.dwg_gemm_last_plan(plan)
            if plan[0] != bn or plan[1] != ks:
                continue
            res.append((timeit(fn), bn, ks))
    L.dwg_gemm_tune(0, 0)
    res.sort()
    best = res[0]
    top = ' '.join(f'{t:.1f}@{bn}/{ks}' for t, bn, ks in res[:4])
    print(f'{name:38s} auto {t_auto:6.1f} us (BN{auto[0]:3d} ks{auto[1]:2d} st{auto[2]})  best {best[0]:6.1f} (BN{best[1]:3d} ks{best[2]:2d})  '
          f'loss {t_auto / best[0]:4.2f}x   top: {top}', flush=True)
    return t_auto, best[0]


def main():
    quick = '--quick' in sys.argv
    dev = 'cuda'
    torch.manual_seed(0)
    tot_a = tot_b = 0.0
    for M, N, K, act, res in GEMMS:
        a = torch.randn(M, K, device=dev).half()
        b = torch.randn(N, K, device=dev).half()
        bias = torch.randn(N, device=dev)
        r = torch.randn(M, N, device=dev).half() if res else None
        ta, tb = sweep(f'gemm M{M} N{N} K{K} {act or "-"} res={int(res)}', lambda: ops.gemm(a, b, bias=bias, act=act, residual=r), act == 'geglu', quick)
        tot_a += ta; tot_b += tb
    for Ni, H, W, Ci, Co, k, res in CONVS:
        x = torch.randn(Ni, H, W, Ci, device=dev).half()
        w = torch.randn(Co, k, k, Ci, device=dev).half()
        bias = torch.randn(Co, device=dev)
        r = torch.randn(Ni, H, W, Co, device=dev).half() if res else None
        ta, tb = sweep(f'conv{k}x{k} {Ni}x{H}x{W} {Ci}->{Co} res={int(res)}', lambda: ops.conv2d_nhwc(x, w, bias=bias, residual=r, padding=k // 2), False, quick)
        tot_a += ta; tot_b += tb
    print(f'sum auto {tot_a:.1f} us, sum best {tot_b:.1f} us')


if __name__ == '__main__':
    main()
