"""Timeline of one GEMM launch (CTA 0, globaltimer): where the fixed cost of a small launch goes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402
from dwg._lib import lib  # noqa: E402

L = lib()
dev = 'cuda'
tr = torch.zeros(16, dtype=torch.int64, device=dev)
names = ['entry', 'setup', 'pdl_wait', 'first_stage', 'last_mma', 'acc_visible', 'last_store', 'drained', 'c0_ld', 'c0_math', 'c0_sts', 'c0_fence', 'c0_store']
for M, N, K, res in ((128, 32, 64, False), (8192, 320, 320, False), (8192, 320, 320, True), (512, 1280, 1280, True), (128, 1280, 1280, False), (2048, 640, 640, True)):
    a = torch.randn(M, K, device=dev).half()
    b = torch.randn(N, K, device=dev).half()
    r = torch.randn(M, N, device=dev).half() if res else None
    bias = torch.randn(N, device=dev)
    for _ in range(3):
        ops.gemm(a, b, bias=bias, residual=r)
    L.dwg_gemm_trace(tr.data_ptr())
    rows = []
    for _ in range(5):
        ops.gemm(a, b, bias=bias, residual=r)
        ops.gemm(a, b, bias=bias, residual=r)
        torch.cuda.synchronize()
        t = tr.cpu().tolist()
        rows.append([t[i] - t[0] for i in range(13)])
    L.dwg_gemm_trace(None)
    med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(13)]
    print(f'M{M} N{N} K{K} res={int(res)}: ' + '  '.join(f'{n}={v / 1e3:.2f}us' for n, v in zip(names, med)))

# ---- weight-bound convolutions (2 x 8 x 8 and 2 x 16 x 16, 1280 channels): timeline of CTA 0 + warm / cold launch times
plan = (__import__('ctypes').c_int * 3)()
for Ni, H, C in ((2, 8, 1280), (2, 16, 1280), (2, 32, 640)):
    x = torch.randn(Ni, H, H, C, device=dev).half()
    ws = [(torch.randn(C, 3, 3, C, device=dev) * 0.02).half() for _ in range(8)]
    bias = torch.randn(C, device=dev)
    for _ in range(3):
        ops.conv2d_nhwc(x, ws[0], bias=bias)
    L.dwg_gemm_last_plan(plan)
    L.dwg_gemm_trace(tr.data_ptr())
    rows = []
    for _ in range(5):
        ops.conv2d_nhwc(x, ws[0], bias=bias)
        ops.conv2d_nhwc(x, ws[0], bias=bias)
        torch.cuda.synchronize()
        t = tr.cpu().tolist()
        rows.append([t[i] - t[0] for i in range(13)])
    L.dwg_gemm_trace(None)
    med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(13)]

    def timed(fn, n=40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / n
    warm = timed(lambda i: ops.conv2d_nhwc(x, ws[0], bias=bias))
    cold = timed(lambda i: ops.conv2d_nhwc(x, ws[i % 8], bias=bias))
    print(f'conv3x3 {Ni}x{H}x{H} {C}->{C} plan BN={plan[0]} ks={plan[1]} pair={L.dwg_gemm_last_pair()} halo={L.dwg_gemm_last_halo()}: '
          f'warm {warm:.1f} us, cold-weights {cold:.1f} us (back-to-back);  ' + '  '.join(f'{n}={v / 1e3:.2f}us' for n, v in zip(names, med)))
