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
