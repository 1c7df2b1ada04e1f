"""One GEMM shape, a few launches (for `ncu -k regex:gemm_kernel -s 3 -c 1`).  python tools/gemm_one.py M N K [res]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
res = len(sys.argv) > 4
a = torch.randn(M, K, device='cuda').half()
b = torch.randn(N, K, device='cuda').half()
bias = torch.randn(N, device='cuda')
r = torch.randn(M, N, device='cuda').half() if res else None
for _ in range(5):
    ops.gemm(a, b, bias=bias, residual=r)
torch.cuda.synchronize()
