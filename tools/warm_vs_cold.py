"""How much do the weight-bound GEMMs / convs of the 1280-channel levels gain when their weights are already in L2?
cold = 16 rotating weight copies (> L2 for the big ones, and always evicted by the flush), warm = one copy reused."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402

DEV = 'cuda'
flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, ncopy, warm):
    n = 8
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.cuda.graph(g):
        for i in range(n):
            fn(i % ncopy)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        if not warm:
            flush.zero_()
        else:
            for i in range(ncopy):
                fn(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / n)
    return best


torch.manual_seed(0)
tot_c = tot_w = 0.0
for (M, N, K, cnt) in ((512, 1280, 1280, 35), (512, 10240, 1280, 7), (512, 1280, 5120, 7), (128, 1280, 1280, 12), (2048, 640, 640, 35), (2048, 5120, 640, 7),
                       (2048, 640, 2560, 7), (8192, 320, 320, 35)):
    a = torch.randn(M, K, device=DEV).half()
    bs = [torch.randn(N, K, device=DEV).half() for _ in range(8)]
    c = timeit(lambda i: ops.gemm(a, bs[i]), 8, False)
    w = timeit(lambda i: ops.gemm(a, bs[i]), 1, True)
    tot_c += c * cnt; tot_w += w * cnt
    print(f'gemm M{M} N{N} K{K}: cold {c:6.1f} us  warm {w:6.1f} us  x{cnt}', flush=True)
for (Ni, H, Ci, Co, cnt) in ((2, 8, 1280, 1280, 19), (2, 16, 1280, 1280, 10), (2, 8, 2560, 1280, 3), (2, 16, 2560, 1280, 2), (2, 32, 640, 640, 9), (2, 32, 1280, 640, 1)):
    x = torch.randn(Ni, H, H, Ci, device=DEV).half()
    ws = [(torch.randn(Co, 3, 3, Ci, device=DEV) * 0.02).half() for _ in range(8)]
    c = timeit(lambda i: ops.conv2d_nhwc(x, ws[i]), 8, False)
    w = timeit(lambda i: ops.conv2d_nhwc(x, ws[i]), 1, True)
    tot_c += c * cnt; tot_w += w * cnt
    print(f'conv3x3 {Ni}x{H}x{H} {Ci}->{Co}: cold {c:6.1f} us  warm {w:6.1f} us  x{cnt}', flush=True)
print(f'weighted per step: cold {tot_c / 1e3:.3f} ms  warm {tot_w / 1e3:.3f} ms')
