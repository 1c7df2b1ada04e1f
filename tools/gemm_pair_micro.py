"""Single CTA vs CTA pair (tcgen05 cta_group::2) on the large GEMM / conv shapes of the SDS step: cold weights
(rotating copies), CUDA-graph replays, CUDA events.      python tools/gemm_pair_micro.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402
from dwg._lib import lib  # noqa: E402

DEV, NCOPY = 'cuda', 8
L = lib()


def timeit(fn):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.cuda.graph(g):
        for i in range(NCOPY):
            fn(i)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / NCOPY)
    return best


def run(name, fn, flops, bns, conv=False):
    plan = (ctypes.c_int * 3)()
    out = []
    for halo in ((0, 1) if conv else (0,)):
        L.dwg_gemm_tune_halo(halo, 0)
        for pair in (0, 1):
            L.dwg_gemm_tune_pair(pair)
            for bn in bns:
                L.dwg_gemm_tune(bn, 1)
                fn(0)
                if L.dwg_gemm_last_pair() != pair or (conv and L.dwg_gemm_last_halo() != halo):
                    continue
                L.dwg_gemm_last_plan(plan)
                t = timeit(fn)
                out.append((t, ('H' if halo else '') + ('P' if pair else 'S'), plan[0], plan[2]))
    L.dwg_gemm_tune(0, 0); L.dwg_gemm_tune_pair(-1); L.dwg_gemm_tune_halo(-1, 0)
    t_auto = timeit(fn)
    s = '  '.join(f"{m}{bn}/{st}st {t:6.1f}" for t, m, bn, st in out)
    best = min(out)
    print(f'{name:32s} auto {t_auto:6.1f}us ({flops / t_auto / 1e6:5.0f} TF/s) | best {best[1]:2s} BN{best[2]} {best[0]:6.1f}us ({flops / best[0] / 1e6:5.0f} TF/s) | {s}', flush=True)


def main():
    torch.manual_seed(0)
    for M, N, K in () if '--conv' in sys.argv else ((8192, 4096, 4096), (8192, 320, 320), (8192, 2560, 320), (8192, 320, 1280), (2048, 640, 640), (2048, 5120, 640),
                    (2048, 640, 2560), (512, 1280, 1280), (512, 10240, 1280), (512, 1280, 5120), (4096, 512, 512), (4096, 512, 4096)):
        a = torch.randn(M, K, device=DEV).half()
        bs = [torch.randn(N, K, device=DEV).half() for _ in range(NCOPY)]
        run(f'gemm M{M} N{N} K{K}', lambda i: ops.gemm(a, bs[i % NCOPY]), 2.0 * M * N * K, (128, 160, 256))
    if '--tiny-channels' in sys.argv:
        # the 3/8/16/32-channel convolutions at 512^2 / 256^2 (VAE conv_in and its dgrad, ControlNet condition embedding)
        for Ni, H, Ci, Co in ((1, 512, 128, 3), (1, 512, 8, 128), (1, 512, 8, 16), (1, 512, 16, 16), (1, 256, 32, 32), (1, 128, 96, 96)):
            x = torch.randn(Ni, H, H, Ci, device=DEV).half()
            ws = [(torch.randn(Co, 3, 3, Ci, device=DEV) * 0.02).half() for _ in range(NCOPY)]
            run(f'conv3x3 {Ni}x{H}x{H} {Ci}->{Co}', lambda i: ops.conv2d_nhwc(x, ws[i % NCOPY], out_dtype=torch.float32 if Co == 3 else torch.float16),
                2.0 * Ni * H * H * Ci * Co * 9, (32, 64, 96, 128), conv=True)
        return
    for Ni, H, Ci, Co in ((1, 512, 128, 128), (1, 256, 256, 256), (1, 128, 512, 512), (1, 64, 512, 512), (2, 64, 320, 320), (2, 32, 640, 640),
                          (2, 16, 1280, 1280), (2, 8, 1280, 1280), (2, 64, 640, 320), (2, 32, 1280, 640)):
        x = torch.randn(Ni, H, H, Ci, device=DEV).half()
        ws = [(torch.randn(Co, 3, 3, Ci, device=DEV) * 0.02).half() for _ in range(NCOPY)]
        run(f'conv3x3 {Ni}x{H}x{H} {Ci}->{Co}', lambda i: ops.conv2d_nhwc(x, ws[i % NCOPY]), 2.0 * Ni * H * H * Ci * Co * 9, (128, 160, 256), conv=True)


if __name__ == '__main__':
    main()
