"""GEMM micro-benchmark on the UNet / ControlNet / VAE shapes: dwg tcgen05 kernel vs cuBLAS
(torch.matmul fp16) under identical conditions (back-to-back launches, CUDA events).
python tools/gemm_micro.py [--ncu]   (--ncu: a few launches per shape only, for an ncu capture)"""
import os
import sys

import torch

NCU = '--ncu' in sys.argv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402

SHAPES = [  # M, N, K, act, residual
    (8192, 320, 320, None, False), (8192, 320, 320, None, True), (8192, 2560, 320, 'geglu', False), (8192, 320, 1280, None, True),
    (2048, 640, 640, None, False), (2048, 5120, 640, 'geglu', False), (2048, 640, 2560, None, True),
    (512, 1280, 1280, None, False), (512, 10240, 1280, 'geglu', False), (512, 1280, 5120, None, True),
    (128, 1280, 1280, None, False), (4096, 512, 512, None, False), (640, 1024, 640, None, False), (1280, 256, 1280, None, False),
    (8192, 4096, 4096, None, False),
]
CONVS = [  # Nimg, H, W, Cin, Cout, k
    (1, 512, 512, 128, 128, 3), (1, 64, 64, 512, 512, 3), (2, 64, 64, 320, 320, 3), (2, 32, 32, 640, 640, 3), (2, 16, 16, 1280, 1280, 3),
    (2, 8, 8, 1280, 1280, 3), (2, 16, 16, 1280, 1280, 1), (2, 64, 64, 320, 320, 1),
]


def timeit(fn, iters):
    """us per launch of `iters` back-to-back launches replayed from a CUDA graph (no host overhead)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if NCU:
        return 0.0
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


def main():
    ncu = NCU
    iters = 20
    shapes = [SHAPES[i] for i in (0, 2, 7)] if ncu else SHAPES
    convs = CONVS[2:3] if ncu else CONVS
    dev = 'cuda'
    torch.manual_seed(0)
    print(f'{"shape":40s} {"dwg us":>8s} {"TFLOP/s":>8s} {"cublas us":>9s} {"ratio":>6s}')
    for M, N, K, act, res in shapes:
        a = torch.randn(M, K, device=dev).half()
        b = torch.randn(N, K, device=dev).half()
        bias = torch.randn(N, device=dev)
        r = torch.randn(M, N, device=dev).half() if res else None
        t = timeit(lambda: ops.gemm(a, b, bias=bias, act=act, residual=r), iters)
        tc = 0.0 if ncu else timeit(lambda: torch.addmm(bias.half(), a, b.t()), iters)
        fl = 2.0 * M * N * K
        print(f'{f"gemm M{M} N{N} K{K} {act or chr(45)} res={int(res)}":40s} {t:8.1f} {fl / t / 1e6:8.1f} {tc:9.1f} {t / max(tc, 1e-9):6.2f}', flush=True)
    for Ni, H, W, Ci, Co, k in convs:
        x = torch.randn(Ni, H, W, Ci, device=dev).half()
        w = torch.randn(Co, k, k, Ci, device=dev).half()
        bias = torch.randn(Co, device=dev)
        t = timeit(lambda: ops.conv2d_nhwc(x, w, bias=bias, padding=k // 2), iters)
        tc = 0.0
        if not ncu:
            xc = x.permute(0, 3, 1, 2)          # channels_last NCHW view
            wc = w.permute(0, 3, 1, 2)
            tc = timeit(lambda: torch.nn.functional.conv2d(xc, wc, bias.half(), padding=k // 2), iters)
        fl = 2.0 * Ni * H * W * Ci * Co * k * k
        print(f'{f"conv{k}x{k} {Ni}x{H}x{W} {Ci}->{Co}":40s} {t:8.1f} {fl / t / 1e6:8.1f} {tc:9.1f} {t / max(tc, 1e-9):6.2f}', flush=True)

    if not ncu:
        for B, h, T, Tk, hd in ((2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 4096, 77, 40), (2, 8, 256, 256, 160 // 1 if False else 128)):
            C = h * hd
            q = torch.randn(B, T, C, device=dev).half()
            k = torch.randn(B, Tk, C, device=dev).half()
            Tkp = (Tk + 7) // 8 * 8
            vt = torch.randn(B, C, Tkp, device=dev).half()
            t = timeit(lambda: ops.attention(q, k, vt, h, Tk), iters)
            sp = lambda x: x.view(B, -1, h, hd).transpose(1, 2)
            v = vt[:, :, :Tk].transpose(1, 2).contiguous()
            tc = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(sp(q), sp(k), sp(v)), iters)
            fl = 4.0 * B * h * T * Tk * hd
            print(f'{f"attention B{B} h{h} T{T} Tk{Tk} d{hd}":40s} {t:8.1f} {fl / t / 1e6:8.1f} {tc:9.1f} {t / max(tc, 1e-9):6.2f}', flush=True)


if __name__ == '__main__':
    main()
