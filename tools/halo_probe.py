"""Halo-mode 3x3 convolution: correctness under both descriptor base-offset conventions, then timing vs the per-tap path.
    python tools/halo_probe.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from dwg import ops  # noqa: E402
from dwg._lib import lib  # noqa: E402

DEV = 'cuda'
L = lib()


def rel(a, b):
    return float((a.float() - b).abs().max() / (b.abs().max() + 1e-12))


def check(bo, pair, N, H, Ci, Co):
    torch.manual_seed(1)
    x = torch.randn(N, H, H, Ci, device=DEV).half()
    w = (torch.randn(Co, 3, 3, Ci, device=DEV) / (9 * Ci) ** 0.5).half()
    bias = torch.randn(Co, device=DEV)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1).permute(0, 2, 3, 1)
    L.dwg_gemm_tune_halo(1, bo)
    L.dwg_gemm_tune_pair(pair)
    y = ops.conv2d_nhwc(x, w, bias=bias, out_dtype=torch.float32)
    torch.cuda.synchronize()
    took = (L.dwg_gemm_last_halo(), L.dwg_gemm_last_pair())
    L.dwg_gemm_tune_halo(-1, 1)
    L.dwg_gemm_tune_pair(-1)
    return rel(y, ref), took


if __name__ == '__main__':
    for bo in (1, 0):
        for pair in (0, 1):
            for shape in ((1, 32, 64, 64), (2, 64, 320, 320), (1, 128, 128, 128), (1, 64, 8, 32)):
                e, took = check(bo, pair, *shape)
                print(f'base_offset_mode={bo} pair={pair} shape={shape}: rel err {e:.2e} (halo, pair)={took}', flush=True)
