"""Does alternating big-shared-memory tcgen05 GEMMs with small streaming kernels cost an L1/shared re-partition per switch?
Times a captured graph of [gemm, groupnorm(stats+apply), layernorm] x N with the streaming kernels' carve-out preference at
the driver default and at 100 %, and the same GEMMs back to back."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dreamwaltz-g_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

from dwg import ops  # noqa: E402
from dwg._lib import lib  # noqa: E402

dev = 'cuda'
torch.manual_seed(0)
N = 40
x = torch.randn(8192, 320, device=dev).half()
w = [torch.randn(320, 320, device=dev).half() * 0.05 for _ in range(N)]
g, b = torch.ones(320, device=dev), torch.zeros(320, device=dev)


def chain(with_norms):
    h = x
    for i in range(N):
        h = ops.gemm(h, w[i])
        if with_norms:
            h = ops.group_norm(h.view(2, 64, 64, 320), g, b, 32, 1e-5, silu=True).view(8192, 320)
            h = ops.layer_norm(h, g, b)
    return h


def timed(with_norms, label):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain(with_norms)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain(with_norms)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'{label:60s} {e0.elapsed_time(e1) / 20 / N * 1000:8.2f} us per iteration', flush=True)


timed(False, 'gemm only (M8192 N320 K320)')
lib().dwg_nn_set_carveout(-1)
timed(True, 'gemm + gn(stats, apply) + layernorm, default carve-out')
lib().dwg_nn_set_carveout(100)
timed(True, 'gemm + gn(stats, apply) + layernorm, carve-out 100')
lib().dwg_nn_set_carveout(0)
timed(True, 'gemm + gn(stats, apply) + layernorm, carve-out 0')
os.environ['X'] = '1'
