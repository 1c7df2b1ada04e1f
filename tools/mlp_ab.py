"""A/B of the avatar-MLP kernels: fp32 SIMT vs tensor-core (fp16 hi+lo split) forward and backward, N = 150k."""
import sys, os, torch, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dreamwaltz-g_b200'))
from dwg import ops, _lib
L = _lib.lib()
dev = 'cuda'
torch.manual_seed(0)
N, Nu = 150000, 135000
shapes = [(64, 32), (64,), (64, 64), (64,), (4, 64), (4,), (64, 95), (64,), (64, 64), (64,), (64, 64), (64,), (64, 64), (64,), (3, 64), (3,), (3, 64), (3,)]
params = [(torch.randn(*s) * 0.3).to(dev).requires_grad_(True) for s in shapes]
enc = ((torch.rand(N, 32) - 0.5)).to(dev).requires_grad_(True)
pos = torch.randn(Nu, 3).to(dev).requires_grad_(True)
pose = (torch.randn(1, 63) * 0.5).to(dev)
res = {}
for tc in (0, 1):
    L.dwg_avatar_mlp_set_tc(tc)
    out = ops.avatar_mlp(enc, pos, pose, params, Nu)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        with torch.no_grad():
            o2 = ops.avatar_mlp(enc, pos, pose, params, Nu)
    b.record(); torch.cuda.synchronize()
    loss = sum((o * (i + 1)).sum() for i, o in enumerate(out))
    g = torch.autograd.grad(loss, [enc, pos] + params)
    res[tc] = ([o.detach().clone() for o in out], [x.clone() for x in g], a.elapsed_time(b) / 20)
print('fwd ms: simt %.4f tc %.4f' % (res[0][2], res[1][2]))
for tc in (0, 1):
    L.dwg_avatar_mlp_set_tc(tc)
    ws = [torch.ones_like(o) for o in ops.avatar_mlp(enc, pos, pose, params, Nu)]
    def fb():
        out = ops.avatar_mlp(enc, pos, pose, params, Nu)
        torch.autograd.grad(out, [enc, pos] + params, ws)
    for _ in range(3):
        fb()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fb()
    b.record(); torch.cuda.synchronize()
    print('fwd+bwd ms (incl. host overhead):', 'tc' if tc else 'simt', a.elapsed_time(b) / 20)
for i, n in enumerate(('colors', 'opac', 'pos', 'scales')):
    d = (res[0][0][i] - res[1][0][i]).abs().max().item(); m = res[0][0][i].abs().max().item()
    print(n, 'max abs diff', d, 'max', m)
for i in range(len(res[0][1])):
    d = (res[0][1][i] - res[1][1][i]).abs().max().item(); m = res[0][1][i].abs().max().item()
    print('grad', i, 'rel diff', d / (m + 1e-30))
from torch.profiler import ProfilerActivity, profile
for tc in (0, 1):
    L.dwg_avatar_mlp_set_tc(tc)
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        out = ops.avatar_mlp(enc, pos, pose, params, Nu)
        torch.autograd.grad(out, [enc, pos] + params, ws)
        torch.cuda.synchronize()
    for e in pr.events():
        if 'mlp' in e.name:
            print('tc' if tc else 'simt', e.name[:60], '%.1f us' % e.device_time)
