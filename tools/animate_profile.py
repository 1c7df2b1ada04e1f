"""Kernel table of avatar.animate forward + backward alone (un-graphed).  python tools/animate_profile.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
import bench  # noqa: E402

dev = 'cuda:0'
sc = bench.Scene.__new__(bench.Scene)
from dwg import avatar as dav, synth  # noqa: E402
model = synth.make_body_model(0)
av = synth.make_avatar(model, bench.N_UNCONSTRAINED, bench.N_MESH_TRI, seed=0)
avatar = dav.DreamWaltzGAvatar(model, av, device=dev)
pose = {k: v.to(dev) for k, v in synth.pose_from_row(bench.poses()[0]).items()}
params = [p for p in avatar.parameters() if p.requires_grad]


def step():
    for p in params:
        p.grad = None
    gs = avatar.animate(pose)
    loss = gs.positions.sum() + gs.colors.sum() + gs.opacities.sum() + gs.scales.sum() + gs.quaternions.sum()
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s), torch.cuda.graph(g):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f'animate fwd+bwd (graphed): {a.elapsed_time(b) / 10:.3f} ms')
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
    step()
    torch.cuda.synchronize()
print(pr.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=60))
