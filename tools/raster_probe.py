"""Per-tile timeline of the forward render kernel (dwg_raster_probe): where the 0.36 ms of render_fwd_kernel at the
benchmark's 150k / 512^2 view goes -- start / end of every tile relative to the first start, instances per tile, the
largest number of instances a warp's 8x4 block had to evaluate and the longest per-pixel blend chain.
    python tools/raster_probe.py [image_size]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import _lib, avatar as dav, camera, synth  # noqa: E402

DEV = 'cuda'
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model = synth.make_body_model(0)
av = synth.make_avatar(model, 135000, 2500, seed=0)
m = dav.DreamWaltzGAvatar(model, av, device=DEV)
with torch.no_grad():
    m.nerf_encoder.embeddings.uniform_(-0.5, 0.5)
rng = np.random.default_rng(0)
obs = {k: v.to(DEV) for k, v in synth.random_pose(rng).items()}
r = dav.GaussianRenderer()
L = _lib.lib()
T = 4 * ((S + 15) // 16) ** 2          # one probe row per 8x8-quadrant CTA
for view in ((2.0, 30.0, 85.0, 50.0), (2.2, 120.0, 70.0, 45.0)):
    data = camera.make_camera(*view, S, S)
    with torch.no_grad():
        gs = m.animate(obs)
        for _ in range(2):
            r.render(data, gs)
        buf = torch.zeros(T, 6, dtype=torch.int64, device=DEV)
        L.dwg_raster_probe(buf.data_ptr())
        r.render(data, gs)
        torch.cuda.synchronize()
        L.dwg_raster_probe(None)
    b = buf.cpu().numpy()
    t0 = b[:, 0].min()
    st, en, n, touch, blend, sm = (b[:, 0] - t0) / 1e3, (b[:, 1] - t0) / 1e3, b[:, 2], b[:, 3], b[:, 4], b[:, 5]
    dur = en - st
    print(f'view {view}: tiles {T}, kernel span {en.max():.1f} us, instances {n.sum()}, max tile {n.max()}')
    print(f'  start times: median {np.median(st):.1f} us, p90 {np.percentile(st, 90):.1f}, max {st.max():.1f}')
    order = np.argsort(-en)[:12]
    print('  last tiles to finish:  tile  start  end  dur  n  strip_touched  pixel_blended  sm  cycles/instance(touched)')
    for i in order:
        print(f'    {i:5d} {st[i]:7.1f} {en[i]:7.1f} {dur[i]:7.1f} {n[i]:6d} {touch[i]:6d} {blend[i]:6d} {sm[i]:4d}  ns/touched {1e3 * dur[i] / max(1, touch[i]):.1f}  ns/n {1e3 * dur[i] / max(1, n[i]):.1f}')
    # fit dur ~ a*n + b*touch
    A = np.stack([n, touch, np.ones_like(n)], 1).astype(np.float64)
    coef, *_ = np.linalg.lstsq(A, dur * 1e3, rcond=None)
    print(f'  least squares: dur[ns] = {coef[0]:.2f} * n + {coef[1]:.2f} * strip_touched + {coef[2]:.0f}')
    per_sm = {}
    for i in range(T):
        per_sm.setdefault(int(sm[i]), []).append(i)
    print(f'  SMs used {len(per_sm)}, tiles per SM max {max(len(v) for v in per_sm.values())}; sum of tile durations / span = {dur.sum() / en.max():.1f} tile-equivalents busy on average')
