"""GPU: (f4 / cfg5) the re-enactment inference path: frame_pack against the reference's post-processing arithmetic
(core/trainer.py:1068-1084 + utils/image.py:52-61 restated with numpy, as the reference runs it) and the graph-replayed
Reenactor against the eager per-frame render."""
import os

import numpy as np
import pytest
import torch

from dwg import avatar as dav, camera, inference, step as dstep, synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def tensor2image_np(t):          # utils/image.py:60-61
    return (t.detach().cpu().numpy() * 255.0).clip(0.0, 255.0).astype(np.uint8)


@pytest.mark.parametrize('H,W', [(64, 64), (48, 36), (7, 5)])
def test_frame_pack_matches_reference_postprocessing(H, W):
    g = torch.Generator().manual_seed(H * W)
    img = (torch.rand(3, H, W, generator=g) * 1.4 - 0.2).to(DEV)          # includes values outside [0,1]
    fg = (torch.rand(3, H, W, generator=g) * 1.2 - 0.1).to(DEV)
    depth = (torch.rand(H, W, generator=g) * 4.0).to(DEV)
    alpha = torch.rand(H, W, generator=g).to(DEV)
    img[0, 0, 0], img[1, 0, 0] = 1.0, 254.5 / 255.0
    out = inference.frame_pack(img, fg, depth, alpha, depth_div=3.0)
    assert np.array_equal(out['image'].cpu().numpy(), tensor2image_np(img.permute(1, 2, 0)))
    assert np.array_equal(out['image_fg'].cpu().numpy(), tensor2image_np(torch.cat([fg.permute(1, 2, 0), alpha[..., None]], dim=2)))     # concat_alpha
    assert np.array_equal(out['depth'].cpu().numpy(), tensor2image_np(depth / 3.0))
    assert np.array_equal(out['alpha'].cpu().numpy(), tensor2image_np(alpha))


def test_reenactor_graph_replay_delivers_every_frame():
    model = synth.make_body_model(0)
    sc = dstep.Scene(dav.DreamWaltzGAvatar(model, synth.make_avatar(model, 2500, 150, seed=5), device=DEV))
    with torch.no_grad():
        sc.avatar.nerf_encoder.embeddings.uniform_(-0.5, 0.5)
        sc.avatar._scales.fill_(np.log(0.02))
    rows = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'poses.npz'))['rows']
    frames = []
    for i in range(7):
        d = camera.make_camera(2.2, 25.0 * i, 85.0, 50.0, 96, 96)
        d['smpl_inputs'] = synth.pose_from_row(rows[i % len(rows)])
        frames.append(d)
    re = inference.Reenactor(sc, bg_mode='white', ring=3)
    eager = []
    for d in frames:
        dd = dict(d, smpl_inputs={k: v.to(DEV) for k, v in d['smpl_inputs'].items()})
        eager.append({k: v.cpu().numpy().copy() for k, v in re.render(dd).items()})
    re.capture(frames[0])
    got = {}
    n = re.run(frames, on_frame=lambda i, fr: got.__setitem__(i, {k: v.numpy().copy() for k, v in fr.items()}))
    assert n == 7 and sorted(got) == list(range(7))
    for i in range(7):
        assert set(got[i]) == {'image', 'image_fg', 'depth', 'alpha'}
        for k in got[i]:
            assert np.array_equal(got[i][k], eager[i][k]), (i, k)
    assert got[0]['image'].shape == (96, 96, 3) and got[0]['image_fg'].shape == (96, 96, 4)
    assert not np.array_equal(got[0]['image'], got[3]['image'])
