"""(f3) Checkpoint / weight loading: the safetensors container, diffusers-format state dicts checked against the
architecture's key set, and the reference's avatar checkpoint format with the reset-to-checkpoint-size semantics
(core/trainer.py:194-259, core/system/scene.py:188-207, core/system/avatar.py:1254-1281)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dreamwaltz-g_b200'))
from dwg import checkpoint as ck  # noqa: E402
from dwg.diffusion import weights as W  # noqa: E402


def test_safetensors_roundtrip_and_diffusers_key_check(tmp_path):
    sd = W.make_unet(W.TINY)
    p32, p16 = str(tmp_path / 'unet.safetensors'), str(tmp_path / 'diffusion_pytorch_model.fp16.safetensors')
    ck.save_safetensors(p32, sd, metadata={'format': 'pt'})
    ck.save_safetensors(p16, {k: v.half() for k, v in sd.items()})
    back = ck.load_safetensors(p32)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    # the fp16 variant found inside a model directory, widened to fp32, extra keys dropped, key set enforced
    extra = dict(back)
    extra['decoder.conv_in.weight'] = torch.zeros(2, 2)
    ck.save_safetensors(str(tmp_path / 'with_extra.safetensors'), extra)
    got = ck.load_diffusers_state_dict(str(tmp_path), expected_keys=sd.keys())
    assert set(got) == set(sd) and all(v.dtype == torch.float32 for v in got.values())
    k0 = 'down_blocks.0.resnets.0.conv1.weight'
    torch.testing.assert_close(got[k0], sd[k0].half().float(), rtol=0, atol=0)
    assert set(ck.load_diffusers_state_dict(str(tmp_path / 'with_extra.safetensors'), expected_keys=sd.keys())) == set(sd)
    part = {k: v for k, v in sd.items() if not k.startswith('mid_block')}
    ck.save_safetensors(str(tmp_path / 'part.safetensors'), part)
    with pytest.raises(KeyError, match='missing'):
        ck.load_diffusers_state_dict(str(tmp_path / 'part.safetensors'), expected_keys=sd.keys())
    # torch pickle (.bin) path
    torch.save(sd, str(tmp_path / 'unet.bin'))
    assert set(ck.load_diffusers_state_dict(str(tmp_path / 'unet.bin'), expected_keys=sd.keys())) == set(sd)


@pytest.mark.gpu
def test_avatar_checkpoint_roundtrip_resets_to_checkpoint_size(tmp_path):
    from dwg import avatar as dav, camera, optim, parallel, step as dstep, synth
    dev = 'cuda'
    model = synth.make_body_model(0)
    a = dstep.Scene(dav.DreamWaltzGAvatar(model, synth.make_avatar(model, 900, 60, seed=1), device=dev))
    b = dstep.Scene(dav.DreamWaltzGAvatar(model, synth.make_avatar(model, 500, 40, seed=2), device=dev))
    with torch.no_grad():
        a.avatar.nerf_encoder.embeddings.uniform_(-0.3, 0.3)
    sd = a.state_dict()
    assert {'avatar._positions', 'avatar._scales', 'avatar._quaternions', 'avatar._lbs_weights', 'avatar.nerf_encoder.embeddings',
            'avatar.nerf_opacity_and_color_net.net.0.weight', 'avatar.nerf_scale_and_quaternion_net.gaussian_warp.bias',
            'avatar.mesh_binding_gaussians.hands._bary_coords'} <= set(sd)              # reference key names
    bucket = parallel.GradBucket([p for p in a.parameters() if p.requires_grad])
    opt = optim.AvatarOptimizer(a.avatar, bucket, iterations=5000)
    bucket.flat.normal_()
    opt.step()
    path = ck.save_checkpoint(str(tmp_path / 'step_000007.pth'), a, 7, past_checkpoints=['step_000007.pth'], optimizers=[opt], full=True)
    step, missing, unexpected = ck.load_checkpoint(path, b, model_only=True)
    assert step == 7 and not missing and not unexpected
    assert b.avatar._positions.shape == (900, 3) and b.avatar._positions.requires_grad and not b.avatar._lbs_weights.requires_grad
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k
    d = camera.make_camera(2.0, 20.0, 85.0, 50.0, 64, 64)
    pose = {k: v.to(dev) for k, v in synth.pose_from_row(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'poses.npz'))['rows'][3]).items()}
    with torch.no_grad():
        ia = a(d, smpl_observed_inputs=pose, use_densifier=False)['image']
        ib = b(d, smpl_observed_inputs=pose, use_densifier=False)['image']
    assert torch.equal(ia, ib)
    # bare state-dict form (trainer.py:203-206)
    torch.save(a.state_dict(), str(tmp_path / 'bare.pth'))
    c = dstep.Scene(dav.DreamWaltzGAvatar(model, synth.make_avatar(model, 300, 40, seed=3), device=dev))
    step, missing, unexpected = ck.load_checkpoint(str(tmp_path / 'bare.pth'), c)
    assert step is None and c.avatar._positions.shape == (900, 3)
    # optimiser state travels with full checkpoints
    bucket_b = parallel.GradBucket([p for p in b.parameters() if p.requires_grad])
    opt_b = optim.AvatarOptimizer(b.avatar, bucket_b, iterations=5000)
    ck.load_checkpoint(path, b, optimizers=[opt_b])
    assert int(opt_b.step_dev) == 1 and torch.equal(opt_b.exp_avg, opt.exp_avg)
