"""Host-side pieces of the guidance surface that need no GPU (reference: core/guidance/controlnet.py:33-55, basic.py:404-418)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))


def _guidance_cls():
    from dwg.diffusion import guidance as G
    return G.ControlNetScoreDistillation


def test_prepare_image_accepts_what_the_reference_passes():
    """PIL image, list of PIL images (resized with LANCZOS to the model's size, /255, NCHW), tensor, list of tensors."""
    from PIL import Image
    cls = _guidance_cls()
    me = types.SimpleNamespace(device='cpu', default_image_size=64)
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, size=(48, 80, 3), dtype=np.uint8)
    pil = Image.fromarray(arr)
    out = cls.prepare_image(me, pil)
    assert out.shape == (1, 3, 64, 64) and out.dtype == torch.float32 and 0.0 <= float(out.min()) and float(out.max()) <= 1.0
    ref = np.array(pil.resize((64, 64), resample=Image.Resampling.LANCZOS)).astype(np.float32).transpose(2, 0, 1) / 255.0
    assert np.array_equal(out[0].numpy(), ref)
    out2 = cls.prepare_image(me, [pil, pil], width=32, height=16)
    assert out2.shape == (2, 3, 16, 32)
    t = torch.rand(1, 3, 64, 64)
    assert cls.prepare_image(me, t) is t or torch.equal(cls.prepare_image(me, t), t)
    assert cls.prepare_image(me, [t[0], t[0]]).shape == (2, 3, 64, 64)


def test_guidance_scale_schedules():
    """basic.py:404-418: constant / linear / linear_reverse between 7.5 and the initial scale."""
    cls = _guidance_cls()
    me = types.SimpleNamespace(initial_guidance_scale=50.0, guidance_adjust='constant')
    assert cls.get_guidance_scale(me, 1, 100) == 50.0
    me.guidance_adjust = 'linear'
    assert cls.get_guidance_scale(me, 1, 100) == 50.0 and abs(cls.get_guidance_scale(me, 100, 100) - 7.5) < 1e-9
    me.guidance_adjust = 'linear_reverse'
    assert cls.get_guidance_scale(me, 1, 100) == 7.5 and abs(cls.get_guidance_scale(me, 100, 100) - 50.0) < 1e-9
    me.guidance_adjust = 'uniform'
    np.random.seed(0)
    s = [cls.get_guidance_scale(me, 1, 100) for _ in range(50)]
    assert min(s) >= 7.5 and max(s) <= 50.0


def test_timestep_sampling_modes_follow_the_reference_formulas():
    """TimePrioritizedScheduler.get_timestep (core/guidance/time_prior.py:322-351) for min / max timestep 0.02 / 0.98."""
    cls = _guidance_cls()
    me = types.SimpleNamespace(t_lo=20, t_hi=980, time_sampling='constant', time_annealing='linear', device='cpu', use_default_generator=True, gen=None)
    me.scheduled_timestep = lambda a, b: cls.scheduled_timestep(me, a, b)
    assert cls.get_timestep(me, 2, 5, 100).tolist() == [500, 500]
    me.time_sampling = 'linear'
    delta = (980 - 20) / 99
    for step in (1, 2, 50, 100):
        assert cls.get_timestep(me, 1, step, 100).item() == int(980 - (step - 1) * delta)
    me.time_sampling = 'annealed'
    for step in (0, 10, 100):
        assert cls.get_timestep(me, 1, step, 100).item() == int(980 - 960 * (step / 100) ** 1.0)
    me.time_annealing = 'hifa,900,100'
    assert cls.get_timestep(me, 1, 25, 100).item() == int(900 - 800 * 0.25 ** 0.5)
    me.time_annealing = 'linear,900,100,2.0'
    assert cls.get_timestep(me, 1, 50, 100).item() == int(900 - 800 * 0.5 ** 2.0)
    me.time_sampling = 'stage-3'
    torch.manual_seed(0)
    per = (980 - 20) // 3
    for step, top in ((0, 20 + 3 * per), (40, 20 + 2 * per), (99, 20 + per)):
        t = cls.get_timestep(me, 512, step, 100)
        assert int(t.min()) >= 20 and int(t.max()) <= top and int(t.max()) > top - per // 4
    me.time_sampling = 'uniform'
    t = cls.get_timestep(me, 2048, 7, 100)
    assert int(t.min()) >= 20 and int(t.max()) <= 980 and int(t.max()) > 900 and int(t.min()) < 100
