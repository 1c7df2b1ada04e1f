"""(f2) The fused optimiser step against the reference's own optimisers: torch.optim.Adam with the reference's group
settings and the reference's get_expon_lr_func schedule (tests/golden/make_optim_golden.py -> optim.npz)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'dreamwaltz-g_b200'))
G = np.load(os.path.join(HERE, 'golden', 'optim.npz'))
_spec = importlib.util.spec_from_file_location('make_optim_golden', os.path.join(HERE, 'golden', 'make_optim_golden.py'))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)


def test_position_lr_schedule_matches_reference_function():
    from dwg import optim
    got = np.array([optim.expon_lr(int(i), 0.00016, 0.0000016, 10000) for i in G['lr_its']])
    np.testing.assert_allclose(got, G['lr_sched'], rtol=1e-12)
    assert got[0] == pytest.approx(0.00016) and got[-1] == pytest.approx(0.0000016)


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam_with_reference_groups():
    from dwg import optim, parallel
    dev = 'cuda'
    params, grads = mk.make_inputs()
    P = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in params.items()}
    bucket = parallel.GradBucket(list(P.values()))
    e15 = dict(eps=1e-15)
    groups = [dict(name='positions', params=[P['positions']], lr=0.00016, **e15), dict(name='scales', params=[P['scales']], lr=0.0025, **e15),
              dict(name='quaternions', params=[P['quaternions']], lr=0.001, **e15),
              dict(name='grid', params=[P['grid']], lr=1e-2, betas=(0.9, 0.99), **e15),
              dict(name='mlps', params=[P['mlp_w'], P['mlp_b']], lr=1e-3, betas=(0.9, 0.99), **e15),
              dict(name='bary', params=[P['bary']], lr=0.00016, **e15), dict(name='mesh_scales', params=[P['mesh_scales']], lr=0.0025, **e15)]
    opt = optim.FusedAdam(bucket, groups)
    assert all(p.data_ptr() == opt.flat_params.data_ptr() + o * 4 for p, o in zip(bucket.params, bucket.offsets))     # parameters re-homed as views
    for t in range(mk.STEPS):
        groups[0]['lr'] = optim.expon_lr(t + 1, 0.00016, 0.0000016, 10000) * 1.7
        groups[1]['lr'] = 0.0025 * 1.7
        opt.set_lrs()
        opt.zero_grad()
        for k in P:
            P[k].grad.copy_(grads[t][k].to(dev))
        opt.step()
        for k in P:
            # eps = 1e-15 makes the first steps sign-like (|update| = lr): compare on the parameter scale
            np.testing.assert_allclose(P[k].detach().cpu().numpy(), G[f'step{t}.{k}'], rtol=2e-6, atol=2e-8, err_msg=f'step {t} {k}')
    assert int(opt.step_dev) == mk.STEPS
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    assert int(opt.step_dev) == mk.STEPS
