"""Golden vectors for (f2) the optimiser step: the reference's own learning-rate schedule
(core/optim/optim_utils.py get_expon_lr_func, extracted with ``ast``) and torch.optim.Adam itself (the optimiser the
reference instantiates, core/gaussian/gaussian_optimizer.py:95, core/system/avatar.py:1626,1091) with the reference's
group settings, on seeded parameters / gradients.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_optim_golden.py      -> tests/golden/optim.npz
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/core/optim/optim_utils.py'

SHAPES = {'positions': (37, 3), 'scales': (37, 3), 'quaternions': (37, 4), 'grid': (90, 2), 'mlp_w': (64, 32), 'mlp_b': (64,),
          'bary': (11, 6, 3), 'mesh_scales': (66, 3)}
STEPS = 6


def make_inputs():
    g = torch.Generator().manual_seed(123)
    params = {k: torch.randn(*s, generator=g) * 0.1 for k, s in SHAPES.items()}
    grads = [{k: torch.randn(*s, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g))) for k, s in SHAPES.items()}
             for _ in range(STEPS)]
    return params, grads


def main():
    ns = {'np': np}
    for node in ast.parse(open(SRC).read()).body:
        if isinstance(node, ast.FunctionDef) and node.name == 'get_expon_lr_func':
            exec(compile(ast.Module(body=[node], type_ignores=[]), SRC, 'exec'), ns)
    iters = 5000
    sched = ns['get_expon_lr_func'](lr_init=0.00016, lr_final=0.0000016, lr_delay_mult=0.01, max_steps=iters * 2)
    its = np.array([0, 1, 2, 10, 100, 1234, 4999, 5000, 9999, 10000, 20000], np.int64)
    lr_sched = np.array([sched(int(i)) for i in its], np.float64)
    params, grads = make_inputs()
    P = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
    avatar = torch.optim.Adam([{'params': [P['positions']], 'lr': 0.00016, 'name': 'positions'},
                               {'params': [P['scales']], 'lr': 0.0025, 'name': 'scales'},
                               {'params': [P['quaternions']], 'lr': 0.001, 'name': 'quaternions'}], lr=0.0, eps=1e-15)
    nerf = torch.optim.Adam([{'params': [P['grid']], 'lr': 1e-3 * 10}, {'params': [P['mlp_w'], P['mlp_b']], 'lr': 1e-3}],
                            betas=(0.9, 0.99), eps=1e-15, weight_decay=0)
    mesh = torch.optim.Adam([{'params': [P['bary']], 'lr': 0.00016, 'name': 'bary_coords'},
                             {'params': [P['mesh_scales']], 'lr': 0.0025, 'name': 'scales'}], lr=0.0, eps=1e-15)
    out = {'lr_its': its, 'lr_sched': lr_sched}
    spatial_scale = 1.7
    for t in range(STEPS):
        # trainer.py:863-868: update_learning_rate(iteration=train_step, spatial_scale) before backward (train_step starts at 1)
        for gparam in avatar.param_groups:
            if gparam['name'] == 'positions':
                gparam['lr'] = sched(t + 1) * spatial_scale
            elif gparam['name'] == 'scales':
                gparam['lr'] = 0.0025 * spatial_scale
        for k in P:
            P[k].grad = grads[t][k].clone()
        for o in (avatar, nerf, mesh):
            o.step()
        for k in P:
            out[f'step{t}.{k}'] = P[k].detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'optim.npz'), **out)
    print('wrote optim.npz')


if __name__ == '__main__':
    main()
