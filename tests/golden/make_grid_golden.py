"""Golden vectors for row R6 (multi-resolution grid encoder) from the reference's OWN CUDA kernel.

The reference's gridencoder has no CPU path, so the vectors are produced on a B200 box:

    python -m oracle.build_ref                                   # here: compiles the reference sources
                                                                 # (core/nerf/gridencoder/src/*, unmodified) -> oracle/_ref/
    gpurun -- python tests/golden/make_grid_golden.py            # there: runs it, writes gpurun_out/grid.npz
    cp gpurun_out/grid.npz tests/golden/grid.npz                 # here: commit the fixture

Calls follow core/nerf/gridencoder/grid.py:28-94 (_grid_encode.forward/backward): [L,B,C] staging
layout, dy_dx buffer, zero-initialised grad_embeddings / grad_inputs.  The embedding table is a
closed-form integer-hash pattern (table_pattern) so the 48 MB avatar table never has to be stored;
grad_embeddings is stored sparsely (flat indices of the non-zero entries + values) -- the index
set is exactly the set of grid corners the kernel touched (bit-exact uint32 indices).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def table_pattern(rows, C=2):
    """Deterministic, version-independent table in [-0.5, 0.5): exact in float32."""
    i = np.arange(rows * C, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + (i >> np.uint64(7)) * np.uint64(40503)) & np.uint64(0xFFFF)
    return (h.astype(np.float32) / np.float32(65536.0) - np.float32(0.5)).reshape(rows, C)


def level_offsets(L, base, pls, log2_hashmap, align_corners, D=3):
    """grid.py:120-133."""
    offs, off = [], 0
    for i in range(L):
        res = int(np.ceil(base * pls ** i))
        params = min(2 ** log2_hashmap, (res if align_corners else res + 1) ** D)
        params = int(np.ceil(params / 8) * 8)
        offs.append(off)
        off += params
    offs.append(off)
    return np.array(offs, np.int32)


CASES = {
    # the avatar's encoder (avatar.py:1141-1150 -> GridEncoder(tiled, smoothstep, L16, C2, base 16, desired 4096, 2^19))
    'avatar': dict(L=16, base=16, desired=4096, log2=19, gridtype=1, align=False, interp=1, B=160, seed=11),
    'hash_linear': dict(L=8, base=16, desired=512, log2=14, gridtype=0, align=False, interp=0, B=128, seed=12),
    'tiled_align': dict(L=6, base=8, desired=128, log2=16, gridtype=1, align=True, interp=0, B=96, seed=13),
    'hash_smooth': dict(L=12, base=16, desired=2048, log2=15, gridtype=0, align=False, interp=1, B=96, seed=14),
}


def case_inputs(c):
    rng = np.random.default_rng(c['seed'])
    B = c['B']
    x = rng.uniform(0.02, 0.98, size=(B, 3)).astype(np.float32)
    x[0] = (0.0, 0.0, 0.0)
    x[1] = (1.0, 1.0, 1.0)
    x[2] = (0.5, 0.5, 0.5)
    x[3] = (1.0 + 1e-3, 0.3, 0.3)            # out of range -> zeros (gridencoder.cu:110-135)
    x[4] = (0.2, -1e-4, 0.9)
    x[5] = (0.999999, 1e-7, 0.25)
    grad = rng.standard_normal(size=(B, c['L'] * 2)).astype(np.float32)
    grad[np.abs(grad) < 1e-3] = 1e-3         # every touched corner receives a non-zero gradient
    return x, grad


def main():
    import torch
    from oracle import build_ref
    ge = build_ref.load_module()
    dev = 'cuda'
    out = {}
    for name, c in CASES.items():
        L, C, D = c['L'], 2, 3
        pls = float(np.exp2(np.log2(c['desired'] / c['base']) / (L - 1)))
        offsets = level_offsets(L, c['base'], pls, c['log2'], c['align'])
        table = table_pattern(int(offsets[-1]), C)
        x, grad = case_inputs(c)
        B = x.shape[0]
        S = float(np.log2(pls))
        t = lambda a: torch.from_numpy(a).to(dev)
        xt, tt, ot = t(x), t(table), t(offsets)
        outputs = torch.empty(L, B, C, device=dev)
        dy_dx = torch.empty(B, L * D * C, device=dev)
        ge.grid_encode_forward(xt, tt, ot, outputs, B, D, C, L, S, c['base'], dy_dx, c['gridtype'], c['align'], c['interp'])
        gl = t(grad).view(B, L, C).permute(1, 0, 2).contiguous()
        g_emb = torch.zeros_like(tt)
        g_in = torch.zeros_like(xt)
        ge.grid_encode_backward(gl, xt, tt, ot, g_emb, B, D, C, L, S, c['base'], dy_dx, g_in, c['gridtype'], c['align'], c['interp'])
        torch.cuda.synchronize()
        ge_flat = g_emb.reshape(-1).cpu().numpy()
        nz = np.flatnonzero(ge_flat).astype(np.int64)
        out[f'{name}.x01'] = x
        out[f'{name}.grad'] = grad
        out[f'{name}.outputs_LBC'] = outputs.cpu().numpy()
        out[f'{name}.dy_dx'] = dy_dx.cpu().numpy()
        out[f'{name}.grad_inputs'] = g_in.cpu().numpy()
        out[f'{name}.grad_emb_idx'] = nz
        out[f'{name}.grad_emb_val'] = ge_flat[nz]
        out[f'{name}.offsets'] = offsets
        out[f'{name}.S'] = np.float64(S)
        # exp2f(level * S) * H - 1 as THIS GPU evaluates it (torch.exp2 -> the same libdevice exp2f the reference kernel calls,
        # gridencoder.cu:138); stored because a host libm exp2 differs from MUFU.EX2 by 1 ulp at some levels
        lv = torch.arange(L, dtype=torch.float32, device=dev)
        out[f'{name}.level_scale'] = (torch.exp2(lv * torch.tensor(S, dtype=torch.float32, device=dev)) * float(c['base']) - 1.0).cpu().numpy()
        print(name, 'B', B, 'rows', int(offsets[-1]), 'touched', nz.size, 'out abs mean', float(np.abs(out[f"{name}.outputs_LBC"]).mean()))
    dst = os.path.join(ROOT, 'gpurun_out', 'grid.npz')
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print('wrote', dst, os.path.getsize(dst), 'bytes')


if __name__ == '__main__':
    main()
