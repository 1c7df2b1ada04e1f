"""Autotune the tcgen05 GEMM planner on the launches of one SDS guidance pass.

Records every GEMM / convolution call of one VAE + ControlNet + UNet forward and VAE backward, then, for
every distinct planner key, sweeps the tile width BN and the split-K factor with COLD weights (each launch of
the timed CUDA graph uses its own copy of the weight operand, the L2 is flushed before the replay) and WARM
activations (touched inside the graph right before the launches) -- the situation of the real step -- and
writes the winners to dreamwaltz-g_b200/csrc/gemm_plan_table.inc.

    python tools/gemm_autotune.py [--tiny] [--out path]
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402
from dwg._lib import lib  # noqa: E402
from dwg.diffusion import guidance as G, weights as W  # noqa: E402

DEV = 'cuda'
NCOPY = 8
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def record_calls(tiny):
    cfg, vcfg = (W.TINY, W.TINY_VAE) if tiny else (W.SD15, W.VAE15)
    g = G.ControlNetScoreDistillation(W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg), cfg, vcfg, DEV, seed=1)
    gen = torch.Generator().manual_seed(7)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=gen).to(DEV), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=gen).to(DEV)}
    S = 64 if tiny else 512
    cond = (torch.rand(1, 3, S, S, generator=gen) > 0.97).float().to(DEV)
    img = torch.rand(1, 3, S, S, device=DEV, requires_grad=True)
    ops.TUNE_RECORD = []
    res = g(img, emb, cond_inputs=cond)
    res['diffusion_loss'].backward()
    torch.cuda.synchronize()
    rec, ops.TUNE_RECORD = ops.TUNE_RECORD, None
    del g
    torch.cuda.empty_cache()
    return rec


def make_problem(call):
    """-> (fn(copy_index), touch()) for one recorded call, with NCOPY copies of the weight operand."""
    torch.manual_seed(0)
    if call[0] == 'gemm':
        _, M, N, K, nb1, nb2, act, has_res, cdt, has_bias, has_b2 = call
        a = torch.randn(nb2, nb1, M, K, device=DEV).half()
        bs = [torch.randn(nb2, nb1, N, K, device=DEV).half() for _ in range(NCOPY)]
        No = N // 2 if act == 'geglu' else N
        r = torch.randn(nb2, nb1, M, No, device=DEV).half() if has_res else None
        bias = torch.randn(N, device=DEV) if has_bias else None
        b2 = torch.randn(M, N, device=DEV) if has_b2 else None
        fn = lambda i: ops.gemm(a, bs[i % NCOPY], bias=bias, bias2=b2, bias2_rows_per=1 if has_b2 else 0, residual=r, act=act, out_dtype=cdt)
        touch = [a] + ([r] if r is not None else [])
    else:
        _, Ni, H, Wd, Ci, Co, k, stride, ph, pw, Ho, Wo, has_res, odt, has_b2 = call
        x = torch.randn(Ni, H, Wd, Ci, device=DEV).half()
        ws = [torch.randn(Co, k, k, Ci, device=DEV).half() for _ in range(NCOPY)]
        r = torch.randn(Ni, Ho, Wo, Co, device=DEV).half() if has_res else None
        bias = torch.randn(Co, device=DEV)
        b2 = torch.randn(Ni, Co, device=DEV) if has_b2 else None
        fn = lambda i: ops.conv2d_nhwc(x, ws[i % NCOPY], bias=bias, bias2=b2, residual=r, stride=stride, padding=(ph, pw), out_hw=(Ho, Wo), out_dtype=odt)
        touch = [x] + ([r] if r is not None else [])
    return fn, touch


def time_config(fn, touch, iters=NCOPY, reps=3):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for t in touch:
                t.add_(0)                       # activations are L2-warm in the real step
            for i in range(iters):
                fn(i)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / iters)
    return best


def main():
    tiny = '--tiny' in sys.argv
    out = sys.argv[sys.argv.index('--out') + 1] if '--out' in sys.argv else os.path.join(ROOT, 'dreamwaltz-g_b200', 'csrc', 'gemm_plan_table.inc')
    L = lib()
    rec = record_calls(tiny)
    plan, key = (ctypes.c_int * 3)(), (ctypes.c_int * 6)()
    seen, rows = {}, []
    tot_auto = tot_best = 0.0
    def force(bn, ks, pair, halo):
        L.dwg_gemm_tune(bn, ks)
        L.dwg_gemm_tune_pair(pair)
        L.dwg_gemm_tune_halo(halo, 0)

    def taken():
        L.dwg_gemm_last_plan(plan)
        return (plan[0], plan[1], L.dwg_gemm_last_pair(), L.dwg_gemm_last_halo())

    import time
    t_begin, budget = time.time(), float(os.environ.get('DWG_TUNE_BUDGET_S', '1e9'))
    for call in rec:
        if time.time() - t_begin > budget:
            print('time budget reached: remaining keys keep the analytic plan', flush=True)
            break
        fn, touch = make_problem(call)
        force(-1, 0, -1, -1)                    # -1: analytic model only (ignore the table)
        fn(0)
        L.dwg_gemm_last_key(key)
        k = tuple(key)
        if k in seen:
            seen[k][0] += 1
            continue
        auto = taken()
        t_auto = time_config(fn, touch)
        geglu = k[4] == 2
        conv3 = call[0] == 'conv' and call[6] == 3 and call[7] == 1
        res = {auto: t_auto}
        bns = (64, 128, 192, 256) if geglu else (32, 64, 96, 128, 160, 192, 224, 256)
        kss = (1,) if geglu else (1, 2, 3, 4, 6, 8, 12, 16)
        cands = [(bn, ks, 0, 0) for bn in bns for ks in kss]
        cands += [(bn, ks, 1, 0) for bn in bns for ks in kss if ks <= 6 and k[0] % 2 == 0]
        if conv3:
            cands += [(bn, 1, pr, 1) for bn in bns for pr in (0, 1)]
        for bn, ks, pr, hl in cands:
            if ks > 1:
                n_tiles = (k[2] + bn - 1) // bn
                if k[0] * k[1] * n_tiles * ks > 2 * 148:
                    continue
            force(bn, ks, pr, hl)
            fn(0)
            got = taken()
            if got != (bn, ks, pr, hl) or got in res:
                continue
            res[got] = time_config(fn, touch)
        force(0, 0, -1, -1)
        ranked = sorted((t, c) for c, t in res.items())
        best = ranked[0]
        # keep the model's choice unless the measured winner is clearly (3%) better
        if best[0] > 0.97 * t_auto:
            best = (t_auto, auto)
        seen[k] = [1, t_auto, best]
        print(f'{str(call[:12]):70s} key={k} auto {t_auto:7.1f} us {auto} -> best {best[0]:7.1f} us {best[1]}', flush=True)
    with open(out, 'w') as fh:
        fh.write('// generated by tools/gemm_autotune.py on a 148-SM B200: {m_tiles, nz, N, k_iters, epi, has_res, BN, ksplit, pair, halo}\n')
        for k, (n, t_auto, best) in sorted(seen.items()):
            bn, ks, pr, hl = best[1]
            fh.write(f'    {{{k[0]}, {k[1]}, {k[2]}, {k[3]}, {k[4]}, {k[5]}, {bn}, {ks}, {pr}, {hl}}},   // x{n}: {t_auto:.1f} -> {best[0]:.1f} us\n')
            tot_auto += n * t_auto
            tot_best += n * best[0]
    print(f'{len(seen)} keys; per step: model {tot_auto / 1e3:.3f} ms -> tuned {tot_best / 1e3:.3f} ms; wrote {out}')


if __name__ == '__main__':
    main()
