"""One un-graphed guidance pass (VAE encode -> ControlNet + UNet -> SDS gradient -> VAE backward) at the benchmark
shapes, bracketed by cudaProfilerStart/Stop so Nsight Compute (--profile-from-start off) sees exactly one step's
tensor-core launches:
    ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        -k regex:gemm_kernel\\|fa_fwd --csv --log-file gpurun_out/<tag>_gemm_traffic.csv python tools/gemm_pass.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg.diffusion import guidance as G, weights as W  # noqa: E402

DEV = 'cuda'
tiny = '--tiny' in sys.argv
cfg, vcfg = (W.TINY, W.TINY_VAE) if tiny else (W.SD15, W.VAE15)
g = G.ControlNetScoreDistillation(W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg), cfg, vcfg, DEV, seed=1)
g.two_streams = False                     # one stream: ncu serialises launches anyway
gen = torch.Generator().manual_seed(7)
S = 64 if tiny else 512
emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=gen).to(DEV), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=gen).to(DEV)}
cond = (torch.rand(1, 3, S, S, generator=gen) > 0.97).float().to(DEV)
for it in range(2):
    img = torch.rand(1, 3, S, S, device=DEV, requires_grad=True)
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    res = g(img, emb, cond_inputs=cond)
    res['diffusion_loss'].backward()
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.cudart().cudaProfilerStop()
print('done')
