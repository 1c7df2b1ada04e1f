"""Oracle: Stable-Diffusion UNet / ControlNet / VAE-encoder forward and the SDS arithmetic
(rows R14-R16), plain torch fp32 (CPU, or any device the state dict / inputs live on), NCHW, functional over a diffusers-style state dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: ``diffusers`` is an un-vendored
third-party dependency of the reference (requirements.txt:4 pins 0.24.0, scripts/install.sh:27
installs git HEAD) and is not installed here; this restates the published architectures
(UNet2DConditionModel, ControlNetModel, AutoencoderKL encoder) anchored on
the reference's call sites core/guidance/controlnet.py:83-114, core/guidance/vae.py:34-40,
core/guidance/basic.py:354-383,546-663,778-917.  Module / parameter names are diffusers' so that a
real checkpoint's state dict can be fed to both this oracle and the CUDA implementation.
"""
import math

import torch
import torch.nn.functional as F

SD15 = dict(block_out=(320, 640, 1280, 1280), layers_per_block=2, heads=8, ctx_dim=768, in_ch=4, out_ch=4,
            cond_embed=(16, 32, 96, 256), groups=32)
VAE15 = dict(block_out=(128, 256, 512, 512), layers_per_block=2, latent=4, groups=32, scaling_factor=0.18215)


# ------------------------------------------------------------------------------------ blocks
def timestep_embedding(t, dim):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + '.weight'], sd.get(name + '.bias'), stride=stride, padding=padding)


def _lin(sd, name, x):
    return F.linear(x, sd[name + '.weight'], sd.get(name + '.bias'))


def _gn(sd, name, x, groups, eps):
    return F.group_norm(x, groups, sd[name + '.weight'], sd[name + '.bias'], eps)


def resnet(sd, p, x, temb, groups, eps):
    h = _conv(sd, p + '.conv1', F.silu(_gn(sd, p + '.norm1', x, groups, eps)))
    if temb is not None:
        h = h + _lin(sd, p + '.time_emb_proj', F.silu(temb))[:, :, None, None]
    h = _conv(sd, p + '.conv2', F.silu(_gn(sd, p + '.norm2', h, groups, eps)))
    if (p + '.conv_shortcut.weight') in sd:
        x = _conv(sd, p + '.conv_shortcut', x, padding=0)
    return x + h


def attention(sd, p, x, ctx, heads):
    q, k, v = _lin(sd, p + '.to_q', x), _lin(sd, p + '.to_k', ctx), _lin(sd, p + '.to_v', ctx)
    B, T, C = q.shape
    hd = C // heads
    sp = lambda t: t.view(B, -1, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))
    return _lin(sd, p + '.to_out.0', o.transpose(1, 2).reshape(B, T, C))


def _heads_at(cfg, level):
    """attention_head_dim of the diffusers configs: an int (SD1.5: 8 heads everywhere) or per-level heads (SD2.1)."""
    h = cfg['heads']
    return h if isinstance(h, int) else h[level]


def transformer(sd, p, x, ctx, heads, groups):
    B, C, H, W = x.shape
    res = x
    n0 = _gn(sd, p + '.norm', x, groups, 1e-6)
    lin_proj = sd[p + '.proj_in.weight'].dim() == 2            # use_linear_projection (SD2.1): Linear after the reshape
    if lin_proj:
        h = _lin(sd, p + '.proj_in', n0.permute(0, 2, 3, 1).reshape(B, H * W, C))
    else:
        h = _conv(sd, p + '.proj_in', n0, padding=0).permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = p + '.transformer_blocks.0'
    n = F.layer_norm(h, (C,), sd[b + '.norm1.weight'], sd[b + '.norm1.bias'])
    h = h + attention(sd, b + '.attn1', n, n, heads)
    n = F.layer_norm(h, (C,), sd[b + '.norm2.weight'], sd[b + '.norm2.bias'])
    h = h + attention(sd, b + '.attn2', n, ctx, heads)
    n = F.layer_norm(h, (C,), sd[b + '.norm3.weight'], sd[b + '.norm3.bias'])
    g = _lin(sd, b + '.ff.net.0.proj', n)
    a, gate = g.chunk(2, dim=-1)
    h = h + _lin(sd, b + '.ff.net.2', a * F.gelu(gate))
    if lin_proj:
        return _lin(sd, p + '.proj_out', h).reshape(B, H, W, C).permute(0, 3, 1, 2) + res
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _conv(sd, p + '.proj_out', h, padding=0) + res


def _time_embed(sd, t, cfg, B):
    temb = timestep_embedding(t.reshape(-1).expand(B), cfg['block_out'][0]).to(sd['time_embedding.linear_1.weight'].dtype)
    return _lin(sd, 'time_embedding.linear_2', F.silu(_lin(sd, 'time_embedding.linear_1', temb)))


def _down_path(sd, cfg, h, temb, ctx):
    """conv_in output + down blocks -> (h, skips)."""
    G, nb = cfg['groups'], len(cfg['block_out'])
    skips = [h]
    for i in range(nb):
        has_attn = i < nb - 1
        for j in range(cfg['layers_per_block']):
            h = resnet(sd, f'down_blocks.{i}.resnets.{j}', h, temb, G, 1e-5)
            if has_attn:
                h = transformer(sd, f'down_blocks.{i}.attentions.{j}', h, ctx, _heads_at(cfg, i), G)
            skips.append(h)
        if i < nb - 1:
            h = _conv(sd, f'down_blocks.{i}.downsamplers.0.conv', h, stride=2, padding=1)
            skips.append(h)
    return h, skips


def _mid(sd, cfg, h, temb, ctx):
    G = cfg['groups']
    h = resnet(sd, 'mid_block.resnets.0', h, temb, G, 1e-5)
    h = transformer(sd, 'mid_block.attentions.0', h, ctx, _heads_at(cfg, len(cfg['block_out']) - 1), G)
    return resnet(sd, 'mid_block.resnets.1', h, temb, G, 1e-5)


def controlnet_forward(sd, cfg, sample, t, ctx, cond, conditioning_scale=1.0):
    """ControlNetModel.forward -> (12 down residuals, mid residual).  cond [B,3,8h,8w] in [0,1]."""
    B = sample.shape[0]
    temb = _time_embed(sd, t, cfg, B)
    h = _conv(sd, 'conv_in', sample)
    c = F.silu(_conv(sd, 'controlnet_cond_embedding.conv_in', cond))
    nblk = 2 * (len(cfg['cond_embed']) - 1)
    for k in range(nblk):
        c = F.silu(_conv(sd, f'controlnet_cond_embedding.blocks.{k}', c, stride=2 if k % 2 == 1 else 1))
    c = _conv(sd, 'controlnet_cond_embedding.conv_out', c)
    h = h + c
    h, skips = _down_path(sd, cfg, h, temb, ctx)
    h = _mid(sd, cfg, h, temb, ctx)
    down = [_conv(sd, f'controlnet_down_blocks.{i}', s, padding=0) * conditioning_scale for i, s in enumerate(skips)]
    mid = _conv(sd, 'controlnet_mid_block', h, padding=0) * conditioning_scale
    return down, mid


def unet_forward(sd, cfg, sample, t, ctx, down_residuals=None, mid_residual=None):
    """UNet2DConditionModel.forward (timestep tensor of shape [1] broadcast over the batch,
    as the reference passes it: controlnet.py:100,109)."""
    B = sample.shape[0]
    G, nb = cfg['groups'], len(cfg['block_out'])
    temb = _time_embed(sd, t, cfg, B)
    h = _conv(sd, 'conv_in', sample)
    h, skips = _down_path(sd, cfg, h, temb, ctx)
    if down_residuals is not None:
        skips = [s + r for s, r in zip(skips, down_residuals)]
    h = _mid(sd, cfg, h, temb, ctx)
    if mid_residual is not None:
        h = h + mid_residual
    for i in range(nb):
        has_attn = i > 0
        for j in range(cfg['layers_per_block'] + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(sd, f'up_blocks.{i}.resnets.{j}', h, temb, G, 1e-5)
            if has_attn:
                h = transformer(sd, f'up_blocks.{i}.attentions.{j}', h, ctx, _heads_at(cfg, nb - 1 - i), G)
        if i < nb - 1:
            h = F.interpolate(h, scale_factor=2.0, mode='nearest')
            h = _conv(sd, f'up_blocks.{i}.upsamplers.0.conv', h)
    h = F.silu(_gn(sd, 'conv_norm_out', h, G, 1e-5))
    return _conv(sd, 'conv_out', h)


def vae_encode_moments(sd, cfg, x):
    """AutoencoderKL.encode: encoder + quant_conv -> (mean, logvar).  x [B,3,H,W] in [-1,1]."""
    G, nb = cfg['groups'], len(cfg['block_out'])
    h = _conv(sd, 'encoder.conv_in', x)
    for i in range(nb):
        for j in range(cfg['layers_per_block']):
            h = resnet(sd, f'encoder.down_blocks.{i}.resnets.{j}', h, None, G, 1e-6)
        if i < nb - 1:
            h = F.pad(h, (0, 1, 0, 1))
            h = _conv(sd, f'encoder.down_blocks.{i}.downsamplers.0.conv', h, stride=2, padding=0)
    h = resnet(sd, 'encoder.mid_block.resnets.0', h, None, G, 1e-6)
    B, C, H, W = h.shape
    p = 'encoder.mid_block.attentions.0'
    n = _gn(sd, p + '.group_norm', h, G, 1e-6).permute(0, 2, 3, 1).reshape(B, H * W, C)
    q, k, v = _lin(sd, p + '.to_q', n), _lin(sd, p + '.to_k', n), _lin(sd, p + '.to_v', n)
    a = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1) @ v
    h = h + _lin(sd, p + '.to_out.0', a).reshape(B, H, W, C).permute(0, 3, 1, 2)
    h = resnet(sd, 'encoder.mid_block.resnets.1', h, None, G, 1e-6)
    h = _conv(sd, 'encoder.conv_out', F.silu(_gn(sd, 'encoder.conv_norm_out', h, G, 1e-6)))
    m = _conv(sd, 'quant_conv', h, padding=0)
    return m.chunk(2, dim=1)


def vae_encode_latents(sd, cfg, images01, eps):
    """AutoEncoderSD.encode_images (vae.py:34-40): normalise to [-1,1], sample with the given
    standard-normal eps, scale by 0.18215."""
    mean, logvar = vae_encode_moments(sd, cfg, 2.0 * images01 - 1.0)
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return (mean + std * eps) * cfg['scaling_factor']


def alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """scaled_linear schedule of the SD schedulers (DDPMScheduler)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(latents, noise, t, acp=None):
    acp = alphas_cumprod().to(latents.device) if acp is None else acp
    a = acp[t].reshape(-1, 1, 1, 1)
    return a.sqrt() * latents + (1 - a).sqrt() * noise


def sds_gradient(unet_sd, cn_sd, cfg, latents_noisy, noise, t, emb_uncond, emb_text, cond_image01, guidance_scale=50.0,
                 conditioning_scale=1.0):
    """calc_gradients (basic.py:546-663) for loss_type 'sds', weight 'sjc' (= 1), CFG on, with
    ControlNetScoreDistillation._predict (controlnet.py:83-114).  Returns (gradients, noise_pred)."""
    ctx = torch.cat([emb_uncond, emb_text], dim=0)
    x2 = torch.cat([latents_noisy] * 2, dim=0)
    cond2 = cond_image01.repeat_interleave(2 // cond_image01.shape[0], dim=0) if cond_image01.shape[0] == 1 else cond_image01
    down, mid = controlnet_forward(cn_sd, cfg, x2, t, ctx, cond2, conditioning_scale)
    eps = unet_forward(unet_sd, cfg, x2, t, ctx, down, mid)
    e_u, e_c = eps.chunk(2)
    noise_pred = e_u + guidance_scale * (e_c - e_u)
    return noise_pred - noise, noise_pred
