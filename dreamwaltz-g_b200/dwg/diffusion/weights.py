"""Random-init state dicts with the parameter names and shapes of the public diffusers modules
(UNet2DConditionModel, ControlNetModel, AutoencoderKL encoder) for the SD1.5 / SD2.1 configs.

No checkpoint exists in this environment (no network), so bench and tests run on synthetic
weights of the right architecture; a real checkpoint's state dict has the
same keys and can be passed to the same loaders."""
import torch

SD15 = dict(block_out=(320, 640, 1280, 1280), layers_per_block=2, heads=8, ctx_dim=768, in_ch=4, out_ch=4,
            cond_embed=(16, 32, 96, 256), groups=32)
# SD2.1: attention_head_dim (5, 10, 20, 20) = HEADS per level (head width 64), OpenCLIP context 1024,
# use_linear_projection=True (proj_in / proj_out of every Transformer2DModel are nn.Linear)
SD21 = dict(block_out=(320, 640, 1280, 1280), layers_per_block=2, heads=(5, 10, 20, 20), ctx_dim=1024, in_ch=4, out_ch=4,
            cond_embed=(16, 32, 96, 256), groups=32, linear_proj=True)
VAE15 = dict(block_out=(128, 256, 512, 512), layers_per_block=2, latent=4, groups=32, scaling_factor=0.18215)
TINY = dict(block_out=(64, 128, 256, 256), layers_per_block=2, heads=8, ctx_dim=96, in_ch=4, out_ch=4,
            cond_embed=(16, 32, 32, 64), groups=32)
TINY21 = dict(block_out=(64, 128, 256, 256), layers_per_block=2, heads=(1, 2, 4, 4), ctx_dim=128, in_ch=4, out_ch=4,
              cond_embed=(16, 32, 32, 64), groups=32, linear_proj=True)
TINY_VAE = dict(block_out=(32, 64, 64, 64), layers_per_block=2, latent=4, groups=32, scaling_factor=0.18215)


class _Init:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.sd = {}

    def conv(self, name, cin, cout, k=3, gain=1.0):
        fan = cin * k * k
        self.sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=self.g) * (gain / fan ** 0.5)
        self.sd[name + '.bias'] = torch.randn(cout, generator=self.g) * 0.02

    def lin(self, name, cin, cout, bias=True, gain=1.0):
        self.sd[name + '.weight'] = torch.randn(cout, cin, generator=self.g) * (gain / cin ** 0.5)
        if bias:
            self.sd[name + '.bias'] = torch.randn(cout, generator=self.g) * 0.02

    def norm(self, name, c):
        self.sd[name + '.weight'] = 1.0 + 0.1 * torch.randn(c, generator=self.g)
        self.sd[name + '.bias'] = 0.05 * torch.randn(c, generator=self.g)

    def resnet(self, p, cin, cout, temb_dim):
        self.norm(p + '.norm1', cin)
        self.conv(p + '.conv1', cin, cout)
        if temb_dim:
            self.lin(p + '.time_emb_proj', temb_dim, cout)
        self.norm(p + '.norm2', cout)
        self.conv(p + '.conv2', cout, cout, gain=0.5)
        if cin != cout:
            self.conv(p + '.conv_shortcut', cin, cout, k=1)

    def transformer(self, p, c, ctx_dim, linear_proj=False):
        self.norm(p + '.norm', c)
        if linear_proj:
            self.lin(p + '.proj_in', c, c)
        else:
            self.conv(p + '.proj_in', c, c, k=1)
        b = p + '.transformer_blocks.0'
        for n in ('norm1', 'norm2', 'norm3'):
            self.norm(f'{b}.{n}', c)
        for a, kd in (('attn1', c), ('attn2', ctx_dim)):
            self.lin(f'{b}.{a}.to_q', c, c, bias=False)
            self.lin(f'{b}.{a}.to_k', kd, c, bias=False)
            self.lin(f'{b}.{a}.to_v', kd, c, bias=False)
            self.lin(f'{b}.{a}.to_out.0', c, c, gain=0.5)
        self.lin(f'{b}.ff.net.0.proj', c, 8 * c)
        self.lin(f'{b}.ff.net.2', 4 * c, c, gain=0.5)
        if linear_proj:
            self.lin(p + '.proj_out', c, c, gain=0.5)
        else:
            self.conv(p + '.proj_out', c, c, k=1, gain=0.5)

    def encoder_path(self, cfg):
        bo, temb_dim = cfg['block_out'], cfg['block_out'][0] * 4
        self.conv('conv_in', cfg['in_ch'], bo[0])
        self.lin('time_embedding.linear_1', bo[0], temb_dim)
        self.lin('time_embedding.linear_2', temb_dim, temb_dim)
        cin = bo[0]
        for i, c in enumerate(bo):
            for j in range(cfg['layers_per_block']):
                self.resnet(f'down_blocks.{i}.resnets.{j}', cin, c, temb_dim)
                if i < len(bo) - 1:
                    self.transformer(f'down_blocks.{i}.attentions.{j}', c, cfg['ctx_dim'], cfg.get('linear_proj', False))
                cin = c
            if i < len(bo) - 1:
                self.conv(f'down_blocks.{i}.downsamplers.0.conv', c, c)
        self.resnet('mid_block.resnets.0', bo[-1], bo[-1], temb_dim)
        self.transformer('mid_block.attentions.0', bo[-1], cfg['ctx_dim'], cfg.get('linear_proj', False))
        self.resnet('mid_block.resnets.1', bo[-1], bo[-1], temb_dim)


def skip_channels(cfg):
    """Channel count of every skip tensor (conv_in output, each resnet/attention, each downsample)."""
    bo = cfg['block_out']
    ch = [bo[0]]
    for i, c in enumerate(bo):
        ch += [c] * cfg['layers_per_block']
        if i < len(bo) - 1:
            ch.append(c)
    return ch


def make_unet(cfg=SD15, seed=1):
    w = _Init(seed)
    w.encoder_path(cfg)
    bo, temb_dim = cfg['block_out'], cfg['block_out'][0] * 4
    skips = skip_channels(cfg)
    rev = list(reversed(bo))
    cin = bo[-1]
    for i, c in enumerate(rev):
        for j in range(cfg['layers_per_block'] + 1):
            sk = skips.pop()
            w.resnet(f'up_blocks.{i}.resnets.{j}', cin + sk, c, temb_dim)
            if i > 0:
                w.transformer(f'up_blocks.{i}.attentions.{j}', c, cfg['ctx_dim'], cfg.get('linear_proj', False))
            cin = c
        if i < len(bo) - 1:
            w.conv(f'up_blocks.{i}.upsamplers.0.conv', c, c)
    w.norm('conv_norm_out', bo[0])
    w.conv('conv_out', bo[0], cfg['out_ch'], gain=0.5)
    return w.sd


def make_controlnet(cfg=SD15, seed=2):
    w = _Init(seed)
    w.encoder_path(cfg)
    ce = cfg['cond_embed']
    w.conv('controlnet_cond_embedding.conv_in', 3, ce[0])
    for i in range(len(ce) - 1):
        w.conv(f'controlnet_cond_embedding.blocks.{2 * i}', ce[i], ce[i])
        w.conv(f'controlnet_cond_embedding.blocks.{2 * i + 1}', ce[i], ce[i + 1])
    w.conv('controlnet_cond_embedding.conv_out', ce[-1], cfg['block_out'][0], gain=0.3)
    for i, c in enumerate(skip_channels(cfg)):
        w.conv(f'controlnet_down_blocks.{i}', c, c, k=1, gain=0.3)
    w.conv('controlnet_mid_block', cfg['block_out'][-1], cfg['block_out'][-1], k=1, gain=0.3)
    return w.sd


def make_vae_encoder(cfg=VAE15, seed=3):
    w = _Init(seed)
    bo = cfg['block_out']
    w.conv('encoder.conv_in', 3, bo[0])
    cin = bo[0]
    for i, c in enumerate(bo):
        for j in range(cfg['layers_per_block']):
            w.resnet(f'encoder.down_blocks.{i}.resnets.{j}', cin, c, 0)
            cin = c
        if i < len(bo) - 1:
            w.conv(f'encoder.down_blocks.{i}.downsamplers.0.conv', c, c)
    w.resnet('encoder.mid_block.resnets.0', bo[-1], bo[-1], 0)
    p = 'encoder.mid_block.attentions.0'
    w.norm(p + '.group_norm', bo[-1])
    for n in ('to_q', 'to_k', 'to_v'):
        w.lin(f'{p}.{n}', bo[-1], bo[-1])
    w.lin(p + '.to_out.0', bo[-1], bo[-1], gain=0.5)
    w.resnet('encoder.mid_block.resnets.1', bo[-1], bo[-1], 0)
    w.norm('encoder.conv_norm_out', bo[-1])
    w.conv('encoder.conv_out', bo[-1], 2 * cfg['latent'])
    w.conv('quant_conv', 2 * cfg['latent'], 2 * cfg['latent'], k=1)
    return w.sd
