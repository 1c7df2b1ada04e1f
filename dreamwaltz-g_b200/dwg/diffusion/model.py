"""UNet2DConditionModel / ControlNetModel / AutoencoderKL-encoder on the dwg tcgen05 kernels.

Activations are NHWC fp16 end to end ([B,H,W,C] == token layout [B,HW,C] for free); every conv /
linear / attention matmul is dwg_gemm_f16 or dwg_conv2d_nhwc_f16 (implicit GEMM, TMA-fed
tcgen05, fp32 accumulation in TMEM) with bias / time-embedding / residual fused into the
epilogue; GroupNorm(+SiLU), LayerNorm, softmax, GEGLU are the streaming kernels of nn_kernels.cu.
torch is used for allocation and pure data movement (cat, pad, nearest upsample, transposes).

The public architectures (diffusers; reference call sites core/guidance/controlnet.py:83-114,
core/guidance/vae.py:34-40) are consumed as diffusers-format state dicts (weights.py).
"""
import math
import os

import torch
import torch.nn.functional as F

from .. import ops

F16 = torch.float16


def _pad_c(t, mult=8, dim=-1):
    c = t.shape[dim]
    pad = (-c) % mult
    if pad == 0:
        return t
    shape = list(t.shape)
    shape[dim] = pad
    return torch.cat([t, torch.zeros(shape, dtype=t.dtype, device=t.device)], dim=dim)


class Weights:
    """Device-side, layout-converted copy of a diffusers state dict.
    conv  [Cout,Cin,k,k] fp32 -> [Cout,k,k,Cin8] fp16 (Cin zero-padded to a multiple of 8)
    linear [out,in] -> fp16;  biases / norm affine -> fp32."""

    def __init__(self, sd, device='cuda', with_dgrad=False):
        self.w, self.b, self.dg = {}, {}, {}
        for k, v in sd.items():
            v = v.detach().to(device=device, dtype=torch.float32)
            if k.endswith('.weight'):
                name = k[:-7]
                if v.dim() == 4:
                    self.w[name] = _pad_c(v.permute(0, 2, 3, 1)).to(F16).contiguous()
                    if with_dgrad:
                        # dgrad weights: [Cin, k, k, Cout8] with the taps flipped
                        self.dg[name] = _pad_c(v.flip(2, 3).permute(1, 2, 3, 0)).to(F16).contiguous()
                elif v.dim() == 2:
                    if name.endswith('.ff.net.0.proj'):
                        # GEGLU projection: interleave (value_i, gate_i) rows so the GEMM epilogue can fuse hidden * gelu(gate)
                        inner = v.shape[0] // 2
                        v = torch.stack([v[:inner], v[inner:]], dim=1).reshape(2 * inner, v.shape[1])
                    self.w[name] = v.to(F16).contiguous()
                    if with_dgrad:
                        self.dg[name] = v.t().to(F16).contiguous()
                else:
                    self.w[name] = v.contiguous()          # norm scale (fp32)
            elif k.endswith('.bias'):
                if k.endswith('.ff.net.0.proj.bias'):
                    inner = v.shape[0] // 2
                    v = torch.stack([v[:inner], v[inner:]], dim=1).reshape(2 * inner)
                self.b[k[:-5]] = v.contiguous()

    def has(self, name):
        return name in self.w


STATS_IN_EPILOGUE = os.environ.get('DWG_NO_EPILOGUE_STATS') != '1'      # GroupNorm statistics from the producing epilogue (A/B switch)
CAT_FREE = os.environ.get('DWG_NO_CATFREE') != '1'                      # decoder skip concatenations never materialised (A/B switch)
# LayerNorms folded into the consuming GEMM epilogues (ops.gemm_ln).  Parity-neutral and 69 fewer launches per step, but MEASURED
# 0.1 ms SLOWER than the three LayerNorm kernels per block at the benchmark's size (13.88 vs 13.77 ms/step, same box, DESIGN.md
# section 6): off by default, DWG_LN_FOLD=1 enables it.
LN_FOLD = os.environ.get('DWG_LN_FOLD') == '1'


def conv(W, name, x, stride=1, padding=1, bias2=None, residual=None, out_dtype=F16, out_hw=None, use_bias=True, stats=False, act=None):
    """stats=True: the output feeds a GroupNorm -- its epilogue accumulates that layer's statistics (y._cs)."""
    w = W.w[name]
    if x.shape[-1] != w.shape[-1]:
        x = _pad_c(x)
    return ops.conv2d_nhwc(x, w, bias=W.b.get(name) if use_bias else None, bias2=bias2, residual=residual, stride=stride,
                           padding=padding, out_hw=out_hw, out_dtype=out_dtype, stats=stats and STATS_IN_EPILOGUE, act=act)


def linear(W, name, x2d, residual=None, act=None, out_dtype=F16, alpha=1.0, stats_rows=None):
    return ops.gemm(x2d, W.w[name], bias=W.b.get(name), residual=residual, act=act, out_dtype=out_dtype, alpha=alpha,
                    colstats_rows=stats_rows if STATS_IN_EPILOGUE else None)


def _view_keep_stats(t, *shape):
    v = t.view(*shape)
    if getattr(t, '_cs', None) is not None:
        v._cs = t._cs
    return v


def gn(W, name, x, groups, eps, silu, return_stats=False):
    return ops.group_norm(x, W.w[name], W.b[name], groups=groups, eps=eps, silu=silu, return_stats=return_stats,
                          colstats=getattr(x, '_cs', None))


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def attention(W, p, x, ctx, heads, residual, kv=None, helper=None):
    """x [B,T,C] fp16 (queries), ctx [B,Tk,Ck] fp16.  Returns to_out(softmax(QK^T/sqrt(d)) V) + residual.
    V is produced transposed ([B,C,Tk], the K-major operand of the PV matmul) by swapping the
    operands of its projection GEMM, so no transpose pass exists.  ``kv`` = precomputed (K, V^T)
    slices of the batched cross-attention projection (DiffusionNet.project_context)."""
    B, T, C = x.shape
    Tk = ctx.shape[1]
    hd = C // heads
    Tkp = (Tk + 7) // 8 * 8
    def v_transposed():
        vT = torch.zeros(B, C, Tkp, device=x.device, dtype=F16) if Tkp != Tk else torch.empty(B, C, Tkp, device=x.device, dtype=F16)
        wv = W.w[p + '.to_v']
        # ONE batched launch: the weight matrix is the (broadcast, batch stride 0) A operand, V^T[b] = Wv ctx[b]^T
        ops.gemm(wv.unsqueeze(0).expand(B, -1, -1), ctx, out=vT[:, :, :Tk] if Tkp != Tk else vT)
        if (p + '.to_v') in W.b:
            vT += W.b[p + '.to_v'].to(F16)[None, :, None]
        return vT
    k3 = None
    if kv is None and ctx is x and (p + '.to_qk') in W.w:
        # self-attention.  V^T is independent of the Q | K projection: it is forked onto the helper stream first (when there is
        # one), then Q and K projections of the same input run as ONE GEMM (weights concatenated at load time); the attention
        # kernel takes the two halves as strided views
        with ops.forked(helper) as fk:
            vT = v_transposed()
        qk = ops.gemm(x.reshape(B * T, C), W.w[p + '.to_qk']).view(B, T, 2 * C)
        q3, k3 = qk[:, :, :C], qk[:, :, C:]
        fk.join(vT)
    else:
        q3 = ops.gemm(x.reshape(B * T, C), W.w[p + '.to_q'], bias=W.b.get(p + '.to_q')).view(B, T, C)
        if kv is not None:
            k3, vT = kv
        else:
            k3 = ops.gemm(ctx.reshape(B * Tk, -1), W.w[p + '.to_k'], bias=W.b.get(p + '.to_k')).view(B, Tk, C)
            vT = v_transposed()
    o = attn_core(q3, k3, vT, heads, Tk)
    return linear(W, p + '.to_out.0', o.reshape(B * T, C), residual=residual.reshape(B * T, C)).view(B, T, C)


def attn_core(q3, k3, vT, heads, Tk):
    """softmax(Q K^T / sqrt(d)) V for q3 [B,T,C], k3 [B,Tk,C] (strided views allowed), vT [B,C,Tkp] -> [B,T,C]."""
    B, T, C = q3.shape
    hd = C // heads
    Tkp = vT.shape[2]
    if hd <= 128:
        return ops.attention(q3, k3, vT, heads, Tk)                    # fused: scores never reach HBM
    q, k = q3.unflatten(2, (heads, hd)).permute(0, 2, 1, 3), k3.unflatten(2, (heads, hd)).permute(0, 2, 1, 3)
    S = torch.empty(B, heads, T, Tkp, device=q3.device, dtype=F16)
    ops.gemm(q, k, alpha=hd ** -0.5, out=S[..., :Tk] if Tkp != Tk else S)
    ops.softmax_rows_(S, Tk)
    o = torch.empty(B, T, C, device=q3.device, dtype=F16)
    ops.gemm(S, vT.unflatten(1, (heads, hd)), out=o.view(B, T, heads, hd).permute(0, 2, 1, 3))
    return o


class DiffusionNet:
    """Shared machinery of the UNet and the ControlNet (same encoder path)."""

    def __init__(self, sd, cfg, device='cuda'):
        self.cfg, self.dev = cfg, device
        self.W = Weights(sd, device)
        self.G = cfg['groups']
        self._ctx_kv = None
        self._sc_split = {}
        self.helper = None                      # (stream, split-K lane) for launches forked beside the main chain (set by the guidance object)
        # all time_emb_proj layers batched into ONE GEMM per step
        names = sorted(n[:-len('.time_emb_proj')] for n in self.W.w if n.endswith('.time_emb_proj'))
        self._tproj_names = names
        self._tproj_w = torch.cat([self.W.w[n + '.time_emb_proj'] for n in names], dim=0).contiguous()
        self._tproj_b = torch.cat([self.W.b[n + '.time_emb_proj'] for n in names], dim=0).contiguous()
        offs, o = {}, 0
        for n in names:
            c = self.W.w[n + '.time_emb_proj'].shape[0]
            offs[n] = (o, o + c)
            o += c
        self._tproj_off = offs
        # all cross-attention K / V projections of the text context batched into 1 + B GEMMs per step
        xn = sorted(n[:-len('.to_k')] for n in self.W.w if n.endswith('.attn2.to_k'))
        self._xattn_wk = torch.cat([self.W.w[n + '.to_k'] for n in xn], dim=0).contiguous()
        self._xattn_wv = torch.cat([self.W.w[n + '.to_v'] for n in xn], dim=0).contiguous()
        offs, o = {}, 0
        for n in xn:
            c = self.W.w[n + '.to_k'].shape[0]
            offs[n] = (o, o + c)
            o += c
        self._xattn_off, self._xattn_total = offs, o
        # LayerNorm folding (ops.gemm_ln): per transformer block, gamma-scaled weights W' = W * gamma (fp16), c1 = row sums of W'
        # (of the ROUNDED weights: exactly what the tensor cores multiply), c2 = W beta (+ bias)
        self._lnf = {}
        if LN_FOLD:
            f32 = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)

            def fold(wname, gamma, beta, bias=None, interleave=False):
                w = f32(wname)
                bv = f32(bias) if bias is not None and bias in sd else None
                if interleave:                                            # GEGLU: (value_i, gate_i) rows interleaved, as Weights does
                    inner = w.shape[0] // 2
                    w = torch.stack([w[:inner], w[inner:]], dim=1).reshape(2 * inner, w.shape[1])
                    if bv is not None:
                        bv = torch.stack([bv[:inner], bv[inner:]], dim=1).reshape(2 * inner)
                wp = (w * gamma[None, :]).to(F16).contiguous()
                c1 = wp.float().sum(1).contiguous()
                c2 = w @ beta
                if bv is not None:
                    c2 = c2 + bv
                return wp, c1, c2.contiguous()
            for blk in sorted(n[:-len('.norm1.weight')] for n in sd if n.endswith('.transformer_blocks.0.norm1.weight')):
                g1, b1 = f32(blk + '.norm1.weight'), f32(blk + '.norm1.bias')
                g2, b2 = f32(blk + '.norm2.weight'), f32(blk + '.norm2.bias')
                g3, b3 = f32(blk + '.norm3.weight'), f32(blk + '.norm3.bias')
                wq, c1q, c2q = fold(blk + '.attn1.to_q.weight', g1, b1, blk + '.attn1.to_q.bias')
                wk, c1k, c2k = fold(blk + '.attn1.to_k.weight', g1, b1, blk + '.attn1.to_k.bias')
                self._lnf[blk] = {
                    'qk': (torch.cat([wq, wk], 0).contiguous(), torch.cat([c1q, c1k]).contiguous(), torch.cat([c2q, c2k]).contiguous()),
                    'v': fold(blk + '.attn1.to_v.weight', g1, b1, blk + '.attn1.to_v.bias'),
                    'q2': fold(blk + '.attn2.to_q.weight', g2, b2, blk + '.attn2.to_q.bias'),
                    'ff': fold(blk + '.ff.net.0.proj.weight', g3, b3, blk + '.ff.net.0.proj.bias', interleave=True),
                }
        # self-attention Q | K projection weights concatenated (bias-free in the public SD architectures)
        for n in sorted(n[:-len('.to_q')] for n in self.W.w if n.endswith('.attn1.to_q')):
            if (n + '.to_q') not in self.W.b and (n + '.to_k') not in self.W.b and os.environ.get('DWG_NO_QK') != '1':
                self.W.w[n + '.to_qk'] = torch.cat([self.W.w[n + '.to_q'], self.W.w[n + '.to_k']], dim=0).contiguous()

    def project_context(self, ctx):
        """ctx [B,Tk,Ck] -> {attn2 prefix: (K [B,Tk,C] view, V^T [B,C,Tkp] view)}."""
        B, Tk, Ck = ctx.shape
        Tkp = (Tk + 7) // 8 * 8
        S = self._xattn_total
        K_all = ops.gemm(ctx.reshape(B * Tk, Ck), self._xattn_wk).view(B, Tk, S)
        vT_all = torch.zeros(B, S, Tkp, device=ctx.device, dtype=F16)
        ops.gemm(self._xattn_wv.unsqueeze(0).expand(B, -1, -1), ctx, out=vT_all[:, :, :Tk])
        return {n: (K_all[:, :, a:c], vT_all[:, a:c, :]) for n, (a, c) in self._xattn_off.items()}

    def heads_at(self, level):
        h = self.cfg['heads']
        return h if isinstance(h, int) else h[level]

    def time_embed(self, t, B):
        W = self.W
        # TMA boxes that are mostly out of bounds are slow: run the M = B (= 2) GEMMs on a zero-padded
        # 128-row operand and slice the two valid rows afterwards
        e = torch.zeros(128, self.cfg['block_out'][0], device=t.device, dtype=F16)
        e[:B] = timestep_embedding(t.reshape(-1).expand(B), self.cfg['block_out'][0]).to(F16)
        e = linear(W, 'time_embedding.linear_1', e, act='silu')
        temb = linear(W, 'time_embedding.linear_2', e)
        tp = ops.gemm(ops.silu(temb), self._tproj_w, bias=self._tproj_b, out_dtype=torch.float32)[:B]      # [B, sum Cout]
        return {n: tp[:, a:b].contiguous() for n, (a, b) in self._tproj_off.items()}

    @torch.no_grad()
    def prepare(self, t, ctx, B):
        """Everything that depends only on the timestep and the prompt embeddings (time-embedding projections of
        every resnet, K / V^T of every cross-attention): can be computed while the avatar is still being rendered."""
        ctx = ctx.to(F16).contiguous()
        return {'tproj': self.time_embed(t, B), 'ctx': ctx, 'ctx_kv': self.project_context(ctx)}

    def resnet(self, p, x, tproj, x2=None):
        """x2 (decoder): the ResNet input is the channel concatenation [x | x2] (skip connection), which is never built: norm1 reads
        both tensors (ops.group_norm_cat) and the 1x1 shortcut is the sum of two convolutions over the two channel ranges."""
        W, G = self.W, self.G
        if x2 is not None and not (CAT_FREE and STATS_IN_EPILOGUE and getattr(x, '_cs', None) is not None and getattr(x2, '_cs', None) is not None
                                   and W.has(p + '.conv_shortcut')):
            x, x2 = ops.cat_channels(x, x2), None
        sc, fk = x, None
        if W.has(p + '.conv_shortcut'):
            # the 1x1 shortcut only meets the main chain again in conv2's epilogue: forked onto the helper stream (if any)
            with ops.forked(self.helper) as fk:
                if x2 is None:
                    sc = conv(W, p + '.conv_shortcut', x, padding=0)
                else:
                    wa, wb = self._shortcut_halves(p + '.conv_shortcut', x.shape[-1])
                    sc = ops.conv2d_nhwc(x2, wb, padding=0)
                    sc = ops.conv2d_nhwc(x, wa, bias=W.b.get(p + '.conv_shortcut'), residual=sc, padding=0)
        if x2 is None:
            h = gn(W, p + '.norm1', x, G, 1e-5, True)
        else:
            h = ops.group_norm_cat(x, x2, W.w[p + '.norm1'], W.b[p + '.norm1'], G, 1e-5, True)
        h = conv(W, p + '.conv1', h, bias2=tproj.get(p) if tproj else None, stats=True)
        h = gn(W, p + '.norm2', h, G, 1e-5, True)
        if fk is not None:
            fk.join(sc)
        return conv(W, p + '.conv2', h, residual=sc, stats=True)

    def transformer(self, p, x, ctx, heads):
        W = self.W
        B, H, Wd, C = x.shape
        h = gn(W, p + '.norm', x, self.G, 1e-6, False)
        lin_proj = W.w[p + '.proj_in'].dim() == 2               # SD2.1 use_linear_projection: nn.Linear on the token layout (same memory)
        b = p + '.transformer_blocks.0'
        if b in self._lnf:
            return self._transformer_folded(p, b, x, h, ctx, heads, lin_proj)
        if lin_proj:
            h = linear(W, p + '.proj_in', h.view(B * H * Wd, C)).view(B, H * Wd, C)
        else:
            h = conv(W, p + '.proj_in', h, padding=0).view(B, H * Wd, C)
        n = ops.layer_norm(h, W.w[b + '.norm1'], W.b[b + '.norm1'])
        h = attention(W, b + '.attn1', n, n, heads, h, helper=self.helper)
        n = ops.layer_norm(h, W.w[b + '.norm2'], W.b[b + '.norm2'])
        h = attention(W, b + '.attn2', n, ctx, heads, h, kv=self._ctx_kv.get(b + '.attn2') if self._ctx_kv else None)
        n = ops.layer_norm(h, W.w[b + '.norm3'], W.b[b + '.norm3'])
        g = linear(W, b + '.ff.net.0.proj', n.reshape(B * H * Wd, C), act='geglu')        # GEGLU fused in the epilogue
        h = linear(W, b + '.ff.net.2', g, residual=h.reshape(B * H * Wd, C)).view(B, H, Wd, C)
        if lin_proj:
            return _view_keep_stats(linear(W, p + '.proj_out', h.view(B * H * Wd, C), residual=x.view(B * H * Wd, C), stats_rows=H * Wd), B, H, Wd, C)
        return conv(W, p + '.proj_out', h, padding=0, residual=x, stats=True)

    def _shortcut_halves(self, name, C1):
        """The 1x1 shortcut weight [Cout,1,1,C1+C2] split along its input channels (cached)."""
        key = (name, C1)
        if key not in self._sc_split:
            w = self.W.w[name]
            self._sc_split[key] = (w[..., :C1].contiguous(), w[..., C1:].contiguous())
        return self._sc_split[key]

    def _transformer_folded(self, p, b, x, hn, ctx, heads, lin_proj):
        """The transformer block without LayerNorm kernels: every producer of a normalised tensor leaves the row statistics
        (sum, sum of squares) in its epilogue, every consumer applies the normalisation algebraically in ITS epilogue
        (ops.gemm_ln).  x = block input (residual of proj_out), hn = GroupNorm(x)."""
        W, F = self.W, self._lnf[b]
        B, H, Wd, C = x.shape
        T = H * Wd
        eps = 1e-5
        w_in = W.w[p + '.proj_in']
        h0 = ops.gemm_ln(hn.view(B * T, C), w_in.reshape(w_in.shape[0], -1)[:, :C] if w_in.dim() == 4 else w_in, bias=W.b.get(p + '.proj_in'), rowstats=True)
        # ---- self-attention: Q | K (rows = tokens) and V^T (columns = tokens) straight from the un-normalised h0
        wqk, c1qk, c2qk = F['qk']
        qk = ops.gemm_ln(h0, wqk, bias=c2qk, ln=(h0._rs, c1qk, C, eps)).view(B, T, 2 * C)
        wv, c1v, c2v = F['v']
        Tp = (T + 7) // 8 * 8                                     # 16-byte rows for the TMA store (only the 2 x 2 test level pads)
        vT = torch.empty(B, C, T, device=x.device, dtype=F16) if Tp == T else torch.zeros(B, C, Tp, device=x.device, dtype=F16)
        ops.gemm_ln(wv.unsqueeze(0).expand(B, -1, -1), h0.view(B, T, C), ln_cols=(h0._rs, c1v, c2v, C, eps), out=vT if Tp == T else vT[:, :, :T])
        o = attn_core(qk[:, :, :C], qk[:, :, C:], vT, heads, T)
        h1 = ops.gemm_ln(o.reshape(B * T, C), W.w[b + '.attn1.to_out.0'], bias=W.b.get(b + '.attn1.to_out.0'), residual=h0, rowstats=True)
        # ---- cross-attention (K / V^T of the text context come from prepare())
        wq2, c1q2, c2q2 = F['q2']
        q2 = ops.gemm_ln(h1, wq2, bias=c2q2, ln=(h1._rs, c1q2, C, eps)).view(B, T, C)
        kv = self._ctx_kv.get(b + '.attn2') if self._ctx_kv else None
        if kv is None:
            kv = self.project_context(ctx.to(F16).contiguous())[b + '.attn2']
        k2, vT2 = kv
        o2 = attn_core(q2, k2, vT2, heads, ctx.shape[1])
        h2 = ops.gemm_ln(o2.reshape(B * T, C), W.w[b + '.attn2.to_out.0'], bias=W.b.get(b + '.attn2.to_out.0'), residual=h1, rowstats=True)
        # ---- feed-forward (GEGLU fused in the epilogue, after the folded LayerNorm)
        wff, c1ff, c2ff = F['ff']
        g = ops.gemm_ln(h2, wff, bias=c2ff, act='geglu', ln=(h2._rs, c1ff, C, eps))
        h3 = linear(W, b + '.ff.net.2', g, residual=h2).view(B, H, Wd, C)
        if lin_proj:
            return _view_keep_stats(linear(W, p + '.proj_out', h3.view(B * T, C), residual=x.view(B * T, C), stats_rows=T), B, H, Wd, C)
        return conv(W, p + '.proj_out', h3, padding=0, residual=x, stats=True)

    def down_path(self, h, tproj, ctx):
        cfg, nb = self.cfg, len(self.cfg['block_out'])
        skips = [h]
        for i in range(nb):
            for j in range(cfg['layers_per_block']):
                h = self.resnet(f'down_blocks.{i}.resnets.{j}', h, tproj)
                if i < nb - 1:
                    h = self.transformer(f'down_blocks.{i}.attentions.{j}', h, ctx, self.heads_at(i))
                skips.append(h)
            if i < nb - 1:
                h = conv(self.W, f'down_blocks.{i}.downsamplers.0.conv', h, stride=2, padding=1, stats=True)
                skips.append(h)
        return h, skips

    def mid(self, h, tproj, ctx):
        h = self.resnet('mid_block.resnets.0', h, tproj)
        h = self.transformer('mid_block.attentions.0', h, ctx, self.heads_at(len(self.cfg['block_out']) - 1))
        return self.resnet('mid_block.resnets.1', h, tproj)


def to_nhwc_f16(x_nchw, scale=1.0, shift=0.0):
    """[N,C,H,W] -> channels-last fp16, channels zero-padded to a multiple of 8 (one kernel: dwg_nchw_f32_to_nhwc_f16)."""
    if x_nchw.is_cuda and x_nchw.dtype == torch.float32:
        return ops.nchw_to_nhwc_f16(x_nchw, scale, shift)
    return _pad_c((x_nchw * scale + shift).permute(0, 2, 3, 1)).to(F16).contiguous()


class ControlNet(DiffusionNet):
    @torch.no_grad()
    def embed_condition(self, cond_nchw01):
        """controlnet_cond_embedding of the condition image(s) -> [Bc,h,w,C0] (before the residual add)."""
        W = self.W
        c = conv(W, 'controlnet_cond_embedding.conv_in', to_nhwc_f16(cond_nchw01), act='silu')      # SiLU in the conv epilogue
        nblk = 2 * (len(self.cfg['cond_embed']) - 1)
        for k in range(nblk):
            c = conv(W, f'controlnet_cond_embedding.blocks.{k}', c, stride=2 if k % 2 == 1 else 1, act='silu')
        return conv(W, 'controlnet_cond_embedding.conv_out', c)

    @torch.no_grad()
    def prepare(self, t, ctx, B, cond_nchw01=None):
        pre = super().prepare(t, ctx, B)
        if cond_nchw01 is not None:
            pre['cond_emb'] = self.embed_condition(cond_nchw01)
        return pre

    @torch.no_grad()
    def forward(self, sample_nchw, t, ctx, cond_nchw01, conditioning_scale=1.0, pre=None):
        """-> (list of 12 down residuals, mid residual), NHWC fp16.  ``pre`` = prepare(...) results computed earlier."""
        W = self.W
        B = sample_nchw.shape[0]
        if pre is None:
            pre = self.prepare(t, ctx, B, cond_nchw01)
        skips, h = self.features(sample_nchw, pre, cond_nchw01)
        return self.residuals(skips, h, conditioning_scale=conditioning_scale)

    @torch.no_grad()
    def features(self, sample_nchw, pre, cond_nchw01=None):
        """conv_in + condition embedding + down blocks + mid block -> (13 skip features, mid feature): the part that can run
        beside the UNet encoder on another stream."""
        W = self.W
        B = sample_nchw.shape[0]
        tproj, ctx = pre['tproj'], pre['ctx']
        self._ctx_kv = pre['ctx_kv']
        h = conv(W, 'conv_in', to_nhwc_f16(sample_nchw))
        # condition embedding: with classifier-free guidance the reference feeds the SAME condition image
        # for every sample (controlnet.py:33-55 prepare_image duplicates it); a batch-1 condition is
        # embedded once and added to every sample
        c = pre['cond_emb'] if 'cond_emb' in pre else self.embed_condition(cond_nchw01)
        assert c.shape[0] in (1, B), 'condition batch must be 1 or the sample batch'
        h = ops.add(h, c) if c.shape[0] == B else h + c
        h, skips = self.down_path(h, tproj, ctx)
        h = self.mid(h, tproj, ctx)
        return skips, h

    @torch.no_grad()
    def residuals(self, skips, h, conditioning_scale=1.0, add_to=None):
        """The zero convolutions.  ``add_to`` = (UNet skip list, UNet mid output): the UNet tensors the residuals are added to
        (diffusers: down_block_additional_residuals / mid_block_additional_residual) become the RESIDUAL operand of the zero
        convolutions' epilogues, which return the sums directly -- no separate add kernels, and the epilogue also accumulates
        the GroupNorm statistics the decoder needs for them."""
        W = self.W
        if add_to is not None and conditioning_scale == 1.0:
            u_skips, u_mid = add_to
            down = [conv(W, f'controlnet_down_blocks.{i}', s, padding=0, residual=u, stats=True) for i, (s, u) in enumerate(zip(skips, u_skips))]
            mid = conv(W, 'controlnet_mid_block', h, padding=0, residual=u_mid, stats=True)
            return down, mid
        down = [conv(W, f'controlnet_down_blocks.{i}', s, padding=0) for i, s in enumerate(skips)]
        mid = conv(W, 'controlnet_mid_block', h, padding=0)
        if conditioning_scale != 1.0:
            down = [d * conditioning_scale for d in down]
            mid = mid * conditioning_scale
        if add_to is not None:
            u_skips, u_mid = add_to
            return [ops.add(u, d) for u, d in zip(u_skips, down)], ops.add(u_mid, mid)
        return down, mid


class UNet(DiffusionNet):
    @torch.no_grad()
    def forward(self, sample_nchw, t, ctx, down_residuals=None, mid_residual=None, pre=None):
        """-> eps [B,out_ch,H,W] fp32 (NCHW, the reference layout)."""
        return self.decode(self.encode(sample_nchw, t, ctx, pre), down_residuals, mid_residual)

    @torch.no_grad()
    def encode(self, sample_nchw, t, ctx, pre=None):
        """conv_in + down blocks + mid block: everything that does not need the ControlNet residuals
        (diffusers adds them to the skip list / the mid output afterwards), so it can run beside the
        ControlNet on another stream (guidance._predict)."""
        W = self.W
        B = sample_nchw.shape[0]
        if pre is None:
            pre = self.prepare(t, ctx, B)
        tproj, ctx = pre['tproj'], pre['ctx']
        self._ctx_kv = pre['ctx_kv']
        h = conv(W, 'conv_in', to_nhwc_f16(sample_nchw), stats=True)
        h, skips = self.down_path(h, tproj, ctx)
        h = self.mid(h, tproj, ctx)
        return h, skips, tproj, ctx

    @torch.no_grad()
    def decode(self, state, down_residuals=None, mid_residual=None, summed=False):
        """summed=True: down_residuals / mid_residual already ARE skip + residual (ControlNet.residuals(add_to=...))."""
        W, cfg = self.W, self.cfg
        nb = len(cfg['block_out'])
        h, skips, tproj, ctx = state
        if summed:
            skips, h = list(down_residuals), mid_residual
        else:
            if down_residuals is not None:
                skips = [ops.add(s, r) for s, r in zip(skips, down_residuals)]
            if mid_residual is not None:
                h = ops.add(h, mid_residual)
        skips = list(skips)
        for i in range(nb):
            for j in range(cfg['layers_per_block'] + 1):
                h = self.resnet(f'up_blocks.{i}.resnets.{j}', h, tproj, x2=skips.pop())      # [h | skip] is never materialised
                if i > 0:
                    h = self.transformer(f'up_blocks.{i}.attentions.{j}', h, ctx, self.heads_at(nb - 1 - i))
            if i < nb - 1:
                h = h.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)            # nearest 2x (data movement)
                h = conv(W, f'up_blocks.{i}.upsamplers.0.conv', h, stats=True)
        h = gn(W, 'conv_norm_out', h, self.G, 1e-5, True)
        out = conv(W, 'conv_out', h, out_dtype=torch.float32)
        return out.permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------------------------ VAE encoder
class VAEEncoder:
    """AutoencoderKL.encode (+ quant_conv) forward and input-gradient backward."""

    def __init__(self, sd, cfg, device='cuda'):
        self.cfg, self.dev, self.G = cfg, device, cfg['groups']
        self.W = Weights(sd, device, with_dgrad=True)

    # ---- forward pieces that record what the backward needs
    def _resnet_fwd(self, p, x, tape):
        W, G = self.W, self.G
        a, st1 = gn(W, p + '.norm1', x, G, 1e-6, True, return_stats=True)
        h1 = conv(W, p + '.conv1', a, stats=True)
        b, st2 = gn(W, p + '.norm2', h1, G, 1e-6, True, return_stats=True)
        has_sc = W.has(p + '.conv_shortcut')
        sc = conv(W, p + '.conv_shortcut', x, padding=0) if has_sc else x
        out = conv(W, p + '.conv2', b, residual=sc, stats=True)
        tape.append(('resnet', p, x, st1, h1, st2, has_sc))
        return out

    def _resnet_bwd(self, rec, g):
        _, p, x, st1, h1, st2, has_sc = rec
        W, G = self.W, self.G
        g_b = self._dgrad(p + '.conv2', g)
        g_h1 = ops.group_norm_bwd(h1, g_b, st2, W.w[p + '.norm2'], W.b[p + '.norm2'], G, 1e-6, True)
        g_a = self._dgrad(p + '.conv1', g_h1)
        g_sc = self._dgrad(p + '.conv_shortcut', g, k1=True) if has_sc else g
        return ops.group_norm_bwd(x, g_a, st1, W.w[p + '.norm1'], W.b[p + '.norm1'], G, 1e-6, True, dx_add=g_sc)

    def _dgrad(self, name, g, k1=False):
        """Input gradient of a stride-1 'same' convolution: conv with flipped, transposed taps."""
        w = self.W.dg[name]
        gp = _pad_c(g) if g.shape[-1] != w.shape[-1] else g
        y = ops.conv2d_nhwc(gp, w, stride=1, padding=0 if w.shape[1] == 1 else 1)
        return y

    def forward(self, images01_nchw, eps_nchw, tape=None):
        """latents = (mean + exp(0.5 clamp(logvar)) * eps) * scaling_factor   [B,4,h,w] fp32.
        With ``tape`` (a list) the activations needed by backward() are recorded."""
        W, cfg, G = self.W, self.cfg, self.G
        tape = [] if tape is None else tape
        nb = len(cfg['block_out'])
        x = to_nhwc_f16(images01_nchw, 2.0, -1.0)                          # [0,1] -> [-1,1] in the same kernel
        h = conv(W, 'encoder.conv_in', x, stats=True)
        for i in range(nb):
            for j in range(cfg['layers_per_block']):
                h = self._resnet_fwd(f'encoder.down_blocks.{i}.resnets.{j}', h, tape)
            if i < nb - 1:
                Hh, Ww = h.shape[1], h.shape[2]
                tape.append(('down', f'encoder.down_blocks.{i}.downsamplers.0.conv', (Hh, Ww)))
                h = conv(W, f'encoder.down_blocks.{i}.downsamplers.0.conv', h, stride=2, padding=(0, 0), out_hw=(Hh // 2, Ww // 2), stats=True)
        h = self._resnet_fwd('encoder.mid_block.resnets.0', h, tape)
        h = self._attn_fwd('encoder.mid_block.attentions.0', h, tape)
        h = self._resnet_fwd('encoder.mid_block.resnets.1', h, tape)
        a, st = gn(W, 'encoder.conv_norm_out', h, G, 1e-6, True, return_stats=True)
        tape.append(('norm_out', h, st))
        m = conv(W, 'encoder.conv_out', a)
        m = conv(W, 'quant_conv', m, padding=0, out_dtype=torch.float32)            # [B,h,w,8] fp32
        L = cfg['latent']
        mean, logvar = m[..., :L].permute(0, 3, 1, 2), m[..., L:2 * L].permute(0, 3, 1, 2)
        lv = torch.clamp(logvar, -30.0, 20.0)
        std = torch.exp(0.5 * lv)
        tape.append(('sample', std, (logvar > -30.0) & (logvar < 20.0), eps_nchw))
        return (mean + std * eps_nchw) * cfg['scaling_factor']

    def _attn_fwd(self, p, h, tape):
        W = self.W
        B, H, Wd, C = h.shape
        T = H * Wd
        n, st = gn(W, p + '.group_norm', h, self.G, 1e-6, False, return_stats=True)
        n2 = n.view(B * T, C)
        q = linear(W, p + '.to_q', n2).view(B, T, C)
        k = linear(W, p + '.to_k', n2).view(B, T, C)
        v = linear(W, p + '.to_v', n2).view(B, T, C)
        P = ops.gemm(q, k, alpha=C ** -0.5)                                  # [B,T,T] fp16
        ops.softmax_rows_(P, T)
        vT = v.transpose(1, 2).contiguous()
        a = ops.gemm(P, vT)                                                  # [B,T,C]
        out = _view_keep_stats(linear(W, p + '.to_out.0', a.view(B * T, C), residual=h.view(B * T, C), stats_rows=T), B, H, Wd, C)
        tape.append(('attn', p, h, st, n, q, k, v, P))
        return out

    def _attn_bwd(self, rec, g):
        _, p, h, st, n, q, k, v, P = rec
        W = self.W
        B, H, Wd, C = h.shape
        T = H * Wd
        s = C ** -0.5
        g2 = g.view(B * T, C)
        g_a = ops.gemm(g2, W.dg[p + '.to_out.0']).view(B, T, C)              # dL/da = g @ Wo
        g_P = ops.gemm(g_a, v)                                               # [B,T,T] = g_a @ v^T
        g_v = ops.gemm(P.transpose(1, 2).contiguous(), g_a.transpose(1, 2).contiguous())        # P^T g_a
        ops.softmax_rows_bwd_(P, g_P)                                        # g_P <- dS (unscaled)
        g_q = ops.gemm(g_P, k.transpose(1, 2).contiguous(), alpha=s)        # dS k
        g_k = ops.gemm(g_P.transpose(1, 2).contiguous(), q.transpose(1, 2).contiguous(), alpha=s)   # dS^T q
        g_n = ops.gemm(g_q.view(B * T, C), W.dg[p + '.to_q'])
        g_n = ops.gemm(g_k.view(B * T, C), W.dg[p + '.to_k'], residual=g_n)
        g_n = ops.gemm(g_v.view(B * T, C), W.dg[p + '.to_v'], residual=g_n).view(B, H, Wd, C)
        return ops.group_norm_bwd(h, g_n, st, W.w[p + '.group_norm'], W.b[p + '.group_norm'], self.G, 1e-6, False, dx_add=g)

    def backward(self, tape, g_latents_nchw):
        """dL/d(images01) [B,3,H,W] fp32 from dL/d(latents)."""
        W, cfg, G = self.W, self.cfg, self.G
        L = cfg['latent']
        tape = list(tape)
        _, std, unclamped, eps = tape.pop()
        gl = g_latents_nchw.float() * cfg['scaling_factor']
        g_mean = gl
        g_logvar = gl * eps * std * 0.5 * unclamped.float()
        gm = torch.cat([g_mean, g_logvar], dim=1).permute(0, 2, 3, 1).to(F16).contiguous()         # [B,h,w,8]
        g = ops.conv2d_nhwc(gm, W.dg['quant_conv'], padding=0)
        g = self._dgrad('encoder.conv_out', g)
        _, h, st = tape.pop()
        g = ops.group_norm_bwd(h, g, st, W.w['encoder.conv_norm_out'], W.b['encoder.conv_norm_out'], G, 1e-6, True)
        while tape:
            rec = tape.pop()
            if rec[0] == 'resnet':
                g = self._resnet_bwd(rec, g)
            elif rec[0] == 'attn':
                g = self._attn_bwd(rec, g)
            elif rec[0] == 'down':
                _, name, (Hh, Ww) = rec
                up = torch.zeros(g.shape[0], Hh, Ww, g.shape[-1], device=g.device, dtype=F16)
                up[:, ::2, ::2] = g                                       # zero-insertion (data movement)
                g = ops.conv2d_nhwc(up, W.dg[name], stride=1, padding=(2, 2), out_hw=(Hh, Ww))
        g = self._dgrad('encoder.conv_in', g)                               # [B,H,W,8] (3 valid channels)
        return ops.nhwc_f16_to_nchw(g, 3, 2.0)                              # d(2x - 1)/dx = 2, layout + type change in one kernel


class _VaeEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images01, eps, enc):
        tape = []
        with torch.no_grad():
            lat = enc.forward(images01, eps, tape)
        ctx.enc, ctx.tape = enc, tape
        return lat

    @staticmethod
    def backward(ctx, g):
        with torch.no_grad():
            gi = ctx.enc.backward(ctx.tape, g)
        ctx.tape = None
        return gi, None, None


def vae_encode(enc: VAEEncoder, images01_nchw, eps_nchw):
    """Differentiable (w.r.t. the image) encode_images (core/guidance/vae.py:34-40)."""
    return _VaeEncodeFn.apply(images01_nchw, eps_nchw, enc)
