"""Per-op determinism + accuracy inside one transformer block / resnet (SD1.5 level-0 shapes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dreamwaltz-g_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    from dwg import ops
    from dwg.diffusion import model as M, weights as W
    dev = 'cuda'
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    cfg = W.SD15
    sd = W.make_unet(cfg)
    un = M.UNet(sd, cfg, dev)
    Wt = un.W
    C = cfg['block_out'][level]
    hw = 64 >> level
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, hw, hw, C, generator=g).to(dev).half()
    ctx = torch.randn(2, 77, cfg['ctx_dim'], generator=g).to(dev).half()
    p = f'down_blocks.{level}.attentions.0'
    b = p + '.transformer_blocks.0'
    heads = 8
    B, T = 2, hw * hw
    sdg = {k: v.to(dev) for k, v in sd.items() if k.startswith(p) or k.startswith(f'down_blocks.{level}.resnets.0')}

    def twice(name, fn, ref=None):
        a = fn(); torch.cuda.synchronize()
        worst, eq = 0.0, True
        for _ in range(4):
            bb = fn(); torch.cuda.synchronize()
            eq = eq and torch.equal(a, bb)
            worst = max(worst, rel(a, bb))
        acc = '' if ref is None else f' err_vs_fp32={rel(a.float(), ref):.3e}'
        print(f'{name:30s} equal={eq} run_to_run={worst:.3e}{acc}', flush=True)
        return a

    xf = x.float()
    n_ref = F.group_norm(xf.permute(0, 3, 1, 2), 32, sdg[p + '.norm.weight'], sdg[p + '.norm.bias'], 1e-6).permute(0, 2, 3, 1)
    h0 = twice('gn (no silu)', lambda: M.gn(Wt, p + '.norm', x, 32, 1e-6, False), n_ref)
    ref = F.conv2d(h0.float().permute(0, 3, 1, 2), sdg[p + '.proj_in.weight'], sdg[p + '.proj_in.bias']).permute(0, 2, 3, 1)
    h = twice('proj_in 1x1', lambda: M.conv(Wt, p + '.proj_in', h0, padding=0), ref).view(B, T, C)
    ref = F.layer_norm(h.float(), (C,), sdg[b + '.norm1.weight'], sdg[b + '.norm1.bias'])
    n = twice('layernorm', lambda: ops.layer_norm(h, Wt.w[b + '.norm1'], Wt.b[b + '.norm1']), ref)
    wq, wk, wv = sdg[b + '.attn1.to_q.weight'], sdg[b + '.attn1.to_k.weight'], sdg[b + '.attn1.to_v.weight']
    qk = twice('qk gemm', lambda: ops.gemm(n.reshape(B * T, C), Wt.w[b + '.attn1.to_qk']), torch.cat([n.float().reshape(B * T, C) @ wq.t(), n.float().reshape(B * T, C) @ wk.t()], 1)).view(B, T, 2 * C)
    q3, k3 = qk[:, :, :C], qk[:, :, C:]

    def vproj():
        vT = torch.empty(B, C, T, device=dev, dtype=torch.float16)
        for i in range(B):
            ops.gemm(Wt.w[b + '.attn1.to_v'], n[i], out=vT[i])
        return vT
    vT = twice('v^T gemm', vproj, (n.float() @ wv.t()).transpose(1, 2))
    sp = lambda t: t.float().reshape(B, -1, heads, C // heads).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q3), sp(k3), sp(vT.transpose(1, 2))).transpose(1, 2).reshape(B, T, C)
    o = twice('self attention (fused)', lambda: ops.attention(q3, k3, vT, heads, T), ref)
    ref = o.float().reshape(B * T, C) @ sdg[b + '.attn1.to_out.0.weight'].t() + sdg[b + '.attn1.to_out.0.bias'] + h.float().reshape(B * T, C)
    h1 = twice('to_out + residual', lambda: M.linear(Wt, b + '.attn1.to_out.0', o.reshape(B * T, C), residual=h.reshape(B * T, C)), ref).view(B, T, C)
    kv = un.project_context(ctx)[b + '.attn2']
    n2 = ops.layer_norm(h1, Wt.w[b + '.norm2'], Wt.b[b + '.norm2'])
    q2 = twice('cross q gemm', lambda: ops.gemm(n2.reshape(B * T, C), Wt.w[b + '.attn2.to_q']), n2.float().reshape(B * T, C) @ sdg[b + '.attn2.to_q.weight'].t()).view(B, T, C)
    ref = F.scaled_dot_product_attention(sp(q2), sp(kv[0]), sp(kv[1][:, :, :77].transpose(1, 2))).transpose(1, 2).reshape(B, T, C)
    o2 = twice('cross attention (fused)', lambda: ops.attention(q2, kv[0], kv[1], heads, 77), ref)
    n3 = ops.layer_norm(h1, Wt.w[b + '.norm3'], Wt.b[b + '.norm3'])
    gg = n3.float().reshape(B * T, C) @ sdg[b + '.ff.net.0.proj.weight'].t() + sdg[b + '.ff.net.0.proj.bias']
    a_, gate = gg.chunk(2, -1)
    gl = twice('geglu gemm', lambda: M.linear(Wt, b + '.ff.net.0.proj', n3.reshape(B * T, C), act='geglu'), a_ * F.gelu(gate))
    ref = gl.float() @ sdg[b + '.ff.net.2.weight'].t() + sdg[b + '.ff.net.2.bias'] + h1.float().reshape(B * T, C)
    h2 = twice('ff.net.2 + residual', lambda: M.linear(Wt, b + '.ff.net.2', gl, residual=h1.reshape(B * T, C)), ref).view(B, hw, hw, C)
    ref = F.conv2d(h2.float().permute(0, 3, 1, 2), sdg[p + '.proj_out.weight'], sdg[p + '.proj_out.bias']).permute(0, 2, 3, 1) + xf
    twice('proj_out + residual', lambda: M.conv(Wt, p + '.proj_out', h2, padding=0, residual=x), ref)
    un._ctx_kv = None
    twice('whole transformer', lambda: un.transformer(p, x, ctx, heads))
    # resnet pieces
    r = f'down_blocks.{level}.resnets.0'
    xin = torch.randn(2, hw, hw, Wt.w[r + '.conv1'].shape[-1], generator=g).to(dev).half()
    a1 = twice('gn+silu', lambda: M.gn(Wt, r + '.norm1', xin, 32, 1e-5, True),
               F.silu(F.group_norm(xin.float().permute(0, 3, 1, 2), 32, sdg[r + '.norm1.weight'], sdg[r + '.norm1.bias'], 1e-5)).permute(0, 2, 3, 1))
    ref = F.conv2d(a1.float().permute(0, 3, 1, 2), sdg[r + '.conv1.weight'], sdg[r + '.conv1.bias'], padding=1).permute(0, 2, 3, 1)
    twice('conv3x3', lambda: M.conv(Wt, r + '.conv1', a1), ref)
    twice('whole resnet', lambda: un.resnet(r, xin, None))


if __name__ == '__main__':
    main()
