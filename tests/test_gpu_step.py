"""GPU: the step glue as a package API (rows R13 / R17): Scene.forward's background composite fused into the blend
epilogue, SDSTrainStep eager vs whole-step CUDA-graph replay, the deferred rasteriser overflow check, and the
bilinear input resize of prepare_latents."""
import numpy as np
import pytest
import torch

from dwg import avatar as dav, camera, ops, step as dstep, synth
from dwg.diffusion import guidance as G, weights as W

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _small_scene(n=3000, tri=200):
    model = synth.make_body_model(0)
    av = synth.make_avatar(model, n, tri, seed=3)
    m = dav.DreamWaltzGAvatar(model, av, device=DEV)
    torch.manual_seed(0)
    with torch.no_grad():
        m.nerf_encoder.embeddings.uniform_(-0.5, 0.5)
        for p in list(m.nerf_opacity_and_color_net.parameters()) + list(m.nerf_scale_and_quaternion_net.parameters()):
            p.copy_(torch.randn_like(p) * 0.3)
        m._scales.fill_(np.log(0.02))
    return dstep.Scene(m, dav.GaussianRenderer())


def _data(img=128, seed=0, row=2):
    rows = np.load(__import__('os').path.join(__import__('os').path.dirname(__file__), 'golden', 'poses.npz'))['rows']
    d = camera.random_camera(np.random.default_rng(seed), img, img)
    d['smpl_inputs'] = {k: v.to(DEV) for k, v in synth.pose_from_row(rows[row]).items()}
    return d


def test_scene_background_composite_is_fused_and_differentiable():
    """scene.py:153-166: image = image_fg + image_bg * (1 - alpha), forward and backward, against the same composite
    done with torch ops on the plain rasteriser outputs."""
    sc = _small_scene()
    data = _data()
    bg = torch.rand(1, 128, 128, 3, device=DEV).requires_grad_(True)
    sc.background = lambda d, shape: bg
    out = sc(data, smpl_observed_inputs=data['smpl_inputs'], use_densifier=False)
    assert set(('image', 'image_fg', 'image_bg', 'alpha', 'depth')) <= set(out)
    ref = out['image_fg'] + bg.detach() * (1 - out['alpha'].detach())
    torch.testing.assert_close(out['image'].detach(), ref, rtol=0, atol=1e-6)
    w = torch.randn_like(out['image'])
    params = [p for p in sc.parameters() if p.requires_grad]
    g_fused = torch.autograd.grad((out['image'] * w).sum(), params + [bg], allow_unused=True)
    sc.background = None
    out2 = sc(data, smpl_observed_inputs=data['smpl_inputs'], use_densifier=False)
    assert out2['image_fg'] is out2['image']
    comp = out2['image'] + bg * (1 - out2['alpha'])
    g_ref = torch.autograd.grad((comp * w).sum(), params + [bg], allow_unused=True)
    for a, b in zip(g_fused, g_ref):
        assert (a is None) == (b is None)
        if a is not None and float(b.abs().max()) > 0:
            assert rel(a, b) < 2e-4, rel(a, b)
    # pure-colour mode of the reference (background.py:14-27)
    out3 = sc(data, smpl_observed_inputs=data['smpl_inputs'], use_densifier=False, bg_mode='white')
    torch.testing.assert_close(out3['image'], out3['image_fg'] + (1 - out3['alpha']), rtol=0, atol=1e-6)


def test_train_step_graph_replay_equals_eager_step():
    """trainer.train_forward + backward (core/trainer.py:856-1017) through SDSTrainStep: one whole-step CUDA-graph replay
    gives the same SDS gradient (bitwise: every reduction of the diffusion path has a fixed order) and the same parameter
    gradients (fp32 atomics in the raster / skinning / grid backward: 1e-4) as the eager step, for fixed draws."""
    cfg, vcfg = W.TINY, W.TINY_VAE
    gd = G.ControlNetScoreDistillation(W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg), cfg, vcfg, DEV,
                                       guidance_scale=50.0, default_image_size=128)
    sc = _small_scene()
    g = torch.Generator().manual_seed(4)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV)}
    cond = (torch.rand(1, 3, 128, 128, generator=g) > 0.95).float().to(DEV)
    tr = dstep.SDSTrainStep(sc, gd, emb, allreduce=False)
    tr.fixed_draws = {'timestep': torch.tensor([400], device=DEV), 'noise': torch.randn(1, 4, 16, 16, generator=g).to(DEV),
                      'vae_eps': torch.randn(1, 4, 16, 16, generator=g).to(DEV)}
    d0, d1 = _data(seed=1, row=1), _data(seed=2, row=4)
    d0['cond_images'], d1['cond_images'] = cond, cond
    loss, ro, so, text = tr.step(d1)                                            # eager
    assert float(loss) == 1.0 and text is None and ro['regularizations'] == {}
    eager_sds, eager_flat, eager_img = so['gradients'].clone(), tr.bucket.flat.clone(), ro['image'].detach().clone()
    assert all(p.grad.data_ptr() == tr.bucket.flat.data_ptr() + o * 4 for p, o in zip(tr.params, tr.bucket.offsets))
    assert float(eager_flat.abs().sum()) > 0
    tr.capture(d0)                                                              # captured on ANOTHER view / pose
    loss, ro, so, _ = tr.step(d1)                                               # replay with d1's camera + pose
    torch.cuda.synchronize()
    assert torch.equal(ro['image'], eager_img)
    assert torch.equal(so['gradients'], eager_sds)
    assert rel(tr.bucket.flat, eager_flat) < 1e-4, rel(tr.bucket.flat, eager_flat)
    assert tr.graph_launches > 100 and set(tr.host_ms) == {'inputs', 'graph_launch', 'post'}
    prev = (float(loss), float(so['gradients'].abs().mean()), 400)
    assert tr.fetch_result(lag=0) == pytest.approx(prev, rel=1e-6)              # asynchronous 12-byte read-back of the step's scalars
    loss, ro, so, _ = tr.step(d0)                                               # and another view through the same graph
    assert tr.fetch_result(lag=1) == pytest.approx(prev, rel=1e-6)              # one step late: no stall on the step just issued
    torch.cuda.synchronize()
    assert not torch.equal(ro['image'], eager_img)


def test_rasteriser_overflow_is_reported_by_the_next_call():
    gs = synth.random_gaussians(4000, seed=1, extent=0.45, scale_range=(0.02, 0.06))
    d = camera.make_camera(2.2, 30.0, 80.0, 45.0, 128, 128)
    view, proj, campos, tfx, tfy = camera.raster_matrices(d)
    t = {k: v.to(DEV) for k, v in gs.items()}
    kw = dict(image_height=128, image_width=128, tanfovx=tfx, tanfovy=tfy, viewmatrix=view, projmatrix=proj, bg=torch.zeros(3))
    m2 = torch.zeros(4000, 3, device=DEV)
    st = []
    ops.rasterize(t['positions'], m2, t['colors'], t['opacities'], t['scales'], t['quaternions'], instance_capacity=2048, state_out=st, **kw)
    torch.cuda.synchronize()
    status = st[0].status.cpu().numpy()
    assert status[0] == 1 and status[1] > 2048                                  # flagged on the device, P reported
    with pytest.raises(RuntimeError, match='instance capacity exceeded'):
        ops.rasterize(t['positions'], m2, t['colors'], t['opacities'], t['scales'], t['quaternions'], **kw)
    assert ops.default_instance_capacity(10, DEV) >= int(status[1])             # the next default allocation is large enough
    ops.STATUS_MONITOR.reset()
    color, *_ = ops.rasterize(t['positions'], m2, t['colors'], t['opacities'], t['scales'], t['quaternions'], **kw)      # clean again
    torch.cuda.synchronize()
    assert torch.isfinite(color).all()


def test_guidance_resizes_non_default_renders_bilinearly():
    """basic.py:354-366 prepare_latents(input_interpolate=True): a render that is not default_image_size^2 is resized
    (bilinear, align_corners=False) in front of the VAE, and the gradient flows back through the resize."""
    cfg, vcfg = W.TINY, W.TINY_VAE
    gd = G.ControlNetScoreDistillation(W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg), cfg, vcfg, DEV,
                                       guidance_scale=7.5, default_image_size=128)
    g = torch.Generator().manual_seed(9)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV)}
    cond = torch.rand(1, 3, 128, 128, generator=g).to(DEV)
    big = torch.rand(1, 3, 192, 160, generator=g).to(DEV).requires_grad_(True)
    kw = dict(timestep=torch.tensor([300], device=DEV), noise=torch.randn(1, 4, 16, 16, generator=g).to(DEV),
              vae_eps=torch.randn(1, 4, 16, 16, generator=g).to(DEV))
    out = gd(big, emb, cond_inputs=cond, **kw)
    assert out['latents'].shape == (1, 4, 16, 16)
    out['diffusion_loss'].backward()
    small = torch.nn.functional.interpolate(big.detach(), (128, 128), mode='bilinear', align_corners=False).requires_grad_(True)
    out2 = gd(small, emb, cond_inputs=cond, **kw)
    assert torch.equal(out2['latents'], out['latents'])
    out2['diffusion_loss'].backward()
    ref = torch.autograd.grad(torch.nn.functional.interpolate(big, (128, 128), mode='bilinear', align_corners=False), big, small.grad)[0]
    torch.testing.assert_close(big.grad, ref, rtol=1e-5, atol=1e-7)


def test_train_step_with_device_side_condition_producer():
    """(f1) in the step: the condition image is produced on the device from the posed body's keypoints and the view's own
    depth / alpha, eager and under the whole-step graph (device-resident camera), and equals the producer run by hand."""
    from dwg import condition
    cfg, vcfg = W.TINY, W.TINY_VAE
    gd = G.ControlNetScoreDistillation(W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg), cfg, vcfg, DEV,
                                       guidance_scale=50.0, default_image_size=128)
    sc = _small_scene()
    model = synth.make_body_model(0)
    ks = condition.synthetic_keypoint_source(sc.avatar.lbs_model, model)
    prod = condition.PoseConditionProducer(128, 128, device=DEV)
    g = torch.Generator().manual_seed(4)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV)}
    tr = dstep.SDSTrainStep(sc, gd, emb, allreduce=False, cond_producer=prod, keypoint_source=ks)
    tr.fixed_draws = {'timestep': torch.tensor([400], device=DEV), 'noise': torch.randn(1, 4, 16, 16, generator=g).to(DEV),
                      'vae_eps': torch.randn(1, 4, 16, 16, generator=g).to(DEV)}
    d0, d1 = _data(seed=1, row=1), _data(seed=2, row=4)
    d0['cond_images'] = d1['cond_images'] = torch.zeros(1, 3, 128, 128, device=DEV)      # placeholder: replaced by the produced image
    loss, ro, so, _ = tr.step(d1)
    cond_eager, sds_eager = ro['cond_images'].clone(), so['gradients'].clone()
    assert cond_eager.shape == (1, 3, 128, 128) and float(cond_eager.sum()) > 0           # a skeleton was drawn
    jt = sc.avatar.lbs_model.joint_transforms(**d1['smpl_inputs'])
    by_hand = prod(ks(jt), d1, {'depth': ro['depth'], 'alpha': ro['alpha']})
    assert torch.equal(by_hand, cond_eager)
    kp = ks(jt)
    assert kp.shape == (128, 3) and torch.isfinite(kp).all()
    del loss, ro, so
    tr.capture(d0)
    loss, ro, so, _ = tr.step(d1)
    torch.cuda.synchronize()
    assert (ro['cond_images'] != cond_eager).float().mean() < 1e-3                       # device-side fov arithmetic: boundary pixels at most
    assert rel(so['gradients'], sds_eager) < 2e-2
