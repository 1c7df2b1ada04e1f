"""(f3) Checkpoint / real-weight loading.

Avatar side -- the reference's own checkpoint format (core/trainer.py:194-259): ``step_%06d.pth`` = torch.save of
{'train_step', 'checkpoints', 'model': Scene.state_dict() [, 'optimizers', 'scaler']} (or a bare state dict), keys
prefixed by the Scene's module names ('avatar._positions', 'avatar.nerf_encoder.embeddings', ...).
Loading follows Scene.load_state_dict (core/system/scene.py:188-207): the avatar is first RESET to the checkpoint's
Gaussian count (DreamWaltzG.reset_by_state_dict, core/system/avatar.py:1254-1281: per-Gaussian parameters are re-created
with the stored shapes), then load_state_dict(strict=False) reports missing / unexpected keys.

Diffusion side -- diffusers-format weights of UNet2DConditionModel / ControlNetModel / AutoencoderKL as the reference
obtains them through from_pretrained (core/guidance/basic.py:119-210): ``*.safetensors`` (parsed here: the library is
not a dependency) or torch pickles (``*.bin`` / ``*.pth``), fp32 or the fp16 variant, checked against the key set of the
architecture (dwg/diffusion/weights.py) before they are handed to dwg.diffusion.model.
"""
import json
import os
import struct

import numpy as np
import torch

PER_GAUSSIAN = ('_positions', '_scales', '_quaternions', '_lbs_weights', '_opacities', '_sh_features_dc', '_sh_features_rest')

_ST_DTYPES = {'F32': (np.float32, torch.float32), 'F16': (np.float16, torch.float16), 'F64': (np.float64, torch.float64),
              'I64': (np.int64, torch.int64), 'I32': (np.int32, torch.int32), 'U8': (np.uint8, torch.uint8), 'BOOL': (np.bool_, torch.bool),
              'BF16': (np.uint16, torch.bfloat16)}


# ------------------------------------------------------------------------------------ safetensors
def load_safetensors(path):
    """Minimal reader of the safetensors container: u64 header length, JSON header {name: {dtype, shape, data_offsets}},
    raw little-endian tensor bytes."""
    with open(path, 'rb') as f:
        n = struct.unpack('<Q', f.read(8))[0]
        header = json.loads(f.read(n).decode('utf-8'))
        base = 8 + n
        data = np.memmap(path, dtype=np.uint8, mode='r', offset=base)
    out = {}
    for name, meta in header.items():
        if name == '__metadata__':
            continue
        npdt, tdt = _ST_DTYPES[meta['dtype']]
        a, b = meta['data_offsets']
        arr = np.frombuffer(data[a:b], dtype=npdt).reshape(meta['shape'])
        t = torch.from_numpy(np.array(arr))
        out[name] = t.view(torch.bfloat16) if meta['dtype'] == 'BF16' else t
    return out


def save_safetensors(path, tensors, metadata=None):
    """Writer of the same container (tests / exporting converted weights)."""
    rev = {v[1]: k for k, v in _ST_DTYPES.items()}
    header, blobs, off = {}, [], 0
    for name, t in tensors.items():
        t = t.detach().cpu().contiguous()
        raw = (t.view(torch.uint16) if t.dtype == torch.bfloat16 else t).numpy().tobytes()
        header[name] = {'dtype': rev[t.dtype], 'shape': list(t.shape), 'data_offsets': [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    if metadata:
        header['__metadata__'] = metadata
    hj = json.dumps(header, separators=(',', ':')).encode('utf-8')
    hj += b' ' * ((8 - len(hj) % 8) % 8)
    with open(path, 'wb') as f:
        f.write(struct.pack('<Q', len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def load_diffusers_state_dict(path, expected_keys=None, dtype=torch.float32):
    """path: a .safetensors / .bin / .pth file, or a diffusers model directory (diffusion_pytorch_model[.fp16].safetensors
    / .bin inside).  Returns {name: tensor(dtype)}; raises KeyError when the key set does not match the architecture."""
    if os.path.isdir(path):
        for cand in ('diffusion_pytorch_model.safetensors', 'diffusion_pytorch_model.fp16.safetensors', 'diffusion_pytorch_model.bin',
                     'diffusion_pytorch_model.fp16.bin'):
            if os.path.exists(os.path.join(path, cand)):
                path = os.path.join(path, cand)
                break
        else:
            raise FileNotFoundError(f'no diffusers weight file in {path}')
    sd = load_safetensors(path) if path.endswith('.safetensors') else torch.load(path, map_location='cpu', weights_only=True)
    if 'state_dict' in sd and isinstance(sd['state_dict'], dict):
        sd = sd['state_dict']
    sd = {k: v.to(dtype) for k, v in sd.items() if torch.is_tensor(v)}
    if expected_keys is not None:
        exp = set(expected_keys)
        missing = sorted(exp - set(sd))
        if missing:
            raise KeyError(f'{path}: {len(missing)} weights of the architecture are missing, e.g. {missing[:4]}')
        sd = {k: v for k, v in sd.items() if k in exp}          # e.g. AutoencoderKL files also hold the decoder
    return sd


# ------------------------------------------------------------------------------------ avatar / scene
def reset_by_state_dict(avatar, avatar_sd):
    """DreamWaltzG.reset_by_state_dict (avatar.py:1254-1281): re-create the per-Gaussian parameters with the checkpoint's
    Gaussian count (keeping each parameter's requires_grad), drop the canonical-pose cache."""
    for name in PER_GAUSSIAN:
        if name in avatar_sd and hasattr(avatar, name):
            old = getattr(avatar, name)
            new = torch.nn.Parameter(torch.empty_like(avatar_sd[name], device=old.device, dtype=old.dtype), requires_grad=old.requires_grad)
            setattr(avatar, name, new)
    for mname, gm in getattr(avatar, 'mesh_binding_gaussians', {}).items():
        pre = f'mesh_binding_gaussians.{mname}.'
        for k in ('_bary_coords', '_vertex_coords', '_scales'):
            if pre + k in avatar_sd:
                old = getattr(gm, k)
                setattr(gm, k, torch.nn.Parameter(torch.empty_like(avatar_sd[pre + k], device=old.device, dtype=old.dtype), requires_grad=old.requires_grad))
        for k in ('predefined_vertex_indices', 'triangles', 'points_to_vertices'):
            if pre + k in avatar_sd:
                gm.register_buffer(k, torch.empty_like(avatar_sd[pre + k], device=getattr(gm, k).device))
    avatar._canonical_cache = None


def organize_state_dict(state_dict):
    """scene.py:188-195."""
    by = {}
    for k, v in state_dict.items():
        mod = k.split('.')[0]
        by.setdefault(mod, {})[k.replace(f'{mod}.', '', 1)] = v
    return by


def load_scene_state_dict(scene, state_dict, strict=False):
    """Scene.load_state_dict (scene.py:202-207)."""
    by = organize_state_dict(state_dict)
    reset_by_state_dict(scene.avatar, by.get('avatar', {}))
    res = scene.load_state_dict(state_dict, strict=strict)
    for gm in getattr(scene.avatar, 'mesh_binding_gaussians', {}).values():      # derived kernel tables follow the loaded mesh
        gm._build_tables(scene.avatar.lbs_model)
    return res


def save_checkpoint(path, scene, train_step, past_checkpoints=None, optimizers=None, full=False):
    """Trainer.save_checkpoint (trainer.py:238-259)."""
    state = {'train_step': int(train_step), 'checkpoints': list(past_checkpoints or [])}
    if full and optimizers is not None:
        state['optimizers'] = [o.state_dict() for o in optimizers]
    state['model'] = scene.state_dict()
    torch.save(state, path)
    return path


def load_checkpoint(path, scene, optimizers=None, model_only=False, map_location=None):
    """Trainer.load_checkpoint (trainer.py:194-236).  Returns (train_step or None, missing_keys, unexpected_keys)."""
    ck = torch.load(path, map_location=map_location or next(scene.parameters()).device, weights_only=False)
    if 'model' not in ck:
        res = load_scene_state_dict(scene, ck, strict=True)
        return None, list(res.missing_keys), list(res.unexpected_keys)
    res = load_scene_state_dict(scene, ck['model'], strict=False)
    if not model_only and optimizers is not None and 'optimizers' in ck:
        for o, sd in zip(optimizers, ck['optimizers']):
            o.load_state_dict(sd)
    return ck.get('train_step'), list(res.missing_keys), list(res.unexpected_keys)
