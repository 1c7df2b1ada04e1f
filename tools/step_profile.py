"""Where one EAGER SDS step (dwg.step.SDSTrainStep, un-graphed, single process) spends its GPU time:
  1. every kernel by CUDA time (CUPTI, one step);
  2. the torch-issued work only (aten ops with their input shapes) -- the glue that is not a dwg kernel, which is what
     the fusion work list is made of.
python tools/step_profile.py [--tiny]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
import bench  # noqa: E402


def main():
    tiny = '--tiny' in sys.argv
    dev = 'cuda:0'
    torch.cuda.set_device(0)
    sc = bench.Workload(dev, 0, tiny=tiny, img=128 if tiny else 512)
    def view():
        data = sc.next_view()
        data['smpl_inputs'] = {k: v.to(dev) for k, v in data['smpl_inputs'].items()}
        return data
    for _ in range(3):
        sc.trainer.step(view())
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as pr:
        sc.trainer.step(view())
        torch.cuda.synchronize()
    ka = pr.key_averages()
    print(ka.table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=80))
    # ---- kernels: dwg vs torch-issued
    tot = collections.defaultdict(lambda: [0.0, 0])
    for e in pr.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            dwg = any(s in e.name for s in ('dwg', 'gemm::', 'nn::', 'attn::', 'raster::', 'mlp::', 'glbs::', 'pose::', 'grid_', 'lbs_skin', 'sh_',
                                            'adam', 'frame_pack', 'splitk'))
            kind = 'memset/memcpy' if ('Memset' in e.name or 'Memcpy' in e.name) else ('dwg kernels' if dwg else 'torch kernels')
            tot[kind][0] += e.device_time
            tot[kind][1] += 1
    print('--- kernel time by origin (one eager step)')
    for k, (us, n) in tot.items():
        print(f'{k:16s} n={n:5d}  {us / 1e3:8.3f} ms')
    # ---- torch ops by (name, shapes)
    print('--- torch-issued ops with device time, by input shapes (top 50)')
    rows = []
    for e in pr.key_averages(group_by_input_shape=True):
        if e.self_device_time_total > 0 and e.key.startswith('aten::'):
            rows.append((e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:110]))
    rows.sort(reverse=True)
    for us, n, name, shp in rows[:50]:
        print(f'{us / 1e3:8.3f} ms  n={n:4d}  {name:28s} {shp}')
    print(f'total torch-op device time {sum(r[0] for r in rows) / 1e3:.3f} ms over {sum(r[1] for r in rows)} ops')


if __name__ == '__main__':
    main()
