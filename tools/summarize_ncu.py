"""Turn the Nsight Compute captures of tools/capture_profiles.sh into the small CSV summaries kept under profiles/.
Runs without a GPU (ncu -i ... --page raw --csv).      python tools/summarize_ncu.py <tag>
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.environ.get('DWG_PROFILES_DIR', os.path.join(ROOT, 'profiles'))          # capture_profiles.sh summarises on the GPU box into gpurun_out/

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('lts__t_bytes.sum', 'l2_bytes'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_pct_active'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct_elapsed'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved_occupancy_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem'),
    ('launch__waves_per_multiprocessor', 'waves'),
]
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}


def short(name):
    name = name.replace('void ', '').replace('dwg::', '').replace('(anonymous namespace)::', '')
    return name.split('(')[0]


def full_reports(tag):
    rows_out = []
    for rep in sorted(glob.glob(os.path.join(OUT, f'{tag}_*.ncu-rep'))):
        r = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for row in rows[2:]:
            d = {'report': os.path.basename(rep), 'kernel': short(row[col['Kernel Name']]), 'grid': row[col['Grid Size']], 'block': row[col['Block Size']]}
            for m, key in METRICS:
                if m not in col:
                    d[key] = ''
                    continue
                v, u = row[col[m]].replace(',', ''), units[col[m]]
                try:
                    v = float(v) * SCALE.get(u, 1.0)
                except ValueError:
                    pass
                d[key] = v
            if isinstance(d.get('dram_read'), float) and isinstance(d.get('dram_write'), float) and d.get('duration'):
                d['dram_traffic_MB'] = round((d['dram_read'] + d['dram_write']) / 1e6, 3)
                d['dram_GBps'] = round((d['dram_read'] + d['dram_write']) / (d['duration'] * 1e-6) / 1e9, 1)
            rows_out.append(d)
    if not rows_out:
        return
    keys = ['report', 'kernel', 'grid', 'block', 'duration', 'dram_traffic_MB', 'dram_GBps'] + [k for _, k in METRICS[1:]]
    path = os.path.join(PROF, f'{tag}_ncu_full_summary.csv')
    with open(path, 'w') as fh:
        fh.write('# ncu --set full --clock-control none --import-source on (tools/capture_profiles.sh); duration in us, bytes in B; '
                 'one row per captured launch of tools/ncu_targets.py\n')
        w = csv.DictWriter(fh, keys, extrasaction='ignore')
        w.writeheader()
        for d in rows_out:
            w.writerow({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})
    print('wrote', path, len(rows_out), 'rows')


def launch_list(tag):
    path = os.path.join(OUT, f'{tag}_launches_bench.csv')
    if not os.path.exists(path):
        return
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.reader(io.StringIO(text[start:])))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    kn, mv, mu = col['Kernel Name'], col['Metric Value'], col['Metric Unit']
    launches = []
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        try:
            launches.append((short(r[kn]), float(r[mv].replace(',', '')) * SCALE.get(r[mu], 1.0)))
        except ValueError:
            pass
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, us in launches:
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    out = os.path.join(PROF, f'{tag}_launch_shares_bench.csv')
    with open(out, 'w') as fh:
        fh.write('# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 --skip-cpu-baseline\n'
                 '# every launch of the whole process (graph warm-up + capture + timed steps + the instrumented roofline pass); per-launch\n'
                 '# times are cold-cache and serialised: compare SHARES, not absolutes\n')
        fh.write('kernel,launches,total_us,share\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f'"{k}",{n},{us:.1f},{us / tot:.4f}\n')
    print('wrote', out, len(launches), 'launches')


def gemm_traffic(tag):
    """Per-launch DRAM bytes / duration of the tensor-core launches of one guidance pass -> profiles/<tag>_gemm_traffic.json"""
    import json
    path = os.path.join(OUT, f'{tag}_gemm_traffic.csv')
    if not os.path.exists(path):
        return
    text = open(path).read()
    rows = list(csv.reader(io.StringIO(text[text.find('"ID"'):])))
    col = {h: i for i, h in enumerate(rows[0])}
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        if len(r) <= col['Metric Value']:
            continue
        try:
            v = float(r[col['Metric Value']].replace(',', '')) * SCALE.get(r[col['Metric Unit']], 1.0)
        except ValueError:
            continue
        per[r[col['ID']]]['kernel'] = short(r[col['Kernel Name']])
        per[r[col['ID']]][r[col['Metric Name']]] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        a = agg['attention' if 'fa_fwd' in d['kernel'] else 'gemm']
        a[0] += 1
        a[1] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
        a[2] += d.get('gpu__time_duration.sum', 0.0)
    out = {k: {'launches': n, 'dram_bytes_total': b, 'dram_bytes_per_launch': b / max(n, 1), 'duration_us_total_under_ncu': t}
           for k, (n, b, t) in agg.items()}
    out['how'] = ('ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum '
                  '-k regex:gemm_kernel|fa_fwd python tools/gemm_pass.py (one un-graphed guidance pass, serialised launches)')
    dst = os.path.join(PROF, f'{tag}_gemm_traffic.json')
    json.dump(out, open(dst, 'w'), indent=1)
    print('wrote', dst, {k: v['launches'] for k, v in out.items() if isinstance(v, dict)})


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1b'
    os.makedirs(PROF, exist_ok=True)
    launch_list(tag)
    full_reports(tag)
    gemm_traffic(tag)
