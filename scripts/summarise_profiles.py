"""Summarises gpurun_out/<tag>_launches.csv (ncu launch list with DRAM bytes) into profiles/<tag>_launches_summary.csv,
profiles/<tag>_traffic.json, and the --set full reports into profiles/<tag>_ncu_full.csv.   usage: summarise_profiles.py <tag>"""
import csv, io, json, re, subprocess, sys, collections, os
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(f'{ROOT}/gpurun_out/{tag}_launches.csv')) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value'); iu = hdr.index('Metric Unit'); iid = hdr.index('ID')
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[iid], {'name': r[ik]})
    v = float(r[iv].replace(',', ''))
    u = r[iu]
    if r[im] == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}.get(u, 1.0)
    else:
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    d[r[im]] = v
def short(n):
    m = re.match(r'(?:void )?(?:orbit::)?(?:tc::|st::|mbs::|seg::|cf::)?([A-Za-z0-9_]+)', n)
    base = m.group(1) if m else n
    t = re.search(r'<(.*)>', n)
    return base + ('<' + t.group(1).replace('(int)', '').replace('(bool)', '').replace(' ', '') + '>' if t else '')
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(short(d['name']), {'launches': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
    a['launches'] += 1; a['us'] += d.get('gpu__time_duration.sum', 0); a['rd'] += d.get('dram__bytes_read.sum', 0); a['wr'] += d.get('dram__bytes_write.sum', 0)
tot = sum(a['us'] for a in agg.values())
with open(f'{ROOT}/profiles/{tag}_launches_summary.csv', 'w') as f:
    f.write('kernel,launches,total_us,share_pct,avg_us,dram_read_GB,dram_write_GB,dram_GBps_under_ncu\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        f.write(f"\"{k}\",{a['launches']},{a['us']:.1f},{100*a['us']/tot:.2f},{a['us']/a['launches']:.1f},{a['rd']/1e9:.3f},{a['wr']/1e9:.3f},"
                f"{(a['rd']+a['wr'])/max(a['us'],1e-9)/1e3:.0f}\n")
fam = {'pointwise_gemm': 'pw_tcgen05|pw_stream', 'depthwise_conv': 'dw2_kernel|dw5s_kernel|mbx_kernel|mbs_kernel', 'stem_conv': 'stem_kernel', 'se_gate': 'se_gate', 'spatial_mean': 'spatial_mean'}
# the TIMED episode only: the bench command runs calibration, 3 warm-up episodes, then ONE timed episode = the launches
# from the second-to-last stem_kernel (support pass; the last one is the query pass) to the end of the process
order = list(launch.values())
stems = [i for i, d in enumerate(order) if 'stem_kernel' in d['name']]
episode = order[stems[-2]:] if len(stems) >= 2 else order
ep_tot = sum(d.get('gpu__time_duration.sum', 0) for d in episode)
traffic = {}
for fname, pat in fam.items():
    sel = [d for d in episode if re.match(pat, short(d['name']))]
    if sel:
        traffic[fname] = {'launches': len(sel),
                          'dram_bytes_per_launch': sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in sel) / len(sel),
                          'dram_bytes_per_episode': sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in sel),
                          'share_of_episode_gpu_time_pct': 100 * sum(d.get('gpu__time_duration.sum', 0) for d in sel) / ep_tot}
json.dump({'command': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 --episodes 0',
           'note': 'launches of the ONE timed episode only (2,240 frames: a 1,600-frame support pass and a 640-frame query pass); per-launch averages',
           'episode_launches': len(episode), 'episode_dram_bytes': sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in episode),
           'families': traffic},
          open(f'{ROOT}/profiles/{tag}_traffic.json', 'w'), indent=1)
print(open(f'{ROOT}/profiles/{tag}_launches_summary.csv').read())
print(json.dumps(traffic, indent=1))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']
with open(f'{ROOT}/profiles/{tag}_ncu_full.csv', 'w') as f:
    first = True
    for rep in (f'{tag}_gemm_full', f'{tag}_stream_full', f'{tag}_dw_full', f'{tag}_dw5s_full', f'{tag}_mbx_full', f'{tag}_se_full'):
        path, exported = f'{ROOT}/gpurun_out/{rep}.ncu-rep', f'{ROOT}/gpurun_out/{rep}_raw.csv'   # (profile_round.sh exports on the box)
        if os.path.exists(exported) and os.path.getsize(exported) > 0:
            raw = open(exported).read()
        elif os.path.exists(path):
            raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        else:
            continue
        rr = list(csv.reader(io.StringIO(raw)))
        h, u = rr[0], rr[1]
        idx = [h.index(k) for k in keys if k in h]
        if first:
            f.write('Kernel Name,' + ','.join(h[i] for i in idx) + '\n,' + ','.join(('Mbyte' if h[i].startswith('dram__bytes') else 'us' if h[i] == 'gpu__time_duration.sum' else u[i]) for i in idx) + '\n'); first = False
        scale = {'Gbyte': 1e3, 'Kbyte': 1e-3, 'byte': 1e-6, 'ms': 1e3, 'ns': 1e-3}   # -> Mbyte / us as in the header row
        for r in rr[2:]:
            vals = []
            for i in idx:
                v = r[i]
                if h[i].startswith('dram__bytes') or h[i] == 'gpu__time_duration.sum':
                    v = f"{float(v.replace(',', '')) * scale.get(u[i], 1.0):.3f}"
                vals.append(v)
            f.write('"' + short(r[h.index('Kernel Name')]) + '",' + ','.join(vals) + '\n')
print(open(f'{ROOT}/profiles/{tag}_ncu_full.csv').read())
