"""Dev helper: per-role instruction / stall summary of a pw_tcgen05 ncu report (source page CSV).
usage: ncu_roles.py <report.ncu-rep>   (roles are split at the USETMAXREG markers and the final barrier)"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__grid_size']:
    print(f"{k:70s} {d.get(k)}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh = rows[1]; ia = hh.index('Source'); ie = hh.index('Instructions Executed'); isamp = hh.index('# Samples')
stall_cols = [i for i, c in enumerate(hh) if c.startswith('stall_') and 'Not Issued' not in c]
data = [r for r in rows[2:] if len(r) > ie]
marks = [i for i, r in enumerate(data) if 'USETMAXREG' in r[ia]]
bounds = [0] + marks + [len(data)]
names = ['prologue', 'ctrl(TMA+MMA)', 'transform+epilogue'] if len(marks) == 2 else [f'seg{i}' for i in range(len(bounds))]
# split transform / epilogue at the first LDTM (tcgen05.ld)
ldtm = next((i for i, r in enumerate(data) if 'LDTM' in r[ia]), None)
if len(marks) == 2 and ldtm:
    # walk back to the nearest backward branch target is hard; use first LDTM - 150 as a rough split
    bounds = [0, marks[0], marks[1], max(marks[1] + 1, ldtm - 120), len(data)]
    names = ['prologue', 'ctrl(TMA+MMA)', 'transform', 'epilogue(+tail)']
tot = sum(int(r[ie] or 0) for r in data)
for n, (a, b) in zip(names, zip(bounds[:-1], bounds[1:])):
    blk = data[a:b]
    ex = sum(int(r[ie] or 0) for r in blk); sm = sum(int(r[isamp] or 0) for r in blk)
    poll = sum(int(r[ie] or 0) for r in blk if 'SYNCS' in r[ia] or 'NANOSLEEP' in r[ia])
    st = collections.Counter()
    for r in blk:
        for i in stall_cols: st[hh[i].replace('stall_', '')] += int(r[i] or 0)
    print(f"{n:18s} instrs {a}-{b} exec {ex:>11d} ({100*ex/tot:4.1f}%) poll-ish {poll:>10d} samples {sm:6d}", st.most_common(6))
