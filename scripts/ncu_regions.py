"""Dev helper: per-source-line-range instruction / stall summary of an ncu report with --import-source on (-lineinfo build).
usage: ncu_regions.py <report.ncu-rep> [kernel-file-substring]
Splits the kernel by the role comment markers of csrc/gemm_tcgen05.cu (TMA producer / MMA issuer / A transform / epilogue)."""
import csv, collections, subprocess, sys, io, re, os
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_xu.sum', 'launch__registers_per_thread',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum',
          'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']:
    if k in d: print(f"{k:70s} {d.get(k)[:90]}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hh = rows[hi]
ia = hh.index('Source'); ie = hh.index('Instructions Executed'); isamp = hh.index('# Samples')
stall_cols = [i for i, c in enumerate(hh) if c.startswith('stall_') and 'Not Issued' not in c]
data = [r for r in rows[hi + 1:] if len(r) > ie]
# role boundaries in SASS: first UTMALDG (producer), first UTCHMMA (issuer), first F2FP (transform), first LDTM (epilogue)
def first(pat):
    return next((i for i, r in enumerate(data) if re.search(pat, r[ia])), None)
marks = sorted([(first(p), n) for p, n in [('UTMALDG', 'producer'), ('UTC.MMA|UTCHMMA', 'mma issuer'), ('F2FP', 'transform'), ('LDTM', 'epilogue')] if first(p) is not None])
bounds = [0] + [max(0, m[0] - 40) for m in marks] + [len(data)]
names = ['prologue'] + [m[1] for m in marks]
tot = sum(int(r[ie] or 0) for r in data); tots = sum(int(r[isamp] or 0) for r in data)
print(f"total warp-instructions {tot}, samples {tots}")
for n, (a, b) in zip(names, zip(bounds[:-1], bounds[1:])):
    blk = data[a:b]
    ex = sum(int(r[ie] or 0) for r in blk); sm = sum(int(r[isamp] or 0) for r in blk)
    st = collections.Counter()
    for r in blk:
        for i in stall_cols: st[hh[i].replace('stall_', '')] += int(r[i] or 0)
    print(f"{n:12s} sass {a}-{b} exec {ex:>11d} ({100*ex/max(tot,1):4.1f}%) samples {sm:6d} ({100*sm/max(tots,1):4.1f}%)", [(k, v) for k, v in st.most_common(6)])
if os.environ.get('TOP'):
    top = sorted(data, key=lambda r: -int(r[isamp] or 0))[:int(os.environ['TOP'])]
    for r in top:
        print(f"{int(r[isamp] or 0):6d} {int(r[ie] or 0):9d}  {r[ia][:110]}")
