"""Summarise an .ncu-rep: key metrics + top stall instructions (SASS) with source line mapping."""
import csv, subprocess, sys, io
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct', 'lts__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'sm__inst_executed.sum.per_cycle_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg', 'launch__registers_per_thread',
        'l1tex__data_bank_conflicts_pipe_lsu', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ']
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want):
        print(f"{h:80s}", [r[i] for r in rows[1:4]])
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
data = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
si = h.index('# Samples'); src = h.index('Source')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
agg = {h[i]: 0 for i in stall_cols}
for r in data:
    for i in stall_cols:
        try: agg[h[i]] += int(r[i])
        except Exception: pass
print('total samples', tot, 'instructions', len(data))
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
idx = sorted(range(len(data)), key=lambda j: -int(data[j][si]) if data[j][si].isdigit() else 0)[:n]
for j in idx:
    r = data[j]
    st = {h[i][6:]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(str(j).rjust(5), r[si].rjust(6), r[src][:64].ljust(64), st)
