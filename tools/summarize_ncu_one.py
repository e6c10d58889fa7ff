"""development aid: summarise ONE ncu --set full report of a step kernel (run here, no GPU needed).
usage: python tools/summarize_ncu_one.py REPORT.ncu-rep OUT.md "title" cells nvar threads_per_cell [traffic_key]"""
import csv, json, re, subprocess, sys, os
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
rep, out, title, cells, nvar, per = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
tkey = sys.argv[7] if len(sys.argv) > 7 else None
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] * 1e9 if os.path.exists('MEASURED_PEAKS.json') else 6.45e12
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__block_size', 'launch__grid_size', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, r = rows[0], rows[1], rows[2]
md = [f'## {title}\n', f'Kernel: `{r[hdr.index("Kernel Name")]}`\n', '| metric | value |\n|---|---|']
vals = {}
for k in keys:
    if k in hdr:
        i = hdr.index(k); md.append(f'| `{k}` | {r[i]} {units[i]} |'); vals[k] = (float(r[i].replace(",", "")), units[i])
sc = lambda k: vals[k][0] * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[vals[k][1]]
traffic = sc('dram__bytes_read.sum') + sc('dram__bytes_write.sum')
fp64 = vals['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'][0]
dur = vals['gpu__time_duration.sum'][0] * {'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'ns': 1e-9}[vals['gpu__time_duration.sum'][1]]
dramp = 100 * traffic / dur / peak
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, x in enumerate(rows) if x and x[0] == 'Address'][0]; h = rows[hi]
data = [x for x in rows[hi + 1:] if len(x) == len(h) and x[0].startswith('0x')]
scols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
tot = Counter()
for x in data:
    for i in scols: tot[h[i]] += int(x[i] or 0)
s = sum(tot.values())
ie, si = h.index('Instructions Executed'), h.index('Source')
mix = Counter()
for x in data:
    op = re.sub(r'^@!?U?P\w+\s+', '', x[si].strip()).split()[0].split('.')[0]; mix[op] += int(x[ie] or 0)
t2 = sum(mix.values()); fp = sum(mix[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
nthreads = cells * per
md.append(f"\nWarp-stall sampling (all samples, %): { {k[6:]: round(100 * v / s, 1) for k, v in tot.most_common(8)} }\n")
md.append(f"Executed instruction mix (% of warp instructions): { {k: round(100 * v / t2, 1) for k, v in mix.most_common(12)} }\n")
md.append(f"Per cell-update: {t2 / (nthreads / 32) * per:.0f} instructions, of which {fp / (nthreads / 32) * per:.0f} FP64 (DFMA+DMUL+DADD+DSETP); DRAM traffic {traffic / cells:.0f} B per cell-update vs {2 * nvar * 8} B algorithmic = {dramp:.0f} % of the measured {peak / 1e12:.2f} TB/s copy peak under ncu; FP64 pipe {fp64:.1f} %.\n")
open(out, 'a').write('\n'.join(md) + '\n')
if tkey:
    tp = 'profiles/step_kernel_traffic.json'
    tj = json.load(open(tp))
    tj[tkey] = {'dram_bytes_per_launch': traffic, 'fp64_pipe_pct_of_peak': round(fp64, 1), 'dram_pct_of_measured_copy_peak': round(dramp, 1),
                'registers_per_thread': int(vals['launch__registers_per_thread'][0]), 'warps_per_sm': 16, 'source': out,
                'fp64_inst_per_cell_update': round(fp / (nthreads / 32) * per, 1), 'inst_per_cell_update': round(t2 / (nthreads / 32) * per, 1)}
    json.dump(tj, open(tp, 'w'), indent=1)
print(traffic, fp64, dramp, dict(mix.most_common(14)))
