"""Turns gpurun_out/{launches_r01.csv, *_full_r01.ncu-rep, bench_*.json} into the committed summaries under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
rows = list(csv.reader(open('gpurun_out/launches_r01.csv', errors='ignore')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0]); durs = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[col['Metric Name']] != 'gpu__time_duration.sum': continue
    name = r[col['Kernel Name']].split('(')[0]
    v = float(r[col['Metric Value']].replace(',', '')); unit = r[col['Metric Unit']]
    ms = v / 1e6 if unit in ('ns', 'nsecond') else (v / 1e3 if unit in ('us', 'usecond') else v)
    agg[name][0] += 1; agg[name][1] += ms; durs[name].append(ms)
tot = sum(v[1] for v in agg.values())
out = [dict(kernel=k, launches=v[0], total_ms=round(v[1], 3), share=round(v[1] / tot, 4), avg_ms=round(v[1] / v[0], 4),
            min_ms=round(min(durs[k]), 4), max_ms=round(max(durs[k]), 4)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
for o in out[:5]: print(o)
json.dump(dict(command="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ess  (ncu --metrics gpu__time_duration.sum --clock-control none -c 400)",
               total_ms=tot, kernels=out), open('profiles/launches_r01_summary.json', 'w'), indent=1)
shutil.copy('gpurun_out/launches_r01.csv', 'profiles/launches_r01.csv')
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__inst_executed.sum', 'dram__bytes_write.sum.per_second']
def summarize(rep, outp, meta):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    m = {h: dict(unit=units[i], values=[v[i] for v in vals]) for i, h in enumerate(hdr) if h in WANT}
    json.dump(dict(**meta, metrics=m), open(outp, 'w'), indent=1)
    return m
m = summarize('gpurun_out/k1_full_r01.ncu-rep', 'profiles/k1_full_r01_summary.json',
              dict(command="ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 25 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ess",
                   kernel="mg::k1_kernel<3 (logistic), 13 (d<=104), 4>", workload="cfg4 N=1e6 d=100 C=10000"))
for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
          'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread'): print(k, m[k])
if os.path.exists('gpurun_out/fused_full_r01.ncu-rep'):
    summarize('gpurun_out/fused_full_r01.ncu-rep', 'profiles/fused_full_r01_summary.json',
              dict(command="ncu --set full --clock-control none --import-source on -k regex:fused_chain -c 1 python bench.py --workload cfg2 --steps 2 --no-cpu-baseline --no-ess",
                   kernel="mg::fused_chain_kernel<0 (normal_fn), 3>", workload="cfg2 65536 chains HMC(0.75), first launch = the 600-step warm-up run"))
lines = {}
for tag, f in (("cfg4_n1", "bench_cfg4"), ("cfg4_n2", "bench_n2"), ("cfg4_n4", "bench_n4"), ("cfg4_n8", "bench_n8"), ("cfg5_n1", "bench_cfg5"),
               ("cfg5_n2", "bench_cfg5_n2"), ("cfg5_n8", "bench_cfg5_n8"), ("cfg3_n1", "bench_cfg3"), ("cfg2_n1", "bench_cfg2"), ("reference_cfg4", "bench_ref")):
    pth = f'gpurun_out/{f}.json'
    if not os.path.exists(pth): continue
    txt = open(pth).read().strip().splitlines()
    l = json.loads([t for t in txt if t.startswith('{')][-1])
    lines[tag] = l
    rf = l.get('roofline', {})
    print(tag, l['n_gpus'], 'value=%.5g' % l['value'], 'e2e=%.5g' % l['e2e']['value'], 'frac=%s' % (round(rf['frac'], 3) if rf else None),
          'ms/launch=%s' % (round(rf.get('ms_per_launch', 0), 2) if rf else None), 'ess=%s' % l.get('min_ess_per_s', {}).get('value'))
json.dump(lines, open('profiles/bench_lines_r01.json', 'w'), indent=1)
