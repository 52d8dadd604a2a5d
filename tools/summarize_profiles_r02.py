"""Turns gpurun_out/{launches_r02.csv, *_r02_raw.csv, bench_r2_*.json} (tools/evidence_r02.sh) into the committed summaries under
profiles/: launches_r02_summary.json, <kernel>_r02_summary.json, bench_lines_r02.json."""
import collections, csv, json, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
G = "gpurun_out"
rows = list(csv.reader(open(f'{G}/launches_r02.csv', errors='ignore')))
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
for o in out[:6]: print(o)
json.dump(dict(command="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ess --no-extras  (ncu --metrics gpu__time_duration.sum --clock-control none -c 400)",
               total_ms=tot, kernels=out), open('profiles/launches_r02_summary.json', 'w'), indent=1)
shutil.copy(f'{G}/launches_r02.csv', 'profiles/launches_r02.csv')

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'smsp__inst_executed.sum', 'dram__bytes_write.sum.per_second',
        'dram__bytes_read.sum.per_second', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed_op_shared_ld.sum', 'sm__sass_inst_executed_op_global_ld.sum']


def summarize(name, meta):
    p = f'{G}/{name}_raw.csv'
    if not os.path.exists(p):
        print("missing", p); return None
    rows = list(csv.reader(open(p, errors='ignore')))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, units, vals = rows[hi], rows[hi + 1], rows[hi + 2:]
    kn = hdr.index('Kernel Name')
    m = {h: dict(unit=units[i], values=[v[i] for v in vals if len(v) > i]) for i, h in enumerate(hdr) if h in WANT}
    json.dump(dict(**meta, kernels=[v[kn].split('(')[0] for v in vals if len(v) > kn], metrics=m), open(f'profiles/{name}_summary.json', 'w'), indent=1)
    for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
              'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
              'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
              'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum'):
        if k in m: print(' ', name, k, m[k]['unit'], m[k]['values'])
    return m


full = "ncu --set full --clock-control none -k regex:%s -s %d -c %d <cmd>; exported on the box with `ncu -i ... --page raw --csv`"
summarize('k1_full_r02', dict(command=full % ("k1_kernel", 25, 1), cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ess --no-extras",
                             kernel="mg::k1_kernel<3 (logistic), 13 (d<=104), 4, 8, 2, 2>", workload="cfg4 N=1e6 d=100 C=10000, interior leapfrog wave"))
summarize('k1_probit_r02', dict(command=full % ("k1_kernel", 8, 1), cmd="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-ess",
                               kernel="mg::k1_kernel<4 (probit), 3 (d<=24), 4, 8, 4 (ring), 2 (CTAs/SM)>", workload="cfg3 N=1e5 d=20 C=16384, MALA wave (value + gradient)"))
summarize('fused_full_r02', dict(command=full % ("fused_chain", 1, 1), cmd="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-ess",
                                kernel="mg::fused_chain_kernel<0 (normal_fn), 3, false>", workload="cfg2 65536 chains HMC(0.75), 400-step timed launch"))
summarize('k1_smalln_r02', dict(command=full % ("k1_kernel", 12, 1), cmd="python tools/smalln_probe.py 4 94720",
                               kernel="mg::k1_kernel<3, 13, ...> with the fused interior leapfrog", workload="smallN N=256 d=100 C=94720"))
summarize('transition_r02', dict(command=full % ("transition_coop", 7, 2), cmd="python tools/smalln_probe.py 4 94720",
                                kernel="mg::transition_coop_kernel (launch 1: interior wave, counters only; launch 2 or 1: decision wave)", workload="smallN N=256 d=100 C=94720"))
summarize('stats_r02', dict(command=full % ("stats_", 4, 4), cmd="python tools/stats_probe.py imse",
                           kernel="stats_mean_kernel, stats_var_kernel<IMSE>, stats_gather_kernel, stats_more_warp_kernel", workload="65536 chains x 3 parameters x 9000 kept draws"))
lines = {}
for tag, f in (("default_n1", "bench_r2_final_n1"), ("default_n2", "bench_r2_n2"), ("default_n4", "bench_r2_n4"), ("default_n8", "bench_r2_n8"), ("reference_n1", "bench_r2_ref")):
    pth = f'{G}/{f}.json'
    if not os.path.exists(pth): continue
    txt = open(pth).read().strip().splitlines()
    lines[tag] = json.loads([t for t in txt if t.startswith('{')][-1])
json.dump(lines, open('profiles/bench_lines_r02.json', 'w'), indent=1)
print("bench lines:", list(lines))
