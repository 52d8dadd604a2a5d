# Round-2 evidence run (one gpurun call): launch list of the headline bench command + one `ncu --set full` capture per hot kernel,
# each only after the same command exited 0 without ncu.  The .ncu-rep files are exported to raw CSV on the box and removed
# (gpurun brings back at most 64 MiB); tools/summarize_profiles_r02.py turns the CSVs into profiles/*_r02_summary.json.
set -x
O=gpurun_out
full() {  # full <name> <kernel regex> <skip> <count> <command...>
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none -k regex:$rx -s $skip -c $cnt -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ess --no-extras"
$B > $O/plain_r02.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02.csv $B > /dev/null 2>&1
full k1_full_r02 k1_kernel 25 1 $B
B3="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-ess"
$B3 > $O/plain_cfg3_r02.log 2>&1 && full k1_probit_r02 k1_kernel 8 1 $B3
B2="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-ess"
$B2 > $O/plain_cfg2_r02.log 2>&1 && full fused_full_r02 fused_chain 1 1 $B2
python tools/smalln_probe.py 4 94720 > $O/plain_smalln_r02.log 2>&1 && full transition_r02 transition_coop 7 2 python tools/smalln_probe.py 4 94720
full k1_smalln_r02 k1_kernel 12 1 python tools/smalln_probe.py 4 94720
python tools/stats_probe.py imse > $O/plain_stats_r02.log 2>&1 && full stats_r02 stats_ 4 4 python tools/stats_probe.py imse
ls -la $O
