#!/bin/bash
# Third profiling pass of round 2: the fused kernel variant that became the k = 20 default last (3 E-warps per S-warp).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python tools/profile_kernels.py fused > $O/r02c_profile_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/r02c_profile_plain.log; exit 1; }
cat $O/r02c_profile_plain.log
ncu --set full --clock-control none --import-source on -k regex:fused_wsd_kernel -s 1 -c 1 -f -o $O/r02c_ncu_fused python tools/profile_kernels.py fused > $O/r02c_ncu_fused.log 2>&1
echo "ncu fused rc=$?"
ncu -i $O/r02c_ncu_fused.ncu-rep --page raw --csv > $O/r02c_ncu_fused_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02c_ncu_fused.ncu-rep > $O/r02c_ncu_fused_summary.txt 2>&1
ncu -i $O/r02c_ncu_fused.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02c_ncu_fused_source.csv.gz
rm -f $O/r02c_ncu_fused.ncu-rep
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02c_bench_plain.json 2> $O/r02c_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02c_bench_under_ncu.log 2>&1
ls -la $O | grep r02c
