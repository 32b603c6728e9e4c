#!/bin/bash
# Runs on the GPU box (gpurun): plain run of every profiled command first, then the ncu captures (profiles/README.md).
# The .ncu-rep files are condensed on the box (raw-page CSV + tools/ncu_summary.py); only the fused and export reports
# travel back (gpurun_out/ is capped at 64 MiB).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
K='regex:fused_wsd_kernel|sample_flat|eval_values|gram_mma|gram_kernel|finalize|scatter'
python tools/profile_kernels.py all > $O/r02_profile_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/r02_profile_plain.log; exit 1; }
cat $O/r02_profile_plain.log
ncu --set full --clock-control none --import-source on -k regex:fused_wsd_kernel -s 1 -c 1 -f -o $O/r02_ncu_fused_k20_n24 python tools/profile_kernels.py fused > $O/r02_ncu_fused.log 2>&1
echo "ncu fused rc=$?"
for w in ishigami export evalpf rk4 gram20 gramreg; do
    ncu --set full --clock-control none --import-source on -k "$K" -c 6 -f -o $O/r02_ncu_$w python tools/profile_kernels.py $w > $O/r02_ncu_$w.log 2>&1
    echo "ncu $w rc=$?"
done
for w in fused_k20_n24 ishigami export evalpf rk4 gram20 gramreg; do
    ncu -i $O/r02_ncu_$w.ncu-rep --page raw --csv > $O/r02_ncu_${w}_raw.csv 2>/dev/null
    python tools/ncu_summary.py $O/r02_ncu_$w.ncu-rep > $O/r02_ncu_${w}_summary.txt 2>&1
done
ncu -i $O/r02_ncu_fused_k20_n24.ncu-rep --page source --csv > $O/r02_ncu_fused_k20_n24_source.csv 2>/dev/null
ncu -i $O/r02_ncu_evalpf.ncu-rep --page source --csv > $O/r02_ncu_evalpf_source.csv 2>/dev/null
ncu -i $O/r02_ncu_export.ncu-rep --page source --csv > $O/r02_ncu_export_source.csv 2>/dev/null
rm -f $O/r02_ncu_ishigami.ncu-rep $O/r02_ncu_rk4.ncu-rep $O/r02_ncu_gram20.ncu-rep $O/r02_ncu_gramreg.ncu-rep $O/r02_ncu_evalpf.ncu-rep
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_bench_plain.json 2> $O/r02_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_bench_under_ncu.log 2>&1
gzip -f $O/r02_ncu_*_source.csv
du -sh $O; ls -la $O | tail -30
