#!/bin/bash
# Second profiling pass of round 2 (after the export / two-phase / general-Gram work): the kernels that changed since
# tools/run_profiles.sh ran.  Same recipe: plain run first, then ncu --set full, condensed on the box.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
K='regex:sample_flat|eval_values|gram_mma|gram_kernel|scatter'
for w in export evalpf gramreg; do python tools/profile_kernels.py $w; done > $O/r02b_profile_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/r02b_profile_plain.log; exit 1; }
python tools/export_trace.py w16 full >> $O/r02b_profile_plain.log 2>&1
cat $O/r02b_profile_plain.log
for w in export evalpf gramreg; do
    ncu --set full --clock-control none --import-source on -k "$K" -c 6 -f -o $O/r02b_ncu_$w python tools/profile_kernels.py $w > $O/r02b_ncu_$w.log 2>&1
    echo "ncu $w rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:sample_flat_bulk -s 1 -c 1 -f -o $O/r02b_ncu_w16 python tools/export_trace.py w16 > $O/r02b_ncu_w16.log 2>&1
echo "ncu w16 rc=$?"
for w in export evalpf gramreg w16; do
    ncu -i $O/r02b_ncu_$w.ncu-rep --page raw --csv > $O/r02b_ncu_${w}_raw.csv 2>/dev/null
    ncu -i $O/r02b_ncu_$w.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02b_ncu_${w}_source.csv.gz
done
rm -f $O/r02b_ncu_*.ncu-rep
ls -la $O | grep r02b
