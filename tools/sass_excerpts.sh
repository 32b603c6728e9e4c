#!/bin/bash
# SASS evidence for profiles/: per kernel of the in-tree library, how many tensor-path (DMMA), bulk-copy (UBLKCP), mbarrier
# (SYNCS) and FP64 instructions the sm_100a code object holds, plus a few lines of each kind in context.  No GPU needed.
#   bash tools/sass_excerpts.sh > profiles/r02_sass_excerpts.txt
cd "$(dirname "$0")/.."
SO=varsens_b200/libvarsens_b200.so
echo "# cuobjdump -sass $SO   ($(cuobjdump -lelf $SO | head -1))"
cuobjdump -sass $SO 2>/dev/null | awk '
  /Function : /{ if (name != "") printf "%-110s DMMA %5d  UBLKCP %3d  SYNCS %4d  DFMA %5d  DADD %5d  DMUL %5d  LDS %5d  total %6d\n", name, c["DMMA"], c["UBLKCP"], c["SYNCS"], c["DFMA"], c["DADD"], c["DMUL"], c["LDS"], tot;
                 name = $3; delete c; tot = 0; next }
  /^[ \t]+\/\*[0-9a-f]+\*\// { tot++; for (m in want) if (index($0, m)) c[m]++ }
  BEGIN { want["DMMA"]; want["UBLKCP"]; want["SYNCS"]; want["DFMA"]; want["DADD"]; want["DMUL"]; want["LDS"] }
  END { printf "%-110s DMMA %5d  UBLKCP %3d  SYNCS %4d  DFMA %5d  DADD %5d  DMUL %5d  LDS %5d  total %6d\n", name, c["DMMA"], c["UBLKCP"], c["SYNCS"], c["DFMA"], c["DADD"], c["DMUL"], c["LDS"], tot }
' | sed 's/_ZN2vs//' | c++filt 2>/dev/null | sort -k3,3nr | cut -c1-230 > /tmp/sass_table.txt
grep -E "fused_wsd_kernelILi20|sample_flat_bulk|gram_mma_kernelILi6ELb0ELb0ELb0|gram_mma_kernelILi2ELb1|eval_values_pf|eval_values_kernelINS_8RK4ChainILi10|p2p_reduce|finalize_kernel|fused_wsd_kernelILi3ENS_11Ishigami" /tmp/sass_table.txt
echo
echo "# arch of the embedded code objects"
cuobjdump -lelf $SO | sed 's/^/  /' | sort | uniq -c | head -5
for pat in "DMMA" "UBLKCP" "SYNCS.ARRIVE" "SYNCS.PHASECHK" "UTMA\|UBLKCP.S.G\|UBLKCP.G.S"; do
  echo; echo "# first occurrences of $pat (any kernel)"
  cuobjdump -sass $SO 2>/dev/null | grep -E "^\s+/\*[0-9a-f]+\*/" | grep -m 4 "$pat" | sed 's/\/\* 0x[0-9a-f]* \*\///' | cut -c1-130
done
