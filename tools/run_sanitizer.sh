#!/bin/bash
# Runs on the GPU box: compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over tools/sanitize_target.py.
#   gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh'            single-GPU kernels
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/run_sanitizer.sh p2p' the peer-memory kernels under torchrun (memcheck + racecheck)
# Logs: gpurun_out/r02_sanitizer_<tool>[_p2p].log; the ERROR SUMMARY lines are echoed.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
MODE=${1:-single}
python tools/sanitize_target.py single > $O/r02_sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -20 $O/r02_sanitizer_plain.log; exit 1; }
tail -2 $O/r02_sanitizer_plain.log
if [ "$MODE" = "p2p" ]; then
    for tool in memcheck racecheck; do
        timeout 400 compute-sanitizer --tool $tool --target-processes all --log-file $O/r02_sanitizer_${tool}_p2p.%p.log \
            python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 \
            tools/sanitize_target.py p2p > $O/r02_sanitizer_${tool}_p2p.out 2>&1
        echo "$tool p2p rc=$?"; grep -h "ERROR SUMMARY\|sanitize target ok" $O/r02_sanitizer_${tool}_p2p.*
    done
    exit 0
fi
for tool in memcheck racecheck synccheck initcheck; do
    timeout 600 compute-sanitizer --tool $tool --log-file $O/r02_sanitizer_$tool.log python tools/sanitize_target.py single > $O/r02_sanitizer_$tool.out 2>&1
    echo "$tool rc=$?"; tail -1 $O/r02_sanitizer_$tool.out; grep -c "=========" $O/r02_sanitizer_$tool.log; grep "ERROR SUMMARY" $O/r02_sanitizer_$tool.log
done
