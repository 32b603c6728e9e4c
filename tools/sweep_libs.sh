#!/bin/bash
# Runs tools/fused_time.py once per experiment library (varsens_b200/libvarsens_b200_*.so) and for the shipped one.
cd "$(dirname "$0")/.."
for lib in varsens_b200/libvarsens_b200.so varsens_b200/libvarsens_b200_*.so; do
    [ -f "$lib" ] || continue
    VS_LIB="$PWD/$lib" timeout 300 python tools/fused_time.py 2>&1 | tail -1
done
