#!/bin/bash
# compute-sanitizer over one small pass through every kernel family.  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun -- 'bash tools/sanitize.sh memcheck'   (then racecheck, synccheck in separate calls)
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 200 python tools/sanitize_target.py > gpurun_out/r02_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_target.py > gpurun_out/r02_sanitize_$TOOL.log 2>&1
echo "exit $?"; tail -n 15 gpurun_out/r02_sanitize_$TOOL.log
