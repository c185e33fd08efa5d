#!/bin/sh
# compute-sanitizer target: memcheck / racecheck / synccheck over a slice of the GPU
# parity tests -- every kernel family through the C ABI on small shapes.  Run on a GPU box:
#     gpurun --timeout 1500 -- sh tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# The resident kernel is switched off for racecheck/synccheck runs that would otherwise only see one
# long-lived launch (FN_NO_RESIDENT=1 evaluates with one launch per evaluation: the same kernels).
TOOL=${1:-memcheck}
EXPR=${2:-"golden_forced_param or bh_bit_exact or lane_mapping or resident_matches"}
mkdir -p gpurun_out
[ "$TOOL" = memcheck ] || export FN_NO_RESIDENT=1
compute-sanitizer --tool "$TOOL" --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -m gpu -x -q -k "$EXPR" \
    > "gpurun_out/sanitize_$TOOL.log" 2>&1
RC=$?
tail -5 "gpurun_out/sanitize_$TOOL.log"
echo "compute-sanitizer $TOOL exit code $RC"
exit $RC
