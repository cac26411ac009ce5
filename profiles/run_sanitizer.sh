#!/bin/bash
# compute-sanitizer over the -m gpu parity suite at CI sizes (SURVEY section 5 / VERDICT r1 item 9).
#   bash profiles/run_sanitizer.sh [tool ...]        default: memcheck racecheck synccheck initcheck
# Logs go to gpurun_out/sanitizer_<tool>.log; a one-line summary per tool to gpurun_out/sanitizer_summary.txt.
# The full-size tests (1080p / 4K / 8K) are left out: under instrumentation they would take hours and exercise the
# same kernels as the CI-size cases.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck initcheck}
TESTS="tests/test_gpu_parity.py tests/test_gpu_decoder.py tests/test_rate_control.py tests/test_gpu_round2.py"
SEL='not beyond_resident_capacity and not cif and not 288'
: > gpurun_out/sanitizer_summary.txt
for tool in $TOOLS; do
    log=gpurun_out/sanitizer_${tool}.log
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    [ "$tool" = racecheck ] && extra="--racecheck-report all"
    t0=$(date +%s)
    timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 40 --error-exitcode 3 --log-file $log \
        python -m pytest $TESTS -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_${tool}.pytest.txt 2>&1
    rc=$?
    t1=$(date +%s)
    nerr=$(grep -c "^========= .*\(Error\|error\|hazard\|Hazard\|Invalid\|Uninitialized\)" $log 2>/dev/null || true)
    tailline=$(tail -n 1 gpurun_out/sanitizer_${tool}.pytest.txt | tr -d '\n')
    summ=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -n 1 | tr -d '\n')
    echo "$tool rc=$rc seconds=$((t1 - t0)) flagged_lines=$nerr | $summ | pytest: $tailline" | tee -a gpurun_out/sanitizer_summary.txt
done
