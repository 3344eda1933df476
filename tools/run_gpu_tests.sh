#!/bin/bash
# Runs every test FUNCTION of the given test files in its own process (a trapped kernel poisons the CUDA context of
# its process only), each under a timeout; logs to gpurun_out/t_<file>_<function>.log and prints a summary.
# usage: tools/run_gpu_tests.sh tests/test_gpu_16bit.py [more files ...]
mkdir -p gpurun_out
summary=gpurun_out/gpu_tests_summary.txt
: > $summary
for f in "$@"; do
  base=$(basename $f .py)
  for fn in $(grep -oE "^def (test_[a-zA-Z0-9_]+)" $f | awk '{print $2}'); do
    log=gpurun_out/t_${base}_${fn}.log
    timeout ${TEST_TIMEOUT:-900} python -m pytest "$f" -q -m gpu -k "$fn" -s -p no:cacheprovider > $log 2>&1
    rc=$?
    echo "$rc $base::$fn :: $(tail -1 $log)" | tee -a $summary
  done
done
