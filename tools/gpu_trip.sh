#!/usr/bin/env bash
# One GPU-box visit: isolated parity tests, smoke, a short bench, then (only if the plain runs
# exited 0) the ncu launch list and one full-set capture of the main kernels.
# Usage (under gpurun):  bash tools/gpu_trip.sh [tests|bench|profile ...]   (default: all)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
WHAT="${*:-tests smoke bench profile}"
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1

if [[ " $WHAT " == *" tests "* ]]; then
  : > $OUT/tests_summary.txt
  # probe experiments one per process (a bad descriptor poisons the CUDA context); PROBE_FILTER narrows them
  for id in $(python -m pytest tests/test_umma_probe.py --collect-only -q -m gpu -k "${PROBE_FILTER:-test_}" 2>/dev/null | grep "::"); do
    timeout 300 python -m pytest "$id" -q -m gpu -x -p no:cacheprovider > $OUT/t.log 2>&1
    rc=$?
    echo "$rc $id" >> $OUT/tests_summary.txt
    if [ $rc -ne 0 ]; then { echo "=== $id"; tail -40 $OUT/t.log; } >> $OUT/tests_failures.txt; fi
  done
  if [ "${ISOLATE_MODEL:-0}" = "1" ]; then
    for id in $(python -m pytest tests/test_gpu_model.py --collect-only -q -m gpu 2>/dev/null | grep "::"); do
      timeout 300 python -m pytest "$id" -q -m gpu -x -p no:cacheprovider > $OUT/t.log 2>&1
      rc=$?
      echo "$rc $id" >> $OUT/tests_summary.txt
      if [ $rc -ne 0 ]; then { echo "=== $id"; grep -v "^$" $OUT/t.log | tail -45; } >> $OUT/tests_failures.txt; fi
    done
    MODEL_FILES=""
  else
    MODEL_FILES="tests/test_gpu_model.py"
  fi
  for f in tests/test_gpu_counts.py tests/test_gpu_preprocess.py tests/test_gpu_preprocess_tv.py tests/test_umma_i8_probe.py $MODEL_FILES; do
    timeout 900 python -m pytest "$f" -q -m gpu -p no:cacheprovider > $OUT/$(basename $f .py).log 2>&1
    echo "$? $f" >> $OUT/tests_summary.txt
    tail -60 $OUT/$(basename $f .py).log >> $OUT/tests_failures.txt
  done
  cat $OUT/tests_summary.txt
fi

if [[ " $WHAT " == *" smoke "* ]]; then
  timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
  echo "smoke rc=$?"; tail -5 $OUT/smoke.log
fi

if [[ " $WHAT " == *" bench "* ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.log 2> $OUT/bench.err
  echo "bench rc=$?"; tail -3 $OUT/bench.log; tail -5 $OUT/bench.err
fi

if [[ " $WHAT " == *" profile "* ]]; then
  timeout 600 python tools/profile_target.py 256 2 > $OUT/profile_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/launches.csv python tools/profile_target.py 256 2 > $OUT/ncu_launches.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_bench.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k "regex:^(preprocess_mma_kernel|preprocess_tc_kernel|preprocess_tc2_kernel|preprocess_kernel|preprocess_tv_fast_kernel|conv1_kernel|conv3x3_pair_kernel|conv3x3_kernel|conv3x3_stream_kernel|linear_splitk_kernel|head_tail_cluster_kernel|head_tail_kernel)$" -c 9 \
      -f -o $OUT/prof python tools/profile_target.py 256 1 > $OUT/ncu_full.log 2>&1
  echo "profile rc=$?"; tail -3 $OUT/profile_plain.log
fi
echo trip-done
