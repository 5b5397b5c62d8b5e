#!/bin/bash
# compute-sanitizer over the small-shape workload in tools/sanitize_driver.py (run on the GPU box):
#   memcheck  - out-of-bounds / misaligned global, shared (incl. DSMEM) and TMEM-adjacent accesses
#   racecheck - shared-memory hazards inside a CTA (the mbarrier / TMA / tcgen05 pipeline of conv_umma_kernel,
#               the staged tiles of the Activation1d kernels)
#   synccheck - invalid barrier usage (bar.sync 1,128 of the epilogue warps, cluster barriers)
# Logs go to gpurun_out/sanitize_<tool>.log; the last lines ("ERROR SUMMARY") are what profiles/r2_sanitize.txt quotes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
rc=0
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "== compute-sanitizer --tool $tool"
  timeout ${SAN_TIMEOUT:-1500} "$SAN" --tool "$tool" --error-exitcode 17 --print-limit 20 \
      python tools/sanitize_driver.py > "gpurun_out/sanitize_$tool.log" 2>&1
  code=$?
  tail -n 4 "gpurun_out/sanitize_$tool.log"
  echo "== $tool exit code $code"
  [ $code -ne 0 ] && rc=1
done
exit $rc
