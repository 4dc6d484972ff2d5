#!/bin/bash
# compute-sanitizer over the small unmold parity tests (the team kernel synchronises through
# named barriers, shared-memory atomics and bulk async copies).  Run on a GPU box:
#   bash tools/sanitize.sh   -> gpurun_out/sanitizer_<tool>.log
cd "$(dirname "$0")/.."
SEL="more_instances or every_box or no_detections or small_boxes or chunk_size or leading_unit"
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest tests/test_gpu_unmold.py -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
