#!/bin/bash
# Run on the GPU box: compute-sanitizer over the smoke sequence (3 frames, every kernel of the frame path) and the stage tests.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/san_$tool.log | tr '\n' ' ')"
done
