#!/bin/bash
for tool in memcheck racecheck synccheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_r02.py > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run done" gpurun_out/r02_sanitize_$tool.log | tail -3
done
