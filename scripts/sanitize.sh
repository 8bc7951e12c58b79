#!/bin/bash
# compute-sanitizer memcheck on a small build + kernel parity run (one tool per gpurun call)
set -u
python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head
