#!/bin/bash
# run every GPU test in its own process so that one CUDA fault cannot poison the others
cd "$(dirname "$0")/.."
out=gpurun_out/each.log; : > $out
for t in $(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::"); do
  if timeout 300 python -m pytest "$t" -q -s -x > gpurun_out/_one.log 2>&1; then echo "PASS $t" >> $out; else echo "FAIL $t" >> $out; grep -E "Error|error|assert|cg parity|   ref" gpurun_out/_one.log | head -45 >> $out; fi
done
grep -c PASS $out; grep FAIL $out
