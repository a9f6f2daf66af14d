#!/bin/bash
# compute-sanitizer over the int8-route kernels incl. the CTA-pair and split-ring variants (subprocess test => --target-processes all)
mkdir -p gpurun_out
{
echo '$ compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 python -m pytest tests/test_gpu_gram_i8.py -x -q -k "not cfg2 and not pipelined"'
timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 python -m pytest tests/test_gpu_gram_i8.py -x -q -k "not cfg2 and not pipelined" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | sort | uniq -c | head -20
echo '$ compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_gram_i8.py -x -q -k "forced and (128 or 200)"'
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_gram_i8.py -x -q -k "forced and (128 or 200)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head
echo '$ PMB_I8_PAIR=1 compute-sanitizer --tool memcheck --error-exitcode 7 tools/probe/i8_probe 520 1000 2 5 1 1 ; ... 300 500 1 6 0 1'
PMB_I8_PAIR=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 tools/probe/i8_probe 520 1000 2 5 1 1 2>&1 | grep -E "PROBE|ERROR SUMMARY"
PMB_I8_PAIR=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 tools/probe/i8_probe 300 500 1 6 0 1 2>&1 | grep -E "PROBE|ERROR SUMMARY"
echo '$ PMB_I8_KC=128 compute-sanitizer --tool memcheck --error-exitcode 7 tools/probe/i8_probe 520 1000 2 5 1 1'
PMB_I8_KC=128 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 tools/probe/i8_probe 520 1000 2 5 1 1 2>&1 | grep -E "PROBE|ERROR SUMMARY"
} > gpurun_out/r01o_sanitizer.txt 2>&1
cat gpurun_out/r01o_sanitizer.txt
