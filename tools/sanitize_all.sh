#!/bin/bash
# compute-sanitizer memcheck over the small-model GPU tests (parity, export, result path, edge cases, golden fixtures, fuzz)
mkdir -p gpurun_out
{
echo '$ compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_export.py tests/test_gpu_result_path.py tests/test_gpu_edge.py tests/test_golden.py tests/test_gpu_fuzz.py -q -m gpu'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_export.py tests/test_gpu_result_path.py tests/test_gpu_edge.py tests/test_golden.py tests/test_gpu_fuzz.py -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -20
} > gpurun_out/r01o_sanitizer_all.txt 2>&1
cat gpurun_out/r01o_sanitizer_all.txt
