#!/bin/bash
# GPU run of the int8-slice Gram probe: small shapes first (a hang costs 60 s), then cfg2 / cfg3 shapes.
# args of i8_probe: n k batch ns(0 = adaptive) dist(0 uniform, 1 normal, 2 wide range) reps
mkdir -p gpurun_out
P=./tools/probe/i8_probe
{
timeout 60 $P 128 64 1 0 0 1; echo "exit $?"
timeout 60 $P 520 1000 2 0 1 2; echo "exit $?"
timeout 120 $P 4096 4096 1 0 0 5; echo "exit $?"
timeout 120 $P 4096 4096 1 0 1 3; echo "exit $?"
timeout 120 $P 4096 4096 1 6 0 3; echo "exit $?"
timeout 120 $P 256 256 1024 0 0 3; echo "exit $?"
timeout 120 $P 512 512 1 0 0 3; echo "exit $?"
timeout 60 $P 300 500 1 0 2 1; echo "exit $?"
} > gpurun_out/i8_probe.txt 2>&1
grep -E "^run|PROBE|exit|device" gpurun_out/i8_probe.txt
