#!/bin/bash
T=r02x
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
