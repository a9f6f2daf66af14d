#!/bin/bash
# what the driver runs at round end, in one call: GPU suite, smoke, bench (both arms)
T=${1:-round}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; cut -c1-160 gpurun_out/${T}_ref.json
timeout -s KILL 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-260 gpurun_out/${T}_bench.json
