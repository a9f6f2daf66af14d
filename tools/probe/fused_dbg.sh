#!/bin/bash
# timing variants of the fused small-instance kernel (results wrong by construction except dbg=0): which resource bounds it?
for d in 0 1 2 3 4 5 6 7 8 16 18 22; do
  echo "== PMB_FUSED_DBG=$d"; PMB_FUSED_DBG=$d timeout -s KILL 60 tools/probe/fused_probe 8192 256 256 3 2>&1 | grep -E "sync" | tail -1
done
