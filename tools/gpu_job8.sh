#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
echo "== default TMA policy" > $O/r02h_tma256.log
timeout 200 python tools/l2_probe.py 256 2 >> $O/r02h_tma256.log 2>&1
echo "== PTX_TMA_GATHER=all" >> $O/r02h_tma256.log
PTX_TMA_GATHER=all timeout 200 python tools/l2_probe.py 256 2 >> $O/r02h_tma256.log 2>&1
echo "== PTX_TMA_GATHER=off" >> $O/r02h_tma256.log
PTX_TMA_GATHER=off timeout 200 python tools/l2_probe.py 256 2 >> $O/r02h_tma256.log 2>&1
cat $O/r02h_tma256.log
