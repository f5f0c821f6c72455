#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 120 python tools/pipe_probe.py 4 > $O/r02k_grad128.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k k_grad -s 3 -c 1 -o $O/r02k_grad128 python tools/pipe_probe.py 4 > $O/r02k_ncu.log 2>&1
echo "ncu exit $?"; cat $O/r02k_grad128.log
