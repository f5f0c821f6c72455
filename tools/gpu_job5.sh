#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 120 python tools/pipe_probe.py 4 > $O/r02e_pipe.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grad_pipe -s 3 -c 1 -o $O/r02e_pipe python tools/pipe_probe.py 4 > $O/r02e_ncu.log 2>&1
echo "ncu exit $?"; cat $O/r02e_pipe.log; tail -3 $O/r02e_ncu.log
