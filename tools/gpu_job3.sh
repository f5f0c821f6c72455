#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
for d in 0 1 2 3 4 5 6 7; do PTX_PIPE_DEBUG=$d timeout 120 python tools/pipe_probe.py 4; done > $O/r02c_pipe_ablation.log 2>&1
PTX_PIPE=0 timeout 120 python tools/pipe_probe.py 4 >> $O/r02c_pipe_ablation.log 2>&1
cat $O/r02c_pipe_ablation.log
PTX_PIPE=0 timeout 400 python tests/tools/cg_fullsize_probe.py > $O/r02c_c5128_nopipe.log 2>&1
timeout 400 python tests/tools/cg_fullsize_probe.py > $O/r02c_c5128_pipe.log 2>&1
grep -v "^#\|iteration\|^ *0," $O/r02c_c5128_nopipe.log $O/r02c_c5128_pipe.log
