#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 300 python tests/tools/kbench.py 128 4 > $O/r02i_kbench128.log 2>&1
timeout 300 python tests/tools/kbench.py 256 2 > $O/r02i_kbench256.log 2>&1
timeout 300 python tests/tools/kbench.py 64 8 > $O/r02i_kbench64.log 2>&1
timeout 300 python tests/tools/kbench.py 512 1 > $O/r02i_kbench512.log 2>&1
PTX_PIPE=1 timeout 120 python tools/pipe_probe.py 4 > $O/r02i_pipe.log 2>&1
for f in 64 128 256 512; do echo "== $f"; grep "API\|cg_\|CG (mine)" $O/r02i_kbench$f.log; done; cat $O/r02i_pipe.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r02i_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -4 $O/r02i_pytest_all.log
