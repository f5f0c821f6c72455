#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 120 python tools/pipe_probe.py 4 > $O/r02d_pipe.log 2>&1
PTX_PIPE=0 timeout 120 python tools/pipe_probe.py 4 >> $O/r02d_pipe.log 2>&1
cat $O/r02d_pipe.log
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r02d_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -5 $O/r02d_pytest_all.log
timeout 600 python tests/tools/cg_error_curves.py > $O/r02d_cg_error_curves.log 2>&1; echo "curves exit $?"
cat $O/r02d_cg_error_curves.log
timeout 300 python tests/tools/kbench.py 128 4 > $O/r02d_kbench128.log 2>&1
grep "CG (mine)\|cg_grad" $O/r02d_kbench128.log
