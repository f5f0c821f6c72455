#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 120 python tools/pipe_probe.py 4 > $O/r02f_pipe.log 2>&1; e1=$?
PTX_PIPE=0 timeout 120 python tools/pipe_probe.py 4 >> $O/r02f_pipe.log 2>&1
cat $O/r02f_pipe.log
timeout 900 python -m pytest tests -m gpu -q -x -k "grad_ptycho or cg_vs_reference or full_size or golden or skipped" > $O/r02f_pytest.log 2>&1; echo "pytest exit $?"
tail -4 $O/r02f_pytest.log
[ $e1 -eq 0 ] && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grad_pipe -s 3 -c 1 -o $O/r02f_pipe python tools/pipe_probe.py 4 > $O/r02f_ncu.log 2>&1
echo "ncu exit $?"
