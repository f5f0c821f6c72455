#!/bin/bash
# round 2, GPU call 2: pipelined 128^2 kernel (parity + timing A/B), f64 referee, regression check of CG rates
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -s -x -k "operators or grad_ptycho or f64 or plugin or bindings" > $O/r02b_pytest_a.log 2>&1; echo "pytest A exit $?"
tail -3 $O/r02b_pytest_a.log
timeout 300 python tests/tools/kbench.py 128 4 > $O/r02b_kbench128_pipe.log 2>&1; echo "kbench128 pipe exit $?"
PTX_PIPE=0 timeout 300 python tests/tools/kbench.py 128 4 > $O/r02b_kbench128_nopipe.log 2>&1; echo "kbench128 nopipe exit $?"
timeout 300 python tests/tools/kbench.py 256 2 > $O/r02b_kbench256.log 2>&1; echo "kbench256 exit $?"
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r02b_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -5 $O/r02b_pytest_all.log
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-extras > $O/r02b_bench_c2.json 2> $O/r02b_bench_c2.err; echo "bench c2 exit $?"
