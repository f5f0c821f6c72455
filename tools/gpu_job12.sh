#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
for s in 1 0; do
  for n in 128 64; do
    echo "== PTX_STRIP=$s ndet=$n"; PTX_STRIP=$s timeout 300 python tests/tools/kbench.py $n $((512/n)) 2>&1 | grep "API\|cg_\|CG (mine)"
  done
done > $O/r02l_strip.log 2>&1
cat $O/r02l_strip.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r02l_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -4 $O/r02l_pytest_all.log
