#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02n_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -3 $O/r02n_pytest_all.log
for n in 128 256; do echo "== ndet=$n"; timeout 300 python tests/tools/kbench.py $n $((512/n)) 2>&1 | grep "cg_\|CG (mine)"; done > $O/r02n_kbench.log 2>&1
cat $O/r02n_kbench.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/r02n_bench_ref.json 2> $O/r02n_bench_ref.err; echo "bench ref exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02n_bench.json 2> $O/r02n_bench.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/r02n_bench.json", "gpurun_out/r02n_bench_ref.json"):
    d = json.load(open(f))
    print(f, d["value"], d["e2e"]["value"], d["e2e"].get("pageable"), d["e2e_cg"]["value"], d.get("cg", {}).get("iters_per_s"))
PY
