#!/bin/bash
# 8 GPUs: the strong-scaling line the driver will ask for at round end
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02j_gpus.txt; nproc >> $O/r02j_gpus.txt; nvidia-smi topo -m >> $O/r02j_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02j_bench_n8.json 2> $O/r02j_bench_n8.err; echo "bench n8 exit $?"
tail -c 600 $O/r02j_bench_n8.json; tail -5 $O/r02j_bench_n8.err
