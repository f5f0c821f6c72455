#!/bin/bash
# 2 GPUs: the torchrun bench path (strong scaling shard, coupled CG over NCCL, ScalarComm check) + the 2-GPU test
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02g_gpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_fullsize.py::test_pipelined_kernel_parity_subprocess -m gpu -q -s > $O/r02g_pytest_dist.log 2>&1; echo "dist test exit $?"; tail -3 $O/r02g_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02g_bench_n2.json 2> $O/r02g_bench_n2.err; echo "bench n2 exit $?"
tail -c 1500 $O/r02g_bench_n2.json; tail -5 $O/r02g_bench_n2.err
