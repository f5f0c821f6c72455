#!/bin/bash
# round 2, GPU call 1: new full-size tests, CG parity A/B, persisting-L2 A/B, first bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/r02a_gpu.txt 2>&1
free -g >> $O/r02a_gpu.txt; nproc >> $O/r02a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_fullsize.py > $O/r02a_pytest_old.log 2>&1; echo "old tests exit $?" 
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > $O/r02a_pytest_fullsize.log 2>&1; echo "fullsize exit $?"
timeout 600 python tests/tools/cg_parity_probe.py > $O/r02a_cg_probe_sfu.log 2>&1; echo "probe sfu exit $?"
PTYCHOFFT_B200_LIB=$PWD/libtike-cufft_b200/libtike/cufft/libptychofft_b200_ieee.so timeout 600 python tests/tools/cg_parity_probe.py > $O/r02a_cg_probe_ieee.log 2>&1; echo "probe ieee exit $?"
timeout 300 python tools/l2_probe.py 256 2 > $O/r02a_l2_on.log 2>&1; echo "l2 on exit $?"
PTX_L2_PERSIST=0 timeout 300 python tools/l2_probe.py 256 2 > $O/r02a_l2_off.log 2>&1; echo "l2 off exit $?"
timeout 300 python tools/l2_probe.py 512 1 > $O/r02a_l2_on512.log 2>&1
PTX_L2_PERSIST=0 timeout 300 python tools/l2_probe.py 512 1 > $O/r02a_l2_off512.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02a_bench_ref.json 2> $O/r02a_bench_ref.err; echo "bench ref exit $?"
# DRAM traffic of the 256^2 kernels with / without the window (one ncu tool per call)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'k_grad|k_adj|k_fwd|k_intensity' -c 70 --csv --log-file $O/r02a_l2_on_ncu.csv python tools/l2_probe.py 256 2 > $O/r02a_ncu_on.log 2>&1; echo "ncu on exit $?"
PTX_L2_PERSIST=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'k_grad|k_adj|k_fwd|k_intensity' -c 70 --csv --log-file $O/r02a_l2_off_ncu.csv python tools/l2_probe.py 256 2 > $O/r02a_ncu_off.log 2>&1; echo "ncu off exit $?"
tail -3 $O/r02a_pytest_old.log $O/r02a_pytest_fullsize.log
