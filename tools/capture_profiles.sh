#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + one ncu --set full capture per
# kernel family.  Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into profiles/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 3 --no-cg > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --no-cg > gpurun_out/ncu_launch.log 2>&1
python tools/prof.py 128 4 > gpurun_out/plain_prof128.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_fwd|k_adj|k_intensity|k_grad|k_linesearch" \
    -s 7 -c 7 -f -o /tmp/all128 python tools/prof.py 128 4 > gpurun_out/ncu_all128.log 2>&1
ncu -i /tmp/all128.ncu-rep --page raw --csv > gpurun_out/all128_raw.csv
ncu -i /tmp/all128.ncu-rep --page details > gpurun_out/all128_details.txt
python tools/prof.py 256 1 > gpurun_out/plain_prof256.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_grad|k_intensity" \
    -s 3 -c 3 -f -o /tmp/all256 python tools/prof.py 256 1 > gpurun_out/ncu_all256.log 2>&1
ncu -i /tmp/all256.ncu-rep --page raw --csv > gpurun_out/all256_raw.csv
ncu -i /tmp/all256.ncu-rep --page details > gpurun_out/all256_details.txt
python bench.py --steps 3 --no-cg > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_grad" -s 4 -c 1 -f -o /tmp/bench_grad \
    python bench.py --steps 3 --no-cg > gpurun_out/ncu_bench_grad.log 2>&1
ncu -i /tmp/bench_grad.ncu-rep --page raw --csv > gpurun_out/bench_grad_raw.csv
ncu -i /tmp/bench_grad.ncu-rep --page details > gpurun_out/bench_grad_details.txt
ls -la gpurun_out
