#!/bin/bash
# one gpurun job: launch list, ncu --set full captures of the VGICP and GICP kernels, full bench line (outputs under gpurun_out/)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r02_launches_bench.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:factor_kernel -s 2 -c 1 -f -o gpurun_out/r02_vgicp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r02_vgicp_ncu.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:factor_kernel -s 2 -c 1 -f -o gpurun_out/r02_gicp python scripts/bench_configs.py --configs cfg3 > gpurun_out/r02_gicp_ncu.log 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
tail -c 600 gpurun_out/r02_bench_1gpu.json
ls -la gpurun_out/*.ncu-rep
