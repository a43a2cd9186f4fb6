#!/bin/bash
# ncu launch list + full capture of the top kernels for the bench workload (value leg only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 4 --warmup 2 --streams ${NCU_STREAMS:-1024} > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_inter|k_intra}" -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-6} -f -o gpurun_out/prof python bench.py --profile --steps 4 --warmup 2 --streams ${NCU_STREAMS:-1024} > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
