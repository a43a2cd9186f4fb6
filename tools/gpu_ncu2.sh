#!/bin/bash
# Round-2 capture: launch list (cheap pass) + `--set full` capture of every kernel of two steady-state steps, per workload.
#   tools/gpu_ncu2.sh [workload ...]      outputs under gpurun_out/ncu_<workload>.*
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for wl in "${@:-moflex_400x240}"; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_${wl}_launches.csv \
      python bench.py --profile --steps 4 --warmup 2 --workload $wl > gpurun_out/ncu_${wl}_launch.log 2>&1
  tail -1 gpurun_out/ncu_${wl}_launch.log
  # bench.py --profile: warm-up replay (2 steps), then 4 steps; skip the staging-time and warm-up launches, capture the kernels of steps 3-4
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_inter|k_intra|k_bgra|k_mc|k_res" -s ${NCU_SKIP:-10} -c ${NCU_COUNT:-8} -f \
      -o gpurun_out/ncu_${wl} python bench.py --profile --steps 4 --warmup 2 --workload $wl > gpurun_out/ncu_${wl}_full.log 2>&1
  tail -1 gpurun_out/ncu_${wl}_full.log
done
ls -la gpurun_out | grep ncu_
