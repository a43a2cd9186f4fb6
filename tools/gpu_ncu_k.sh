#!/bin/bash
# One full ncu capture (source counters included) of the kernel(s) matching $NCU_REGEX (default k_inter), $NCU_SKIP launches skipped.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_REGEX:-k_inter}" -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-1} -f -o gpurun_out/prof_${NCU_NAME:-k} python bench.py --profile --steps 2 --warmup 2 > gpurun_out/ncu_full_${NCU_NAME:-k}.log 2>&1
tail -2 gpurun_out/ncu_full_${NCU_NAME:-k}.log
