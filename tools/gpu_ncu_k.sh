#!/bin/bash
# One full ncu capture (source counters included) of the inter kernel per variant named in NCU_KERNELS.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in ${NCU_KERNELS:-warp run4 run2}; do
  echo "=== ncu full $k"
  MOBI_INTER_KERNEL=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_inter' -s 2 -c 1 -f -o gpurun_out/prof_$k python bench.py --profile --steps 2 --warmup 2 > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log
done
ls -la gpurun_out
