#!/bin/bash
# Parity + kernel-only bench of the inter kernel (MOBI_INTER_KERNEL: anything but "warp" = k_inter_chunk), then one full ncu capture with source counters.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L; nproc
for k in ${KERNELS:-v3}; do
  echo "=== pytest -m gpu MOBI_INTER_KERNEL=$k"; MOBI_INTER_KERNEL=$k timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -${TAIL:-6}
done
for k in ${BENCH_KERNELS:-v3 chunk}; do
  echo "=== bench $k"; MOBI_INTER_KERNEL=$k timeout 400 python bench.py --no-e2e --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_$k.json 2> gpurun_out/bench_$k.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$k.json'))
    print('$k', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'k_inter ms', round(d['roofline']['launch_ms'], 4), 'frac', round(d['roofline']['frac'], 4), d['roofline']['step_ms_by_kernel'])
except Exception as e:
    print('$k', 'bench failed', e); print(open('gpurun_out/bench_$k.err').read()[-1500:])
PY
done
for k in ${NCU_KERNELS:-v3}; do
  echo "=== ncu full $k"
  MOBI_INTER_KERNEL=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_inter' -s 2 -c 1 -f -o gpurun_out/prof_$k python bench.py --profile --steps 2 --warmup 2 > gpurun_out/ncu_full_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_$k.log
done
ls -la gpurun_out | head -30
