#!/bin/bash
# Parity of the default inter kernel, then a kernel-only bench line per variant ("name:ENV=val,ENV=val" words in $VARIANTS),
# then (optionally) one full ncu capture of the variant named $NCU_VARIANT.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -1
if [ "${PYTEST:-1}" = "1" ]; then
  echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} 2>&1 | tail -${TAIL:-8}
fi
for v in ${VARIANTS:-v3:}; do
  name=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
  env $envs timeout 400 python bench.py --no-e2e --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$name.json'))
    r = d['roofline']
    print('%-10s value %8d  ms/step %.4f  inter %.4f ms  frac %.4f  ' % ('$name', d['value'], d['ms_per_step'], r['launch_ms'], r['frac']), {k: round(v, 4) for k, v in r['step_ms_by_kernel'].items()})
except Exception as e:
    print('$name', 'bench failed', e); print(open('gpurun_out/bench_$name.err').read()[-1500:])
PY
done
if [ -n "${NCU_VARIANT:-}" ]; then
  for v in ${VARIANTS}; do
    name=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
    if [ "$name" = "$NCU_VARIANT" ]; then
      env $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_REGEX:-k_inter}" -s ${NCU_SKIP:-2} -c 1 -f -o gpurun_out/prof_$name python bench.py --profile --steps 2 --warmup 2 ${BENCH_ARGS:-} > gpurun_out/ncu_full_$name.log 2>&1
      tail -2 gpurun_out/ncu_full_$name.log
    fi
  done
fi
