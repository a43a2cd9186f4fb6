#!/bin/bash
# One GPU session: parity tests, smoke, bench line, ncu launch list + full capture of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== bench reference"; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 4 --warmup 2 --streams ${NCU_STREAMS:-1024} > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
echo "=== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_inter|k_intra' -s 4 -c 4 -f -o gpurun_out/prof python bench.py --profile --steps 4 --warmup 2 --streams ${NCU_STREAMS:-1024} > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
fi
