#!/bin/bash
# compute-sanitizer over the hot path: memcheck on a few GPU tests, racecheck on smoke() (I- and P-pictures, all kernels).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_geometries or submit_packed or error_behaviour" 2>&1 | tail -8
echo "=== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "=== synccheck"; timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
