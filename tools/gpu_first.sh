#!/bin/bash
# first GPU contact: parity tests, then whatever else is passed
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -40
