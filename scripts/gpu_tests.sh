#!/bin/bash
# GPU tests only (no -x: list every failure).
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -60 gpurun_out/pytest_gpu.log
