#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/pytest_dropin.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_dropin.log | cut -c1-300
