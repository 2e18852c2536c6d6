#!/bin/bash
# GPU box, round 2 call A: the whole -m gpu suite (new K4b / prefilter / BASELINE-shape tests included).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
