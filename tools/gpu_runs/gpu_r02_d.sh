#!/bin/bash
# round 2, call D: whole GPU suite on the restructured library (per-context launch state, ranges / multi API, variants build)
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02_d.log 2>&1
tail -15 gpurun_out/pytest_gpu_r02_d.log
