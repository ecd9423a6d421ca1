#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/nproc_x.txt
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02_x.log 2>&1; tail -3 gpurun_out/pytest_gpu_r02_x.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/pytest_gpu_r02_x.log
