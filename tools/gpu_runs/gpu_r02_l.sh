#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02_l.log 2>&1; tail -4 gpurun_out/pytest_gpu_r02_l.log
SAN_TIMEOUT=240 tools/sanitize.sh
SWB_STAMPS=1 python tools/bench_wgs.py --bgzf --reads-per-file 16000000 --devices 1 --dir /tmp/synwgs > gpurun_out/wgs_stamps_n1.json 2> gpurun_out/wgs_stamps_n1.err; cat gpurun_out/wgs_stamps_n1.json | cut -c1-900; cat gpurun_out/wgs_stamps_n1.err | tail -20
