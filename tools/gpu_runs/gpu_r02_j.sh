#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_r02_j.log 2>&1; tail -6 gpurun_out/pytest_gpu_r02_j.log
python tools/probe_mid.py > gpurun_out/mid_path_r02.json 2> gpurun_out/mid_path.err; cat gpurun_out/mid_path_r02.json
SWB_DEBUG=1 python tools/bench_wgs.py --bgzf --reads-per-file 2000000 --devices 1 --dir /tmp/synwgs > gpurun_out/wgs_probe.json 2> gpurun_out/wgs_probe.err; cat gpurun_out/wgs_probe.json | cut -c1-700; grep -E "^\[main\]|^\[wgs\]" gpurun_out/wgs_probe.err
