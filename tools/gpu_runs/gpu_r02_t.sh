#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ranges_multi.py -x -q -m gpu 2>&1 | tail -4
python tools/probe_rows128.py > gpurun_out/rows128_r02.json 2> gpurun_out/rows128.err; cat gpurun_out/rows128_r02.json
